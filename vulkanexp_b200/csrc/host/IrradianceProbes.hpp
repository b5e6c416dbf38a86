// The reference's IrradianceProbes class (reference src/IrradianceProbes.hpp:16-129) above the C ABI: same public
// members and method names, same host logic (orientation RNG, hysteresis ramp, probe scheduler), CUDA underneath.
#pragma once
#include <cstdint>
#include <vector>
#include "Renderer.hpp"

namespace vkx {

struct LightBuffer { // reference src/Light.hpp:6-9
    float direction[4] = {0.09901475f, 0.99014754f, 0.09901475f, 1.0f}; // normalize(0.2, 2, 0.2), 1
    float color[4] = {10.0f, 10.0f, 10.0f, 10.0f};
};

class RollingBuffer { // reference src/RollingBuffer.hpp, reduced
  public:
    void add(float v) { _v.push_back(v); if (_v.size() > 256) _v.erase(_v.begin()); }
    const std::vector<float>& get() const { return _v; }
    float last() const { return _v.empty() ? 0.f : _v.back(); }
  private:
    std::vector<float> _v;
};

class IrradianceProbes {
  public:
    void init(const Device& device, vec3 min, vec3 max);                  // reference :18 (queue family arguments dropped)
    void initProbes();                                                    // reference :19
    void createPipeline() {}                                              // kernels are compiled into the library
    void writeDescriptorSet(const Renderer& renderer, const LightBuffer& lightBuffer) { _renderer = &renderer; _lightBuffer = &lightBuffer; }
    void setLightBuffer(const LightBuffer& lightBuffer) { _lightBuffer = &lightBuffer; }
    void updateUniforms() { _deviceGrid = GridParameters; }               // reference :24: what the device sees
    void update();                                                        // reference :25
    void destroy() { _device = nullptr; }

    uint32_t getProbeCount() const { return uint32_t(GridParameters.resolution[0] * GridParameters.resolution[1] * GridParameters.resolution[2]); }
    // getIrradiance / getDepth / getProbeInfoBuffer: host copies of the sampled atlases and the state buffer
    void download(std::vector<uint32_t>& irradiance, std::vector<uint32_t>& depth, std::vector<uint32_t>& state) const;

    static const uint32_t MaxRaysPerProbe = VKX_MAX_RAYS_PER_PROBE;
    uint32_t ProbesPerUpdate = 0;
    // Not in the reference: run selectProbesToUpdate on the device (vkx_probes_schedule) instead of reading the states back.
    // Same lists, same counters; only the list length returns to the host.
    bool DeviceScheduler = false;
    float TargetHysteresis = 0.98f;
    using GridInfo = vkx_grid_info;
    GridInfo GridParameters{{0, 0, 0}, 12.0f, {0, 0, 0}, 0.0f, {32, 16, 32}, 192, 8, 16, 0.3f, 0};

    const RollingBuffer& getComputeTimes() const { return _computeTimes; }
    const RollingBuffer& getTraceTimes() const { return _traceTimes; }
    const RollingBuffer& getUpdateTimes() const { return _updateTimes; }
    const RollingBuffer& getBorderCopyTimes() const { return _borderCopyTimes; }
    const RollingBuffer& getCopyTimes() const { return _copyTimes; }
    uint32_t lastUpdatedProbeCount() const { return _lastCount; }

  private:
    uint32_t selectProbesToUpdate(std::vector<uint32_t>& toUpdate); // reference :396-424
    const Device* _device = nullptr;
    const Renderer* _renderer = nullptr;
    const LightBuffer* _lightBuffer = nullptr;
    GridInfo _deviceGrid{};
    std::vector<uint32_t> _probesState;
    uint32_t _lastUpdateOffset = 0, _loopIndex = 0, _updatedProbes = 0, _rngState = 1, _lastCount = 0;
    bool _deviceSchedulerSeeded = false;
    bool _haveTimings = false;
    RollingBuffer _computeTimes, _traceTimes, _updateTimes, _borderCopyTimes, _copyTimes;
};

} // namespace vkx
