#include "IrradianceProbes.hpp"
#include <cmath>

namespace vkx {

void IrradianceProbes::init(const Device& device, vec3 min, vec3 max) { // reference src/IrradianceProbes.cpp:12-104
    _device = &device;
    for (int a = 0; a < 3; ++a) { GridParameters.extentMin[a] = min[a]; GridParameters.extentMax[a] = max[a]; }
    check(device.ctx(), vkx_probes_init(device.ctx(), &GridParameters));
    updateUniforms();
    _lastUpdateOffset = _loopIndex = _updatedProbes = 0;
}

void IrradianceProbes::initProbes() { // reference src/IrradianceProbes.cpp:357-394
    GridParameters.hysteresis = 0.0f;
    float orientation[16];
    vkx_host_next_orientation(&_rngState, orientation);
    check(_device->ctx(), vkx_probes_classify(_device->ctx(), orientation));
}

uint32_t IrradianceProbes::selectProbesToUpdate(std::vector<uint32_t>& toUpdate) {
    // Get probe states back (the reference maps the host-visible state buffer every update, :399-403)
    _probesState.resize(getProbeCount());
    check(_device->ctx(), vkx_probes_download(_device->ctx(), nullptr, nullptr, _probesState.data(), nullptr, 0));
    toUpdate.resize(_probesState.size());
    uint32_t n = vkx_host_select_probes(&_loopIndex, &_lastUpdateOffset, _probesState.data(), getProbeCount(), ProbesPerUpdate, toUpdate.data());
    toUpdate.resize(n);
    return n;
}

void IrradianceProbes::update() { // reference src/IrradianceProbes.cpp:426-594
    vkx_ctx* ctx = _device->ctx();
    check(ctx, vkx_sync(ctx)); // vkWaitForFences
    std::vector<uint32_t> toUpdate;
    uint32_t probeCount = 0;
    if (DeviceScheduler) {
        if (!_deviceSchedulerSeeded) { // hand the host counters over once, so the switch can be flipped mid-run
            const uint32_t st[2] = {_loopIndex, _lastUpdateOffset};
            check(ctx, vkx_probes_scheduler_state(ctx, st, nullptr));
            _deviceSchedulerSeeded = true;
        }
        check(ctx, vkx_probes_schedule(ctx, ProbesPerUpdate, &probeCount));
        uint32_t st[2];
        check(ctx, vkx_probes_scheduler_state(ctx, nullptr, st));
        _loopIndex = st[0]; _lastUpdateOffset = st[1];
    } else {
        _deviceSchedulerSeeded = false;
        probeCount = selectProbesToUpdate(toUpdate);
    }
    if (_haveTimings) { // the five timestamp differences of the previous update (:441-452)
        float ms[5];
        check(ctx, vkx_probes_timings(ctx, ms));
        _computeTimes.add(ms[0]); _traceTimes.add(ms[1]); _updateTimes.add(ms[2]); _borderCopyTimes.add(ms[3]); _copyTimes.add(ms[4]);
    }
    float orientation[16];
    vkx_host_next_orientation(&_rngState, orientation);
    // Get closer to the target hysteresis (:462-476). Quirk A.5.8: the first branch uploads the UBO before incrementing.
    if (_updatedProbes >= getProbeCount()) {
        if (std::abs(GridParameters.hysteresis - TargetHysteresis) > 0.05) {
            updateUniforms();
            GridParameters.hysteresis += 0.1f * (TargetHysteresis - GridParameters.hysteresis);
        } else if (GridParameters.hysteresis != TargetHysteresis) {
            GridParameters.hysteresis = TargetHysteresis;
            updateUniforms();
        }
        _updatedProbes -= getProbeCount();
    }
    _updatedProbes += probeCount;
    _lastCount = probeCount;
    if (probeCount == 0) return;
    vkx_light light;
    for (int i = 0; i < 4; ++i) { light.direction[i] = _lightBuffer->direction[i]; light.color[i] = _lightBuffer->color[i]; }
    // raysPerProbe edits take effect immediately in the reference (the UBO field is shared with GridParameters only through
    // updateUniforms, but raysPerProbe also sizes the dispatch); keep the device copy's hysteresis lag, refresh the rest.
    GridInfo g = _deviceGrid;
    g.raysPerProbe = GridParameters.raysPerProbe;
    if (DeviceScheduler) check(ctx, vkx_probes_update_scheduled(ctx, &g, &light, orientation, 0));
    else check(ctx, vkx_probes_update(ctx, &g, &light, orientation, toUpdate.data(), probeCount, 0));
    _haveTimings = true;
}

void IrradianceProbes::download(std::vector<uint32_t>& irradiance, std::vector<uint32_t>& depth, std::vector<uint32_t>& state) const {
    const size_t plane = size_t(GridParameters.resolution[0]) * size_t(GridParameters.resolution[1]), rz = size_t(GridParameters.resolution[2]);
    irradiance.resize(64 * plane * rz); depth.resize(256 * plane * rz); state.resize(plane * rz);
    check(_device->ctx(), vkx_probes_download(_device->ctx(), irradiance.data(), depth.data(), state.data(), nullptr, 0));
}

} // namespace vkx
