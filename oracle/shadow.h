// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
#pragma once
#include <cstdint>
#include <vector>
#include "ddgi.h"

namespace oshadow {

struct State {
    uint32_t w = 0, h = 0;
    std::vector<float> positionDepth, normalMetalness; // RGBA32F G-buffer
    std::vector<float> albedoRoughness, emissive;      // RGBA32F G-buffer targets only the composite reads (GBuffer.frag:66-67)
    std::vector<float> gathered;                       // RGBA32F output of finalGather
    std::vector<float> reflRaw, reflX, reflFinal, reflPrevious; // reflection pass images (RGBA32F)
    std::vector<float> reflDirs;                       // jittered reflection direction per pixel (debug / parity)
    std::vector<vkx_hit> reflHits;                     // closest hit of the reflection ray (t = -1: miss or no ray)
    std::vector<uint8_t> reflMask;                     // 0 no ray, 1 miss, 2 back face, 3 front lit, 4 front shadowed
    std::vector<float> raw, filteredX, final_, previous; // RGBA32F
    std::vector<float> dirs;   // jittered light direction per pixel (debug / parity)
    std::vector<uint8_t> mask; // 0 not traced, 1 lit, 2 shadowed
    std::vector<float> noise;  // [slices][h][w][4]
    uint32_t noiseW = 0, noiseH = 0, noiseSlices = 0;
};

void init(State& st, uint32_t w, uint32_t h);
void gbufferGenerate(const oddgi::Scene& s, State& st, const vkx_camera& cam);
// dirOverride (optional, [h][w][3]): use these jittered directions instead of computing them (bit-exact mask tests).
void frame(const oddgi::Scene& s, State& st, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light, const float* dirOverride);

// reflection.rgen:117-186 + reflectionFilterX/Y (reflectionFilter.glsl) with the history copy of the editor (previous <- final).
// dirOverride (optional, [h][w][3]): use these jittered directions instead of computing them (bit-exact hit tests).
void reflectionFrame(const oddgi::Scene& s, const oddgi::Probes& probes, State& st, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light, const float* dirOverride);

// FinalGather.frag:38-77: sky on empty pixels, else direct * shadow + specular * reflection + sampleProbes * diffuse + emissive.
// reflection: optional RGBA32F [h][w][4] (e.g. reflFinal of reflectionFrame; nullptr = black).
void finalGather(const oddgi::Scene& s, const oddgi::Probes& probes, State& st, const vkx_camera& cam, const vkx_light& light, const float* reflection);

float gaussian(float stdDev, float dist);  // directLightFilter.glsl:29-31
float rgaussian(float stdDev, float dist); // reflectionFilter.glsl:37-39

} // namespace oshadow
