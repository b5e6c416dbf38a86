// Shading kernels of the DDGI update: sky for missed rays, material + sampleProbes + direct term for front-face hits
// (reference src/shaders/miss.rmiss, sky.glsl, closesthit.glsl NO_REFLECTION, irradiance.glsl::sampleProbes,
// pbrMetallicRoughness.glsl). Separate translation unit so that its floating-point flags can differ from the traversal and
// blend kernels; everything that feeds a later traversal (probe origin, hit position) uses explicit _rn intrinsics.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "shade.cuh"
#include "ddgi_common.cuh"
#include "hitshade.cuh"

namespace {

// miss.rmiss + sky.glsl over the dense miss queue
__global__ void __launch_bounds__(128) k_shade_miss(ShadeParams sp, const float4* __restrict__ origins, const float4* __restrict__ dirs,
                                                    const uint32_t* __restrict__ queue, const uint32_t* __restrict__ counters, float4* __restrict__ rays) {
    const uint32_t n = counters[3];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t ri = queue[i];
        const uint32_t slot = ri / sp.raysPerProbe, ray = ri - slot * sp.raysPerProbe;
        const float4 o4 = __ldg(origins + slot);
        const v3 origin = mk3(o4.x, o4.y, o4.z);
        const float4 d4 = __ldg(dirs + ray);
        const v3 c = skyColor(origin, mk3(d4.x, d4.y, d4.z), mk3(sp.light.direction[0], sp.light.direction[1], sp.light.direction[2]),
                              mk3(sp.light.color[0], sp.light.color[1], sp.light.color[2]), sp.light.color[3]);
        rays[ri] = make_float4(c.x, c.y, c.z, -1.0f);
    }
}

// closesthit.glsl:143-288 (NO_REFLECTION) over the dense front-hit queue; appends the shadow rays. TEXTURED: the scene has a texture
// list (the host picks the variant, so untextured scenes run the kernel without any texture code).
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 7 // resident 128-thread CTAs per SM k_shade_front is compiled for (72 registers; measured on B200, cfg2, shade phase: 6 CTAs 0.834 ms, 7: 0.804, 8: 0.808)
#endif
template <bool TEXTURED>
__global__ void __launch_bounds__(128, SHADE_MIN_BLOCKS) k_shade_front(DeviceScene sc, DeviceProbes pr, ShadeParams sp, const GridConsts gc, const float4* __restrict__ origins,
                                                     const float4* __restrict__ dirs, const vkx_hit* __restrict__ hits, const uint32_t* __restrict__ frontQueue,
                                                     uint32_t* __restrict__ counters, float4* __restrict__ rays, float4* __restrict__ queue) {
    const uint32_t n = counters[4];
    const v3 lightDir = mk3(sp.light.direction[0], sp.light.direction[1], sp.light.direction[2]);
    const v3 lightColor = mk3(sp.light.color[0], sp.light.color[1], sp.light.color[2]);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t ri = frontQueue[i];
        const uint32_t slot = ri / sp.raysPerProbe, ray = ri - slot * sp.raysPerProbe;
        const float4 o4 = __ldg(origins + slot);
        const v3 origin = mk3(o4.x, o4.y, o4.z);
        const float4 d4 = __ldg(dirs + ray);
        const v3 direction = mk3(d4.x, d4.y, d4.z);
        const vkx_hit h = hits[ri];
        const v3 position = pointOnRayExact(origin, direction, h.t);
        v3 color, lit;
        if (TEXTURED) { // payload.raydx / raydy, traceProbes.rgen:40-41
            const v3 raydx = rotateAxisH(direction, norm3(cross3(direction, mk3(1.0f, 0.0f, 0.0f))), 0.001f);
            const v3 raydy = rotateAxisH(direction, norm3(cross3(direction, mk3(0.0f, 1.0f, 0.0f))), 0.001f);
            shadeFrontHit<true>(sc, pr, gc, lightDir, lightColor, direction, position, h, color, lit, origin, raydx, raydy);
        } else shadeFrontHit<false>(sc, pr, gc, lightDir, lightColor, direction, position, h, color, lit);
        rays[ri] = make_float4(color.x, color.y, color.z, h.t); // value if the sun is occluded; k_trace_shadow writes `lit` if the shadow ray escapes
        // Every front hit spawns exactly one shadow ray, so its slot is the item's own index: no atomic append, and the shadow queue
        // keeps the cell-sorted order of the front queue exactly (coherent origins for k_trace_shadow).
        const uint32_t qi = i;
        if (i == 0u) counters[0] = n;
        queue[2 * size_t(qi)] = make_float4(position.x, position.y, position.z, __uint_as_float(ri));
        queue[2 * size_t(qi) + 1] = make_float4(lit.x, lit.y, lit.z, h.t); // the complete ray record if the shadow ray escapes
    }
}

} // namespace

void launchShadeMiss(unsigned blocks, cudaStream_t st, const ShadeParams& sp, const float4* origins, const float4* dirs, const uint32_t* missQueue,
                     const uint32_t* counters, float4* rays) {
    k_shade_miss<<<blocks, 128, 0, st>>>(sp, origins, dirs, missQueue, counters, rays);
}
void launchShadeFront(unsigned blocks, cudaStream_t st, const DeviceScene& sc, const DeviceProbes& pr, const ShadeParams& sp, const float4* origins,
                      const float4* dirs, const vkx_hit* hits, const uint32_t* frontQueue, uint32_t* counters, float4* rays, float4* shadowQueue) {
    // One queue item per thread (the loop only matters if the grid is capped). Items are not equal (a hit near the volume's border or
    // next to disabled probes skips probes) and the sky kernel shares the SMs, so the hardware block scheduler balances short blocks
    // better than a grid-stride loop over few long ones; the shadow queue is also appended closer to the sorted order, which makes the
    // shadow rays more coherent. Measured on B200, cfg2 (profiles/r01d_shade_grid_sweep.txt), sort + shading / k_trace_shadow:
    // 6 blocks per SM (one wave) 0.907 ms, 16 (the old grid) 0.854 / 0.375 ms, 36: 0.841, 64: 0.832 / 0.360, 128: 0.825 / 0.357,
    // uncapped: 0.823 / 0.351 ms. VKX_SHADE_BLOCKS_PER_SM caps the grid (tuning).
    static const int capPerSm = [] { const char* e = getenv("VKX_SHADE_BLOCKS_PER_SM"); return e && atoi(e) > 0 ? atoi(e) : 0; }(); // an environment setting, not device state
    blocks = unsigned((sp.numRays + 127u) / 128u);
    if (capPerSm) { // the SM count belongs to the current device (one context per device; several contexts may live in one process)
        int dev = 0, smCount = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, dev);
        const unsigned long long cap = (unsigned long long)(smCount) * (unsigned long long)(capPerSm);
        if ((unsigned long long)(blocks) > cap) blocks = unsigned(cap);
    }
    const GridConsts gc = makeGridConstsHost(sp.grid);
    if (sc.numTextures) k_shade_front<true><<<blocks, 128, 0, st>>>(sc, pr, sp, gc, origins, dirs, hits, frontQueue, counters, rays, shadowQueue);
    else k_shade_front<false><<<blocks, 128, 0, st>>>(sc, pr, sp, gc, origins, dirs, hits, frontQueue, counters, rays, shadowQueue);
}
