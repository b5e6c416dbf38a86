// Declarations shared by ddgi.cu (traversal, blend; exact arithmetic) and ddgi_shade.cu (shading kernels).
#pragma once
#include "common.cuh"

// Thread -> ray mapping shared by the trace and shade kernels. A warp handles a tile of 8 probes x 4 directions: the 8 probes are
// a 2x2x2 block of the grid (`order` lists slots block by block) and the 4 directions are neighbours on the sphere (`perm`
// sorts the frame's direction table along a Morton curve of the octahedral map). Rays of a warp are then near-parallel with
// nearby origins, instead of 32 consecutive spherical-Fibonacci directions of one probe. Results are stored by (slot, ray), so
// the mapping changes scheduling only.
struct RayMap {
    uint32_t count, raysPerProbe, numDirGroups, numThreads;
    uint32_t dgShift;      // log2(numDirGroups) if it is a power of two (256 rays -> 64 groups), else 0xFFFFFFFF
    const uint32_t* order; // [count] position -> slot
    const uint32_t* perm;  // [raysPerProbe] position -> ray index
};
__device__ __forceinline__ bool mapRay(const RayMap& m, uint32_t t, uint32_t& slot, uint32_t& ray) {
    const uint32_t tile = t >> 5, lane = t & 31u;
    const uint32_t pg = m.dgShift != 0xFFFFFFFFu ? tile >> m.dgShift : tile / m.numDirGroups, dg = tile - pg * m.numDirGroups;
    const uint32_t j = pg * 8u + (lane & 7u), k = dg * 4u + (lane >> 3);
    if (j >= m.count || k >= m.raysPerProbe) return false;
    slot = __ldg(m.order + j); ray = __ldg(m.perm + k);
    return true;
}

__device__ __forceinline__ uint32_t warpAppend(uint32_t* counter) { // warp-aggregated queue slot allocation
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(int(m)) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(m, base, leader);
    return base + uint32_t(__popc(m & ((1u << lane) - 1u)));
}

struct ShadeParams {
    vkx_grid_info grid;
    vkx_light light;
    uint32_t raysPerProbe, numRays;
};

void launchShadeMiss(unsigned blocks, cudaStream_t st, const ShadeParams& sp, const float4* origins, const float4* dirs, const uint32_t* missQueue,
                     const uint32_t* counters, float4* rays);
void launchShadeFront(unsigned blocks, cudaStream_t st, const DeviceScene& sc, const DeviceProbes& pr, const ShadeParams& sp, const float4* origins,
                      const float4* dirs, const vkx_hit* hits, const uint32_t* frontQueue, uint32_t* counters, float4* rays, float4* shadowQueue);
