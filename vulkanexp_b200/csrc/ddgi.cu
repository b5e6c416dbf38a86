// DDGI probe update on the device. Replaces the command buffer IrradianceProbes::update records (reference
// src/IrradianceProbes.cpp:486-576): traceProbes.rgen -> {closesthit_noreflection.rchit | miss.rmiss, shadow.rmiss}
// -> probesUpdateIrradiance.comp + probesUpdateDepth.comp -> probesCopyBorders.comp -> 2x vkCmdCopyImage.
//
// Kernel pipeline per chunk of probes (wavefront, rays grouped by what they need next):
//   k_trace_primary  closest-hit traversal of probe rays                              -> hit records
//   k_shade          miss: sky | back face: 0.8 t | front: material + 2x sampleProbes -> ray records, shadow queue
//   k_trace_shadow   any-hit traversal of the compacted shadow-ray queue, adds the direct term
//   k_blend          one CTA per probe: ray records staged in shared memory, 196 depth + 36 irradiance texels,
//                    hysteresis mix against the work atlas, state machine, border texels, packed tile stores
//   k_publish        work -> sampled for the updated probes (the reference copies both whole atlases)
#include "common.cuh"
#include "traverse.cuh"
#include "ptrace.cuh"
#include "shade.cuh"
#include "ddgi_common.cuh"
#include "blend_tc.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

struct TraceParams {
    vkx_grid_info grid;
    float tmin, tmax;
    float invCell[3];
    uint32_t raysPerProbe, numRays; // numRays = chunk probes * raysPerProbe
};

// Loop-invariant halves of the ray set-up, tabulated once instead of once per ray: the reciprocal direction / octant of each of the
// frame's directions (what makeRay() derives from a direction) and the world position of each probe of the chunk's list.
__global__ void k_dir_table(uint32_t N, const float4* __restrict__ dirs, float4* __restrict__ invDirs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 d = dirs[i];
    const Ray r = makeRay(0.f, 0.f, 0.f, d.x, d.y, d.z);
    invDirs[i] = make_float4(r.ix, r.iy, r.iz, __uint_as_float(r.oct));
}
__global__ void k_origin_table(vkx_grid_info grid, const uint32_t* __restrict__ probeIndices, uint32_t n, float4* __restrict__ origins) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int ix, iy, iz; probeGridIndex(__ldg(probeIndices + s), grid, ix, iy, iz);
    const v3 o = probeWorldPos(ix, iy, iz, grid);
    origins[s] = make_float4(o.x, o.y, o.z, 0.0f);
}

// What a finished primary ray needs next, as one word per ray: a miss goes to the sky queue, a back-face hit is final, a front-face
// hit is shaded and carries its grouping key = (grid cell of the hit point) >> binShift (scheduling only: rays that shade from the
// same 8 probes end up in the same warps of k_shade_front, and their shadow rays start close together).
#define KEY_MISS 0xFFFFFFFFu
#define KEY_BACK 0xFFFFFFFEu
__device__ __forceinline__ uint32_t hitKey(const TraceParams& tp, uint32_t binShift, float ox, float oy, float oz, float dx, float dy, float dz, float t, uint32_t prim) {
    if (prim == 0xFFFFFFFFu) return KEY_MISS;
    if (prim & 0x80000000u) return KEY_BACK;
    const int cx = min(max(int((ox + dx * t - tp.grid.extentMin[0]) * tp.invCell[0]), 0), tp.grid.resolution[0] - 1);
    const int cy = min(max(int((oy + dy * t - tp.grid.extentMin[1]) * tp.invCell[1]), 0), tp.grid.resolution[1] - 1);
    const int cz = min(max(int((oz + dz * t - tp.grid.extentMin[2]) * tp.invCell[2]), 0), tp.grid.resolution[2] - 1);
    return uint32_t(cx + tp.grid.resolution[0] * (cy + tp.grid.resolution[1] * cz)) >> binShift;
}

// Rays are sorted by what they need next when their traversal ends: misses -> sky queue, front-face hits -> shading queue,
// back-face hits are final (closesthit.glsl:137-141) and written here. The two shading kernels then run on dense queues.
struct PrimarySrc {
    TraceParams tp; RayMap rm; const float4* origins; const float4* dirs; const float4* invDirs; vkx_hit* hits;
    float4* rays;
    uint32_t ri;
    __device__ __forceinline__ bool load(uint32_t item, Ray& r, float& tmin, float& tmax, uint32_t& cullMask) {
        uint32_t slot, ray;
        if (!mapRay(rm, item, slot, ray)) return false;
        ri = slot * tp.raysPerProbe + ray;
        const float4 o = __ldg(origins + slot), d = __ldg(dirs + ray), id = __ldg(invDirs + ray);
        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.ix = id.x; r.iy = id.y; r.iz = id.z; r.oct = __float_as_uint(id.w);
        tmin = tp.tmin; tmax = tp.tmax; cullMask = VKX_INSTANCE_STATIC | VKX_INSTANCE_DYNAMIC;
        return true;
    }
    // The hit record only (plus the final value of a back-face hit, closesthit.glsl:137-141). Which queue the ray goes to next is
    // decided by the classification pass: appending here cost a global atomic round trip per finishing ray with the whole warp
    // waiting on it (16 % of the kernel's stall samples, profiles/r02f_src_k_trace_primary.txt), and even computing the ray's
    // grouping key here (the ray is still in registers) loses: the finishing path runs at ~2 of 32 lanes, so its ~25 instructions
    // cost 12 warp-instructions per ray (k_trace_primary 0.759 -> 0.794 ms) against one in the dense pass.
    __device__ __forceinline__ void store(uint32_t, const HitRec& h, bool) {
        vkx_hit out; out.t = h.t; out.instance = h.inst; out.primitive = h.prim; out.u = h.u; out.v = h.v;
        hits[ri] = out;
        if (h.found && (h.prim & 0x80000000u)) rays[ri] = make_float4(0.f, 0.f, 0.f, h.t * 0.80f);
    }
};

// Rays are sorted by what they need next: misses -> sky queue, front-face hits -> shading queue with the sort key of the hit
// (= grid cell of the hit point; scheduling only: rays that shade from the same 8 probes end up in the same warps of
// k_shade_front, and their shadow rays start close together). Back-face hits are final. A dense pass over the hit records
// (84 MB at cfg2), warp-aggregated appends.
__global__ void __launch_bounds__(256) k_classify_hits(TraceParams tp, const float4* __restrict__ origins, const float4* __restrict__ dirs, const vkx_hit* __restrict__ hits,
                                                       uint32_t* __restrict__ missQueue, uint32_t* __restrict__ frontQueue, uint32_t* __restrict__ keys, uint32_t* __restrict__ counters) {
    // Appends are aggregated per block: one atomic per queue per 256 rays (per-warp atomics on the two counters serialised in L2:
    // the pass took 0.14 ms instead of 0.02).
    __shared__ uint32_t sCount[2][8], sBase[2];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = blockIdx.x * 256u; base < tp.numRays; base += gridDim.x * 256u) { // block-uniform trip count
        const uint32_t ri = base + threadIdx.x;
        bool miss = false, front = false; float t = 0.0f;
        if (ri < tp.numRays) {
            t = hits[ri].t; const uint32_t prim = hits[ri].primitive;
            miss = prim == 0xFFFFFFFFu; front = !miss && !(prim & 0x80000000u);
        }
        const uint32_t mm = __ballot_sync(0xFFFFFFFFu, miss), fm = __ballot_sync(0xFFFFFFFFu, front);
        if (lane == 0) { sCount[0][warp] = uint32_t(__popc(mm)); sCount[1][warp] = uint32_t(__popc(fm)); }
        __syncthreads();
        if (threadIdx.x < 2) {
            uint32_t total = 0;
            for (int w = 0; w < 8; ++w) { const uint32_t c = sCount[threadIdx.x][w]; sCount[threadIdx.x][w] = total; total += c; }
            sBase[threadIdx.x] = total ? atomicAdd(counters + 3 + threadIdx.x, total) : 0u;
        }
        __syncthreads();
        const uint32_t lt = (1u << lane) - 1u;
        if (miss) missQueue[sBase[0] + sCount[0][warp] + uint32_t(__popc(mm & lt))] = ri;
        if (front) {
            const uint32_t slot = ri / tp.raysPerProbe, ray = ri - slot * tp.raysPerProbe;
            const float4 o = __ldg(origins + slot);
            const float4 d = __ldg(dirs + ray);
            const int cx = min(max(int((o.x + d.x * t - tp.grid.extentMin[0]) * tp.invCell[0]), 0), tp.grid.resolution[0] - 1);
            const int cy = min(max(int((o.y + d.y * t - tp.grid.extentMin[1]) * tp.invCell[1]), 0), tp.grid.resolution[1] - 1);
            const int cz = min(max(int((o.z + d.z * t - tp.grid.extentMin[2]) * tp.invCell[2]), 0), tp.grid.resolution[2] - 1);
            const uint32_t i = sBase[1] + sCount[1][warp] + uint32_t(__popc(fm & lt));
            frontQueue[i] = ri;
            keys[i] = uint32_t(cx + tp.grid.resolution[0] * (cy + tp.grid.resolution[1] * cz));
        }
        __syncthreads(); // sCount / sBase are rewritten by the next iteration
    }
}

// ---- front hits grouped by grid cell without a sort library: counting sort with block-private histograms
// The key of a front hit is its bin = (grid cell of the hit point) >> binShift, at most BIN_MAX bins (64 KB of shared-memory
// counters). k_bin_count reads the hit records, writes one key word per ray (hitKey), appends the misses of each tile to the
// sky queue, counts the front hits of BIN_TILE rays at a time in shared memory and adds the block's counts to the global histogram
// once per block (hit points cluster in few cells: per-ray or per-warp global atomics on the hot cells serialise in L2, the reason
// an earlier counting sort lost to the radix sort). k_bin_scan turns the histogram into first positions; k_bin_scatter counts each
// tile again in shared memory (a ray's rank inside its tile's bin is what the shared-memory atomic returns), reserves the tile's
// range of every bin it touched with one global atomic, and writes the ray indices. The order inside a bin depends on scheduling;
// every queue item is shaded independently and results are stored by ray index, so outputs do not.
#define BIN_TILE 4096u
#define BIN_MAX 16384u
#define BIN_PER (BIN_TILE / 256u)
__global__ void __launch_bounds__(256) k_bin_count(TraceParams tp, uint32_t numBins, uint32_t binShift, const float4* __restrict__ origins, const float4* __restrict__ dirs,
                                                   const vkx_hit* __restrict__ hits, uint32_t* __restrict__ missQueue, uint32_t* __restrict__ keys, uint32_t* __restrict__ hist,
                                                   uint32_t* __restrict__ counters) {
    extern __shared__ uint32_t sBins[];
    __shared__ uint32_t sMissBase, sMissCount, sFront;
    for (uint32_t b = threadIdx.x; b < numBins; b += 256u) sBins[b] = 0u;
    if (threadIdx.x == 0) sFront = 0u;
    const uint32_t numTiles = (tp.numRays + BIN_TILE - 1u) / BIN_TILE;
    uint32_t myFront = 0;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) { // block-uniform trip count
        if (threadIdx.x == 0) sMissCount = 0u;
        __syncthreads();
        float t[BIN_PER]; uint32_t prim[BIN_PER];
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) { // the two words of all the thread's hit records in flight at once (one after the other, behind the branches below, they were a chain of DRAM round trips: 51 us for 84 MB)
            const uint32_t ri = tile * BIN_TILE + j * 256u + threadIdx.x;
            const bool in = ri < tp.numRays;
            t[j] = in ? hits[ri].t : 0.0f; prim[j] = in ? hits[ri].primitive : 0x80000000u; // padding counts as a (final) back face
        }
        uint32_t missMask = 0, missRank = 0;
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) {
            const uint32_t ri = tile * BIN_TILE + j * 256u + threadIdx.x;
            const uint32_t slot = ri / tp.raysPerProbe, ray = ri - slot * tp.raysPerProbe;
            uint32_t key = KEY_BACK;
            if (prim[j] == 0xFFFFFFFFu) { missMask |= 1u << j; key = KEY_MISS; }
            else if (!(prim[j] & 0x80000000u)) {
                const float4 o = __ldg(origins + slot), d = __ldg(dirs + ray);
                key = hitKey(tp, binShift, o.x, o.y, o.z, d.x, d.y, d.z, t[j], prim[j]);
                atomicAdd(&sBins[key], 1u);
                ++myFront;
            }
            if (ri < tp.numRays) keys[ri] = key;
        }
        // misses of the tile: one reservation per block, positions by (thread, ray) inside the tile
        const uint32_t nMiss = uint32_t(__popc(missMask));
        if (nMiss) missRank = atomicAdd(&sMissCount, nMiss);
        __syncthreads();
        if (threadIdx.x == 0) sMissBase = sMissCount ? atomicAdd(counters + 3, sMissCount) : 0u;
        __syncthreads();
        uint32_t pos = sMissBase + missRank;
        while (missMask) { const uint32_t j = uint32_t(__ffs(int(missMask))) - 1u; missMask &= missMask - 1u; missQueue[pos++] = tile * BIN_TILE + j * 256u + threadIdx.x; }
    }
    if (myFront) atomicAdd(&sFront, myFront);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < numBins; b += 256u) { const uint32_t c = sBins[b]; if (c) atomicAdd(hist + b, c); }
    if (threadIdx.x == 0 && sFront) atomicAdd(counters + 4, sFront);
}
// exclusive scan of the bin histogram (<= BIN_MAX entries): one block, each thread scans its <= 16 consecutive bins; the loads are
// unrolled so that all of them are in flight at once (a runtime-bounded loop of global loads took 18 us)
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t numBins, const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t sWarp[32];
    constexpr uint32_t PER = BIN_MAX / 1024u;
    const uint32_t first = threadIdx.x * PER;
    uint32_t v[PER], sum = 0;
    // 128-bit loads: a thread's 16 bins are 64 contiguous bytes; word loads made one SM serve 16 k sector requests (15 us)
#pragma unroll
    for (uint32_t k = 0; k < PER; k += 4u) {
        if (first + k + 3u < numBins) { const uint4 q = *reinterpret_cast<const uint4*>(hist + first + k); v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w; }
        else { for (uint32_t j = 0; j < 4u; ++j) v[k + j] = first + k + j < numBins ? hist[first + k + j] : 0u; }
    }
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) sum += v[k];
    uint32_t incl = sum; // inclusive scan over the block
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (int(threadIdx.x & 31u) >= o) incl += u; }
    if ((threadIdx.x & 31u) == 31u) sWarp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32u) {
        uint32_t w = sWarp[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, w, o); if (int(threadIdx.x) >= o) w += u; }
        sWarp[threadIdx.x] = w;
    }
    __syncthreads();
    uint32_t run = incl - sum + ((threadIdx.x >> 5) ? sWarp[(threadIdx.x >> 5) - 1u] : 0u);
#pragma unroll
    for (uint32_t k = 0; k < PER; k += 4u) {
        uint4 q; q.x = run; q.y = q.x + v[k]; q.z = q.y + v[k + 1]; q.w = q.z + v[k + 2]; run = q.w + v[k + 3];
        if (first + k + 3u < numBins) *reinterpret_cast<uint4*>(cursor + first + k) = q;
        else { const uint32_t w[4] = {q.x, q.y, q.z, q.w}; for (uint32_t j = 0; j < 4u; ++j) if (first + k + j < numBins) cursor[first + k + j] = w[j]; }
    }
}
// A tile touches few of the bins (4096 rays of 16 probes): instead of sweeping all of them, the thread whose shared-memory atomic
// returned rank 0 for a bin speaks for it - it reserves the tile's range of that bin (all of a thread's reservations in flight
// together), leaves the range's start in the bin's counter for the writers, and clears the counter for the next tile afterwards.
// (The sweep over 16 K counters per tile - 8 dependent rounds of global atomics plus the clear - was most of the 51 us.)
// Tiles of the scatter are larger than those of the count (SCAT_THREADS x BIN_PER rays): hit points cluster, so every tile reserves a
// range in the same few hot bins, and those reservations - atomics with a return value on one address - are served one after the
// other by L2. Measured by leaving phases out (cfg2, under ncu, cold caches): 23 us for the key loads and the shared-memory ranks,
// + 11 us for the scattered 4-byte stores, + 5-10 us for the reservations; 4096-ray tiles: 59 us, the per-tile sweep before: 53 us.
#define SCAT_THREADS 1024u
#define SCAT_TILE (SCAT_THREADS * BIN_PER)
__global__ void __launch_bounds__(SCAT_THREADS) k_bin_scatter(uint32_t numRays, uint32_t numBins, const uint32_t* __restrict__ keys, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
    extern __shared__ uint32_t sBins[];
    const uint32_t numTiles = (numRays + SCAT_TILE - 1u) / SCAT_TILE;
    for (uint32_t b = threadIdx.x; b < numBins; b += SCAT_THREADS) sBins[b] = 0u;
    __syncthreads();
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        uint32_t key[BIN_PER], rank[BIN_PER];
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) {
            const uint32_t ri = tile * SCAT_TILE + j * SCAT_THREADS + threadIdx.x;
            key[j] = ri < numRays ? __ldg(keys + ri) : KEY_BACK;
        }
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) rank[j] = key[j] < KEY_BACK ? atomicAdd(&sBins[key[j]], 1u) : 0xFFFFFFFFu;
        __syncthreads();
        uint32_t lead = 0, base[BIN_PER];
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) if (rank[j] == 0u) { lead |= 1u << j; base[j] = atomicAdd(cursor + key[j], sBins[key[j]]); }
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) if (lead & (1u << j)) sBins[key[j]] = base[j]; // only this thread touches the counter of a bin it leads until the barrier
        __syncthreads();
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j)
            if (key[j] < KEY_BACK) sorted[sBins[key[j]] + rank[j]] = tile * SCAT_TILE + j * SCAT_THREADS + threadIdx.x;
        __syncthreads();
#pragma unroll
        for (uint32_t j = 0; j < BIN_PER; ++j) if (lead & (1u << j)) sBins[key[j]] = 0u;
        __syncthreads();
    }
}

#ifndef PT_DEFER_PRIMARY_DEFAULT
#define PT_DEFER_PRIMARY_DEFAULT 0 // closest-hit rays lose with deferral: 0.949 ms immediate, 1.038 / 0.995 / 0.978 ms with 8 / 12 / 16
#endif
#ifndef PT_DEFER_SHADOW_DEFAULT
#define PT_DEFER_SHADOW_DEFAULT 12 // measured on B200, cfg2 (gpurun_out r01b sweep): 0.400 ms immediate, 0.378 / 0.374 / 0.382 ms with 8 / 12 / 16
#endif
// Persistent warps; see ptrace.cuh. DEFER = 0: triangles tested right after their node; DEFER > 0: parked until DEFER lanes hold some.
#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 8 // resident CTAs per SM the traversal kernels are compiled for (64 registers)
#endif
template <int DEFER>
__global__ void __launch_bounds__(128, PT_MIN_BLOCKS) k_trace_primary(DeviceScene sc, TraceParams tp, RayMap rm, const float4* __restrict__ origins,
                                                       const float4* __restrict__ dirs, const float4* __restrict__ invDirs, vkx_hit* __restrict__ hits, float4* __restrict__ rays,
                                                       uint32_t* __restrict__ counters) {
    PrimarySrc src; src.tp = tp; src.rm = rm; src.origins = origins; src.dirs = dirs; src.invDirs = invDirs; src.hits = hits; src.ri = 0;
    src.rays = rays;
    if (DEFER == 0) persistentTrace<false>(sc.nodes, sc.tris, src, rm.numThreads, counters + 1);
    else if (DEFER < 0) persistentTracePF<false>(sc.nodes, sc.tris, src, rm.numThreads, counters + 1);
    else persistentTraceDeferred<false, (DEFER > 0 ? DEFER : 1)>(sc.nodes, sc.tris, src, rm.numThreads, counters + 1);
}

struct ShadowSrc {
    vkx_light light; const float4* queue; float4* rays; uint8_t* shadowFlags; uint32_t count;
    uint32_t ri; float lx, ly, lz, lt;
    __device__ __forceinline__ bool load(uint32_t item, Ray& r, float& tmin, float& tmax, uint32_t& cullMask) {
        if (item >= count) return false;
        const float4 q0 = queue[2 * size_t(item)], q1 = queue[2 * size_t(item) + 1];
        ri = __float_as_uint(q0.w); lx = q1.x; ly = q1.y; lz = q1.z; lt = q1.w;
        r = makeRay(q0.x, q0.y, q0.z, light.direction[0], light.direction[1], light.direction[2]);
        tmin = 0.1f; tmax = 10000.0f; cullMask = 0xFFu; // closesthit.glsl:270-281
        return true;
    }
    __device__ __forceinline__ void store(uint32_t, const HitRec& h, bool) {
        if (!h.found) rays[ri] = make_float4(lx, ly, lz, lt); // the queue entry carries the ray's depth: no read-modify-write (the load was 8 % of this kernel's stall samples)
        if (shadowFlags) shadowFlags[ri] = h.found ? 2 : 1;
    }
};

template <int DEFER>
__global__ void __launch_bounds__(128, PT_MIN_BLOCKS) k_trace_shadow(DeviceScene sc, vkx_light light, const float4* __restrict__ queue, const uint32_t* __restrict__ queueCount,
                                                      float4* __restrict__ rays, uint8_t* __restrict__ shadowFlags, uint32_t* __restrict__ counter) {
    ShadowSrc src; src.light = light; src.queue = queue; src.rays = rays; src.shadowFlags = shadowFlags; src.count = *queueCount; src.ri = 0; src.lx = src.ly = src.lz = src.lt = 0.f;
    if (DEFER == 0) persistentTrace<true>(sc.nodes, sc.tris, src, src.count, counter);
    else if (DEFER < 0) persistentTracePF<true>(sc.nodes, sc.tris, src, src.count, counter);
    else persistentTraceDeferred<true, (DEFER > 0 ? DEFER : 1)>(sc.nodes, sc.tris, src, src.count, counter);
}

// ---- ray-pool variants (ptrace.cuh::poolTrace): K rays per lane in shared memory, node passes and triangle passes at full width
struct PrimaryPoolSrc {
    TraceParams tp; RayMap rm; const float4* origins; const float4* dirs; const float4* invDirs; vkx_hit* hits; float4* rays;
    __device__ __forceinline__ uint32_t rayOf(uint32_t ri) const { return (tp.raysPerProbe & (tp.raysPerProbe - 1u)) == 0u ? ri & (tp.raysPerProbe - 1u) : ri % tp.raysPerProbe; }
    __device__ __forceinline__ bool begin(uint32_t item, uint32_t& tag, float& ox, float& oy, float& oz, float& tmax) {
        uint32_t slot, ray;
        if (!mapRay(rm, item, slot, ray)) return false;
        tag = slot * tp.raysPerProbe + ray;
        const float4 o = __ldg(origins + slot);
        ox = o.x; oy = o.y; oz = o.z; tmax = tp.tmax;
        return true;
    }
    __device__ __forceinline__ void nodeRay(uint32_t tag, float& ix, float& iy, float& iz, uint32_t& octw) const { const float4 id = __ldg(invDirs + rayOf(tag)); ix = id.x; iy = id.y; iz = id.z; octw = __float_as_uint(id.w); }
    __device__ __forceinline__ void triRay(uint32_t tag, float& dx, float& dy, float& dz) const { const float4 d = __ldg(dirs + rayOf(tag)); dx = d.x; dy = d.y; dz = d.z; }
    __device__ __forceinline__ float tmin() const { return tp.tmin; }
    __device__ __forceinline__ uint32_t cullMask() const { return VKX_INSTANCE_STATIC | VKX_INSTANCE_DYNAMIC; }
    __device__ __forceinline__ void finish(uint32_t ri, bool found, float t, float u, float v, uint32_t inst, uint32_t prim) {
        vkx_hit out; out.t = t; out.instance = inst; out.primitive = prim; out.u = u; out.v = v;
        hits[ri] = out;
        if (found && (prim & 0x80000000u)) rays[ri] = make_float4(0.f, 0.f, 0.f, t * 0.80f); // closesthit.glsl:137-141
    }
};
// Shadow rays share one direction: reciprocal / octant are per-kernel constants; the visibility goes to a byte per queue item and
// k_apply_shadow writes the lit colour of the unoccluded ones (no per-ray colour in the pool, no loads on the finishing path).
struct ShadowPoolSrc {
    const float4* queue; uint8_t* visibility; uint32_t count; float dx, dy, dz, ix, iy, iz; uint32_t octw;
    __device__ __forceinline__ bool begin(uint32_t item, uint32_t& tag, float& ox, float& oy, float& oz, float& tmax) {
        if (item >= count) return false;
        const float4 q0 = queue[2 * size_t(item)];
        tag = item; ox = q0.x; oy = q0.y; oz = q0.z; tmax = 10000.0f; // closesthit.glsl:270-281
        return true;
    }
    __device__ __forceinline__ void nodeRay(uint32_t, float& x, float& y, float& z, uint32_t& o) const { x = ix; y = iy; z = iz; o = octw; }
    __device__ __forceinline__ void triRay(uint32_t, float& x, float& y, float& z) const { x = dx; y = dy; z = dz; }
    __device__ __forceinline__ float tmin() const { return 0.1f; }
    __device__ __forceinline__ uint32_t cullMask() const { return 0xFFu; }
    __device__ __forceinline__ void finish(uint32_t item, bool found, float, float, float, uint32_t, uint32_t) { visibility[item] = found ? 2 : 1; }
};

#ifndef POOL_K
#define POOL_K 2
#endif
#ifndef POOL_TRI_THRESH
#define POOL_TRI_THRESH 20
#endif
template <int STK>
__global__ void __launch_bounds__(128) k_trace_primary_pool(DeviceScene sc, TraceParams tp, RayMap rm, const float4* __restrict__ origins, const float4* __restrict__ dirs,
                                                           const float4* __restrict__ invDirs, vkx_hit* __restrict__ hits, float4* __restrict__ rays, uint32_t* __restrict__ counters) {
    extern __shared__ uint32_t sPool[];
    PrimaryPoolSrc src; src.tp = tp; src.rm = rm; src.origins = origins; src.dirs = dirs; src.invDirs = invDirs; src.hits = hits; src.rays = rays;
    poolTrace<false, POOL_K, STK, POOL_TRI_THRESH>(sc.nodes, sc.tris, src, rm.numThreads, counters + 1, sPool + (threadIdx.x >> 5) * PoolLayout<POOL_K, STK>::wordsPerWarp);
}
template <int STK>
__global__ void __launch_bounds__(128) k_trace_shadow_pool(DeviceScene sc, vkx_light light, const float4* __restrict__ queue, const uint32_t* __restrict__ queueCount,
                                                          uint8_t* __restrict__ visibility, uint32_t* __restrict__ counter) {
    extern __shared__ uint32_t sPool[];
    const Ray lr = makeRay(0.f, 0.f, 0.f, light.direction[0], light.direction[1], light.direction[2]);
    ShadowPoolSrc src; src.queue = queue; src.visibility = visibility; src.count = *queueCount;
    src.dx = lr.dx; src.dy = lr.dy; src.dz = lr.dz; src.ix = lr.ix; src.iy = lr.iy; src.iz = lr.iz; src.octw = lr.oct;
    poolTrace<true, POOL_K, STK, POOL_TRI_THRESH>(sc.nodes, sc.tris, src, src.count, counter, sPool + (threadIdx.x >> 5) * PoolLayout<POOL_K, STK>::wordsPerWarp);
}
// Shadow rays that escaped: the ray record becomes the lit colour the shading kernel left in the queue (closesthit.glsl:282-286)
__global__ void k_apply_shadow(const float4* __restrict__ queue, const uint32_t* __restrict__ queueCount, const uint8_t* __restrict__ visibility, float4* __restrict__ rays, uint8_t* __restrict__ shadowFlags) {
    const uint32_t n = *queueCount;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t vis = visibility[i];
        const uint32_t ri = __float_as_uint(queue[2 * size_t(i)].w);
        if (vis == 1u) rays[ri] = queue[2 * size_t(i) + 1];
        if (shadowFlags) shadowFlags[ri] = uint8_t(vis);
    }
}

// ------------------------------------------------------------------------------------------------ blend
#ifndef BLEND_P
#define BLEND_P 8
#endif
#define BLEND_SMEM_BYTES (BLEND_P * (VKX_MAX_RAYS_PER_PROBE * 5 * 4 + 256 * 4 + 64 * 4))
__global__ void __launch_bounds__(BLEND_COLS) k_blend_weights(float depthSharpness, uint32_t N, const float4* __restrict__ dirs, float* __restrict__ W) {
    const uint32_t i = blockIdx.x, col = threadIdx.x;
    if (i >= N) { W[size_t(i) * BLEND_COLS + col] = 0.0f; return; } // padding rows up to a multiple of 4 rays
    const float4 dd = __ldg(dirs + i);
    float w = 0.0f;
    if (col < 196u) {
        const int lx = int(col % 14u), ly = int(col / 14u);
        const v3 td = octDecode(0.142857f * (float(lx) - 6.5f), 0.142857f * (float(ly) - 6.5f));
        w = powf(maxS(0.0f, td.x * dd.x + td.y * dd.y + td.z * dd.z), depthSharpness);
    } else if (col >= BLEND_IRR_COL0 && col < BLEND_IRR_COL0 + 36u) {
        const uint32_t t = col - BLEND_IRR_COL0;
        const int lx = int(t % 6u), ly = int(t / 6u);
        const v3 td = octDecode(0.33333f * (float(lx) - 2.5f), 0.33333f * (float(ly) - 2.5f));
        w = maxS(0.0f, td.x * dd.x + td.y * dd.y + td.z * dd.z);
    }
    W[size_t(i) * BLEND_COLS + col] = w;
}

// Sum of a texel's weights over the rays, in ray order (the `result.w` every blend thread of that texel would accumulate): row BLEND_WSUM_ROW.
__global__ void __launch_bounds__(BLEND_COLS) k_blend_weight_sums(uint32_t N, float* __restrict__ W) {
    float rw = 0.0f;
    uint32_t i = 0;
    for (; i + 64u <= N; i += 64u) { // 64 loads in flight, then the additions in ray order (one block: the loop is latency-bound, 13 us with 16)
        float w[64];
#pragma unroll
        for (uint32_t k = 0; k < 64u; ++k) w[k] = W[size_t(i + k) * BLEND_COLS + threadIdx.x];
#pragma unroll
        for (uint32_t k = 0; k < 64u; ++k) rw = rw + w[k];
    }
    for (; i + 8u <= N; i += 8u) {
        float w[8];
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) w[k] = W[size_t(i + k) * BLEND_COLS + threadIdx.x];
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) rw = rw + w[k];
    }
    for (; i < N; ++i) rw = rw + W[size_t(i) * BLEND_COLS + threadIdx.x];
    W[size_t(BLEND_WSUM_ROW) * BLEND_COLS + threadIdx.x] = rw;
}

// One CTA blends BLEND_P probes. Their ray records are staged in shared memory as five planes (d, d^2, r, g, b); a thread owns
// one (texel, plane) pair of every probe: 2 x 196 depth-moment threads + 3 x 36 irradiance-channel threads (warp-uniform roles,
// BLEND_THREADS = 576). All threads run the same loop: the weight of their texel is loaded once per ray and applied to BLEND_P
// probes from registers, and one 128-bit shared-memory load feeds four rays of a probe, so the loop is 80 % FFMA and every warp
// carries the same load. Accumulation over rays is sequential (i = 0..N-1) like the oracle's, with fused multiply-adds (one
// rounding instead of the oracle's two per term; w * d^2 instead of (w * d) * d: fp32 results differ by ~1e-6 relative, which
// flips an 11-bit packed code on ~1e-5 of the texels). The normalised sums meet again in shared memory for the hysteresis mix.
#define BLEND_THREADS 576
#define BLEND_DEPTH_T0 0     // threads [0, 224): first depth moment, texel = tid
#define BLEND_DEPTH_T1 224   // threads [224, 448): second depth moment
#define BLEND_IRR_T 448      // threads [448, 576): irradiance, channel = t / 36, texel = t % 36
__global__ void __launch_bounds__(BLEND_THREADS) k_blend(BlendParams bp, DeviceProbes pr, const uint32_t* __restrict__ probeIndices, const float4* __restrict__ rays,
                                                         const float* __restrict__ W, float* __restrict__ irrUnpacked, float* __restrict__ depUnpacked, uint32_t slotBase, PeerTargets pt) {
    // dynamic shared memory (50 KB): 5 planes [BLEND_P][256] of ray data, re-used for the normalised sums after the ray loop; output tiles
    extern __shared__ float4 sBlend[];
    float* sPlane = reinterpret_cast<float*>(sBlend);                                // [5][BLEND_P][256]
    float* sResD = sPlane;                                                           // [BLEND_P][2][196]  (aliases the planes)
    float* sResI = sPlane + BLEND_P * 2 * 196;                                       // [BLEND_P][3][36]
    uint32_t (*sDep)[256] = reinterpret_cast<uint32_t (*)[256]>(sPlane + 5 * BLEND_P * VKX_MAX_RAYS_PER_PROBE);
    uint32_t (*sIrr)[64] = reinterpret_cast<uint32_t (*)[64]>(sDep + BLEND_P);
    __shared__ uint32_t sMaxChange[BLEND_P];
    __shared__ uint32_t sOutOfRange[BLEND_P];
    __shared__ uint32_t sLinear[BLEND_P];
    const uint32_t tid = threadIdx.x, N = bp.raysPerProbe;
    const uint32_t N4 = (N + 3u) & ~3u;
    const uint32_t slot0 = blockIdx.x * BLEND_P;
    const uint32_t np = min(uint32_t(BLEND_P), bp.count - slot0);
    const float cellLen = bp.gridCellLen;
    constexpr uint32_t PS = BLEND_P * VKX_MAX_RAYS_PER_PROBE; // plane stride
    if (tid < BLEND_P) { sMaxChange[tid] = 0u; sOutOfRange[tid] = 0u; sLinear[tid] = tid < np ? __ldg(probeIndices + slot0 + tid) : 0u; }
    __syncthreads();
    // stage ray records; count out-of-range rays per probe (probesUpdate.glsl:74) and clamp depths (:78-79) once per ray
    for (uint32_t e = tid; e < BLEND_P * N4; e += BLEND_THREADS) {
        const uint32_t p = e / N4, i = e - p * N4;
        float4 rd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < np && i < N) {
            rd = rays[size_t(slot0 + p) * N + i];
            if (rd.w < 0.0f || rd.w > cellLen) atomicAdd(&sOutOfRange[p], 1u);
            float depth = minS(cellLen, rd.w);
            if (depth < 0.0f) depth = cellLen;
            rd.w = depth;
        }
        float* q = sPlane + p * VKX_MAX_RAYS_PER_PROBE + i;
        q[0] = rd.w; q[PS] = rd.w * rd.w; q[2 * PS] = rd.x; q[3 * PS] = rd.y; q[4 * PS] = rd.z;
    }
    __syncthreads();
    // role of this thread: weight column, data plane, texel
    int plane = -1; uint32_t col = 0, texel = 0;
    if (tid < BLEND_DEPTH_T1) { if (tid < 196u) { plane = 0; texel = tid; col = tid; } }
    else if (tid < BLEND_IRR_T) { if (tid - BLEND_DEPTH_T1 < 196u) { plane = 1; texel = tid - BLEND_DEPTH_T1; col = texel; } }
    else if (tid - BLEND_IRR_T < 108u) { const uint32_t t = tid - BLEND_IRR_T; plane = 2 + int(t / 36u); texel = t % 36u; col = BLEND_IRR_COL0 + texel; }
    float acc[BLEND_P], rw = 0.0f;
#pragma unroll
    for (int p = 0; p < BLEND_P; ++p) acc[p] = 0.0f;
    if (plane >= 0) {
        const float* Wc = W + col;
        const float* src = sPlane + size_t(plane) * PS;
        rw = __ldg(Wc + size_t(BLEND_WSUM_ROW) * BLEND_COLS);
#pragma unroll 1
        for (uint32_t i = 0; i < N4; i += 4) { // rays beyond N are zero in the planes and in the (padded) weight table
            const float w0 = __ldg(Wc + size_t(i) * BLEND_COLS), w1 = __ldg(Wc + size_t(i + 1) * BLEND_COLS);
            const float w2 = __ldg(Wc + size_t(i + 2) * BLEND_COLS), w3 = __ldg(Wc + size_t(i + 3) * BLEND_COLS);
#pragma unroll
            for (int p = 0; p < BLEND_P; ++p) {
                const float4 m = *reinterpret_cast<const float4*>(src + p * VKX_MAX_RAYS_PER_PROBE + i);
                acc[p] = fmaf(w0, m.x, acc[p]); acc[p] = fmaf(w1, m.y, acc[p]); acc[p] = fmaf(w2, m.z, acc[p]); acc[p] = fmaf(w3, m.w, acc[p]);
            }
        }
    }
    __syncthreads(); // the planes are dead from here on: their memory now holds the normalised sums
    if (plane >= 0) {
#pragma unroll
        for (int p = 0; p < BLEND_P; ++p) {
            float r = acc[p];
            if (rw > 1e-3f) r = r / rw;
            if (plane < 2) sResD[(p * 2 + plane) * 196 + texel] = r; else sResI[(p * 3 + (plane - 2)) * 36 + texel] = r;
        }
    }
    __syncthreads();
    const float hysteresis = bp.grid.hysteresis;
    if (tid < 196u) { // ---- depth texels: hysteresis mix against the work atlas, pack
        const int lx = int(tid % 14u), ly = int(tid / 14u);
        for (uint32_t p = 0; p < np; ++p) {
            const float r0 = sResD[(p * 2 + 0) * 196 + tid], r1 = sResD[(p * 2 + 1) * 196 + tid];
            int ix, iy, iz; probeGridIndex(sLinear[p], bp.grid, ix, iy, iz);
            const int tile = iy * bp.grid.resolution[0] + ix;
            const size_t gi = size_t(16 * iz + 1 + ly) * pr.depW + size_t(16 * tile + 1 + lx);
            const float2 prev = unpackRG16F(pr.depWork[gi]);
            const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis);
            sDep[p][(ly + 1) * 16 + (lx + 1)] = packRG16F(o0, o1);
            if (depUnpacked) { float* up = depUnpacked + (size_t(slotBase + slot0 + p) * 196 + tid) * 2; up[0] = o0; up[1] = o1; }
        }
    } else if (tid >= BLEND_DEPTH_T1 && tid < BLEND_DEPTH_T1 + 36u) { // ---- irradiance texels (a different warp than the depth texels)
        const uint32_t t = tid - BLEND_DEPTH_T1;
        const int lx = int(t % 6u), ly = int(t / 6u);
        for (uint32_t p = 0; p < np; ++p) {
            const float r0 = sResI[(p * 3 + 0) * 36 + t], r1 = sResI[(p * 3 + 1) * 36 + t], r2 = sResI[(p * 3 + 2) * 36 + t];
            int ix, iy, iz; probeGridIndex(sLinear[p], bp.grid, ix, iy, iz);
            const int tile = iy * bp.grid.resolution[0] + ix;
            const size_t gi = size_t(8 * iz + 1 + ly) * pr.irrW + size_t(8 * tile + 1 + lx);
            const float3 prev = unpackR11G11B10(pr.irrWork[gi]);
            const float maxChange = maxS(maxS(fabsf(r0 - prev.x), fabsf(r1 - prev.y)), fabsf(r2 - prev.z));
            const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis), o2 = mixf(r2, prev.z, hysteresis);
            sIrr[p][(ly + 1) * 8 + (lx + 1)] = packR11G11B10(o0, o1, o2);
            if (irrUnpacked) { float* up = irrUnpacked + (size_t(slotBase + slot0 + p) * 36 + t) * 3; up[0] = o0; up[1] = o1; up[2] = o2; }
            atomicMax(&sMaxChange[p], __float_as_uint(maxChange)); // probesUpdate.glsl:106-107 (non-negative floats order as uints)
        }
    }
    __syncthreads();
    if (tid < np) { // state machine, probesUpdate.glsl:110-119 (decree A.5.3: full max over the 36 texels)
        const uint32_t linearIndex = sLinear[tid];
        uint32_t st = pr.stateWork[linearIndex];
        if (sOutOfRange[tid] >= N) st = 8;
        else {
            const float maxChange = __uint_as_float(sMaxChange[tid]);
            if (maxChange < 0.02f / float(st)) st = min(st + 1u, 8u);
            else if (maxChange > 0.04f / float(st)) st = max(st - 1u, 1u);
            else if (maxChange > 0.25f) st = 1;
        }
        pr.stateWork[linearIndex] = st;
        for (int r = 0; r < pt.n; ++r) pt.state[r][linearIndex] = st; // sharded + peer memory: straight into every rank's next state array
    }
    // ---- borders (probesCopyBorders.comp) from the shared tiles: 60 depth + 28 irradiance texels per probe
    for (uint32_t e = tid; e < np * 88u; e += BLEND_THREADS) {
        const uint32_t p = e / 88u, b = e - p * 88u;
        int x, y, sx, sy;
        if (b < 60u) {
            if (b < 16u) { x = int(b); y = 0; } else if (b < 32u) { x = int(b) - 16; y = 15; } else if (b < 46u) { x = 0; y = int(b) - 32 + 1; } else { x = 15; y = int(b) - 46 + 1; }
            blendBorderSource(16, x, y, sx, sy);
            sDep[p][y * 16 + x] = sDep[p][sy * 16 + sx];
        } else {
            const int c = int(b) - 60;
            if (c < 8) { x = c; y = 0; } else if (c < 16) { x = c - 8; y = 7; } else if (c < 22) { x = 0; y = c - 16 + 1; } else { x = 7; y = c - 22 + 1; }
            blendBorderSource(8, x, y, sx, sy);
            sIrr[p][y * 8 + x] = sIrr[p][sy * 8 + sx];
        }
    }
    __syncthreads();
    // ---- vectorised tile stores: per probe depth 16 rows x 64 B (64 uint4), irradiance 8 rows x 32 B (16 uint4)
    for (uint32_t e = tid; e < np * 80u; e += BLEND_THREADS) {
        const uint32_t p = e / 80u, k = e - p * 80u;
        int ix, iy, iz; probeGridIndex(sLinear[p], bp.grid, ix, iy, iz);
        const int tile = iy * bp.grid.resolution[0] + ix;
        if (k < 64u) {
            const int row = int(k >> 2), q = int(k & 3u);
            const uint4 v = *reinterpret_cast<const uint4*>(&sDep[p][row * 16 + q * 4]);
            const size_t off = size_t(16 * iz + row) * pr.depW + size_t(16 * tile + q * 4);
            *reinterpret_cast<uint4*>(pr.depWork + off) = v;
            for (int r = 0; r < pt.n; ++r) *reinterpret_cast<uint4*>(pt.dep[r] + off) = v; // NVLink peer stores (or the local next set)
        } else {
            const int kk = int(k) - 64; const int row = kk >> 1, q = kk & 1;
            const uint4 v = *reinterpret_cast<const uint4*>(&sIrr[p][row * 8 + q * 4]);
            const size_t off = size_t(8 * iz + row) * pr.irrW + size_t(8 * tile + q * 4);
            *reinterpret_cast<uint4*>(pr.irrWork + off) = v;
            for (int r = 0; r < pt.n; ++r) *reinterpret_cast<uint4*>(pt.irr[r] + off) = v;
        }
    }
    // (no fence here: the arrival flag is raised by the next kernel on this stream, i.e. after this grid and all its stores have completed;
    // a per-CTA __threadfence_system() made every CTA wait for its NVLink write acknowledgements: blend 0.25 -> 0.32 ms on 2 GPUs)
}

// Arrival flags of the peer-memory exchange. After its last blend of frame g a rank writes g + 1 into slot [rank] of every rank's flag
// array; before the first kernel that reads the sampled atlases a rank waits until all slots of its own array have reached the frame
// it is about to shade. No host round trip and no collective: the wait is a one-warp kernel polling local memory.
__global__ void k_p2p_signal(PeerFlags pf, int n, int self, uint32_t value) {
    const int r = int(threadIdx.x);
    if (r < n) { __threadfence_system(); *reinterpret_cast<volatile uint32_t*>(pf.flags[r] + self) = value; }
}
__global__ void k_p2p_wait(const uint32_t* flags, int n, uint32_t need, uint32_t* err) {
    const int r = int(threadIdx.x);
    if (r >= n) return;
    const volatile uint32_t* f = flags + r;
    const long long t0 = clock64();
    while (*f < need) {
        if (clock64() - t0 > (20LL << 30)) { *err = 1u + uint32_t(r); break; } // ~10 s at 2 GHz: a peer died; do not hang the GPU
        __nanosleep(500);
    }
}

// Multi-chunk updates: probe indices gathered in block order, so that a chunk is a contiguous piece of one list.
__global__ void k_gather_list(uint32_t* __restrict__ out, const uint32_t* __restrict__ list, const uint32_t* __restrict__ order, uint32_t n) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = list[order[j]];
}

// work -> sampled for the updated probes (tiles + state word). One warp per probe: 16 depth rows + 8 irradiance rows.
__global__ void k_publish(DeviceProbes pr, uint32_t* __restrict__ irrSampled, uint32_t* __restrict__ depSampled, uint32_t* __restrict__ stateSampled,
                          const uint32_t* __restrict__ probeIndices, uint32_t count) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= count) return;
    const uint32_t linearIndex = __ldg(probeIndices + warp);
    int ix, iy, iz; probeGridIndex(linearIndex, pr.grid, ix, iy, iz);
    const int tile = iy * pr.grid.resolution[0] + ix;
    // 64 uint4 of depth (2 per lane) + 16 uint4 of irradiance (lanes 0..15)
    for (int k = int(lane); k < 64; k += 32) {
        const int row = k >> 2, q = k & 3;
        const size_t off = size_t(16 * iz + row) * pr.depW + size_t(16 * tile + q * 4);
        *reinterpret_cast<uint4*>(depSampled + off) = *reinterpret_cast<const uint4*>(pr.depWork + off);
    }
    if (lane < 16) {
        const int row = int(lane >> 1), q = int(lane & 1);
        const size_t off = size_t(8 * iz + row) * pr.irrW + size_t(8 * tile + q * 4);
        *reinterpret_cast<uint4*>(irrSampled + off) = *reinterpret_cast<const uint4*>(pr.irrWork + off);
    }
    if (lane == 0) stateSampled[linearIndex] = pr.stateWork[linearIndex];
}

// Sharded list update: the tiles of an updated probe as one 1296-byte record (64 uint4 depth rows, 16 uint4 irradiance rows, state),
// so that ranks can exchange the probes they updated with one all-gather. One warp per probe, like k_publish.
#define VKX_TILE_RECORD 81u // uint4 per probe
__global__ void k_pack_tiles(DeviceProbes pr, const uint32_t* __restrict__ probeIndices, uint32_t count, uint4* __restrict__ packed) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= count) return;
    const uint32_t linearIndex = __ldg(probeIndices + warp);
    int ix, iy, iz; probeGridIndex(linearIndex, pr.grid, ix, iy, iz);
    const int tile = iy * pr.grid.resolution[0] + ix;
    uint4* rec = packed + size_t(warp) * VKX_TILE_RECORD;
    for (int k = int(lane); k < 64; k += 32) rec[k] = *reinterpret_cast<const uint4*>(pr.depWork + size_t(16 * iz + (k >> 2)) * pr.depW + size_t(16 * tile + (k & 3) * 4));
    if (lane < 16) rec[64 + lane] = *reinterpret_cast<const uint4*>(pr.irrWork + size_t(8 * iz + int(lane >> 1)) * pr.irrW + size_t(8 * tile + int(lane & 1) * 4));
    if (lane == 16) rec[80] = make_uint4(pr.stateWork[linearIndex], linearIndex, 0u, 0u);
}
// List position s was processed by rank s / perRank as its item s % perRank; its record goes to the work and the sampled set.
__global__ void k_unpack_tiles(DeviceProbes pr, uint32_t* __restrict__ irrSampled, uint32_t* __restrict__ depSampled, uint32_t* __restrict__ stateSampled,
                               const uint32_t* __restrict__ probeIndices, uint32_t count, uint32_t perRank, const uint4* __restrict__ packed) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= count) return;
    const uint32_t linearIndex = __ldg(probeIndices + warp);
    int ix, iy, iz; probeGridIndex(linearIndex, pr.grid, ix, iy, iz);
    const int tile = iy * pr.grid.resolution[0] + ix;
    const uint4* rec = packed + size_t(warp) * VKX_TILE_RECORD; // records are stored rank after rank with perRank slots each = list order
    (void)perRank;
    for (int k = int(lane); k < 64; k += 32) {
        const size_t off = size_t(16 * iz + (k >> 2)) * pr.depW + size_t(16 * tile + (k & 3) * 4);
        const uint4 v = rec[k];
        *reinterpret_cast<uint4*>(pr.depWork + off) = v; *reinterpret_cast<uint4*>(depSampled + off) = v;
    }
    if (lane < 16) {
        const size_t off = size_t(8 * iz + int(lane >> 1)) * pr.irrW + size_t(8 * tile + int(lane & 1) * 4);
        const uint4 v = rec[64 + lane];
        *reinterpret_cast<uint4*>(pr.irrWork + off) = v; *reinterpret_cast<uint4*>(irrSampled + off) = v;
    }
    if (lane == 16) { const uint32_t st = rec[80].x; pr.stateWork[linearIndex] = st; stateSampled[linearIndex] = st; }
}

// ------------------------------------------------------------------------------------------------ classification
// probesInit.rgen:31-64 + backfaceTest.rchit + probeInitMiss.rmiss. One CTA of 128 threads per probe, 4 rays each.
__global__ void __launch_bounds__(128) k_classify(DeviceScene sc, vkx_grid_info grid, const float4* __restrict__ dirs512, uint32_t* __restrict__ stateWork,
                                                  uint32_t* __restrict__ stateSampled) {
    __shared__ uint32_t sBack, sAffect;
    const uint32_t li = blockIdx.x;
    if (threadIdx.x == 0) { sBack = 0; sAffect = 0; }
    __syncthreads();
    int ix, iy, iz; probeGridIndex(li, grid, ix, iy, iz);
    const v3 origin = probeWorldPos(ix, iy, iz, grid);
    const v3 cell = gridCellSize(grid);
    const float maxDistance = len3(cell);
    const float tmax = 1.5f * maxDistance;
    uint32_t back = 0, affect = 0;
    for (uint32_t i = threadIdx.x; i < 512u; i += blockDim.x) {
        const float4 d = __ldg(dirs512 + i);
        const Ray r = makeRay(origin.x, origin.y, origin.z, d.x, d.y, d.z);
        HitRec h;
        float depth = 3.402823466e+38f; bool isBack = false;
        if (traverse<false>(sc.nodes, sc.tris, r, 0.01f, tmax, VKX_INSTANCE_STATIC, h)) { depth = h.t; isBack = (h.prim & 0x80000000u) != 0u; }
        if (depth < maxDistance) {
            if (isBack) ++back;
            const v3 position = origin + depth * mk3(d.x, d.y, d.z);
            const v3 dist = abs3(position - origin);
            if (dist.x < cell.x && dist.y < cell.y && dist.z < cell.z) affect = 1;
        }
    }
    if (back) atomicAdd(&sBack, back);
    if (affect) atomicOr(&sAffect, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t st;
        if (float(sBack) > 0.5f * float(grid.raysPerProbe)) st = 0;
        else st = sAffect ? 1u : 8u;
        stateWork[li] = st; stateSampled[li] = st;
    }
}

} // namespace

DeviceScene deviceScene(const vkx_ctx* ctx) {
    DeviceScene s;
    s.vertices = ctx->dVertices; s.indices = ctx->dIndices; s.offsets = ctx->dOffsets; s.materials = ctx->dMaterials; s.instances = ctx->dInstances;
    s.worldToObject = ctx->dWorldToObject; s.nodes = ctx->dNodes; s.tris = ctx->dTris;
    s.texels = ctx->dTexels; s.texelsDecoded = ctx->dTexelsDecoded; s.textures = ctx->dTextures; s.srgbLut = ctx->dSrgbLut; s.numTextures = uint32_t(ctx->hTextures.size());
    return s;
}

DeviceProbes deviceProbes(const vkx_ctx* ctx) {
    DeviceProbes p;
    p.grid = ctx->grid; p.irrW = ctx->irrW; p.irrH = ctx->irrH; p.depW = ctx->depW; p.depH = ctx->depH; p.probeCount = ctx->probeCount;
    p.irrSampled = ctx->dIrrSampled; p.depSampled = ctx->dDepSampled; p.stateSampled = ctx->dStateSampled;
    p.irrWork = ctx->dIrrWork; p.depWork = ctx->dDepWork; p.stateWork = ctx->dStateWork;
    return p;
}

int ddgiClassify(vkx_ctx* ctx, const float* dirs512) {
    cudaStream_t st = ctx->stream;
    std::vector<float4> d(512);
    for (int i = 0; i < 512; ++i) d[i] = make_float4(dirs512[3 * i], dirs512[3 * i + 1], dirs512[3 * i + 2], 0.f);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dDirs, d.data(), 512 * sizeof(float4), cudaMemcpyHostToDevice, st));
    k_classify<<<ctx->probeCount, 128, 0, st>>>(deviceScene(ctx), ctx->grid, ctx->dDirs, ctx->dStateWork, ctx->dStateSampled);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return VKX_OK;
}

// Traces + blends `count` probes whose indices are already in ctx->dIndicesList[listOffset ...]. Directions are in ctx->dDirs.
int ddgiUpdate(vkx_ctx* ctx, const vkx_light& light, const uint32_t* /*unused*/, uint32_t count, uint32_t listOffset, bool /*unused*/) {
    cudaStream_t st = ctx->stream;
    const uint32_t N = ctx->grid.raysPerProbe;
    const DeviceScene sc = deviceScene(ctx);
    const DeviceProbes pr = deviceProbes(ctx);
    const float ex = ctx->grid.extentMax[0] - ctx->grid.extentMin[0], ey = ctx->grid.extentMax[1] - ctx->grid.extentMin[1], ez = ctx->grid.extentMax[2] - ctx->grid.extentMin[2];
    const float tmax = sqrtf(ex * ex + ey * ey + ez * ez); // traceProbes.rgen:33
    const float cx = ex / float(ctx->grid.resolution[0] - 1), cy = ey / float(ctx->grid.resolution[1] - 1), cz = ez / float(ctx->grid.resolution[2] - 1);
    BlendParams bp; bp.grid = ctx->grid; bp.raysPerProbe = N; bp.gridCellLen = sqrtf(cx * cx + cy * cy + cz * cz);
    // function attributes and occupancy are per device: cached in the context, not in process-wide statics
    if (!ctx->blendAttrSet) { CUDA_TRY(ctx, cudaFuncSetAttribute(k_blend, cudaFuncAttributeMaxDynamicSharedMemorySize, BLEND_SMEM_BYTES)); ctx->blendAttrSet = true; }
    // Leaf deferral of the two persistent traversals (ptrace.cuh): 0 = off, else the number of waiting lanes that triggers a triangle phase.
    // Tuning knobs (results are identical for every value): VKX_PT_DEFER / VKX_PT_DEFER_SHADOW in {0, 8, 12, 16}.
    static int deferPrimary = -1, deferShadow = -1;
    static const bool poolMode = [] { const char* e = getenv("VKX_PT_POOL"); return e ? atoi(e) != 0 : false; }();
    if (deferPrimary < 0) {
        auto pick = [](const char* name, int dflt) { const char* e = getenv(name); int v = e ? atoi(e) : dflt; return v < 0 ? -1 : v == 0 ? 0 : v <= 8 ? 8 : v <= 12 ? 12 : 16; }; // -1: prefetching variant
        deferPrimary = pick("VKX_PT_DEFER", PT_DEFER_PRIMARY_DEFAULT); deferShadow = pick("VKX_PT_DEFER_SHADOW", PT_DEFER_SHADOW_DEFAULT);
    }
    if (!ctx->traceBlocksPerSm) { int a = 0, b = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace_primary<0>, 128, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_shadow<0>, 128, 0); ctx->traceBlocksPerSm = std::max(1, std::min(a, b)); }
    const unsigned persistentBlocks = unsigned(ctx->smCount * ctx->traceBlocksPerSm);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
    // The frame's blend weights (three small kernels, one of them a single block) are only read by the blend at the end of the
    // update: they run on the second stream, beside the set-up and the traversal, and are joined together with the sky kernel.
    cudaStream_t ax = ctx->auxStream;
    CUDA_TRY(ctx, cudaEventRecord(ctx->auxEvent[0], st)); // after the directions' upload and after the previous update's blend (which reads the tables)
    CUDA_TRY(ctx, cudaStreamWaitEvent(ax, ctx->auxEvent[0], 0));
    k_blend_weights<<<(N + 3u) & ~3u, BLEND_COLS, 0, ax>>>(ctx->grid.depthSharpness, N, ctx->dDirs, ctx->dBlendW); LAUNCH_CHECK(ctx);
    k_blend_weight_sums<<<1, BLEND_COLS, 0, ax>>>(N, ctx->dBlendW); LAUNCH_CHECK(ctx);
    // blend on the tensor cores (blend_tc.cu) unless tiles go straight to peer memory (fused exchange) or VKX_BLEND=simt asks for the CUDA-core kernel
    static const bool blendSimt = [] { const char* e = getenv("VKX_BLEND"); return e && !strcmp(e, "simt"); }();
    const bool blendTc = !blendSimt && !ctx->blendToPeers;
    if (blendTc) { int rc = blendTcWeights(ctx, ax); if (rc != VKX_OK) return rc; }
    k_dir_table<<<divUp(N, 128), 128, 0, st>>>(N, ctx->dDirs, ctx->dInvDirs); LAUNCH_CHECK(ctx);
    // One chunk: slots are the caller's list positions (ray/hit buffers are laid out [slot][ray]) and `order` only schedules them.
    // Several chunks: the list is first gathered in block order, a chunk is then a contiguous piece of it with identity order.
    const bool multi = count > ctx->chunkProbes;
    if (multi) { k_gather_list<<<divUp(count, 256), 256, 0, st>>>(ctx->dPermList, ctx->dIndicesList + listOffset, ctx->dOrder + listOffset, count); LAUNCH_CHECK(ctx); }
    for (uint32_t base = 0; base < count; base += ctx->chunkProbes) {
        const uint32_t n = std::min(ctx->chunkProbes, count - base);
        const uint32_t numRays = n * N;
        const uint32_t* idx = multi ? ctx->dPermList + base : ctx->dIndicesList + listOffset;
        RayMap rm; rm.count = n; rm.raysPerProbe = N; rm.numDirGroups = (N + 3u) / 4u; rm.order = multi ? ctx->dIota : ctx->dOrder + listOffset; rm.perm = ctx->dPerm;
        rm.numThreads = ((n + 7u) / 8u) * rm.numDirGroups * 32u;
        rm.dgShift = 0xFFFFFFFFu; for (uint32_t b = 0; b < 31; ++b) if ((1u << b) == rm.numDirGroups) rm.dgShift = b;
        TraceParams tp; tp.grid = ctx->grid; tp.tmin = 0.01f; tp.tmax = tmax; tp.raysPerProbe = N; tp.numRays = numRays;
        tp.invCell[0] = 1.0f / cx; tp.invCell[1] = 1.0f / cy; tp.invCell[2] = 1.0f / cz;
        ShadeParams sp; sp.grid = ctx->grid; sp.light = light; sp.raysPerProbe = N; sp.numRays = numRays;
        const bool timed = base == 0; // per-kernel events on the first chunk
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->dQueueCount, 0, 32, st)); // [0] shadow queue length, [1] primary work counter, [2] shadow work counter, [3] misses, [4] front hits
        k_origin_table<<<divUp(n, 128), 128, 0, st>>>(ctx->grid, idx, n, ctx->dOrigins); LAUNCH_CHECK(ctx);
        if (timed) { CUDA_TRY(ctx, cudaEventRecord(ctx->kev[0], st)); ctx->kevProbes = n; }
        // ray-pool traversal (VKX_PT_POOL, default on): needs a shared-memory stack of depth - 1 group entries per ray
        const int poolStk = ctx->bvh.depth <= 9u ? 8 : ctx->bvh.depth <= 13u ? 12 : 0;
        const bool pool = poolMode && poolStk;
        const size_t poolBytes = size_t(4) * (poolStk == 8 ? PoolLayout<POOL_K, 8>::wordsPerWarp : PoolLayout<POOL_K, 12>::wordsPerWarp) * 4;
        unsigned poolBlocks = 0;
        if (pool) {
            int& cached = poolStk == 8 ? ctx->poolBlocksPerSm[0] : ctx->poolBlocksPerSm[1];
            if (!cached) {
                int a = 0, b = 0;
                if (poolStk == 8) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace_primary_pool<8>, 128, poolBytes); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_shadow_pool<8>, 128, poolBytes); }
                else { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace_primary_pool<12>, 128, poolBytes); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_shadow_pool<12>, 128, poolBytes); }
                cached = std::max(1, std::min(a, b));
            }
            poolBlocks = unsigned(ctx->smCount * cached);
        }
        if (base == 0) { const int rc = flushCopyRequests(ctx); if (rc != VKX_OK) return rc; } // requested read-backs of the sampled atlases run alongside the traversal (api.cu)
        // classification: misses -> sky queue, front hits -> key (grid cell of the hit point) for the grouping below
        static const bool radixSort = [] { const char* e = getenv("VKX_SORT"); return e && !strcmp(e, "radix"); }(); // A/B: the round-1 path (cub::DeviceRadixSort)
        uint32_t binShift = 0; while (((ctx->probeCount - 1u) >> binShift) + 1u > BIN_MAX) ++binShift;
        const uint32_t numBins = ((ctx->probeCount - 1u) >> binShift) + 1u;
        const size_t binBytes = size_t(numBins) * 4;
        if (pool) {
            if (poolStk == 8) k_trace_primary_pool<8><<<poolBlocks, 128, poolBytes, st>>>(sc, tp, rm, ctx->dOrigins, ctx->dDirs, ctx->dInvDirs, ctx->dHits, ctx->dRays, ctx->dQueueCount);
            else k_trace_primary_pool<12><<<poolBlocks, 128, poolBytes, st>>>(sc, tp, rm, ctx->dOrigins, ctx->dDirs, ctx->dInvDirs, ctx->dHits, ctx->dRays, ctx->dQueueCount);
        } else
#define VKX_LAUNCH_PRIMARY(D) k_trace_primary<D><<<persistentBlocks, 128, 0, st>>>(sc, tp, rm, ctx->dOrigins, ctx->dDirs, ctx->dInvDirs, ctx->dHits, ctx->dRays, ctx->dQueueCount)
        switch (deferPrimary) { case -1: VKX_LAUNCH_PRIMARY(-1); break; case 8: VKX_LAUNCH_PRIMARY(8); break; case 12: VKX_LAUNCH_PRIMARY(12); break; case 16: VKX_LAUNCH_PRIMARY(16); break; default: VKX_LAUNCH_PRIMARY(0); }
#undef VKX_LAUNCH_PRIMARY
        LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[1], st));
        if (radixSort) {
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->dFrontKeys, 0xFF, size_t(numRays) * 4, st)); // unused slots sort to the end
            k_classify_hits<<<std::min<unsigned>(divUp(numRays, 256), unsigned(ctx->smCount) * 16u), 256, 0, st>>>(tp, ctx->dOrigins, ctx->dDirs, ctx->dHits, ctx->dMissQueue, ctx->dFrontQueue, ctx->dFrontKeys, ctx->dQueueCount); LAUNCH_CHECK(ctx);
        } else {
            if (!ctx->binAttrSet) {
                CUDA_TRY(ctx, cudaFuncSetAttribute(k_bin_count, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BIN_MAX * 4)));
                CUDA_TRY(ctx, cudaFuncSetAttribute(k_bin_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BIN_MAX * 4)));
                ctx->binAttrSet = true;
            }
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->dCellHist, 0, binBytes, st));
            const unsigned binBlocks = std::min<unsigned>(divUp(numRays, BIN_TILE), unsigned(ctx->smCount) * 3u);
            k_bin_count<<<binBlocks, 256, binBytes, st>>>(tp, numBins, binShift, ctx->dOrigins, ctx->dDirs, ctx->dHits, ctx->dMissQueue, ctx->dFrontKeys, ctx->dCellHist, ctx->dQueueCount); LAUNCH_CHECK(ctx);
        }
        // The sky kernel only needs the miss queue: it runs on a second stream, concurrently with the sort and the front-hit shading.
        const unsigned shadeBlocks = std::min<unsigned>(divUp(numRays, 128), unsigned(ctx->smCount) * 16u);
        CUDA_TRY(ctx, cudaEventRecord(ctx->auxEvent[0], st));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->auxStream, ctx->auxEvent[0], 0));
        launchShadeMiss(shadeBlocks, ctx->auxStream, sp, ctx->dOrigins, ctx->dDirs, ctx->dMissQueue, ctx->dQueueCount, ctx->dRays); LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaEventRecord(ctx->auxEvent[1], ctx->auxStream));
        if (radixSort) { // front-hit queue sorted by grid cell (radix sort over just the bits a cell index needs)
            uint32_t cells = ctx->probeCount, bits = 1; while ((1u << bits) <= cells) ++bits;
            size_t need = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->dFrontKeys, ctx->dFrontKeysOut, ctx->dFrontQueue, ctx->dFrontQueueSorted, int(numRays), 0, int(bits), st);
            if (need > ctx->sortTempBytes) { if (ctx->dSortTemp) cudaFree(ctx->dSortTemp); CUDA_TRY(ctx, cudaMalloc(&ctx->dSortTemp, need)); ctx->sortTempBytes = need; }
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->dSortTemp, need, ctx->dFrontKeys, ctx->dFrontKeysOut, ctx->dFrontQueue, ctx->dFrontQueueSorted, int(numRays), 0, int(bits), st));
            ctx->launches += 3;
        } else { // front-hit queue grouped by bin: first positions, then the scatter (see k_bin_count)
            k_bin_scan<<<1, 1024, 0, st>>>(numBins, ctx->dCellHist, ctx->dFrontKeysOut); LAUNCH_CHECK(ctx);
            k_bin_scatter<<<std::min<unsigned>(divUp(numRays, SCAT_TILE), unsigned(ctx->smCount) * 2u), SCAT_THREADS, binBytes, st>>>(numRays, numBins, ctx->dFrontKeys, ctx->dFrontKeysOut, ctx->dFrontQueueSorted); LAUNCH_CHECK(ctx);
        }
        { int rc = waitGather(ctx); if (rc != VKX_OK) return rc; } // sharded path: the previous frame's atlas all-gather must have landed
        launchShadeFront(shadeBlocks, st, sc, pr, sp, ctx->dOrigins, ctx->dDirs, ctx->dHits, ctx->dFrontQueueSorted, ctx->dQueueCount, ctx->dRays, ctx->dShadowQueue); LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[2], st));
        if (ctx->debugBuffers) CUDA_TRY(ctx, cudaMemsetAsync(ctx->dShadowFlags, 0, numRays, st));
        if (pool) {
            if (poolStk == 8) k_trace_shadow_pool<8><<<poolBlocks, 128, poolBytes, st>>>(sc, light, ctx->dShadowQueue, ctx->dQueueCount, ctx->dShadowVis, ctx->dQueueCount + 2);
            else k_trace_shadow_pool<12><<<poolBlocks, 128, poolBytes, st>>>(sc, light, ctx->dShadowQueue, ctx->dQueueCount, ctx->dShadowVis, ctx->dQueueCount + 2);
            LAUNCH_CHECK(ctx);
            k_apply_shadow<<<std::min<unsigned>(divUp(numRays, 256), unsigned(ctx->smCount) * 16u), 256, 0, st>>>(ctx->dShadowQueue, ctx->dQueueCount, ctx->dShadowVis, ctx->dRays, ctx->debugBuffers ? ctx->dShadowFlags : nullptr);
        } else
#define VKX_LAUNCH_SHADOW(D) k_trace_shadow<D><<<persistentBlocks, 128, 0, st>>>(sc, light, ctx->dShadowQueue, ctx->dQueueCount, ctx->dRays, ctx->debugBuffers ? ctx->dShadowFlags : nullptr, ctx->dQueueCount + 2)
        switch (deferShadow) { case -1: VKX_LAUNCH_SHADOW(-1); break; case 8: VKX_LAUNCH_SHADOW(8); break; case 12: VKX_LAUNCH_SHADOW(12); break; case 16: VKX_LAUNCH_SHADOW(16); break; default: VKX_LAUNCH_SHADOW(0); }
#undef VKX_LAUNCH_SHADOW
        LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[3], st));
        if (base + n >= count) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
        bp.count = n;
        CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->auxEvent[1], 0)); // sky results
        if (ctx->copyPending && ctx->copyReadsWork) { CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->evCopyDone, 0)); ctx->copyPending = false; ctx->copyReadsWork = false; } // a queued read-back still reads the work atlases
        if (blendTc) { int rc = blendTcLaunch(ctx, bp, pr, idx, n, base, st); if (rc != VKX_OK) return rc; }
        else k_blend<<<divUp(n, BLEND_P), BLEND_THREADS, BLEND_SMEM_BYTES, st>>>(bp, pr, idx, ctx->dRays, ctx->dBlendW, ctx->debugBuffers ? ctx->dIrrUnpacked : nullptr, ctx->debugBuffers ? ctx->dDepUnpacked : nullptr, base, ctx->blendToPeers ? ctx->blendPeers : PeerTargets{}); LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[4], st));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
    ctx->lastCount = count; ctx->lastRays = count * N;
    return VKX_OK;
}

int launchP2pSignal(vkx_ctx* ctx) {
    PeerFlags pf{};
    for (int r = 0; r < ctx->nranks; ++r) pf.flags[r] = reinterpret_cast<uint32_t*>(ctx->peerSlab[r] + ctx->p2pFlagsOff);
    k_p2p_signal<<<1, 32, 0, ctx->stream>>>(pf, ctx->nranks, ctx->rank, ctx->p2pFrame + 1u); LAUNCH_CHECK(ctx);
    return VKX_OK;
}
int launchP2pWait(vkx_ctx* ctx) {
    uint32_t* flags = reinterpret_cast<uint32_t*>(ctx->p2pSlab + ctx->p2pFlagsOff);
    k_p2p_wait<<<1, 32, 0, ctx->stream>>>(flags, ctx->nranks, ctx->p2pFrame, flags + 64); LAUNCH_CHECK(ctx);
    return VKX_OK;
}

int ddgiPublish(vkx_ctx* ctx, uint32_t count) {
    cudaStream_t st = ctx->stream;
    const DeviceProbes pr = deviceProbes(ctx);
    if (count) { k_publish<<<divUp(size_t(count) * 32, 256), 256, 0, st>>>(pr, ctx->dIrrSampled, ctx->dDepSampled, ctx->dStateSampled, ctx->dIndicesList, count); LAUNCH_CHECK(ctx); }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    return VKX_OK;
}

int ddgiPackTiles(vkx_ctx* ctx, const uint32_t* probeIndices, uint32_t count, uint4* packed, cudaStream_t st) {
    if (count) { k_pack_tiles<<<divUp(size_t(count) * 32, 256), 256, 0, st>>>(deviceProbes(ctx), probeIndices, count, packed); LAUNCH_CHECK(ctx); }
    return VKX_OK;
}
int ddgiUnpackTiles(vkx_ctx* ctx, const uint32_t* probeIndices, uint32_t count, uint32_t perRank, const uint4* packed, cudaStream_t st) {
    if (count) { k_unpack_tiles<<<divUp(size_t(count) * 32, 256), 256, 0, st>>>(deviceProbes(ctx), ctx->dIrrSampled, ctx->dDepSampled, ctx->dStateSampled, probeIndices, count, perRank, packed); LAUNCH_CHECK(ctx); }
    return VKX_OK;
}
