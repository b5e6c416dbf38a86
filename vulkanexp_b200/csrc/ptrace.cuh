// Persistent-thread traversal with warp-level ray replacement ("dynamic fetch"): warps stay resident, and a lane whose ray has
// terminated takes the next ray from the work list instead of idling until the slowest ray of its warp finishes. Measured
// motivation (profiles/r01a): with one ray per thread 11 of 32 lanes were active in the node test because traversal lengths
// inside a warp differ by 3x. The per-ray operation sequence is unchanged (it is the one of traverse<> in traverse.cuh);
// only the assignment of rays to lanes is dynamic.
//
// Work distribution: each warp owns a private chunk of PT_CHUNK consecutive work items taken from a global counter with one
// atomicAdd per chunk; idle lanes take consecutive items of the chunk, so rays that are neighbours in the (coherent) work
// order still run in the same warp.
#pragma once
#include "traverse.cuh"

#ifndef PT_CHUNK
#define PT_CHUNK 32u
#endif
#ifndef PT_REFILL_MIN
#define PT_REFILL_MIN 8 // fetch when at least this many lanes are idle (or all remaining lanes are idle)
#endif

// Src must provide:
//   __device__ bool load(uint32_t item, Ray& r, float& tmin, float& tmax, uint32_t& cullMask);   false: item is padding
//   __device__ void store(uint32_t item, const HitRec& h, bool anyHit);
// One candidate triangle of the lane's current ray (the acceptance rule of traverse<>). Returns true if an any-hit ray is finished.
template <bool ANY>
__device__ __forceinline__ bool testTriangle(const float4* __restrict__ tris, uint32_t triIndex, const Ray& r, float tmin, float& tbest, uint32_t cullMask, HitRec& hit) {
    const float4* tp = tris + size_t(triIndex) * 3;
    const float4 q2 = __ldg(tp + 2), q0 = __ldg(tp + 0), q1 = __ldg(tp + 1); // all three words at once: one memory round trip per triangle
    const uint32_t instW = __float_as_uint(q2.y), primW = __float_as_uint(q2.z);
    if (!((instW >> 24) & cullMask)) return false;
    float t, u, v, det;
    if (!intersectTri(q0, q1, q2, r, t, u, v, det)) return false;
    if (!(t > tmin)) return false;
    const uint32_t inst = instW & 0x00FFFFFFu, prim = primW & 0x7FFFFFFFu;
    const bool closer = t < tbest || (hit.found && t == tbest && (inst < hit.inst || (inst == hit.inst && prim < (hit.prim & 0x7FFFFFFFu))));
    if (!closer) return false;
    hit.found = true;
    if (ANY) return true;
    const bool back = (det > 0.0f) == ((primW & 0x80000000u) != 0u);
    tbest = t; hit.t = t; hit.inst = inst; hit.prim = prim | (back ? 0x80000000u : 0u); hit.u = u; hit.v = v;
    return false;
}

// Deferred-leaf variant (DEFER > 0). In persistentTrace<> the triangles of a node are tested right after the node, by the few
// lanes (3 of 32 on cfg2, profiles/r01) whose node step happened to reach a leaf. Here a lane that has found candidate triangles
// parks them (pBase, pBits) and waits; the warp keeps running node steps with the other lanes until at least DEFER lanes hold
// triangles (or as many lanes wait as can still step), then all waiting lanes test their triangles together. The closest accepted
// hit of a ray does not depend on the order of its triangle tests (equal t is resolved by (instance, primitive)), and the node a
// lane visits next is decided only after its parked triangles have been tested, so every ray executes exactly the per-ray
// operation sequence of traverse<>: results are bit-identical, only the interleaving of rays inside a warp changes.
// Measured on B200, cfg2 (profiles/r01c_defer_sweep.txt): any-hit rays gain (k_trace_shadow 0.400 -> 0.374 ms at DEFER = 12, the
// default); closest-hit rays lose (k_trace_primary 0.949 -> 0.978 ms at 16, 1.04 ms at 8): a parked lane does not step, and with
// 5-6 of 28 lanes receiving triangles per node step the node phase thins out faster than the triangle phase fills. A variant
// without the threshold (every iteration one node step or one parked triangle per lane) was slower still (1.22 ms) and is gone.
template <bool ANY, int DEFER, class Src>
__device__ __forceinline__ void persistentTraceDeferred(const uint4* __restrict__ nodes, const float4* __restrict__ tris, Src& src, uint32_t total, uint32_t* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    uint2 stack[VKX_STACK];
    int sp = 0;
    uint2 g = make_uint2(0u, 0u);
    Ray r; float tmin = 0.f, tbest = 0.f; uint32_t cullMask = 0, item = 0;
    HitRec hit; hit.found = false; hit.t = -1.0f; hit.u = hit.v = 0.f; hit.inst = hit.prim = 0xFFFFFFFFu;
    bool active = false;
    uint32_t pBase = 0, pBits = 0, pValid = 0; // parked triangles of this lane's ray (hit bits + the node's validity word)
    uint32_t cur = 0, end = 0;
    bool more = true;
    for (;;) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        const int nIdle = __popc(idle);
        if (nIdle == 32 && !more) break;
        if (more && (nIdle >= PT_REFILL_MIN)) {
            if (cur >= end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counter, PT_CHUNK);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                cur = base; end = min(base + PT_CHUNK, total);
                if (base >= total) { more = false; cur = end = 0; }
            }
            if (more) {
                const uint32_t avail = end - cur;
                const uint32_t rank = uint32_t(__popc(idle & ltMask));
                if (!active && rank < avail) {
                    const uint32_t it = cur + rank;
                    float tmax;
                    if (src.load(it, r, tmin, tmax, cullMask)) {
                        item = it; tbest = tmax; sp = 0; g = make_uint2(0u, VKX_ROOT_GROUP); pBits = 0;
                        hit.found = false; hit.t = -1.0f; hit.u = hit.v = 0.f; hit.inst = hit.prim = 0xFFFFFFFFu;
                        active = true;
                    }
                }
                cur += min(uint32_t(nIdle), avail);
            }
        }
        const bool parked = active && pBits != 0u;
        bool advance = false, done = false; // advance: this lane has no untested triangles left and must pick its next node
        const int nParked = __popc(__ballot_sync(0xFFFFFFFFu, parked));
        const int nStep = __popc(__ballot_sync(0xFFFFFFFFu, active && pBits == 0u));
        if (nParked >= DEFER || (nParked > 0 && nParked >= nStep)) { // triangle phase
            if (parked) {
                do {
                    const uint32_t b = uint32_t(__ffs(int(pBits))) - 1u;
                    pBits &= pBits - 1u;
                    if (testTriangle<ANY>(tris, pBase + triangleOffset(pValid, b), r, tmin, tbest, cullMask, hit)) { done = true; pBits = 0u; }
                } while (pBits);
                advance = true;
            }
        } else if (active && pBits == 0u) { // node phase (a lane without parked triangles always has an inner child to visit)
            const uint32_t slot = nextSlot(g.y >> 24, r.oct);
            g.y &= ~(0x01000000u << slot);
            if (g.y & 0xFF000000u) { if (sp < VKX_STACK) stack[sp++] = g; }
            const uint32_t rel = uint32_t(__popc(g.y & 0xFFu & ((1u << slot) - 1u)));
            uint4 w0, w1, w2, w3, w4;
            loadNode(nodes, g.x + rel, w0, w1, w2, w3, w4);
            const uint32_t m = intersectNode(w0, w1, w2, w3, w4, r, tmin, tbest);
            g.x = w1.x; g.y = (m & 0xFF000000u) | (w0.w >> 24);
            pBase = w1.y; pBits = m & 0x00FFFFFFu; pValid = w1.z;
            advance = pBits == 0u;
        }
        if (advance) {
            if (!done && !(g.y & 0xFF000000u)) {
                if (sp == 0) done = true;
                else g = stack[--sp];
            }
            if (done) { src.store(item, hit, ANY); active = false; }
        }
    }
}

template <bool ANY, class Src>
__device__ __forceinline__ void persistentTrace(const uint4* __restrict__ nodes, const float4* __restrict__ tris, Src& src, uint32_t total, uint32_t* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    uint2 stack[VKX_STACK];
    int sp = 0;
    uint2 g = make_uint2(0u, 0u);
    Ray r; float tmin = 0.f, tbest = 0.f; uint32_t cullMask = 0, item = 0;
    HitRec hit; hit.found = false; hit.t = -1.0f; hit.u = hit.v = 0.f; hit.inst = hit.prim = 0xFFFFFFFFu;
    bool active = false;
    uint32_t cur = 0, end = 0; // this warp's chunk (warp-uniform)
    bool more = true;
    for (;;) {
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        const int nIdle = __popc(idle);
        if (nIdle == 32 && !more) break;
        if (more && (nIdle >= PT_REFILL_MIN)) {
            if (cur >= end) { // take a new chunk
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counter, PT_CHUNK);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                cur = base; end = min(base + PT_CHUNK, total);
                if (base >= total) { more = false; cur = end = 0; }
            }
            if (more) {
                const uint32_t avail = end - cur;
                const uint32_t rank = uint32_t(__popc(idle & ltMask));
                if (!active && rank < avail) {
                    const uint32_t it = cur + rank;
                    float tmax;
                    if (src.load(it, r, tmin, tmax, cullMask)) {
                        item = it; tbest = tmax; sp = 0; g = make_uint2(0u, VKX_ROOT_GROUP);
                        hit.found = false; hit.t = -1.0f; hit.u = hit.v = 0.f; hit.inst = hit.prim = 0xFFFFFFFFu;
                        active = true;
                    }
                }
                cur += min(uint32_t(nIdle), avail);
            }
        }
        if (active) { // one node step + its triangles (same order as traverse<>)
            uint32_t triBase = 0, triBits = 0, triValid = 0;
            bool done = false;
            if (g.y & 0xFF000000u) {
                const uint32_t slot = nextSlot(g.y >> 24, r.oct);
                g.y &= ~(0x01000000u << slot);
                if (g.y & 0xFF000000u) { if (sp < VKX_STACK) stack[sp++] = g; }
                const uint32_t rel = uint32_t(__popc(g.y & 0xFFu & ((1u << slot) - 1u)));
                uint4 w0, w1, w2, w3, w4;
                loadNode(nodes, g.x + rel, w0, w1, w2, w3, w4);
                const uint32_t m = intersectNode(w0, w1, w2, w3, w4, r, tmin, tbest);
                g.x = w1.x; g.y = (m & 0xFF000000u) | (w0.w >> 24);
                triBase = w1.y; triBits = m & 0x00FFFFFFu; triValid = w1.z;
            }
            while (triBits) {
                const uint32_t b = uint32_t(__ffs(int(triBits))) - 1u;
                triBits &= triBits - 1u;
                const float4* tp = tris + size_t(triBase + triangleOffset(triValid, b)) * 3;
                const float4 q2 = __ldg(tp + 2); // (loading all three words before the mask test measured the same: 0.951 vs 0.944 ms)
                const uint32_t instW = __float_as_uint(q2.y), primW = __float_as_uint(q2.z);
                if (!((instW >> 24) & cullMask)) continue;
                const float4 q0 = __ldg(tp + 0), q1 = __ldg(tp + 1);
                float t, u, v, det;
                if (!intersectTri(q0, q1, q2, r, t, u, v, det)) continue;
                if (!(t > tmin)) continue;
                const uint32_t inst = instW & 0x00FFFFFFu, prim = primW & 0x7FFFFFFFu;
                const bool closer = t < tbest || (hit.found && t == tbest && (inst < hit.inst || (inst == hit.inst && prim < (hit.prim & 0x7FFFFFFFu))));
                if (!closer) continue;
                hit.found = true;
                if (ANY) { done = true; break; }
                const bool back = (det > 0.0f) == ((primW & 0x80000000u) != 0u);
                tbest = t; hit.t = t; hit.inst = inst; hit.prim = prim | (back ? 0x80000000u : 0u); hit.u = u; hit.v = v;
            }
            if (!done && !(g.y & 0xFF000000u)) {
                if (sp == 0) done = true;
                else g = stack[--sp];
            }
            if (done) { src.store(item, hit, ANY); active = false; }
        }
    }
}
