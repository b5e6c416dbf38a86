// Persistent-thread traversal with warp-level ray replacement ("dynamic fetch"): warps stay resident, and a lane whose ray has
// terminated takes the next ray from the work list instead of idling until the slowest ray of its warp finishes. Measured
// motivation (profiles/r01a): with one ray per thread 11 of 32 lanes were active in the node test because traversal lengths
// inside a warp differ by 3x. The per-ray operation sequence is unchanged (it is the one of traverse<> in traverse.cuh);
// only the assignment of rays to lanes is dynamic.
//
// A finished ray is stored at once by its lane. Holding the record back until the lane takes its next ray, so that the >= PT_REFILL_MIN
// idle lanes store together, was measured and is slower (k_trace_primary 0.760 -> 0.781 ms, k_trace_shadow 0.291 -> 0.301 ms).
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

// Prefetching variant. The node a ray visits next is decided by the hit mask and the stack alone, never by the outcome of the
// triangle tests (those only shrink tbest, which the *next* node test reads). So the next child is popped right after the node
// test, its two cache lines are requested with prefetch.global.L1, and only then are the node's triangles tested: the node's
// memory round trip (34 % of the node fetches miss L1, profiles/r01d) overlaps the triangle phase instead of following it.
// No extra registers are live across the triangle loop (the r01 attempt loaded the next node into registers: 1.10 -> 1.32 ms).
// Same per-ray operation sequence as traverse<>: results are bit-identical.
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <bool ANY, class Src>
__device__ __forceinline__ void persistentTracePF(const uint4* __restrict__ nodes, const float4* __restrict__ tris, Src& src, uint32_t total, uint32_t* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    uint2 stack[VKX_STACK];
    int sp = 0;
    uint2 g = make_uint2(0u, 0u);
    Ray r; float tmin = 0.f, tbest = 0.f; uint32_t cullMask = 0, item = 0;
    HitRec hit; hit.found = false; hit.t = -1.0f; hit.u = hit.v = 0.f; hit.inst = hit.prim = 0xFFFFFFFFu;
    uint32_t next = 0xFFFFFFFFu; // node to test in the next step (already popped and prefetched); 0xFFFFFFFF: the ray is finished
    bool active = false;
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
                        item = it; tbest = tmax; sp = 0; g = make_uint2(0u, 0u); next = 0u; // the root
                        hit.found = false; hit.t = -1.0f; hit.u = hit.v = 0.f; hit.inst = hit.prim = 0xFFFFFFFFu;
                        active = true;
                    }
                }
                cur += min(uint32_t(nIdle), avail);
            }
        }
        if (active) {
            uint4 w0, w1, w2, w3, w4;
            loadNode(nodes, next, w0, w1, w2, w3, w4);
            const uint32_t m = intersectNode(w0, w1, w2, w3, w4, r, tmin, tbest);
            g.x = w1.x; g.y = (m & 0xFF000000u) | (w0.w >> 24);
            const uint32_t triBase = w1.y, triValid = w1.z;
            uint32_t triBits = m & 0x00FFFFFFu;
            // choose and prefetch the next node before the triangles
            if (!(g.y & 0xFF000000u)) {
                if (sp == 0) next = 0xFFFFFFFFu;
                else g = stack[--sp];
            }
            if (g.y & 0xFF000000u) {
                const uint32_t slot = nextSlot(g.y >> 24, r.oct);
                g.y &= ~(0x01000000u << slot);
                next = g.x + uint32_t(__popc(g.y & 0xFFu & ((1u << slot) - 1u)));
                if (g.y & 0xFF000000u) { if (sp < VKX_STACK) stack[sp++] = g; }
                const char* np = reinterpret_cast<const char*>(nodes + size_t(next) * 5);
                prefetchL1(np); prefetchL1(np + 64); // 80 bytes from a 16-byte aligned address: at most two 128-byte lines
            }
            bool done = false;
            while (triBits) {
                const uint32_t b = uint32_t(__ffs(int(triBits))) - 1u;
                triBits &= triBits - 1u;
                if (testTriangle<ANY>(tris, triBase + triangleOffset(triValid, b), r, tmin, tbest, cullMask, hit)) { done = true; break; }
            }
            if (next == 0xFFFFFFFFu) done = true;
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
                const float4 q2 = __ldg(tp + 2), q0 = __ldg(tp + 0), q1 = __ldg(tp + 1); // one round trip: the mask test no longer gates the other two words (r02f: the second wait alone was 8 % of the stall samples)
                const uint32_t instW = __float_as_uint(q2.y), primW = __float_as_uint(q2.z);
                if (!((instW >> 24) & cullMask)) continue;
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

// ---------------------------------------------------------------------------------------------------------------------------
// Ray-pool traversal: K rays per lane, their state in shared memory, one phase per pass.
//
// In persistentTrace<> a node step is followed by the triangle tests of the few lanes whose step reached a leaf: measured on cfg2
// (profiles/r02f_src_k_trace_primary.txt) the triangle loop runs 2.5 trips per node step at 3.0 of 32 lanes, a third of the issued
// instructions and 40 % of the stall samples (every trip pays a memory round trip for three lanes' worth of work). Here a lane owns
// K rays. Each pass the warp votes: a node pass (every lane advances one of its rays that wants a node step) or, once enough lanes
// hold a ray with untested triangles, a triangle pass (every such lane tests its ray's pending triangles). A ray that found
// triangles simply waits in shared memory while its lane steps another ray, so both passes run with (nearly) all lanes.
// The per-ray operation sequence is that of traverse<> (a ray's triangles are tested before its next node step, in slot order,
// and the next node is chosen by the same rule): results are bit-identical. A lane touches only its own shared-memory column
// (word index [field][slot][lane]): no bank conflicts, no synchronisation beyond the votes.
//
// PSrc provides:  bool begin(item, tag&, ox&, oy&, oz&, tmax&)   false: padding item
//                 void nodeRay(tag, ix&, iy&, iz&, octw&)         per node pass: reciprocal direction + octant word
//                 void triRay(tag, dx&, dy&, dz&)                 per triangle pass
//                 float tmin(); uint32_t cullMask();
//                 void finish(tag, found, t, u, v, inst, prim)
enum PoolField { PF_OX, PF_OY, PF_OZ, PF_TAG, PF_TBEST, PF_HU, PF_HV, PF_HINST, PF_HPRIM, PF_GX, PF_GY, PF_SP, PF_PBASE, PF_PBITS, PF_PVALID, PF_STACK };
template <int K, int STK> struct PoolLayout { static constexpr int fields = PF_STACK + 2 * STK; static constexpr int wordsPerWarp = fields * K * 32; };

template <bool ANY, int K, int STK, int TRI_THRESH, class PSrc>
__device__ __forceinline__ void poolTrace(const uint4* __restrict__ nodes, const float4* __restrict__ tris, PSrc& src, uint32_t total, uint32_t* __restrict__ counter, uint32_t* warpPool) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    uint32_t* const P = warpPool + lane;
    constexpr int FS = K * 32; // words between fields
    const float tmin = src.tmin();
    const uint32_t cullMask = src.cullMask();
    uint32_t states = 0; // 2 bits per slot: 0 empty, 1 wants a node step, 2 has untested triangles
    uint32_t cur = 0, end = 0;
    bool more = true;
    for (;;) {
        int kE = -1, kN = -1, kT = -1;
#pragma unroll
        for (int k = K - 1; k >= 0; --k) { const uint32_t s = (states >> (2 * k)) & 3u; if (s == 0u) kE = k; else if (s == 1u) kN = k; else kT = k; }
        const unsigned emptyLanes = __ballot_sync(0xFFFFFFFFu, kE >= 0);
        const unsigned busyLanes = __ballot_sync(0xFFFFFFFFu, states != 0u);
        if (!busyLanes && !more) break;
        if (more && (__popc(emptyLanes) >= PT_REFILL_MIN || !busyLanes)) { // ---- refill: one new ray per lane with a free slot
            if (cur >= end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counter, PT_CHUNK);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                cur = base; end = min(base + PT_CHUNK, total);
                if (base >= total) { more = false; cur = end = 0; }
            }
            if (more) {
                const uint32_t avail = end - cur;
                const uint32_t rank = uint32_t(__popc(emptyLanes & ltMask));
                if (kE >= 0 && rank < avail) {
                    uint32_t tag; float ox, oy, oz, tmax;
                    if (src.begin(cur + rank, tag, ox, oy, oz, tmax)) {
                        uint32_t* S = P + kE * 32;
                        S[PF_OX * FS] = __float_as_uint(ox); S[PF_OY * FS] = __float_as_uint(oy); S[PF_OZ * FS] = __float_as_uint(oz); S[PF_TAG * FS] = tag;
                        S[PF_TBEST * FS] = __float_as_uint(tmax); S[PF_HU * FS] = 0u; S[PF_HV * FS] = 0u; S[PF_HINST * FS] = 0xFFFFFFFFu; S[PF_HPRIM * FS] = 0xFFFFFFFFu;
                        S[PF_GX * FS] = 0u; S[PF_GY * FS] = VKX_ROOT_GROUP; S[PF_SP * FS] = 0u;
                        states |= 1u << (2 * kE);
                        if (kN < 0) kN = kE;
                    }
                }
                cur += min(uint32_t(__popc(emptyLanes)), avail);
            }
        }
        const int nT = __popc(__ballot_sync(0xFFFFFFFFu, kT >= 0));
        const int nN = __popc(__ballot_sync(0xFFFFFFFFu, kN >= 0));
        if (nT == 0 && nN == 0) continue;
        if (nT >= TRI_THRESH || nN == 0) { // ---- triangle pass
            if (kT >= 0) {
                uint32_t* S = P + kT * 32;
                Ray r;
                r.ox = __uint_as_float(S[PF_OX * FS]); r.oy = __uint_as_float(S[PF_OY * FS]); r.oz = __uint_as_float(S[PF_OZ * FS]);
                const uint32_t tag = S[PF_TAG * FS];
                src.triRay(tag, r.dx, r.dy, r.dz);
                float tbest = __uint_as_float(S[PF_TBEST * FS]);
                HitRec hit; hit.inst = S[PF_HINST * FS]; hit.prim = S[PF_HPRIM * FS]; hit.found = hit.inst != 0xFFFFFFFFu; hit.t = tbest;
                hit.u = __uint_as_float(S[PF_HU * FS]); hit.v = __uint_as_float(S[PF_HV * FS]);
                const uint32_t pBase = S[PF_PBASE * FS], pValid = S[PF_PVALID * FS];
                uint32_t pBits = S[PF_PBITS * FS];
                bool done = false;
                do {
                    const uint32_t b = uint32_t(__ffs(int(pBits))) - 1u;
                    pBits &= pBits - 1u;
                    if (testTriangle<ANY>(tris, pBase + triangleOffset(pValid, b), r, tmin, tbest, cullMask, hit)) { done = true; break; }
                } while (pBits);
                uint32_t next = 1u; // back to node steps
                if (!done) {
                    S[PF_TBEST * FS] = __float_as_uint(tbest); S[PF_HU * FS] = __float_as_uint(hit.u); S[PF_HV * FS] = __float_as_uint(hit.v); S[PF_HINST * FS] = hit.inst; S[PF_HPRIM * FS] = hit.prim;
                    if (!(S[PF_GY * FS] & 0xFF000000u) && S[PF_SP * FS] == 0u) done = true; // nothing left to visit
                }
                if (done) {
                    src.finish(tag, hit.found, hit.found ? tbest : -1.0f, hit.u, hit.v, hit.inst, hit.prim);
                    next = 0u;
                }
                states = (states & ~(3u << (2 * kT))) | (next << (2 * kT));
            }
        } else if (kN >= 0) { // ---- node pass
            uint32_t* S = P + kN * 32;
            Ray r;
            r.ox = __uint_as_float(S[PF_OX * FS]); r.oy = __uint_as_float(S[PF_OY * FS]); r.oz = __uint_as_float(S[PF_OZ * FS]);
            const uint32_t tag = S[PF_TAG * FS];
            src.nodeRay(tag, r.ix, r.iy, r.iz, r.oct);
            const float tbest = __uint_as_float(S[PF_TBEST * FS]);
            uint2 g = make_uint2(S[PF_GX * FS], S[PF_GY * FS]);
            uint32_t sp = S[PF_SP * FS];
            if (!(g.y & 0xFF000000u)) { --sp; g.x = S[(PF_STACK + 2 * sp) * FS]; g.y = S[(PF_STACK + 2 * sp + 1) * FS]; } // a ray in this state always has a group to pop
            const uint32_t slot = nextSlot(g.y >> 24, r.oct);
            g.y &= ~(0x01000000u << slot);
            if ((g.y & 0xFF000000u) && sp < uint32_t(STK)) { S[(PF_STACK + 2 * sp) * FS] = g.x; S[(PF_STACK + 2 * sp + 1) * FS] = g.y; ++sp; }
            const uint32_t rel = uint32_t(__popc(g.y & 0xFFu & ((1u << slot) - 1u)));
            uint4 w0, w1, w2, w3, w4;
            loadNode(nodes, g.x + rel, w0, w1, w2, w3, w4);
            const uint32_t m = intersectNode(w0, w1, w2, w3, w4, r, tmin, tbest);
            g.x = w1.x; g.y = (m & 0xFF000000u) | (w0.w >> 24);
            S[PF_GX * FS] = g.x; S[PF_GY * FS] = g.y; S[PF_SP * FS] = sp;
            const uint32_t pBits = m & 0x00FFFFFFu;
            uint32_t next = 1u;
            if (pBits) { S[PF_PBASE * FS] = w1.y; S[PF_PBITS * FS] = pBits; S[PF_PVALID * FS] = w1.z; next = 2u; }
            else if (!(g.y & 0xFF000000u) && sp == 0u) { // finished without pending triangles
                const uint32_t inst = S[PF_HINST * FS];
                const bool found = inst != 0xFFFFFFFFu;
                src.finish(tag, found, found ? tbest : -1.0f, __uint_as_float(S[PF_HU * FS]), __uint_as_float(S[PF_HV * FS]), inst, S[PF_HPRIM * FS]);
                next = 0u;
            }
            states = (states & ~(3u << (2 * kN))) | (next << (2 * kN));
        }
    }
}
