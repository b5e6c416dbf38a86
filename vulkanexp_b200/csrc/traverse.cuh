// Per-ray traversal of the 8-wide compressed BVH ("BVH spec v2"). Replaces the RT-core
// traversal behind traceRayEXT in the reference's shaders (reference src/shaders/traceProbes.rgen:43,
// closesthit.glsl:270-281, directLight.rgen:79-90, probesInit.rgen:45). The float operation sequence per ray is the
// one oracle/bvh.cpp executes; only the scheduling of rays onto lanes differs.
#pragma once
#include "common.cuh"
#include "texture.cuh"

#define VKX_STACK 48
#define VKX_ROOT_GROUP 0x01000000u // the root as "slot 0" of a virtual parent whose first child is node 0
// index (within the node's triangle range) of the triangle behind bit b of a hit mask: leaf triangles are contiguous in slot order
__device__ __forceinline__ uint32_t triangleOffset(uint32_t valid, uint32_t b) { return uint32_t(__popc(valid & 0x00FFFFFFu & ((1u << b) - 1u))); }

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz; // 1 / zero-fixed direction
    uint32_t oct;     // bit 2: dx < 0, bit 1: dy < 0, bit 0: dz < 0; bytes 1..3: the slot-preference masks of nextSlot() for this octant
};

__device__ __forceinline__ float fixZero(float d) { return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d; }

// Octant word: the 3 sign bits plus, per bit k of the octant, the slots whose (slot ^ octant) has bit k set (see nextSlot).
__device__ __forceinline__ uint32_t octantWord(uint32_t oct) { return oct | ((0xF0u >> (oct & 4u)) << 8) | ((0xCCu >> (oct & 2u)) << 16) | ((0xAAu >> (oct & 1u)) << 24); }

__device__ __forceinline__ Ray makeRay(float ox, float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    float fx = fixZero(dx), fy = fixZero(dy), fz = fixZero(dz);
    r.ix = __fdiv_rn(1.0f, fx); r.iy = __fdiv_rn(1.0f, fy); r.iz = __fdiv_rn(1.0f, fz);
    r.oct = octantWord((fx < 0.0f ? 4u : 0u) | (fy < 0.0f ? 2u : 0u) | (fz < 0.0f ? 1u : 0u));
    return r;
}

// The pending inner child to visit next: the slot s with the largest (s ^ octant). Children sit in the slot of their octant
// relative to the node centre, so this order is front to back along the ray (spec section 3.5; oracle/bvh.cpp::traverse runs the
// same rule as a loop). Three binary choices, most significant key bit first; `pending` is the 8-bit set of hit inner slots.
__device__ __forceinline__ uint32_t nextSlot(uint32_t pending, uint32_t octw) {
    uint32_t m = pending, t;
    t = m & (octw >> 8);  m = t ? t : m; // (the bytes above the selected one are harmless: m has only 8 bits)
    t = m & (octw >> 16); m = t ? t : m;
    t = m & (octw >> 24); m = t ? t : m;
    return uint32_t(__ffs(int(m))) - 1u;
}

// Two quantised plane bytes -> two floats q * 2^-24: PRMT places each byte in the low half of a binary16 (a subnormal, value
// q * 2^-24, exact) and the two conversions run on the FMA pipe (HADD2.F32). With the axis scale multiplied by 2^24 (exact) the
// plane distance fma(q * 2^-24, a * 2^24, b) is bit-identical to the spec's fma(float(q), a, b). Replaces 24 I2F.U8 (quarter-rate
// conversion pipe, 94 % busy in profiles/r01c) + 24 PRMT/FADD pairs by 24 PRMT + 48 HADD2.F32.
__device__ __forceinline__ void bytePairToFloats(uint32_t w, uint32_t sel, float& f0, float& f1) {
    const uint32_t pr = __byte_perm(w, 0u, sel);
    const __half2 h = *reinterpret_cast<const __half2*>(&pr);
    f0 = __low2float(h); f1 = __high2float(h);
}

// 8 quantised child boxes of one axis: near plane bytes (n0: slots 0-3, n1: slots 4-7) and far plane bytes.
struct AxisQ { uint32_t n0, n1, f0, f1; };

// Returns the hit mask of one node (BVH spec v2): bits 24..31 inner children by slot, bits 0..23 triangles (3 bits per leaf slot),
// already restricted to what the node holds (w1.z = the node's validity word).
__device__ __forceinline__ uint32_t intersectNode(const uint4 w0, const uint4 w1, const uint4 w2, const uint4 w3, const uint4 w4,
                                                   const Ray& r, float tmin, float tmax) {
    const float px = __uint_as_float(w0.x), py = __uint_as_float(w0.y), pz = __uint_as_float(w0.z);
    const uint32_t ew = w0.w;
    const float S = 16777216.0f; // 2^24, exact: a is rounded first like the spec's, then scaled
    const float ax = __fmul_rn(__fmul_rn(__uint_as_float((ew & 0xFFu) << 23), r.ix), S);
    const float ay = __fmul_rn(__fmul_rn(__uint_as_float(((ew >> 8) & 0xFFu) << 23), r.iy), S);
    const float az = __fmul_rn(__fmul_rn(__uint_as_float(((ew >> 16) & 0xFFu) << 23), r.iz), S);
    const float bx = __fmul_rn(__fsub_rn(px, r.ox), r.ix);
    const float by = __fmul_rn(__fsub_rn(py, r.oy), r.iy);
    const float bz = __fmul_rn(__fsub_rn(pz, r.oz), r.iz);
    // words: w2 = qlo.x[0..7] (x,y) qlo.y[0..7] (z,w); w3 = qlo.z, qhi.x; w4 = qhi.y, qhi.z
    AxisQ qx, qy, qz;
    if (r.oct & 4u) { qx.n0 = w3.z; qx.n1 = w3.w; qx.f0 = w2.x; qx.f1 = w2.y; } else { qx.n0 = w2.x; qx.n1 = w2.y; qx.f0 = w3.z; qx.f1 = w3.w; }
    if (r.oct & 2u) { qy.n0 = w4.x; qy.n1 = w4.y; qy.f0 = w2.z; qy.f1 = w2.w; } else { qy.n0 = w2.z; qy.n1 = w2.w; qy.f0 = w4.x; qy.f1 = w4.y; }
    if (r.oct & 1u) { qz.n0 = w4.z; qz.n1 = w4.w; qz.f0 = w3.x; qz.f1 = w3.y; } else { qz.n0 = w3.x; qz.n1 = w3.y; qz.f0 = w4.z; qz.f1 = w4.w; }
    uint32_t acc = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t nx = h ? qx.n1 : qx.n0, fx = h ? qx.f1 : qx.f0;
        const uint32_t ny = h ? qy.n1 : qy.n0, fy = h ? qy.f1 : qy.f0;
        const uint32_t nz = h ? qz.n1 : qz.n0, fz = h ? qz.f1 : qz.f0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t sel = j ? 0x4342u : 0x4140u; // bytes (2j, 2j+1) into the low bytes of the two halves
            float nxa, nxb, fxa, fxb, nya, nyb, fya, fyb, nza, nzb, fza, fzb;
            bytePairToFloats(nx, sel, nxa, nxb); bytePairToFloats(fx, sel, fxa, fxb);
            bytePairToFloats(ny, sel, nya, nyb); bytePairToFloats(fy, sel, fya, fyb);
            bytePairToFloats(nz, sel, nza, nzb); bytePairToFloats(fz, sel, fza, fzb);
            {
                const int s = 4 * h + 2 * j;
                const float tn = fmaxf(fmaxf(__fmaf_rn(nxa, ax, bx), __fmaf_rn(nya, ay, by)), fmaxf(__fmaf_rn(nza, az, bz), tmin));
                const float tf = fminf(fminf(__fmaf_rn(fxa, ax, bx), __fmaf_rn(fya, ay, by)), fminf(__fmaf_rn(fza, az, bz), tmax));
                if (tn <= tf) acc |= (7u << (3 * s)) | (1u << (24 + s));
            }
            {
                const int s = 4 * h + 2 * j + 1;
                const float tn = fmaxf(fmaxf(__fmaf_rn(nxb, ax, bx), __fmaf_rn(nyb, ay, by)), fmaxf(__fmaf_rn(nzb, az, bz), tmin));
                const float tf = fminf(fminf(__fmaf_rn(fxb, ax, bx), __fmaf_rn(fyb, ay, by)), fminf(__fmaf_rn(fzb, az, bz), tmax));
                if (tn <= tf) acc |= (7u << (3 * s)) | (1u << (24 + s));
            }
        }
    }
    return acc & w1.z;
}

// Moeller-Trumbore with the fixed operation order of oracle/bvh.cpp::intersectTri (spec v2.1: the barycentric tests compare the
// sign-adjusted numerators with |det|, the division is paid only by candidates that pass them).
__device__ __forceinline__ float flipBy(float x, float s) { return __uint_as_float(__float_as_uint(x) ^ (__float_as_uint(s) & 0x80000000u)); }
__device__ __forceinline__ bool intersectTri(const float4 q0, const float4 q1, const float4 q2, const Ray& r, float& t, float& u, float& v, float& det) {
    const float v0x = q0.x, v0y = q0.y, v0z = q0.z;
    const float e1x = q0.w, e1y = q1.x, e1z = q1.y;
    const float e2x = q1.z, e2y = q1.w, e2z = q2.x;
    const float px = __fmaf_rn(r.dy, e2z, -__fmul_rn(r.dz, e2y));
    const float py = __fmaf_rn(r.dz, e2x, -__fmul_rn(r.dx, e2z));
    const float pz = __fmaf_rn(r.dx, e2y, -__fmul_rn(r.dy, e2x));
    det = __fmaf_rn(e1x, px, __fmaf_rn(e1y, py, __fmul_rn(e1z, pz)));
    if (det == 0.0f) return false;
    const float ad = fabsf(det);
    const float tx = __fsub_rn(r.ox, v0x), ty = __fsub_rn(r.oy, v0y), tz = __fsub_rn(r.oz, v0z);
    const float un = __fmaf_rn(tx, px, __fmaf_rn(ty, py, __fmul_rn(tz, pz)));
    const float uns = flipBy(un, det);
    if (!(uns >= 0.0f && uns <= ad)) return false;
    const float qx = __fmaf_rn(ty, e1z, -__fmul_rn(tz, e1y));
    const float qy = __fmaf_rn(tz, e1x, -__fmul_rn(tx, e1z));
    const float qz = __fmaf_rn(tx, e1y, -__fmul_rn(ty, e1x));
    const float vn = __fmaf_rn(r.dx, qx, __fmaf_rn(r.dy, qy, __fmul_rn(r.dz, qz)));
    const float vns = flipBy(vn, det);
    if (!(vns >= 0.0f && __fadd_rn(uns, vns) <= ad)) return false;
    const float inv = __fdiv_rn(1.0f, det);
    u = __fmul_rn(un, inv);
    v = __fmul_rn(vn, inv);
    t = __fmul_rn(__fmaf_rn(e2x, qx, __fmaf_rn(e2y, qy, __fmul_rn(e2z, qz))), inv);
    return true;
}

struct HitRec {
    float t, u, v;
    uint32_t inst, prim; // prim bit 31: back face
    bool found;
};

__device__ __forceinline__ void loadNode(const uint4* __restrict__ nodes, uint32_t idx, uint4& w0, uint4& w1, uint4& w2, uint4& w3, uint4& w4) {
    const uint4* p = nodes + size_t(idx) * 5;
    w0 = __ldg(p + 0); w1 = __ldg(p + 1); w2 = __ldg(p + 2); w3 = __ldg(p + 3); w4 = __ldg(p + 4);
}

// One full traversal. ANY: terminate on first accepted hit, returns true if occluded.
// ALPHA: the pipeline's hit group has anyhit.rahit (direct light, reflection; not the probe pipelines): a candidate that would be
// accepted is first shown to the cut-out test and dropped if its albedo alpha is below 0.01 (needs alphaScene). The test runs only
// on would-be-accepted candidates, so the result is the closest / any non-ignored hit whatever the order of the triangle tests.
template <bool ANY, bool ALPHA = false>
__device__ __forceinline__ bool traverse(const uint4* __restrict__ nodes, const float4* __restrict__ tris, const Ray& r, float tmin, float tmax,
                                         uint32_t cullMask, HitRec& hit, const DeviceScene* alphaScene = nullptr) {
    float tbest = tmax;
    hit.found = false; hit.inst = 0xFFFFFFFFu; hit.prim = 0xFFFFFFFFu; hit.u = 0.f; hit.v = 0.f; hit.t = -1.0f;
    uint2 stack[VKX_STACK];
    int sp = 0;
    uint2 g = make_uint2(0u, VKX_ROOT_GROUP);
    for (;;) {
        uint32_t triBase = 0, triBits = 0, triValid = 0;
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
            const float4 q2 = __ldg(tp + 2);
            const uint32_t instW = __float_as_uint(q2.y), primW = __float_as_uint(q2.z);
            if (!((instW >> 24) & cullMask)) continue;
            const float4 q0 = __ldg(tp + 0), q1 = __ldg(tp + 1);
            float t, u, v, det;
            if (!intersectTri(q0, q1, q2, r, t, u, v, det)) continue;
            if (!(t > tmin)) continue;
            const uint32_t inst = instW & 0x00FFFFFFu, prim = primW & 0x7FFFFFFFu;
            const bool closer = t < tbest || (hit.found && t == tbest && (inst < hit.inst || (inst == hit.inst && prim < (hit.prim & 0x7FFFFFFFu))));
            if (!closer) continue;
            if (ALPHA) { if (anyHitIgnores(*alphaScene, inst, prim, u, v)) continue; }
            if (ANY) return true;
            const bool back = (det > 0.0f) == ((primW & 0x80000000u) != 0u);
            hit.found = true; tbest = t; hit.t = t; hit.inst = inst; hit.prim = prim | (back ? 0x80000000u : 0u); hit.u = u; hit.v = v;
        }
        if (!(g.y & 0xFF000000u)) {
            if (sp == 0) break;
            g = stack[--sp];
        }
    }
    return hit.found;
}
