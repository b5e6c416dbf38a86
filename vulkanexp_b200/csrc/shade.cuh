// Device restatement of the reference's shading code for probe rays: sampleProbes (reference
// src/shaders/irradiance.glsl:145-237), sky (src/shaders/sky.glsl:59-126), pbrMetallicRoughness
// (src/shaders/pbrMetallicRoughness.glsl:43-84). Compiled with --fmad=false, so the arithmetic is the same
// sequence of IEEE operations the oracle executes; remaining differences come from libm (exp/pow/sqrt are IEEE,
// expf/powf are not bit-identical between glibc and CUDA).
#pragma once
#include "common.cuh"

struct v3 { float x, y, z; };
__device__ __forceinline__ v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 mk3(float s) { return mk3(s, s, s); }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ v3 operator*(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ v3 operator/(v3 a, v3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3 operator*(float s, v3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ v3 operator/(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ v3 operator+(v3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
__device__ __forceinline__ v3 operator-(float s, v3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
__device__ __forceinline__ float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ v3 cross3(v3 x, v3 y) { return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ float len3(v3 v) { return sqrtf(dot3(v, v)); }
__device__ __forceinline__ v3 norm3(v3 v) { return v * (1.0f / sqrtf(dot3(v, v))); }
__device__ __forceinline__ float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
__device__ __forceinline__ v3 mix3(v3 x, v3 y, float a) { return x * (1.0f - a) + y * a; }
__device__ __forceinline__ v3 mix3(v3 x, v3 y, v3 a) { return x * (1.0f - a) + y * a; }
__device__ __forceinline__ float maxS(float a, float b) { return a < b ? b : a; } // std::max
__device__ __forceinline__ float minS(float a, float b) { return b < a ? b : a; } // std::min
__device__ __forceinline__ float clampS(float x, float lo, float hi) { return minS(maxS(x, lo), hi); }
__device__ __forceinline__ float signS(float x) { return float((0.0f < x) - (x < 0.0f)); }
__device__ __forceinline__ v3 abs3(v3 v) { return mk3(fabsf(v.x), fabsf(v.y), fabsf(v.z)); }
__device__ __forceinline__ v3 reflect3(v3 I, v3 N) { return I - N * dot3(N, I) * 2.0f; }

// ---- exact helpers: one IEEE operation each, in the oracle's (= the shader's) order, whatever the translation unit's flags.
// The chain position -> normal / biased position -> octahedral coordinate -> atlas texel coordinate -> bilinear weights, and the
// depth moments that feed the Chebyshev test, are ill-conditioned: the texel coordinate is formed at the magnitude of the atlas
// (ulp(512 tiles) = 1e-3 depth texels), and variance = |mean^2 - mean2| cancels on flat walls, then enters the weight cubed and
// crushed (x9). Measured on cfg2 / cfg4 (profiles/r02_parity_flags.txt): with FMA contraction on this chain single rays differ from
// the oracle by 0.4 % ... 16 %; with it exact the largest difference is 3e-5. Everything smooth (colour filtering, weights, BRDF)
// keeps fused / approximate arithmetic.
__device__ __forceinline__ v3 xadd3(v3 a, v3 b) { return mk3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ v3 xsub3(v3 a, v3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ v3 xmul3(v3 a, float s) { return mk3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ float xdot3(v3 a, v3 b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }
__device__ __forceinline__ v3 xnorm3(v3 v) { return xmul3(v, __frcp_rn(__fsqrt_rn(xdot3(v, v)))); } // (rcp.rn = the correctly rounded 1 / x, the same value as div.rn(1, x)) // v * inversesqrt(dot(v, v))
__device__ __forceinline__ v3 xreflect3(v3 I, v3 N) { return xsub3(I, xmul3(xmul3(N, xdot3(N, I)), 2.0f)); } // I - N * dot(N, I) * 2

#define VKX_PI 3.1415926538f

// ---- grid helpers (irradiance.glsl:7-38)
__device__ __forceinline__ v3 gridCellSize(const vkx_grid_info& g) {
    return mk3(g.extentMax[0] - g.extentMin[0], g.extentMax[1] - g.extentMin[1], g.extentMax[2] - g.extentMin[2]) /
           mk3(float(g.resolution[0] - 1), float(g.resolution[1] - 1), float(g.resolution[2] - 1));
}
__device__ __forceinline__ v3 probeWorldPos(int ix, int iy, int iz, const vkx_grid_info& g) {
    return mk3(float(ix), float(iy), float(iz)) * gridCellSize(g) + mk3(g.extentMin[0], g.extentMin[1], g.extentMin[2]);
}
// Exact versions (explicit _rn intrinsics: immune to the translation unit's fmad / prec-div flags). Probe origins and hit
// positions feed traversal, whose results must stay bit-identical to the oracle.
__device__ __forceinline__ v3 gridCellSizeExact(const vkx_grid_info& g) {
    return mk3(__fdiv_rn(__fsub_rn(g.extentMax[0], g.extentMin[0]), float(g.resolution[0] - 1)), __fdiv_rn(__fsub_rn(g.extentMax[1], g.extentMin[1]), float(g.resolution[1] - 1)),
               __fdiv_rn(__fsub_rn(g.extentMax[2], g.extentMin[2]), float(g.resolution[2] - 1)));
}
__device__ __forceinline__ v3 probeWorldPosExact(int ix, int iy, int iz, const vkx_grid_info& g) {
    const v3 c = gridCellSizeExact(g);
    return mk3(__fadd_rn(__fmul_rn(float(ix), c.x), g.extentMin[0]), __fadd_rn(__fmul_rn(float(iy), c.y), g.extentMin[1]), __fadd_rn(__fmul_rn(float(iz), c.z), g.extentMin[2]));
}
__device__ __forceinline__ v3 pointOnRayExact(v3 origin, v3 direction, float t) { // direction * t + origin
    return mk3(__fadd_rn(__fmul_rn(direction.x, t), origin.x), __fadd_rn(__fmul_rn(direction.y, t), origin.y), __fadd_rn(__fmul_rn(direction.z, t), origin.z));
}
__device__ __forceinline__ void probeGridIndex(uint32_t index, const vkx_grid_info& g, int& ix, int& iy, int& iz) {
    uint32_t rx = uint32_t(g.resolution[0]), ry = uint32_t(g.resolution[1]);
    ix = int(index % rx); iy = int((index % (rx * ry)) / rx); iz = int(index / (rx * ry));
}

__device__ __forceinline__ float signNotZero(float k) { return (k >= 0.0f) ? 1.0f : -1.0f; }
__device__ __forceinline__ v3 octDecode(float ox, float oy) { // irradiance.glsl:107-112
    v3 v = mk3(ox, oy, 1.0f - fabsf(ox) - fabsf(oy));
    if (v.z < 0.0f) {
        float nx = (1.0f - fabsf(v.y)) * signNotZero(v.x);
        float ny = (1.0f - fabsf(v.x)) * signNotZero(v.y);
        v.x = nx; v.y = ny;
    }
    return norm3(v);
}
__device__ __forceinline__ float2 sphereToOctUV(v3 direction) { // irradiance.glsl:119-138
    v3 octant = mk3(signS(direction.x), signS(direction.y), signS(direction.z));
    float sum = dot3(direction, octant);
    v3 o = direction / sum;
    if (o.z < 0.0f) {
        v3 a = abs3(o);
        o.x = octant.x * (1.0f - a.y);
        o.y = octant.y * (1.0f - a.x);
    }
    return make_float2(o.x * 0.5f + 0.5f, o.y * 0.5f + 0.5f);
}

// REPEAT addressing without integer division or branches: atlas/noise coordinates produced on this path lie in [-size, 2*size)
// (uv in [-1, 2)), where one conditional add/subtract equals the oracle's ((i % size) + size) % size.
__device__ __forceinline__ void bilinearSetup(float u, uint32_t size, int& i0, int& i1, float& f) {
    float x = __fsub_rn(__fmul_rn(u, float(size)), 0.5f); // u * size - 0.5 with both roundings (exact chain, see above)
    float fl = floorf(x);
    f = __fsub_rn(x, fl);
    const int isz = int(size);
    int i = int(fl);
    i += (i < 0) ? isz : 0;
    i -= (i >= isz) ? isz : 0;
    i0 = i;
    i1 = (i0 + 1 == isz) ? 0 : i0 + 1;
}

__device__ __forceinline__ v3 sampleIrradianceTex(const DeviceProbes& p, float u, float v) {
    int x0, x1, y0, y1; float fx, fy;
    bilinearSetup(u, p.irrW, x0, x1, fx); bilinearSetup(v, p.irrH, y0, y1, fy);
    const uint32_t* r0 = p.irrSampled + size_t(y0) * p.irrW; const uint32_t* r1 = p.irrSampled + size_t(y1) * p.irrW;
    float3 t00 = unpackR11G11B10(__ldg(r0 + x0)), t10 = unpackR11G11B10(__ldg(r0 + x1)), t01 = unpackR11G11B10(__ldg(r1 + x0)), t11 = unpackR11G11B10(__ldg(r1 + x1));
    float gx = 1.0f - fx, gy = 1.0f - fy;
    v3 r;
    r.x = (t00.x * gx + t10.x * fx) * gy + (t01.x * gx + t11.x * fx) * fy;
    r.y = (t00.y * gx + t10.y * fx) * gy + (t01.y * gx + t11.y * fx) * fy;
    r.z = (t00.z * gx + t10.z * fx) * gy + (t01.z * gx + t11.z * fx) * fy;
    return r;
}
__device__ __forceinline__ float2 sampleDepthTex(const DeviceProbes& p, float u, float v) {
    int x0, x1, y0, y1; float fx, fy;
    bilinearSetup(u, p.depW, x0, x1, fx); bilinearSetup(v, p.depH, y0, y1, fy);
    const uint32_t* r0 = p.depSampled + size_t(y0) * p.depW; const uint32_t* r1 = p.depSampled + size_t(y1) * p.depW;
    float2 t00 = unpackRG16F(__ldg(r0 + x0)), t10 = unpackRG16F(__ldg(r0 + x1)), t01 = unpackRG16F(__ldg(r1 + x0)), t11 = unpackRG16F(__ldg(r1 + x1));
    float gx = 1.0f - fx, gy = 1.0f - fy;
    return make_float2((t00.x * gx + t10.x * fx) * gy + (t01.x * gx + t11.x * fx) * fy, (t00.y * gx + t10.y * fx) * gy + (t01.y * gx + t11.y * fx) * fy);
}

// Loop-invariant grid quantities, computed once per thread (same IEEE operations the per-call helpers perform).
struct GridConsts {
    v3 cell, acell, extentMin;
    float usx, usy, invUsx, invUsy; // uvScaling; inverses are only used when the scale is a power of two (exact)
    bool pow2x, pow2y;
    float cscale, dscale;
    int rx, ry, rz;
    float irrWf, irrHf, depWf, depHf; // atlas sizes as floats (8 / 16 texels per probe tile)
};
__device__ __forceinline__ bool isPow2f(float x) { return (__float_as_uint(x) & 0x007FFFFFu) == 0u && x > 0.0f; }
__device__ __forceinline__ GridConsts makeGridConsts(const vkx_grid_info& grid) {
    GridConsts c;
    c.cell = gridCellSizeExact(grid); c.acell = abs3(c.cell);
    c.extentMin = mk3(grid.extentMin[0], grid.extentMin[1], grid.extentMin[2]);
    c.usx = float(grid.resolution[0] * grid.resolution[1]); c.usy = float(grid.resolution[2]);
    c.pow2x = isPow2f(c.usx); c.pow2y = isPow2f(c.usy);
    c.invUsx = 1.0f / c.usx; c.invUsy = 1.0f / c.usy;
    c.cscale = float(grid.colorRes - 2) / float(grid.colorRes); c.dscale = float(grid.depthRes - 2) / float(grid.depthRes);
    c.rx = grid.resolution[0]; c.ry = grid.resolution[1]; c.rz = grid.resolution[2];
    c.irrWf = float(8 * c.rx * c.ry); c.irrHf = float(8 * c.rz); c.depWf = float(16 * c.rx * c.ry); c.depHf = float(16 * c.rz);
    return c;
}
// The same values computed on the host (every operation above is a single IEEE fp32 operation - subtraction, division, conversion -
// so x86-64 SSE arithmetic gives the same bits): passed to k_shade_front as a kernel parameter, the 24 values are read from the
// constant bank where they are used instead of occupying registers across the probe loop.
inline GridConsts makeGridConstsHost(const vkx_grid_info& grid) {
    GridConsts c;
    const volatile float ex = grid.extentMax[0] - grid.extentMin[0], ey = grid.extentMax[1] - grid.extentMin[1], ez = grid.extentMax[2] - grid.extentMin[2];
    c.cell.x = ex / float(grid.resolution[0] - 1); c.cell.y = ey / float(grid.resolution[1] - 1); c.cell.z = ez / float(grid.resolution[2] - 1);
    c.acell.x = c.cell.x < 0.0f ? -c.cell.x : c.cell.x; c.acell.y = c.cell.y < 0.0f ? -c.cell.y : c.cell.y; c.acell.z = c.cell.z < 0.0f ? -c.cell.z : c.cell.z;
    c.extentMin.x = grid.extentMin[0]; c.extentMin.y = grid.extentMin[1]; c.extentMin.z = grid.extentMin[2];
    c.usx = float(grid.resolution[0] * grid.resolution[1]); c.usy = float(grid.resolution[2]);
    auto pow2 = [](float x) { uint32_t b; memcpy(&b, &x, 4); return (b & 0x007FFFFFu) == 0u && x > 0.0f; };
    c.pow2x = pow2(c.usx); c.pow2y = pow2(c.usy);
    c.invUsx = 1.0f / c.usx; c.invUsy = 1.0f / c.usy;
    c.cscale = float(grid.colorRes - 2) / float(grid.colorRes); c.dscale = float(grid.depthRes - 2) / float(grid.depthRes);
    c.rx = grid.resolution[0]; c.ry = grid.resolution[1]; c.rz = grid.resolution[2];
    c.irrWf = float(8 * c.rx * c.ry); c.irrHf = float(8 * c.rz); c.depWf = float(16 * c.rx * c.ry); c.depHf = float(16 * c.rz);
    return c;
}
// x / scale, exactly: a division by a power of two equals the multiplication by its (exact) reciprocal.
__device__ __forceinline__ float divScale(float x, float scale, float inv, bool pow2) { return pow2 ? __fmul_rn(x, inv) : __fdiv_rn(x, scale); }
// (tileOrigin + 1) / res + localScale * oct, all over uvScaling (irradiance.glsl:171-175): the sum is formed at atlas magnitude, exact chain
__device__ __forceinline__ float atlasU(float base, float localScale, float oct, float scale, float inv, bool pow2) {
    return divScale(__fadd_rn(base, __fmul_rn(localScale, oct)), scale, inv, pow2);
}

// spherePointToOctohedralUV (irradiance.glsl:119-138), exact chain, without the z division: octahedron.z < 0 <=> direction.z < 0
// because the divisor dot(direction, sign(direction)) = |x| + |y| + |z| is positive (a quotient that underflows to -0 would need
// |z| < 2^-149). x and y are the shader's IEEE quotients.
__device__ __forceinline__ float2 sphereToOctUVxy(v3 direction) {
    const v3 octant = mk3(signS(direction.x), signS(direction.y), signS(direction.z));
    const float sum = xdot3(direction, octant);
    // direction.xy / sum: two IEEE quotients by the same divisor through one correctly rounded reciprocal r = RN(1 / sum) and
    // Markstein's correction (q0 = RN(n r), e = n - sum q0 exactly, q = RN(q0 + e r) = RN(n / sum); sum = |x| + |y| + |z| lies in
    // [1, sqrt 3] for the unit vectors passed here, |n| <= sum): 14 instructions instead of two ~15-instruction divisions with their
    // range checks.
    const float r = __frcp_rn(sum);
    const float qx = __fmul_rn(direction.x, r), qy = __fmul_rn(direction.y, r);
    float ox = __fmaf_rn(__fmaf_rn(-sum, qx, direction.x), r, qx), oy = __fmaf_rn(__fmaf_rn(-sum, qy, direction.y), r, qy);
    if (direction.z < 0.0f) {
        const float ax = fabsf(ox), ay = fabsf(oy);
        ox = __fmul_rn(octant.x, __fsub_rn(1.0f, ay));
        oy = __fmul_rn(octant.y, __fsub_rn(1.0f, ax));
    }
    return make_float2(__fadd_rn(__fmul_rn(ox, 0.5f), 0.5f), __fadd_rn(__fmul_rn(oy, 0.5f), 0.5f));
}

// Two sampleProbes calls of one closest hit (closesthit.glsl:241 with the reflected direction, :248 with the normal) share the
// position, hence the 8 probes, their positions, the direction to each probe, the trilinear weights and the state look-ups.
// This evaluates both in one pass; per call the arithmetic is the sequence of irradiance.glsl:145-237.
struct ProbeAccum { v3 finalColor, fallbackColor; float totalWeight, totalFallbackWeight; };

struct BilinearTaps { int o00, o10, o01, o11; float fx, fy; }; // texel offsets + weights of one bilinear fetch
__device__ __forceinline__ BilinearTaps makeTaps(float u, float v, uint32_t w, uint32_t h) {
    int x0, x1, y0, y1; BilinearTaps t;
    bilinearSetup(u, w, x0, x1, t.fx); bilinearSetup(v, h, y0, y1, t.fy);
    t.o00 = y0 * int(w) + x0; t.o10 = y0 * int(w) + x1; t.o01 = y1 * int(w) + x0; t.o11 = y1 * int(w) + x1;
    return t;
}
// Atlas look-ups of sampleProbes never wrap: the octahedral coordinate lies in [0, 1], so the unnormalised texel coordinate
// tile * res + 1 + (res - 2) * oct - 0.5 stays inside [tile * res + 0.5, tile * res + res - 1.5] (rounding cannot carry it to the
// next integer), i.e. x0 >= 0 and x0 + 1 <= size - 1: the REPEAT arithmetic of bilinearSetup is the identity there. The four texels
// are base, base + 1, base + w, base + w + 1; fx / fy are bilinearSetup's (same two roundings).
struct AtlasTaps { int base; float fx, fy; };
__device__ __forceinline__ AtlasTaps makeAtlasTaps(float u, float v, float wf, float hf, int w) {
    AtlasTaps t;
    const float x = __fsub_rn(__fmul_rn(u, wf), 0.5f), y = __fsub_rn(__fmul_rn(v, hf), 0.5f);
    const float flx = floorf(x), fly = floorf(y);
    t.fx = __fsub_rn(x, flx); t.fy = __fsub_rn(y, fly);
    t.base = int(fly) * w + int(flx);
    return t;
}
__device__ __forceinline__ float lerp2X(float t00, float t10, float t01, float t11, float gx, float fx, float gy, float fy) { // sampleDepth's arithmetic, op for op
    const float top = __fadd_rn(__fmul_rn(t00, gx), __fmul_rn(t10, fx)), bot = __fadd_rn(__fmul_rn(t01, gx), __fmul_rn(t11, fx));
    return __fadd_rn(__fmul_rn(top, gy), __fmul_rn(bot, fy));
}
template <class Taps>
__device__ __forceinline__ float2 lerpDepth(const Taps& t, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { // exact: the moments feed the variance
    const float2 t00 = unpackRG16F(a), t10 = unpackRG16F(b), t01 = unpackRG16F(c), t11 = unpackRG16F(d);
    const float gx = __fsub_rn(1.0f, t.fx), gy = __fsub_rn(1.0f, t.fy);
    return make_float2(lerp2X(t00.x, t10.x, t01.x, t11.x, gx, t.fx, gy, t.fy), lerp2X(t00.y, t10.y, t01.y, t11.y, gx, t.fx, gy, t.fy));
}
template <class Taps>
__device__ __forceinline__ v3 lerpIrradiance(const Taps& t, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const float3 t00 = unpackR11G11B10(a), t10 = unpackR11G11B10(b), t01 = unpackR11G11B10(c), t11 = unpackR11G11B10(d);
    const float gx = 1.0f - t.fx, gy = 1.0f - t.fy;
    return mk3((t00.x * gx + t10.x * t.fx) * gy + (t01.x * gx + t11.x * t.fx) * t.fy, (t00.y * gx + t10.y * t.fx) * gy + (t01.y * gx + t11.y * t.fx) * t.fy,
               (t00.z * gx + t10.z * t.fx) * gy + (t01.z * gx + t11.z * t.fx) * t.fy);
}
__device__ __forceinline__ void accumulateProbe(ProbeAccum& acc, v3 normal, v3 directionToProbe, float tri, float biasedDistToProbe, float2 depth, v3 color) {
    float weight = 1.0f;
    const float backfaceweight = maxS(0.0001f, (dot3(directionToProbe, normal) + 1.0f) * 0.5f);
    weight *= backfaceweight * backfaceweight + 0.2f;
    float fallbackWeight = weight;
    const float mean = depth.x;
    const float variance = fabsf(__fsub_rn(__fmul_rn(depth.x, depth.x), depth.y)); // two roundings like the shader: a contracted FMA changes this cancellation-prone difference by orders of magnitude on flat walls (mean^2 ~ mean2), and the Chebyshev ratio with it
    const float dd = maxS(biasedDistToProbe - mean, 0.0001f);
    float chebyshevWeight = variance / (variance + dd * dd);
    chebyshevWeight = maxS(chebyshevWeight * chebyshevWeight * chebyshevWeight, 0.0f); // pow(x, 3.0): within 2 ulp of powf
    weight *= (biasedDistToProbe <= mean) ? 1.0f : chebyshevWeight;
    weight = maxS(0.000001f, weight);
    const float crushThreshold = 0.2f;
    if (weight < crushThreshold) weight *= weight * weight * (1.0f / (crushThreshold * crushThreshold));
    weight *= tri;
    fallbackWeight *= tri;
    color = mk3(sqrtf(color.x), sqrtf(color.y), sqrtf(color.z));
    acc.finalColor = acc.finalColor + weight * color;
    acc.totalWeight += weight;
    acc.fallbackColor = acc.fallbackColor + fallbackWeight * color;
    acc.totalFallbackWeight += fallbackWeight;
}

// One probe, both sampleProbes calls: all 16 atlas texels (2 x (4 depth + 4 irradiance)) are requested before any is used, so
// the loads overlap instead of forming four dependent round trips to L2.
// One probe of one sampleProbes call (irradiance.glsl:158-222): both bilinear footprints are requested before either is used.
__device__ __forceinline__ void sampleProbeOne(const DeviceProbes& p, const GridConsts& gc, ProbeAccum& acc, v3 normal, float2 loc, v3 biased, v3 probePosition, v3 directionToProbe, float tri,
                                               float cu0, float cv0, float du0, float dv0) {
    // loc: (colorRes - 2) / colorRes * spherePointToOctohedralUV(normal) (the same for all eight probes of the call)
    const v3 b = xsub3(probePosition, biased);
    const float len = __fsqrt_rn(xdot3(b, b)); // biasedDistToProbe
    const float2 octD = sphereToOctUVxy(-xmul3(b, __frcp_rn(len))); // -normalize(biasedDirectionToProbe)
    const int irrW = int(p.irrW), depW = int(p.depW);
    const AtlasTaps tc = makeAtlasTaps(divScale(__fadd_rn(cu0, loc.x), gc.usx, gc.invUsx, gc.pow2x), divScale(__fadd_rn(cv0, loc.y), gc.usy, gc.invUsy, gc.pow2y), gc.irrWf, gc.irrHf, irrW);
    const AtlasTaps td = makeAtlasTaps(atlasU(du0, gc.dscale, octD.x, gc.usx, gc.invUsx, gc.pow2x), atlasU(dv0, gc.dscale, octD.y, gc.usy, gc.invUsy, gc.pow2y), gc.depWf, gc.depHf, depW);
    const uint32_t* D = p.depSampled + td.base; const uint32_t* C = p.irrSampled + tc.base;
    const uint32_t d0 = __ldg(D), d1 = __ldg(D + 1), d2 = __ldg(D + depW), d3 = __ldg(D + depW + 1);
    const uint32_t c0 = __ldg(C), c1 = __ldg(C + 1), c2 = __ldg(C + irrW), c3 = __ldg(C + irrW + 1);
    accumulateProbe(acc, normal, directionToProbe, tri, len, lerpDepth(td, d0, d1, d2, d3), lerpIrradiance(tc, c0, c1, c2, c3));
}
// One probe, both sampleProbes calls of a closest hit, one after the other. (Requesting all 16 texels of the pair before any use kept
// 16 more values live: 96 instead of 56 bytes of spills at 80 registers, and measured 0.008 ms slower for the shade phase.)
__device__ __forceinline__ void sampleProbePair(const DeviceProbes& p, const GridConsts& gc, ProbeAccum& accA, ProbeAccum& accB, v3 normalA, v3 normalB, float2 locA, float2 locB,
                                                v3 biasedA, v3 biasedB, v3 probePosition, v3 directionToProbe, float tri, int tile, int cz) {
    // (tileOrigin + 1) / res = tile + 1 / res: exact in fp32 (tile counts are far below 2^20)
    const float tileF = float(tile), czF = float(cz);
    const float cu0 = __fadd_rn(tileF, 0.125f), cv0 = __fadd_rn(czF, 0.125f), du0 = __fadd_rn(tileF, 0.0625f), dv0 = __fadd_rn(czF, 0.0625f);
    sampleProbeOne(p, gc, accA, normalA, locA, biasedA, probePosition, directionToProbe, tri, cu0, cv0, du0, dv0);
    asm volatile("" ::: "memory"); // keeps the compiler from interleaving the two calls again
    sampleProbeOne(p, gc, accB, normalB, locB, biasedB, probePosition, directionToProbe, tri, cu0, cv0, du0, dv0);
}
__device__ __forceinline__ v3 finishProbes(ProbeAccum a) {
    if (a.totalWeight > 1e-3f) a.finalColor = a.finalColor * (1.0f / a.totalWeight);
    if (a.totalFallbackWeight > 1e-3f) a.fallbackColor = a.fallbackColor * (1.0f / a.totalFallbackWeight);
    a.finalColor = a.finalColor * a.finalColor;
    a.fallbackColor = a.fallbackColor * a.fallbackColor;
    return mix3(a.fallbackColor, a.finalColor, 8.0f * clampS(a.totalWeight, 0.0f, 1.0f / 8.0f));
}

// resultA = sampleProbes(position, normalA, toCamera), resultB = sampleProbes(position, normalB, toCamera)
__device__ inline void sampleProbes2(const DeviceProbes& p, const GridConsts& gc, v3 position, v3 normalA, v3 normalB, v3 toCamera, v3& resultA, v3& resultB) {
    const vkx_grid_info& grid = p.grid;
    // exact quotient: int(gridCoords) selects the 8 probes
    const v3 gridCoords = mk3(__fdiv_rn(__fsub_rn(position.x, gc.extentMin.x), gc.acell.x), __fdiv_rn(__fsub_rn(position.y, gc.extentMin.y), gc.acell.y), __fdiv_rn(__fsub_rn(position.z, gc.extentMin.z), gc.acell.z));
    resultA = mk3(0.0f); resultB = mk3(0.0f);
    if (gridCoords.x < 0.0f || gridCoords.y < 0.0f || gridCoords.z < 0.0f) return;
    const v3 biasedA = xadd3(position, xmul3(xadd3(normalA, toCamera), grid.shadowBias));
    const v3 biasedB = xadd3(position, xmul3(xadd3(normalB, toCamera), grid.shadowBias));
    const int fx = int(gridCoords.x), fy = int(gridCoords.y), fz = int(gridCoords.z);
    v3 alpha = (position - (mk3(float(fx), float(fy), float(fz)) * gc.cell + gc.extentMin)) / gc.acell;
    alpha = mk3(clampS(alpha.x, 0.0f, 1.0f), clampS(alpha.y, 0.0f, 1.0f), clampS(alpha.z, 0.0f, 1.0f));
    const float2 octA = sphereToOctUVxy(normalA), octB = sphereToOctUVxy(normalB);
    const float2 locA = make_float2(__fmul_rn(gc.cscale, octA.x), __fmul_rn(gc.cscale, octA.y)), locB = make_float2(__fmul_rn(gc.cscale, octB.x), __fmul_rn(gc.cscale, octB.y));
    ProbeAccum accA, accB;
    accA.finalColor = accA.fallbackColor = mk3(0.0f); accA.totalWeight = accA.totalFallbackWeight = 0.0f;
    accB = accA;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        const int cx = fx + ox, cy = fy + oy, cz = fz + oz;
        if (cx > gc.rx - 1 || cy > gc.ry - 1 || cz > gc.rz - 1) continue;
        const uint32_t li = uint32_t(cx + gc.rx * cy + gc.rx * gc.ry * cz);
        if (__ldg(p.stateSampled + li) == 0u) continue;
        const v3 probePosition = mk3(__fadd_rn(__fmul_rn(float(cx), gc.cell.x), gc.extentMin.x), __fadd_rn(__fmul_rn(float(cy), gc.cell.y), gc.extentMin.y), __fadd_rn(__fmul_rn(float(cz), gc.cell.z), gc.extentMin.z));
        const v3 directionToProbe = norm3(probePosition - position);
        const v3 trilinear = mix3(1.0f - alpha, alpha, mk3(float(ox), float(oy), float(oz)));
        const float tri = trilinear.x * trilinear.y * trilinear.z + 0.001f;
        const int tile = cy * gc.rx + cx;
        sampleProbePair(p, gc, accA, accB, normalA, normalB, locA, locB, biasedA, biasedB, probePosition, directionToProbe, tri, tile, cz);
    }
    resultA = finishProbes(accA);
    resultB = finishProbes(accB);
}

// Single look-up (FinalGather.frag:68): same per-probe sequence as one half of sampleProbePair.
__device__ inline v3 sampleProbes1(const DeviceProbes& p, const GridConsts& gc, v3 position, v3 normal, v3 toCamera) {
    const vkx_grid_info& grid = p.grid;
    const v3 gridCoords = mk3(__fdiv_rn(__fsub_rn(position.x, gc.extentMin.x), gc.acell.x), __fdiv_rn(__fsub_rn(position.y, gc.extentMin.y), gc.acell.y), __fdiv_rn(__fsub_rn(position.z, gc.extentMin.z), gc.acell.z));
    if (gridCoords.x < 0.0f || gridCoords.y < 0.0f || gridCoords.z < 0.0f) return mk3(0.0f);
    const v3 biased = xadd3(position, xmul3(xadd3(normal, toCamera), grid.shadowBias));
    const int fx = int(gridCoords.x), fy = int(gridCoords.y), fz = int(gridCoords.z);
    v3 alpha = (position - (mk3(float(fx), float(fy), float(fz)) * gc.cell + gc.extentMin)) / gc.acell;
    alpha = mk3(clampS(alpha.x, 0.0f, 1.0f), clampS(alpha.y, 0.0f, 1.0f), clampS(alpha.z, 0.0f, 1.0f));
    const float2 oct = sphereToOctUVxy(normal);
    ProbeAccum acc;
    acc.finalColor = acc.fallbackColor = mk3(0.0f); acc.totalWeight = acc.totalFallbackWeight = 0.0f;
    const uint32_t* D = p.depSampled; const uint32_t* C = p.irrSampled;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        const int cx = fx + ox, cy = fy + oy, cz = fz + oz;
        if (cx > gc.rx - 1 || cy > gc.ry - 1 || cz > gc.rz - 1) continue;
        const uint32_t li = uint32_t(cx + gc.rx * cy + gc.rx * gc.ry * cz);
        if (__ldg(p.stateSampled + li) == 0u) continue;
        const v3 probePosition = mk3(__fadd_rn(__fmul_rn(float(cx), gc.cell.x), gc.extentMin.x), __fadd_rn(__fmul_rn(float(cy), gc.cell.y), gc.extentMin.y), __fadd_rn(__fmul_rn(float(cz), gc.cell.z), gc.extentMin.z));
        const v3 directionToProbe = norm3(probePosition - position);
        const v3 trilinear = mix3(1.0f - alpha, alpha, mk3(float(ox), float(oy), float(oz)));
        const float tri = trilinear.x * trilinear.y * trilinear.z + 0.001f;
        const int tile = cy * gc.rx + cx;
        const v3 b = xsub3(probePosition, biased);
        const float len = __fsqrt_rn(xdot3(b, b));
        const float2 octD = sphereToOctUVxy(-xmul3(b, __frcp_rn(len)));
        const float cu0 = float(8 * tile + 1) * 0.125f, cv0 = float(8 * cz + 1) * 0.125f, du0 = float(16 * tile + 1) * 0.0625f, dv0 = float(16 * cz + 1) * 0.0625f;
        const BilinearTaps tc = makeTaps(atlasU(cu0, gc.cscale, oct.x, gc.usx, gc.invUsx, gc.pow2x), atlasU(cv0, gc.cscale, oct.y, gc.usy, gc.invUsy, gc.pow2y), p.irrW, p.irrH);
        const BilinearTaps td = makeTaps(atlasU(du0, gc.dscale, octD.x, gc.usx, gc.invUsx, gc.pow2x), atlasU(dv0, gc.dscale, octD.y, gc.usy, gc.invUsy, gc.pow2y), p.depW, p.depH);
        const uint32_t d0 = __ldg(D + td.o00), d1 = __ldg(D + td.o10), d2 = __ldg(D + td.o01), d3 = __ldg(D + td.o11);
        const uint32_t c0 = __ldg(C + tc.o00), c1 = __ldg(C + tc.o10), c2 = __ldg(C + tc.o01), c3 = __ldg(C + tc.o11);
        accumulateProbe(acc, normal, directionToProbe, tri, len, lerpDepth(td, d0, d1, d2, d3), lerpIrradiance(tc, c0, c1, c2, c3));
    }
    return finishProbes(acc);
}

// ---- sky.glsl. The 64-step integral dominates missed rays; inside the loop this uses explicit FMAs, ex2.approx and
// approximate reciprocal/rsqrt (the function is smooth: measured deviation from the oracle ~1e-6 relative, tolerance 1e-3).
__device__ __forceinline__ float fastExp(float x) { return __expf(x); }
__device__ __forceinline__ float fastRsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float skyScaleFast(float fCos) {
    const float x = 1.0f - fCos;
    return 0.25f * fastExp(fmaf(x, fmaf(x, fmaf(x, fmaf(x, 5.25f, -6.80f), 3.83f), 0.459f), -0.00287f));
}
__device__ __forceinline__ float skyScale(float fCos) {
    float x = 1.0f - fCos;
    return 0.25f * expf(-0.00287f + x * (0.459f + x * (3.83f + x * (-6.80f + x * 5.25f))));
}
__device__ inline float traceSphereOutside(v3 center, float radius, v3 origin, v3 direction) {
    v3 d = origin - center;
    float a = dot3(direction, direction), b = dot3(direction, d), c = dot3(d, d) - radius * radius;
    float g = b * b - a * c;
    if (g > 0.0f) { float dis = (-sqrtf(g) - b) / a; if (dis > 0.0f) return dis; }
    return -1.0f;
}
__device__ inline float traceSphereInside(v3 center, float radius, v3 origin, v3 direction) {
    v3 oc = center - origin;
    float docdir = dot3(oc, direction);
    v3 pc = origin + docdir * direction;
    float dist = sqrtf(radius * radius - len3(pc - center) * len3(pc - center));
    if (docdir > 0.0f) return dist - len3(pc - origin);
    else return dist + len3(pc - origin);
}

__device__ inline v3 skyColor(v3 rayOrigin, v3 rayDirection, v3 sunPosition, v3 sunColor, float sunBrightnessFactor) { // sky.glsl:59-126, showSun = true
    const float AvegerageDensityAltitude = 0.25f, InnerRadius = 100000.0f, OuterRadius = 2500.0f + InnerRadius;
    const float Scale = 1.0f / (OuterRadius - InnerRadius);
    const float Kr = 0.0025f, Kr4PI = Kr * 4.0f * VKX_PI, Km = 0.0010f, Km4PI = Km * 4.0f * VKX_PI, g = -0.990f;
    // 1 / pow(w, 4) for w = 0.650, 0.570, 0.475 as glibc's powf evaluates them in fp32
    const v3 InvWaveLengths = mk3(__uint_as_float(0x40b343f5u), __uint_as_float(0x41179293u), __uint_as_float(0x419d2682u)); // 5.60204554, 9.47328472, 19.6438026
    sunColor = sunColor * sunBrightnessFactor;
    const v3 planetCenter = mk3(0.0f, -InnerRadius - 100.0f, 0.0f);
    v3 position = rayOrigin - planetCenter;
    float height = len3(position);
    const v3 lightDir = norm3(sunPosition);
    if (fabsf(height - InnerRadius) < 1e-3f) { position = position + 1e-2f * norm3(position); height = len3(position); }
    if (height < OuterRadius) {
        if (height > InnerRadius) {
            float planetDistance = traceSphereOutside(mk3(0.0f), InnerRadius, position, rayDirection);
            if (planetDistance >= 0.0f) return maxS(0.1f, dot3(lightDir, norm3(position + planetDistance * rayDirection))) * mk3(0.05f);
        } else return mk3(0.0f);
        const float rayDepth = traceSphereInside(mk3(0.0f), OuterRadius, position, rayDirection);
        if (isinf(rayDepth) || isnan(rayDepth)) return mk3(0.0f);
        const float kDepth = Scale / AvegerageDensityAltitude;
        const float depth = expf(kDepth * (InnerRadius - height));
        const float startAngle = dot3(rayDirection, position) / height;
        const float startOffset = depth * skyScale(startAngle);
        const float sampleLength = rayDepth / 64.0f;
        const float scaledLength = sampleLength * Scale;
        const v3 sampleRay = rayDirection * sampleLength;
        v3 samplePoint = position + 0.5f * sampleRay;
        v3 color = mk3(0.0f);
        const v3 kk = InvWaveLengths * Kr4PI + Km4PI;
#pragma unroll 2
        for (uint32_t i = 0; i < 64u; ++i) {
            const float h2 = fmaf(samplePoint.x, samplePoint.x, fmaf(samplePoint.y, samplePoint.y, samplePoint.z * samplePoint.z));
            const float ih = fastRsqrt(h2);
            const float h = h2 * ih;
            const float dpt = fastExp(kDepth * (InnerRadius - h));
            const float lightAngle = fmaf(lightDir.x, samplePoint.x, fmaf(lightDir.y, samplePoint.y, lightDir.z * samplePoint.z)) * ih;
            const float cameraAngle = fmaf(rayDirection.x, samplePoint.x, fmaf(rayDirection.y, samplePoint.y, rayDirection.z * samplePoint.z)) * ih;
            const float scatter = fmaf(dpt, skyScaleFast(lightAngle) - skyScaleFast(cameraAngle), startOffset);
            const v3 attenuate = mk3(fastExp(-scatter * kk.x), fastExp(-scatter * kk.y), fastExp(-scatter * kk.z));
            // sky.glsl:103 `if(any(isinf(attenuate)) || any(isnan(attenuate))) continue;` - the components are exponentials (>= 0 or NaN), so
            // their sum is finite exactly when all three are: one addition chain and one comparison instead of six classifications
            if (!((attenuate.x + attenuate.y) + attenuate.z < __uint_as_float(0x7F800000u))) continue;
            const float s = dpt * scaledLength;
            color = mk3(fmaf(attenuate.x, s, color.x), fmaf(attenuate.y, s, color.y), fmaf(attenuate.z, s, color.z));
            samplePoint = samplePoint + sampleRay;
        }
        const v3 secondary = color * Km * sunColor;
        color = color * (InvWaveLengths * Kr * sunColor);
        {
            const float miecos = dot3(lightDir, -rayDirection);
            const float miePhase = 1.5f * ((1.0f - g * g) / (2.0f + g * g)) * (1.0f + miecos * miecos) / powf(maxS(1e-3f, 1.0f + g * g - 2.0f * g * miecos), 1.5f);
            color = color + miePhase * secondary;
        }
        if (!(isinf(color.x) || isinf(color.y) || isinf(color.z))) return color;
    } else {
        float depth = traceSphereOutside(mk3(0.0f), OuterRadius, position, rayDirection);
        if (depth > 0.0f) return dot3(lightDir, norm3(position + depth * rayDirection)) * 0.5f * mk3(0.5294117647f, 0.80784313725f, 0.92156862745f);
    }
    return mk3(0.0f);
}

// ---- pbrMetallicRoughness.glsl:43-84
__device__ inline v3 pbrMetallicRoughness(v3 normal, v3 view, v3 lightColor, v3 lightDirection, v3 albedo, float metalness, float roughness) {
    const v3 f0 = mk3(0.04f);
    v3 diffuseColor = albedo * (1.0f - f0);
    diffuseColor = diffuseColor * (1.0f - metalness);
    const float alphaRoughness = roughness * roughness;
    const v3 specularColor = mix3(f0, albedo, metalness);
    const float reflectance = maxS(maxS(specularColor.x, specularColor.y), specularColor.z);
    const float reflectance90 = clampS(reflectance * 25.0f, 0.0f, 1.0f);
    const v3 R0 = specularColor, R90 = mk3(1.0f) * reflectance90;
    const v3 n = normal, v = view;
    const v3 l = norm3(lightDirection);
    const v3 h = norm3(l + v);
    const float NdotL = clampS(dot3(n, l), 0.001f, 1.0f);
    const float NdotV = clampS(fabsf(dot3(n, v)), 0.001f, 1.0f);
    const float NdotH = clampS(dot3(n, h), 0.0f, 1.0f);
    const float VdotH = clampS(dot3(v, h), 0.0f, 1.0f);
    const v3 F = R0 + (R90 - R0) * powf(clampS(1.0f - VdotH, 0.0f, 1.0f), 5.0f);
    const float ar2 = alphaRoughness * alphaRoughness;
    const float attenuationL = 2.0f * NdotL / (NdotL + sqrtf(ar2 + (1.0f - ar2) * (NdotL * NdotL)));
    const float attenuationV = 2.0f * NdotV / (NdotV + sqrtf(ar2 + (1.0f - ar2) * (NdotV * NdotV)));
    const float G = attenuationL * attenuationV;
    const float a = NdotH * alphaRoughness;
    const float k = alphaRoughness / ((1.0f - NdotH * NdotH) + a * a);
    const float D = clampS(k * k * (1.0f / VKX_PI), 0.0f, 4.0f);
    const v3 diffuseContrib = (1.0f - F) * diffuseColor / VKX_PI;
    const v3 specContrib = F * G * D / (4.0f * NdotL * NdotV);
    return NdotL * lightColor * (diffuseContrib + specContrib);
}
