// 1-spp ray-traced sun shadows with depth-aware separable Gaussian filter and reprojected temporal accumulation.
// Replaces (reference): src/shaders/directLight.rgen:42-99 (+ shadow.rmiss, anyhit.rahit for untextured scenes),
// src/shaders/directLightFilter.glsl:53-142 (X and Y variants), dispatch order of src/SwapchainManagement.cpp:409-438 and
// the history copy of src/Editor.cpp:287-316 (here a ping-pong, no copy). Out-of-bounds image loads return 0 and
// out-of-bounds stores are dropped.
#include "common.cuh"
#include "traverse.cuh"
#include "shade.cuh"

namespace {

struct M4 { float m[16]; }; // column-major

__device__ __forceinline__ float4 mulM4(const M4& M, float x, float y, float z, float w) { // glm: (m0*x + m1*y) + (m2*z + m3*w)
    float4 r;
    r.x = (M.m[0] * x + M.m[4] * y) + (M.m[8] * z + M.m[12] * w);
    r.y = (M.m[1] * x + M.m[5] * y) + (M.m[9] * z + M.m[13] * w);
    r.z = (M.m[2] * x + M.m[6] * y) + (M.m[10] * z + M.m[14] * w);
    r.w = (M.m[3] * x + M.m[7] * y) + (M.m[11] * z + M.m[15] * w);
    return r;
}

// G-buffer fixture: primary rays set up as raygen.rgen:27-33, outputs laid out as GBuffer.frag:64-68.
// ALPHA: the scene has a texture list; cut-outs follow anyhit.rahit (the raster pass discards below 0.05, GBuffer.frag:38: this is a fixture).
template <bool ALPHA>
__global__ void __launch_bounds__(128) k_gbuffer(DeviceScene sc, M4 invView, M4 invProj, float3 camOrigin, uint32_t W, uint32_t H,
                                                 float4* __restrict__ posDepth, float4* __restrict__ normalMetal, float4* __restrict__ albedoRough, float4* __restrict__ emissive) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const size_t pix = size_t(y) * W + x;
    const float4 o4 = mulM4(invView, 0.f, 0.f, 0.f, 1.f);
    const float ux = (float(x) + 0.5f) / float(W), uy = (float(y) + 0.5f) / float(H);
    const float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
    const float4 target = mulM4(invProj, dx, dy, 1.f, 1.f);
    const v3 tn = norm3(mk3(target.x, target.y, target.z));
    const float4 d4 = mulM4(invView, tn.x, tn.y, tn.z, 0.f);
    const Ray r = makeRay(o4.x, o4.y, o4.z, d4.x, d4.y, d4.z);
    HitRec h;
    if (!traverse<false, ALPHA>(sc.nodes, sc.tris, r, 0.001f, 100000.0f, 0xFFu, h, &sc)) {
        posDepth[pix] = make_float4(0.f, 0.f, 0.f, 0.f); normalMetal[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        albedoRough[pix] = make_float4(0.f, 0.f, 0.f, 0.f); emissive[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const v3 origin = mk3(o4.x, o4.y, o4.z), dir = mk3(d4.x, d4.y, d4.z);
    const v3 position = dir * h.t + origin;
    const uint32_t meshEntry = sc.instances[h.inst].meshEntry;
    const vkx_offset_entry oe = sc.offsets[meshEntry];
    const uint32_t prim = h.prim & 0x7FFFFFFFu;
    v3 n[3], col[3];
    for (int c = 0; c < 3; ++c) {
        const vkx_vertex& vx = sc.vertices[oe.vertexOffset + sc.indices[oe.indexOffset + 3 * prim + c]];
        n[c] = mk3(vx.normal[0], vx.normal[1], vx.normal[2]); col[c] = mk3(vx.color[0], vx.color[1], vx.color[2]);
    }
    const v3 on = norm3(n[0] * (1.0f - h.u - h.v) + n[1] * h.u + n[2] * h.v);
    const v3 vcolor = col[0] * (1.0f - h.u - h.v) + col[1] * h.u + col[2] * h.v; // the `color` varying of GBuffer.vert
    const float* Wm = sc.worldToObject + size_t(h.inst) * 9;
    const v3 normal = norm3(mk3(dot3(on, mk3(Wm[0], Wm[3], Wm[6])), dot3(on, mk3(Wm[1], Wm[4], Wm[7])), dot3(on, mk3(Wm[2], Wm[5], Wm[8]))));
    posDepth[pix] = make_float4(position.x, position.y, position.z, len3(position - mk3(camOrigin.x, camOrigin.y, camOrigin.z)));
    const vkx_material& mat = sc.materials[oe.materialIndex];
    normalMetal[pix] = make_float4(normal.x, normal.y, normal.z, mat.metallicFactor);
    albedoRough[pix] = make_float4(vcolor.x * mat.baseColorFactor[0], vcolor.y * mat.baseColorFactor[1], vcolor.z * mat.baseColorFactor[2], mat.roughnessFactor); // GBuffer.frag:35,66
    emissive[pix] = make_float4(mat.emissiveFactor[0], mat.emissiveFactor[1], mat.emissiveFactor[2], 1.0f);                                                       // GBuffer.frag:59,67
}

__device__ __forceinline__ v3 rotateAxis(v3 p, v3 axis, float angle) { // common.glsl:6-8
    float sn, cs; sincosf(angle, &sn, &cs);
    return mix3(dot3(axis, p) * axis, p, cs) + cross3(axis, p) * sn;
}

// general REPEAT wrap (uv = pixel / 64 is far outside [0, 1), unlike the atlas coordinates bilinearSetup() handles)
__device__ __forceinline__ void bilinearSetupRepeat(float u, uint32_t size, int& i0, int& i1, float& f) {
    const float x = u * float(size) - 0.5f;
    const float fl = floorf(x);
    f = x - fl;
    const int isz = int(size), i = int(fl);
    if ((size & (size - 1u)) == 0u) { i0 = i & (isz - 1); i1 = (i0 + 1) & (isz - 1); } // power-of-two noise tile (the reference's is 64 x 64)
    else { i0 = ((i % isz) + isz) % isz; i1 = (i0 + 1) % isz; }
}

__device__ __forceinline__ float4 sampleNoise(const float* __restrict__ tex, uint32_t nw, uint32_t nh, float u, float v) { // linear, REPEAT
    int x0, x1, y0, y1; float fx, fy;
    bilinearSetupRepeat(u, nw, x0, x1, fx); bilinearSetupRepeat(v, nh, y0, y1, fy);
    const float4* t = reinterpret_cast<const float4*>(tex);
    const float4 t00 = __ldg(t + size_t(y0) * nw + x0), t10 = __ldg(t + size_t(y0) * nw + x1), t01 = __ldg(t + size_t(y1) * nw + x0), t11 = __ldg(t + size_t(y1) * nw + x1);
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    float4 r;
    r.x = (t00.x * gx + t10.x * fx) * gy + (t01.x * gx + t11.x * fx) * fy;
    r.y = (t00.y * gx + t10.y * fx) * gy + (t01.y * gx + t11.y * fx) * fy;
    r.z = (t00.z * gx + t10.z * fx) * gy + (t01.z * gx + t11.z * fx) * fy;
    r.w = (t00.w * gx + t10.w * fx) * gy + (t01.w * gx + t11.w * fx) * fy;
    return r;
}

// directLight.rgen:42-99. ALPHA: the scene has a texture list; the pipeline's only hit group is anyhit.rahit
// (src/RenderPasses/DirectLightPipeline.cpp:43-51), i.e. the shadow ray passes through texels with alpha < 0.01.
template <bool ALPHA>
__global__ void __launch_bounds__(128) k_direct_light(DeviceScene sc, vkx_light light, const float* __restrict__ noiseSlice, uint32_t nw, uint32_t nh, uint32_t W, uint32_t H,
                                                      const float4* __restrict__ posDepth, const float4* __restrict__ normalMetal, const float4* __restrict__ previous,
                                                      float4* __restrict__ out, float4* __restrict__ dbgDirs, uint8_t* __restrict__ dbgMask) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    const size_t pix = size_t(y) * W + x;
    const float4 pd = posDepth[pix], nm = normalMetal[pix];
    if (dbgMask) dbgMask[pix] = 0;
    if (pd.w <= 0.0f) { out[pix] = make_float4(-1.f, -1.f, -1.f, -1.f); return; }
    const v3 normal = mk3(nm.x, nm.y, nm.z);
    v3 direction = norm3(mk3(light.direction[0], light.direction[1], light.direction[2]));
    const float angle = 0.02f;
    const float4 noise = sampleNoise(noiseSlice, nw, nh, float(x) / 64.0f, float(y) / 64.0f);
    const v3 temp = rotateAxis(direction, norm3(cross3(normal, direction)), 2.0f * (noise.x - 0.5f) * angle);
    direction = rotateAxis(temp, direction, 2.0f * VKX_PI * noise.y);
    if (dbgDirs) dbgDirs[pix] = make_float4(direction.x, direction.y, direction.z, 0.f);
    if (dot3(direction, normal) > 0.0f) {
        const Ray r = makeRay(pd.x, pd.y, pd.z, direction.x, direction.y, direction.z);
        HitRec h;
        const bool shadowed = traverse<true, ALPHA>(sc.nodes, sc.tris, r, 0.01f, 10000.0f, 0xFFu, h, &sc);
        if (dbgMask) dbgMask[pix] = shadowed ? 2 : 1;
        float outColor = 0.0f;
        if (!shadowed) { outColor = 1.0f; if (direction.y < 0.0f) outColor *= 1.0f - clampS(-direction.y, 0.0f, 0.1f) / 0.1f; }
        const float4 pv = previous[pix];
        out[pix] = make_float4(outColor, pv.y, pv.z, 1.0f);
    } else out[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// gaussian(stdDev, dist) of directLightFilter.glsl:29-31 = norm(stdDev) * exp(-(dist * dist) * invTwoVar(stdDev)); the two factors that
// only depend on stdDev are hoisted out of the tap loops and exp is ex2.approx (tolerance of the filter outputs: 1e-3 absolute).
__device__ __forceinline__ float gaussNorm(float stdDev) { return 1.0f / (sqrtf(2.0f * 3.14159f) * stdDev); }
__device__ __forceinline__ float gaussInvTwoVar(float stdDev) { return 1.0f / (2.0f * stdDev * stdDev); }
#define MAX_DEV 7.0f
#define I_MAX_DEV 8
#define DEPTH_FACTOR (1.0f / 0.5f)
#define BASE_HYST 0.94f
#define DEPTH_STD 0.01f
#define HIST_THRESH 0.05f

__device__ __forceinline__ int filterWindow(float depth, float& stdDev) {
    stdDev = 1.0f + maxS(1.0f, MAX_DEV / (maxS(1.0f, DEPTH_FACTOR * depth)));
    return int(clampS(ceilf(sqrtf(-2.0f * stdDev * stdDev * logf(0.01f * stdDev * sqrtf(2.0f * 3.14159f)))), 1.0f, float(I_MAX_DEV)));
}

// Tap loop shared by both passes. The reference's factor gaussian(stdDev, i) * gaussian(depthStdDev, |dz|) is evaluated as
// w[|i|] * ex2(dz^2 * cd): w[] holds the spatial gaussian times both normalisations (9 values per pixel, zero beyond the window),
// cd = -log2(e) / (2 depthStdDev^2). Taps the reference skips at the image border are given depth = +inf in shared memory
// (ex2(-inf) = 0, an exact zero factor); the one out-of-bounds tap it does read (coordinate == size) is stored as
// depth 0 / value 0. Taps are accumulated in the reference's order (ascending offset) with FMAs.
__device__ __forceinline__ float ex2Approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define LOG2E 1.4426950408889634f
__device__ __forceinline__ void spatialWeights(float depth, float (&w)[I_MAX_DEV + 1]) {
    float stdDev; const int window = filterWindow(depth, stdDev);
    const float gn = gaussNorm(stdDev) * gaussNorm(DEPTH_STD);
    const float gv = -gaussInvTwoVar(stdDev) * LOG2E;
#pragma unroll
    for (int k = 0; k <= I_MAX_DEV; ++k) w[k] = k <= window ? gn * ex2Approx(float(k * k) * gv) : 0.0f;
}
template <int STRIDE>
__device__ __forceinline__ float4 filterTaps(const float4* __restrict__ sIn, int c, float depth) { // sIn: (value.xyz, depth)
    float w[I_MAX_DEV + 1];
    spatialWeights(depth, w);
    const float cd = -gaussInvTwoVar(DEPTH_STD) * LOG2E;
    float totalFactor = 0.0f; float4 fin = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = -I_MAX_DEV; i <= I_MAX_DEV; ++i) {
        const float4 v = sIn[c + i * STRIDE];
        const float dz = depth - v.w;
        const float factor = w[i < 0 ? -i : i] * ex2Approx(dz * dz * cd);
        totalFactor += factor;
        fin.x = __fmaf_rn(factor, v.x, fin.x); fin.y = __fmaf_rn(factor, v.y, fin.y); fin.z = __fmaf_rn(factor, v.z, fin.z);
    }
    // the temporal stage tests fin.z == 1.0f: sum / sum must stay exactly 1 although the quotient is a multiplication by the reciprocal
    if (totalFactor > 1e-2f) {
        const float inv = 1.0f / totalFactor;
        fin.x = fin.x == totalFactor ? 1.0f : fin.x * inv; fin.y = fin.y == totalFactor ? 1.0f : fin.y * inv; fin.z = fin.z == totalFactor ? 1.0f : fin.z * inv;
    } else { fin.x = fin.y = fin.z = 0.f; }
    return fin;
}
// halo depth for coordinate q of an axis of length n: inside -> the image, q == n -> 0 (read by the reference as an out-of-bounds load), else +inf (skipped by it)
__device__ __forceinline__ float haloDepth(int q, int n) { return q == n ? 0.0f : __int_as_float(0x7F800000); }

// X pass: one CTA = 256 consecutive pixels of one row; depth + input staged in shared memory with an 8-texel halo.
__global__ void __launch_bounds__(256) k_filter_x(uint32_t W, uint32_t H, const float4* __restrict__ posDepth, const float4* __restrict__ in, float4* __restrict__ out) {
    __shared__ float4 sIn[256 + 2 * I_MAX_DEV];
    const int y = int(blockIdx.y), x0 = int(blockIdx.x) * 256, tid = int(threadIdx.x);
    for (int i = tid; i < 256 + 2 * I_MAX_DEV; i += 256) {
        const int x = x0 + i - I_MAX_DEV;
        float4 v = make_float4(0.f, 0.f, 0.f, haloDepth(x, int(W)));
        if (x >= 0 && x < int(W)) { v = in[size_t(y) * W + x]; v.w = posDepth[size_t(y) * W + x].w; }
        sIn[i] = v;
    }
    __syncthreads();
    const int x = x0 + tid;
    if (x >= int(W)) return;
    const float depth = sIn[tid + I_MAX_DEV].w;
    const float4 fin = filterTaps<1>(sIn, tid + I_MAX_DEV, depth);
    out[size_t(y) * W + x] = make_float4(fin.x, fin.y, fin.z, depth);
}

// Y pass + temporal accumulation: one CTA = 16 columns x 64 rows, (64 + 16) x 16 tile in shared memory.
__global__ void __launch_bounds__(256) k_filter_y(uint32_t W, uint32_t H, const float4* __restrict__ posDepth, const float4* __restrict__ in, const float4* __restrict__ prevImg,
                                                  M4 prevView, M4 prevProj, float3 prevOrigin, float4* __restrict__ out) {
    constexpr int TW = 16;
    __shared__ float4 sIn[(64 + 2 * I_MAX_DEV) * TW];
    const int x0 = int(blockIdx.x) * TW, y0 = int(blockIdx.y) * 64;
    const int tx = int(threadIdx.x) & (TW - 1), ty = int(threadIdx.x) / TW; // 16 rows of 16
    const int x = x0 + tx;
    for (int r = ty; r < 64 + 2 * I_MAX_DEV; r += 256 / TW) {
        const int y = y0 + r - I_MAX_DEV;
        float4 v = make_float4(0.f, 0.f, 0.f, haloDepth(y, int(H)));
        if (x < int(W) && y >= 0 && y < int(H)) v = in[size_t(y) * W + x]; // .w of the X pass output is the depth
        sIn[r * TW + tx] = v;
    }
    __syncthreads();
    if (x >= int(W)) return;
    for (int ry = ty; ry < 64; ry += 256 / TW) {
        const int y = y0 + ry;
        if (y >= int(H)) break;
        const int c = (ry + I_MAX_DEV) * TW + tx;
        const float4 pd = posDepth[size_t(y) * W + x];
        const float depth = pd.w;
        float4 fin = filterTaps<TW>(sIn, c, depth);
        // temporal accumulation, directLightFilter.glsl:110-142
        fin.x = clampS(fin.x, 0.0f, 1.0f);
        fin.y = fin.x * fin.x;
        const v3 position = mk3(pd.x, pd.y, pd.z);
        const float4 vp = mulM4(prevView, position.x, position.y, position.z, 1.0f);
        float4 pc = mulM4(prevProj, vp.x, vp.y, vp.z, vp.w);
        pc.x = pc.x / pc.w; pc.y = pc.y / pc.w;
        pc.x = (0.5f * pc.x + 0.5f) * float(W);
        pc.y = (0.5f * pc.y + 0.5f) * float(H);
        float4 previousValue = make_float4(0.f, 0.f, 0.f, 0.f);
        float hysteresis = BASE_HYST;
        if (fin.z > 0.0f) hysteresis = fin.z == 1.0f ? 0.5f : 0.0f;
        if (pc.x >= float(W) || pc.x < 0.0f || pc.y >= float(H) || pc.y < 0.0f) hysteresis = 0.0f;
        else {
            previousValue = prevImg[size_t(int(pc.y)) * W + size_t(int(pc.x))];
            const v3 po = mk3(prevOrigin.x, prevOrigin.y, prevOrigin.z);
            const v3 previousPosition = po + previousValue.w * norm3(position - po);
            const float factor = clampS(len3(position - previousPosition), 0.0f, HIST_THRESH) / HIST_THRESH;
            hysteresis *= 1.0f - clampS(factor, 0.0f, 1.0f);
            const float variance = fabsf(previousValue.x * previousValue.x - previousValue.y);
            if (variance < 0.25f && fabsf(previousValue.x - fin.x) > 0.75f) { hysteresis = 0.0f; fin.z = 1.0f; }
            else fin.z = 0.0f;
        }
        out[size_t(y) * W + x] = make_float4(hysteresis * previousValue.x + (1.0f - hysteresis) * fin.x, hysteresis * previousValue.y + (1.0f - hysteresis) * fin.y,
                                             hysteresis * previousValue.z + (1.0f - hysteresis) * fin.z, depth);
    }
}

void inverse4(const float* a, float* out) { // same cofactor expansion as the oracle, fp32
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    float id = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
}

} // namespace

int shadowGBuffer(vkx_ctx* ctx, const vkx_camera& cam) {
    M4 iv, ip;
    inverse4(cam.view, iv.m); inverse4(cam.proj, ip.m);
    dim3 grid(divUp(ctx->shW, 128), ctx->shH);
    if (!ctx->hTextures.empty()) k_gbuffer<true><<<grid, 128, 0, ctx->stream>>>(deviceScene(ctx), iv, ip, make_float3(cam.origin[0], cam.origin[1], cam.origin[2]), ctx->shW, ctx->shH, ctx->dPosDepth, ctx->dNormalMetal, ctx->dAlbedoRough, ctx->dEmissive);
    else k_gbuffer<false><<<grid, 128, 0, ctx->stream>>>(deviceScene(ctx), iv, ip, make_float3(cam.origin[0], cam.origin[1], cam.origin[2]), ctx->shW, ctx->shH, ctx->dPosDepth, ctx->dNormalMetal, ctx->dAlbedoRough, ctx->dEmissive);
    LAUNCH_CHECK(ctx);
    return VKX_OK;
}

int shadowFrame(vkx_ctx* ctx, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light) {
    cudaStream_t st = ctx->stream;
    const uint32_t W = ctx->shW, H = ctx->shH;
    const float* slice = ctx->dNoise + size_t(cur.frameIndex % ctx->noiseSlices) * ctx->noiseW * ctx->noiseH * 4;
    float4* previous = ctx->dShFinal[ctx->shCur];
    float4* next = ctx->dShFinal[ctx->shCur ^ 1];
    M4 pv, pp; memcpy(pv.m, prev.view, 64); memcpy(pp.m, prev.proj, 64);
    CUDA_TRY(ctx, cudaEventRecord(ctx->sev[0], st));
    if (!ctx->hTextures.empty()) k_direct_light<true><<<dim3(divUp(W, 128), H), 128, 0, st>>>(deviceScene(ctx), light, slice, ctx->noiseW, ctx->noiseH, W, H, ctx->dPosDepth, ctx->dNormalMetal, previous, ctx->dShRaw, ctx->dShDirs, ctx->dShMask);
    else k_direct_light<false><<<dim3(divUp(W, 128), H), 128, 0, st>>>(deviceScene(ctx), light, slice, ctx->noiseW, ctx->noiseH, W, H, ctx->dPosDepth, ctx->dNormalMetal, previous, ctx->dShRaw, ctx->dShDirs, ctx->dShMask);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->sev[1], st));
    k_filter_x<<<dim3(divUp(W, 256), H), 256, 0, st>>>(W, H, ctx->dPosDepth, ctx->dShRaw, ctx->dShX); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->sev[2], st));
    k_filter_y<<<dim3(divUp(W, 16), divUp(H, 64)), 256, 0, st>>>(W, H, ctx->dPosDepth, ctx->dShX, previous, pv, pp, make_float3(prev.origin[0], prev.origin[1], prev.origin[2]), next); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->sev[3], st));
    ctx->shCur ^= 1;
    return VKX_OK;
}
