// 1-spp ray-traced reflections with a roughness- and depth-aware separable Gaussian filter and reprojected temporal accumulation
// (SURVEY 8(f) rank 3). Replaces (reference): src/shaders/reflection.rgen:117-189 (the `#if 1` single-sample branch), the
// closest-hit / miss / shadow shaders it invokes (closesthit.glsl with recursionDepth = 1, miss.rmiss, shadow.rmiss),
// src/shaders/reflectionFilter.glsl:58-147 (X and Y variants), the dispatch order of src/SwapchainManagement.cpp:401-455 and the
// editor's history copy (here a ping-pong). Motion vectors are zero (static scenes, GBuffer.vert.glsl:46-48); out-of-bounds image
// loads return 0 and out-of-bounds stores are dropped. Compiled with --fmad=false like the traversal it contains.
#include <algorithm>
#include "common.cuh"
#include "traverse.cuh"
#include "hitshade.cuh"
#include "ddgi_common.cuh"

namespace {

struct M4r { float m[16]; }; // column-major
__device__ __forceinline__ float4 mulM4r(const M4r& M, float x, float y, float z, float w) { // glm: (m0*x + m1*y) + (m2*z + m3*w)
    float4 r;
    r.x = (M.m[0] * x + M.m[4] * y) + (M.m[8] * z + M.m[12] * w);
    r.y = (M.m[1] * x + M.m[5] * y) + (M.m[9] * z + M.m[13] * w);
    r.z = (M.m[2] * x + M.m[6] * y) + (M.m[10] * z + M.m[14] * w);
    r.w = (M.m[3] * x + M.m[7] * y) + (M.m[11] * z + M.m[15] * w);
    return r;
}

__device__ __forceinline__ v3 rotateAxisR(v3 p, v3 axis, float angle) { // common.glsl:6-8
    float sn, cs; sincosf(angle, &sn, &cs);
    return mix3(dot3(axis, p) * axis, p, cs) + cross3(axis, p) * sn;
}

__device__ __forceinline__ void wrapSetup(float u, uint32_t size, int& i0, int& i1, float& f) { // linear filter, REPEAT addressing
    const float x = u * float(size) - 0.5f;
    const float fl = floorf(x);
    f = x - fl;
    const int isz = int(size), i = int(fl);
    if ((size & (size - 1u)) == 0u) { i0 = i & (isz - 1); i1 = (i0 + 1) & (isz - 1); }
    else { i0 = ((i % isz) + isz) % isz; i1 = (i0 + 1) % isz; }
}
__device__ __forceinline__ float2 sampleNoiseXY(const float* __restrict__ tex, uint32_t nw, uint32_t nh, float u, float v) {
    int x0, x1, y0, y1; float fx, fy;
    wrapSetup(u, nw, x0, x1, fx); wrapSetup(v, nh, y0, y1, fy);
    const float4* t = reinterpret_cast<const float4*>(tex);
    const float4 t00 = __ldg(t + size_t(y0) * nw + x0), t10 = __ldg(t + size_t(y0) * nw + x1), t01 = __ldg(t + size_t(y1) * nw + x0), t11 = __ldg(t + size_t(y1) * nw + x1);
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    return make_float2((t00.x * gx + t10.x * fx) * gy + (t01.x * gx + t11.x * fx) * fy, (t00.y * gx + t10.y * fx) * gy + (t01.y * gx + t11.y * fx) * fy);
}

// reflection.rgen:117-163, ray generation: one thread per pixel (16 x 8 tiles). Pixels that do not reflect (sky, or rough dielectrics:
// roughness >= 0.4 and metalness <= 0.01) get their zero output here; the others get their jittered direction and a slot in the
// compact ray queue, so the heavy trace + shade kernel only runs over pixels that have work.
__global__ void __launch_bounds__(128) k_reflect_gen(float3 camOrigin, const float* __restrict__ noiseSlice, uint32_t nw, uint32_t nh, int offx, int offy, uint32_t W, uint32_t H,
                                                     const float4* __restrict__ posDepth, const float4* __restrict__ normalMetal, const float4* __restrict__ albedoRough,
                                                     float4* __restrict__ out, float4* __restrict__ dirs, uint32_t* __restrict__ queue, uint32_t* __restrict__ queueCount,
                                                     vkx_hit* __restrict__ dbgHits, uint8_t* __restrict__ dbgMask) {
    const uint32_t x = blockIdx.x * 16u + (threadIdx.x & 15u), y = blockIdx.y * 8u + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    const size_t pix = size_t(y) * W + x;
    const float4 pd = __ldg(posDepth + pix), nm = __ldg(normalMetal + pix);
    const float depth = pd.w, metalness = nm.w, roughness = __ldg(albedoRough + pix).w;
    if (dbgMask) { dbgMask[pix] = 0; vkx_hit z; z.t = -1.0f; z.u = z.v = 0.f; z.instance = z.primitive = 0u; dbgHits[pix] = z; }
    if (!(depth > 0.0f && (roughness < 0.4f || metalness > 0.01f))) { out[pix] = make_float4(0.f, 0.f, 0.f, 0.f); dirs[pix] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const v3 position = mk3(pd.x, pd.y, pd.z), normal = mk3(nm.x, nm.y, nm.z);
    const v3 toOrigin = norm3(mk3(camOrigin.x, camOrigin.y, camOrigin.z) - position);
    const v3 reflectDir = norm3(reflect3(-toOrigin, normal));
    const float2 noise = sampleNoiseXY(noiseSlice, nw, nh, float(offx + int(x)) / 64.0f, float(offy + int(y)) / 64.0f);
    const float theta = roughness * (noise.x - 0.5f) * 2.0f * VKX_PI;
    const float phi = (noise.y - 0.5f) * 2.0f * VKX_PI;
    v3 tangent;
    if (dot3(reflectDir, normal) < 0.9f) tangent = norm3(cross3(reflectDir, normal));
    else tangent = norm3(cross3(reflectDir, mk3(1.0f, 0.0f, 0.0f)));
    v3 direction = rotateAxisR(reflectDir, tangent, theta);
    direction = rotateAxisR(direction, reflectDir, phi);
    dirs[pix] = make_float4(direction.x, direction.y, direction.z, roughness);
    queue[warpAppend(queueCount)] = uint32_t(pix);
}

// reflection.rgen:186-187 with closesthit.glsl / miss.rmiss / shadow.rmiss: closest hit, shading from the irradiance volume, sun
// shadow ray, colour compression. One thread per queued pixel.
// TEX: the scene has a texture list. The reflection pipeline's hit group has anyhit.rahit (src/RenderPasses/ReflectionPipeline.cpp:51),
// so both traversals run the alpha cut-out test, and the closest hit shades with textures (ray differentials reflection.rgen:169-170).
template <bool TEX>
__global__ void __launch_bounds__(128) k_reflect_shade(DeviceScene sc, DeviceProbes pr, vkx_light light, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ queueCount,
                                                       const float4* __restrict__ posDepth, const float4* __restrict__ normalMetal, const float4* __restrict__ dirs, float4* __restrict__ out,
                                                       vkx_hit* __restrict__ dbgHits, uint8_t* __restrict__ dbgMask) {
    const uint32_t n = *queueCount;
    const v3 lightDir = mk3(light.direction[0], light.direction[1], light.direction[2]);
    const v3 lightColor = mk3(light.color[0], light.color[1], light.color[2]);
    const GridConsts gc = makeGridConsts(pr.grid);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pix = queue[i];
        const float4 pd = __ldg(posDepth + pix), d4 = dirs[pix];
        const v3 position = mk3(pd.x, pd.y, pd.z), direction = mk3(d4.x, d4.y, d4.z);
        const Ray r = makeRay(position.x, position.y, position.z, direction.x, direction.y, direction.z);
        HitRec h;
        v3 color;
        if (!traverse<false, TEX>(sc.nodes, sc.tris, r, 0.1f, 10000.0f, 0xFFu, h, &sc)) { // miss.rmiss
            color = skyColor(position, direction, lightDir, lightColor, light.color[3]);
            if (dbgMask) dbgMask[pix] = 1;
        } else {
            vkx_hit vh; vh.t = h.t; vh.u = h.u; vh.v = h.v; vh.instance = h.inst; vh.primitive = h.prim;
            if (dbgHits) dbgHits[pix] = vh;
            if (h.prim & 0x80000000u) { color = mk3(0.0f); if (dbgMask) dbgMask[pix] = 2; } // closesthit.glsl:137-141
            else {
                const v3 hitPos = pointOnRayExact(position, direction, h.t);
                v3 base, lit;
                if (TEX) {
                    const float4 nm = __ldg(normalMetal + pix);
                    const v3 normal = mk3(nm.x, nm.y, nm.z);
                    const v3 raydx = rotateAxisH(direction, normal, 0.001f), raydy = rotateAxisH(direction, cross3(normal, direction), 0.001f);
                    shadeFrontHit<true>(sc, pr, gc, lightDir, lightColor, direction, hitPos, vh, base, lit, position, raydx, raydy);
                } else shadeFrontHit<false>(sc, pr, gc, lightDir, lightColor, direction, hitPos, vh, base, lit);
                const Ray sr = makeRay(hitPos.x, hitPos.y, hitPos.z, lightDir.x, lightDir.y, lightDir.z);
                HitRec sh;
                const bool shadowed = traverse<true, TEX>(sc.nodes, sc.tris, sr, 0.1f, 10000.0f, 0xFFu, sh, &sc);
                color = shadowed ? base : lit;
                if (dbgMask) dbgMask[pix] = shadowed ? 4 : 3;
            }
        }
        // colorCompression = reinhard_whitepoint(v, 1.0) (reflection.rgen:95-110)
        const float maxValue = 1.0f;
        const v3 comp = color * ((color / mk3(maxValue * maxValue)) + 1.0f) / (color + 1.0f);
        out[pix] = make_float4(comp.x, comp.y, comp.z, d4.w);
    }
}

#define R_MAX_DEV 5.0f
#define R_I_MAX_DEV 5
#define R_DEPTH_FACTOR (1.0f / 20.0f)
#define R_BASE_HYST 0.98f
#define R_DEPTH_STD 0.1f

__device__ __forceinline__ float rgaussian(float stdDev, float dist) { // reflectionFilter.glsl:37-39
    return (1.0f / (sqrtf(2.0f * 3.14159f) * stdDev)) * expf(-(dist * dist) / (2.0f * stdDev * stdDev));
}
__device__ __forceinline__ float4 loadImg(const float4* __restrict__ img, int W, int H, int x, int y) {
    return (x < 0 || y < 0 || x >= W || y >= H) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(img + size_t(y) * W + x);
}

// reflectionFilter.glsl:58-147; most pixels carry no reflection (roughness 0 in the cleared output) and leave through the
// stdDev == 0 pass-through, so the taps are read straight from L1/L2 instead of a staged tile.
template <int DIR>
__global__ void __launch_bounds__(128) k_refl_filter(int W, int H, const float4* __restrict__ posDepth, const float4* __restrict__ in, const float4* __restrict__ prevImg,
                                                     M4r prevView, M4r prevProj, float3 curOrigin, float3 prevOrigin, float4* __restrict__ out) {
    const int x = int(blockIdx.x * 16u + (threadIdx.x & 15u)), y = int(blockIdx.y * 8u + (threadIdx.x >> 4));
    if (x >= W || y >= H) return;
    const size_t pix = size_t(y) * W + x;
    const float4 pd = __ldg(posDepth + pix), center = __ldg(in + pix);
    const float depth = pd.w, roughness = center.w;
    const float stdDev = maxS(0.0f, R_MAX_DEV * roughness / maxS(1.0f, R_DEPTH_FACTOR * depth));
    if (stdDev == 0.0f) { out[pix] = center; return; }
    const float sqrDev = stdDev * stdDev;
    const int window = int(clampS(ceilf(sqrtf(-2.0f * sqrDev * logf(0.01f * stdDev * sqrtf(2.0f * 3.14159f)))), 1.0f, R_MAX_DEV));
    const int c = DIR == 0 ? x : y, n = DIR == 0 ? W : H;
    const int minOffset = -min(window, c), maxOffset = min(window, n - c);
    float totalFactor = 0.0f; float4 fin = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = minOffset; i <= maxOffset; ++i) {
        const int ox = x + (DIR == 0 ? i : 0), oy = y + (DIR == 1 ? i : 0);
        float factor = rgaussian(stdDev, float(i));
        factor *= rgaussian(R_DEPTH_STD, fabsf(depth - loadImg(posDepth, W, H, ox, oy).w));
        totalFactor += factor;
        const float4 v = loadImg(in, W, H, ox, oy);
        fin.x += factor * v.x; fin.y += factor * v.y; fin.z += factor * v.z;
    }
    if (totalFactor > 1e-2f) { fin.x = fin.x / totalFactor; fin.y = fin.y / totalFactor; fin.z = fin.z / totalFactor; } else { fin.x = fin.y = fin.z = 0.f; }
    if (DIR == 0) { out[pix] = make_float4(fin.x, fin.y, fin.z, roughness); return; }
    float hysteresis = R_BASE_HYST;
    float4 previousValue = make_float4(0.f, 0.f, 0.f, 0.f);
    const v3 position = mk3(pd.x, pd.y, pd.z);
    const v3 co = mk3(curOrigin.x, curOrigin.y, curOrigin.z), po = mk3(prevOrigin.x, prevOrigin.y, prevOrigin.z);
    const float cameraMovement = len3(co - po);
    hysteresis *= maxS(0.0f, 1.0f - cameraMovement);
    if (hysteresis > 0.0f) {
        const float4 vp = mulM4r(prevView, position.x, position.y, position.z, 1.0f);
        float4 pc = mulM4r(prevProj, vp.x, vp.y, vp.z, vp.w);
        pc.x = pc.x / pc.w; pc.y = pc.y / pc.w;
        pc.x = (0.5f * pc.x + 0.5f) * float(W);
        pc.y = (0.5f * pc.y + 0.5f) * float(H);
        if (pc.x > float(W) || pc.x < 0.0f || pc.y > float(H) || pc.y < 0.0f) hysteresis = 0.0f; // `>`: a coordinate equal to the extent loads 0 (:128)
        else {
            previousValue = loadImg(prevImg, W, H, int(pc.x), int(pc.y));
            const v3 previousPosition = po + previousValue.w * norm3(position - po);
            const float factor = len3(position - previousPosition);
            hysteresis *= 1.0f - clampS(factor, 0.0f, 1.0f);
        }
    }
    out[pix] = make_float4(mixf(fin.x, previousValue.x, hysteresis), mixf(fin.y, previousValue.y, hysteresis), mixf(fin.z, previousValue.z, hysteresis), depth);
}

} // namespace

int reflectionFrame(vkx_ctx* ctx, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light) {
    cudaStream_t st = ctx->stream;
    const uint32_t W = ctx->shW, H = ctx->shH;
    const float* slice = ctx->dNoise + size_t(cur.frameIndex % ctx->noiseSlices) * ctx->noiseW * ctx->noiseH * 4;
    const float4* previous = ctx->dReflFinal[ctx->reflCur];
    float4* final_ = ctx->dReflFinal[ctx->reflCur ^ 1];
    if (!ctx->rev[0]) for (auto& e : ctx->rev) CUDA_TRY(ctx, cudaEventCreate(&e));
    const dim3 grid(divUp(W, 16), divUp(H, 8));
    M4r pv, pp; for (int i = 0; i < 16; ++i) { pv.m[i] = prev.view[i]; pp.m[i] = prev.proj[i]; }
    const float3 co = make_float3(cur.origin[0], cur.origin[1], cur.origin[2]), po = make_float3(prev.origin[0], prev.origin[1], prev.origin[2]);
    CUDA_TRY(ctx, cudaEventRecord(ctx->rev[0], st));
    vkx_hit* dbgHits = ctx->debugBuffers ? ctx->dReflHits : nullptr; uint8_t* dbgMask = ctx->debugBuffers ? ctx->dReflMask : nullptr;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dReflCount, 0, 4, st));
    k_reflect_gen<<<grid, 128, 0, st>>>(co, slice, ctx->noiseW, ctx->noiseH, int(cur.frameIndex / 64u), int(cur.frameIndex / 64u / 64u), W, H, ctx->dPosDepth, ctx->dNormalMetal,
                                        ctx->dAlbedoRough, ctx->dReflRaw, ctx->dReflDirs, ctx->dReflQueue, ctx->dReflCount, dbgHits, dbgMask);
    LAUNCH_CHECK(ctx);
    const unsigned shadeBlocks = std::min<unsigned>(divUp(size_t(W) * H, 128), unsigned(ctx->smCount) * 16u);
    if (!ctx->hTextures.empty()) k_reflect_shade<true><<<shadeBlocks, 128, 0, st>>>(deviceScene(ctx), deviceProbes(ctx), light, ctx->dReflQueue, ctx->dReflCount, ctx->dPosDepth, ctx->dNormalMetal, ctx->dReflDirs, ctx->dReflRaw, dbgHits, dbgMask);
    else k_reflect_shade<false><<<shadeBlocks, 128, 0, st>>>(deviceScene(ctx), deviceProbes(ctx), light, ctx->dReflQueue, ctx->dReflCount, ctx->dPosDepth, ctx->dNormalMetal, ctx->dReflDirs, ctx->dReflRaw, dbgHits, dbgMask);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->rev[1], st));
    k_refl_filter<0><<<grid, 128, 0, st>>>(int(W), int(H), ctx->dPosDepth, ctx->dReflRaw, nullptr, pv, pp, co, po, ctx->dReflX); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->rev[2], st));
    k_refl_filter<1><<<grid, 128, 0, st>>>(int(W), int(H), ctx->dPosDepth, ctx->dReflX, previous, pv, pp, co, po, final_); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->rev[3], st));
    ctx->reflCur ^= 1; ctx->reflValid = true;
    return VKX_OK;
}
