// Texture look-ups of the ray-tracing shaders ("sampler spec v1"). Replaces the fixed-function sampler behind
//   textureGrad(textures[i], texCoord, grad.xy, grad.zw)   reference src/shaders/closesthit.glsl:163-192
//   texture(textures[m.albedoTexture], texCoord)           reference src/shaders/anyhit.rahit:43
// with the sampler state of reference src/Resources.cpp:88-124 (glTF filters / wrap modes, mipLodBias 0, minLod 0, maxLod = levels).
// The arithmetic is the Vulkan specification's texel-filtering equations with exact fp32 weights and an isotropic level of detail;
// the weight / coordinate arithmetic is written with _rn intrinsics so that it does not depend on the translation unit's
// floating-point flags: base-level look-ups (the alpha cut-out decisions) are bit-identical to oracle/texture.h.
#pragma once
#include "common.cuh"

#define VKX_TEX_SRGB 1u
#define VKX_TEX_MAG_LINEAR 2u
#define VKX_TEX_MIN_LINEAR 4u
#define VKX_TEX_MIP_LINEAR 8u
#define VKX_TEX_WRAP_S_SHIFT 4 // 2 bits: 0 REPEAT, 1 CLAMP_TO_EDGE, 2 MIRRORED_REPEAT
#define VKX_TEX_WRAP_T_SHIFT 6

__device__ __forceinline__ int texWrap(int i, int size, uint32_t mode) { // Vulkan "Wrapping Operation"
    if (mode == 1u) return min(max(i, 0), size - 1);
    if (mode == 0u) { const int m = i % size; return m < 0 ? m + size : m; }
    const int p = 2 * size; int m = i % p; if (m < 0) m += p;
    int n = m - size; n = n >= 0 ? n : -(1 + n);
    return (size - 1) - n;
}

// floor of a texel-space coordinate; pinned to +-2^30 (NaN -> -2^30) so that the int conversion is defined
__device__ __forceinline__ float texPinnedFloor(float x) { return fminf(fmaxf(floorf(x), -1073741824.0f), 1073741824.0f); }

__device__ __forceinline__ float texLerp(float a, float b, float w) { return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, w)), __fmul_rn(b, w)); }
__device__ __forceinline__ float4 texLerp4(float4 a, float4 b, float w) { return make_float4(texLerp(a.x, b.x, w), texLerp(a.y, b.y, w), texLerp(a.z, b.z, w), texLerp(a.w, b.w, w)); }

// lut: 512 floats, [0, 256) sRGB code -> linear (decree T4), [256, 512) code / 255 (the same IEEE quotient as float(code) / 255.0f,
// tabulated on the host: four table look-ups per texel instead of up to four IEEE divisions).
__device__ __forceinline__ float4 texDecode(const float* __restrict__ lut, uint32_t flags, uint32_t w) {
    const float* rgb = (flags & VKX_TEX_SRGB) ? lut : lut + 256;
    return make_float4(__ldg(rgb + (w & 0xFFu)), __ldg(rgb + ((w >> 8) & 0xFFu)), __ldg(rgb + ((w >> 16) & 0xFFu)), __ldg(lut + 256 + (w >> 24)));
}

__device__ __forceinline__ float4 texFetch(const DeviceScene& sc, const DeviceTexture& t, uint32_t level, int x, int y) {
    const int w = int(max(1u, t.width >> level)), h = int(max(1u, t.height >> level));
    const int xx = texWrap(x, w, (t.flags >> VKX_TEX_WRAP_S_SHIFT) & 3u), yy = texWrap(y, h, (t.flags >> VKX_TEX_WRAP_T_SHIFT) & 3u);
    return __ldg(sc.texelsDecoded + size_t(t.levelOffset[level]) + size_t(yy) * size_t(w) + size_t(xx)); // = texDecode(lut, flags, texels[...]), decoded at upload
}

__device__ __forceinline__ float4 texSampleLevel(const DeviceScene& sc, const DeviceTexture& t, uint32_t level, float s, float tt, bool linear) {
    const float w = float(max(1u, t.width >> level)), h = float(max(1u, t.height >> level));
    if (!linear) return texFetch(sc, t, level, int(texPinnedFloor(__fmul_rn(s, w))), int(texPinnedFloor(__fmul_rn(tt, h))));
    const float u = __fsub_rn(__fmul_rn(s, w), 0.5f), v = __fsub_rn(__fmul_rn(tt, h), 0.5f);
    const float fu = texPinnedFloor(u), fv = texPinnedFloor(v);
    const float a = __fsub_rn(u, fu), b = __fsub_rn(v, fv);
    const int i0 = int(fu), j0 = int(fv);
    const float4 top = texLerp4(texFetch(sc, t, level, i0, j0), texFetch(sc, t, level, i0 + 1, j0), a);
    const float4 bot = texLerp4(texFetch(sc, t, level, i0, j0 + 1), texFetch(sc, t, level, i0 + 1, j0 + 1), a);
    return texLerp4(top, bot, b);
}

// texture(sampler2D, uv) in a ray-tracing stage: no implicit derivatives, base level
__device__ __forceinline__ float4 texSampleBase(const DeviceScene& sc, uint32_t index, float s, float tt) {
    const DeviceTexture t = sc.textures[index];
    return texSampleLevel(sc, t, 0u, s, tt, (t.flags & VKX_TEX_MAG_LINEAR) != 0u);
}

// textureGrad(sampler2D, uv, dPdx, dPdy)
__device__ __forceinline__ float4 texSampleGrad(const DeviceScene& sc, uint32_t index, float s, float tt, float dudx, float dvdx, float dudy, float dvdy) {
    const DeviceTexture t = sc.textures[index];
    const float w = float(t.width), h = float(t.height);
    const float ax = __fmul_rn(dudx, w), bx = __fmul_rn(dvdx, h), ay = __fmul_rn(dudy, w), by = __fmul_rn(dvdy, h);
    const float rhoX = __fsqrt_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(bx, bx))), rhoY = __fsqrt_rn(__fadd_rn(__fmul_rn(ay, ay), __fmul_rn(by, by)));
    const float rho = fmaxf(rhoX, rhoY);
    float lambda = 0.0f;
    if (rho > 0.0f) lambda = log2f(rho);
    if (!(lambda > 0.0f)) return texSampleLevel(sc, t, 0u, s, tt, (t.flags & VKX_TEX_MAG_LINEAR) != 0u);
    const float q = float(t.levels - 1u);
    const float d = fminf(fminf(lambda, float(t.levels)), q);
    const bool minLinear = (t.flags & VKX_TEX_MIN_LINEAR) != 0u;
    if (!(t.flags & VKX_TEX_MIP_LINEAR)) {
        const uint32_t level = uint32_t(fminf(fmaxf(__fsub_rn(ceilf(__fadd_rn(d, 0.5f)), 1.0f), 0.0f), q));
        return texSampleLevel(sc, t, level, s, tt, minLinear);
    }
    const float dhi = floorf(d);
    const uint32_t lhi = uint32_t(dhi), llo = min(lhi + 1u, t.levels - 1u);
    const float delta = __fsub_rn(d, dhi);
    const float4 c0 = texSampleLevel(sc, t, lhi, s, tt, minLinear);
    if (delta == 0.0f) return c0;
    return texLerp4(c0, texSampleLevel(sc, t, llo, s, tt, minLinear), delta);
}

// anyhit.rahit:24-48 for one candidate hit: true = ignoreIntersectionEXT
__device__ __forceinline__ bool anyHitIgnores(const DeviceScene& sc, uint32_t instance, uint32_t primitive, float u, float v) {
    const vkx_offset_entry oe = sc.offsets[__ldg(&sc.instances[instance].meshEntry)];
    const uint32_t tex = __ldg(&sc.materials[oe.materialIndex].albedoTexture);
    if (tex == VKX_INVALID_TEXTURE) return false;
    float tu[3], tv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* tc = sc.vertices[oe.vertexOffset + __ldg(sc.indices + oe.indexOffset + 3u * primitive + c)].texCoord;
        tu[c] = __ldg(tc); tv[c] = __ldg(tc + 1);
    }
    const float bx = __fsub_rn(__fsub_rn(1.0f, u), v);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(tu[0], bx), __fmul_rn(tu[1], u)), __fmul_rn(tu[2], v));
    const float t = __fadd_rn(__fadd_rn(__fmul_rn(tv[0], bx), __fmul_rn(tv[1], u)), __fmul_rn(tv[2], v));
    return texSampleBase(sc, tex, s, t).w < 1e-2f;
}
