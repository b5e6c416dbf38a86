// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
//
// "Sampler spec v1": the texture path of the reference's ray-tracing shaders, restated on the CPU.
//   upload + mip chain   src/vulkan/Image.cpp:62-111 (upload), :195-275 (generateMipmaps: levels = floor(log2(max(w, h))) + 1,
//                        src/vulkan/Image.cpp:29-31; level i = vkCmdBlitImage(level i-1, VK_FILTER_LINEAR) of max(1, w/2) x max(1, h/2))
//   formats              R8G8B8A8_SRGB for albedo / emissive, R8G8B8A8_UNORM for normal / metallic-roughness (src/Scene.cpp:43,671,677)
//   sampler              glTF sampler description -> VkSampler (src/Resources.cpp:8-43,88-124): mag/min filter, mipmap mode, wrap S/T,
//                        mipLodBias 0, minLod 0, maxLod = mip levels, anisotropy enabled
//   use                  textureGrad(textures[i], uv, grad.xy, grad.zw) in closesthit.glsl:163-192, texture(textures[i], uv) in anyhit.rahit:43
//
// Texel filtering is fixed-function in the reference (the driver's), so the arithmetic below is the Vulkan specification's
// ("Texel Input Operations" / "Image Sample Operations": scale factor, level-of-detail, (un)normalised coordinates, wrapping,
// linear filter weights), with these decrees where the specification leaves a choice:
//   T1  weights and texel values are exact fp32 (hardware uses ~8-bit fixed-point weights);
//   T2  level of detail is isotropic: lambda = log2(max(rho_x, rho_y)); the reference's anisotropic filtering (maxAnisotropy = the
//       device limit) has an implementation-defined footprint and is not modelled;
//   T3  rho = 0 or NaN (degenerate ray differentials, closesthit.glsl:73-74 divide by dot(raydx, normal)) selects lambda = 0 and the
//       magnification filter; texture() outside a fragment shader (anyhit.rahit) has no implicit derivatives: base level;
//   T4  sRGB decode by the exact piecewise formula evaluated in double precision per code (a 256-entry table); mip levels of sRGB
//       images are filtered in linear space and re-encoded to the nearest 8-bit code of the exact formula;
//   T5  the LINEAR blit of generateMipmaps follows the specification's blit equations (for even sizes a 2x2 box average).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "../include/vkx.h"

namespace otex {

enum Wrap : uint32_t { REPEAT = 0, CLAMP = 1, MIRROR = 2 };

struct Texture {
    uint32_t width = 0, height = 0, levels = 0;
    bool srgb = false, magLinear = true, minLinear = true, mipLinear = true;
    uint32_t wrapS = REPEAT, wrapT = REPEAT;
    std::vector<std::vector<uint32_t>> mip; // [level][y * w_level + x], r | g << 8 | b << 16 | a << 24
    uint32_t levelWidth(uint32_t l) const { return std::max(1u, width >> l); }
    uint32_t levelHeight(uint32_t l) const { return std::max(1u, height >> l); }
};

struct RGBA { float r, g, b, a; };

struct SrgbTables {
    float toLinear[256]; // T4
    float threshold[256]; // threshold[k] (k >= 1): smallest linear value that encodes to code >= k
    SrgbTables() {
        for (int i = 0; i < 256; ++i) {
            toLinear[i] = float(decode(double(i) / 255.0));
            threshold[i] = i == 0 ? 0.0f : float(decode((double(i) - 0.5) / 255.0));
        }
    }
    static double decode(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }
    uint32_t encode(float x) const { // number of thresholds <= x
        uint32_t lo = 0, hi = 255; // invariant: threshold[lo] <= x (threshold[0] = 0 and x >= 0 after the clamp)
        if (!(x > 0.0f)) return 0;
        while (lo < hi) { uint32_t mid = (lo + hi + 1) / 2; if (threshold[mid] <= x) lo = mid; else hi = mid - 1; }
        return lo;
    }
};
inline const SrgbTables& srgbTables() { static const SrgbTables t; return t; }

inline RGBA decodeTexel(const Texture& t, uint32_t w) {
    const SrgbTables& s = srgbTables();
    RGBA c;
    if (t.srgb) { c.r = s.toLinear[w & 0xFFu]; c.g = s.toLinear[(w >> 8) & 0xFFu]; c.b = s.toLinear[(w >> 16) & 0xFFu]; }
    else { c.r = float(w & 0xFFu) / 255.0f; c.g = float((w >> 8) & 0xFFu) / 255.0f; c.b = float((w >> 16) & 0xFFu) / 255.0f; }
    c.a = float(w >> 24) / 255.0f;
    return c;
}
inline uint32_t unormEncode(float x) { x = std::fmin(std::fmax(x, 0.0f), 1.0f); return uint32_t(x * 255.0f + 0.5f); }
inline uint32_t encodeTexel(const Texture& t, RGBA c) {
    const SrgbTables& s = srgbTables();
    uint32_t r, g, b;
    if (t.srgb) { r = s.encode(c.r); g = s.encode(c.g); b = s.encode(c.b); }
    else { r = unormEncode(c.r); g = unormEncode(c.g); b = unormEncode(c.b); }
    return r | (g << 8) | (b << 16) | (unormEncode(c.a) << 24);
}

inline float lerp1(float a, float b, float w) { return a * (1.0f - w) + b * w; }
inline RGBA lerp4(RGBA a, RGBA b, float w) { return RGBA{lerp1(a.r, b.r, w), lerp1(a.g, b.g, w), lerp1(a.b, b.b, w), lerp1(a.a, b.a, w)}; }

// Integer texel coordinate -> [0, size) (Vulkan "Wrapping Operation")
inline int wrapCoord(int i, int size, uint32_t mode) {
    if (mode == CLAMP) return std::min(std::max(i, 0), size - 1);
    if (mode == REPEAT) { int m = i % size; return m < 0 ? m + size : m; }
    int p = 2 * size, m = i % p; if (m < 0) m += p; // MIRROR: (size - 1) - mirror((i mod 2 size) - size), mirror(n) = n >= 0 ? n : -(1 + n)
    int n = m - size; n = n >= 0 ? n : -(1 + n);
    return (size - 1) - n;
}

// floor of a texel-space coordinate as an int; coordinates beyond +-2^30 (and NaN) are pinned so that the conversion is defined
inline float pinnedFloor(float x) { return std::fmin(std::fmax(std::floor(x), -1073741824.0f), 1073741824.0f); }

inline RGBA fetch(const Texture& t, uint32_t level, int x, int y) {
    const int w = int(t.levelWidth(level)), h = int(t.levelHeight(level));
    return decodeTexel(t, t.mip[level][size_t(wrapCoord(y, h, t.wrapT)) * size_t(w) + size_t(wrapCoord(x, w, t.wrapS))]);
}

// One level, normalised coordinates (s, t)
inline RGBA sampleLevel(const Texture& t, uint32_t level, float s, float tt, bool linear) {
    const float w = float(t.levelWidth(level)), h = float(t.levelHeight(level));
    if (!linear) return fetch(t, level, int(pinnedFloor(s * w)), int(pinnedFloor(tt * h)));
    const float u = s * w - 0.5f, v = tt * h - 0.5f;
    const float fu = pinnedFloor(u), fv = pinnedFloor(v);
    const float a = u - fu, b = v - fv;
    const int i0 = int(fu), j0 = int(fv);
    const RGBA top = lerp4(fetch(t, level, i0, j0), fetch(t, level, i0 + 1, j0), a);
    const RGBA bot = lerp4(fetch(t, level, i0, j0 + 1), fetch(t, level, i0 + 1, j0 + 1), a);
    return lerp4(top, bot, b);
}

// textureGrad(sampler2D, uv, dPdx, dPdy)
inline RGBA sampleGrad(const Texture& t, float s, float tt, float dudx, float dvdx, float dudy, float dvdy) {
    const float w = float(t.width), h = float(t.height);
    const float ax = dudx * w, bx = dvdx * h, ay = dudy * w, by = dvdy * h;
    const float rhoX = std::sqrt(ax * ax + bx * bx), rhoY = std::sqrt(ay * ay + by * by);
    const float rho = std::fmax(rhoX, rhoY); // T2 (fmax drops a NaN operand)
    float lambda = 0.0f;
    if (rho > 0.0f) lambda = std::log2(rho); // T3
    if (!(lambda > 0.0f)) return sampleLevel(t, 0, s, tt, t.magLinear);
    const float q = float(t.levels - 1);
    const float d = std::fmin(std::fmin(lambda, float(t.levels)), q); // clamp(lambda, minLod 0, maxLod = levels), then to the last level
    if (!t.mipLinear) {
        const uint32_t level = uint32_t(std::fmin(std::fmax(std::ceil(d + 0.5f) - 1.0f, 0.0f), q));
        return sampleLevel(t, level, s, tt, t.minLinear);
    }
    const float dhi = std::floor(d);
    const uint32_t lhi = uint32_t(dhi), llo = std::min(lhi + 1u, t.levels - 1u);
    const float delta = d - dhi;
    const RGBA c0 = sampleLevel(t, lhi, s, tt, t.minLinear);
    if (delta == 0.0f) return c0;
    return lerp4(c0, sampleLevel(t, llo, s, tt, t.minLinear), delta);
}

// texture(sampler2D, uv) in a ray-tracing stage (T3)
inline RGBA sampleBase(const Texture& t, float s, float tt) { return sampleLevel(t, 0, s, tt, t.magLinear); }

inline uint32_t mipLevels(uint32_t w, uint32_t h) { uint32_t m = std::max(w, h), l = 0; while (m > 1) { m >>= 1; ++l; } return l + 1; }

// T5: one LINEAR blit, whole level -> whole next level, edge texels clamped
inline void blitHalf(const Texture& t, const std::vector<uint32_t>& src, uint32_t sw, uint32_t sh, std::vector<uint32_t>& dst, uint32_t dw, uint32_t dh) {
    dst.resize(size_t(dw) * dh);
    const float scaleU = float(sw) / float(dw), scaleV = float(sh) / float(dh);
    for (uint32_t j = 0; j < dh; ++j)
        for (uint32_t i = 0; i < dw; ++i) {
            const float u = (float(i) + 0.5f) * scaleU - 0.5f, v = (float(j) + 0.5f) * scaleV - 0.5f;
            const float fu = std::floor(u), fv = std::floor(v);
            const float a = u - fu, b = v - fv;
            auto at = [&](int x, int y) {
                x = std::min(std::max(x, 0), int(sw) - 1); y = std::min(std::max(y, 0), int(sh) - 1);
                return decodeTexel(t, src[size_t(y) * sw + size_t(x)]);
            };
            const int i0 = int(fu), j0 = int(fv);
            const RGBA top = lerp4(at(i0, j0), at(i0 + 1, j0), a), bot = lerp4(at(i0, j0 + 1), at(i0 + 1, j0 + 1), a);
            dst[size_t(j) * dw + i] = encodeTexel(t, lerp4(top, bot, b));
        }
}

inline Texture makeTexture(const vkx_texture& d) {
    Texture t;
    t.width = d.width; t.height = d.height; t.levels = mipLevels(d.width, d.height); t.srgb = d.srgb != 0;
    const uint32_t mag = d.magFilter ? d.magFilter : 9729u, mn = d.minFilter ? d.minFilter : 9729u; // Resources.cpp:88-90 defaults
    t.magLinear = (mag == 9729u || mag == 9987u);                 // glTFToVkFilter, src/Resources.cpp:8-19
    t.minLinear = (mn == 9729u || mn == 9987u);
    t.mipLinear = (mn == 9729u || mn == 9986u || mn == 9987u);    // glTFToVkSamplerMipmapMode, :21-32
    auto wrap = [](uint32_t e) { return e == 33071u ? CLAMP : e == 33648u ? MIRROR : REPEAT; }; // :34-43
    t.wrapS = wrap(d.wrapS); t.wrapT = wrap(d.wrapT);
    t.mip.resize(t.levels);
    t.mip[0].resize(size_t(d.width) * d.height);
    for (size_t i = 0; i < t.mip[0].size(); ++i)
        t.mip[0][i] = uint32_t(d.pixels[4 * i]) | (uint32_t(d.pixels[4 * i + 1]) << 8) | (uint32_t(d.pixels[4 * i + 2]) << 16) | (uint32_t(d.pixels[4 * i + 3]) << 24);
    for (uint32_t l = 1; l < t.levels; ++l) blitHalf(t, t.mip[l - 1], t.levelWidth(l - 1), t.levelHeight(l - 1), t.mip[l], t.levelWidth(l), t.levelHeight(l));
    return t;
}

} // namespace otex
