// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
//
// Atlas texel formats of the reference (src/IrradianceProbes.cpp:40-41,62-63):
//   irradiance: VK_FORMAT_B10G11R11_UFLOAT_PACK32  (R bits 0..10 [5e6m], G bits 11..21 [5e6m], B bits 22..31 [5e5m])
//   depth:      VK_FORMAT_R16G16_SFLOAT            (R low half, G high half)
// Vulkan leaves the float->small-float rounding to the implementation; SURVEY section 7 (hard part 4) decrees
// round-to-nearest-even, saturate to the largest finite value, NaN / negative -> 0. Pure integer code so that the
// CUDA kernels can reproduce it bit for bit.
#pragma once
#include <cstdint>
#include <cstring>

namespace opack {

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Unsigned small float with 5 exponent bits (bias 15) and MB mantissa bits.
template <int MB>
inline uint32_t packUF(float f) {
    if (!(f > 0.0f)) return 0; // negative, zero, NaN
    const uint32_t maxCode = (31u << MB) - 1u;
    uint32_t b = f2u(f);
    int e = int(b >> 23) - 127;
    uint32_t m = b & 0x7FFFFFu;
    if (e > 15) return maxCode; // includes +inf
    uint32_t code;
    if (e >= -14) {
        const int sh = 23 - MB;
        uint32_t q = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        code = (uint32_t(e + 15) << MB) + q;
        if (rem > half || (rem == half && (q & 1u))) code += 1; // carry may bump the exponent: intended
    } else {
        // denormal target: value = mant * 2^(-14-MB)
        int sh = (23 - MB) + (-14 - e);
        if (sh > 24) return 0;
        uint32_t full = m | 0x800000u;
        uint32_t q = full >> sh, rem = full & ((1u << sh) - 1u), half = 1u << (sh - 1);
        code = q;
        if (rem > half || (rem == half && (q & 1u))) code += 1;
    }
    return code > maxCode ? maxCode : code;
}

template <int MB>
inline float unpackUF(uint32_t c) {
    uint32_t e = c >> MB, m = c & ((1u << MB) - 1u);
    if (e == 0) return float(m) * u2f(uint32_t(127 - 14 - MB) << 23);
    if (e == 31) return m ? u2f(0x7FC00000u) : u2f(0x7F800000u);
    return u2f(((e + 112u) << 23) | (m << (23 - MB)));
}

inline uint32_t packR11G11B10(float r, float g, float b) { return packUF<6>(r) | (packUF<6>(g) << 11) | (packUF<5>(b) << 22); }
inline void unpackR11G11B10(uint32_t p, float out[3]) {
    out[0] = unpackUF<6>(p & 0x7FFu);
    out[1] = unpackUF<6>((p >> 11) & 0x7FFu);
    out[2] = unpackUF<5>(p >> 22);
}

// IEEE binary16, round-to-nearest-even, overflow -> inf, NaN -> quiet NaN (same as CUDA __float2half_rn).
inline uint16_t packHalf(float f) {
    uint32_t b = f2u(f);
    uint32_t sign = (b >> 16) & 0x8000u;
    uint32_t a = b & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return uint16_t(sign | 0x7FFFu); // NaN (CUDA returns 0x7FFF)
    if (a >= 0x47800000u) return uint16_t(sign | 0x7C00u); // >= 65536 -> inf (values in [65520,65536) round to inf below)
    int e = int(a >> 23) - 127;
    uint32_t m = a & 0x7FFFFFu;
    uint32_t code;
    if (e >= -14) {
        uint32_t q = m >> 13, rem = m & 0x1FFFu;
        code = (uint32_t(e + 15) << 10) + q;
        if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) code += 1;
    } else {
        int sh = 13 + (-14 - e);
        if (sh > 24) return uint16_t(sign);
        uint32_t full = m | 0x800000u;
        uint32_t q = full >> sh, rem = full & ((1u << sh) - 1u), half = 1u << (sh - 1);
        code = q;
        if (rem > half || (rem == half && (q & 1u))) code += 1;
    }
    return uint16_t(sign | code);
}
inline float unpackHalf(uint16_t h) {
    uint32_t sign = uint32_t(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 31u, m = h & 0x3FFu;
    if (e == 0) { float v = float(m) * u2f(uint32_t(127 - 24) << 23); return (sign ? -v : v); }
    if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}
inline uint32_t packRG16F(float r, float g) { return uint32_t(packHalf(r)) | (uint32_t(packHalf(g)) << 16); }
inline void unpackRG16F(uint32_t p, float out[2]) { out[0] = unpackHalf(uint16_t(p & 0xFFFFu)); out[1] = unpackHalf(uint16_t(p >> 16)); }

} // namespace opack
