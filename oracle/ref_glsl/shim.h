// TEST INFRASTRUCTURE — GLSL-as-C++ shim for oracle/_ref. Lets the reference's pure shader functions
// (src/shaders/{common,ProbeGrid,irradiance,sky,pbrMetallicRoughness}.glsl and gaussian() of directLightFilter.glsl)
// compile as C++ against the reference's own vendored GLM (ext/glm, 0.9.9.8), so that the arithmetic that runs is the
// reference's text, not a restatement. Nothing here re-implements shader logic: the shim only supplies what the GLSL
// language provides implicitly (implicit int -> float conversions in mixed expressions, the `in` qualifier, swizzles on
// r-values, sampler objects and the `Probes` storage buffer).
#pragma once
#define GLM_FORCE_SWIZZLE
#ifndef _MSC_EXTENSIONS
#define _MSC_EXTENSIONS 1 // GLM only checks this to allow anonymous structs (swizzle operators v.xy as in GLSL); g++ supports them
#endif
#define GLM_FORCE_PURE          // no SIMD code paths: plain scalar IEEE operations in GLM's documented order
#define GLM_FORCE_SILENT_WARNINGS
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>

namespace glslref {
using namespace glm;
typedef unsigned int uint;

#define in
#define out_param &

// ---- implicit int -> float conversions of GLSL in mixed scalar expressions
inline float sqrt(int x) { return std::sqrt(float(x)); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float clamp(float x, int lo, int hi) { return glm::clamp(x, float(lo), float(hi)); }
inline float clamp(float x, int lo, float hi) { return glm::clamp(x, float(lo), hi); }
inline float clamp(float x, float lo, float hi) { return glm::clamp(x, lo, hi); }
inline float max(float a, float b) { return glm::max(a, b); }
inline float min(float a, float b) { return glm::min(a, b); }
inline float pow(float a, float b) { return glm::pow(a, b); }
inline float exp(float a) { return glm::exp(a); }
inline float cos(float a) { return glm::cos(a); }
inline float sin(float a) { return glm::sin(a); }
inline float acos(float a) { return glm::acos(a); }
inline float atan(float a) { return glm::atan(a); }
inline float floor(float a) { return glm::floor(a); }
inline float abs(float a) { return glm::abs(a); }
// swizzle proxies as function arguments (GLSL: abs(v.yx), signNotZero(v.xy), vec2 r = v.xy * s)
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> abs(glm::detail::_swizzle<N, T, Q, E0, E1, E2, E3> const& s) { return glm::abs(s()); }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> mix(vec<N, T, Q> const& a, glm::detail::_swizzle<N, T, Q, E0, E1, E2, E3> const& b, T t) { return glm::mix(a, b(), t); }
// (vector arguments reach glm::abs / sqrt / exp / clamp / mix ... through argument-dependent lookup)

// ---- mixed int / float vector arithmetic (GLSL converts the integer operand)
inline vec3 operator*(ivec3 const& a, vec3 const& b) { return vec3(a) * b; }
inline vec3 operator/(vec3 const& a, ivec3 const& b) { return a / vec3(b); }
inline vec2 operator/(vec2 const& a, uint b) { return a / float(b); }
inline ivec2 operator*(uint a, ivec2 const& b) { return ivec2(int(a) * b.x, int(a) * b.y); } // uint * ivec2 -> (GLSL: uvec2) -> ivec2(...) at the call site
inline vec2 operator*(vec2 const& a, int b) { return a * float(b); }

// ---- resources the shader text refers to
struct sampler2D { int kind; }; // 0: irradiance atlas, 1: depth atlas
typedef void (*FetchFn)(const void* user, int kind, float u, float v, float out4[4]);
extern FetchFn gFetch;        // fixed-function bilinear fetch (not shader code): supplied by the caller
extern const void* gFetchUser;
inline vec4 textureLod(sampler2D s, vec2 uv, int /*lod*/) { float o[4] = {0, 0, 0, 0}; gFetch(gFetchUser, s.kind, uv.x, uv.y, o); return vec4(o[0], o[1], o[2], o[3]); }
extern const uint32_t* Probes; // layout(binding = 13) buffer ProbesBlock { uint Probes[]; }

} // namespace glslref
