// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
//
// Minimal GLSL/GLM-style vector math for the shader transliteration. Operation order follows GLM 0.9.9.8
// (the reference's host math library, /root/reference/ext/glm) so that host-side functions (genBasis,
// sphericalRand) are bit-identical to the reference's; tests/test_glm_pin.py checks that against the real GLM.
// Compile with -ffp-contract=off -fno-fast-math.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace ovm {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct vec4 { float x, y, z, w; float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct ivec2 { int x, y; };
struct ivec3 { int x, y, z; int operator[](int i) const { return (&x)[i]; } };

inline vec2 V2(float x, float y) { return vec2{x, y}; }
inline vec3 V3(float x, float y, float z) { return vec3{x, y, z}; }
inline vec3 V3(float s) { return vec3{s, s, s}; }
inline vec4 V4(float x, float y, float z, float w) { return vec4{x, y, z, w}; }
inline vec4 V4(vec3 v, float w) { return vec4{v.x, v.y, v.z, w}; }
inline vec3 xyz(vec4 v) { return vec3{v.x, v.y, v.z}; }

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, vec2 b) { return {a.x * b.x, a.y * b.y}; }
inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline vec2 operator*(float s, vec2 a) { return {a.x * s, a.y * s}; }
inline vec2 operator/(vec2 a, vec2 b) { return {a.x / b.x, a.y / b.y}; }
inline vec2 operator/(vec2 a, float s) { return {a.x / s, a.y / s}; }
inline vec2 operator+(vec2 a, float s) { return {a.x + s, a.y + s}; }
inline vec2 operator-(vec2 a, float s) { return {a.x - s, a.y - s}; }

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator-(float s, vec3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, vec3 b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator-(vec4 a, vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline vec4 operator*(vec4 a, vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(float s, vec4 a) { return {s * a.x, s * a.y, s * a.z, s * a.w}; }
inline vec4 operator/(vec4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
inline vec4& operator+=(vec4& a, vec4 b) { a = a + b; return a; }

// glm::compute_dot: vec3 -> tmp.x + tmp.y + tmp.z ; vec4 -> (x + y) + (z + w)
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline float length(vec2 v) { return std::sqrt(dot(v, v)); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(vec3 v) { return v * inversesqrt(dot(v, v)); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, vec3 a) { return x * (1.0f - a) + y * a; }
inline float clampf(float x, float lo, float hi) { return std::min(std::max(x, lo), hi); }
inline vec3 clamp(vec3 v, vec3 lo, vec3 hi) { return {clampf(v.x, lo.x, hi.x), clampf(v.y, lo.y, hi.y), clampf(v.z, lo.z, hi.z)}; }
inline vec3 abs(vec3 v) { return {std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)}; }
inline vec2 abs(vec2 v) { return {std::fabs(v.x), std::fabs(v.y)}; }
inline float signf(float x) { return float((0.0f < x) - (x < 0.0f)); } // glm::sign / GLSL sign: 0 for 0
inline vec3 sign(vec3 v) { return {signf(v.x), signf(v.y), signf(v.z)}; }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 min(vec3 a, vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 max(vec3 a, vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline vec3 sqrt(vec3 v) { return {std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)}; }
inline vec3 exp(vec3 v) { return {std::exp(v.x), std::exp(v.y), std::exp(v.z)}; }

// Column-major matrices (as glm): m[col][row].
struct mat3 { vec3 c[3]; vec3& operator[](int i) { return c[i]; } const vec3& operator[](int i) const { return c[i]; } };
struct mat4 { vec4 c[4]; vec4& operator[](int i) { return c[i]; } const vec4& operator[](int i) const { return c[i]; } };

// glm mat3 * vec3: m[0][0]*v.x + m[1][0]*v.y + m[2][0]*v.z
inline vec3 operator*(const mat3& m, vec3 v) {
    return {m[0].x * v.x + m[1].x * v.y + m[2].x * v.z, m[0].y * v.x + m[1].y * v.y + m[2].y * v.z,
            m[0].z * v.x + m[1].z * v.y + m[2].z * v.z};
}
// glm mat4 * vec4: (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
inline vec4 operator*(const mat4& m, vec4 v) {
    vec4 a0 = (m[0] * v.x) + (m[1] * v.y);
    vec4 a1 = (m[2] * v.z) + (m[3] * v.w);
    return a0 + a1;
}
inline mat3 transpose(const mat3& m) {
    mat3 r;
    r[0] = {m[0].x, m[1].x, m[2].x};
    r[1] = {m[0].y, m[1].y, m[2].y};
    r[2] = {m[0].z, m[1].z, m[2].z};
    return r;
}
inline mat3 mat3_from_mat4(const float* m16) { // GLSL mat3(mat4): upper-left 3x3, column-major
    mat3 r;
    for (int c = 0; c < 3; ++c) r[c] = {m16[4 * c + 0], m16[4 * c + 1], m16[4 * c + 2]};
    return r;
}
inline mat4 mat4_from(const float* m16) {
    mat4 r;
    for (int c = 0; c < 4; ++c) r[c] = {m16[4 * c + 0], m16[4 * c + 1], m16[4 * c + 2], m16[4 * c + 3]};
    return r;
}

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

} // namespace ovm
