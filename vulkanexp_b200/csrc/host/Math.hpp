// Small column-major vector/matrix types for the host facade (the reference uses GLM; operation order follows
// glm 0.9.9.8 where results feed the device: mat4 * mat4, mat4 * vec4, dot, normalize, cross).
#pragma once
#include <algorithm>
#include <cmath>

namespace vkx {

struct vec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    vec3() = default;
    vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct ivec3 { int x = 0, y = 0, z = 0; int& operator[](int i) { return (&x)[i]; } int operator[](int i) const { return (&x)[i]; } };
struct vec4 { float x = 0.f, y = 0.f, z = 0.f, w = 0.f; };

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline vec3 min(vec3 a, vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 max(vec3 a, vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }

struct mat4 { // m[col][row], as glm::mat4
    float m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    float* operator[](int c) { return m[c]; }
    const float* operator[](int c) const { return m[c]; }
};
inline mat4 operator*(const mat4& a, const mat4& b) { // glm: col_j = ((A0*b0j + A1*b1j) + A2*b2j) + A3*b3j
    mat4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) r.m[j][i] = ((a.m[0][i] * b.m[j][0] + a.m[1][i] * b.m[j][1]) + a.m[2][i] * b.m[j][2]) + a.m[3][i] * b.m[j][3];
    return r;
}
inline vec4 operator*(const mat4& m, vec4 v) { // glm: (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
    vec4 r;
    float* o = &r.x;
    for (int i = 0; i < 4; ++i) o[i] = (m.m[0][i] * v.x + m.m[1][i] * v.y) + (m.m[2][i] * v.z + m.m[3][i] * v.w);
    return r;
}

struct Bounds { // reference src/Bounds.hpp
    vec3 min, max;
    Bounds& operator+=(const Bounds& o) { min = vkx::min(min, o.min); max = vkx::max(max, o.max); return *this; }
};
inline Bounds operator*(const mat4& t, const Bounds& b) { // two-corner transform, src/Bounds.hpp:38-45
    vec4 a = t * vec4{b.min.x, b.min.y, b.min.z, 1.0f}, c = t * vec4{b.max.x, b.max.y, b.max.z, 1.0f};
    vec3 p{a.x, a.y, a.z}, q{c.x, c.y, c.z};
    return {vkx::min(p, q), vkx::max(p, q)};
}

} // namespace vkx
