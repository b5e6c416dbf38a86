// Host logic of the reference's IrradianceProbes.cpp that the device never sees: the per-frame random orientation
// (reference src/IrradianceProbes.cpp:347-355, 455-460: glm::sphericalRand + genBasis, unseeded MSVC rand()) and the
// probe scheduler (src/IrradianceProbes.cpp:396-424). Exported through the C ABI so non-C++ hosts reuse them.
#include <cmath>
#include <cstdint>
#include <limits>
#include "../../../include/vkx.h"
#include "Math.hpp"

namespace {
inline int msvcRand(uint32_t& x) { x = x * 214013u + 2531011u; return int((x >> 16) & 0x7FFFu); }
// glm::detail::compute_rand<1, uint32> (ext/glm/glm/gtc/random.inl:19-85): four rand() % 255 bytes; the first draw ends up
// in the lowest byte with the operand evaluation order g++ uses (MSVC's is unverifiable here).
inline uint32_t randU32(uint32_t& x) {
    uint32_t b0 = uint32_t(msvcRand(x) % 255), b1 = uint32_t(msvcRand(x) % 255), b2 = uint32_t(msvcRand(x) % 255), b3 = uint32_t(msvcRand(x) % 255);
    return (b3 << 24) | (b2 << 16) | (b1 << 8) | b0;
}
inline float linearRand(uint32_t& x, float lo, float hi) { return float(randU32(x)) / float(std::numeric_limits<uint32_t>::max()) * (hi - lo) + lo; }
} // namespace

extern "C" {

void vkx_host_next_orientation(uint32_t* rngState, float orientation[16]) {
    using namespace vkx;
    const float theta = linearRand(*rngState, 0.0f, 6.283185307179586476925286766559f);
    const float phi = std::acos(linearRand(*rngState, -1.0f, 1.0f));
    vec3 Z{std::sin(phi) * std::cos(theta), std::sin(phi) * std::sin(theta), std::cos(phi)};
    Z = Z * 1.0f;
    vec3 b1 = Z.x > 0.9f ? vec3{0.0f, 1.0f, 0.0f} : vec3{1.0f, 0.0f, 0.0f}; // genBasis
    b1 = b1 - Z * dot(b1, Z);
    b1 = normalize(b1);
    const vec3 b2 = cross(Z, b1);
    // mat4(transpose(mat3(X, Y, Z))), column-major
    const vec3 cols[3] = {b1, b2, Z};
    for (int i = 0; i < 16; ++i) orientation[i] = 0.0f;
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) orientation[4 * c + r] = cols[r][c];
    orientation[15] = 1.0f;
}

uint32_t vkx_host_select_probes(uint32_t* loopIndex, uint32_t* lastUpdateOffset, const uint32_t* state, uint32_t probeCount, uint32_t probesPerUpdate, uint32_t* out) {
    uint32_t n = 0, idx = *lastUpdateOffset, checked = 0;
    if (idx >= probeCount) idx = 0;
    while (checked < probeCount && (probesPerUpdate == 0 || n < probesPerUpdate)) {
        if (state[idx] != 0 && ((idx + *loopIndex) % state[idx]) == 0) out[n++] = idx;
        ++idx;
        if (idx >= probeCount) { idx = 0; ++*loopIndex; }
        ++checked;
    }
    *lastUpdateOffset = idx;
    return n;
}

} // extern "C"
