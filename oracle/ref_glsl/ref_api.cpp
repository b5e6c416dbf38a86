// TEST INFRASTRUCTURE — C entry points of oracle/_ref/libglslref.so: the reference's own shader functions, compiled from
// the files where they lie under /root/reference/src/shaders (after prep.py's literal-suffix transform, into oracle/_ref/gen/).
// Used only by tests/test_glsl_ref_pin.py and tools/gen_golden_glsl.py to pin oracle/ against the reference's text.
#include "shim.h"

namespace glslref {
FetchFn gFetch = nullptr;
const void* gFetchUser = nullptr;
const uint32_t* Probes = nullptr;

#include "gen/irradiance.glsl"           // includes common.glsl, ProbeGrid.glsl
#include "gen/sky.glsl"
#include "gen/pbrMetallicRoughness.glsl"
namespace filter {
#include "gen/gaussian.glsl"             // directLightFilter.glsl:29-31
}
namespace reflfilter {
#include "gen/gaussian_refl.glsl"        // reflectionFilter.glsl gaussian (same text, kept separately pinned)
}
} // namespace glslref

using namespace glslref;

extern "C" {
struct RefGrid { float extentMin[3]; float depthSharpness; float extentMax[3]; float hysteresis; int resolution[3]; unsigned raysPerProbe, colorRes, depthRes; float shadowBias; unsigned pad; };
static ProbeGrid toGrid(const RefGrid* g) {
    ProbeGrid r;
    r.extentMin = vec3(g->extentMin[0], g->extentMin[1], g->extentMin[2]); r.depthSharpness = g->depthSharpness;
    r.extentMax = vec3(g->extentMax[0], g->extentMax[1], g->extentMax[2]); r.hysteresis = g->hysteresis;
    r.resolution = ivec3(g->resolution[0], g->resolution[1], g->resolution[2]); r.raysPerProbe = g->raysPerProbe;
    r.colorRes = g->colorRes; r.depthRes = g->depthRes; r.shadowBias = g->shadowBias; r.padding[0] = 0;
    return r;
}
void ref_set_fetch(void* fn, const void* user) { gFetch = reinterpret_cast<FetchFn>(fn); gFetchUser = user; }
void ref_set_probes(const uint32_t* states) { Probes = states; }

void ref_sky(const float* o, const float* d, const float* sun, const float* sunColor, float brightness, int showSun, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        vec3 c = sky(vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), vec3(sun[0], sun[1], sun[2]), vec3(sunColor[0], sunColor[1], sunColor[2]), brightness, showSun != 0);
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
void ref_sample_probes(const RefGrid* g, const float* pos, const float* nrm, const float* toCam, size_t n, float* out) {
    const ProbeGrid grid = toGrid(g);
    sampler2D colorTex{0}, depthTex{1};
    for (size_t i = 0; i < n; ++i) {
        vec3 c = sampleProbes(vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), vec3(toCam[3 * i], toCam[3 * i + 1], toCam[3 * i + 2]), grid, colorTex, depthTex);
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
void ref_pbr(const float* nrm, const float* view, const float* lightColor, const float* lightDir, const float* albedo, const float* metalRough, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        vec4 c = pbrMetallicRoughness(vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), vec3(view[3 * i], view[3 * i + 1], view[3 * i + 2]), vec3(lightColor[0], lightColor[1], lightColor[2]),
                                      vec3(lightDir[0], lightDir[1], lightDir[2]), vec4(albedo[4 * i], albedo[4 * i + 1], albedo[4 * i + 2], albedo[4 * i + 3]), metalRough[2 * i], metalRough[2 * i + 1]);
        out[4 * i] = c.x; out[4 * i + 1] = c.y; out[4 * i + 2] = c.z; out[4 * i + 3] = c.w;
    }
}
void ref_spherical_fibonacci(const float* i_, float nn, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { vec3 v = sphericalFibonacci(i_[i], nn); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
void ref_oct_decode(const float* o, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { vec3 v = octDecode(vec2(o[2 * i], o[2 * i + 1])); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
void ref_oct_encode(const float* d, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { vec2 v = octEncode(vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2])); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
void ref_sphere_to_oct_uv(const float* d, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { vec2 v = spherePointToOctohedralUV(vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2])); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
void ref_rotate_axis(const float* p, const float* axis, const float* angle, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { vec3 v = rotateAxis(vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), vec3(axis[3 * i], axis[3 * i + 1], axis[3 * i + 2]), angle[i]); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
void ref_gaussian(const float* stdDev, const float* dist, size_t n, float* out) { for (size_t i = 0; i < n; ++i) out[i] = filter::gaussian(stdDev[i], dist[i]); }
void ref_gaussian_refl(const float* stdDev, const float* dist, size_t n, float* out) { for (size_t i = 0; i < n; ++i) out[i] = reflfilter::gaussian(stdDev[i], dist[i]); }
// irradiance.glsl:7-38 index helpers: out = (gridIndex xyz, linear index round trip, colourUV xy, depthUV xy) + world position
void ref_probe_helpers(const RefGrid* g, const uint32_t* index, size_t n, int* outI, float* outF) {
    const ProbeGrid grid = toGrid(g);
    for (size_t i = 0; i < n; ++i) {
        const ivec3 gi = probeLinearIndexToGridIndex(index[i], grid);
        const ivec2 cu = probeIndexToColorUVOffset(gi, grid), du = probeIndexToDepthUVOffset(gi, grid);
        int* o = outI + 8 * i; o[0] = gi.x; o[1] = gi.y; o[2] = gi.z; o[3] = int(probeLinearIndex(gi, grid)); o[4] = cu.x; o[5] = cu.y; o[6] = du.x; o[7] = du.y;
        const vec3 w = probeIndexToWorldPosition(index[i], grid), cs = probeGridCellSize(grid);
        float* f = outF + 6 * i; f[0] = w.x; f[1] = w.y; f[2] = w.z; f[3] = cs.x; f[4] = cs.y; f[5] = cs.z;
    }
}
void ref_normalize_local_texel(const int* coord, unsigned res, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { vec2 v = normalizeLocalTexelCoord(ivec2(coord[2 * i], coord[2 * i + 1]), res); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
}
