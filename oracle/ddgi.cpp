// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
//
// CPU transliteration of the reference's DDGI probe update (all citations relative to /root/reference):
//   src/shaders/traceProbes.rgen, closesthit.glsl (NO_REFLECTION), irradiance.glsl, miss.rmiss, sky.glsl,
//   shadow.rmiss, pbrMetallicRoughness.glsl, probesUpdate.glsl, probesCopyBorders.comp, probesInit.rgen,
//   backfaceTest.rchit, probeInitMiss.rmiss and the host logic of src/IrradianceProbes.cpp.
// PARITY UNPINNED by the reference: VulkanExp ships no tests, golden images or known-answer vectors for this
// path. What *is* pinned: the border tables (tests/golden/border_tables.json, extracted from
// probesCopyBorders.comp), the POD layouts, and glm::sphericalRand/genBasis (tests/golden/glm_pin.json, produced by
// compiling the reference's vendored GLM). Undefined behaviour in the shaders is resolved as SURVEY A.5 decrees.
// Textures: closesthit.glsl:161-192 through "sampler spec v1" (texture.h; SURVEY A.8: filtering is implementation-defined in the
// reference, the decrees are listed there).
#include "ddgi.h"
#include "packing.h"
#include <cmath>
#include <cstring>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ovm;

namespace oddgi {

static const float pi = 3.1415926538f; // common.glsl:4

// ---------------------------------------------------------------- irradiance.glsl
static inline ivec3 probeLinearIndexToGridIndex(uint32_t index, const vkx_grid_info& g) { // irradiance.glsl:7-13
    uint32_t rx = uint32_t(g.resolution[0]), ry = uint32_t(g.resolution[1]);
    return ivec3{int(index % rx), int((index % (rx * ry)) / rx), int(index / (rx * ry))};
}
static inline uint32_t probeLinearIndex(ivec3 i, const vkx_grid_info& g) { // :15-17
    return uint32_t(i.x + g.resolution[0] * i.y + g.resolution[0] * g.resolution[1] * i.z);
}
static inline vec3 probeGridCellSize(const vkx_grid_info& g) { // :19-21
    vec3 mx = V3(g.extentMax[0], g.extentMax[1], g.extentMax[2]), mn = V3(g.extentMin[0], g.extentMin[1], g.extentMin[2]);
    return (mx - mn) / V3(float(g.resolution[0] - 1), float(g.resolution[1] - 1), float(g.resolution[2] - 1));
}
static inline vec3 probeIndexToWorldPosition(ivec3 i, const vkx_grid_info& g) { // :23-26
    return V3(float(i.x), float(i.y), float(i.z)) * probeGridCellSize(g) + V3(g.extentMin[0], g.extentMin[1], g.extentMin[2]);
}
static inline ivec2 probeIndexToColorUVOffset(ivec3 i, const vkx_grid_info& g) { // :32-34
    return ivec2{int(g.colorRes) * (i.y * g.resolution[0] + i.x), int(g.colorRes) * i.z};
}
static inline ivec2 probeIndexToDepthUVOffset(ivec3 i, const vkx_grid_info& g) { // :36-38
    return ivec2{int(g.depthRes) * (i.y * g.resolution[0] + i.x), int(g.depthRes) * i.z};
}

void probeHelpers(const vkx_grid_info& g, uint32_t index, int outI[8], float outF[6]) { // irradiance.glsl:7-38, every helper on one index (pin hook)
    const ivec3 gi = probeLinearIndexToGridIndex(index, g);
    const ivec2 cu = probeIndexToColorUVOffset(gi, g), du = probeIndexToDepthUVOffset(gi, g);
    outI[0] = gi.x; outI[1] = gi.y; outI[2] = gi.z; outI[3] = int(probeLinearIndex(gi, g)); outI[4] = cu.x; outI[5] = cu.y; outI[6] = du.x; outI[7] = du.y;
    const vec3 w = probeIndexToWorldPosition(gi, g), cs = probeGridCellSize(g);
    outF[0] = w.x; outF[1] = w.y; outF[2] = w.z; outF[3] = cs.x; outF[4] = cs.y; outF[5] = cs.z;
}
vec3 sphericalFibonacci(float i, float n) { // irradiance.glsl:52-64
    const float PHI = std::sqrt(5.0f) * 0.5f + 0.5f;
    float ab = i * (PHI - 1.0f);
    float phi = 2.0f * pi * (ab - std::floor(ab));
    float cosTheta = 1.0f - (2.0f * i + 1.0f) * (1.0f / n);
    float sinTheta = std::sqrt(clampf(1.0f - cosTheta * cosTheta, 0.0f, 1.0f));
    return V3(std::cos(phi) * sinTheta, std::sin(phi) * sinTheta, cosTheta);
}

static inline float signNotZero(float k) { return (k >= 0.0f) ? 1.0f : -1.0f; } // :88-90

vec2 octEncode(vec3 v) { // :97-103 (not called on the hot path; restated so that row a12 is pinned as a whole)
    float l1norm = std::fabs(v.x) + std::fabs(v.y) + std::fabs(v.z);
    vec2 result = V2(v.x, v.y) * (1.0f / l1norm);
    if (v.z < 0.0f) result = V2((1.0f - std::fabs(result.y)) * signNotZero(result.x), (1.0f - std::fabs(result.x)) * signNotZero(result.y));
    return result;
}
vec3 octDecode(vec2 o) { // :107-112
    vec3 v = V3(o.x, o.y, 1.0f - std::fabs(o.x) - std::fabs(o.y));
    if (v.z < 0.0f) {
        float nx = (1.0f - std::fabs(v.y)) * signNotZero(v.x);
        float ny = (1.0f - std::fabs(v.x)) * signNotZero(v.y);
        v.x = nx; v.y = ny;
    }
    return normalize(v);
}

vec2 spherePointToOctohedralUV(vec3 direction) { // :119-138 (sign() is 0 for 0)
    vec3 octant = sign(direction);
    float sum = dot(direction, octant);
    vec3 octahedron = direction / sum;
    if (octahedron.z < 0.0f) {
        vec3 absolute = abs(octahedron);
        octahedron.x = octant.x * (1.0f - absolute.y);
        octahedron.y = octant.y * (1.0f - absolute.x);
    }
    return V2(octahedron.x * 0.5f + 0.5f, octahedron.y * 0.5f + 0.5f);
}

// textureLod(sampler2D, uv, 0): LINEAR filter, REPEAT addressing (src/RaytracingDescriptors.hpp:63,70), exact fp32 weights
static inline void bilinearSetup(float u, float size, int& i0, int& i1, float& f) {
    float x = u * size - 0.5f;
    float fl = std::floor(x);
    f = x - fl;
    int isz = int(size);
    int i = int(fl);
    i0 = ((i % isz) + isz) % isz;
    i1 = (i0 + 1) % isz;
}
vec3 sampleIrradiance(const Probes& p, vec2 uv) {
    int x0, x1, y0, y1; float fx, fy;
    bilinearSetup(uv.x, float(p.irrW), x0, x1, fx);
    bilinearSetup(uv.y, float(p.irrH), y0, y1, fy);
    float t[4][3];
    opack::unpackR11G11B10(p.irrSampled[size_t(y0) * p.irrW + x0], t[0]);
    opack::unpackR11G11B10(p.irrSampled[size_t(y0) * p.irrW + x1], t[1]);
    opack::unpackR11G11B10(p.irrSampled[size_t(y1) * p.irrW + x0], t[2]);
    opack::unpackR11G11B10(p.irrSampled[size_t(y1) * p.irrW + x1], t[3]);
    vec3 r;
    for (int c = 0; c < 3; ++c) {
        float top = t[0][c] * (1.0f - fx) + t[1][c] * fx;
        float bot = t[2][c] * (1.0f - fx) + t[3][c] * fx;
        r[c] = top * (1.0f - fy) + bot * fy;
    }
    return r;
}
vec2 sampleDepth(const Probes& p, vec2 uv) {
    int x0, x1, y0, y1; float fx, fy;
    bilinearSetup(uv.x, float(p.depW), x0, x1, fx);
    bilinearSetup(uv.y, float(p.depH), y0, y1, fy);
    float t[4][2];
    opack::unpackRG16F(p.depSampled[size_t(y0) * p.depW + x0], t[0]);
    opack::unpackRG16F(p.depSampled[size_t(y0) * p.depW + x1], t[1]);
    opack::unpackRG16F(p.depSampled[size_t(y1) * p.depW + x0], t[2]);
    opack::unpackRG16F(p.depSampled[size_t(y1) * p.depW + x1], t[3]);
    float r[2];
    for (int c = 0; c < 2; ++c) {
        float top = t[0][c] * (1.0f - fx) + t[1][c] * fx;
        float bot = t[2][c] * (1.0f - fx) + t[3][c] * fx;
        r[c] = top * (1.0f - fy) + bot * fy;
    }
    return V2(r[0], r[1]);
}

vec3 sampleProbes(const Probes& p, vec3 position, vec3 normal, vec3 toCamera) { // irradiance.glsl:145-237
    const vkx_grid_info& grid = p.grid;
    vec3 gridCellSize = probeGridCellSize(grid);
    vec3 extentMin = V3(grid.extentMin[0], grid.extentMin[1], grid.extentMin[2]);
    vec3 gridCoords = (position - extentMin) / abs(gridCellSize);
    if (gridCoords.x < 0.0f || gridCoords.y < 0.0f || gridCoords.z < 0.0f) return V3(0.0f);
    vec3 biasVector = (normal + toCamera) * grid.shadowBias;
    vec3 biasedPosition = position + biasVector;
    ivec3 firstProbeIdx = ivec3{int(gridCoords.x), int(gridCoords.y), int(gridCoords.z)};
    vec3 alpha = clamp((position - probeIndexToWorldPosition(firstProbeIdx, grid)) / abs(gridCellSize), V3(0.0f), V3(1.0f));

    vec3 finalColor = V3(0.0f);
    float totalWeight = 0.0f;
    vec3 fallbackColor = V3(0.0f);
    float totalFallbackWeight = 0.0f;
    vec2 uvScaling = V2(float(grid.resolution[0] * grid.resolution[1]), float(grid.resolution[2]));

    for (int i = 0; i < 8; ++i) {
        ivec3 offset = ivec3{i & 1, (i >> 1) & 1, (i >> 2) & 1};
        ivec3 probeCoords = ivec3{firstProbeIdx.x + offset.x, firstProbeIdx.y + offset.y, firstProbeIdx.z + offset.z};
        if (probeCoords.x > grid.resolution[0] - 1 || probeCoords.y > grid.resolution[1] - 1 || probeCoords.z > grid.resolution[2] - 1) continue;
        if (p.state[probeLinearIndex(probeCoords, grid)] == 0) continue;
        vec3 probePosition = probeIndexToWorldPosition(probeCoords, grid);
        vec3 directionToProbe = normalize(probePosition - position);
        vec3 biasedDirectionToProbe = probePosition - biasedPosition;
        vec2 localColorUV = (float(grid.colorRes - 2) / float(grid.colorRes)) * spherePointToOctohedralUV(normal);
        vec2 localDepthUV = (float(grid.depthRes - 2) / float(grid.depthRes)) * spherePointToOctohedralUV(-normalize(biasedDirectionToProbe));
        ivec2 co = probeIndexToColorUVOffset(probeCoords, grid), dofs = probeIndexToDepthUVOffset(probeCoords, grid);
        vec2 colorUV = (V2(float(co.x + 1), float(co.y + 1)) / float(grid.colorRes) + localColorUV) / uvScaling;
        vec2 depthUV = (V2(float(dofs.x + 1), float(dofs.y + 1)) / float(grid.depthRes) + localDepthUV) / uvScaling;
        vec3 trilinear = mix(1.0f - alpha, alpha, V3(float(offset.x), float(offset.y), float(offset.z)));
        float weight = 1.0f;

        float backfaceweight = std::max(0.0001f, (dot(directionToProbe, normal) + 1.0f) * 0.5f);
        weight *= backfaceweight * backfaceweight + 0.2f;
        float fallbackWeight = weight;

        vec2 depth = sampleDepth(p, depthUV);
        float mean = depth.x;
        float variance = std::fabs(depth.x * depth.x - depth.y);
        float biasedDistToProbe = length(probePosition - biasedPosition);
        float dd = std::max(biasedDistToProbe - mean, 0.0001f);
        float chebyshevWeight = variance / (variance + dd * dd);
        chebyshevWeight = std::max(std::pow(chebyshevWeight, 3.0f), 0.0f);
        weight *= (biasedDistToProbe <= mean) ? 1.0f : chebyshevWeight;
        weight = std::max(0.000001f, weight);

        const float crushThreshold = 0.2f;
        if (weight < crushThreshold) weight *= weight * weight * (1.0f / (crushThreshold * crushThreshold));

        float tri = trilinear.x * trilinear.y * trilinear.z + 0.001f;
        weight *= tri;
        fallbackWeight *= tri;

        vec3 color = sampleIrradiance(p, colorUV);
        color = sqrt(color);

        finalColor += weight * color;
        totalWeight += weight;
        fallbackColor += fallbackWeight * color;
        totalFallbackWeight += fallbackWeight;
    }
    if (totalWeight > 1e-3f) finalColor *= 1.0f / totalWeight;
    if (totalFallbackWeight > 1e-3f) fallbackColor *= 1.0f / totalFallbackWeight;
    finalColor *= finalColor;
    fallbackColor *= fallbackColor;
    return mix(fallbackColor, finalColor, 8.0f * clampf(totalWeight, 0.0f, 1.0f / 8.0f));
}

// ---------------------------------------------------------------- sky.glsl
namespace skyc {
static const float AvegerageDensityAltitude = 0.25f;
static const float InnerRadius = 100000.0f;
static const float OuterRadius = 2500.0f + InnerRadius;
static const float Scale = 1.0f / (OuterRadius - InnerRadius);
static const uint32_t SampleCount = 64;
static const float Kr = 0.0025f;
static const float Kr4PI = Kr * 4.0f * pi;
static const float Km = 0.0010f;
static const float Km4PI = Km * 4.0f * pi;
static const float g = -0.990f;
}

static float traceSphereOutside(vec3 center, float radius, vec3 origin, vec3 direction) { // sky.glsl:13-29
    vec3 d = origin - center;
    float a = dot(direction, direction);
    float b = dot(direction, d);
    float c = dot(d, d) - radius * radius;
    float gg = b * b - a * c;
    if (gg > 0.0f) {
        float dis = (-std::sqrt(gg) - b) / a;
        if (dis > 0.0f) return dis;
    }
    return -1.0f;
}
static float traceSphereInside(vec3 center, float radius, vec3 origin, vec3 direction) { // :32-41
    vec3 oc = center - origin;
    float docdir = dot(oc, direction);
    vec3 pc = origin + docdir * direction;
    float dist = std::sqrt(radius * radius - length(pc - center) * length(pc - center));
    if (docdir > 0.0f) return dist - length(pc - origin);
    else return dist + length(pc - origin);
}
static float scaleFn(float fCos) { // :43-47
    float x = 1.0f - fCos;
    return skyc::AvegerageDensityAltitude * std::exp(-0.00287f + x * (0.459f + x * (3.83f + x * (-6.80f + x * 5.25f))));
}

vec3 sky(vec3 rayOrigin, vec3 rayDirection, vec3 sunPosition, vec3 sunColor, float sunBrightnessFactor, bool showSun) { // :59-126
    using namespace skyc;
    const vec3 InvWaveLengths = V3(1.0f / std::pow(0.650f, 4.0f), 1.0f / std::pow(0.570f, 4.0f), 1.0f / std::pow(0.475f, 4.0f));
    sunColor *= sunBrightnessFactor;
    const vec3 planetCenter = V3(0.0f, -InnerRadius - 100.0f, 0.0f);
    vec3 position = rayOrigin - planetCenter;
    float height = length(position);
    vec3 lightDir = normalize(sunPosition);
    if (std::fabs(height - InnerRadius) < 1e-3f) {
        position += 1e-2f * normalize(position);
        height = length(position);
    }
    if (height < OuterRadius) {
        if (height > InnerRadius) {
            float planetDistance = traceSphereOutside(V3(0.0f), InnerRadius, position, rayDirection);
            if (planetDistance >= 0.0f) return std::max(0.1f, dot(lightDir, normalize(position + planetDistance * rayDirection))) * V3(0.05f);
        } else {
            return V3(0.0f);
        }
        float rayDepth = traceSphereInside(V3(0.0f), OuterRadius, position, rayDirection);
        if (std::isinf(rayDepth) || std::isnan(rayDepth)) return V3(0.0f);

        float depth = std::exp(Scale / AvegerageDensityAltitude * (InnerRadius - height));
        float startAngle = dot(rayDirection, position) / height;
        float startOffset = depth * scaleFn(startAngle);

        float sampleLength = rayDepth / float(SampleCount);
        float scaledLength = sampleLength * Scale;
        vec3 sampleRay = rayDirection * sampleLength;
        vec3 samplePoint = position + 0.5f * sampleRay;

        vec3 color = V3(0.0f);
        for (uint32_t i = 0; i < SampleCount; ++i) {
            float h = length(samplePoint);
            float dpt = std::exp(Scale / AvegerageDensityAltitude * (InnerRadius - h));
            float lightAngle = dot(lightDir, samplePoint) / h;
            float cameraAngle = dot(rayDirection, samplePoint) / h;
            float scatter = startOffset + dpt * (scaleFn(lightAngle) - scaleFn(cameraAngle));
            vec3 attenuate = exp(-scatter * (InvWaveLengths * Kr4PI + Km4PI));
            if (std::isinf(attenuate.x) || std::isinf(attenuate.y) || std::isinf(attenuate.z) || std::isnan(attenuate.x) || std::isnan(attenuate.y) || std::isnan(attenuate.z))
                continue; // (the reference's `continue` also skips the samplePoint advance)
            color += attenuate * (dpt * scaledLength);
            samplePoint += sampleRay;
        }
        vec3 secondary = color * Km * sunColor;
        color *= InvWaveLengths * Kr * sunColor;
        if (showSun) {
            float miecos = dot(lightDir, -rayDirection);
            float miePhase = 1.5f * ((1.0f - g * g) / (2.0f + g * g)) * (1.0f + miecos * miecos) / std::pow(std::max(1e-3f, 1.0f + g * g - 2.0f * g * miecos), 1.5f);
            color += miePhase * secondary;
        }
        if (!(std::isinf(color.x) || std::isinf(color.y) || std::isinf(color.z))) return color;
    } else {
        float depth = traceSphereOutside(V3(0.0f), OuterRadius, position, rayDirection);
        if (depth > 0.0f) return dot(lightDir, normalize(position + depth * rayDirection)) * 0.5f * V3(0.5294117647f, 0.80784313725f, 0.92156862745f);
    }
    return V3(0.0f);
}

// ---------------------------------------------------------------- pbrMetallicRoughness.glsl
vec4 pbrMetallicRoughness(vec3 normal, vec3 view, vec3 lightColor, vec3 lightDirection, vec4 albedo, float metalness, float roughness) { // :43-84
    vec3 f0 = V3(0.04f);
    vec3 alb = xyz(albedo);
    vec3 diffuseColor = alb * (1.0f - f0);
    diffuseColor *= (1.0f - metalness);
    float alphaRoughness = roughness * roughness;
    vec3 specularColor = mix(f0, alb, metalness);
    float reflectance = std::max(std::max(specularColor.x, specularColor.y), specularColor.z);
    float reflectance90 = clampf(reflectance * 25.0f, 0.0f, 1.0f);
    vec3 R0 = specularColor;
    vec3 R90 = V3(1.0f) * reflectance90;
    vec3 n = normal, v = view;
    vec3 l = normalize(lightDirection);
    vec3 h = normalize(l + v);
    float NdotL = clampf(dot(n, l), 0.001f, 1.0f);
    float NdotV = clampf(std::fabs(dot(n, v)), 0.001f, 1.0f);
    float NdotH = clampf(dot(n, h), 0.0f, 1.0f);
    float VdotH = clampf(dot(v, h), 0.0f, 1.0f);
    vec3 F = R0 + (R90 - R0) * std::pow(clampf(1.0f - VdotH, 0.0f, 1.0f), 5.0f);
    float ar2 = alphaRoughness * alphaRoughness;
    float attenuationL = 2.0f * NdotL / (NdotL + std::sqrt(ar2 + (1.0f - ar2) * (NdotL * NdotL)));
    float attenuationV = 2.0f * NdotV / (NdotV + std::sqrt(ar2 + (1.0f - ar2) * (NdotV * NdotV)));
    float G = attenuationL * attenuationV;
    float a = NdotH * alphaRoughness;
    float k = alphaRoughness / ((1.0f - NdotH * NdotH) + a * a);
    float D = clampf(k * k * (1.0f / pi), 0.0f, 4.0f);
    vec3 diffuseContrib = (1.0f - F) * diffuseColor / pi;
    vec3 specContrib = F * G * D / (4.0f * NdotL * NdotV);
    vec3 color = NdotL * lightColor * (diffuseContrib + specContrib);
    return V4(color, albedo.w);
}

// ---------------------------------------------------------------- scene / probes setup
static mat3 inverse3(const float* M) { // M: row-major 3x4; returns W = inverse(M3) stored so that c[col][row]
    float a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], i = M[10];
    float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    float det = a * A + b * B + c * C;
    float id = 1.0f / det;
    // inverse row-major entries
    float r00 = A * id, r01 = -(b * i - c * h) * id, r02 = (b * f - c * e) * id;
    float r10 = B * id, r11 = (a * i - c * g) * id, r12 = -(a * f - c * d) * id;
    float r20 = C * id, r21 = -(a * h - b * g) * id, r22 = (a * e - b * d) * id;
    mat3 W;
    W[0] = V3(r00, r10, r20);
    W[1] = V3(r01, r11, r21);
    W[2] = V3(r02, r12, r22);
    return W;
}

void sceneFinalize(Scene& s) {
    s.worldToObject.resize(s.instances.size());
    for (size_t k = 0; k < s.instances.size(); ++k) s.worldToObject[k] = inverse3(s.instances[k].transform);
    std::vector<obvh::Tri48> flat; std::vector<float> lo, hi;
    obvh::flatten(s.vertices.data(), s.indices.data(), s.offsets.data(), s.meshIndexCounts.data(), s.instances.data(), s.instances.size(), flat, lo, hi);
    obvh::build(flat, lo, hi, s.bvh);
}

void sceneRefit(Scene& s) { // Renderer::updateAccelerationStructureInstances + updateTLAS (src/Renderer.cpp:671-742) on the wide BVH
    for (size_t k = 0; k < s.instances.size(); ++k) s.worldToObject[k] = inverse3(s.instances[k].transform);
    obvh::refit(s.vertices.data(), s.indices.data(), s.offsets.data(), s.instances.data(), s.bvh);
}

void probesInit(Probes& p, const vkx_grid_info& g) { // IrradianceProbes::init, src/IrradianceProbes.cpp:12-104
    p.grid = g;
    p.probeCount = uint32_t(g.resolution[0]) * uint32_t(g.resolution[1]) * uint32_t(g.resolution[2]);
    p.irrW = g.colorRes * uint32_t(g.resolution[0] * g.resolution[1]); p.irrH = g.colorRes * uint32_t(g.resolution[2]);
    p.depW = g.depthRes * uint32_t(g.resolution[0] * g.resolution[1]); p.depH = g.depthRes * uint32_t(g.resolution[2]);
    p.irrWork.assign(size_t(p.irrW) * p.irrH, 0); p.irrSampled = p.irrWork;
    p.depWork.assign(size_t(p.depW) * p.depH, 0); p.depSampled = p.depWork;
    p.state.assign(p.probeCount, 0);
}

void rayDirections(const float orientation[16], uint32_t count, float n, std::vector<float>& out) {
    mat3 R = mat3_from_mat4(orientation);
    out.resize(size_t(count) * 3);
    for (uint32_t i = 0; i < count; ++i) {
        vec3 d = R * sphericalFibonacci(float(i), n);
        out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
    }
}

// ---------------------------------------------------------------- probesInit.rgen
void classify(const Scene& s, Probes& p, const float orientation[16]) {
    const vkx_grid_info& grid = p.grid;
    std::vector<float> dirs;
    rayDirections(orientation, 512, float(grid.raysPerProbe), dirs);
    vec3 gridCellSize = probeGridCellSize(grid);
    float maxDistance = length(gridCellSize);
    float tmax = 1.5f * maxDistance;
    const float tmin = 0.01f;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t li = 0; li < int64_t(p.probeCount); ++li) {
        ivec3 pi3 = probeLinearIndexToGridIndex(uint32_t(li), grid);
        vec3 origin = probeIndexToWorldPosition(pi3, grid);
        bool affectGeometry = false;
        uint32_t backfaceHits = 0;
        for (int i = 0; i < 512; ++i) {
            vec3 direction = V3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
            vkx_hit h;
            float depth = tmax; bool isBackface = false;
            if (obvh::traceClosest(s.bvh, &origin.x, &direction.x, tmin, tmax, VKX_INSTANCE_STATIC, h)) { depth = h.t; isBackface = (h.primitive & 0x80000000u) != 0; }
            else depth = 3.402823466e+38f;
            if (depth < maxDistance) {
                if (isBackface) ++backfaceHits;
                vec3 position = origin + depth * direction;
                vec3 dist = abs(position - origin);
                if (dist.x < gridCellSize.x && dist.y < gridCellSize.y && dist.z < gridCellSize.z) affectGeometry = true;
            }
        }
        uint32_t st;
        if (float(backfaceHits) > 0.5f * float(grid.raysPerProbe)) st = 0;
        else st = affectGeometry ? 1 : 8;
        p.state[probeLinearIndex(pi3, grid)] = st;
    }
}

// ---------------------------------------------------------------- vertexSkinning.comp
void skinVertices(Scene& s, const float* jointTransforms, const uint16_t* skinJoints, const float* skinWeights, uint32_t srcOffset, uint32_t dstOffset, uint32_t size, float* motionVectors) {
    for (uint32_t i = 0; i < size; ++i) { // one invocation per vertex (:39-40)
        const float* w = skinWeights + 4 * size_t(i);
        mat4 J[4];
        for (int k = 0; k < 4; ++k) J[k] = mat4_from(jointTransforms + 16 * size_t(skinJoints[4 * size_t(i) + k]));
        mat4 skinMatrix; // :42-45, component-wise, left to right
        for (int c = 0; c < 4; ++c) skinMatrix[c] = ((J[0][c] * w[0] + J[1][c] * w[1]) + J[2][c] * w[2]) + J[3][c] * w[3];
        vkx_vertex& src = s.vertices[srcOffset + i];
        vkx_vertex& dst = s.vertices[dstOffset + i];
        vec4 np = skinMatrix * V4(src.pos[0], src.pos[1], src.pos[2], 1.0f);                     // :47
        vec3 motionVector = xyz(np) - V3(dst.pos[0], dst.pos[1], dst.pos[2]);                    // :48
        mat3 m3; for (int c = 0; c < 3; ++c) m3[c] = xyz(skinMatrix[c]);
        // :52-53 read the normal / tangent of the *destination* vertex (the bind-pose copy) ...
        vec3 normal = m3 * V3(dst.normal[0], dst.normal[1], dst.normal[2]);
        vec3 tangent = m3 * V3(dst.tangent[0], dst.tangent[1], dst.tangent[2]);
        dst.pos[0] = np.x; dst.pos[1] = np.y; dst.pos[2] = np.z;                                 // :49
        // ... and :54-57 store them at the *source* vertex (sic)
        src.normal[0] = normal.x; src.normal[1] = normal.y; src.normal[2] = normal.z;
        src.tangent[0] = tangent.x; src.tangent[1] = tangent.y; src.tangent[2] = tangent.z;
        if (motionVectors) { motionVectors[4 * size_t(i)] = motionVector.x; motionVectors[4 * size_t(i) + 1] = motionVector.y; motionVectors[4 * size_t(i) + 2] = motionVector.z; motionVectors[4 * size_t(i) + 3] = 1.0f; } // :59
    }
}

// ---------------------------------------------------------------- textures: anyhit.rahit, texDerivative
vec3 rotateAxis(vec3 p, vec3 axis, float angle) { // common.glsl:6-8
    return mix(dot(axis, p) * axis, p, std::cos(angle)) + cross(axis, p) * std::sin(angle);
}

static inline mat3 objectToWorld3(const vkx_instance& inst) { // mat3(gl_ObjectToWorldEXT): column c = (T[0][c], T[1][c], T[2][c])
    const float* T = inst.transform;
    mat3 m;
    for (int c = 0; c < 3; ++c) m[c] = V3(T[c], T[4 + c], T[8 + c]);
    return m;
}

bool anyHitIgnores(const Scene& s, uint32_t instance, uint32_t primitive, float u, float v) { // anyhit.rahit:24-48
    const vkx_offset_entry& oe = s.offsets[s.instances[instance].meshEntry];
    const vkx_material& m = s.materials[oe.materialIndex];
    if (m.albedoTexture == VKX_INVALID_TEXTURE) return false; // :29
    const vkx_vertex* vx[3];
    for (int c = 0; c < 3; ++c) vx[c] = &s.vertices[oe.vertexOffset + s.indices[oe.indexOffset + 3 * primitive + c]];
    vec3 bary = V3(1.0f - u - v, u, v); // :40
    vec2 texCoord = V2(vx[0]->texCoord[0], vx[0]->texCoord[1]) * bary.x + V2(vx[1]->texCoord[0], vx[1]->texCoord[1]) * bary.y + V2(vx[2]->texCoord[0], vx[2]->texCoord[1]) * bary.z; // :41
    otex::RGBA texColor = otex::sampleBase(s.textures[m.albedoTexture], texCoord.x, texCoord.y); // :43
    return texColor.a < 1e-2f; // :45-46
}
static bool anyHitThunk(const void* user, uint32_t instance, uint32_t primitive, float u, float v) { return anyHitIgnores(*static_cast<const Scene*>(user), instance, primitive, u, v); }
obvh::AnyHitFilter anyHitFilter(const Scene& s) { return obvh::AnyHitFilter{anyHitThunk, &s}; }

vec4 texDerivative(vec3 worldPosition, vec3 rayOrigin, const mat3& o2w, const vkx_vertex& v0, const vkx_vertex& v1, const vkx_vertex& v2, vec3 raydx, vec3 raydy) { // closesthit.glsl:50-107
    auto P = [](const vkx_vertex& v) { return V3(v.pos[0], v.pos[1], v.pos[2]); };
    auto UV = [](const vkx_vertex& v) { return V2(v.texCoord[0], v.texCoord[1]); };
    vec3 dpdu, dpdv;
    vec3 p01 = o2w * (P(v1) - P(v0));
    vec3 p02 = o2w * (P(v2) - P(v0));
    vec3 normal = normalize(cross(p01, p02));
    vec2 tex01 = UV(v1) - UV(v0);
    vec2 tex02 = UV(v2) - UV(v0);
    float det = tex01.x * tex02.y - tex01.y * tex02.x;
    if (std::fabs(det) < 1e-10f) {
        dpdu = normalize(std::fabs(normal.x) > std::fabs(normal.y) ? V3(-normal.z, 0, normal.x) : V3(0, -normal.z, normal.y));
        dpdv = cross(normal, dpdu);
    } else {
        float inv_det = 1.0f / det;
        dpdu = (tex02.y * p01 - tex01.y * p02) * inv_det;
        dpdv = (-tex02.x * p01 + tex01.x * p02) * inv_det;
    }
    float tx = dot(worldPosition - rayOrigin, normal) / dot(raydx, normal);
    float ty = dot(worldPosition - rayOrigin, normal) / dot(raydy, normal);
    vec3 dpdx = (rayOrigin + raydx * tx) - worldPosition;
    vec3 dpdy = (rayOrigin + raydy * ty) - worldPosition;
    float dudx = 0, dvdx = 0, dudy = 0, dvdy = 0;
    {
        int dim0 = 0, dim1 = 1;
        vec3 a = abs(normal);
        if (a.x > a.y && a.x > a.z) { dim0 = 1; dim1 = 2; }
        else if (a.y > a.z) { dim0 = 0; dim1 = 2; }
        float a00 = dpdu[dim0], a01 = dpdv[dim0], a10 = dpdu[dim1], a11 = dpdv[dim1];
        float det2 = a00 * a11 - a01 * a10;
        if (std::fabs(det2) > 1e-10f) {
            float inv_det = 1.0f / det2;
            dudx = (a11 * dpdx[dim0] - a01 * dpdx[dim1]) * inv_det;
            dvdx = (-a10 * dpdx[dim0] - a00 * dpdx[dim1]) * inv_det; // sic (:98)
            dudy = (a11 * dpdy[dim0] - a01 * dpdy[dim1]) * inv_det;
            dvdy = (-a10 * dpdy[dim0] - a00 * dpdy[dim1]) * inv_det; // sic (:101)
        }
    }
    return V4(dudx, dvdx, dudy, dvdy);
}

// ---------------------------------------------------------------- traceProbes.rgen + closesthit.glsl + miss.rmiss
static vec4 shadeRay(const Scene& s, const Probes& p, const vkx_light& light, vec3 origin, vec3 direction, float tmax,
                     vkx_hit& hit, uint8_t& shadowFlag, obvh::Counters* ctr, obvh::Counters* sctr, uint64_t& front,
                     float tmin, uint32_t cullMask, vec3 raydx, vec3 raydy, bool anyHit) {
    shadowFlag = 0;
    vec3 lightDir = V3(light.direction[0], light.direction[1], light.direction[2]);
    vec3 lightColor = V3(light.color[0], light.color[1], light.color[2]);
    const obvh::AnyHitFilter filter = anyHitFilter(s);
    const obvh::AnyHitFilter* flt = (anyHit && !s.textures.empty()) ? &filter : nullptr;
    if (!obvh::traceClosest(s.bvh, &origin.x, &direction.x, tmin, tmax, cullMask, hit, ctr, flt)) {
        vec3 c = sky(origin, direction, lightDir, lightColor, light.color[3], true); // miss.rmiss:19-22
        return V4(c, -1.0f);
    }
    float depth = hit.t;
    if (hit.primitive & 0x80000000u) return V4(0.0f, 0.0f, 0.0f, depth * 0.80f); // closesthit.glsl:137-141
    front++;
    float u = hit.u, v = hit.v;
    vec3 bary = V3(1.0f - u - v, u, v);
    vec3 position = direction * depth + origin; // :145
    const vkx_instance& inst = s.instances[hit.instance];
    const vkx_offset_entry& oe = s.offsets[inst.meshEntry];
    uint32_t prim = hit.primitive & 0x7FFFFFFFu;
    const vkx_vertex* vx[3];
    for (int c = 0; c < 3; ++c) vx[c] = &s.vertices[oe.vertexOffset + s.indices[oe.indexOffset + 3 * prim + c]];
    const vkx_material& m = s.materials[oe.materialIndex];
    auto N = [&](int c) { return V3(vx[c]->normal[0], vx[c]->normal[1], vx[c]->normal[2]); };
    vec3 tangentSpaceNormal = normalize(N(0) * bary.x + N(1) * bary.y + N(2) * bary.z); // :158
    const mat3& W = s.worldToObject[hit.instance];
    // vec3(tangentSpaceNormal * gl_WorldToObjectEXT): row vector times matrix -> dot with each column  (:159)
    vec3 normal = normalize(V3(dot(tangentSpaceNormal, W[0]), dot(tangentSpaceNormal, W[1]), dot(tangentSpaceNormal, W[2])));
    vec4 albedo = V4(m.baseColorFactor[0], m.baseColorFactor[1], m.baseColorFactor[2], 1.0f);
    float metalness = m.metallicFactor, roughness = m.roughnessFactor;
    vec3 emissiveLight = V3(m.emissiveFactor[0], m.emissiveFactor[1], m.emissiveFactor[2]);
    const bool textured = m.albedoTexture != VKX_INVALID_TEXTURE || m.normalTexture != VKX_INVALID_TEXTURE ||
                          m.metallicRoughnessTexture != VKX_INVALID_TEXTURE || m.emissiveTexture != VKX_INVALID_TEXTURE;
    if (textured) { // :157,161-192 (without a texture none of this reaches the result)
        auto UV = [&](int c) { return V2(vx[c]->texCoord[0], vx[c]->texCoord[1]); };
        vec2 texCoord = UV(0) * bary.x + UV(1) * bary.y + UV(2) * bary.z; // :157
        vec4 grad = texDerivative(position, origin, objectToWorld3(inst), *vx[0], *vx[1], *vx[2], raydx, raydy); // :161
        auto tex = [&](uint32_t index) {
            otex::RGBA c = otex::sampleGrad(s.textures[index], texCoord.x, texCoord.y, grad.x, grad.y, grad.z, grad.w); // textureGrad(.., grad.xy, grad.zw)
            return V4(c.r, c.g, c.b, c.a);
        };
        if (m.albedoTexture != VKX_INVALID_TEXTURE) albedo = albedo * tex(m.albedoTexture); // :163-166
        if (m.normalTexture != VKX_INVALID_TEXTURE) { // :169-177
            auto T = [&](int c) { return V4(vx[c]->tangent[0], vx[c]->tangent[1], vx[c]->tangent[2], vx[c]->tangent[3]); };
            vec4 tangentData = T(0) * bary.x + T(1) * bary.y + T(2) * bary.z;
            vec3 td = xyz(tangentData);
            vec3 tangent = normalize(V3(dot(td, W[0]), dot(td, W[1]), dot(td, W[2])));
            float tangentHandedness = tangentData.w;
            vec3 bitangent = cross(normal, tangent) * tangentHandedness;
            vec3 mappedNormal = normalize(2.0f * xyz(tex(m.normalTexture)) - 1.0f);
            normal = normalize(tangent * mappedNormal.x + bitangent * mappedNormal.y + normal * mappedNormal.z); // mat3(tangent, bitangent, normal) * mappedNormal
        }
        if (m.metallicRoughnessTexture != VKX_INVALID_TEXTURE) { // :181-185
            vec4 mr = tex(m.metallicRoughnessTexture);
            metalness *= mr.z; roughness *= mr.y;
        }
        if (m.emissiveTexture != VKX_INVALID_TEXTURE) emissiveLight *= xyz(tex(m.emissiveTexture)); // :188-190
    }
    vec3 color = V3(0.0f) + emissiveLight; // :194
    vec3 f0 = V3(0.04f);
    vec3 diffuseColor = xyz(albedo) * (1.0f - f0);
    diffuseColor *= (1.0f - metalness);
    vec3 specularColor = mix(f0, xyz(albedo), metalness);
    vec3 reflectDir = reflect(direction, normal);
    vec3 reflection = sampleProbes(p, position, reflectDir, -direction); // :241
    color += specularColor * reflection;
    vec3 indirectLight = sampleProbes(p, position, normal, -direction); // :248
    color += indirectLight * diffuseColor;
    // shadow ray :252-281 (tmin 0.1, tmax 10000, cull mask 0xFF, un-normalised light direction)
    bool isShadowed = obvh::traceAny(s.bvh, &position.x, &lightDir.x, 0.1f, 10000.0f, 0xFFu, sctr, flt);
    shadowFlag = isShadowed ? 2 : 1;
    if (!isShadowed) {
        vec4 pbr = pbrMetallicRoughness(normal, normalize(-direction), lightColor, lightDir, albedo, metalness, roughness);
        color += xyz(pbr);
        if (lightDir.y < 0.0f) color *= 1.0f - clampf(-lightDir.y, 0.0f, 0.1f) / 0.1f;
    }
    return V4(color, depth);
}

// One ray through closest-hit / miss shading with the caller's ray interval and cull mask (reflection.rgen:186 traces with
// tmin 0.1, tmax 10000, mask 0xff and payload.recursionDepth = 1, i.e. the same shading as the probe rays).
vec4 traceAndShade(const Scene& s, const Probes& p, const vkx_light& light, vec3 origin, vec3 direction, float tmin, float tmax, uint32_t cullMask,
                   vkx_hit& hit, uint8_t& shadowFlag, vec3 raydx, vec3 raydy, bool anyHit) {
    uint64_t front = 0;
    return shadeRay(s, p, light, origin, direction, tmax, hit, shadowFlag, nullptr, nullptr, front, tmin, cullMask, raydx, raydy, anyHit);
}

// ---------------------------------------------------------------- probesUpdate.glsl + probesCopyBorders.comp
static inline void borderSource(int T, int x, int y, int& sx, int& sy) {
    // Mirrored interior texel of a border texel of a T x T tile (tables in probesCopyBorders.comp:21-220).
    const int L = T - 1;
    bool bx = (x == 0 || x == L), by = (y == 0 || y == L);
    if (bx && by) { sx = x == 0 ? L - 1 : 1; sy = y == 0 ? L - 1 : 1; }
    else if (bx) { sx = x == 0 ? 1 : L - 1; sy = L - y; }
    else { sx = L - x; sy = y == 0 ? 1 : L - 1; }
}

void update(const Scene& s, Probes& p, const vkx_grid_info& g, const vkx_light& light, const float orientation[16],
            const uint32_t* indices, uint32_t count, int threads) {
    p.grid = g; // updateUniforms
    const vkx_grid_info& grid = p.grid;
    const uint32_t N = grid.raysPerProbe;
    std::vector<uint32_t> all;
    if (!indices) { all.resize(p.probeCount); for (uint32_t i = 0; i < p.probeCount; ++i) all[i] = i; indices = all.data(); count = p.probeCount; }
    rayDirections(orientation, N, float(N), p.dirs);
    p.rays.assign(size_t(count) * N * 4, 0.0f);
    p.hits.assign(size_t(count) * N, vkx_hit{});
    p.shadow.assign(size_t(count) * N, 0);
    p.irrUnpacked.assign(size_t(count) * 36 * 3, 0.0f);
    p.depUnpacked.assign(size_t(count) * 196 * 2, 0.0f);
    vec3 emax = V3(grid.extentMax[0], grid.extentMax[1], grid.extentMax[2]), emin = V3(grid.extentMin[0], grid.extentMin[1], grid.extentMin[2]);
    float tmax = length(emax - emin); // traceProbes.rgen:33
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    obvh::Counters total, stotal; uint64_t frontTotal = 0;
    // ---- trace + shade (traceProbes.rgen)
#pragma omp parallel
    {
        obvh::Counters ctr, sctr; uint64_t front = 0;
#pragma omp for schedule(dynamic, 4)
        for (int64_t slot = 0; slot < int64_t(count); ++slot) {
            ivec3 probeIndex = probeLinearIndexToGridIndex(indices[slot], grid);
            vec3 origin = probeIndexToWorldPosition(probeIndex, grid);
            for (uint32_t r = 0; r < N; ++r) {
                vec3 direction = V3(p.dirs[3 * r], p.dirs[3 * r + 1], p.dirs[3 * r + 2]);
                size_t ri = size_t(slot) * N + r;
                // payload.raydx / raydy, traceProbes.rgen:40-41 (NaN when the direction is parallel to the axis; only textured hits read them)
                vec3 raydx = V3(0.0f), raydy = V3(0.0f);
                if (!s.textures.empty()) {
                    raydx = rotateAxis(direction, normalize(cross(direction, V3(1, 0, 0))), 0.001f);
                    raydy = rotateAxis(direction, normalize(cross(direction, V3(0, 1, 0))), 0.001f);
                }
                vec4 c = shadeRay(s, p, light, origin, direction, tmax, p.hits[ri], p.shadow[ri], &ctr, &sctr, front, 0.01f /* traceProbes.rgen:27 */,
                                  VKX_INSTANCE_STATIC | VKX_INSTANCE_DYNAMIC /* :43 */, raydx, raydy, false);
                p.rays[4 * ri + 0] = c.x; p.rays[4 * ri + 1] = c.y; p.rays[4 * ri + 2] = c.z; p.rays[4 * ri + 3] = c.w;
            }
        }
#pragma omp critical
        { total.nodes += ctr.nodes; total.tris += ctr.tris; total.rays += ctr.rays; stotal.nodes += sctr.nodes; stotal.tris += sctr.tris; stotal.rays += sctr.rays; frontTotal += front; }
    }
    p.counters = total; p.shadowCounters = stotal; p.frontHits = frontTotal;

    // ---- blend (probesUpdate.glsl), decrees A.5.1-3
    float gridCellSize = length(probeGridCellSize(grid));
    const float hysteresis = grid.hysteresis;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t slot = 0; slot < int64_t(count); ++slot) {
        uint32_t linearIndex = indices[slot];
        ivec3 probeIndex = probeLinearIndexToGridIndex(linearIndex, grid);
        const float* rays = &p.rays[size_t(slot) * N * 4];
        // IRRADIANCE
        float globalMaxChange = 0.0f; uint32_t outOfRange00 = 0;
        ivec2 cbase = probeIndexToColorUVOffset(probeIndex, grid);
        for (int ly = 0; ly < 6; ++ly) for (int lx = 0; lx < 6; ++lx) {
            // localFragCoord = gl_LocalInvocationID.yz -> (lx, ly)
            vec2 nc = V2(0.33333f * (float(lx) - 2.5f), 0.33333f * (float(ly) - 2.5f));
            vec3 texelDirection = octDecode(nc);
            vec4 result = V4(0, 0, 0, 0); uint32_t outOfRange = 0;
            for (uint32_t i = 0; i < N; ++i) {
                vec4 rayData = V4(rays[4 * i], rays[4 * i + 1], rays[4 * i + 2], rays[4 * i + 3]);
                vec3 direction = V3(p.dirs[3 * i], p.dirs[3 * i + 1], p.dirs[3 * i + 2]);
                if (rayData.w < 0.0f || rayData.w > gridCellSize) ++outOfRange;
                float weight = std::max(0.0f, dot(texelDirection, direction));
                result += V4(weight * rayData.x, weight * rayData.y, weight * rayData.z, weight);
            }
            if (result.w > 1e-3f) { result.x /= result.w; result.y /= result.w; result.z /= result.w; }
            size_t gi = size_t(cbase.y + 1 + ly) * p.irrW + size_t(cbase.x + 1 + lx);
            float prev[3]; opack::unpackR11G11B10(p.irrWork[gi], prev);
            float maxChange = std::max(std::max(std::fabs(result.x - prev[0]), std::fabs(result.y - prev[1])), std::fabs(result.z - prev[2]));
            float out[3] = {mix(result.x, prev[0], hysteresis), mix(result.y, prev[1], hysteresis), mix(result.z, prev[2], hysteresis)};
            p.irrWork[gi] = opack::packR11G11B10(out[0], out[1], out[2]);
            float* up = &p.irrUnpacked[(size_t(slot) * 36 + size_t(ly * 6 + lx)) * 3];
            up[0] = out[0]; up[1] = out[1]; up[2] = out[2];
            // atomicMax on the uint bit pattern (non-negative floats; NaN patterns compare above +inf)
            if (opack::f2u(maxChange) > opack::f2u(globalMaxChange)) globalMaxChange = maxChange;
            if (lx == 0 && ly == 0) outOfRange00 = outOfRange;
        }
        { // state machine, texel (0,0) (:110-119)
            uint32_t st = p.state[linearIndex];
            if (outOfRange00 >= N) st = 8;
            else {
                float maxChange = globalMaxChange;
                if (maxChange < 0.02f / float(st)) st = std::min(st + 1, 8u);
                else if (maxChange > 0.04f / float(st)) st = std::max(st - 1, 1u);
                else if (maxChange > 0.25f) st = 1;
            }
            p.state[linearIndex] = st;
        }
        // DEPTH
        ivec2 dbase = probeIndexToDepthUVOffset(probeIndex, grid);
        for (int ly = 0; ly < 14; ++ly) for (int lx = 0; lx < 14; ++lx) {
            vec2 nc = V2(0.142857f * (float(lx) - 6.5f), 0.142857f * (float(ly) - 6.5f));
            vec3 texelDirection = octDecode(nc);
            vec4 result = V4(0, 0, 0, 0);
            for (uint32_t i = 0; i < N; ++i) {
                float w4 = rays[4 * i + 3];
                vec3 direction = V3(p.dirs[3 * i], p.dirs[3 * i + 1], p.dirs[3 * i + 2]);
                float depth = std::min(gridCellSize, w4);
                if (depth < 0.0f) depth = gridCellSize;
                float weight = std::pow(std::max(0.0f, dot(texelDirection, direction)), grid.depthSharpness);
                result += V4(weight * depth, weight * depth * depth, 0.0f, weight);
            }
            if (result.w > 1e-3f) { result.x /= result.w; result.y /= result.w; result.z /= result.w; }
            size_t gi = size_t(dbase.y + 1 + ly) * p.depW + size_t(dbase.x + 1 + lx);
            float prev[2]; opack::unpackRG16F(p.depWork[gi], prev);
            float out[2] = {mix(result.x, prev[0], hysteresis), mix(result.y, prev[1], hysteresis)};
            p.depWork[gi] = opack::packRG16F(out[0], out[1]);
            float* up = &p.depUnpacked[(size_t(slot) * 196 + size_t(ly * 14 + lx)) * 2];
            up[0] = out[0]; up[1] = out[1];
        }
        // borders (probesCopyBorders.comp)
        for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
            if (x != 0 && x != 7 && y != 0 && y != 7) continue;
            int sx, sy; borderSource(8, x, y, sx, sy);
            p.irrWork[size_t(cbase.y + y) * p.irrW + size_t(cbase.x + x)] = p.irrWork[size_t(cbase.y + sy) * p.irrW + size_t(cbase.x + sx)];
        }
        for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) {
            if (x != 0 && x != 15 && y != 0 && y != 15) continue;
            int sx, sy; borderSource(16, x, y, sx, sy);
            p.depWork[size_t(dbase.y + y) * p.depW + size_t(dbase.x + x)] = p.depWork[size_t(dbase.y + sy) * p.depW + size_t(dbase.x + sx)];
        }
    }
    // ---- publish (vkCmdCopyImage work -> sampled, src/IrradianceProbes.cpp:531-576)
    p.irrSampled = p.irrWork;
    p.depSampled = p.depWork;
}

// ---------------------------------------------------------------- host logic of IrradianceProbes.cpp
static uint32_t randU32(MsvcRand& r) {
    // glm::detail::compute_rand<1, uint32>: (u16 << 16) | u16, u16 = (u8 << 8) | u8, u8 = rand() % 255
    // (ext/glm/glm/gtc/random.inl:19-85). Operand evaluation order of `|` is unspecified in C++. Decree: the order g++ 13
    // gives the unmodified GLM headers (right operand first, i.e. the first draw is the lowest byte), which is what
    // tests/golden/glm_pin.json pins; MSVC's order (the reference's only compiler) cannot be verified here.
    uint32_t b0 = uint32_t(r.next() % 255), b1 = uint32_t(r.next() % 255), b2 = uint32_t(r.next() % 255), b3 = uint32_t(r.next() % 255);
    return (((b3 << 8) | b2) << 16) | ((b1 << 8) | b0);
}
static float linearRand(MsvcRand& r, float Min, float Max) { // random.inl:177-183
    return float(randU32(r)) / float(std::numeric_limits<uint32_t>::max()) * (Max - Min) + Min;
}
void sphericalRand(MsvcRand& rng, float out[3]) { // random.inl:290-302
    float theta = linearRand(rng, 0.0f, 6.283185307179586476925286766559f);
    float phi = std::acos(linearRand(rng, -1.0f, 1.0f));
    out[0] = std::sin(phi) * std::cos(theta);
    out[1] = std::sin(phi) * std::sin(theta);
    out[2] = std::cos(phi);
    for (int i = 0; i < 3; ++i) out[i] = out[i] * 1.0f;
}
void orientationFromZ(const float Zf[3], float out16[16]) { // genBasis, src/IrradianceProbes.cpp:347-355,455-460
    vec3 n = V3(Zf[0], Zf[1], Zf[2]);
    vec3 b1 = n.x > 0.9f ? V3(0.0f, 1.0f, 0.0f) : V3(1.0f, 0.0f, 0.0f);
    b1 -= n * dot(b1, n);
    b1 = normalize(b1);
    vec3 b2 = cross(n, b1);
    mat3 M; M[0] = b1; M[1] = b2; M[2] = n; // glm::mat3(X, Y, Z): columns
    mat3 Tm = transpose(M);
    for (int i = 0; i < 16; ++i) out16[i] = 0.0f;
    for (int c = 0; c < 3; ++c) { out16[4 * c + 0] = Tm[c].x; out16[4 * c + 1] = Tm[c].y; out16[4 * c + 2] = Tm[c].z; }
    out16[15] = 1.0f;
}
uint32_t selectProbesToUpdate(Scheduler& s, const uint32_t* state, uint32_t probeCount, uint32_t probesPerUpdate, uint32_t* out) { // :396-424
    uint32_t n = 0, idx = s.lastUpdateOffset, checked = 0;
    while (checked < probeCount && (probesPerUpdate == 0 || n < probesPerUpdate)) {
        if (state[idx] != 0 && ((idx + s.loopIndex) % state[idx]) == 0) out[n++] = idx;
        ++idx;
        if (idx >= probeCount) { idx = 0; ++s.loopIndex; }
        ++checked;
    }
    s.lastUpdateOffset = idx;
    return n;
}

} // namespace oddgi
