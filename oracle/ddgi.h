// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
// CPU transliteration of the reference's DDGI probe update shaders and the host logic around them.
#pragma once
#include <cstdint>
#include <vector>
#include "../include/vkx.h"
#include "bvh.h"
#include "texture.h"
#include "vmath.h"

namespace oddgi {

struct Scene {
    std::vector<vkx_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<vkx_offset_entry> offsets;
    std::vector<uint32_t> meshIndexCounts;
    std::vector<vkx_material> materials;
    std::vector<vkx_instance> instances;
    std::vector<ovm::mat3> worldToObject; // per instance, inverse of the 3x3 part (math convention W[row][col] = c[col][row])
    std::vector<otex::Texture> textures;  // Scene's texture list with generated mip chains (sampler spec v1, texture.h)
    obvh::Bvh bvh;
};

// anyhit.rahit:24-48 for one candidate: true = ignoreIntersectionEXT (albedo alpha < 1e-2 at the hit's texture coordinate)
bool anyHitIgnores(const Scene& s, uint32_t instance, uint32_t primitive, float u, float v);
// The filter object for obvh::traceClosest / traceAny in pipelines whose hit group has the any-hit shader (direct light, reflection)
obvh::AnyHitFilter anyHitFilter(const Scene& s);
ovm::vec3 rotateAxis(ovm::vec3 p, ovm::vec3 axis, float angle); // common.glsl:6-8
// texDerivative, closesthit.glsl:50-107 -> (dudx, dvdx, dudy, dvdy)
ovm::vec4 texDerivative(ovm::vec3 worldPosition, ovm::vec3 rayOrigin, const ovm::mat3& objectToWorld, const vkx_vertex& v0, const vkx_vertex& v1, const vkx_vertex& v2,
                        ovm::vec3 raydx, ovm::vec3 raydy);

struct Probes {
    vkx_grid_info grid{};
    uint32_t probeCount = 0, irrW = 0, irrH = 0, depW = 0, depH = 0;
    std::vector<uint32_t> irrWork, irrSampled;   // B10G11R11
    std::vector<uint32_t> depWork, depSampled;   // RG16F
    std::vector<uint32_t> state;
    // last update (debug / parity side buffers)
    std::vector<float> rays;        // [count][N][4]
    std::vector<vkx_hit> hits;      // [count][N]
    std::vector<uint8_t> shadow;    // [count][N] 0 not traced, 1 lit, 2 shadowed
    std::vector<float> irrUnpacked; // [count][36][3]  post-hysteresis fp32
    std::vector<float> depUnpacked; // [count][196][2]
    std::vector<float> dirs;        // [N][3] rotated ray directions of the last update
    obvh::Counters counters;        // traversal counters of the primary rays of the last update
    obvh::Counters shadowCounters;  // ... and of the shadow rays
    uint64_t frontHits = 0;
};

void sceneFinalize(Scene& s); // worldToObject + BVH
void sceneRefit(Scene& s);    // worldToObject + topology-preserving BVH refit after instance transforms / vertices changed
// vertexSkinning.comp:37-60 over `size` vertices (push constants srcOffset / dstOffset); motionVectors: optional, 4 floats per vertex.
// The BVH must be rebuilt afterwards (sceneFinalize), as Renderer::updateSkinnedBLAS does for the skinned BLASes.
void skinVertices(Scene& s, const float* jointTransforms, const uint16_t* skinJoints, const float* skinWeights, uint32_t srcOffset, uint32_t dstOffset, uint32_t size, float* motionVectors);
void probesInit(Probes& p, const vkx_grid_info& g);
// mat3(orientation) * sphericalFibonacci(i, n) for i in [0, count)  (traceProbes.rgen:36, probesInit.rgen:41)
void rayDirections(const float orientation[16], uint32_t count, float n, std::vector<float>& out);
void classify(const Scene& s, Probes& p, const float orientation[16]);
void update(const Scene& s, Probes& p, const vkx_grid_info& g, const vkx_light& light, const float orientation[16],
            const uint32_t* indices, uint32_t count, int threads);

// closest-hit / miss shading of one ray (closesthit.glsl with recursionDepth >= 1, miss.rmiss); returns (rgb, depth) like the payload.
// raydx / raydy: payload ray differentials set by the ray-generation shader; anyHit: the pipeline's hit group has anyhit.rahit
// (reflection pipeline: yes, src/RenderPasses/ReflectionPipeline.cpp:51; probe pipeline: no, src/IrradianceProbes.cpp:244-254).
ovm::vec4 traceAndShade(const Scene& s, const Probes& p, const vkx_light& light, ovm::vec3 origin, ovm::vec3 direction, float tmin, float tmax, uint32_t cullMask,
                        vkx_hit& hit, uint8_t& shadowFlag, ovm::vec3 raydx, ovm::vec3 raydy, bool anyHit);
ovm::vec3 sky(ovm::vec3 rayOrigin, ovm::vec3 rayDirection, ovm::vec3 sunPosition, ovm::vec3 sunColor, float sunBrightnessFactor, bool showSun);
ovm::vec4 pbrMetallicRoughness(ovm::vec3 normal, ovm::vec3 view, ovm::vec3 lightColor, ovm::vec3 lightDirection, ovm::vec4 albedo, float metalness, float roughness);
ovm::vec3 sampleProbes(const Probes& p, ovm::vec3 position, ovm::vec3 normal, ovm::vec3 toCamera);
ovm::vec3 sampleIrradiance(const Probes& p, ovm::vec2 uv); // textureLod(colorTex, uv, 0): decreed exact-fp32 bilinear, REPEAT
ovm::vec2 sampleDepth(const Probes& p, ovm::vec2 uv);
ovm::vec3 sphericalFibonacci(float i, float n);
ovm::vec2 octEncode(ovm::vec3 v);
void probeHelpers(const vkx_grid_info& g, uint32_t index, int outI[8], float outF[6]);
ovm::vec3 octDecode(ovm::vec2 o);
ovm::vec2 spherePointToOctohedralUV(ovm::vec3 direction);

// Host logic of IrradianceProbes.cpp
struct MsvcRand { uint32_t x = 1; int next() { x = x * 214013u + 2531011u; return int((x >> 16) & 0x7FFFu); } };
void sphericalRand(MsvcRand& rng, float out[3]);            // glm::sphericalRand(1.0f)
void orientationFromZ(const float Z[3], float out16[16]);   // genBasis + mat4(transpose(mat3(X,Y,Z)))
struct Scheduler { uint32_t loopIndex = 0, lastUpdateOffset = 0; };
uint32_t selectProbesToUpdate(Scheduler& s, const uint32_t* state, uint32_t probeCount, uint32_t probesPerUpdate, uint32_t* out);

} // namespace oddgi
