// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
//
// Sequential CPU restatement of "BVH spec v2" (v2 = v1 with the per-slot meta bytes replaced by one validity word; v2.1: division-free barycentric acceptance tests): the deterministic binned-SAH binary build,
// the greedy collapse into 8-wide nodes, the 8-bit child-box quantisation, and the per-ray traversal order.
// The reference delegates all of this to the Vulkan driver / RT cores (reference src/Renderer.cpp:272-449,
// 525-642), so nothing here follows reference code: this file *defines* the behaviour the
// CUDA builder and traversal kernels must reproduce bit-exactly (topology, quantised bytes, hit ids, t, u, v).
#pragma once
#include <cstdint>
#include <vector>
#include "../include/vkx.h"

namespace obvh {

struct Tri48 {          // 48 B, three 128-bit words
    float v0[3];
    float e1[3];
    float e2[3];
    uint32_t inst;      // bits 0..23 instance index, bits 24..31 instance mask
    uint32_t prim;      // bits 0..30 primitive index in the mesh, bit 31: winding flipped (negative-determinant transform)
    uint32_t pad;
};
static_assert(sizeof(Tri48) == 48, "Tri48");

struct Node80 {         // 80 B, five 128-bit words
    float p[3];         // quantisation origin = node AABB min
    uint8_t e[3];       // biased exponents: cell size on axis a = asfloat(e[a] << 23)
    uint8_t imask;      // bit s set: slot s holds an inner child
    uint32_t childBase; // wide-node index of the first inner child (inner children contiguous, in slot order)
    uint32_t primBase;  // index of the node's first triangle (leaf children contiguous, in slot order)
    uint32_t valid;     // spec v2: bits 24..31 = imask; bits [3s, 3s + count) set for leaf slot s (count <= 3); 0 for empty slots.
                        // Triangle of bit b = primBase + popcount(valid & ((1 << b) - 1) & 0xFFFFFF): leaf triangles are contiguous in slot order.
    uint32_t pad;       // 0
    uint8_t qlo[3][8];  // quantised child AABB mins  [axis][slot]
    uint8_t qhi[3][8];  // quantised child AABB maxs
};
static_assert(sizeof(Node80) == 80, "Node80");

struct Bvh {
    std::vector<Node80> nodes;
    std::vector<Tri48> tris;
    uint32_t numBinaryNodes = 0;
    uint32_t depth = 0;
    float sceneMin[3] = {0, 0, 0}, sceneMax[3] = {0, 0, 0};
};

extern int gMaxStack;
struct Counters { uint64_t nodes = 0, tris = 0, rays = 0; };

// anyhit.rahit as a candidate filter: ignore(user, instance, primitive, u, v) == true drops the candidate (ignoreIntersectionEXT).
// It is asked only for candidates that would otherwise be accepted, so the result is the closest / any non-ignored hit whatever
// the order of the triangle tests (Vulkan leaves the order of any-hit invocations unspecified too).
struct AnyHitFilter { bool (*ignore)(const void* user, uint32_t instance, uint32_t primitive, float u, float v); const void* user; };

// Flatten instances into world-space triangles (spec section 3.1) in (instance, primitive) order.
void flatten(const vkx_vertex* vertices, const uint32_t* indices, const vkx_offset_entry* offsets,
             const uint32_t* meshIndexCounts, const vkx_instance* instances, size_t numInstances,
             std::vector<Tri48>& out, std::vector<float>& lo, std::vector<float>& hi);

void build(const std::vector<Tri48>& flat, const std::vector<float>& lo, const std::vector<float>& hi, Bvh& out);

// Topology-preserving refit after the vertices / instance transforms changed (same instance list, same meshes): see bvh.cpp.
void refit(const vkx_vertex* vertices, const uint32_t* indices, const vkx_offset_entry* offsets, const vkx_instance* instances, Bvh& bvh);

// Closest hit. Returns true on hit. hit.t < 0 on miss.
bool traceClosest(const Bvh& bvh, const float o[3], const float d[3], float tmin, float tmax, uint32_t cullMask,
                  vkx_hit& hit, Counters* ctr = nullptr, const AnyHitFilter* filter = nullptr);
// Terminate-on-first-hit occlusion query.
bool traceAny(const Bvh& bvh, const float o[3], const float d[3], float tmin, float tmax, uint32_t cullMask,
              Counters* ctr = nullptr, const AnyHitFilter* filter = nullptr);

} // namespace obvh
