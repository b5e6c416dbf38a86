// Closest-hit shading of one front-face hit, shared by the probe rays (k_shade_front) and the reflection rays (reflection.cu).
// Replaces (reference): src/shaders/closesthit.glsl:143-288 with payload.recursionDepth >= 1 (the NO_REFLECTION variant and the
// reflection pipeline's variant shade identically: reflection.rgen:125 starts at recursion depth 1), untextured materials.
#pragma once
#include "common.cuh"
#include "shade.cuh"

// base = emissive + specular * sampleProbes(reflectDir) + diffuse * sampleProbes(normal)   (the colour if the sun is occluded)
// lit  = (base + direct PBR term) * night fade                                             (the colour if the shadow ray escapes)
__device__ __forceinline__ void shadeFrontHit(const DeviceScene& sc, const DeviceProbes& pr, const GridConsts& gc, v3 lightDir, v3 lightColor, v3 direction, v3 position,
                                              const vkx_hit& h, v3& base, v3& lit) {
    const float u = h.u, v = h.v;
    const float bx = 1.0f - u - v, by = u, bz = v;
    const uint32_t meshEntry = __ldg(&sc.instances[h.instance].meshEntry);
    const vkx_offset_entry oe = sc.offsets[meshEntry];
    const uint32_t prim = h.primitive & 0x7FFFFFFFu;
    v3 n3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t vi = oe.vertexOffset + __ldg(sc.indices + oe.indexOffset + 3 * prim + c);
        const float* nn = sc.vertices[vi].normal;
        n3[c] = mk3(__ldg(nn), __ldg(nn + 1), __ldg(nn + 2));
    }
    const vkx_material m = sc.materials[oe.materialIndex];
    const v3 tsn = norm3(n3[0] * bx + n3[1] * by + n3[2] * bz);
    const float* W = sc.worldToObject + size_t(h.instance) * 9; // W[row][col]
    // vec3(tsn * worldToObject): component j = dot(tsn, column j)
    const v3 normal = norm3(mk3(dot3(tsn, mk3(W[0], W[3], W[6])), dot3(tsn, mk3(W[1], W[4], W[7])), dot3(tsn, mk3(W[2], W[5], W[8]))));
    const v3 albedo = mk3(m.baseColorFactor[0], m.baseColorFactor[1], m.baseColorFactor[2]);
    const float metalness = m.metallicFactor, roughness = m.roughnessFactor;
    v3 color = mk3(0.0f) + mk3(m.emissiveFactor[0], m.emissiveFactor[1], m.emissiveFactor[2]);
    const v3 f0 = mk3(0.04f);
    v3 diffuseColor = albedo * (1.0f - f0);
    diffuseColor = diffuseColor * (1.0f - metalness);
    const v3 specularColor = mix3(f0, albedo, metalness);
    const v3 reflectDir = reflect3(direction, normal);
    v3 reflection, indirectLight;
    sampleProbes2(pr, gc, position, reflectDir, normal, -direction, reflection, indirectLight);
    color = color + specularColor * reflection;
    color = color + indirectLight * diffuseColor;
    base = color;
    lit = color + pbrMetallicRoughness(normal, norm3(-direction), lightColor, lightDir, albedo, metalness, roughness);
    if (lightDir.y < 0.0f) lit = lit * (1.0f - clampS(-lightDir.y, 0.0f, 0.1f) / 0.1f);
}
