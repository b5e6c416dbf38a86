// Closest-hit shading of one front-face hit, shared by the probe rays (k_shade_front) and the reflection rays (reflection.cu).
// Replaces (reference): src/shaders/closesthit.glsl:143-288 with payload.recursionDepth >= 1 (the NO_REFLECTION variant and the
// reflection pipeline's variant shade identically: reflection.rgen:125 starts at recursion depth 1). TEXTURED = false is the
// kernel variant for scenes without a texture list (no texture code, no extra registers); TEXTURED = true adds
// closesthit.glsl:157,161-192 (texDerivative from the payload's ray differentials, four textureGrad look-ups, normal mapping).
#pragma once
#include "common.cuh"
#include "shade.cuh"
#include "texture.cuh"

__device__ __forceinline__ v3 rotateAxisH(v3 p, v3 axis, float angle) { // common.glsl:6-8
    return mix3(dot3(axis, p) * axis, p, cosf(angle)) + cross3(axis, p) * sinf(angle);
}
__device__ __forceinline__ float comp3(v3 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : v.z; }

// texDerivative, closesthit.glsl:50-107 -> (dudx, dvdx, dudy, dvdy). p01 / p02: world-space edges mat3(gl_ObjectToWorldEXT) * (v1 - v0).
// The two determinants that select a branch are evaluated without contraction so that the branch matches the oracle's.
__device__ __forceinline__ float4 texDerivativeD(v3 worldPosition, v3 rayOrigin, v3 p01, v3 p02, float2 uv0, float2 uv1, float2 uv2, v3 raydx, v3 raydy) {
    v3 dpdu, dpdv;
    const v3 normal = norm3(cross3(p01, p02));
    const float t01x = uv1.x - uv0.x, t01y = uv1.y - uv0.y, t02x = uv2.x - uv0.x, t02y = uv2.y - uv0.y;
    const float det = __fsub_rn(__fmul_rn(t01x, t02y), __fmul_rn(t01y, t02x));
    if (fabsf(det) < 1e-10f) {
        dpdu = norm3(fabsf(normal.x) > fabsf(normal.y) ? mk3(-normal.z, 0.0f, normal.x) : mk3(0.0f, -normal.z, normal.y));
        dpdv = cross3(normal, dpdu);
    } else {
        const float inv_det = 1.0f / det;
        dpdu = (t02y * p01 - t01y * p02) * inv_det;
        dpdv = (-t02x * p01 + t01x * p02) * inv_det;
    }
    const float tx = dot3(worldPosition - rayOrigin, normal) / dot3(raydx, normal);
    const float ty = dot3(worldPosition - rayOrigin, normal) / dot3(raydy, normal);
    const v3 dpdx = (rayOrigin + raydx * tx) - worldPosition;
    const v3 dpdy = (rayOrigin + raydy * ty) - worldPosition;
    float dudx = 0.0f, dvdx = 0.0f, dudy = 0.0f, dvdy = 0.0f;
    int dim0 = 0, dim1 = 1;
    const v3 a = abs3(normal);
    if (a.x > a.y && a.x > a.z) { dim0 = 1; dim1 = 2; }
    else if (a.y > a.z) { dim0 = 0; dim1 = 2; }
    const float a00 = comp3(dpdu, dim0), a01 = comp3(dpdv, dim0), a10 = comp3(dpdu, dim1), a11 = comp3(dpdv, dim1);
    const float det2 = __fsub_rn(__fmul_rn(a00, a11), __fmul_rn(a01, a10));
    if (fabsf(det2) > 1e-10f) {
        const float inv_det = 1.0f / det2;
        dudx = (a11 * comp3(dpdx, dim0) - a01 * comp3(dpdx, dim1)) * inv_det;
        dvdx = (-a10 * comp3(dpdx, dim0) - a00 * comp3(dpdx, dim1)) * inv_det; // sic (closesthit.glsl:98)
        dudy = (a11 * comp3(dpdy, dim0) - a01 * comp3(dpdy, dim1)) * inv_det;
        dvdy = (-a10 * comp3(dpdy, dim0) - a00 * comp3(dpdy, dim1)) * inv_det; // sic (:101)
    }
    return make_float4(dudx, dvdx, dudy, dvdy);
}

// base = emissive + specular * sampleProbes(reflectDir) + diffuse * sampleProbes(normal)   (the colour if the sun is occluded)
// lit  = (base + direct PBR term) * night fade                                             (the colour if the shadow ray escapes)
// rayOrigin / raydx / raydy (gl_WorldRayOriginEXT and the payload's ray differentials) are read by the TEXTURED variant only.
template <bool TEXTURED>
__device__ __forceinline__ void shadeFrontHit(const DeviceScene& sc, const DeviceProbes& pr, const GridConsts& gc, v3 lightDir, v3 lightColor, v3 direction, v3 position,
                                              const vkx_hit& h, v3& base, v3& lit, v3 rayOrigin = mk3(0.0f), v3 raydx = mk3(0.0f), v3 raydy = mk3(0.0f)) {
    const float u = h.u, v = h.v;
    const float bx = __fsub_rn(__fsub_rn(1.0f, u), v), by = u, bz = v;
    const uint32_t meshEntry = __ldg(&sc.instances[h.instance].meshEntry);
    const vkx_offset_entry oe = sc.offsets[meshEntry];
    const uint32_t prim = h.primitive & 0x7FFFFFFFu;
    v3 n3[3];
    uint32_t vidx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t vi = oe.vertexOffset + __ldg(sc.indices + oe.indexOffset + 3 * prim + c);
        vidx[c] = vi;
        const float* nn = sc.vertices[vi].normal;
        n3[c] = mk3(__ldg(nn), __ldg(nn + 1), __ldg(nn + 2));
    }
    const vkx_material m = sc.materials[oe.materialIndex];
    // exact chain (shade.cuh): the normal selects atlas texels through its octahedral coordinate
    const v3 tsn = xnorm3(xadd3(xadd3(xmul3(n3[0], bx), xmul3(n3[1], by)), xmul3(n3[2], bz)));
    const float* W = sc.worldToObject + size_t(h.instance) * 9; // W[row][col]
    // vec3(tsn * worldToObject): component j = dot(tsn, column j)
    v3 normal = xnorm3(mk3(xdot3(tsn, mk3(W[0], W[3], W[6])), xdot3(tsn, mk3(W[1], W[4], W[7])), xdot3(tsn, mk3(W[2], W[5], W[8]))));
    v3 albedo = mk3(m.baseColorFactor[0], m.baseColorFactor[1], m.baseColorFactor[2]);
    float metalness = m.metallicFactor, roughness = m.roughnessFactor;
    v3 emissive = mk3(m.emissiveFactor[0], m.emissiveFactor[1], m.emissiveFactor[2]);
    if (TEXTURED) {
        if (m.albedoTexture != VKX_INVALID_TEXTURE || m.normalTexture != VKX_INVALID_TEXTURE || m.metallicRoughnessTexture != VKX_INVALID_TEXTURE || m.emissiveTexture != VKX_INVALID_TEXTURE) {
            v3 p3[3]; float2 uv3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const vkx_vertex* vx = sc.vertices + vidx[c];
                p3[c] = mk3(__ldg(vx->pos), __ldg(vx->pos + 1), __ldg(vx->pos + 2));
                uv3[c] = make_float2(__ldg(vx->texCoord), __ldg(vx->texCoord + 1));
            }
            const float2 texCoord = make_float2(uv3[0].x * bx + uv3[1].x * by + uv3[2].x * bz, uv3[0].y * bx + uv3[1].y * by + uv3[2].y * bz); // closesthit.glsl:157
            const float* T = sc.instances[h.instance].transform; // row-major 3x4: mat3(gl_ObjectToWorldEXT) * e = rows dot e
            const v3 e1 = p3[1] - p3[0], e2 = p3[2] - p3[0];
            const v3 r0 = mk3(__ldg(T), __ldg(T + 1), __ldg(T + 2)), r1 = mk3(__ldg(T + 4), __ldg(T + 5), __ldg(T + 6)), r2 = mk3(__ldg(T + 8), __ldg(T + 9), __ldg(T + 10));
            // column-major product as the GLSL evaluates it: col0 * e.x + col1 * e.y + col2 * e.z
            const v3 p01 = mk3(r0.x * e1.x + r0.y * e1.y + r0.z * e1.z, r1.x * e1.x + r1.y * e1.y + r1.z * e1.z, r2.x * e1.x + r2.y * e1.y + r2.z * e1.z);
            const v3 p02 = mk3(r0.x * e2.x + r0.y * e2.y + r0.z * e2.z, r1.x * e2.x + r1.y * e2.y + r1.z * e2.z, r2.x * e2.x + r2.y * e2.y + r2.z * e2.z);
            const float4 grad = texDerivativeD(position, rayOrigin, p01, p02, uv3[0], uv3[1], uv3[2], raydx, raydy); // :161
            if (m.albedoTexture != VKX_INVALID_TEXTURE) { // :163-166 (albedo.a only reaches payload.color.a, which no caller stores)
                const float4 t = texSampleGrad(sc, m.albedoTexture, texCoord.x, texCoord.y, grad.x, grad.y, grad.z, grad.w);
                albedo = albedo * mk3(t.x, t.y, t.z);
            }
            if (m.normalTexture != VKX_INVALID_TEXTURE) { // :169-177
                float4 tg[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { const float* tp = sc.vertices[vidx[c]].tangent; tg[c] = make_float4(__ldg(tp), __ldg(tp + 1), __ldg(tp + 2), __ldg(tp + 3)); }
                const v3 td = mk3(tg[0].x * bx + tg[1].x * by + tg[2].x * bz, tg[0].y * bx + tg[1].y * by + tg[2].y * bz, tg[0].z * bx + tg[1].z * by + tg[2].z * bz);
                const float handedness = tg[0].w * bx + tg[1].w * by + tg[2].w * bz;
                const v3 tangent = norm3(mk3(dot3(td, mk3(W[0], W[3], W[6])), dot3(td, mk3(W[1], W[4], W[7])), dot3(td, mk3(W[2], W[5], W[8]))));
                const v3 bitangent = cross3(normal, tangent) * handedness;
                const float4 t = texSampleGrad(sc, m.normalTexture, texCoord.x, texCoord.y, grad.x, grad.y, grad.z, grad.w);
                const v3 mapped = norm3(2.0f * mk3(t.x, t.y, t.z) + (-1.0f));
                normal = norm3(tangent * mapped.x + bitangent * mapped.y + normal * mapped.z);
            }
            if (m.metallicRoughnessTexture != VKX_INVALID_TEXTURE) { // :181-185
                const float4 t = texSampleGrad(sc, m.metallicRoughnessTexture, texCoord.x, texCoord.y, grad.x, grad.y, grad.z, grad.w);
                metalness *= t.z; roughness *= t.y;
            }
            if (m.emissiveTexture != VKX_INVALID_TEXTURE) { // :188-190
                const float4 t = texSampleGrad(sc, m.emissiveTexture, texCoord.x, texCoord.y, grad.x, grad.y, grad.z, grad.w);
                emissive = emissive * mk3(t.x, t.y, t.z);
            }
        }
    }
    v3 color = mk3(0.0f) + emissive;
    const v3 f0 = mk3(0.04f);
    v3 diffuseColor = albedo * (1.0f - f0);
    diffuseColor = diffuseColor * (1.0f - metalness);
    const v3 specularColor = mix3(f0, albedo, metalness);
    const v3 reflectDir = xreflect3(direction, normal);
    v3 reflection, indirectLight;
    sampleProbes2(pr, gc, position, reflectDir, normal, -direction, reflection, indirectLight);
    color = color + specularColor * reflection;
    color = color + indirectLight * diffuseColor;
    base = color;
    lit = color + pbrMetallicRoughness(normal, norm3(-direction), lightColor, lightDir, albedo, metalness, roughness);
    if (lightDir.y < 0.0f) lit = lit * (1.0f - clampS(-lightDir.y, 0.0f, 0.1f) / 0.1f);
}
