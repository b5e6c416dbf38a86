// Vertex skinning on the device + the hook for the skinned-geometry rebuild.
// Replaces (reference): src/shaders/vertexSkinning.comp:37-60 as dispatched by Renderer::updateSkinnedVertexBuffer
// (src/Renderer.cpp:201-240), followed by Renderer::updateSkinnedBLAS (src/Renderer.cpp:644-669). The reference rebuilds the skinned
// BLASes in place; here all geometry lives in one world-space BVH, so the structure is marked stale and the next vkx_bvh_build
// rebuilds it (the deterministic build: the result equals a fresh upload of the skinned vertices).
// The shader's quirk is kept: normals / tangents are read from the *destination* range (the bind-pose copy made by
// Renderer::allocateSkinnedMeshes) and written to the *source* range (Output[VertexStride * (i + srcOffset) + 1..2], :54-57), so the
// skinned instance keeps bind-pose normals and the source mesh receives the skinned ones.
// Every float operation is a single _rn operation in the order of oracle/ddgi.cpp::skinVertices (bit-identical vertices).
#include "common.cuh"

namespace {

__global__ void k_vertex_skinning(vkx_vertex* __restrict__ vertices, const float* __restrict__ jointTransforms, const uint16_t* __restrict__ skinJoints,
                                  const float4* __restrict__ skinWeights, uint32_t srcOffset, uint32_t dstOffset, uint32_t size, float4* __restrict__ motionVectors) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const float4 w = skinWeights[i];
    const float* J0 = jointTransforms + 16u * skinJoints[4 * i + 0];
    const float* J1 = jointTransforms + 16u * skinJoints[4 * i + 1];
    const float* J2 = jointTransforms + 16u * skinJoints[4 * i + 2];
    const float* J3 = jointTransforms + 16u * skinJoints[4 * i + 3];
    float m[16]; // column-major: m[4 * col + row]
#pragma unroll
    for (int e = 0; e < 16; ++e)
        m[e] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w.x, J0[e]), __fmul_rn(w.y, J1[e])), __fmul_rn(w.z, J2[e])), __fmul_rn(w.w, J3[e]));
    vkx_vertex* src = vertices + srcOffset + i;
    vkx_vertex* dst = vertices + dstOffset + i;
    const float px = src->pos[0], py = src->pos[1], pz = src->pos[2];
    float np[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) // mat4 * vec4(p, 1): (m0 * x + m1 * y) + (m2 * z + m3 * 1)
        np[r] = __fadd_rn(__fadd_rn(__fmul_rn(m[r], px), __fmul_rn(m[4 + r], py)), __fadd_rn(__fmul_rn(m[8 + r], pz), __fmul_rn(m[12 + r], 1.0f)));
    const float mvx = __fsub_rn(np[0], dst->pos[0]), mvy = __fsub_rn(np[1], dst->pos[1]), mvz = __fsub_rn(np[2], dst->pos[2]);
    const float nx = dst->normal[0], ny = dst->normal[1], nz = dst->normal[2];
    const float tx = dst->tangent[0], ty = dst->tangent[1], tz = dst->tangent[2];
    float nn[3], tt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { // mat3(skinMatrix) * v: (m0 * x + m1 * y) + m2 * z
        nn[r] = __fadd_rn(__fadd_rn(__fmul_rn(m[r], nx), __fmul_rn(m[4 + r], ny)), __fmul_rn(m[8 + r], nz));
        tt[r] = __fadd_rn(__fadd_rn(__fmul_rn(m[r], tx), __fmul_rn(m[4 + r], ty)), __fmul_rn(m[8 + r], tz));
    }
    dst->pos[0] = np[0]; dst->pos[1] = np[1]; dst->pos[2] = np[2];
    src->normal[0] = nn[0]; src->normal[1] = nn[1]; src->normal[2] = nn[2];
    src->tangent[0] = tt[0]; src->tangent[1] = tt[1]; src->tangent[2] = tt[2];
    if (motionVectors) motionVectors[i] = make_float4(mvx, mvy, mvz, 1.0f);
}

} // namespace

extern "C" int vkx_skin_vertices(vkx_ctx* ctx, const float* jointTransforms, size_t numJoints, const uint16_t* skinJoints, const float* skinWeights,
                                 uint32_t srcOffset, uint32_t dstOffset, uint32_t size, float* motionVectors) {
    if (!ctx) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (size == 0) return VKX_OK;
    if (!jointTransforms || !skinJoints || !skinWeights || numJoints == 0) return vkx_fail(ctx, VKX_E_INVALID, "vkx_skin_vertices: null array");
    if (size_t(srcOffset) + size > ctx->numVertices || size_t(dstOffset) + size > ctx->numVertices) return vkx_fail(ctx, VKX_E_INVALID, "vkx_skin_vertices: vertex range out of bounds (%zu vertices uploaded)", ctx->numVertices);
    const bool overlap = srcOffset < dstOffset ? srcOffset + size > dstOffset : dstOffset + size > srcOffset;
    if (overlap) return vkx_fail(ctx, VKX_E_INVALID, "vkx_skin_vertices: source and destination ranges overlap");
    for (size_t k = 0; k < size_t(size) * 4; ++k) if (skinJoints[k] >= numJoints) return vkx_fail(ctx, VKX_E_INVALID, "vkx_skin_vertices: joint index %u of vertex %zu out of range (%zu joints)", unsigned(skinJoints[k]), k / 4, numJoints);
    float* dJ = nullptr; uint16_t* dI = nullptr; float4* dW = nullptr; float4* dM = nullptr;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = cudaMalloc(&dJ, numJoints * 64)) != cudaSuccess) break;
        if ((e = cudaMalloc(&dI, size_t(size) * 8)) != cudaSuccess) break;
        if ((e = cudaMalloc(&dW, size_t(size) * 16)) != cudaSuccess) break;
        if (motionVectors && (e = cudaMalloc(&dM, size_t(size) * 16)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dJ, jointTransforms, numJoints * 64, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dI, skinJoints, size_t(size) * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dW, skinWeights, size_t(size) * 16, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
        k_vertex_skinning<<<divUp(size, 128), 128, 0, ctx->stream>>>(ctx->dVertices, dJ, dI, dW, srcOffset, dstOffset, size, dM); // local_size_x = 128 in the shader, too
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if (motionVectors && (e = cudaMemcpyAsync(motionVectors, dM, size_t(size) * 16, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
        e = cudaStreamSynchronize(ctx->stream);
    } while (0);
    cudaFree(dJ); cudaFree(dI); cudaFree(dW); cudaFree(dM);
    if (e != cudaSuccess) return vkx_fail(ctx, VKX_E_CUDA, "vkx_skin_vertices: %s", cudaGetErrorString(e));
    ctx->bvhBuilt = false; // updateSkinnedBLAS: the next vkx_bvh_build picks the new positions up
    return VKX_OK;
}

extern "C" int vkx_vertices_download(vkx_ctx* ctx, size_t firstVertex, size_t count, vkx_vertex* out) {
    if (!ctx) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (firstVertex + count > ctx->numVertices || (count && !out)) return vkx_fail(ctx, VKX_E_INVALID, "vkx_vertices_download: range out of bounds");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (count) CUDA_TRY(ctx, cudaMemcpy(out, ctx->dVertices + firstVertex, count * sizeof(vkx_vertex), cudaMemcpyDeviceToHost));
    return VKX_OK;
}
