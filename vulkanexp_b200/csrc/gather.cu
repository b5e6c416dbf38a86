// Final composite of a frame: the consumer of both hot-path outputs (filtered sun shadow + irradiance volume).
// Replaces (reference): src/shaders/FinalGather.frag:38-77 with FullScreenQuad.vert (fragPosition = pixel centre / extent),
// drawn by src/SwapchainManagement.cpp:466-474. One thread per pixel; empty pixels get the sky, the others
// direct * shadow + specular * reflection + sampleProbes * diffuse + emissive. Compiled with the shading flags of
// ddgi_shade.cu (smooth functions, tolerance 1e-3); the quotient that selects the 8 probes stays exact (sampleProbes1).
#include "common.cuh"
#include "shade.cuh"

namespace {

struct M4g { float m[16]; }; // column-major

__device__ __forceinline__ float4 mulM4g(const M4g& M, float x, float y, float z, float w) {
    float4 r;
    r.x = (M.m[0] * x + M.m[4] * y) + (M.m[8] * z + M.m[12] * w);
    r.y = (M.m[1] * x + M.m[5] * y) + (M.m[9] * z + M.m[13] * w);
    r.z = (M.m[2] * x + M.m[6] * y) + (M.m[10] * z + M.m[14] * w);
    r.w = (M.m[3] * x + M.m[7] * y) + (M.m[11] * z + M.m[15] * w);
    return r;
}

// 2D tiles of 16 x 8 pixels: neighbouring pixels shade from the same 8 probes, so a warp's atlas fetches stay in a few tiles
__global__ void __launch_bounds__(128) k_final_gather(DeviceProbes pr, vkx_light light, M4g invView, M4g invProj, uint32_t W, uint32_t H,
                                                      const float4* __restrict__ posDepth, const float4* __restrict__ normalMetal, const float4* __restrict__ albedoRough,
                                                      const float4* __restrict__ emissive, const float4* __restrict__ directLight, const float4* __restrict__ reflection,
                                                      float4* __restrict__ out) {
    const uint32_t x = blockIdx.x * 16u + (threadIdx.x & 15u), y = blockIdx.y * 8u + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    const size_t pix = size_t(y) * W + x;
    const float4 o4 = mulM4g(invView, 0.f, 0.f, 0.f, 1.f);
    const v3 origin = mk3(o4.x, o4.y, o4.z);
    const v3 lightDir = mk3(light.direction[0], light.direction[1], light.direction[2]);
    const v3 lightColor = mk3(light.color[0], light.color[1], light.color[2]);
    const float4 pd = __ldg(posDepth + pix);
    v3 color = mk3(0.0f);
    if (pd.w <= 0.0f) {
        const float fx = (float(x) + 0.5f) / float(W), fy = (float(y) + 0.5f) / float(H);
        const float4 t = mulM4g(invProj, 2.0f * fx - 1.0f, 2.0f * fy - 1.0f, 0.0f, 1.0f);
        const v3 tn = norm3(mk3(t.x, t.y, t.z));
        const float4 d4 = mulM4g(invView, tn.x, tn.y, tn.z, 0.0f);
        color = skyColor(origin, mk3(d4.x, d4.y, d4.z), lightDir, lightColor, 1.0f);
    } else {
        const float4 nm = __ldg(normalMetal + pix), ar = __ldg(albedoRough + pix), em = __ldg(emissive + pix);
        const v3 position = mk3(pd.x, pd.y, pd.z);
        const v3 normal = norm3(mk3(nm.x, nm.y, nm.z));
        const float metalness = nm.w, roughness = ar.w;
        const v3 albedo = mk3(ar.x, ar.y, ar.z);
        const v3 view = norm3(origin - position);
        const float direct = __ldg(directLight + pix).x;
        color = color + direct * pbrMetallicRoughness(normal, view, lightColor, lightDir, albedo, metalness, roughness);
        const v3 f0 = mk3(0.004f);
        v3 diffuseColor = albedo * (1.0f - f0);
        diffuseColor = diffuseColor * (1.0f - metalness);
        const v3 specularColor = mix3(f0, albedo, metalness);
        if (reflection) { const float4 rf = __ldg(reflection + pix); color = color + specularColor * mk3(rf.x, rf.y, rf.z); }
        const GridConsts gc = makeGridConsts(pr.grid);
        const v3 indirect = sampleProbes1(pr, gc, position, normal, view);
        color = color + indirect * diffuseColor;
        color = color + mk3(em.x, em.y, em.z);
    }
    out[pix] = make_float4(color.x, color.y, color.z, 1.0f);
}

void inverse4g(const float* a, float* out) { // cofactor expansion, fp32 (same sequence as the oracle's inverse4)
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    const float id = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
}

} // namespace

int finalGather(vkx_ctx* ctx, const vkx_camera& cam, const vkx_light& light, bool haveReflection) {
    M4g iv, ip;
    inverse4g(cam.view, iv.m); inverse4g(cam.proj, ip.m);
    const uint32_t W = ctx->shW, H = ctx->shH;
    if (!ctx->gev[0]) for (auto& e : ctx->gev) CUDA_TRY(ctx, cudaEventCreate(&e));
    CUDA_TRY(ctx, cudaEventRecord(ctx->gev[0], ctx->stream));
    k_final_gather<<<dim3(divUp(W, 16), divUp(H, 8)), 128, 0, ctx->stream>>>(deviceProbes(ctx), light, iv, ip, W, H, ctx->dPosDepth, ctx->dNormalMetal, ctx->dAlbedoRough,
                                                                          ctx->dEmissive, ctx->dShFinal[ctx->shCur], haveReflection ? ctx->dReflection : nullptr, ctx->dGathered);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->gev[1], ctx->stream));
    return VKX_OK;
}
