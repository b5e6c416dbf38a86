// C ABI, sun-shadow half (include/vkx.h): host-side orchestration of shadow.cu.
#include "common.cuh"
#include <cstring>

#define BIND(ctx) do { if (!(ctx)) return VKX_E_INVALID; cudaError_t _e = cudaSetDevice((ctx)->device); if (_e != cudaSuccess) return vkx_fail((ctx), VKX_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(_e)); } while (0)

extern "C" {

int vkx_shadow_set_noise(vkx_ctx* ctx, const float* rgba, uint32_t w, uint32_t h, uint32_t slices) {
    BIND(ctx);
    if (!rgba || !w || !h || !slices) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shadow_set_noise: bad arguments");
    if (ctx->dNoise) { cudaFree(ctx->dNoise); ctx->dNoise = nullptr; }
    const size_t bytes = size_t(w) * h * slices * 16;
    CUDA_TRY(ctx, cudaMalloc(&ctx->dNoise, bytes));
    CUDA_TRY(ctx, cudaMemcpy(ctx->dNoise, rgba, bytes, cudaMemcpyHostToDevice));
    ctx->noiseW = w; ctx->noiseH = h; ctx->noiseSlices = slices;
    return VKX_OK;
}

int vkx_shadow_init(vkx_ctx* ctx, uint32_t width, uint32_t height) {
    BIND(ctx);
    if (!width || !height) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shadow_init: empty image");
    void* old[] = {ctx->dPosDepth, ctx->dNormalMetal, ctx->dShRaw, ctx->dShX, ctx->dShFinal[0], ctx->dShFinal[1], ctx->dShDirs, ctx->dShMask, ctx->dAlbedoRough, ctx->dEmissive, ctx->dReflection, ctx->dGathered,
                   ctx->dReflRaw, ctx->dReflX, ctx->dReflFinal[0], ctx->dReflFinal[1], ctx->dReflDirs, ctx->dReflHits, ctx->dReflMask, ctx->dReflQueue, ctx->dReflCount};
    for (void* p : old) if (p) cudaFree(p);
    ctx->dPosDepth = ctx->dNormalMetal = ctx->dShRaw = ctx->dShX = ctx->dShFinal[0] = ctx->dShFinal[1] = ctx->dShDirs = nullptr; ctx->dShMask = nullptr;
    ctx->dAlbedoRough = ctx->dEmissive = ctx->dReflection = ctx->dGathered = nullptr;
    ctx->dReflRaw = ctx->dReflX = ctx->dReflFinal[0] = ctx->dReflFinal[1] = ctx->dReflDirs = nullptr; ctx->dReflHits = nullptr; ctx->dReflMask = nullptr; ctx->dReflQueue = ctx->dReflCount = nullptr; ctx->reflValid = false; ctx->reflCur = 0;
    const size_t px = size_t(width) * height;
    float4** imgs[] = {&ctx->dPosDepth, &ctx->dNormalMetal, &ctx->dShRaw, &ctx->dShX, &ctx->dShFinal[0], &ctx->dShFinal[1], &ctx->dShDirs, &ctx->dAlbedoRough, &ctx->dEmissive, &ctx->dGathered,
                       &ctx->dReflRaw, &ctx->dReflX, &ctx->dReflFinal[0], &ctx->dReflFinal[1], &ctx->dReflDirs};
    for (float4** p : imgs) { CUDA_TRY(ctx, cudaMalloc(p, px * 16)); CUDA_TRY(ctx, cudaMemsetAsync(*p, 0, px * 16, ctx->stream)); }
    CUDA_TRY(ctx, cudaMalloc(&ctx->dShMask, px)); CUDA_TRY(ctx, cudaMemsetAsync(ctx->dShMask, 0, px, ctx->stream));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dReflQueue, px * 4)); CUDA_TRY(ctx, cudaMalloc(&ctx->dReflCount, 4));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dReflMask, px)); CUDA_TRY(ctx, cudaMemsetAsync(ctx->dReflMask, 0, px, ctx->stream));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dReflHits, px * sizeof(vkx_hit))); CUDA_TRY(ctx, cudaMemsetAsync(ctx->dReflHits, 0, px * sizeof(vkx_hit), ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->shW = width; ctx->shH = height; ctx->shCur = 0;
    return VKX_OK;
}

int vkx_gbuffer_generate(vkx_ctx* ctx, const vkx_camera* cam) {
    BIND(ctx);
    if (!cam || !ctx->shW || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_gbuffer_generate: shadow images or BVH not ready");
    int rc = shadowGBuffer(ctx, *cam);
    if (rc != VKX_OK) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_gbuffer_upload(vkx_ctx* ctx, const float* positionDepth, const float* normalMetalness) {
    BIND(ctx);
    if (!ctx->shW || !positionDepth || !normalMetalness) return vkx_fail(ctx, VKX_E_INVALID, "vkx_gbuffer_upload: bad arguments");
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(ctx->dPosDepth, positionDepth, bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->dNormalMetal, normalMetalness, bytes, cudaMemcpyHostToDevice));
    return VKX_OK;
}

int vkx_gbuffer_download(vkx_ctx* ctx, float* positionDepth, float* normalMetalness) {
    BIND(ctx);
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "shadow images not initialised");
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (positionDepth) CUDA_TRY(ctx, cudaMemcpy(positionDepth, ctx->dPosDepth, bytes, cudaMemcpyDeviceToHost));
    if (normalMetalness) CUDA_TRY(ctx, cudaMemcpy(normalMetalness, ctx->dNormalMetal, bytes, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_gbuffer_upload_material(vkx_ctx* ctx, const float* albedoRoughness, const float* emissive) {
    BIND(ctx);
    if (!ctx->shW || !albedoRoughness || !emissive) return vkx_fail(ctx, VKX_E_INVALID, "vkx_gbuffer_upload_material: bad arguments");
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(ctx->dAlbedoRough, albedoRoughness, bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->dEmissive, emissive, bytes, cudaMemcpyHostToDevice));
    return VKX_OK;
}

int vkx_gbuffer_download_material(vkx_ctx* ctx, float* albedoRoughness, float* emissive) {
    BIND(ctx);
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "shadow images not initialised");
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (albedoRoughness) CUDA_TRY(ctx, cudaMemcpy(albedoRoughness, ctx->dAlbedoRough, bytes, cudaMemcpyDeviceToHost));
    if (emissive) CUDA_TRY(ctx, cudaMemcpy(emissive, ctx->dEmissive, bytes, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_final_gather(vkx_ctx* ctx, const vkx_camera* cam, const vkx_light* light, const float* reflection, int sync) {
    BIND(ctx);
    if (!cam || !light) return vkx_fail(ctx, VKX_E_INVALID, "vkx_final_gather: null argument");
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "vkx_final_gather: call vkx_shadow_init first");
    if (!ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "vkx_final_gather: call vkx_probes_init first");
    int rc = waitGather(ctx); // a sharded update may still be gathering the atlases this pass samples
    if (rc != VKX_OK) return rc;
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    if (reflection) {
        if (!ctx->dReflection) CUDA_TRY(ctx, cudaMalloc(&ctx->dReflection, bytes));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dReflection, reflection, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!reflection && ctx->reflValid) { // the filtered result of the last vkx_reflection_frame
        float4* keep = ctx->dReflection; ctx->dReflection = ctx->dReflFinal[ctx->reflCur];
        rc = finalGather(ctx, *cam, *light, true);
        ctx->dReflection = keep;
    } else rc = finalGather(ctx, *cam, *light, reflection != nullptr);
    if (rc != VKX_OK) return rc;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_final_gather_download(vkx_ctx* ctx, float* rgba, float* ms) {
    BIND(ctx);
    if (!ctx->shW || !ctx->gev[0]) return vkx_fail(ctx, VKX_E_INVALID, "vkx_final_gather_download: nothing gathered yet");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (rgba) CUDA_TRY(ctx, cudaMemcpy(rgba, ctx->dGathered, size_t(ctx->shW) * ctx->shH * 16, cudaMemcpyDeviceToHost));
    if (ms) CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->gev[0], ctx->gev[1]));
    return VKX_OK;
}

int vkx_reflection_frame(vkx_ctx* ctx, const vkx_camera* cur, const vkx_camera* prev, const vkx_light* light, int sync) {
    BIND(ctx);
    if (!cur || !prev || !light) return vkx_fail(ctx, VKX_E_INVALID, "vkx_reflection_frame: null argument");
    if (!ctx->shW || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_reflection_frame: screen images (vkx_shadow_init) or BVH not ready");
    if (!ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "vkx_reflection_frame: call vkx_probes_init first (hits are shaded from the irradiance volume)");
    if (!ctx->dNoise) return vkx_fail(ctx, VKX_E_INVALID, "vkx_reflection_frame: call vkx_shadow_set_noise first");
    int rc = waitGather(ctx);
    if (rc != VKX_OK) return rc;
    rc = reflectionFrame(ctx, *cur, *prev, *light);
    if (rc != VKX_OK) return rc;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_reflection_download(vkx_ctx* ctx, int stage, float* rgba) {
    BIND(ctx);
    if (!ctx->shW || !rgba) return vkx_fail(ctx, VKX_E_INVALID, "vkx_reflection_download: bad arguments");
    const float4* src = stage == 0 ? ctx->dReflRaw : stage == 1 ? ctx->dReflX : ctx->dReflFinal[ctx->reflCur];
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(rgba, src, size_t(ctx->shW) * ctx->shH * 16, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_reflection_download_debug(vkx_ctx* ctx, float* dirs4, vkx_hit* hits, uint8_t* mask) {
    BIND(ctx);
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "screen images not initialised");
    if ((hits || mask) && !ctx->debugBuffers) return vkx_fail(ctx, VKX_E_INVALID, "enable vkx_probes_debug before the reflection frame (hit / mask side buffers)");
    const size_t px = size_t(ctx->shW) * ctx->shH;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (dirs4) CUDA_TRY(ctx, cudaMemcpy(dirs4, ctx->dReflDirs, px * 16, cudaMemcpyDeviceToHost));
    if (hits) CUDA_TRY(ctx, cudaMemcpy(hits, ctx->dReflHits, px * sizeof(vkx_hit), cudaMemcpyDeviceToHost));
    if (mask) CUDA_TRY(ctx, cudaMemcpy(mask, ctx->dReflMask, px, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_reflection_reset_history(vkx_ctx* ctx) {
    BIND(ctx);
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "screen images not initialised");
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dReflFinal[0], 0, bytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dReflFinal[1], 0, bytes, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->reflValid = false;
    return VKX_OK;
}

int vkx_reflection_timings(vkx_ctx* ctx, float ms[4]) {
    BIND(ctx);
    if (!ms || !ctx->rev[0]) return vkx_fail(ctx, VKX_E_INVALID, "vkx_reflection_timings: no reflection frame yet");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[0], ctx->rev[0], ctx->rev[3]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[1], ctx->rev[0], ctx->rev[1]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[2], ctx->rev[1], ctx->rev[2]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[3], ctx->rev[2], ctx->rev[3]));
    return VKX_OK;
}

int vkx_shadow_frame(vkx_ctx* ctx, const vkx_camera* cur, const vkx_camera* prev, const vkx_light* light, int sync) {
    BIND(ctx);
    if (!cur || !prev || !light) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shadow_frame: null argument");
    if (!ctx->shW || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shadow_frame: shadow images or BVH not ready");
    if (!ctx->dNoise) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shadow_frame: call vkx_shadow_set_noise first");
    int rc = shadowFrame(ctx, *cur, *prev, *light);
    if (rc != VKX_OK) return rc;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_shadow_download(vkx_ctx* ctx, int stage, float* rgba) {
    BIND(ctx);
    if (!ctx->shW || !rgba) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shadow_download: bad arguments");
    const float4* src = stage == 0 ? ctx->dShRaw : stage == 1 ? ctx->dShX : ctx->dShFinal[ctx->shCur];
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(rgba, src, size_t(ctx->shW) * ctx->shH * 16, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

/* Parity side buffers of the last frame: jittered light directions [h][w][4] and the mask (0 not traced, 1 lit, 2 shadowed). */
int vkx_shadow_download_debug(vkx_ctx* ctx, float* dirs4, uint8_t* mask) {
    BIND(ctx);
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "shadow images not initialised");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (dirs4) CUDA_TRY(ctx, cudaMemcpy(dirs4, ctx->dShDirs, size_t(ctx->shW) * ctx->shH * 16, cudaMemcpyDeviceToHost));
    if (mask) CUDA_TRY(ctx, cudaMemcpy(mask, ctx->dShMask, size_t(ctx->shW) * ctx->shH, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_shadow_reset_history(vkx_ctx* ctx) {
    BIND(ctx);
    if (!ctx->shW) return vkx_fail(ctx, VKX_E_INVALID, "shadow images not initialised");
    const size_t bytes = size_t(ctx->shW) * ctx->shH * 16;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dShFinal[0], 0, bytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dShFinal[1], 0, bytes, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_shadow_timings(vkx_ctx* ctx, float ms[4]) {
    BIND(ctx);
    if (!ms) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[0], ctx->sev[0], ctx->sev[3]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[1], ctx->sev[0], ctx->sev[1]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[2], ctx->sev[1], ctx->sev[2]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[3], ctx->sev[2], ctx->sev[3]));
    return VKX_OK;
}

} // extern "C"
