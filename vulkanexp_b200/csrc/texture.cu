// Scene textures on the device: upload, mip-chain generation and the parity entry points of "sampler spec v1" (DESIGN.md 9f).
// Replaces (reference): uploadTextures (src/Resources.cpp:46-95) -> Image::upload (src/vulkan/Image.cpp:62-111) ->
// Image::generateMipmaps (src/vulkan/Image.cpp:195-275: levels = floor(log2(max(w, h))) + 1, level i = LINEAR vkCmdBlitImage of level
// i-1 into max(1, w/2) x max(1, h/2)) and the sampler creation of src/Resources.cpp:97-124.
// Every float operation of the blit is a single _rn operation in the order of oracle/texture.h::blitHalf, so the generated texels
// are bit-identical to the oracle's.
#include <algorithm>
#include <cmath>
#include "common.cuh"
#include "texture.cuh"

namespace {

// encode to the nearest 8-bit sRGB code of the exact transfer function: number of thresholds <= x (threshold[k] = linear((k - 0.5) / 255))
__device__ __forceinline__ uint32_t srgbEncode(const float* __restrict__ threshold, float x) {
    if (!(x > 0.0f)) return 0u;
    uint32_t lo = 0u, hi = 255u;
    while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (__ldg(threshold + mid) <= x) lo = mid; else hi = mid - 1u; }
    return lo;
}
__device__ __forceinline__ uint32_t unormEncode(float x) { x = fminf(fmaxf(x, 0.0f), 1.0f); return uint32_t(__fadd_rn(__fmul_rn(x, 255.0f), 0.5f)); }

// One LINEAR blit: whole source level -> whole destination level, edge texels clamped (Vulkan blit equations).
__global__ void k_mip_blit(uint32_t* __restrict__ texels, uint32_t srcOff, uint32_t sw, uint32_t sh, uint32_t dstOff, uint32_t dw, uint32_t dh, uint32_t flags,
                           const float* __restrict__ srgbLut, const float* __restrict__ threshold) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= dw || j >= dh) return;
    const float scaleU = __fdiv_rn(float(sw), float(dw)), scaleV = __fdiv_rn(float(sh), float(dh));
    const float u = __fsub_rn(__fmul_rn(__fadd_rn(float(i), 0.5f), scaleU), 0.5f), v = __fsub_rn(__fmul_rn(__fadd_rn(float(j), 0.5f), scaleV), 0.5f);
    const float fu = floorf(u), fv = floorf(v);
    const float a = __fsub_rn(u, fu), b = __fsub_rn(v, fv);
    const int i0 = int(fu), j0 = int(fv);
    auto at = [&](int x, int y) {
        x = min(max(x, 0), int(sw) - 1); y = min(max(y, 0), int(sh) - 1);
        return texDecode(srgbLut, flags, texels[size_t(srcOff) + size_t(y) * sw + size_t(x)]);
    };
    const float4 top = texLerp4(at(i0, j0), at(i0 + 1, j0), a), bot = texLerp4(at(i0, j0 + 1), at(i0 + 1, j0 + 1), a);
    const float4 c = texLerp4(top, bot, b);
    uint32_t r, g, bl;
    if (flags & VKX_TEX_SRGB) { r = srgbEncode(threshold, c.x); g = srgbEncode(threshold, c.y); bl = srgbEncode(threshold, c.z); }
    else { r = unormEncode(c.x); g = unormEncode(c.y); bl = unormEncode(c.z); }
    texels[size_t(dstOff) + size_t(j) * dw + i] = r | (g << 8) | (bl << 16) | (unormEncode(c.w) << 24);
}

// Decoded copy of a texture's whole mip chain: the values texDecode yields, stored once (sampling is 4x the memory of RGBA8, with
// 180 GB of HBM that buys one 128-bit load per bilinear tap instead of a word and four table look-ups).
__global__ void k_decode_texels(const uint32_t* __restrict__ texels, float4* __restrict__ decoded, uint32_t first, uint32_t count, uint32_t flags, const float* __restrict__ lut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) decoded[first + i] = texDecode(lut, flags, texels[first + i]);
}

__global__ void k_texture_sample(DeviceScene sc, uint32_t index, const float* __restrict__ uv, const float* __restrict__ grads, uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = grads ? texSampleGrad(sc, index, uv[2 * i], uv[2 * i + 1], grads[4 * i], grads[4 * i + 1], grads[4 * i + 2], grads[4 * i + 3])
                   : texSampleBase(sc, index, uv[2 * i], uv[2 * i + 1]);
}

double srgbDecode(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }

uint32_t samplerFlags(const vkx_texture& d) {
    const uint32_t mag = d.magFilter ? d.magFilter : 9729u, mn = d.minFilter ? d.minFilter : 9729u; // defaults of src/Resources.cpp:88-90
    uint32_t f = d.srgb ? VKX_TEX_SRGB : 0u;
    if (mag == 9729u || mag == 9987u) f |= VKX_TEX_MAG_LINEAR;                 // glTFToVkFilter, src/Resources.cpp:8-19
    if (mn == 9729u || mn == 9987u) f |= VKX_TEX_MIN_LINEAR;
    if (mn == 9729u || mn == 9986u || mn == 9987u) f |= VKX_TEX_MIP_LINEAR;    // glTFToVkSamplerMipmapMode, :21-32
    auto wrap = [](uint32_t e) { return e == 33071u ? 1u : e == 33648u ? 2u : 0u; }; // glTFtoVkSamplerAddressMode, :34-43
    return f | (wrap(d.wrapS) << VKX_TEX_WRAP_S_SHIFT) | (wrap(d.wrapT) << VKX_TEX_WRAP_T_SHIFT);
}

} // namespace

void freeTextures(vkx_ctx* ctx) {
    if (ctx->dTexels) cudaFree(ctx->dTexels);
    if (ctx->dTextures) cudaFree(ctx->dTextures);
    if (ctx->dTexelsDecoded) cudaFree(ctx->dTexelsDecoded);
    ctx->dTexels = nullptr; ctx->dTextures = nullptr; ctx->dTexelsDecoded = nullptr; ctx->hTextures.clear(); ctx->numTexels = 0;
}

extern "C" int vkx_scene_textures(vkx_ctx* ctx, const vkx_texture* textures, size_t numTextures) {
    if (!ctx) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (numTextures && !textures) return vkx_fail(ctx, VKX_E_INVALID, "vkx_scene_textures: null array");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<DeviceTexture> desc(numTextures);
    size_t total = 0;
    for (size_t t = 0; t < numTextures; ++t) {
        const vkx_texture& d = textures[t];
        if (!d.pixels || d.width == 0 || d.height == 0) return vkx_fail(ctx, VKX_E_INVALID, "texture %zu: empty image", t);
        if (d.width > VKX_MAX_TEXTURE_SIZE || d.height > VKX_MAX_TEXTURE_SIZE) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "texture %zu: %u x %u exceeds %u texels on a side", t, d.width, d.height, VKX_MAX_TEXTURE_SIZE);
        DeviceTexture& o = desc[t];
        o.width = d.width; o.height = d.height; o.flags = samplerFlags(d);
        uint32_t m = std::max(d.width, d.height), levels = 1; while (m > 1) { m >>= 1; ++levels; } // src/vulkan/Image.cpp:29-31
        o.levels = levels;
        for (uint32_t l = 0; l < 16; ++l) o.levelOffset[l] = 0;
        for (uint32_t l = 0; l < levels; ++l) {
            if (total >= (1ull << 32)) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "texture arena exceeds 2^32 texels");
            o.levelOffset[l] = uint32_t(total);
            total += size_t(std::max(1u, d.width >> l)) * std::max(1u, d.height >> l);
        }
    }
    if (total >= (1ull << 32)) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "texture arena exceeds 2^32 texels");
    if (numTextures < ctx->texturesUsed) { // the uploaded materials point past the new list: that scene is gone until the next vkx_scene_upload
        ctx->numInstances = 0; ctx->numFlatTris = 0; ctx->texturesUsed = 0; ctx->bvhBuilt = false;
    }
    freeTextures(ctx);
    if (!ctx->dSrgbLut) { // T4: exact transfer function per code, evaluated in double precision
        float lut[512], thr[256]; // lut: sRGB -> linear, then code / 255 (texDecode)
        for (int i = 0; i < 256; ++i) { lut[i] = float(srgbDecode(double(i) / 255.0)); lut[256 + i] = float(i) / 255.0f; thr[i] = i == 0 ? 0.0f : float(srgbDecode((double(i) - 0.5) / 255.0)); }
        CUDA_TRY(ctx, cudaMalloc(&ctx->dSrgbLut, sizeof(lut))); CUDA_TRY(ctx, cudaMalloc(&ctx->dSrgbThreshold, sizeof(thr)));
        CUDA_TRY(ctx, cudaMemcpy(ctx->dSrgbLut, lut, sizeof(lut), cudaMemcpyHostToDevice)); CUDA_TRY(ctx, cudaMemcpy(ctx->dSrgbThreshold, thr, sizeof(thr), cudaMemcpyHostToDevice));
    }
    if (numTextures == 0) return VKX_OK;
    CUDA_TRY(ctx, cudaMalloc(&ctx->dTexels, total * 4));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dTexelsDecoded, total * 16));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dTextures, numTextures * sizeof(DeviceTexture)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dTextures, desc.data(), numTextures * sizeof(DeviceTexture), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t t = 0; t < numTextures; ++t) {
        const DeviceTexture& o = desc[t];
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dTexels + o.levelOffset[0], textures[t].pixels, size_t(o.width) * o.height * 4, cudaMemcpyHostToDevice, ctx->stream));
        for (uint32_t l = 1; l < o.levels; ++l) {
            const uint32_t sw = std::max(1u, o.width >> (l - 1)), sh = std::max(1u, o.height >> (l - 1)), dw = std::max(1u, o.width >> l), dh = std::max(1u, o.height >> l);
            k_mip_blit<<<dim3(divUp(dw, 16), divUp(dh, 16)), dim3(16, 16), 0, ctx->stream>>>(ctx->dTexels, o.levelOffset[l - 1], sw, sh, o.levelOffset[l], dw, dh, o.flags, ctx->dSrgbLut, ctx->dSrgbThreshold);
            LAUNCH_CHECK(ctx);
        }
        const size_t first = o.levelOffset[0], count = (t + 1 < numTextures ? size_t(desc[t + 1].levelOffset[0]) : total) - first;
        k_decode_texels<<<divUp(count, 256), 256, 0, ctx->stream>>>(ctx->dTexels, ctx->dTexelsDecoded, uint32_t(first), uint32_t(count), o.flags, ctx->dSrgbLut);
        LAUNCH_CHECK(ctx);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // the caller may free the images now
    ctx->hTextures = desc; ctx->numTexels = total;
    return VKX_OK;
}

extern "C" int vkx_texture_download(vkx_ctx* ctx, uint32_t texture, void* texels, size_t texelsBytes, uint32_t* numLevels) {
    if (!ctx) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (texture >= ctx->hTextures.size()) return vkx_fail(ctx, VKX_E_INVALID, "vkx_texture_download: texture %u of %zu", texture, ctx->hTextures.size());
    const DeviceTexture& o = ctx->hTextures[texture];
    if (numLevels) *numLevels = o.levels;
    if (!texels) return VKX_OK;
    size_t count = 0;
    for (uint32_t l = 0; l < o.levels; ++l) count += size_t(std::max(1u, o.width >> l)) * std::max(1u, o.height >> l);
    if (texelsBytes < count * 4) return vkx_fail(ctx, VKX_E_INVALID, "vkx_texture_download: buffer too small (%zu < %zu bytes)", texelsBytes, count * 4);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(texels, ctx->dTexels + o.levelOffset[0], count * 4, cudaMemcpyDeviceToHost)); // the levels of a texture are contiguous
    return VKX_OK;
}

extern "C" int vkx_texture_sample(vkx_ctx* ctx, uint32_t texture, const float* uv, const float* grads, size_t n, float* out) {
    if (!ctx) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (texture >= ctx->hTextures.size()) return vkx_fail(ctx, VKX_E_INVALID, "vkx_texture_sample: texture %u of %zu", texture, ctx->hTextures.size());
    if (n == 0) return VKX_OK;
    if (!uv || !out || n > 0x7FFFFFFFu) return vkx_fail(ctx, VKX_E_INVALID, "vkx_texture_sample: bad arguments");
    float *dUv = nullptr, *dGrads = nullptr; float4* dOut = nullptr;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = cudaMalloc(&dUv, n * 8)) != cudaSuccess) break;
        if ((e = cudaMalloc(&dOut, n * 16)) != cudaSuccess) break;
        if (grads && (e = cudaMalloc(&dGrads, n * 16)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dUv, uv, n * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
        if (grads && (e = cudaMemcpyAsync(dGrads, grads, n * 16, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) break;
        k_texture_sample<<<divUp(n, 128), 128, 0, ctx->stream>>>(deviceScene(ctx), texture, dUv, dGrads, uint32_t(n), dOut);
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(out, dOut, n * 16, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
        e = cudaStreamSynchronize(ctx->stream);
    } while (0);
    cudaFree(dUv); cudaFree(dGrads); cudaFree(dOut);
    if (e != cudaSuccess) return vkx_fail(ctx, VKX_E_CUDA, "vkx_texture_sample: %s", cudaGetErrorString(e));
    return VKX_OK;
}
