// vkx_trace: parity primitive that pushes host rays through the device BVH (closest-hit or any-hit).
#include "common.cuh"
#include "traverse.cuh"

namespace {
template <bool ANY, bool ALPHA>
__global__ void __launch_bounds__(128) k_trace_host(DeviceScene sc, const float* __restrict__ o, const float* __restrict__ d, uint32_t n, float tmin, float tmax,
                                                    uint32_t cullMask, vkx_hit* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ray r = makeRay(o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    HitRec h;
    const bool hit = traverse<ANY, ALPHA>(sc.nodes, sc.tris, r, tmin, tmax, cullMask, h, &sc);
    vkx_hit res;
    if (ANY) { res.t = hit ? 1.0f : -1.0f; res.instance = 0xFFFFFFFFu; res.primitive = 0xFFFFFFFFu; res.u = 0.f; res.v = 0.f; }
    else { res.t = h.t; res.instance = h.inst; res.primitive = h.prim; res.u = h.u; res.v = h.v; }
    out[i] = res;
}
} // namespace

int traceHostRays(vkx_ctx* ctx, const float* origins, const float* directions, size_t n, float tmin, float tmax, uint32_t cullMask, int anyHit, vkx_hit* out, bool alphaTest) {
    if (n == 0) return VKX_OK;
    if (n > 0x7FFFFFFFu) return vkx_fail(ctx, VKX_E_INVALID, "too many rays");
    cudaStream_t st = ctx->stream;
    float *dO = nullptr, *dD = nullptr; vkx_hit* dH = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&dO, n * 12)) != cudaSuccess || (e = cudaMalloc(&dD, n * 12)) != cudaSuccess || (e = cudaMalloc(&dH, n * sizeof(vkx_hit))) != cudaSuccess) {
        cudaFree(dO); cudaFree(dD); cudaFree(dH);
        return vkx_fail(ctx, VKX_E_NOMEM, "vkx_trace: %s", cudaGetErrorString(e));
    }
    int rc = VKX_OK;
    do {
        if ((e = cudaMemcpyAsync(dO, origins, n * 12, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dD, directions, n * 12, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        const DeviceScene sc = deviceScene(ctx);
        const bool alpha = alphaTest && sc.numTextures != 0u;
        if (anyHit) { if (alpha) k_trace_host<true, true><<<divUp(n, 128), 128, 0, st>>>(sc, dO, dD, uint32_t(n), tmin, tmax, cullMask, dH); else k_trace_host<true, false><<<divUp(n, 128), 128, 0, st>>>(sc, dO, dD, uint32_t(n), tmin, tmax, cullMask, dH); }
        else { if (alpha) k_trace_host<false, true><<<divUp(n, 128), 128, 0, st>>>(sc, dO, dD, uint32_t(n), tmin, tmax, cullMask, dH); else k_trace_host<false, false><<<divUp(n, 128), 128, 0, st>>>(sc, dO, dD, uint32_t(n), tmin, tmax, cullMask, dH); }
        ctx->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(out, dH, n * sizeof(vkx_hit), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        e = cudaStreamSynchronize(st);
    } while (0);
    if (e != cudaSuccess) rc = vkx_fail(ctx, VKX_E_CUDA, "vkx_trace: %s", cudaGetErrorString(e));
    cudaFree(dO); cudaFree(dD); cudaFree(dH);
    return rc;
}
