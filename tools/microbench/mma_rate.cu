// Micro-benchmark (diagnostics for blend_tc.cu): cycles per tcgen05.mma for the shapes / operand layouts the probe blend uses, one CTA
// per SM, one issuing thread, operands resident in shared memory (contents irrelevant). nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count)); }
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int KIND> // 0 tf32, 1 f16 (bf16 inputs)
__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory"); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return uint64_t((addr >> 4) & 0x3FFFu) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(layout) << 61);
}
// idesc: c format f32 (1 << 4); a/b format: tf32 = 2, bf16 = 1 (bits 7.., 10..); N >> 3 at 17; M >> 4 at 24
__host__ __device__ constexpr uint32_t idescOf(int kind, uint32_t n) { return (1u << 4) | ((kind == 0 ? 2u : 1u) << 7) | ((kind == 0 ? 2u : 1u) << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

template <int KIND>
__global__ void __launch_bounds__(288, 1) k_rate(uint32_t n, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t kstep, uint32_t iters, uint32_t accs, uint32_t batch, uint32_t background, unsigned long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t sTmem;
    __shared__ volatile uint32_t sStop;
    if (threadIdx.x == 0) sStop = 0u;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&sTmem)), "r"(512) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcFenceBefore(); __syncthreads(); tcFenceAfter();
    const uint32_t tmem = sTmem;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smemAddr(smem), b0 = a0 + 64 * 1024;
        const uint32_t id = idescOf(KIND, n);
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; it += batch) {
            for (uint32_t j = 0; j < batch; ++j) {
                const uint32_t ks = j & 1u, acc = (j % accs) * 128u;
                umma<KIND>(tmem + acc, desc(a0 + ks * kstep + (j & 3u) * 8192u, lbo, sbo, layout), desc(b0 + ks * kstep + (j & 3u) * 12288u, lbo, sbo, layout), id, j >= accs ? 1u : 0u);
            }
            ummaCommit(&bar);
            mbarWait(&bar, parity); parity ^= 1u;
        }
        const long long t1 = clock64();
        out[blockIdx.x] = (unsigned long long)(t1 - t0);
        sStop = 1u;
    } else if (background && threadIdx.x >= 32) { // background shared-memory stores while the MMAs run (the producer warps of the blend)
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem + 140 * 1024) + (threadIdx.x - 32);
        uint32_t v = threadIdx.x;
        while (!sStop) {
#pragma unroll
            for (int k = 0; k < 40; ++k) dst[(k & 7) * 512] = v + k;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    tcFenceBefore(); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int KIND>
static void run(const char* what, uint32_t n, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t kstep, uint32_t accs, uint32_t batch, int blocks, uint32_t background = 0) {
    unsigned long long* out; cudaMallocManaged(&out, 8 * blocks);
    const uint32_t iters = 4608;
    cudaFuncSetAttribute(k_rate<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int rep = 0; rep < 2; ++rep) { k_rate<KIND><<<blocks, background ? 288 : 128, 160 * 1024>>>(n, lbo, sbo, layout, kstep, iters, accs, batch, background, out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; } }
    double mean = 0; for (int b = 0; b < blocks; ++b) mean += double(out[b]); mean /= blocks;
    printf("%-64s N=%3u accs=%u batch=%3u blocks=%3d bg=%u : %7.1f cycles per MMA\n", what, n, accs, batch, blocks, background, mean / iters);
    cudaFree(out);
}
int main() {
    for (uint32_t bg : {0u, 1u}) { // the blend's mix with and without 8 warps of background shared-memory stores
        run<0>("tf32 M128 K8, SWIZZLE_64B SBO 512 (blend mix)", 128, 16, 512, 4, 32, 3, 18, 148, bg);
        run<0>("tf32 M128 K8, SWIZZLE_64B SBO 512 (blend mix)", 192, 16, 512, 4, 32, 2, 18, 148, bg);
    }
    for (int blocks : {148}) {
        for (uint32_t batch : {18u}) {
            run<0>("tf32 M128 K8, interleaved LBO 128 SBO 528", 128, 128, 528, 0, 256, 3, batch, blocks);
            run<0>("tf32 M128 K8, interleaved LBO 128 SBO 512", 128, 128, 512, 0, 256, 3, batch, blocks);
            run<0>("tf32 M128 K8, SWIZZLE_64B SBO 512", 128, 16, 512, 4, 32, 3, batch, blocks);
            run<0>("tf32 M128 K8, SWIZZLE_128B SBO 1024", 128, 16, 1024, 2, 32, 3, batch, blocks);
            run<0>("tf32 M128 K8, SWIZZLE_128B SBO 1024", 192, 16, 1024, 2, 32, 2, batch, blocks);
            run<0>("tf32 M128 K8, SWIZZLE_128B SBO 1024", 256, 16, 1024, 2, 32, 2, batch, blocks);
            run<0>("tf32 M128 K8, SWIZZLE_128B, one accumulator", 128, 16, 1024, 2, 32, 1, batch, blocks);
            run<1>("bf16 M128 K16, SWIZZLE_128B SBO 1024", 128, 16, 1024, 2, 32, 3, batch, blocks);
            run<1>("bf16 M128 K16, SWIZZLE_128B SBO 1024", 256, 16, 1024, 2, 32, 2, batch, blocks);
            run<1>("bf16 M128 K16, interleaved LBO 128 SBO 512", 128, 128, 512, 0, 256, 3, batch, blocks);
        }
    }
    return 0;
}
