// Probe blend on the 5th-generation tensor cores (tcgen05). Replaces the accumulation loop of k_blend (ddgi.cu), i.e. the ray
// sums of probesUpdate.glsl:68-86 (reference src/shaders/probesUpdate.glsl): every probe uses the same rotated ray directions, so
// the cosine / pow(cosine, sharpness) weights form ONE table W[texel][ray] per frame and the blend of a batch of probes is
//     depth moments     S[texel 0..195][(probe, d | d^2)] = Wd[196 x 256] . Dd[256 x 2P]
//     irradiance sums   S[texel 0..35 ][(probe, r|g|b)]   = Wi[ 36 x 256] . Dc[256 x 3P]
// (4.2 GFLOP per full-volume update at cfg2). One CTA blends BTC_P = 64 probes:
//   * A operand = weights, M = texel rows in two 128-row tiles (tile 0: depth texels 0..127; tile 1: depth texels 128..195 in rows
//     0..67 and the 36 irradiance texels in rows 68..103). The per-frame table is laid out by k_blend_weight_image as the exact
//     shared-memory image of the K-major, unswizzled UMMA operand (hi and lo TF32 parts), 16 rays per chunk, so one
//     cp.async.bulk (TMA bulk copy, UBLKCP) per chunk brings it in, completing on an mbarrier.
//   * B operand = ray data, N = (probe, plane) rows: the ray records (rgb, depth) are read once from global memory, depth is clamped
//     and squared (probesUpdate.glsl:74,78-79), every value is split into TF32 hi + lo and stored in the same UMMA layout.
//   * 3xTF32: D += Ahi.Bhi + Ahi.Blo + Alo.Bhi with fp32 accumulation in tensor memory (448 of 512 columns): depth tile 0 -> columns
//     [0,128), depth tile 1 -> [128,256), tile 1 x colour planes -> [256,448) (only its irradiance rows are read back).
//   * two shared-memory stages: the MMAs of chunk c (issued by one thread, completion signalled with tcgen05.commit on an
//     mbarrier) run while all threads produce chunk c+1.
//   * epilogue: tcgen05.ld (one accumulator row per thread) -> shared memory in the layout the tail of the blend expects, eight
//     probes at a time; then normalisation, hysteresis mix against the work atlas, state machine, border texels and 128-bit tile
//     stores exactly as in k_blend.
// Accuracy: the split keeps 21 mantissa bits per operand; measured against the oracle the pre-pack fp32 texels agree like the
// CUDA-core kernel's (tests/test_ddgi_parity.py), so the packed-code flip rate is unchanged.
#include "common.cuh"
#include "shade.cuh"
#include "ddgi_common.cuh"
#include "blend_tc.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count)); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { // TMA bulk copy (1-D), completes on the mbarrier
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); } // generic-proxy stores -> visible to the tensor core
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmemAlloc(uint32_t* slot, uint32_t cols) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(slot)), "r"(cols) : "memory"); }
__device__ __forceinline__ void tmemRelinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmemFree(uint32_t addr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory"); }
__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) { // D[tmem] (+)= A[smem] . B[smem]^T, TF32
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory"); }
__device__ __forceinline__ void tmemLoad8(uint32_t addr, uint32_t (&v)[8]) { // this thread's accumulator row (lane of its warp's quarter), 8 consecutive columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmemLoadWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, unswizzled UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address, leading byte offset (between the
// two 16-byte K cores of one MMA), stride byte offset (between 8-row groups), all in 16-byte units; version 1 (Blackwell).
__device__ __forceinline__ uint64_t ummaDesc(uint32_t smemByteAddr) {
    return uint64_t((smemByteAddr >> 4) & 0x3FFFu) | (uint64_t(BTC_LBO >> 4) << 16) | (uint64_t(BTC_SBO >> 4) << 32) | (uint64_t(1) << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, both operands K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t ummaIdesc(uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

// x = hi + lo exactly; hi = x rounded to the nearest TF32 (10 mantissa bits), so lo is signed and the tensor core's truncation of
// lo to TF32 errs in both directions. (Truncating x instead made every lo positive and every product err low: a bias of 2e-6
// relative on sums of 256 positive terms, enough to flip 0.5 % of the RG16F depth codes; measured in tests/test_ddgi_parity.py.)
__device__ __forceinline__ void splitTf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = __fsub_rn(x, hi);
}
// byte offset of element (row, k) inside one operand tile of a chunk: [row group][k core][row in group][k in core]
__device__ __forceinline__ uint32_t tileOffset(uint32_t row, uint32_t k) { return (row >> 3) * BTC_SBO + (k >> 2) * BTC_LBO + (row & 7u) * 16u + (k & 3u) * 4u; }

} // namespace

// Re-lays the per-frame weight table W[ray][col] (k_blend_weights) as the shared-memory image of the A operand, chunk by chunk:
// [chunk][hi | lo][tile 0 | 1][BTC_A_TILE_BYTES]. One thread per (chunk, tile, row, k).
__global__ void k_blend_weight_image(uint32_t N, const float* __restrict__ W, float* __restrict__ image) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = e & 15u, row = (e >> 4) & 127u, tile = (e >> 11) & 1u, chunk = e >> 12;
    if (chunk >= BTC_MAX_CHUNKS) return;
    const uint32_t ray = chunk * BTC_KC + k;
    int col = -1;
    if (tile == 0) col = int(row);
    else if (row < 68u) col = 128 + int(row);
    else if (row < 104u) col = BLEND_IRR_COL0 + int(row - 68u);
    float w = 0.0f;
    if (col >= 0 && ray < N) w = W[size_t(ray) * BLEND_COLS + uint32_t(col)];
    float hi, lo; splitTf32(w, hi, lo);
    char* base = reinterpret_cast<char*>(image) + size_t(chunk) * BTC_A_CHUNK_BYTES + size_t(tile) * BTC_A_TILE_BYTES + tileOffset(row, k);
    *reinterpret_cast<float*>(base) = hi;
    *reinterpret_cast<float*>(base + 2 * BTC_A_TILE_BYTES) = lo;
}

__global__ void __launch_bounds__(BTC_THREADS, 1) k_blend_tc(BlendParams bp, DeviceProbes pr, const uint32_t* __restrict__ probeIndices, const float4* __restrict__ rays,
                                                             const float* __restrict__ W, const float* __restrict__ image, float* __restrict__ irrUnpacked,
                                                             float* __restrict__ depUnpacked, uint32_t slotBase) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // stage s: [A chunk image (hi t0, hi t1, lo t0, lo t1)][B depth hi][B depth lo][B colour hi][B colour lo]
    unsigned char* stage[2] = {smem, smem + BTC_STAGE_BYTES};
    __shared__ __align__(8) uint64_t barA[2], barM[2], barDone;
    __shared__ uint32_t sTmem;
    __shared__ uint32_t sMaxChange[BTC_P], sOutOfRange[BTC_P], sLinear[BTC_P];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, N = bp.raysPerProbe;
    const uint32_t slot0 = blockIdx.x * BTC_P;
    const uint32_t np = min(uint32_t(BTC_P), bp.count - slot0);
    const uint32_t chunks = (N + BTC_KC - 1u) / BTC_KC;
    const float cellLen = bp.gridCellLen;
    if (tid < BTC_P) { sMaxChange[tid] = 0u; sOutOfRange[tid] = 0u; sLinear[tid] = tid < np ? __ldg(probeIndices + slot0 + tid) : 0u; }
    if (tid == 0) { mbarInit(&barA[0], 1); mbarInit(&barA[1], 1); mbarInit(&barM[0], 1); mbarInit(&barM[1], 1); mbarInit(&barDone, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { tmemAlloc(&sTmem, 512); tmemRelinquish(); }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmem = sTmem;

    // ---- main loop over chunks of BTC_KC rays. A thread owns the same (probe, ray-in-core) of every K core: element i of a chunk is
    // probe tid / 4, ray 4 i + tid % 4 (a warp covers 8 probes x 4 rays per element: conflict-free stores, see tileOffset). The ray
    // records of chunk c + 1 are requested before chunk c is converted, so their latency overlaps the conversion and the barrier.
    constexpr uint32_t EPT = BTC_P * BTC_KC / BTC_THREADS; // elements per thread per chunk = K cores per chunk
    static_assert(EPT == BTC_KC / 4u && BTC_THREADS == 4u * BTC_P, "one K core per element");
    const uint32_t myP = tid >> 2, myKq = tid & 3u;
    const float4* myRays = rays + size_t(slot0 + myP) * N;
    auto fetch = [&](uint32_t c, float4 (&rd)[EPT]) {
#pragma unroll
        for (uint32_t i = 0; i < EPT; ++i) {
            const uint32_t ray = c * BTC_KC + 4u * i + myKq;
            rd[i] = (myP < np && ray < N) ? myRays[ray] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float4 cur[EPT], nxt[EPT];
    fetch(0, cur);
    uint32_t outOfRange = 0;
    for (uint32_t c = 0; c < chunks; ++c) {
        const uint32_t s = c & 1u;
        unsigned char* st = stage[s];
        if (c + 1u < chunks) fetch(c + 1u, nxt);
        if (c >= 2u) mbarWait(&barM[s], ((c >> 1) - 1u) & 1u); // the MMAs that read this stage two chunks ago are done
        if (tid == 0) { // weights of the chunk: one bulk copy
            mbarExpectTx(&barA[s], BTC_A_CHUNK_BYTES);
            bulkCopyG2S(st, reinterpret_cast<const char*>(image) + size_t(c) * BTC_A_CHUNK_BYTES, BTC_A_CHUNK_BYTES, &barA[s]);
        }
        unsigned char* bD = st + BTC_A_CHUNK_BYTES;                 // depth planes hi, then lo
        unsigned char* bC = bD + 2 * BTC_BD_TILE_BYTES;             // colour planes hi, then lo
#pragma unroll
        for (uint32_t i = 0; i < EPT; ++i) {
            const uint32_t k = 4u * i + myKq, ray = c * BTC_KC + k;
            float4 rd = cur[i];
            if (myP < np && ray < N) {
                if (rd.w < 0.0f || rd.w > cellLen) ++outOfRange;                   // probesUpdate.glsl:74
                float depth = minS(cellLen, rd.w);                                  // :78-79
                if (depth < 0.0f) depth = cellLen;
                rd.w = depth;
            }
            float hi, lo;
            const uint32_t od = tileOffset(2u * myP, k), oc = tileOffset(3u * myP, k);
            splitTf32(rd.w, hi, lo);           *reinterpret_cast<float*>(bD + od) = hi;        *reinterpret_cast<float*>(bD + BTC_BD_TILE_BYTES + od) = lo;
            splitTf32(rd.w * rd.w, hi, lo);    *reinterpret_cast<float*>(bD + od + 16u) = hi;  *reinterpret_cast<float*>(bD + BTC_BD_TILE_BYTES + od + 16u) = lo;
            // colour rows 3p, 3p+1, 3p+2 may cross an 8-row group: address each
            splitTf32(rd.x, hi, lo);           *reinterpret_cast<float*>(bC + oc) = hi;        *reinterpret_cast<float*>(bC + BTC_BC_TILE_BYTES + oc) = lo;
            const uint32_t oc1 = tileOffset(3u * myP + 1u, k), oc2 = tileOffset(3u * myP + 2u, k);
            splitTf32(rd.y, hi, lo);           *reinterpret_cast<float*>(bC + oc1) = hi;       *reinterpret_cast<float*>(bC + BTC_BC_TILE_BYTES + oc1) = lo;
            splitTf32(rd.z, hi, lo);           *reinterpret_cast<float*>(bC + oc2) = hi;       *reinterpret_cast<float*>(bC + BTC_BC_TILE_BYTES + oc2) = lo;
        }
#pragma unroll
        for (uint32_t i = 0; i < EPT; ++i) cur[i] = nxt[i];
        fenceProxyAsync();
        __syncthreads();
        if (tid == 0) {
            mbarWait(&barA[s], (c >> 1) & 1u);
            tcFenceAfter();
            const uint32_t aHi0 = smemAddr(st), aHi1 = aHi0 + BTC_A_TILE_BYTES, aLo0 = aHi0 + 2 * BTC_A_TILE_BYTES, aLo1 = aHi0 + 3 * BTC_A_TILE_BYTES;
            const uint32_t dHi = smemAddr(bD), dLo = dHi + BTC_BD_TILE_BYTES, cHi = smemAddr(bC), cLo = cHi + BTC_BC_TILE_BYTES;
            constexpr uint32_t idD = ummaIdesc(2 * BTC_P), idC = ummaIdesc(3 * BTC_P);
#pragma unroll
            for (uint32_t ks = 0; ks < BTC_KC / 8u; ++ks) {
                const uint32_t ko = ks * 2u * BTC_LBO; // two K cores per MMA
                const uint32_t acc = (c | ks) ? 1u : 0u;
                // depth tile 0 -> columns [0, 2P), depth tile 1 -> [2P, 4P), tile 1 x colour -> [4P, 7P)
                umma(tmem + 0u, ummaDesc(aHi0 + ko), ummaDesc(dHi + ko), idD, acc);
                umma(tmem + 0u, ummaDesc(aHi0 + ko), ummaDesc(dLo + ko), idD, 1u);
                umma(tmem + 0u, ummaDesc(aLo0 + ko), ummaDesc(dHi + ko), idD, 1u);
                umma(tmem + 2u * BTC_P, ummaDesc(aHi1 + ko), ummaDesc(dHi + ko), idD, acc);
                umma(tmem + 2u * BTC_P, ummaDesc(aHi1 + ko), ummaDesc(dLo + ko), idD, 1u);
                umma(tmem + 2u * BTC_P, ummaDesc(aLo1 + ko), ummaDesc(dHi + ko), idD, 1u);
                umma(tmem + 4u * BTC_P, ummaDesc(aHi1 + ko), ummaDesc(cHi + ko), idC, acc);
                umma(tmem + 4u * BTC_P, ummaDesc(aHi1 + ko), ummaDesc(cLo + ko), idC, 1u);
                umma(tmem + 4u * BTC_P, ummaDesc(aLo1 + ko), ummaDesc(cHi + ko), idC, 1u);
            }
            ummaCommit(&barM[s]);                      // stage free when these MMAs have read it
            if (c + 1u == chunks) ummaCommit(&barDone); // accumulators complete
        }
    }
    if (outOfRange) atomicAdd(&sOutOfRange[myP], outOfRange);
    mbarWait(&barDone, 0u);
    tcFenceAfter();
    __syncthreads(); // the stages are free: their memory now holds the epilogue's staging

    // ---- epilogue, 8 probes at a time: accumulators -> shared memory -> the tail of k_blend
    float* sResD = reinterpret_cast<float*>(smem);                                      // [8][2][196]
    float* sResI = sResD + 8 * 2 * 196;                                                 // [8][3][36]
    uint32_t (*sDep)[256] = reinterpret_cast<uint32_t (*)[256]>(sResI + 8 * 3 * 36);    // [8][256]
    uint32_t (*sIrr)[64] = reinterpret_cast<uint32_t (*)[64]>(sDep + 8);                // [8][64]
    float* sRw = reinterpret_cast<float*>(sIrr + 8);                                   // [232] per-texel weight sums (depth 0..195, irradiance 196..231)
    uint32_t* sOrigin = reinterpret_cast<uint32_t*>(sRw + 232);                         // [BTC_P][2]: word offset of the probe's depth / irradiance tile in its atlas
    for (uint32_t t = tid; t < 232u; t += BTC_THREADS) sRw[t] = __ldg(W + size_t(BLEND_WSUM_ROW) * BLEND_COLS + (t < 196u ? t : BLEND_IRR_COL0 + t - 196u));
    if (tid < np) {
        int ix, iy, iz; probeGridIndex(sLinear[tid], bp.grid, ix, iy, iz);
        const int tile = iy * bp.grid.resolution[0] + ix;
        sOrigin[2 * tid] = uint32_t(size_t(16 * iz) * pr.depW + size_t(16 * tile));
        sOrigin[2 * tid + 1] = uint32_t(size_t(8 * iz) * pr.irrW + size_t(8 * tile));
    }
    const float hysteresis = bp.grid.hysteresis;
    const uint32_t laneBase = ((warp & 3u) * 32u) << 16; // this warp's quarter of the 128 accumulator rows
    const uint32_t row = (warp & 3u) * 32u + (tid & 31u);
    for (uint32_t j = 0; j < BTC_P / 8u; ++j) {
        if (j * 8u >= np) break;
        if (warp < 4u) { // depth tile 0: texel = row, columns 16j .. 16j+15 = (probe 8j + q, plane)
            uint32_t v[8];
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                tmemLoad8(tmem + laneBase + 16u * j + 8u * h, v); tmemLoadWait();
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) sResD[((4u * h + (q >> 1)) * 2u + (q & 1u)) * 196u + row] = __uint_as_float(v[q]);
            }
        } else { // warps 4..7: tile 1: rows 0..67 depth texels 128..195 (columns 2P + 16j ..), rows 68..103 irradiance texels (columns 4P + 24j ..)
            uint32_t v[8];
            if (row < 96u) { // warps 4, 5, 6 hold depth rows (row < 68); warp-uniform condition
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h) {
                    tmemLoad8(tmem + laneBase + 2u * BTC_P + 16u * j + 8u * h, v); tmemLoadWait();
                    if (row < 68u) {
#pragma unroll
                        for (uint32_t q = 0; q < 8; ++q) sResD[((4u * h + (q >> 1)) * 2u + (q & 1u)) * 196u + 128u + row] = __uint_as_float(v[q]);
                    }
                }
            }
            if (row >= 64u) { // warps 6, 7 hold irradiance rows (68 <= row < 104)
#pragma unroll
                for (uint32_t h = 0; h < 3; ++h) {
                    tmemLoad8(tmem + laneBase + 4u * BTC_P + 24u * j + 8u * h, v); tmemLoadWait();
                    if (row >= 68u && row < 104u) {
#pragma unroll
                        for (uint32_t q = 0; q < 8; ++q) { const uint32_t col = 8u * h + q; sResI[((col / 3u) * 3u + (col % 3u)) * 36u + (row - 68u)] = __uint_as_float(v[q]); }
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t p0 = j * 8u, npj = min(8u, np - p0);
        // normalisation (probesUpdate.glsl:85-86) + hysteresis mix against the work atlas (:92-103) + pack; same arithmetic as k_blend.
        // Every thread first requests all its previous texels (one round trip per sub-batch instead of one per texel), then mixes.
        constexpr uint32_t IT = (8u * 232u + BTC_THREADS - 1u) / BTC_THREADS;
        uint32_t prevWord[IT];
#pragma unroll
        for (uint32_t i = 0; i < IT; ++i) {
            const uint32_t e = tid + i * BTC_THREADS;
            prevWord[i] = 0u;
            if (e < npj * 232u) {
                const uint32_t p = e / 232u, t = e - p * 232u;
                if (t < 196u) prevWord[i] = pr.depWork[sOrigin[2 * (p0 + p)] + (1u + t / 14u) * pr.depW + 1u + t % 14u];
                else { const uint32_t ti = t - 196u; prevWord[i] = pr.irrWork[sOrigin[2 * (p0 + p) + 1] + (1u + ti / 6u) * pr.irrW + 1u + ti % 6u]; }
            }
        }
#pragma unroll
        for (uint32_t i = 0; i < IT; ++i) {
            const uint32_t e = tid + i * BTC_THREADS;
            if (e >= npj * 232u) continue;
            const uint32_t p = e / 232u, t = e - p * 232u;
            const float rw = sRw[t];
            if (t < 196u) {
                float r0 = sResD[(p * 2u + 0u) * 196u + t], r1 = sResD[(p * 2u + 1u) * 196u + t];
                if (rw > 1e-3f) { r0 = r0 / rw; r1 = r1 / rw; }
                const int lx = int(t % 14u), ly = int(t / 14u);
                const float2 prev = unpackRG16F(prevWord[i]);
                const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis);
                sDep[p][(ly + 1) * 16 + (lx + 1)] = packRG16F(o0, o1);
                if (depUnpacked) { float* up = depUnpacked + (size_t(slotBase + slot0 + p0 + p) * 196 + t) * 2; up[0] = o0; up[1] = o1; }
            } else {
                const uint32_t ti = t - 196u;
                float r0 = sResI[(p * 3u + 0u) * 36u + ti], r1 = sResI[(p * 3u + 1u) * 36u + ti], r2 = sResI[(p * 3u + 2u) * 36u + ti];
                if (rw > 1e-3f) { r0 = r0 / rw; r1 = r1 / rw; r2 = r2 / rw; }
                const int lx = int(ti % 6u), ly = int(ti / 6u);
                const float3 prev = unpackR11G11B10(prevWord[i]);
                const float maxChange = maxS(maxS(fabsf(r0 - prev.x), fabsf(r1 - prev.y)), fabsf(r2 - prev.z));
                const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis), o2 = mixf(r2, prev.z, hysteresis);
                sIrr[p][(ly + 1) * 8 + (lx + 1)] = packR11G11B10(o0, o1, o2);
                if (irrUnpacked) { float* up = irrUnpacked + (size_t(slotBase + slot0 + p0 + p) * 36 + ti) * 3; up[0] = o0; up[1] = o1; up[2] = o2; }
                atomicMax(&sMaxChange[p0 + p], __float_as_uint(maxChange)); // probesUpdate.glsl:106-107
            }
        }
        __syncthreads();
        if (tid < npj) { // state machine, probesUpdate.glsl:110-119 (decree A.5.3: full max over the 36 texels)
            const uint32_t linearIndex = sLinear[p0 + tid];
            uint32_t stt = pr.stateWork[linearIndex];
            if (sOutOfRange[p0 + tid] >= N) stt = 8;
            else {
                const float maxChange = __uint_as_float(sMaxChange[p0 + tid]);
                if (maxChange < 0.02f / float(stt)) stt = min(stt + 1u, 8u);
                else if (maxChange > 0.04f / float(stt)) stt = max(stt - 1u, 1u);
                else if (maxChange > 0.25f) stt = 1;
            }
            pr.stateWork[linearIndex] = stt;
        }
        for (uint32_t e = tid; e < npj * 88u; e += BTC_THREADS) { // borders (probesCopyBorders.comp) from the shared tiles
            const uint32_t p = e / 88u, b = e - p * 88u;
            int x, y, sx, sy;
            if (b < 60u) {
                if (b < 16u) { x = int(b); y = 0; } else if (b < 32u) { x = int(b) - 16; y = 15; } else if (b < 46u) { x = 0; y = int(b) - 32 + 1; } else { x = 15; y = int(b) - 46 + 1; }
                blendBorderSource(16, x, y, sx, sy);
                sDep[p][y * 16 + x] = sDep[p][sy * 16 + sx];
            } else {
                const int cc = int(b) - 60;
                if (cc < 8) { x = cc; y = 0; } else if (cc < 16) { x = cc - 8; y = 7; } else if (cc < 22) { x = 0; y = cc - 16 + 1; } else { x = 7; y = cc - 22 + 1; }
                blendBorderSource(8, x, y, sx, sy);
                sIrr[p][y * 8 + x] = sIrr[p][sy * 8 + sx];
            }
        }
        __syncthreads();
        for (uint32_t e = tid; e < npj * 80u; e += BTC_THREADS) { // 128-bit tile stores
            const uint32_t p = e / 80u, k = e - p * 80u;
            if (k < 64u) {
                const uint32_t rrow = k >> 2, q = k & 3u;
                *reinterpret_cast<uint4*>(pr.depWork + sOrigin[2 * (p0 + p)] + size_t(rrow) * pr.depW + q * 4u) = *reinterpret_cast<const uint4*>(&sDep[p][rrow * 16u + q * 4u]);
            } else {
                const uint32_t kk = k - 64u, rrow = kk >> 1, q = kk & 1u;
                *reinterpret_cast<uint4*>(pr.irrWork + sOrigin[2 * (p0 + p) + 1] + size_t(rrow) * pr.irrW + q * 4u) = *reinterpret_cast<const uint4*>(&sIrr[p][rrow * 8u + q * 4u]);
            }
        }
        __syncthreads();
    }
    tcFenceBefore();
    __syncthreads();
    if (warp == 0) tmemFree(tmem, 512);
}

int blendTcWeights(vkx_ctx* ctx, cudaStream_t st) {
    k_blend_weight_image<<<(BTC_MAX_CHUNKS * 2 * 128 * 16) / 256, 256, 0, st>>>(ctx->grid.raysPerProbe, ctx->dBlendW, ctx->dBlendImage); LAUNCH_CHECK(ctx);
    return VKX_OK;
}

int blendTcLaunch(vkx_ctx* ctx, const BlendParams& bp, const DeviceProbes& pr, const uint32_t* idx, uint32_t n, uint32_t slotBase, cudaStream_t st) {
    if (n == 0) return VKX_OK;
    if (!ctx->blendTcAttrSet) { CUDA_TRY(ctx, cudaFuncSetAttribute(k_blend_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, BTC_SMEM_BYTES)); ctx->blendTcAttrSet = true; }
    k_blend_tc<<<divUp(n, BTC_P), BTC_THREADS, BTC_SMEM_BYTES, st>>>(bp, pr, idx, ctx->dRays, ctx->dBlendW, ctx->dBlendImage, ctx->debugBuffers ? ctx->dIrrUnpacked : nullptr,
                                                                      ctx->debugBuffers ? ctx->dDepUnpacked : nullptr, slotBase); LAUNCH_CHECK(ctx);
    return VKX_OK;
}
