// Probe blend on the 5th-generation tensor cores (tcgen05). Replaces the accumulation loop of k_blend (ddgi.cu), i.e. the ray
// sums of probesUpdate.glsl:68-86 (reference src/shaders/probesUpdate.glsl): every probe uses the same rotated ray directions, so
// the cosine / pow(cosine, sharpness) weights form ONE table W[texel][ray] per frame and the blend of a batch of probes is
//     depth moments     S[texel 0..195][(probe, d | d^2)] = Wd[196 x 256] . Dd[256 x 2P]
//     irradiance sums   S[texel 0..35 ][(probe, r|g|b)]   = Wi[ 36 x 256] . Dc[256 x 3P]
// (4.2 GFLOP per full-volume update at cfg2). Persistent, warp-specialised kernel: one CTA per SM walks over tiles of BTC_P = 64 probes.
//   * A operand = weights, M = texel rows in two 128-row tiles (tile 0: depth texels 0..127; tile 1: depth texels 128..195 in rows
//     0..67 and the 36 irradiance texels in rows 68..103). The per-frame table is laid out by k_blend_weight_image as the exact
//     shared-memory image of the K-major, unswizzled UMMA operand (hi and lo TF32 parts), 16 rays per chunk: the TMA warp brings a
//     chunk in with one cp.async.bulk (UBLKCP) that completes on the stage's `full` mbarrier.
//   * B operand = ray data, N = (plane, probe) rows: eight producer warps read the ray records (rgb, depth) once from global memory
//     (a thread takes four consecutive rays of one probe; the next two chunks' loads are in flight while one is converted), clamp and
//     square the depth (probesUpdate.glsl:74,78-79), split every value into TF32 hi + lo and store each plane's K core with one
//     conflict-free 128-bit store in the same UMMA layout, then arrive on `full`.
//   * 3xTF32: D += Ahi.Bhi + Ahi.Blo + Alo.Bhi with fp32 accumulation in tensor memory (448 of 512 columns): depth tile 0 -> columns
//     [0,128) (d of probe p in column p, d^2 in 64 + p), depth tile 1 -> [128,256), tile 1 x colour planes -> [256,448) (only its irradiance rows are read back). One lane of
//     the MMA warp issues them; tcgen05.commit releases the stage (`empty`) and, after a tile's last chunk, publishes the accumulators
//     (`accFull`). Producers and TMA run ahead into the next tile while the epilogue drains the accumulators.
//   * epilogue: eight warps. A warp reads the TMEM lane quarter warp % 4, so the two warps of a quarter split the tile's probes. While
//     the tensor core works on the tile they load the previous texels of their (texel, probe) pairs into a private shared-memory
//     column; when `accFull` arrives each thread reads its accumulator row with tcgen05.ld (8 probes at a time), normalises (exact
//     quotient through a shared reciprocal), mixes with the previous texel (hysteresis), packs and puts the result back into its
//     column - no global memory traffic, so tensor memory is handed back (`accEmpty`) after ~7 k cycles. Then all eight warps write
//     the tile out: 128-bit pieces of atlas rows, border texels taken from the interior texel probesCopyBorders.comp copies them from,
//     four lanes per 64-byte row = full sectors (scattered 4-byte stores from the accumulator rows measured 100 k cycles per tile),
//     and the per-probe state machine (probesUpdate.glsl:110-119) runs. This overlaps the next tile's main loop.
// Accuracy: the split keeps 21 mantissa bits per operand; pre-pack fp32 texels agree with the oracle like the CUDA-core kernel's
// (tests/test_ddgi_parity.py).
#include "common.cuh"
#include "shade.cuh"
#include "ddgi_common.cuh"
#include "blend_tc.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

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
__device__ __forceinline__ void mbarArrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory"); }
__device__ __forceinline__ void namedBarrier(uint32_t id, uint32_t threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void tmemLoadWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address, leading byte offset (between the two 16-byte
// K cores of one MMA; unswizzled layout only), stride byte offset (between 8-row groups), all in 16-byte units; version 1
// (Blackwell); layout type in bits 61..63 (0 = no swizzle, 4 = SWIZZLE_64B). One K = 8 step advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t ummaDesc(uint32_t smemByteAddr) {
    return uint64_t((smemByteAddr >> 4) & 0x3FFFu) | (uint64_t(BTC_LBO >> 4) << 16) | (uint64_t(BTC_SBO >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(BTC_LAYOUT == 1 ? 4 : 0) << 61);
}
#if BTC_LAYOUT == 1
#define BTC_KSTEP_BYTES 32u
#else
#define BTC_KSTEP_BYTES (2u * BTC_LBO)
#endif
// Instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, both operands K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t ummaIdesc(uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

// x = hi + lo exactly; hi = x rounded to the nearest TF32 (10 mantissa bits), so lo is signed and the tensor core's truncation of
// lo to TF32 errs in both directions. (Truncating x instead made every lo positive and every product err low: a bias of 2e-6
// relative on sums of 256 positive terms, enough to flip 0.5 % of the RG16F depth codes; measured in tests/test_ddgi_parity.py.)
__device__ __forceinline__ void splitTf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = __fsub_rn(x, hi);
}
// byte offset of element (row, k) inside one operand tile of a chunk (see BTC_LAYOUT in blend_tc.cuh)
#if BTC_LAYOUT == 1
__device__ __forceinline__ uint32_t tileOffset(uint32_t row, uint32_t k) { return (row >> 3) * BTC_SBO + (row & 7u) * 64u + (((k >> 2) ^ ((row >> 1) & 3u)) << 4) + (k & 3u) * 4u; }
#else
__device__ __forceinline__ uint32_t tileOffset(uint32_t row, uint32_t k) { return (row >> 3) * BTC_SBO + (k >> 2) * BTC_LBO + (row & 7u) * 16u + (k & 3u) * 4u; }
#endif

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


// VKX_BLEND_PROFILE=1 (diagnostics): per-CTA cycle counts of where each role waits, 16 slots per CTA (see blendTcLaunch)
#define BTC_TIMED(slot, stmt) do { if (prof) { const long long t0_ = clock64(); stmt; pacc[slot] += (unsigned long long)(clock64() - t0_); } else { stmt; } } while (0)

// Word offsets (relative to the probe's tile origin) of an interior texel and of the border texels that copy from it: the inverse of
// blendBorderSource (probesCopyBorders.comp). (ix, iy) in [1, T-2]^2; up to three borders (interior corners feed a row, a column and
// the opposite corner texel).
__device__ __forceinline__ void texelTargets(int T, int ix, int iy, uint32_t pitch, uint32_t& interior, uint32_t (&border)[3], int& nb) {
    const int L = T - 1;
    interior = uint32_t(iy) * pitch + uint32_t(ix);
    const bool ex = (ix == 1 || ix == L - 1), ey = (iy == 1 || iy == L - 1);
    // left / right column: (0, y) <- (1, L - y), (L, y) <- (L - 1, L - y); top / bottom row: (x, 0) <- (L - x, 1), (x, L) <- (L - x, L - 1);
    // corner (x, y) <- (x == 0 ? L - 1 : 1, y == 0 ? L - 1 : 1). ix cannot be both 1 and L - 1 (T >= 8), so at most one of each kind.
    const uint32_t colB = uint32_t(L - iy) * pitch + (ix == 1 ? 0u : uint32_t(L));
    const uint32_t rowB = (iy == 1 ? 0u : uint32_t(L)) * pitch + uint32_t(L - ix);
    const uint32_t cornerB = uint32_t(iy == L - 1 ? 0 : L) * pitch + uint32_t(ix == L - 1 ? 0 : L);
    nb = int(ex) + int(ey) + int(ex && ey);
    border[0] = ex ? colB : rowB;          // first target: the column copy if there is one, else the row copy
    border[1] = (ex && ey) ? rowB : 0u;
    border[2] = cornerB;
}

// n / d for the per-texel weight sum d (the same for every probe of the frame) with r = RN(1 / d) computed once per thread:
// q0 = RN(n r), e = n - d q0 exactly (FMA), q = RN(q0 + e r) is the correctly rounded quotient (Markstein's correction step; d is a
// positive normal number > 1e-3 here), i.e. what the `/` of probesUpdate.glsl:85 gives, at three instructions instead of a division.
__device__ __forceinline__ float divShared(float n, float d, float r) { const float q0 = __fmul_rn(n, r); return __fmaf_rn(__fmaf_rn(-d, q0, n), r, q0); }
__device__ __forceinline__ uint32_t packRG16Fx2(float r, float g) { const __half2 h = __floats2half2_rn(r, g); return *reinterpret_cast<const uint32_t*>(&h); } // one F2FP instead of two F2F + shift/or

struct TileMeta { uint32_t linear[BTC_P]; uint32_t originD[BTC_P]; uint32_t originI[BTC_P]; uint32_t outOfRange[BTC_P]; uint32_t maxChange[BTC_P]; };

__global__ void __launch_bounds__(BTC_THREADS, 1) k_blend_tc(BlendParams bp, DeviceProbes pr, const uint32_t* __restrict__ probeIndices, const float4* __restrict__ rays,
                                                             const float* __restrict__ W, const float* __restrict__ image, float* __restrict__ irrUnpacked,
                                                             float* __restrict__ depUnpacked, uint32_t slotBase, unsigned long long* __restrict__ prof) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // weight ring: ASTAGES x [A chunk image (hi t0, hi t1, lo t0, lo t1)]; ray-data ring: BSTAGES x [B depth hi][B depth lo][B colour hi][B colour lo]
    unsigned char* const smemB = smem + BTC_ASTAGES * BTC_A_CHUNK_BYTES;
    uint32_t* const smemPrev = reinterpret_cast<uint32_t*>(smemB + BTC_BSTAGES * BTC_B_STAGE_BYTES); // [local probe][A | B][epilogue thread]
    __shared__ __align__(8) uint64_t barFullA[BTC_ASTAGES], barEmptyA[BTC_ASTAGES], barFullB[BTC_BSTAGES], barEmptyB[BTC_BSTAGES], barAccFull, barAccEmpty, barMetaFree[2];
    __shared__ uint32_t sTmem;
    __shared__ TileMeta sMeta[2];
    __shared__ float sRw[232]; // per-texel weight sums (depth 0..195, irradiance 196..231)
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, N = bp.raysPerProbe;
    unsigned long long pacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // diagnostics (VKX_BLEND_PROFILE): cycle counters kept in registers, flushed once at the end
    const uint32_t chunks = (N + BTC_KC - 1u) / BTC_KC;
    const uint32_t numTiles = (bp.count + BTC_P - 1u) / BTC_P;
    const float cellLen = bp.gridCellLen;
    if (tid == 0) {
        for (uint32_t s = 0; s < BTC_ASTAGES; ++s) { mbarInit(&barFullA[s], 1); mbarInit(&barEmptyA[s], 1); }
        for (uint32_t s = 0; s < BTC_BSTAGES; ++s) { mbarInit(&barFullB[s], BTC_PROD_WARPS); mbarInit(&barEmptyB[s], 1); }
        mbarInit(&barAccFull, 1); mbarInit(&barAccEmpty, BTC_EPI_WARPS); mbarInit(&barMetaFree[0], BTC_EPI_WARPS); mbarInit(&barMetaFree[1], BTC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t t = tid; t < 232u; t += BTC_THREADS) sRw[t] = __ldg(W + size_t(BLEND_WSUM_ROW) * BLEND_COLS + (t < 196u ? t : BLEND_IRR_COL0 + t - 196u));
    for (uint32_t t = tid; t < 2u * BTC_P; t += BTC_THREADS) { sMeta[t / BTC_P].outOfRange[t % BTC_P] = 0u; sMeta[t / BTC_P].maxChange[t % BTC_P] = 0u; }
    if (warp == BTC_MMA_WARP) { tmemAlloc(&sTmem, 512); tmemRelinquish(); }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmem = sTmem;

    if (warp == BTC_TMA_WARP) {
        // ------------------------------------------------------------------------------------------ weights: one bulk copy per chunk
        if (lane == 0) {
            uint32_t g = 0; // chunks issued by this CTA so far (stage = g % STAGES, use = g / STAGES)
            for (uint32_t lt = blockIdx.x; lt < numTiles; lt += gridDim.x)
                for (uint32_t c = 0; c < chunks; ++c, ++g) {
                    const uint32_t s = g % BTC_ASTAGES, use = g / BTC_ASTAGES;
                    if (use) BTC_TIMED(0, mbarWait(&barEmptyA[s], (use - 1u) & 1u));
                    mbarExpectTx(&barFullA[s], BTC_A_CHUNK_BYTES);
                    bulkCopyG2S(smem + s * BTC_A_CHUNK_BYTES, reinterpret_cast<const char*>(image) + size_t(c) * BTC_A_CHUNK_BYTES, BTC_A_CHUNK_BYTES, &barFullA[s]);
                }
        }
    } else if (warp == BTC_MMA_WARP) {
        // ------------------------------------------------------------------------------------------ MMA issue
        if (lane == 0) {
            const long long tStart = clock64();
            uint32_t g = 0, it = 0;
            constexpr uint32_t idD = ummaIdesc(2 * BTC_P), idC = ummaIdesc(3 * BTC_P);
            for (uint32_t lt = blockIdx.x; lt < numTiles; lt += gridDim.x, ++it) {
                if (it) { BTC_TIMED(1, mbarWait(&barAccEmpty, (it - 1u) & 1u)); tcFenceAfter(); } // the epilogue has read the previous tile's accumulators
                for (uint32_t c = 0; c < chunks; ++c, ++g) {
                    const uint32_t sa = g % BTC_ASTAGES, useA = g / BTC_ASTAGES, sb = g % BTC_BSTAGES, useB = g / BTC_BSTAGES;
                    BTC_TIMED(2, mbarWait(&barFullA[sa], useA & 1u));
                    BTC_TIMED(3, mbarWait(&barFullB[sb], useB & 1u));
                    tcFenceAfter();
                    const long long tIssue = prof ? clock64() : 0;
                    const uint32_t aHi0 = smemAddr(smem + sa * BTC_A_CHUNK_BYTES), aHi1 = aHi0 + BTC_A_TILE_BYTES, aLo0 = aHi0 + 2 * BTC_A_TILE_BYTES, aLo1 = aHi0 + 3 * BTC_A_TILE_BYTES;
                    const uint32_t dHi = smemAddr(smemB + sb * BTC_B_STAGE_BYTES), dLo = dHi + BTC_BD_TILE_BYTES, cHi = dHi + 2 * BTC_BD_TILE_BYTES, cLo = cHi + BTC_BC_TILE_BYTES;
#pragma unroll
                    for (uint32_t ks = 0; ks < BTC_KC / 8u; ++ks) {
                        const uint32_t ko = ks * BTC_KSTEP_BYTES; // two K cores per MMA
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
                    const long long tCommit = prof ? clock64() : 0;
                    if (prof) pacc[11] += (unsigned long long)(tCommit - tIssue);
                    ummaCommit(&barEmptyA[sa]);                    // stages free when these MMAs have read them
                    ummaCommit(&barEmptyB[sb]);
                    if (c + 1u == chunks) { ummaCommit(&barAccFull); if (prof && it == 0) pacc[15] = (unsigned long long)(clock64() - tStart); } // accumulators of the tile complete
                    if (prof) pacc[12] += (unsigned long long)(clock64() - tCommit);
                }
            }
            if (prof) pacc[4] = (unsigned long long)(clock64() - tStart);
        }
    } else if (warp >= BTC_EPI_WARPS) {
        // ------------------------------------------------------------------------------------------ producers: ray records -> operand tiles
        // A thread owns one K core (four consecutive rays = 64 contiguous bytes of ray records) of one probe per chunk: probe pt / 4,
        // core pt % 4. Operand rows are plane-major (depth tile: d of probe p in row p, d^2 in row 64 + p; colour tile: r, g, b in rows
        // p, 64 + p, 128 + p), so the eight lanes of a quarter-warp (two neighbouring probes x four cores) write 128 contiguous,
        // swizzled bytes: every 128-bit store is conflict-free, ten stores per thread and chunk instead of forty 32-bit ones.
        const uint32_t pt = tid - BTC_EPI_WARPS * 32u;
        static_assert(BTC_PROD_WARPS * 32u == BTC_P * (BTC_KC / 4u), "one (probe, K core) pair per producer thread");
        const uint32_t myP = pt >> 2, myCore = pt & 3u;
        const uint32_t off0 = tileOffset(myP, 4u * myCore), off1 = tileOffset(BTC_P + myP, 4u * myCore), off2 = tileOffset(2u * BTC_P + myP, 4u * myCore);
        uint32_t g = 0, it = 0;
        for (uint32_t lt = blockIdx.x; lt < numTiles; lt += gridDim.x, ++it) {
            const uint32_t tile = lt;
            const uint32_t slot0 = tile * BTC_P;
            const uint32_t np = min(uint32_t(BTC_P), bp.count - slot0);
            TileMeta& meta = sMeta[it & 1u];
            if (it >= 2u) mbarWait(&barMetaFree[it & 1u], ((it >> 1) - 1u) & 1u); // the epilogue two tiles back is done with this meta block
            const float4* myRays = rays + size_t(slot0 + myP) * N;
            auto fetch = [&](uint32_t c, float4 (&rd)[4]) {
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    const uint32_t ray = c * BTC_KC + 4u * myCore + j;
                    rd[j] = (myP < np && ray < N) ? __ldcs(myRays + ray) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            // the ray records of chunks c + 1 and c + 2 are in flight while chunk c is converted
            float4 cur[4], nxt[4], nx2[4];
            fetch(0, cur);
            if (1u < chunks) fetch(1u, nxt);
            uint32_t outOfRange = 0;
            for (uint32_t c = 0; c < chunks; ++c, ++g) {
                const uint32_t s = g % BTC_BSTAGES, use = g / BTC_BSTAGES;
                if (c + 2u < chunks) fetch(c + 2u, nx2);
                if (use) { if (pt == 0) BTC_TIMED(6, mbarWait(&barEmptyB[s], (use - 1u) & 1u)); else mbarWait(&barEmptyB[s], (use - 1u) & 1u); } // the MMAs that read this stage have completed
                const long long tConv = (prof && pt == 0) ? clock64() : 0;
                unsigned char* bD = smemB + s * BTC_B_STAGE_BYTES;                   // depth planes hi, then lo
                unsigned char* bC = bD + 2 * BTC_BD_TILE_BYTES;                      // colour planes hi, then lo
                {
                    float4 dH, dL, d2H, d2L, rH, rL, gH, gL, bH, bL;
                    float* const dst[10] = {&dH.x, &dL.x, &d2H.x, &d2L.x, &rH.x, &rL.x, &gH.x, &gL.x, &bH.x, &bL.x};
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
                        float4 rd = cur[j];
                        if (myP < np && c * BTC_KC + 4u * myCore + j < N) {
                            if (rd.w < 0.0f || rd.w > cellLen) ++outOfRange;               // probesUpdate.glsl:74
                            float depth = minS(cellLen, rd.w);                              // :78-79
                            if (depth < 0.0f) depth = cellLen;
                            rd.w = depth;
                        }
                        splitTf32(rd.w, dst[0][j], dst[1][j]);
                        splitTf32(rd.w * rd.w, dst[2][j], dst[3][j]);
                        splitTf32(rd.x, dst[4][j], dst[5][j]);
                        splitTf32(rd.y, dst[6][j], dst[7][j]);
                        splitTf32(rd.z, dst[8][j], dst[9][j]);
                    }
                    *reinterpret_cast<float4*>(bD + off0) = dH;   *reinterpret_cast<float4*>(bD + BTC_BD_TILE_BYTES + off0) = dL;
                    *reinterpret_cast<float4*>(bD + off1) = d2H;  *reinterpret_cast<float4*>(bD + BTC_BD_TILE_BYTES + off1) = d2L;
                    *reinterpret_cast<float4*>(bC + off0) = rH;   *reinterpret_cast<float4*>(bC + BTC_BC_TILE_BYTES + off0) = rL;
                    *reinterpret_cast<float4*>(bC + off1) = gH;   *reinterpret_cast<float4*>(bC + BTC_BC_TILE_BYTES + off1) = gL;
                    *reinterpret_cast<float4*>(bC + off2) = bH;   *reinterpret_cast<float4*>(bC + BTC_BC_TILE_BYTES + off2) = bL;
                }
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) { cur[j] = nxt[j]; nxt[j] = nx2[j]; }
                if (c + 1u == chunks && outOfRange) atomicAdd(&meta.outOfRange[myP], outOfRange); // visible to the epilogue through the arrive below
                fenceProxyAsync(); // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbarArrive(&barFullB[s]);
                if (prof && pt == 0) pacc[7] += (unsigned long long)(clock64() - tConv);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue
        // A warp reads the TMEM lane quarter warp % 4, so rows 32 q .. 32 q + 31 of BOTH weight tiles belong to the two warps q and
        // q + 4; they split the tile's probes (warp q: probes 0..31, warp q + 4: probes 32..63), which balances the irradiance rows
        // (three channels, 11-bit packing, the max-change reduction) that all live in quarters 2 and 3 of tile 1. A thread owns up
        // to two texels: row r of tile 0 = depth texel r, and row r of tile 1 = depth texel 128 + r (r < 68) or irradiance texel
        // r - 68 (68 <= r < 104).
        const uint32_t quarter = warp & 3u, half = warp >> 2;
        const uint32_t row = quarter * 32u + lane;                // accumulator row = TMEM lane
        const uint32_t laneBase = (quarter * 32u) << 16;
        const bool bDepth = row < 68u, bIrr = row >= 68u && row < 104u;
        const uint32_t teA = row, teB = bDepth ? 128u + row : row - 68u;
        uint32_t interiorA = 0, borderA[3] = {0, 0, 0}, interiorB = 0, borderB[3] = {0, 0, 0}; int nbA = 0, nbB = 0;
        texelTargets(16, int(teA % 14u) + 1, int(teA / 14u) + 1, pr.depW, interiorA, borderA, nbA);
        if (bDepth) texelTargets(16, int(teB % 14u) + 1, int(teB / 14u) + 1, pr.depW, interiorB, borderB, nbB);
        else if (bIrr) texelTargets(8, int(teB % 6u) + 1, int(teB / 6u) + 1, pr.irrW, interiorB, borderB, nbB);
        const float rwA = sRw[teA], rwB = bDepth ? sRw[teB] : (bIrr ? sRw[196u + teB] : 0.0f);
        const bool normA = rwA > 1e-3f, normB = rwB > 1e-3f;
        const float rwInvA = normA ? __frcp_rn(rwA) : 0.0f, rwInvB = normB ? __frcp_rn(rwB) : 0.0f;
        const float hysteresis = bp.grid.hysteresis;
        const bool warpHasBDepth = quarter * 32u < 68u;          // warp-uniform: quarters 0, 1, 2 hold depth rows of tile 1
        const bool warpHasIrr = quarter >= 2u;                    // warp-uniform: irradiance rows 68..103 live in quarters 2 and 3
        const uint32_t* prevAtlasB = bDepth ? pr.depWork : pr.irrWork;
        const uint32_t pBegin = half * (BTC_P / 2u);
        // Write-out job of this thread (after the drain the staging area holds the tile's new interior texels): 128-bit piece k of a
        // probe's two atlas tiles, k < 64: row k / 4 of the 16 x 16 depth tile, texels 4 (k % 4) .. + 3; k >= 64: row (k - 64) / 2 of the
        // 8 x 8 irradiance tile. Border texels take their value from the interior texel probesCopyBorders.comp copies them from
        // (blendBorderSource). Three probes per pass (240 of the 256 threads); the four lanes of a row write 64 contiguous bytes:
        // full sectors only (scattered 4-byte stores straight from the accumulator rows took 100 k cycles per tile instead of 17 k).
        const uint32_t woProbe = tid / 80u, woK = tid % 80u; // woProbe == 3: idle
        const bool woDepth = woK < 64u;
        const uint32_t woRow = woDepth ? woK >> 2 : (woK - 64u) >> 1, woQuad = woDepth ? woK & 3u : (woK - 64u) & 1u;
        uint32_t woRel[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int T = woDepth ? 16 : 8, x = int(4u * woQuad) + j, y = int(woRow);
            int sx = x, sy = y;
            if (x == 0 || y == 0 || x == T - 1 || y == T - 1) blendBorderSource(T, x, y, sx, sy);
            const uint32_t te = uint32_t((sy - 1) * (T - 2) + (sx - 1));
            // staging slot of an interior texel relative to its probe's block: depth texel te < 128 -> column te of plane A, else column
            // te - 128 of plane B; irradiance texel ti -> column 68 + ti of plane B (a plane = one word per epilogue thread)
            woRel[j] = woDepth ? (te < 128u ? te : BTC_EPI_WARPS * 32u + te - 128u) : BTC_EPI_WARPS * 32u + 68u + te;
        }
        uint32_t it = 0;
        for (uint32_t lt = blockIdx.x; lt < numTiles; lt += gridDim.x, ++it) {
            const uint32_t tile = lt;
            const uint32_t slot0 = tile * BTC_P;
            const uint32_t np = min(uint32_t(BTC_P), bp.count - slot0);
            const uint32_t pEnd = min(np, pBegin + BTC_P / 2u);
            TileMeta& meta = sMeta[it & 1u];
            // While the tensor core accumulates this tile the epilogue warps are idle: they work out where the tile's probes live and pull
            // the previous texels (the work atlases were last touched a frame ago: DRAM) into L2.
            if (tid < BTC_P) {
                uint32_t lin = 0, oD = 0, oI = 0;
                if (tid < np) {
                    lin = __ldg(probeIndices + slot0 + tid);
                    int ix, iy, iz; probeGridIndex(lin, bp.grid, ix, iy, iz);
                    const int tl = iy * bp.grid.resolution[0] + ix;
                    oD = uint32_t(size_t(16 * iz) * pr.depW + size_t(16 * tl));
                    oI = uint32_t(size_t(8 * iz) * pr.irrW + size_t(8 * tl));
                }
                meta.linear[tid] = lin; meta.originD[tid] = oD; meta.originI[tid] = oI;
            }
            namedBarrier(2, BTC_EPI_WARPS * 32u);
            const uint32_t* originB = bDepth ? meta.originD : meta.originI;
            // Previous texels of this thread's probes -> its private column of the staging area (the same thread reads them back: no
            // barrier). Real loads, eight in flight, issued a whole main loop before they are needed: the drain below touches global
            // memory only with stores, and those go to sectors these loads have just brought into L2. (prefetch.global.L2 hints were
            // dropped under load: drains of 17 k cycles with them honoured, 125 k without.)
            for (uint32_t p0 = pBegin; p0 < pEnd; p0 += 8u) {
                uint32_t a[8], b[8];
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) {
                    const bool in = p0 + q < pEnd;
                    a[q] = in ? pr.depWork[meta.originD[p0 + q] + interiorA] : 0u;
                    b[q] = (in && (bDepth || bIrr)) ? prevAtlasB[originB[p0 + q] + interiorB] : 0u;
                }
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) { smemPrev[((p0 - pBegin + q) * 2u + 0u) * (BTC_EPI_WARPS * 32u) + tid] = a[q]; smemPrev[((p0 - pBegin + q) * 2u + 1u) * (BTC_EPI_WARPS * 32u) + tid] = b[q]; }
            }
            if (tid == 0) BTC_TIMED(8, mbarWait(&barAccFull, it & 1u)); else mbarWait(&barAccFull, it & 1u);
            __syncwarp(); // reconverge before the warp-aligned tcgen05.ld below
            tcFenceAfter();
            const long long tEpi = (prof && tid == 0) ? clock64() : 0;
            for (uint32_t p0 = pBegin; p0 < pEnd; p0 += 8u) { // eight probes per pass
                const uint32_t npj = min(8u, pEnd - p0);
                uint32_t prevA[8], prevB[8];
#pragma unroll
                for (uint32_t q = 0; q < 8; ++q) { prevA[q] = smemPrev[((p0 - pBegin + q) * 2u + 0u) * (BTC_EPI_WARPS * 32u) + tid]; prevB[q] = smemPrev[((p0 - pBegin + q) * 2u + 1u) * (BTC_EPI_WARPS * 32u) + tid]; }
                { // tile 0: depth texel teA of probes p0 .. p0 + 7 (columns p: sum of w d, columns P + p: sum of w d^2)
                    uint32_t v0[8], v1[8];
                    const long long tT = (prof && tid == 0) ? clock64() : 0;
                    tmemLoad8(tmem + laneBase + p0, v0);
                    tmemLoad8(tmem + laneBase + BTC_P + p0, v1);
                    tmemLoadWait();
                    if (prof && tid == 0) pacc[5] += (unsigned long long)(clock64() - tT);
#pragma unroll
                    for (uint32_t q = 0; q < 8; ++q) {
                        if (q >= npj) break;
                        float r0 = __uint_as_float(v0[q]), r1 = __uint_as_float(v1[q]);
                        if (normA) { r0 = divShared(r0, rwA, rwInvA); r1 = divShared(r1, rwA, rwInvA); } // probesUpdate.glsl:85-86
                        const float2 prev = unpackRG16F(prevA[q]);
                        const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis); // :103
                        const uint32_t word = packRG16Fx2(o0, o1);
                        smemPrev[((p0 - pBegin + q) * 2u + 0u) * (BTC_EPI_WARPS * 32u) + tid] = word; // the new texel replaces the previous one in this thread's staging column
                        if (depUnpacked) { float* up = depUnpacked + (size_t(slotBase + slot0 + p0 + q) * 196 + teA) * 2; up[0] = o0; up[1] = o1; }
                    }
                }
                if (warpHasBDepth) { // tile 1, depth rows (columns 2 P + p and 3 P + p)
                    uint32_t v0[8], v1[8];
                    tmemLoad8(tmem + laneBase + 2u * BTC_P + p0, v0);
                    tmemLoad8(tmem + laneBase + 3u * BTC_P + p0, v1);
                    tmemLoadWait();
                    if (bDepth) {
#pragma unroll
                        for (uint32_t q = 0; q < 8; ++q) {
                            if (q >= npj) break;
                            float r0 = __uint_as_float(v0[q]), r1 = __uint_as_float(v1[q]);
                            if (normB) { r0 = divShared(r0, rwB, rwInvB); r1 = divShared(r1, rwB, rwInvB); }
                            const float2 prev = unpackRG16F(prevB[q]);
                            const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis);
                            const uint32_t word = packRG16Fx2(o0, o1);
                            smemPrev[((p0 - pBegin + q) * 2u + 1u) * (BTC_EPI_WARPS * 32u) + tid] = word;
                            if (depUnpacked) { float* up = depUnpacked + (size_t(slotBase + slot0 + p0 + q) * 196 + teB) * 2; up[0] = o0; up[1] = o1; }
                        }
                    }
                }
                if (warpHasIrr) { // tile 1, irradiance rows (columns 4 P + p: red, 5 P + p: green, 6 P + p: blue)
                    uint32_t vr[8], vg[8], vb[8];
                    tmemLoad8(tmem + laneBase + 4u * BTC_P + p0, vr);
                    tmemLoad8(tmem + laneBase + 5u * BTC_P + p0, vg);
                    tmemLoad8(tmem + laneBase + 6u * BTC_P + p0, vb);
                    tmemLoadWait();
#pragma unroll
                    for (uint32_t q = 0; q < 8; ++q) {
                        if (q >= npj) break; // warp-uniform
                        float maxChange = 0.0f;
                        if (bIrr) {
                            float r0 = __uint_as_float(vr[q]), r1 = __uint_as_float(vg[q]), r2 = __uint_as_float(vb[q]);
                            if (normB) { r0 = divShared(r0, rwB, rwInvB); r1 = divShared(r1, rwB, rwInvB); r2 = divShared(r2, rwB, rwInvB); }
                            const float3 prev = unpackR11G11B10(prevB[q]);
                            maxChange = maxS(maxS(fabsf(r0 - prev.x), fabsf(r1 - prev.y)), fabsf(r2 - prev.z));
                            const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis), o2 = mixf(r2, prev.z, hysteresis);
                            const uint32_t word = packR11G11B10(o0, o1, o2);
                            smemPrev[((p0 - pBegin + q) * 2u + 1u) * (BTC_EPI_WARPS * 32u) + tid] = word;
                            if (irrUnpacked) { float* up = irrUnpacked + (size_t(slotBase + slot0 + p0 + q) * 36 + teB) * 3; up[0] = o0; up[1] = o1; up[2] = o2; }
                        }
                        // probesUpdate.glsl:106-107: maximum over the probe's 36 texels (non-negative floats order like their bit patterns)
                        const uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(maxChange));
                        if (lane == 0 && wmax) atomicMax(&meta.maxChange[p0 + q], wmax);
                    }
                }
            }
            // tensor memory is free for the next tile
            tcFenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(&barAccEmpty);
            if (prof && tid == 0) { const unsigned long long d = (unsigned long long)(clock64() - tEpi); pacc[9] += d; if (it == 0) pacc[13] = d; else if (it == 1) pacc[14] = d; }
            const long long tBar = (prof && tid == 0) ? clock64() : 0;
            namedBarrier(1, BTC_EPI_WARPS * 32u); // (aligned barrier: every lane of the warp must execute the same instruction - no divergent timing wrapper here)
            if (prof && tid == 0) pacc[10] += (unsigned long long)(clock64() - tBar); // every texel of the tile is mixed: maxChange is complete
            if (tid < np) { // state machine, probesUpdate.glsl:110-119 (decree A.5.3: full max over the 36 texels)
                const uint32_t linearIndex = meta.linear[tid];
                uint32_t stt = pr.stateWork[linearIndex];
                if (meta.outOfRange[tid] >= N) stt = 8;
                else {
                    const float maxChange = __uint_as_float(meta.maxChange[tid]);
                    if (maxChange < 0.02f / float(stt)) stt = min(stt + 1u, 8u);
                    else if (maxChange > 0.04f / float(stt)) stt = max(stt - 1u, 1u);
                    else if (maxChange > 0.25f) stt = 1;
                }
                pr.stateWork[linearIndex] = stt;
            }
            if (woProbe < 3u) {
                for (uint32_t p = woProbe; p < np; p += 3u) {
                    // probe p's block: local probe p % 32 of half p / 32 -> threads 128 (p / 32) .. + 127 of both planes
                    const uint32_t* blk = smemPrev + (p & 31u) * 2u * (BTC_EPI_WARPS * 32u) + (p >> 5) * 128u;
                    const uint4 w = make_uint4(blk[woRel[0]], blk[woRel[1]], blk[woRel[2]], blk[woRel[3]]);
                    if (woDepth) *reinterpret_cast<uint4*>(pr.depWork + meta.originD[p] + size_t(woRow) * pr.depW + 4u * woQuad) = w;
                    else *reinterpret_cast<uint4*>(pr.irrWork + meta.originI[p] + size_t(woRow) * pr.irrW + 4u * woQuad) = w;
                }
            }
            if (tid < BTC_P) { meta.outOfRange[tid] = 0u; meta.maxChange[tid] = 0u; }
            namedBarrier(3, BTC_EPI_WARPS * 32u); // the staging area and the tile's meta block may be rewritten (next tile's previous texels)
            if (lane == 0) mbarArrive(&barMetaFree[it & 1u]); // release semantics: the zeroes above are visible to the producers that wait
        }
    }
    if (prof) { // slot owners: TMA lane (0), MMA lane (1-4, 11, 12), first producer thread (5-7), first epilogue thread (8-10)
        const bool owner[16] = {warp == BTC_TMA_WARP && lane == 0, warp == BTC_MMA_WARP && lane == 0, warp == BTC_MMA_WARP && lane == 0, warp == BTC_MMA_WARP && lane == 0, warp == BTC_MMA_WARP && lane == 0,
                                tid == 0, tid == BTC_EPI_WARPS * 32u, tid == BTC_EPI_WARPS * 32u, tid == 0, tid == 0, tid == 0, warp == BTC_MMA_WARP && lane == 0, warp == BTC_MMA_WARP && lane == 0, tid == 0, tid == 0, warp == BTC_MMA_WARP && lane == 0};
#pragma unroll
        for (int k = 0; k < 16; ++k) if (owner[k]) prof[blockIdx.x * 16 + k] = pacc[k];
    }
    tcFenceBefore();
    __syncthreads();
    if (warp == BTC_MMA_WARP) tmemFree(tmem, 512);
}

int blendTcWeights(vkx_ctx* ctx, cudaStream_t st) {
    k_blend_weight_image<<<(BTC_MAX_CHUNKS * 2 * 128 * 16) / 256, 256, 0, st>>>(ctx->grid.raysPerProbe, ctx->dBlendW, ctx->dBlendImage); LAUNCH_CHECK(ctx);
    return VKX_OK;
}

int blendTcLaunch(vkx_ctx* ctx, const BlendParams& bp, const DeviceProbes& pr, const uint32_t* idx, uint32_t n, uint32_t slotBase, cudaStream_t st) {
    if (n == 0) return VKX_OK;
    if (!ctx->blendTcAttrSet) { CUDA_TRY(ctx, cudaFuncSetAttribute(k_blend_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, BTC_SMEM_BYTES)); ctx->blendTcAttrSet = true; }
    const unsigned grid = std::min<unsigned>(divUp(n, BTC_P), unsigned(ctx->smCount));
    static const bool profile = getenv("VKX_BLEND_PROFILE") != nullptr;
    unsigned long long* prof = nullptr;
    if (profile) { CUDA_TRY(ctx, cudaMallocManaged(&prof, size_t(grid) * 16 * sizeof(unsigned long long))); memset(prof, 0, size_t(grid) * 16 * sizeof(unsigned long long)); }
    k_blend_tc<<<grid, BTC_THREADS, BTC_SMEM_BYTES, st>>>(bp, pr, idx, ctx->dRays, ctx->dBlendW, ctx->dBlendImage, ctx->debugBuffers ? ctx->dIrrUnpacked : nullptr,
                                                                      ctx->debugBuffers ? ctx->dDepUnpacked : nullptr, slotBase, prof); LAUNCH_CHECK(ctx);
    if (prof) { // diagnostics: mean / max cycles per CTA of every instrumented wait (slots: see BTC_TIMED uses)
        cudaStreamSynchronize(st);
        static const char* names[16] = {"tma: wait emptyA", "mma: wait accEmpty", "mma: wait fullA", "mma: wait fullB", "mma: total", "epi: tmem ld tile0 blk", "prod: wait emptyB", "prod: convert+arrive", "epi: wait accFull", "epi: drain+mix", "epi: named barrier", "mma: issue 18 MMAs", "mma: commits", "epi: drain of tile 0", "epi: drain of tile 1", "mma: first tile issued"};
        for (int k = 0; k < 16; ++k) { double sum = 0, mx = 0; for (unsigned b = 0; b < grid; ++b) { const double v = double(prof[b * 16 + k]); sum += v; mx = v > mx ? v : mx; } fprintf(stderr, "[blend_tc profile] %-22s mean %10.0f max %10.0f cycles per CTA (%u CTAs, %u probes)\n", names[k], sum / grid, mx, grid, n); }
        if (getenv("VKX_BLEND_PROFILE")[0] == '2') for (unsigned b = 0; b < grid; b += 12) { fprintf(stderr, "[blend_tc cta %3u]", b); for (int k = 0; k < 16; ++k) fprintf(stderr, " %7llu", prof[b * 16 + k]); fprintf(stderr, "\n"); }
        cudaFree(prof);
    }
    return VKX_OK;
}
