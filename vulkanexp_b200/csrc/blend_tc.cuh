// Shared declarations of the two blend kernels: k_blend (CUDA cores, ddgi.cu; also the variant that stores into peer memory) and
// k_blend_tc (tcgen05 tensor cores, blend_tc.cu).
#pragma once
#include "common.cuh"

struct BlendParams {
    vkx_grid_info grid;
    uint32_t raysPerProbe, count;
    float gridCellLen;   // length(probeGridCellSize)
};

// Per-frame weight table shared by every probe: W[ray][col], col 0..195 depth texels pow(max(0, dot), sharpness), col 224..259
// irradiance texels max(0, dot) (probesUpdate.glsl:60,75,80); row BLEND_WSUM_ROW holds the per-texel weight sums.
#define BLEND_COLS 288
#define BLEND_IRR_COL0 224
#define BLEND_WSUM_ROW VKX_MAX_RAYS_PER_PROBE

__device__ __forceinline__ void blendBorderSource(int T, int x, int y, int& sx, int& sy) { // probesCopyBorders.comp:21-220 as a formula
    const int L = T - 1;
    const bool bx = (x == 0 || x == L), by = (y == 0 || y == L);
    if (bx && by) { sx = x == 0 ? L - 1 : 1; sy = y == 0 ? L - 1 : 1; }
    else if (bx) { sx = x == 0 ? 1 : L - 1; sy = L - y; }
    else { sx = L - x; sy = y == 0 ? 1 : L - 1; }
}

// ---- tensor-core blend geometry (blend_tc.cu)
#define BTC_P 64u            // probes per tile: N = 128 (depth planes) / 192 (colour planes)
#define BTC_KC 16u           // rays per chunk (two K = 8 TF32 MMAs)
#define BTC_ASTAGES 2u       // weight chunks in flight (TMA warp -> MMA warp); measured: the MMA warp waits < 3 % of its time for them
#define BTC_BSTAGES 2u       // ray-data chunks in flight (producer warps -> MMA warp)
#define BTC_EPI_WARPS 8u     // warps 0..3: accumulator rows of weight tile 0, warps 4..7: weight tile 1 (a warp reads the TMEM lane quarter warp % 4)
#define BTC_PROD_WARPS 8u    // warps 8..15: ray records -> TF32 hi / lo operand tiles
#define BTC_MMA_WARP 16u     // one lane issues tcgen05.mma
#define BTC_TMA_WARP 17u     // one lane issues the bulk copies of the weight image
#define BTC_THREADS (32u * (BTC_EPI_WARPS + BTC_PROD_WARPS + 2u))
#define BTC_MAX_CHUNKS (VKX_MAX_RAYS_PER_PROBE / 16)
// Operand tiles are K-major with 16 rays (64 bytes) per row. BTC_LAYOUT 1 (default): the canonical SWIZZLE_64B layout - 8-row groups of
// 512 bytes, the four 16-byte K cores of a row XOR-ed with (row / 2) % 4. BTC_LAYOUT 0: unswizzled "interleaved" core matrices
// (8 rows x 16 bytes contiguous, K cores BTC_LBO apart, row groups BTC_SBO apart) - kept for the measurement in DESIGN.md: its
// tcgen05.mma ran at a quarter of the tensor-core floor (operand fetch), the swizzled layout runs at the floor.
#ifndef BTC_LAYOUT
#define BTC_LAYOUT 1
#endif
#if BTC_LAYOUT == 1
#define BTC_SBO 512u
#define BTC_LBO 16u          // ignored by the hardware for swizzled K-major operands
#else
#define BTC_LBO 128u         // bytes between the 16-byte K cores of an operand tile (core matrix = 8 rows x 16 bytes)
#ifndef BTC_SBO
#define BTC_SBO 528u         // bytes between 8-row groups: 4 K cores + 16 bytes, so that 8 probes x 4 rays of a warp hit 32 different banks
#endif
#endif
#define BTC_A_TILE_BYTES (16u * BTC_SBO)                 // 128 weight rows
#define BTC_A_CHUNK_BYTES (4u * BTC_A_TILE_BYTES)        // hi tile 0, hi tile 1, lo tile 0, lo tile 1
#define BTC_BD_TILE_BYTES (16u * BTC_SBO)                // 128 rows: (probe, d | d^2)
#define BTC_BC_TILE_BYTES (24u * BTC_SBO)                // 192 rows: (probe, r | g | b)
#define BTC_B_STAGE_BYTES (2u * BTC_BD_TILE_BYTES + 2u * BTC_BC_TILE_BYTES)
#define BTC_PREV_BYTES (BTC_EPI_WARPS * 32u * BTC_P * 4u) // previous texels of a tile: (probes / 2) x 2 texels per epilogue thread
#define BTC_SMEM_BYTES (BTC_ASTAGES * BTC_A_CHUNK_BYTES + BTC_BSTAGES * BTC_B_STAGE_BYTES + BTC_PREV_BYTES)
#define BTC_IMAGE_BYTES (size_t(BTC_MAX_CHUNKS) * BTC_A_CHUNK_BYTES)

int blendTcWeights(vkx_ctx* ctx, cudaStream_t st);  // per frame, after k_blend_weights: the A-operand image
int blendTcLaunch(vkx_ctx* ctx, const BlendParams& bp, const DeviceProbes& pr, const uint32_t* idx, uint32_t n, uint32_t slotBase, cudaStream_t st);
