// On-device probe scheduler: IrradianceProbes::selectProbesToUpdate (reference src/IrradianceProbes.cpp:396-424) as a stream
// compaction over the device-resident probe states, so a frame no longer reads P state words back and uploads a list
// (the reference maps the state buffer, loops over it on the host and refills the to-update buffer every update).
//
// The reference scans idx = lastUpdateOffset, lastUpdateOffset + 1, ... (wrapping once, which bumps s_LoopIndex) and stops
// after P probes or, if ProbesPerUpdate != 0, as soon as that many are selected. With j the position in that scan:
//   idx_j = (offset + j) mod P,  loop_j = loopIndex + (offset + j >= P),  selected_j = state != 0 && (idx_j + loop_j) % state == 0
//   slot_j = #selected before j (exclusive scan);  the list is { idx_j : selected_j && (K == 0 || slot_j < K) }
//   checked = P, or j* + 1 where j* holds the K-th selected probe;  the new offset / loop index follow from offset + checked.
// A second compaction, over the cached 2x2x2-block order of the probes, produces the slot schedule (`order`) that
// uploadOrder() builds on the host for host-provided lists. Only 16 bytes (count, checked) return to the host.
#include "common.cuh"
#include <cub/device/device_scan.cuh>

namespace {

__global__ void k_sched_flags(const uint32_t* __restrict__ state, uint32_t P, uint32_t offset, uint32_t loopIndex, uint32_t* __restrict__ flags) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P) return;
    uint32_t idx = offset + j, loop = loopIndex;
    if (idx >= P) { idx -= P; ++loop; }
    const uint32_t s = state[idx];
    flags[j] = (s != 0u && ((idx + loop) % s) == 0u) ? 1u : 0u;
}

// list[slot] = probe, slotOf[probe] = slot + 1, result = {count, checked}
__global__ void k_sched_compact(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ slots, uint32_t P, uint32_t offset, uint32_t K,
                                uint32_t* __restrict__ list, uint32_t* __restrict__ slotOf, uint32_t* __restrict__ result) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P) return;
    const uint32_t f = flags[j], s = slots[j];
    if (f && (K == 0u || s < K)) {
        uint32_t idx = offset + j; if (idx >= P) idx -= P;
        list[s] = idx; slotOf[idx] = s + 1u;
        if (K != 0u && s == K - 1u) result[1] = j + 1u; // the scan stops right after the K-th selected probe
    }
    if (j == P - 1u) { const uint32_t total = s + f; result[0] = (K != 0u && total > K) ? K : total; if (K == 0u || total < K) result[1] = P; }
}

__global__ void k_sched_block_flags(const uint32_t* __restrict__ blockedOrder, const uint32_t* __restrict__ slotOf, uint32_t P, uint32_t* __restrict__ flags) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P) flags[r] = slotOf[blockedOrder[r]] != 0u ? 1u : 0u;
}
__global__ void k_sched_order(const uint32_t* __restrict__ blockedOrder, const uint32_t* __restrict__ slotOf, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, uint32_t P,
                              uint32_t* __restrict__ order) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < P && flags[r]) order[pos[r]] = slotOf[blockedOrder[r]] - 1u;
}

} // namespace

// Fills ctx->dIndicesList / ctx->dOrder for the next update and returns the list length; advances the scheduler state.
int scheduleProbes(vkx_ctx* ctx, uint32_t probesPerUpdate, uint32_t* countOut) {
    cudaStream_t st = ctx->stream;
    const uint32_t P = ctx->probeCount;
    if (!ctx->dSchedFlags) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->dSchedFlags, size_t(P) * 4)); CUDA_TRY(ctx, cudaMalloc(&ctx->dSchedPos, size_t(P) * 4));
        CUDA_TRY(ctx, cudaMalloc(&ctx->dSchedSlotOf, size_t(P) * 4)); CUDA_TRY(ctx, cudaMalloc(&ctx->dSchedResult, 16));
        CUDA_TRY(ctx, cudaMallocHost(&ctx->hSchedResult, 16));
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->dSchedFlags, ctx->dSchedPos, int(P), st);
        ctx->schedTempBytes = need; CUDA_TRY(ctx, cudaMalloc(&ctx->dSchedTemp, need));
    }
    const unsigned blocks = divUp(P, 256);
    size_t need = ctx->schedTempBytes;
    k_sched_flags<<<blocks, 256, 0, st>>>(ctx->dStateSampled, P, ctx->schedOffset, ctx->schedLoopIndex, ctx->dSchedFlags); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->dSchedTemp, need, ctx->dSchedFlags, ctx->dSchedPos, int(P), st)); ctx->launches++;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dSchedSlotOf, 0, size_t(P) * 4, st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dSchedResult, 0, 16, st));
    k_sched_compact<<<blocks, 256, 0, st>>>(ctx->dSchedFlags, ctx->dSchedPos, P, ctx->schedOffset, probesPerUpdate, ctx->dIndicesList, ctx->dSchedSlotOf, ctx->dSchedResult); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hSchedResult, ctx->dSchedResult, 8, cudaMemcpyDeviceToHost, st));
    k_sched_block_flags<<<blocks, 256, 0, st>>>(ctx->dBlockedOrder, ctx->dSchedSlotOf, P, ctx->dSchedFlags); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->dSchedTemp, need, ctx->dSchedFlags, ctx->dSchedPos, int(P), st)); ctx->launches++;
    k_sched_order<<<blocks, 256, 0, st>>>(ctx->dBlockedOrder, ctx->dSchedSlotOf, ctx->dSchedFlags, ctx->dSchedPos, P, ctx->dOrder); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const uint32_t count = ctx->hSchedResult[0], checked = ctx->hSchedResult[1];
    const uint64_t end = uint64_t(ctx->schedOffset) + checked;
    if (end >= P) { ctx->schedOffset = uint32_t(end - P); ctx->schedLoopIndex++; } else ctx->schedOffset = uint32_t(end);
    ctx->schedCount = count;
    if (countOut) *countOut = count;
    return VKX_OK;
}
