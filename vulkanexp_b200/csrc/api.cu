// C ABI of libvkexp_b200.so (include/vkx.h). Host-side orchestration only: every numeric result comes from the CUDA
// kernels in bvh_build.cu / ddgi.cu / shadow.cu. There is deliberately no CPU fallback: without a usable CUDA device
// vkx_create fails and nothing else can be called.
#include "common.cuh"
#include "blend_tc.cuh"
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <nccl.h>
#include <chrono>
#include <cstdlib>

static thread_local std::string g_createError;

int vkx_fail(vkx_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf; else g_createError = buf;
    return code;
}

template <typename T>
static int upload(vkx_ctx* ctx, T** dst, const T* src, size_t n) {
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    CUDA_TRY(ctx, cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CUDA_TRY(ctx, cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return VKX_OK;
}
#define TRY(expr) do { int _rc = (expr); if (_rc != VKX_OK) return _rc; } while (0)
#define BIND(ctx) do { if (!(ctx)) return VKX_E_INVALID; cudaError_t _e = cudaSetDevice((ctx)->device); if (_e != cudaSuccess) return vkx_fail((ctx), VKX_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(_e)); } while (0)

// Asynchronous read-backs. A request for the sampled atlases of a single-GPU context is only noted (ctx->copyRequests) and queued
// later: by the next update right before its primary traversal starts (ddgiUpdate), or by whatever needs the copy ordered first
// (download_wait, any writer of the sampled atlases). Queued right behind the publish, the copy ran while the next update issued its
// dozen small set-up launches, and every step was ~0.1 ms longer than with the copy alongside the 0.76 ms traversal kernel
// (tools/e2e_probe.py: the overhead grew with the bytes read back, not with the number of event interlocks). Nothing writes the
// sampled atlases between the request and that point, so the bytes are the same.
static int ensureCopyStream(vkx_ctx* ctx) {
    if (!ctx->copyStream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evPublished, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evCopyDone, cudaEventDisableTiming));
    }
    return VKX_OK;
}
int flushCopyRequests(vkx_ctx* ctx) {
    if (ctx->copyRequests.empty()) return VKX_OK;
    CUDA_TRY(ctx, cudaEventRecord(ctx->evPublished, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->evPublished, 0));
    for (const auto& op : ctx->copyRequests) CUDA_TRY(ctx, cudaMemcpyAsync(op.dst, op.src, op.bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
    ctx->copyRequests.clear();
    CUDA_TRY(ctx, cudaEventRecord(ctx->evCopyDone, ctx->copyStream));
    ctx->copyPending = true; ctx->copyReadsWork = ctx->copyRequestsReadWork;
    return VKX_OK;
}
// Before anything other than an update overwrites the sampled atlases (upload, classification, re-initialisation): the requested
// read-backs are queued now and the context's stream waits for them.
static int settleReadBack(vkx_ctx* ctx) {
    TRY(flushCopyRequests(ctx));
    if (ctx->copyPending) { CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopyDone, 0)); ctx->copyPending = false; ctx->copyReadsWork = false; }
    return VKX_OK;
}
// Single-GPU contexts: reads of the sampled atlases (next written by the next update's publish). Sharded contexts: reads of the rank's
// own slices from its work atlases (next written by the next update's blend; not with the fused peer stores, whose blend does not keep
// the work atlases current); other reads there go through the exchange bookkeeping (gather / alternating sets) and stay eager.
static bool deferReadBack(const vkx_ctx* ctx, bool fromWork) {
    static const bool off = [] { const char* e = getenv("VKX_READBACK"); return e && !strcmp(e, "eager"); }(); // A/B: queue behind the publish as before
    return !off && (ctx->nranks <= 1 ? !fromWork : (fromWork && (!ctx->p2p || ctx->p2pCopy)));
}
int waitGather(vkx_ctx* ctx) { // orders the context's stream after a pending exchange (all-gather or peer stores) of the sampled atlases
    if (ctx->gatherPending) { CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->gatherDone, 0)); ctx->gatherPending = false; }
    if (ctx->p2pPending) { ctx->p2pPending = false; int rc = launchP2pWait(ctx); if (rc != VKX_OK) return rc; }
    return VKX_OK;
}

extern "C" {

int vkx_abi_version(void) { return VKX_ABI_VERSION; }

int vkx_create(int device, vkx_ctx** out) {
    if (!out) return VKX_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return vkx_fail(nullptr, VKX_E_CUDA, "no CUDA device available (%s); libvkexp_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return vkx_fail(nullptr, VKX_E_INVALID, "device %d out of range (%d devices)", device, n);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return vkx_fail(nullptr, VKX_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    vkx_ctx* ctx = new vkx_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return vkx_fail(nullptr, VKX_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    ctx->smCount = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return vkx_fail(nullptr, VKX_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    for (auto& ev : ctx->sev) cudaEventCreate(&ev);
    for (auto& ev : ctx->kev) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&ctx->auxStream, cudaStreamNonBlocking);
    for (auto& ev : ctx->auxEvent) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    *out = ctx;
    return VKX_OK;
}

static void releaseP2p(vkx_ctx* ctx) { // the six atlas pointers live inside the slab while peer exchange is enabled
    if (!ctx->p2pSlab) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < VKX_MAX_RANKS; ++r) { if (ctx->peerSlab[r] && ctx->peerSlab[r] != ctx->p2pSlab) cudaIpcCloseMemHandle(ctx->peerSlab[r]); ctx->peerSlab[r] = nullptr; }
    cudaFree(ctx->p2pSlab); ctx->p2pSlab = nullptr;
    if (ctx->hFlagRing) { cudaFreeHost(ctx->hFlagRing); ctx->hFlagRing = nullptr; }
    ctx->p2pCopy = false;
    ctx->dIrrSampled = ctx->dDepSampled = ctx->dStateSampled = ctx->dIrrNext = ctx->dDepNext = ctx->dStateNext = nullptr;
    ctx->p2p = ctx->p2pPending = ctx->blendToPeers = false; ctx->p2pFrame = 0; ctx->p2pSampledSet = 0;
}

static void freeProbes(vkx_ctx* ctx) {
    releaseP2p(ctx);
    void* ptrs[] = {ctx->dIrrWork, ctx->dIrrSampled, ctx->dDepWork, ctx->dDepSampled, ctx->dStateWork, ctx->dStateSampled, ctx->dIndicesList, ctx->dDirs, ctx->dRays,
                    ctx->dHits, ctx->dShadowQueue, ctx->dShadowVis, ctx->dQueueCount, ctx->dShadowFlags, ctx->dIrrUnpacked, ctx->dDepUnpacked, ctx->dMissQueue, ctx->dFrontQueue, ctx->dFrontKeys, ctx->dFrontKeysOut, ctx->dFrontQueueSorted, ctx->dSortTemp, ctx->dIrrNext, ctx->dDepNext, ctx->dStateNext,
                    ctx->dPerm, ctx->dOrder, ctx->dBlockedOrder, ctx->dBlendW, ctx->dBlendImage, ctx->dPermList, ctx->dIota, ctx->dCellHist, ctx->dInvDirs, ctx->dOrigins};
    for (void* p : ptrs) if (p) cudaFree(p);
    ctx->dIrrWork = ctx->dIrrSampled = ctx->dDepWork = ctx->dDepSampled = ctx->dStateWork = ctx->dStateSampled = ctx->dIndicesList = nullptr;
    ctx->dIrrNext = ctx->dDepNext = ctx->dStateNext = nullptr;
    ctx->dPerm = ctx->dOrder = ctx->dBlockedOrder = ctx->dPermList = ctx->dIota = ctx->dCellHist = nullptr; ctx->dBlendW = nullptr; ctx->dBlendImage = nullptr; ctx->shardOrderReady = false;
    ctx->dDirs = nullptr; ctx->dInvDirs = nullptr; ctx->dOrigins = nullptr; ctx->dRays = nullptr; ctx->dHits = nullptr; ctx->dShadowQueue = nullptr; ctx->dShadowVis = nullptr; ctx->dQueueCount = nullptr; ctx->dShadowFlags = nullptr;
    ctx->dIrrUnpacked = ctx->dDepUnpacked = nullptr; ctx->dMissQueue = ctx->dFrontQueue = ctx->dFrontKeys = ctx->dFrontKeysOut = ctx->dFrontQueueSorted = nullptr; ctx->dSortTemp = nullptr; ctx->sortTempBytes = 0;
    if (ctx->hListStage) { cudaFreeHost(ctx->hListStage); ctx->hListStage = nullptr; }
    ctx->hLastList.clear();
    void* sched[] = {ctx->dSchedFlags, ctx->dSchedPos, ctx->dSchedSlotOf, ctx->dSchedResult, ctx->dSchedTemp};
    for (void* p : sched) if (p) cudaFree(p);
    if (ctx->hSchedResult) cudaFreeHost(ctx->hSchedResult);
    ctx->dSchedFlags = ctx->dSchedPos = ctx->dSchedSlotOf = ctx->dSchedResult = ctx->hSchedResult = nullptr; ctx->dSchedTemp = nullptr; ctx->schedTempBytes = 0;
    ctx->schedLoopIndex = ctx->schedOffset = ctx->schedCount = 0; ctx->schedValid = false;
    ctx->probesReady = false;
}
static void freeShadow(vkx_ctx* ctx) {
    void* ptrs[] = {ctx->dPosDepth, ctx->dNormalMetal, ctx->dShRaw, ctx->dShX, ctx->dShFinal[0], ctx->dShFinal[1], ctx->dShDirs, ctx->dShMask, ctx->dAlbedoRough, ctx->dEmissive, ctx->dReflection, ctx->dGathered,
                    ctx->dReflRaw, ctx->dReflX, ctx->dReflFinal[0], ctx->dReflFinal[1], ctx->dReflDirs, ctx->dReflHits, ctx->dReflMask, ctx->dReflQueue, ctx->dReflCount};
    for (void* p : ptrs) if (p) cudaFree(p);
    ctx->dPosDepth = ctx->dNormalMetal = ctx->dShRaw = ctx->dShX = ctx->dShFinal[0] = ctx->dShFinal[1] = ctx->dShDirs = nullptr; ctx->dShMask = nullptr;
    ctx->dAlbedoRough = ctx->dEmissive = ctx->dReflection = ctx->dGathered = nullptr;
    ctx->dReflRaw = ctx->dReflX = ctx->dReflFinal[0] = ctx->dReflFinal[1] = ctx->dReflDirs = nullptr; ctx->dReflHits = nullptr; ctx->dReflMask = nullptr; ctx->dReflQueue = ctx->dReflCount = nullptr; ctx->reflValid = false;
    ctx->shW = ctx->shH = 0;
}

void vkx_destroy(vkx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    flushCopyRequests(ctx); // requested read-backs still land in the caller's buffers
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) ncclCommDestroy(reinterpret_cast<ncclComm_t>(ctx->comm));
    freeProbes(ctx); freeShadow(ctx); freeTextures(ctx);
    if (ctx->dSrgbLut) cudaFree(ctx->dSrgbLut);
    if (ctx->dSrgbThreshold) cudaFree(ctx->dSrgbThreshold);
    void* ptrs[] = {ctx->dVertices, ctx->dIndices, ctx->dOffsets, ctx->dMeshCounts, ctx->dMaterials, ctx->dInstances, ctx->dWorldToObject, ctx->dInstTriBase, ctx->dNodes, ctx->dTris, ctx->dNoise, ctx->dRefitScratch};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->sev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->kev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->auxEvent) if (ev) cudaEventDestroy(ev);
    if (ctx->auxStream) cudaStreamDestroy(ctx->auxStream);
    if (ctx->commStream) cudaStreamDestroy(ctx->commStream);
    if (ctx->commEvent) cudaEventDestroy(ctx->commEvent);
    if (ctx->gatherDone) cudaEventDestroy(ctx->gatherDone);
    if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); cudaEventDestroy(ctx->evPublished); cudaEventDestroy(ctx->evCopyDone); }
    if (ctx->hStage) { cudaFreeHost(ctx->hStage); for (auto& e : ctx->stageEvent) if (e) cudaEventDestroy(e); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* vkx_last_error(vkx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }
uint64_t vkx_launch_count(vkx_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* vkx_stream(vkx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int vkx_sync(vkx_ctx* ctx) {
    BIND(ctx); TRY(waitGather(ctx)); CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->p2pSlab) { // the device-side wait gives up after ~10 s instead of hanging the GPU; report it here
        uint32_t err = 0;
        CUDA_TRY(ctx, cudaMemcpy(&err, ctx->p2pSlab + ctx->p2pFlagsOff + 256, 4, cudaMemcpyDeviceToHost));
        if (err) return vkx_fail(ctx, VKX_E_NCCL, "peer exchange: rank %u never signalled its tiles", err - 1u);
    }
    return VKX_OK;
}

// ---------------------------------------------------------------------------------------------------- geometry
int vkx_scene_upload(vkx_ctx* ctx, const vkx_vertex* vertices, size_t numVertices, const uint32_t* indices, size_t numIndices,
                     const vkx_offset_entry* offsets, const uint32_t* meshIndexCounts, size_t numMeshes, const vkx_material* materials,
                     size_t numMaterials, const vkx_instance* instances, size_t numInstances) {
    BIND(ctx);
    if ((numVertices && !vertices) || (numIndices && !indices) || (numMeshes && (!offsets || !meshIndexCounts)) || (numMaterials && !materials) || (numInstances && !instances))
        return vkx_fail(ctx, VKX_E_INVALID, "vkx_scene_upload: null array");
    if (numInstances >= (1u << 24)) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "more than 2^24 instances");
    uint32_t texturesUsed = 0;
    for (size_t m = 0; m < numMaterials; ++m) { // texture indices refer to the list of vkx_scene_textures
        const uint32_t t[4] = {materials[m].albedoTexture, materials[m].normalTexture, materials[m].metallicRoughnessTexture, materials[m].emissiveTexture};
        for (uint32_t i : t)
            if (i != VKX_INVALID_TEXTURE && i >= ctx->hTextures.size())
                return vkx_fail(ctx, VKX_E_INVALID, "material %zu uses texture %u but vkx_scene_textures provided %zu textures", m, i, ctx->hTextures.size());
        for (uint32_t i : t) if (i != VKX_INVALID_TEXTURE) texturesUsed = std::max(texturesUsed, i + 1u);
    }
    for (size_t m = 0; m < numMeshes; ++m) {
        if (meshIndexCounts[m] % 3 != 0) return vkx_fail(ctx, VKX_E_INVALID, "mesh %zu: index count %u is not a multiple of 3", m, meshIndexCounts[m]);
        if (size_t(offsets[m].indexOffset) + meshIndexCounts[m] > numIndices) return vkx_fail(ctx, VKX_E_INVALID, "mesh %zu: index range out of bounds", m);
        if (offsets[m].materialIndex >= numMaterials) return vkx_fail(ctx, VKX_E_INVALID, "mesh %zu: material %u out of range", m, offsets[m].materialIndex);
        if (offsets[m].vertexOffset > numVertices) return vkx_fail(ctx, VKX_E_INVALID, "mesh %zu: vertex offset out of range", m);
    }
    // index validation (a bad index would read out of bounds on the device)
    {
        for (size_t m = 0; m < numMeshes; ++m) {
            uint32_t limit = uint32_t(numVertices) - offsets[m].vertexOffset;
            for (uint32_t i = 0; i < meshIndexCounts[m]; ++i)
                if (indices[offsets[m].indexOffset + i] >= limit) return vkx_fail(ctx, VKX_E_INVALID, "mesh %zu: vertex index out of range", m);
        }
    }
    for (size_t v = 0; v < numVertices; ++v)
        for (int a = 0; a < 3; ++a) if (!std::isfinite(vertices[v].pos[a])) return vkx_fail(ctx, VKX_E_INVALID, "vertex %zu has a non-finite position", v);
    std::vector<float> w2o(numInstances * 9);
    ctx->hInstTriBase.assign(numInstances + 1, 0);
    size_t total = 0;
    for (size_t k = 0; k < numInstances; ++k) {
        if (instances[k].meshEntry >= numMeshes) return vkx_fail(ctx, VKX_E_INVALID, "instance %zu: mesh entry out of range", k);
        ctx->hInstTriBase[k] = uint32_t(total);
        total += meshIndexCounts[instances[k].meshEntry] / 3;
        if (total >= (1ull << 29)) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "more than 2^29 instanced triangles");
        // inverse of the 3x3 part (gl_WorldToObjectEXT), fp32 adjugate / determinant
        const float* M = instances[k].transform;
        float a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], i = M[10];
        float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
        float det = a * A + b * B + c * C;
        float id = 1.0f / det;
        float* W = &w2o[k * 9];
        W[0] = A * id; W[1] = -(b * i - c * h) * id; W[2] = (b * f - c * e) * id;
        W[3] = B * id; W[4] = (a * i - c * g) * id;  W[5] = -(a * f - c * d) * id;
        W[6] = C * id; W[7] = -(a * h - b * g) * id; W[8] = (a * e - b * d) * id;
    }
    ctx->hInstTriBase[numInstances] = uint32_t(total);
    TRY(upload(ctx, &ctx->dVertices, vertices, numVertices));
    TRY(upload(ctx, &ctx->dIndices, indices, numIndices));
    TRY(upload(ctx, &ctx->dOffsets, offsets, numMeshes));
    TRY(upload(ctx, &ctx->dMeshCounts, meshIndexCounts, numMeshes));
    TRY(upload(ctx, &ctx->dMaterials, materials, numMaterials));
    TRY(upload(ctx, &ctx->dInstances, instances, numInstances));
    TRY(upload(ctx, &ctx->dWorldToObject, w2o.data(), w2o.size()));
    TRY(upload(ctx, &ctx->dInstTriBase, ctx->hInstTriBase.data(), ctx->hInstTriBase.size()));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->numVertices = numVertices; ctx->numIndices = numIndices; ctx->numMeshes = numMeshes; ctx->numMaterials = numMaterials;
    ctx->numInstances = numInstances; ctx->numFlatTris = total; ctx->texturesUsed = texturesUsed;
    ctx->bvhBuilt = false; ctx->bvhTopology = false;
    return VKX_OK;
}

int vkx_instances_update(vkx_ctx* ctx, const vkx_instance* instances, size_t numInstances) {
    BIND(ctx);
    if (!instances || numInstances != ctx->numInstances) return vkx_fail(ctx, VKX_E_INVALID, "vkx_instances_update: expected %zu instances (same list as vkx_scene_upload)", ctx->numInstances);
    std::vector<vkx_instance> cur(numInstances);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(cur.data(), ctx->dInstances, numInstances * sizeof(vkx_instance), cudaMemcpyDeviceToHost));
    std::vector<float> w2o(numInstances * 9);
    for (size_t k = 0; k < numInstances; ++k) {
        if (instances[k].meshEntry != cur[k].meshEntry) return vkx_fail(ctx, VKX_E_INVALID, "vkx_instances_update: instance %zu changed its mesh (only transforms, masks and ids may change)", k);
        for (int q = 0; q < 12; ++q) if (!std::isfinite(instances[k].transform[q])) return vkx_fail(ctx, VKX_E_INVALID, "instance %zu has a non-finite transform", k);
        const float* M = instances[k].transform; // inverse of the 3x3 part, same arithmetic as vkx_scene_upload
        float a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], i = M[10];
        float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
        float det = a * A + b * B + c * C;
        float id = 1.0f / det;
        float* W = &w2o[k * 9];
        W[0] = A * id; W[1] = -(b * i - c * h) * id; W[2] = (b * f - c * e) * id;
        W[3] = B * id; W[4] = (a * i - c * g) * id;  W[5] = -(a * f - c * d) * id;
        W[6] = C * id; W[7] = -(a * h - b * g) * id; W[8] = (a * e - b * d) * id;
    }
    CUDA_TRY(ctx, cudaMemcpy(ctx->dInstances, instances, numInstances * sizeof(vkx_instance), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->dWorldToObject, w2o.data(), w2o.size() * 4, cudaMemcpyHostToDevice));
    ctx->bvhBuilt = false; // the flattened world-space BVH is stale: vkx_bvh_build rebuilds it (deterministic, same spec as the first build), vkx_bvh_refit re-fits it
    return VKX_OK;
}

int vkx_bvh_build(vkx_ctx* ctx) { BIND(ctx); return bvhBuildDevice(ctx); }

int vkx_bvh_refit(vkx_ctx* ctx) {
    BIND(ctx);
    if (!ctx->bvhTopology) return vkx_fail(ctx, VKX_E_INVALID, "vkx_bvh_refit: no built BVH for the uploaded scene (call vkx_bvh_build first)");
    return bvhRefitDevice(ctx);
}

int vkx_bvh_info_get(vkx_ctx* ctx, vkx_bvh_info* out) {
    if (!ctx || !out) return VKX_E_INVALID;
    if (!ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "BVH not built");
    *out = ctx->bvh;
    return VKX_OK;
}

int vkx_bvh_download(vkx_ctx* ctx, void* nodes, size_t nodesBytes, void* triangles, size_t trianglesBytes) {
    BIND(ctx);
    if (!ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "BVH not built");
    if (nodes) { size_t need = size_t(ctx->bvh.numNodes) * 80; if (nodesBytes < need) return vkx_fail(ctx, VKX_E_INVALID, "node buffer too small"); CUDA_TRY(ctx, cudaMemcpy(nodes, ctx->dNodes, need, cudaMemcpyDeviceToHost)); }
    if (triangles) { size_t need = size_t(ctx->bvh.numTriangles) * 48; if (trianglesBytes < need) return vkx_fail(ctx, VKX_E_INVALID, "triangle buffer too small"); if (need) CUDA_TRY(ctx, cudaMemcpy(triangles, ctx->dTris, need, cudaMemcpyDeviceToHost)); }
    return VKX_OK;
}

int vkx_trace(vkx_ctx* ctx, const float* origins, const float* directions, size_t n, float tmin, float tmax, uint32_t cullMask, int anyHit, vkx_hit* out) {
    BIND(ctx);
    if (!ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "BVH not built");
    if (n && (!origins || !directions || !out)) return vkx_fail(ctx, VKX_E_INVALID, "vkx_trace: null array");
    return traceHostRays(ctx, origins, directions, n, tmin, tmax, cullMask, anyHit, out, false);
}

int vkx_trace_alpha(vkx_ctx* ctx, const float* origins, const float* directions, size_t n, float tmin, float tmax, uint32_t cullMask, int anyHit, vkx_hit* out) {
    BIND(ctx);
    if (!ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "BVH not built");
    if (n && (!origins || !directions || !out)) return vkx_fail(ctx, VKX_E_INVALID, "vkx_trace_alpha: null array");
    return traceHostRays(ctx, origins, directions, n, tmin, tmax, cullMask, anyHit, out, true);
}

// ---------------------------------------------------------------------------------------------------- DDGI
} // extern "C"
__global__ void k_iota_list(uint32_t* p, uint32_t first, uint32_t n) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = first + i; }
extern "C" {
static int checkGrid(vkx_ctx* ctx, const vkx_grid_info* g) {
    if (!g) return vkx_fail(ctx, VKX_E_INVALID, "null grid");
    if (g->colorRes != 8 || g->depthRes != 16) return vkx_fail(ctx, VKX_E_INVALID, "colorRes/depthRes must be 8/16 (baked into the reference's shaders)");
    if (g->resolution[0] < 2 || g->resolution[1] < 2 || g->resolution[2] < 2) return vkx_fail(ctx, VKX_E_INVALID, "grid resolution must be >= 2 on every axis");
    if (g->raysPerProbe < 1 || g->raysPerProbe > VKX_MAX_RAYS_PER_PROBE) return vkx_fail(ctx, VKX_E_INVALID, "raysPerProbe must be in [1, %d]", VKX_MAX_RAYS_PER_PROBE);
    return VKX_OK;
}

static int allocProbeScratch(vkx_ctx* ctx) {
    // ray-level scratch for one chunk of probes
    void* old[] = {ctx->dRays, ctx->dHits, ctx->dShadowQueue, ctx->dShadowVis, ctx->dShadowFlags, ctx->dIrrUnpacked, ctx->dDepUnpacked, ctx->dMissQueue, ctx->dFrontQueue, ctx->dFrontKeys, ctx->dFrontKeysOut, ctx->dFrontQueueSorted, ctx->dOrigins};
    for (void* p : old) if (p) cudaFree(p);
    ctx->dOrigins = nullptr;
    ctx->dMissQueue = ctx->dFrontQueue = ctx->dFrontKeys = ctx->dFrontKeysOut = ctx->dFrontQueueSorted = nullptr;
    ctx->dRays = nullptr; ctx->dHits = nullptr; ctx->dShadowQueue = nullptr; ctx->dShadowVis = nullptr; ctx->dShadowFlags = nullptr; ctx->dIrrUnpacked = ctx->dDepUnpacked = nullptr;
    const size_t maxRays = size_t(ctx->chunkProbes) * VKX_MAX_RAYS_PER_PROBE;
    CUDA_TRY(ctx, cudaMalloc(&ctx->dRays, maxRays * sizeof(float4)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dOrigins, size_t(ctx->chunkProbes) * sizeof(float4)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dHits, maxRays * sizeof(vkx_hit)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dShadowQueue, maxRays * 2 * sizeof(float4)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dShadowVis, maxRays));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dMissQueue, maxRays * 4));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dFrontQueue, maxRays * 4));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dFrontQueue, 0, maxRays * 4, ctx->stream)); // the sort carries the unused tail (keys all-ones) along: defined values, once
    CUDA_TRY(ctx, cudaMalloc(&ctx->dFrontKeys, maxRays * 4));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dFrontKeysOut, maxRays * 4));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dFrontQueueSorted, maxRays * 4));
    if (ctx->debugBuffers) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->dShadowFlags, maxRays));
        CUDA_TRY(ctx, cudaMalloc(&ctx->dIrrUnpacked, size_t(ctx->probeCount) * 36 * 3 * 4));
        CUDA_TRY(ctx, cudaMalloc(&ctx->dDepUnpacked, size_t(ctx->probeCount) * 196 * 2 * 4));
    }
    return VKX_OK;
}

int vkx_probes_init(vkx_ctx* ctx, const vkx_grid_info* grid) {
    BIND(ctx);
    TRY(checkGrid(ctx, grid));
    TRY(flushCopyRequests(ctx));
    if (ctx->copyStream) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copyStream)); // a queued read-back still reads the atlases freed below
    freeProbes(ctx);
    ctx->grid = *grid;
    ctx->probeCount = uint32_t(grid->resolution[0]) * uint32_t(grid->resolution[1]) * uint32_t(grid->resolution[2]);
    ctx->irrW = 8u * uint32_t(grid->resolution[0] * grid->resolution[1]); ctx->irrH = 8u * uint32_t(grid->resolution[2]);
    ctx->depW = 16u * uint32_t(grid->resolution[0] * grid->resolution[1]); ctx->depH = 16u * uint32_t(grid->resolution[2]);
    const size_t irrBytes = size_t(ctx->irrW) * ctx->irrH * 4, depBytes = size_t(ctx->depW) * ctx->depH * 4, stBytes = size_t(ctx->probeCount) * 4;
    uint32_t** bufs[] = {&ctx->dIrrWork, &ctx->dIrrSampled, &ctx->dDepWork, &ctx->dDepSampled, &ctx->dStateWork, &ctx->dStateSampled};
    const size_t sizes[] = {irrBytes, irrBytes, depBytes, depBytes, stBytes, stBytes};
    for (int i = 0; i < 6; ++i) { CUDA_TRY(ctx, cudaMalloc(bufs[i], sizes[i])); CUDA_TRY(ctx, cudaMemsetAsync(*bufs[i], 0, sizes[i], ctx->stream)); }
    CUDA_TRY(ctx, cudaMalloc(&ctx->dIndicesList, stBytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dDirs, 512 * sizeof(float4)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dInvDirs, 512 * sizeof(float4)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dQueueCount, 32));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dPerm, VKX_MAX_RAYS_PER_PROBE * 4));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dOrder, stBytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dBlockedOrder, stBytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dPermList, stBytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dIota, stBytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->dCellHist, stBytes + 4));
    k_iota_list<<<divUp(ctx->probeCount, 256), 256, 0, ctx->stream>>>(ctx->dIota, 0, ctx->probeCount); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaMalloc(&ctx->dBlendW, size_t(VKX_MAX_RAYS_PER_PROBE + 1) * 288 * 4)); // + the row of weight sums
    CUDA_TRY(ctx, cudaMalloc(&ctx->dBlendImage, BTC_IMAGE_BYTES)); CUDA_TRY(ctx, cudaMemsetAsync(ctx->dBlendImage, 0, BTC_IMAGE_BYTES, ctx->stream));
    { // rank of every probe in 2x2x2-block order (scheduling only: which probes share a warp)
        const uint32_t rx = uint32_t(grid->resolution[0]), ry = uint32_t(grid->resolution[1]);
        const uint32_t nbx = (rx + 1) / 2, nby = (ry + 1) / 2;
        std::vector<std::pair<uint64_t, uint32_t>> keys(ctx->probeCount);
        for (uint32_t p = 0; p < ctx->probeCount; ++p) {
            const uint32_t ix = p % rx, iy = (p % (rx * ry)) / rx, iz = p / (rx * ry);
            const uint64_t block = (ix >> 1) + uint64_t(nbx) * ((iy >> 1) + uint64_t(nby) * (iz >> 1));
            keys[p] = {(block << 3) | ((iz & 1u) << 2) | ((iy & 1u) << 1) | (ix & 1u), p};
        }
        std::sort(keys.begin(), keys.end());
        ctx->hBlockRank.assign(ctx->probeCount, 0);
        std::vector<uint32_t> order(ctx->probeCount);
        for (uint32_t r = 0; r < ctx->probeCount; ++r) { ctx->hBlockRank[keys[r].second] = r; order[r] = keys[r].second; }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dBlockedOrder, order.data(), stBytes, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // One chunk holds at most 32768 probes (8.4 M rays: 128 MiB of ray records) so the ray-level scratch stays L2-sized
    // relative to the atlases; debug buffers force a single chunk.
    ctx->chunkProbes = ctx->debugBuffers ? ctx->probeCount : std::min<uint32_t>(ctx->probeCount, 32768u);
    TRY(allocProbeScratch(ctx));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->probesReady = true;
    ctx->lastCount = 0;
    return VKX_OK;
}

int vkx_probes_debug(vkx_ctx* ctx, int enable) {
    BIND(ctx);
    ctx->debugBuffers = enable != 0;
    if (ctx->probesReady) {
        ctx->chunkProbes = ctx->debugBuffers ? ctx->probeCount : std::min<uint32_t>(ctx->probeCount, 32768u);
        TRY(allocProbeScratch(ctx));
    }
    return VKX_OK;
}

// mat3(orientation) * sphericalFibonacci(i, n), i in [0, count): traceProbes.rgen:36, irradiance.glsl:52-64. Computed on the
// host in fp32 (glibc sinf/cosf) once per update: 256 directions shared by every probe, as the reference's rayDirection image.
static void rayDirections(const float R[16], uint32_t count, float n, float* out) {
    const float pi = 3.1415926538f;
    const float PHI = std::sqrt(5.0f) * 0.5f + 0.5f;
    for (uint32_t k = 0; k < count; ++k) {
        const float i = float(k);
        const float ab = i * (PHI - 1.0f);
        const float phi = 2.0f * pi * (ab - std::floor(ab));
        const float cosTheta = 1.0f - (2.0f * i + 1.0f) * (1.0f / n);
        const float sinTheta = std::sqrt(std::min(std::max(1.0f - cosTheta * cosTheta, 0.0f), 1.0f));
        const float x = std::cos(phi) * sinTheta, y = std::sin(phi) * sinTheta, z = cosTheta;
        // mat3(R) * v, column-major R
        out[3 * k + 0] = R[0] * x + R[4] * y + R[8] * z;
        out[3 * k + 1] = R[1] * x + R[5] * y + R[9] * z;
        out[3 * k + 2] = R[2] * x + R[6] * y + R[10] * z;
    }
}

int vkx_probes_classify(vkx_ctx* ctx, const float orientation[16]) {
    BIND(ctx);
    if (!ctx->probesReady || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_classify: probes or BVH not ready");
    if (!orientation) return vkx_fail(ctx, VKX_E_INVALID, "null orientation");
    TRY(settleReadBack(ctx));
    float dirs[512 * 3];
    rayDirections(orientation, 512, float(ctx->grid.raysPerProbe), dirs);
    return ddgiClassify(ctx, dirs);
}

static int uploadFrameInputs(vkx_ctx* ctx, const vkx_grid_info* grid, const float orientation[16]) {
    TRY(checkGrid(ctx, grid));
    if (grid->resolution[0] != ctx->grid.resolution[0] || grid->resolution[1] != ctx->grid.resolution[1] || grid->resolution[2] != ctx->grid.resolution[2])
        return vkx_fail(ctx, VKX_E_INVALID, "grid resolution changed; call vkx_probes_init again (reference quirk A.10.3)");
    ctx->grid = *grid; // updateUniforms
    const uint32_t N = grid->raysPerProbe;
    float dirs[VKX_MAX_RAYS_PER_PROBE * 3];
    rayDirections(orientation, N, float(N), dirs);
    // pinned staging, 4 slots in rotation so that frames can be queued without waiting for the previous one
    if (!ctx->hStage) { CUDA_TRY(ctx, cudaMallocHost(&ctx->hStage, 4 * sizeof(vkx_ctx::FrameStage))); for (auto& e : ctx->stageEvent) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
    const int slot = int(ctx->stageCursor++ & 3u);
    if (ctx->stageUsed[slot]) CUDA_TRY(ctx, cudaEventSynchronize(ctx->stageEvent[slot]));
    ctx->curSlot = slot; // uploadOrder stages the to-update list / slot order of this frame in the same slot
    vkx_ctx::FrameStage& fs = ctx->hStage[slot];
    float4* d4 = fs.dirs;
    for (uint32_t i = 0; i < N; ++i) d4[i] = make_float4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], 1.0f);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dDirs, d4, N * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    // direction order of the frame: Morton curve over the octahedral map, so 4 consecutive entries are neighbours on the sphere
    std::pair<uint32_t, uint32_t> keys[VKX_MAX_RAYS_PER_PROBE];
    for (uint32_t i = 0; i < N; ++i) {
        const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        const float l1 = std::fabs(x) + std::fabs(y) + std::fabs(z);
        float u = l1 > 0.f ? x / l1 : 0.f, v = l1 > 0.f ? y / l1 : 0.f;
        if (z < 0.f) { const float uu = (1.0f - std::fabs(v)) * (u >= 0.f ? 1.f : -1.f), vv = (1.0f - std::fabs(u)) * (v >= 0.f ? 1.f : -1.f); u = uu; v = vv; }
        uint32_t qx = uint32_t(std::min(std::max((u * 0.5f + 0.5f) * 65535.0f, 0.0f), 65535.0f)), qy = uint32_t(std::min(std::max((v * 0.5f + 0.5f) * 65535.0f, 0.0f), 65535.0f));
        auto spread = [](uint32_t a) { a = (a | (a << 8)) & 0x00FF00FFu; a = (a | (a << 4)) & 0x0F0F0F0Fu; a = (a | (a << 2)) & 0x33333333u; a = (a | (a << 1)) & 0x55555555u; return a; };
        keys[i] = {spread(qx) | (spread(qy) << 1), i};
    }
    std::sort(keys, keys + N);
    uint32_t* perm = fs.perm;
    for (uint32_t i = 0; i < N; ++i) perm[i] = keys[i].second;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dPerm, perm, N * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->stageEvent[slot], ctx->stream));
    ctx->stageUsed[slot] = true;
    return VKX_OK;
}

// position -> slot order for a host-provided to-update list: slots sorted by the 2x2x2-block rank of their probe (O(P))
static int uploadOrder(vkx_ctx* ctx, const uint32_t* probeIndices, uint32_t count, uint32_t listOffset, bool uploadList) {
    // pinned staging of the current frame slot (uploadFrameInputs picked it and waited for its previous use), so both copies are
    // truly asynchronous and the host can queue the next frame while this one runs
    if (!ctx->hListStage) CUDA_TRY(ctx, cudaMallocHost(&ctx->hListStage, size_t(8) * ctx->probeCount * 4));
    uint32_t* list = ctx->hListStage + size_t(ctx->curSlot) * 2 * ctx->probeCount + listOffset;
    uint32_t* order = list + ctx->probeCount;
    std::vector<uint32_t>& mark = ctx->hMark;
    mark.assign(ctx->probeCount, 0u);
    bool dup = false;
    for (uint32_t s = 0; s < count; ++s) { uint32_t& m = mark[ctx->hBlockRank[probeIndices[s]]]; if (m) dup = true; m = s + 1; }
    if (dup) { for (uint32_t s = 0; s < count; ++s) order[s] = s; }
    else { uint32_t n = 0; for (uint32_t r = 0; r < ctx->probeCount; ++r) if (mark[r]) order[n++] = mark[r] - 1; }
    if (count) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dOrder + listOffset, order, size_t(count) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (uploadList && count) {
        memcpy(list, probeIndices, size_t(count) * 4);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dIndicesList + listOffset, list, size_t(count) * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    // the slot's event is recorded by uploadFrameInputs before these copies; record it again so the slot is not reused too early
    CUDA_TRY(ctx, cudaEventRecord(ctx->stageEvent[ctx->curSlot], ctx->stream));
    return VKX_OK;
}

// A sharded full-volume update blends only this rank's z-slab, so its work atlases (the `previous` texels of the next blend) are
// stale everywhere else; the gathered *sampled* set is complete. Any update that may blend probes outside the slab (a host list,
// the unsharded full volume, a sharded list) first brings the work set up to date.
static int syncWorkAtlases(vkx_ctx* ctx) {
    if (!ctx->workStale) return VKX_OK;
    { int rc = waitGather(ctx); if (rc != VKX_OK) return rc; }
    { int rc = settleReadBack(ctx); if (rc != VKX_OK) return rc; } // a read-back of the rank's own slices still reads the work atlases rewritten below
    const size_t irrBytes = size_t(ctx->irrW) * ctx->irrH * 4, depBytes = size_t(ctx->depW) * ctx->depH * 4, stBytes = size_t(ctx->probeCount) * 4;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dIrrWork, ctx->dIrrSampled, irrBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dDepWork, ctx->dDepSampled, depBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dStateWork, ctx->dStateSampled, stBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->workStale = false;
    return VKX_OK;
}

int vkx_probes_update(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16], const uint32_t* probeIndices, uint32_t count, int sync) {
    BIND(ctx);
    TRY(syncWorkAtlases(ctx));
    if (!ctx->probesReady || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_update: probes or BVH not ready");
    if (!light || !orientation) return vkx_fail(ctx, VKX_E_INVALID, "null light/orientation");
    static const bool traceHost = getenv("VKX_TRACE_HOST") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    TRY(uploadFrameInputs(ctx, grid, orientation));
    const double t1 = now();
    if (probeIndices) {
        if (count > ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "more indices than probes");
        // an unchanged to-update list (every frame of a full-volume schedule) is already on the device together with its slot order
        if (ctx->hLastList.size() != count || (count && memcmp(ctx->hLastList.data(), probeIndices, size_t(count) * 4) != 0)) {
            for (uint32_t i = 0; i < count; ++i) if (probeIndices[i] >= ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "probe index %u out of range", probeIndices[i]);
            ctx->hLastList.clear();
            TRY(uploadOrder(ctx, probeIndices, count, 0, true));
            ctx->hLastList.assign(probeIndices, probeIndices + count);
        }
    } else {
        ctx->hLastList.clear();
        count = ctx->probeCount;
        k_iota_list<<<divUp(count, 256), 256, 0, ctx->stream>>>(ctx->dIndicesList, 0, count); LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dOrder, ctx->dBlockedOrder, size_t(count) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->shardOrderReady = false;
    const double t2 = now();
    TRY(ddgiUpdate(ctx, *light, nullptr, count, 0, false));
    const double t3 = now();
    if (traceHost) fprintf(stderr, "[vkx host] inputs %.3f ms, list %.3f ms, enqueue %.3f ms\n", t1 - t0, t2 - t1, t3 - t2);
    if (ctx->copyPending) { CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopyDone, 0)); ctx->copyPending = false; } // a queued read-back still reads the sampled atlases
    TRY(ddgiPublish(ctx, count));
    if (ctx->evPublished) CUDA_TRY(ctx, cudaEventRecord(ctx->evPublished, ctx->stream));
    ctx->shardedLast = false;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_probes_schedule(vkx_ctx* ctx, uint32_t probesPerUpdate, uint32_t* count) {
    BIND(ctx);
    if (!ctx->probesReady) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_schedule: probes not ready");
    TRY(waitGather(ctx)); // the states of a sharded update must have landed
    ctx->hLastList.clear();
    TRY(scheduleProbes(ctx, probesPerUpdate, count));
    ctx->schedValid = true; ctx->shardOrderReady = false;
    return VKX_OK;
}

int vkx_probes_scheduler_state(vkx_ctx* ctx, const uint32_t* set, uint32_t* get) {
    if (!ctx) return VKX_E_INVALID;
    if (set) { if (ctx->probeCount && set[1] >= ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "lastUpdateOffset out of range"); ctx->schedLoopIndex = set[0]; ctx->schedOffset = set[1]; ctx->schedValid = false; }
    if (get) { get[0] = ctx->schedLoopIndex; get[1] = ctx->schedOffset; }
    return VKX_OK;
}

int vkx_probes_scheduled_list(vkx_ctx* ctx, uint32_t* indices, uint32_t capacity, uint32_t* count) {
    BIND(ctx);
    if (!ctx->schedValid) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_scheduled_list: call vkx_probes_schedule first");
    if (count) *count = ctx->schedCount;
    if (indices) {
        if (capacity < ctx->schedCount) return vkx_fail(ctx, VKX_E_INVALID, "list capacity %u < %u", capacity, ctx->schedCount);
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->schedCount) CUDA_TRY(ctx, cudaMemcpy(indices, ctx->dIndicesList, size_t(ctx->schedCount) * 4, cudaMemcpyDeviceToHost));
    }
    return VKX_OK;
}

int vkx_probes_update_scheduled(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16], int sync) {
    BIND(ctx);
    if (!ctx->probesReady || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_update_scheduled: probes or BVH not ready");
    if (!light || !orientation) return vkx_fail(ctx, VKX_E_INVALID, "null light/orientation");
    if (!ctx->schedValid) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_update_scheduled: call vkx_probes_schedule first (one schedule per update)");
    ctx->schedValid = false; // the list is consumed: publish changes the states the next schedule reads
    TRY(syncWorkAtlases(ctx));
    TRY(uploadFrameInputs(ctx, grid, orientation));
    const uint32_t count = ctx->schedCount;
    if (count == 0) return VKX_OK;
    TRY(ddgiUpdate(ctx, *light, nullptr, count, 0, false));
    if (ctx->copyPending) { CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopyDone, 0)); ctx->copyPending = false; }
    TRY(ddgiPublish(ctx, count));
    if (ctx->evPublished) CUDA_TRY(ctx, cudaEventRecord(ctx->evPublished, ctx->stream));
    ctx->shardedLast = false;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return VKX_OK;
}

int vkx_probes_download(vkx_ctx* ctx, uint32_t* irradiance, uint32_t* depth, uint32_t* state, float* rays, size_t raysCapacityBytes) {
    BIND(ctx);
    TRY(flushCopyRequests(ctx));
    TRY(waitGather(ctx));
    if (!ctx->probesReady) return vkx_fail(ctx, VKX_E_INVALID, "probes not initialised");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (irradiance) CUDA_TRY(ctx, cudaMemcpy(irradiance, ctx->dIrrSampled, size_t(ctx->irrW) * ctx->irrH * 4, cudaMemcpyDeviceToHost));
    if (depth) CUDA_TRY(ctx, cudaMemcpy(depth, ctx->dDepSampled, size_t(ctx->depW) * ctx->depH * 4, cudaMemcpyDeviceToHost));
    if (state) CUDA_TRY(ctx, cudaMemcpy(state, ctx->dStateSampled, size_t(ctx->probeCount) * 4, cudaMemcpyDeviceToHost));
    if (rays) {
        if (ctx->lastCount > ctx->chunkProbes) return vkx_fail(ctx, VKX_E_INVALID, "ray buffer only holds one chunk; enable vkx_probes_debug before the update");
        size_t need = size_t(ctx->lastRays) * 16;
        if (raysCapacityBytes < need) return vkx_fail(ctx, VKX_E_INVALID, "ray buffer too small");
        if (need) CUDA_TRY(ctx, cudaMemcpy(rays, ctx->dRays, need, cudaMemcpyDeviceToHost));
    }
    return VKX_OK;
}

/* Asynchronous variant: queues the three device->host copies on a copy stream behind the last publish and returns; the next update's
 * publish waits for them. Host buffers should be pinned. vkx_probes_download_wait blocks until the copies have landed. */
int vkx_probes_download_async(vkx_ctx* ctx, uint32_t* irradiance, uint32_t* depth, uint32_t* state) {
    BIND(ctx);
    if (!ctx->probesReady) return vkx_fail(ctx, VKX_E_INVALID, "probes not initialised");
    return vkx_probes_download_slab_async(ctx, 0, uint32_t(ctx->grid.resolution[2]), irradiance, depth, state);
}
/* Same for the z-slices [z0, z1) only (contiguous atlas rows [8*z0, 8*z1) / [16*z0, 16*z1) and state words): what one rank of a sharded
 * run owns. Destination pointers address the first copied row. */
int vkx_probes_download_slab_async(vkx_ctx* ctx, uint32_t z0, uint32_t z1, uint32_t* irradiance, uint32_t* depth, uint32_t* state) {
    BIND(ctx);
    if (!ctx->probesReady || z0 >= z1 || z1 > uint32_t(ctx->grid.resolution[2])) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_download_slab_async: bad slab");
    // After a sharded update the rank's own slab is complete in its work atlases as soon as its blend has run: reading it from there
    // does not have to wait for the all-gather, which stays hidden behind the next frame's traversal. Any other slab comes from the
    // sampled atlases and needs the gather.
    bool own = false; // inside one of this rank's slice groups (vkx_shard_slices)?
    { uint32_t s = 0, K = 0;
      if (ctx->nranks > 1 && vkx_shard_groups(uint32_t(ctx->grid.resolution[2]), ctx->nranks, &s, &K) == VKX_OK) {
          const uint32_t g = z0 / (s * uint32_t(ctx->nranks)), lo = g * s * uint32_t(ctx->nranks) + uint32_t(ctx->rank) * s;
          own = z0 >= lo && z1 <= lo + s;
      } }
    const bool fromWork = ctx->shardedLast && (ctx->gatherPending || ctx->p2pPending) && own;
    if (!fromWork) TRY(waitGather(ctx));
    TRY(ensureCopyStream(ctx));
    const size_t plane = size_t(ctx->grid.resolution[0]) * size_t(ctx->grid.resolution[1]), nz = z1 - z0;
    const uint32_t* irrSrc = fromWork ? ctx->dIrrWork : ctx->dIrrSampled; const uint32_t* depSrc = fromWork ? ctx->dDepWork : ctx->dDepSampled;
    const uint32_t* stSrc = fromWork ? ctx->dStateWork : ctx->dStateSampled;
    const vkx_ctx::CopyOp ops[3] = {{irradiance, irrSrc + size_t(8 * z0) * ctx->irrW, size_t(8 * nz) * ctx->irrW * 4},
                                    {depth, depSrc + size_t(16 * z0) * ctx->depW, size_t(16 * nz) * ctx->depW * 4},
                                    {state, stSrc + size_t(z0) * plane, nz * plane * 4}};
    if (deferReadBack(ctx, fromWork)) { // noted; queued by the next update (or by whoever needs it ordered first)
        if (!ctx->copyRequests.empty() && ctx->copyRequestsReadWork != fromWork) TRY(flushCopyRequests(ctx));
        for (const auto& op : ops) if (op.dst) ctx->copyRequests.push_back(op);
        ctx->copyRequestsReadWork = fromWork;
        return VKX_OK;
    }
    TRY(flushCopyRequests(ctx));
    if (ctx->copyPending && ctx->copyReadsWork != fromWork) { // one flag says which atlas set the copies in flight read: do not mix the two kinds
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopyDone, 0)); ctx->copyPending = false;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->evPublished, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->evPublished, 0));
    for (const auto& op : ops) if (op.dst) CUDA_TRY(ctx, cudaMemcpyAsync(op.dst, op.src, op.bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->evCopyDone, ctx->copyStream));
    ctx->copyPending = true; ctx->copyReadsWork = fromWork;
    return VKX_OK;
}
int vkx_probes_download_wait(vkx_ctx* ctx) {
    BIND(ctx);
    TRY(flushCopyRequests(ctx));
    if (ctx->evCopyDone) CUDA_TRY(ctx, cudaEventSynchronize(ctx->evCopyDone));
    return VKX_OK;
}

int vkx_probes_upload(vkx_ctx* ctx, const uint32_t* irradiance, const uint32_t* depth, const uint32_t* state) {
    BIND(ctx);
    TRY(settleReadBack(ctx));
    TRY(waitGather(ctx));
    TRY(syncWorkAtlases(ctx)); // a partial upload (e.g. states only) must not leave the other arrays of the work set stale
    if (!ctx->probesReady) return vkx_fail(ctx, VKX_E_INVALID, "probes not initialised");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t irrBytes = size_t(ctx->irrW) * ctx->irrH * 4, depBytes = size_t(ctx->depW) * ctx->depH * 4, stBytes = size_t(ctx->probeCount) * 4;
    if (irradiance) { CUDA_TRY(ctx, cudaMemcpy(ctx->dIrrSampled, irradiance, irrBytes, cudaMemcpyHostToDevice)); CUDA_TRY(ctx, cudaMemcpy(ctx->dIrrWork, irradiance, irrBytes, cudaMemcpyHostToDevice)); }
    if (depth) { CUDA_TRY(ctx, cudaMemcpy(ctx->dDepSampled, depth, depBytes, cudaMemcpyHostToDevice)); CUDA_TRY(ctx, cudaMemcpy(ctx->dDepWork, depth, depBytes, cudaMemcpyHostToDevice)); }
    if (state) { CUDA_TRY(ctx, cudaMemcpy(ctx->dStateSampled, state, stBytes, cudaMemcpyHostToDevice)); CUDA_TRY(ctx, cudaMemcpy(ctx->dStateWork, state, stBytes, cudaMemcpyHostToDevice)); }
    return VKX_OK;
}

int vkx_probes_download_unpacked(vkx_ctx* ctx, float* irr, float* depth) {
    BIND(ctx);
    if (!ctx->probesReady || !ctx->debugBuffers || !ctx->dIrrUnpacked) return vkx_fail(ctx, VKX_E_INVALID, "enable vkx_probes_debug before the update");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (irr) CUDA_TRY(ctx, cudaMemcpy(irr, ctx->dIrrUnpacked, size_t(ctx->lastCount) * 36 * 3 * 4, cudaMemcpyDeviceToHost));
    if (depth) CUDA_TRY(ctx, cudaMemcpy(depth, ctx->dDepUnpacked, size_t(ctx->lastCount) * 196 * 2 * 4, cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_probes_download_hits(vkx_ctx* ctx, vkx_hit* hits, uint8_t* shadow) {
    BIND(ctx);
    if (!ctx->probesReady || !ctx->debugBuffers || !ctx->dShadowFlags) return vkx_fail(ctx, VKX_E_INVALID, "enable vkx_probes_debug before the update");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (hits) CUDA_TRY(ctx, cudaMemcpy(hits, ctx->dHits, size_t(ctx->lastRays) * sizeof(vkx_hit), cudaMemcpyDeviceToHost));
    if (shadow) CUDA_TRY(ctx, cudaMemcpy(shadow, ctx->dShadowFlags, size_t(ctx->lastRays), cudaMemcpyDeviceToHost));
    return VKX_OK;
}

int vkx_probes_timings(vkx_ctx* ctx, float ms[5]) {
    BIND(ctx);
    TRY(waitGather(ctx));
    if (!ms) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 5; ++i) ms[i] = 0.f;
    if (!ctx->lastCount) return VKX_OK;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[0], ctx->shardedLast ? ctx->ev[4] : ctx->ev[0], ctx->ev[3]));
    if (ctx->shardedLast) { // per-stage split is per chunk in the sharded path: set-up before the first chunk's traversal, that chunk's kernels, everything after its blend
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms[1], ctx->ev[4], ctx->kev[0]));
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms[2], ctx->kev[0], ctx->kev[4]));
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms[4], ctx->kev[4], ctx->ev[3]));
        return VKX_OK;
    }
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[1], ctx->ev[0], ctx->ev[1]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[2], ctx->ev[1], ctx->ev[2]));
    ms[3] = 0.f; // borders are written by the blend kernel
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms[4], ctx->ev[2], ctx->ev[3]));
    return VKX_OK;
}

int vkx_probes_kernel_timings(vkx_ctx* ctx, float ms[4], uint32_t* probes, uint32_t* shadowRays) {
    BIND(ctx);
    TRY(waitGather(ctx));
    if (!ms) return VKX_E_INVALID;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; ++i) ms[i] = 0.f;
    if (!ctx->lastCount) return VKX_OK;
    for (int i = 0; i < 4; ++i) CUDA_TRY(ctx, cudaEventElapsedTime(&ms[i], ctx->kev[i], ctx->kev[i + 1]));
    if (probes) *probes = ctx->kevProbes;
    if (shadowRays) CUDA_TRY(ctx, cudaMemcpy(shadowRays, ctx->dQueueCount, 4, cudaMemcpyDeviceToHost)); // queue length of the last chunk
    return VKX_OK;
}

int vkx_probes_device_ptrs(vkx_ctx* ctx, void** irradiance, void** depth, void** state) {
    if (!ctx || !ctx->probesReady) return VKX_E_INVALID;
    if (irradiance) *irradiance = ctx->dIrrSampled;
    if (depth) *depth = ctx->dDepSampled;
    if (state) *state = ctx->dStateSampled;
    return VKX_OK;
}

// ---------------------------------------------------------------------------------------------------- multi-GPU
int vkx_comm_unique_id(void* id128) {
    if (!id128) return VKX_E_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return VKX_E_NCCL;
    memcpy(id128, &id, 128);
    return VKX_OK;
}

int vkx_comm_init(vkx_ctx* ctx, int rank, int nranks, const void* id128) {
    BIND(ctx);
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return vkx_fail(ctx, VKX_E_INVALID, "bad rank/nranks");
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclComm_t comm;
    ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return vkx_fail(ctx, VKX_E_NCCL, "ncclCommInitRank: %s", ncclGetErrorString(r));
    ctx->comm = reinterpret_cast<ncclComm*>(comm); ctx->rank = rank; ctx->nranks = nranks; ctx->shardOrderReady = false;
    if (!ctx->commStream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->commStream, cudaStreamNonBlocking));
    if (!ctx->commEvent) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->commEvent, cudaEventDisableTiming));
    if (!ctx->gatherDone) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->gatherDone, cudaEventDisableTiming));
    return VKX_OK;
}

int vkx_comm_p2p_export(vkx_ctx* ctx, void* handle64) {
    BIND(ctx);
    if (!handle64) return VKX_E_INVALID;
    if (!ctx->probesReady) return vkx_fail(ctx, VKX_E_INVALID, "vkx_comm_p2p_export: call vkx_probes_init first");
    if (ctx->nranks < 2 || ctx->nranks > VKX_MAX_RANKS) return vkx_fail(ctx, VKX_E_INVALID, "vkx_comm_p2p_export: needs vkx_comm_init with 2..%d ranks", VKX_MAX_RANKS);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    TRY(waitGather(ctx));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!ctx->p2pSlab) {
        auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
        const size_t irrBytes = size_t(ctx->irrW) * ctx->irrH * 4, depBytes = size_t(ctx->depW) * ctx->depH * 4, stBytes = size_t(ctx->probeCount) * 4;
        ctx->p2pDepOff = al(irrBytes); ctx->p2pStOff = ctx->p2pDepOff + al(depBytes); ctx->p2pSetBytes = ctx->p2pStOff + al(stBytes);
        ctx->p2pFlagsOff = 2 * ctx->p2pSetBytes;
        char* slab = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&slab, ctx->p2pFlagsOff + 512));
        CUDA_TRY(ctx, cudaMemset(slab, 0, ctx->p2pFlagsOff + 512));
        CUDA_TRY(ctx, cudaMemcpy(slab, ctx->dIrrSampled, irrBytes, cudaMemcpyDeviceToDevice));
        CUDA_TRY(ctx, cudaMemcpy(slab + ctx->p2pDepOff, ctx->dDepSampled, depBytes, cudaMemcpyDeviceToDevice));
        CUDA_TRY(ctx, cudaMemcpy(slab + ctx->p2pStOff, ctx->dStateSampled, stBytes, cudaMemcpyDeviceToDevice));
        void* old[] = {ctx->dIrrSampled, ctx->dDepSampled, ctx->dStateSampled, ctx->dIrrNext, ctx->dDepNext, ctx->dStateNext};
        for (void* p : old) if (p) cudaFree(p);
        ctx->p2pSlab = slab; ctx->p2pSampledSet = 0; ctx->p2pFrame = 0;
        ctx->dIrrSampled = reinterpret_cast<uint32_t*>(slab); ctx->dDepSampled = reinterpret_cast<uint32_t*>(slab + ctx->p2pDepOff); ctx->dStateSampled = reinterpret_cast<uint32_t*>(slab + ctx->p2pStOff);
        char* s1 = slab + ctx->p2pSetBytes;
        ctx->dIrrNext = reinterpret_cast<uint32_t*>(s1); ctx->dDepNext = reinterpret_cast<uint32_t*>(s1 + ctx->p2pDepOff); ctx->dStateNext = reinterpret_cast<uint32_t*>(s1 + ctx->p2pStOff);
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->p2pSlab));
    memcpy(handle64, &h, 64);
    return VKX_OK;
}

int vkx_comm_p2p_import(vkx_ctx* ctx, const void* handles, int count) {
    BIND(ctx);
    if (!handles && count == 0) { TRY(waitGather(ctx)); ctx->p2p = false; return VKX_OK; } // back to the NCCL all-gather (every rank must do the same)
    if (!handles || count != ctx->nranks || !ctx->p2pSlab) return vkx_fail(ctx, VKX_E_INVALID, "vkx_comm_p2p_import: export first, then pass one 64-byte handle per rank");
    for (int r = 0; r < count; ++r) {
        if (r == ctx->rank) { ctx->peerSlab[r] = ctx->p2pSlab; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, static_cast<const char*>(handles) + size_t(r) * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return vkx_fail(ctx, VKX_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        ctx->peerSlab[r] = static_cast<char*>(p);
    }
    ctx->p2p = true; ctx->p2pCopy = false;
    return VKX_OK;
}

int vkx_comm_p2p_mode(vkx_ctx* ctx, int copyEngines) {
    BIND(ctx);
    if (!ctx->p2p) return vkx_fail(ctx, VKX_E_INVALID, "vkx_comm_p2p_mode: import the peer slabs first (vkx_comm_p2p_import)");
    TRY(waitGather(ctx));
    if (copyEngines && !ctx->hFlagRing) CUDA_TRY(ctx, cudaHostAlloc(&ctx->hFlagRing, 4096 * sizeof(uint32_t), cudaHostAllocPortable));
    ctx->p2pCopy = copyEngines != 0;
    return VKX_OK;
}

// Full-volume update, sharded: the z range is cut into K chunks of s*nranks slices; inside chunk k rank r traces and blends
// the s slices [k*s*n + r*s, k*s*n + (r+1)*s). A chunk's atlas rows are contiguous in memory, so one ncclAllGather per
// atlas per chunk (on a second stream, overlapped with the next chunk's tracing) assembles the *next* sampled atlases on
// every rank; they become current by pointer swap at the end. No reduction crosses ranks, so results equal the 1-GPU run.
int vkx_probes_update_sharded(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16], int sync) {
    BIND(ctx);
    if (!ctx->probesReady || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_update_sharded: probes or BVH not ready");
    if (!light || !orientation) return vkx_fail(ctx, VKX_E_INVALID, "null light/orientation");
    if (ctx->nranks == 1 || !ctx->comm) return vkx_probes_update(ctx, grid, light, orientation, nullptr, 0, sync);
    TRY(uploadFrameInputs(ctx, grid, orientation));
    ctx->hLastList.clear();
    const uint32_t n = uint32_t(ctx->nranks), rz = uint32_t(ctx->grid.resolution[2]), plane = uint32_t(ctx->grid.resolution[0] * ctx->grid.resolution[1]);
    if (rz % n != 0) return vkx_fail(ctx, VKX_E_INVALID, "grid z resolution %u is not divisible by %u ranks", rz, n);
    const size_t irrBytes = size_t(ctx->irrW) * ctx->irrH * 4, depBytes = size_t(ctx->depW) * ctx->depH * 4, stBytes = size_t(ctx->probeCount) * 4;
    if (!ctx->dIrrNext) { // (with peer exchange enabled the next set lives in the shared slab)
        CUDA_TRY(ctx, cudaMalloc(&ctx->dIrrNext, irrBytes)); CUDA_TRY(ctx, cudaMalloc(&ctx->dDepNext, depBytes)); CUDA_TRY(ctx, cudaMalloc(&ctx->dStateNext, stBytes)); }
    // One update per frame over all of the rank's slice groups: the all-gathers of frame f are not waited for at the end of the update
    // but before the first kernel of frame f+1 that reads the sampled atlases (k_shade_front), so they overlap frame f+1's primary
    // traversal, which never touches them. (Measured on 4 GPUs: one update per group, to overlap inside the frame, cost more in small
    // launches than it hid.)
    uint32_t s = 0, K = 0; // slices per group, groups per rank
    if (vkx_shard_groups(rz, ctx->nranks, &s, &K) != VKX_OK) return vkx_fail(ctx, VKX_E_INVALID, "vkx_shard_groups");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(ctx->comm);
    cudaStream_t st = ctx->stream, cs = ctx->commStream;
    const bool p2p = ctx->p2p, pushCopies = ctx->p2p && ctx->p2pCopy;
    if (p2p && !pushCopies) { // k_blend stores its tiles straight into every rank's next atlas set (NVLink peer memory): no separate exchange step
        const int nextSet = ctx->p2pSampledSet ^ 1;
        PeerTargets pt{}; pt.n = int(n);
        for (uint32_t r = 0; r < n; ++r) {
            char* base = ctx->peerSlab[r] + size_t(nextSet) * ctx->p2pSetBytes;
            pt.irr[r] = reinterpret_cast<uint32_t*>(base); pt.dep[r] = reinterpret_cast<uint32_t*>(base + ctx->p2pDepOff); pt.state[r] = reinterpret_cast<uint32_t*>(base + ctx->p2pStOff);
        }
        ctx->blendPeers = pt; ctx->blendToPeers = true;
    } else if (!p2p && ctx->copyPending && !ctx->copyReadsWork) { CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->commStream, ctx->evCopyDone, 0)); ctx->copyPending = false; } // read-back of the buffers about to be overwritten
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[4], st));
    const uint32_t groupProbes = s * plane, total = K * groupProbes;
    if (!ctx->shardOrderReady) { // the rank's to-update list (its slices of every group, in z order) and its slot order: the same every frame, built once
        for (uint32_t k = 0; k < K; ++k) {
            const uint32_t first = (k * s * n + uint32_t(ctx->rank) * s) * plane;
            k_iota_list<<<divUp(groupProbes, 256), 256, 0, st>>>(ctx->dIndicesList + size_t(k) * groupProbes, first, groupProbes); LAUNCH_CHECK(ctx);
        }
        std::vector<uint32_t> idx(total);
        for (uint32_t k = 0; k < K; ++k) { const uint32_t first = (k * s * n + uint32_t(ctx->rank) * s) * plane; for (uint32_t i = 0; i < groupProbes; ++i) idx[size_t(k) * groupProbes + i] = first + i; }
        TRY(uploadOrder(ctx, idx.data(), total, 0, false));
    }
    { int rc = ddgiUpdate(ctx, *light, nullptr, total, 0, false); if (rc != VKX_OK) { ctx->blendToPeers = false; return rc; } }
    if (!p2p) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->commEvent, st));
        CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->commEvent, 0));
        const size_t irrChunk = size_t(8 * s) * ctx->irrW, depChunk = size_t(16 * s) * ctx->depW, stChunk = size_t(s) * plane; // elements per rank and group
        static const bool dbgNoGather = getenv("VKX_DEBUG_NO_GATHER") != nullptr; // timing experiments only: the atlases of the other ranks stay stale
        ncclResult_t r = ncclGroupStart();
        for (uint32_t k = 0; k < K && r == ncclSuccess && !dbgNoGather; ++k) { // group k: the slices [k s n, (k + 1) s n) of all ranks are consecutive rows
            const size_t irrOff = size_t(8 * k * s * n) * ctx->irrW, depOff = size_t(16 * k * s * n) * ctx->depW, stOff = size_t(k * s * n) * plane;
            r = ncclAllGather(ctx->dIrrWork + irrOff + irrChunk * ctx->rank, ctx->dIrrNext + irrOff, irrChunk, ncclUint32, comm, cs);
            if (r == ncclSuccess) r = ncclAllGather(ctx->dDepWork + depOff + depChunk * ctx->rank, ctx->dDepNext + depOff, depChunk, ncclUint32, comm, cs);
            if (r == ncclSuccess) r = ncclAllGather(ctx->dStateWork + stOff + stChunk * ctx->rank, ctx->dStateNext + stOff, stChunk, ncclUint32, comm, cs);
        }
        if (r == ncclSuccess) r = ncclGroupEnd();
        if (r != ncclSuccess) return vkx_fail(ctx, VKX_E_NCCL, "ncclAllGather: %s", ncclGetErrorString(r));
    }
    ctx->shardOrderReady = true;
    if (pushCopies) {
        // Copy engines push this rank's rows of every group into the next set of every rank (its own included): one strided copy per
        // array and peer (width = the rank's rows of one group, one row of the copy per group). They are ordered after the blend and
        // run beside the next update's traversal without occupying an SM - the persistent traversal kernel fills every SM to the last
        // register, so a collective's thread blocks only became resident when it ended (2 GPUs, cfg4: 14.06 ms with the NCCL
        // all-gather, 13.54 ms with the exchange left out). The arrival flag follows as a 4-byte copy in the same stream; the flag
        // protocol is the one of the fused path (a rank writes set T only after every rank has raised the previous frame's flag).
        const int nextSet = ctx->p2pSampledSet ^ 1;
        CUDA_TRY(ctx, cudaEventRecord(ctx->commEvent, st));
        CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->commEvent, 0));
        const size_t irrW = size_t(8 * s) * ctx->irrW * 4, depW = size_t(16 * s) * ctx->depW * 4, stW = size_t(s) * plane * 4; // bytes of one rank's rows in one group
        const size_t own = size_t(ctx->rank);
        for (uint32_t r = 0; r < n; ++r) {
            char* base = ctx->peerSlab[r] + size_t(nextSet) * ctx->p2pSetBytes;
            CUDA_TRY(ctx, cudaMemcpy2DAsync(base + own * irrW, irrW * n, reinterpret_cast<const char*>(ctx->dIrrWork) + own * irrW, irrW * n, irrW, K, cudaMemcpyDeviceToDevice, cs));
            CUDA_TRY(ctx, cudaMemcpy2DAsync(base + ctx->p2pDepOff + own * depW, depW * n, reinterpret_cast<const char*>(ctx->dDepWork) + own * depW, depW * n, depW, K, cudaMemcpyDeviceToDevice, cs));
            CUDA_TRY(ctx, cudaMemcpy2DAsync(base + ctx->p2pStOff + own * stW, stW * n, reinterpret_cast<const char*>(ctx->dStateWork) + own * stW, stW * n, stW, K, cudaMemcpyDeviceToDevice, cs));
        }
        // peers may overwrite the set a queued read-back still reads as soon as they see this rank's flag: raise it after the copy
        if (ctx->copyPending && !ctx->copyReadsWork) { CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->evCopyDone, 0)); ctx->copyPending = false; }
        uint32_t* src = ctx->hFlagRing + (ctx->p2pFrame & 4095u);
        *src = ctx->p2pFrame + 1u; // the host is never 4096 frames ahead of the device
        for (uint32_t r = 0; r < n; ++r)
            CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<uint32_t*>(ctx->peerSlab[r] + ctx->p2pFlagsOff) + ctx->rank, src, 4, cudaMemcpyDefault, cs));
        // (the wait of the next reader - waitGather - polls this rank's own flag too: when it has passed, the pushes above have left
        // the work atlases, which the next blend rewrites)
        ctx->p2pFrame++; ctx->p2pSampledSet ^= 1; ctx->p2pPending = true;
    } else if (p2p) {
        ctx->blendToPeers = false;
        // peers may overwrite the set a queued read-back still reads as soon as they see this rank's flag: raise it after the copy
        if (ctx->copyPending && !ctx->copyReadsWork) { CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->evCopyDone, 0)); ctx->copyPending = false; }
        TRY(launchP2pSignal(ctx));
        ctx->p2pFrame++; ctx->p2pSampledSet ^= 1; ctx->p2pPending = true; // waited for by the next reader of the sampled atlases (waitGather)
    } else {
        CUDA_TRY(ctx, cudaEventRecord(ctx->gatherDone, cs));
        ctx->gatherPending = true; // waited for by the next reader of the sampled atlases (waitGather)
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    // publish = swap; the work buffers keep this rank's slices current (they are the only ones it reads as `previous`)
    std::swap(ctx->dIrrSampled, ctx->dIrrNext); std::swap(ctx->dDepSampled, ctx->dDepNext); std::swap(ctx->dStateSampled, ctx->dStateNext);
    ctx->workStale = true;
    ctx->lastCount = total; ctx->lastRays = total * ctx->grid.raysPerProbe; ctx->shardedLast = true;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return VKX_OK;
}


// A rank's z-slices in a sharded full-volume update. The volume is cut into groups of `groupSlices` consecutive slices per rank, dealt
// round robin: group g of rank r = slices [g s n + r s, g s n + (r + 1) s). With s = 2 (whole 2x2x2 probe blocks) every rank gets a
// sample of the whole volume instead of one slab, which evens the load (measured on cfg4, 8 ranks: slowest / mean slab 1.069,
// interleaved 1.036, tools/slab_balance.py), and the slices of all ranks in group g are still consecutive: one all-gather per group
// and atlas, no packing. An odd number of slices per rank falls back to one slab per rank.
int vkx_shard_groups(uint32_t rz, int nranks, uint32_t* groupSlices, uint32_t* numGroups) {
    if (nranks < 1 || !groupSlices || !numGroups || rz == 0u || rz % uint32_t(nranks) != 0u) return VKX_E_INVALID;
    const uint32_t S = rz / uint32_t(nranks);
    *groupSlices = (S % 2u == 0u) ? 2u : S;
    *numGroups = S / *groupSlices;
    return VKX_OK;
}
int vkx_shard_slices(uint32_t rz, int nranks, int rank, uint32_t group, uint32_t* z0, uint32_t* z1) {
    uint32_t s = 0, K = 0;
    if (rank < 0 || rank >= nranks || !z0 || !z1 || vkx_shard_groups(rz, nranks, &s, &K) != VKX_OK || group >= K) return VKX_E_INVALID;
    *z0 = group * s * uint32_t(nranks) + uint32_t(rank) * s; *z1 = *z0 + s;
    return VKX_OK;
}
int vkx_shard_range(uint32_t count, int nranks, int rank, uint32_t* first, uint32_t* n) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !first || !n) return VKX_E_INVALID;
    const uint32_t per = (count + uint32_t(nranks) - 1u) / uint32_t(nranks); // every rank's share is `per` list positions, the tail ranks get what is left
    const uint32_t f = std::min(count, uint32_t(rank) * per);
    *first = f; *n = std::min(per, count - f);
    return VKX_OK;
}
int vkx_stream_wait_exchange(vkx_ctx* ctx) { BIND(ctx); return waitGather(ctx); }

int vkx_probes_update_sharded_list(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16], const uint32_t* probeIndices, uint32_t count, int sync) {
    BIND(ctx);
    if (ctx->nranks == 1 || !ctx->comm) return vkx_probes_update(ctx, grid, light, orientation, probeIndices, count, sync);
    if (!ctx->probesReady || !ctx->bvhBuilt) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_update_sharded_list: probes or BVH not ready");
    if (!light || !orientation || (!probeIndices && count)) return vkx_fail(ctx, VKX_E_INVALID, "null light/orientation/list");
    if (ctx->p2p) return vkx_fail(ctx, VKX_E_INVALID, "vkx_probes_update_sharded_list uses the NCCL exchange; disable the peer-memory exchange first (vkx_comm_p2p_import(NULL, 0))");
    if (count > ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "more indices than probes");
    for (uint32_t i = 0; i < count; ++i) if (probeIndices[i] >= ctx->probeCount) return vkx_fail(ctx, VKX_E_INVALID, "probe index %u out of range", probeIndices[i]);
    TRY(waitGather(ctx)); // a pending full-volume exchange writes the sampled set this update patches
    TRY(syncWorkAtlases(ctx));
    TRY(uploadFrameInputs(ctx, grid, orientation));
    ctx->hLastList.clear(); ctx->shardOrderReady = false;
    cudaStream_t st = ctx->stream, cs = ctx->commStream;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[4], st));
    if (count == 0) { CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st)); ctx->lastCount = 0; ctx->lastRays = 0; ctx->shardedLast = true; return VKX_OK; }
    const uint32_t n = uint32_t(ctx->nranks), per = (count + n - 1u) / n;
    uint32_t first = 0, mine = 0;
    vkx_shard_range(count, ctx->nranks, ctx->rank, &first, &mine);
    const size_t need = size_t(per) * n;
    if (!ctx->dShardList) CUDA_TRY(ctx, cudaMalloc(&ctx->dShardList, size_t(ctx->probeCount) * 4));
    if (need > ctx->packCapacity) {
        CUDA_TRY(ctx, cudaStreamSynchronize(st)); CUDA_TRY(ctx, cudaStreamSynchronize(cs));
        if (ctx->dPackSend) cudaFree(ctx->dPackSend); if (ctx->dPackRecv) cudaFree(ctx->dPackRecv);
        ctx->dPackSend = ctx->dPackRecv = nullptr; ctx->packCapacity = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->dPackSend, size_t(per) * 81 * sizeof(uint4)));
        CUDA_TRY(ctx, cudaMalloc(&ctx->dPackRecv, need * 81 * sizeof(uint4)));
        ctx->packCapacity = need;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dShardList, probeIndices, size_t(count) * 4, cudaMemcpyHostToDevice, st)); // pageable source: returns after staging
    if (mine) {
        TRY(uploadOrder(ctx, probeIndices + first, mine, first, true));
        TRY(ddgiUpdate(ctx, *light, nullptr, mine, first, false));
    }
    TRY(ddgiPackTiles(ctx, ctx->dIndicesList + first, mine, ctx->dPackSend, st));
    if (ctx->copyPending) { CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->evCopyDone, 0)); ctx->copyPending = false; ctx->copyReadsWork = false; } // a queued read-back still reads the atlases about to be patched
    CUDA_TRY(ctx, cudaEventRecord(ctx->commEvent, st));
    CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->commEvent, 0));
    ncclResult_t r = ncclAllGather(ctx->dPackSend, ctx->dPackRecv, size_t(per) * 81 * 4, ncclUint32, reinterpret_cast<ncclComm_t>(ctx->comm), cs);
    if (r != ncclSuccess) return vkx_fail(ctx, VKX_E_NCCL, "ncclAllGather: %s", ncclGetErrorString(r));
    CUDA_TRY(ctx, cudaEventRecord(ctx->gatherDone, cs));
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->gatherDone, 0));
    // rank r's records sit at [r * per, r * per + its share): exactly list order, because every full share is `per` long
    TRY(ddgiUnpackTiles(ctx, ctx->dShardList, count, per, ctx->dPackRecv, st));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    if (ctx->evPublished) CUDA_TRY(ctx, cudaEventRecord(ctx->evPublished, st));
    ctx->lastCount = mine; ctx->lastRays = mine * ctx->grid.raysPerProbe; ctx->shardedLast = true;
    if (sync) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return VKX_OK;
}

} // extern "C"
