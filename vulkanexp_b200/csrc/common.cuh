// Shared declarations of libvkexp_b200: context, error handling, packing helpers.
// All .cu files are compiled with --fmad=false: every float expression is a sequence of single IEEE-rounded
// operations in source order, FMAs only where __fmaf_rn is written. This is what makes hit masks / triangle ids /
// BVH topology bit-identical to oracle/ (compiled with -ffp-contract=off); see DESIGN.md section 4.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/vkx.h"

static_assert(sizeof(vkx_vertex) == 64, "vkx_vertex");
static_assert(sizeof(vkx_material) == 48, "vkx_material");
static_assert(sizeof(vkx_offset_entry) == 12, "vkx_offset_entry");
static_assert(sizeof(vkx_instance) == 56, "vkx_instance");
static_assert(sizeof(vkx_grid_info) == 64, "vkx_grid_info");
static_assert(sizeof(vkx_light) == 32, "vkx_light");
static_assert(sizeof(vkx_camera) == 144, "vkx_camera");
static_assert(sizeof(vkx_hit) == 20, "vkx_hit");

struct ncclComm;

// One texture of the scene's list: all mip levels live in one texel arena (RGBA8 words), levelOffset in texels.
struct DeviceTexture { uint32_t width, height, levels, flags; uint32_t levelOffset[16]; };
#define VKX_MAX_TEXTURE_SIZE 16384u // 15 mip levels

struct DeviceScene {
    const vkx_vertex* vertices;
    const uint32_t* indices;
    const vkx_offset_entry* offsets;
    const vkx_material* materials;
    const vkx_instance* instances;
    const float* worldToObject; // 9 floats per instance, W[row][col] row-major = inverse of the 3x3 part
    const uint4* nodes;         // 5 x uint4 per node
    const float4* tris;         // 3 x float4 per triangle
    const uint32_t* texels;     // texel arena of every texture's mip chain, RGBA8 words (texture.cu)
    const float4* texelsDecoded; // the same texels decoded to linear fp32 (one 128-bit load per tap instead of a word + 4 table look-ups)
    const DeviceTexture* textures;
    const float* srgbLut;       // 512 entries: sRGB code -> linear, then code / 255 (texture.cuh::texDecode)
    uint32_t numTextures;
};

struct DeviceProbes {
    vkx_grid_info grid;
    uint32_t irrW, irrH, depW, depH, probeCount;
    const uint32_t* irrSampled;
    const uint32_t* depSampled;
    const uint32_t* stateSampled;
    uint32_t* irrWork;
    uint32_t* depWork;
    uint32_t* stateWork;
};

// Fused blend + atlas exchange over NVLink peer memory: where k_blend stores its finished tiles (every rank's *next* atlas set)
#define VKX_MAX_RANKS 16
struct PeerTargets { uint32_t* irr[VKX_MAX_RANKS]; uint32_t* dep[VKX_MAX_RANKS]; uint32_t* state[VKX_MAX_RANKS]; int n; };
struct PeerFlags { uint32_t* flags[VKX_MAX_RANKS]; };

struct vkx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int smCount = 148;
    bool blendAttrSet = false; int traceBlocksPerSm = 0; int poolBlocksPerSm[2] = {0, 0}; // per-device kernel attributes / occupancy (set on first use)

    // scene
    vkx_vertex* dVertices = nullptr; uint32_t* dIndices = nullptr; vkx_offset_entry* dOffsets = nullptr; uint32_t* dMeshCounts = nullptr;
    vkx_material* dMaterials = nullptr; vkx_instance* dInstances = nullptr; float* dWorldToObject = nullptr; uint32_t* dInstTriBase = nullptr;
    size_t numVertices = 0, numIndices = 0, numMeshes = 0, numMaterials = 0, numInstances = 0, numFlatTris = 0;
    std::vector<uint32_t> hInstTriBase;
    // textures (texture.cu): arena + descriptors + sRGB tables; hTextures mirrors the descriptors for validation / read-back
    uint32_t* dTexels = nullptr; float4* dTexelsDecoded = nullptr; DeviceTexture* dTextures = nullptr; float* dSrgbLut = nullptr; float* dSrgbThreshold = nullptr;
    std::vector<DeviceTexture> hTextures; size_t numTexels = 0; uint32_t texturesUsed = 0; // texturesUsed: highest texture index of the uploaded materials + 1

    // bvh
    uint4* dNodes = nullptr; float4* dTris = nullptr;
    vkx_bvh_info bvh{};
    bool bvhBuilt = false;
    bool bvhTopology = false;           // a built tree exists for the uploaded scene (vkx_bvh_refit can re-fit it after transforms / vertices changed)
    std::vector<uint32_t> hLevelBase;   // first node of every level of the wide tree + the node count (nodes are stored level by level)
    float* dRefitScratch = nullptr; size_t refitScratchBytes = 0; // triangle + node boxes of the refit

    // probes
    bool probesReady = false;
    vkx_grid_info grid{};
    uint32_t probeCount = 0, irrW = 0, irrH = 0, depW = 0, depH = 0;
    uint32_t *dIrrWork = nullptr, *dIrrSampled = nullptr, *dDepWork = nullptr, *dDepSampled = nullptr, *dStateWork = nullptr, *dStateSampled = nullptr;
    uint32_t* dIndicesList = nullptr;   // to-update list [probeCount]
    float4* dDirs = nullptr;            // rotated ray directions [512]
    float4* dInvDirs = nullptr;         // per direction: reciprocal components + ray octant, the loop-invariant half of makeRay() [512]
    float4* dOrigins = nullptr;         // per list slot of the current chunk: probe world position
    uint32_t* dPerm = nullptr;          // direction sort permutation of the frame [256]
    uint32_t* dOrder = nullptr;         // [probeCount] position -> slot of the to-update list (2x2x2 probe blocks)
    uint32_t* dBlockedOrder = nullptr;  // cached order for the full-volume list
    uint32_t* dPermList = nullptr;      // multi-chunk updates: probe indices in block order
    uint32_t* dIota = nullptr;          // 0..probeCount-1
    float* dBlendW = nullptr;           // per-frame blend weight table [256][288]
    bool binAttrSet = false;            // shared-memory opt-in of the counting-sort kernels (per device, like the other attributes)
    float* dBlendImage = nullptr; bool blendTcAttrSet = false; // the same weights as the tensor-core blend's A-operand image (blend_tc.cu)
    std::vector<uint32_t> hLastList;    // the host list currently resident in dIndicesList / dOrder (empty: none), so an unchanged list is not uploaded again
    std::vector<uint32_t> hBlockRank;   // probe linear index -> rank in 2x2x2-block order
    std::vector<uint32_t> hMark, hOrder; // scratch of uploadOrder
    struct FrameStage { float4 dirs[VKX_MAX_RAYS_PER_PROBE]; uint32_t perm[VKX_MAX_RAYS_PER_PROBE]; };
    FrameStage* hStage = nullptr; uint32_t* hListStage = nullptr; /* pinned [4][2][probeCount]: to-update list and order per slot */ int curSlot = 0; cudaEvent_t stageEvent[4] = {nullptr, nullptr, nullptr, nullptr}; bool stageUsed[4] = {false, false, false, false}; uint32_t stageCursor = 0;
    uint32_t chunkProbes = 0;           // probes traced per chunk
    float4* dRays = nullptr;            // [chunkProbes][N] (rgb, depth)
    vkx_hit* dHits = nullptr;           // [chunkProbes][N]
    float4* dShadowQueue = nullptr;     // [chunkProbes*N][2]
    uint32_t* dQueueCount = nullptr;    // 8 counters, see ddgiUpdate
    uint32_t* dMissQueue = nullptr; uint32_t* dFrontQueue = nullptr; // ray indices sorted by what they need next
    uint32_t *dFrontKeys = nullptr, *dFrontKeysOut = nullptr, *dFrontQueueSorted = nullptr, *dCellHist = nullptr; void* dSortTemp = nullptr; size_t sortTempBytes = 0; // front queue sorted by grid cell
    uint8_t* dShadowFlags = nullptr;    // debug
    uint8_t* dShadowVis = nullptr;      // per shadow-queue item: 1 lit, 2 occluded (ray-pool traversal + k_apply_shadow)
    float* dIrrUnpacked = nullptr; float* dDepUnpacked = nullptr; // debug, full count
    bool debugBuffers = false;
    uint32_t lastCount = 0, lastRays = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t kev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // per-kernel timing of chunk 0
    uint32_t kevProbes = 0, kevShadowRays = 0;

    cudaStream_t auxStream = nullptr; cudaEvent_t auxEvent[2] = {nullptr, nullptr}; // second stream of the update (sky kernel)

    // asynchronous read-back (vkx_probes_download_async)
    cudaStream_t copyStream = nullptr; cudaEvent_t evPublished = nullptr, evCopyDone = nullptr; bool copyPending = false, copyReadsWork = false; // copyReadsWork: the queued read-back reads the work atlases (own slab of a sharded update)
    // Read-backs of the sampled atlases that were requested but not queued yet (single-GPU contexts): they are queued right before the
    // next update's primary traversal starts instead of right behind the publish, see flushCopyRequests (api.cu)
    struct CopyOp { void* dst; const void* src; size_t bytes; };
    std::vector<CopyOp> copyRequests; bool copyRequestsReadWork = false; // all noted requests read the same atlas set (sampled, or the work set: own slices of a sharded update)

    // on-device scheduler (schedule.cu); the two counters are the reference's s_LoopIndex / _lastUpdateOffset
    uint32_t *dSchedFlags = nullptr, *dSchedPos = nullptr, *dSchedSlotOf = nullptr, *dSchedResult = nullptr, *hSchedResult = nullptr; void* dSchedTemp = nullptr; size_t schedTempBytes = 0;
    uint32_t schedLoopIndex = 0, schedOffset = 0, schedCount = 0; bool schedValid = false;

    // multi-GPU
    ncclComm* comm = nullptr; int rank = 0, nranks = 1;
    cudaStream_t commStream = nullptr; cudaEvent_t commEvent = nullptr, gatherDone = nullptr; bool gatherPending = false;
    uint32_t *dIrrNext = nullptr, *dDepNext = nullptr, *dStateNext = nullptr; // all-gather targets (sharded update)
    uint32_t* dShardList = nullptr; uint4 *dPackSend = nullptr, *dPackRecv = nullptr; size_t packCapacity = 0; // sharded list update: whole list, packed tiles (81 uint4 per probe)
    // peer-memory exchange (vkx_comm_p2p_export / _import): one slab per rank = atlas set 0 | atlas set 1 | arrival flags | error word,
    // mapped into every peer through CUDA IPC; sampled / next point into the slab
    bool p2p = false, p2pPending = false, blendToPeers = false;
    bool p2pCopy = false;          // peer exchange by copy engines (vkx_comm_p2p_mode): the blend writes local rows, DMA copies push them to the peers
    uint32_t* hFlagRing = nullptr; // pinned ring of frame numbers: the source of the arrival-flag copies
    char* p2pSlab = nullptr; char* peerSlab[VKX_MAX_RANKS] = {};
    size_t p2pSetBytes = 0, p2pDepOff = 0, p2pStOff = 0, p2pFlagsOff = 0;
    uint32_t p2pFrame = 0; int p2pSampledSet = 0;
    PeerTargets blendPeers = {};
    bool shardedLast = false, shardOrderReady = false;
    bool workStale = false; // after a sharded full-volume update the work atlases are current only inside this rank's slab (see syncWorkAtlases)

    // shadows
    float* dNoise = nullptr; uint32_t noiseW = 0, noiseH = 0, noiseSlices = 0;
    uint32_t shW = 0, shH = 0;
    float4 *dPosDepth = nullptr, *dNormalMetal = nullptr, *dShRaw = nullptr, *dShX = nullptr, *dShFinal[2] = {nullptr, nullptr};
    int shCur = 0; // dShFinal[shCur] = last frame's filtered result
    float4* dShDirs = nullptr; uint8_t* dShMask = nullptr; // debug
    cudaEvent_t sev[4] = {nullptr, nullptr, nullptr, nullptr};
    // composite (FinalGather): the two G-buffer targets only it reads, optional reflection input, output
    float4 *dAlbedoRough = nullptr, *dEmissive = nullptr, *dReflection = nullptr, *dGathered = nullptr;
    cudaEvent_t gev[2] = {nullptr, nullptr};
    // reflection pass (reflection.cu): raw 1-spp, after filter X, final ping-pong (dReflFinal[reflCur] = last frame's result), parity side buffers
    float4 *dReflRaw = nullptr, *dReflX = nullptr, *dReflFinal[2] = {nullptr, nullptr}, *dReflDirs = nullptr; vkx_hit* dReflHits = nullptr; uint8_t* dReflMask = nullptr; uint32_t *dReflQueue = nullptr, *dReflCount = nullptr;
    int reflCur = 0; bool reflValid = false;
    cudaEvent_t rev[4] = {nullptr, nullptr, nullptr, nullptr};
};

int vkx_fail(vkx_ctx* ctx, int code, const char* fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                                  \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return vkx_fail((ctx), VKX_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                                                    \
    do {                                                                                                     \
        (ctx)->launches++;                                                                                   \
        cudaError_t _e = cudaGetLastError();                                                                 \
        if (_e != cudaSuccess) return vkx_fail((ctx), VKX_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)

static inline unsigned divUp(size_t a, size_t b) { return unsigned((a + b - 1) / b); }

// ---- implemented in bvh_build.cu
int bvhBuildDevice(vkx_ctx* ctx);
int bvhRefitDevice(vkx_ctx* ctx);
int waitGather(vkx_ctx* ctx); // api.cu
int flushCopyRequests(vkx_ctx* ctx); // api.cu
int launchP2pWait(vkx_ctx* ctx);   // ddgi.cu: stream-ordered wait for every rank's tiles of the last sharded update
int launchP2pSignal(vkx_ctx* ctx); // ddgi.cu
// ---- ddgi.cu
int ddgiClassify(vkx_ctx* ctx, const float* dirs512);
int ddgiUpdate(vkx_ctx* ctx, const vkx_light& light, const uint32_t* hostIndices, uint32_t count, uint32_t firstProbe, bool publishAll);
int ddgiPublish(vkx_ctx* ctx, uint32_t count);
int ddgiPackTiles(vkx_ctx* ctx, const uint32_t* probeIndices, uint32_t count, uint4* packed, cudaStream_t st);                      // work atlases -> packed records
int ddgiUnpackTiles(vkx_ctx* ctx, const uint32_t* probeIndices, uint32_t count, uint32_t perRank, const uint4* packed, cudaStream_t st); // records of every rank -> work + sampled atlases
// ---- trace_api.cu (alphaTest: run anyhit.rahit's cut-out test on the candidates)
int traceHostRays(vkx_ctx* ctx, const float* origins, const float* directions, size_t n, float tmin, float tmax, uint32_t cullMask, int anyHit, vkx_hit* out, bool alphaTest);
void freeTextures(vkx_ctx* ctx); // texture.cu
// ---- shadow.cu
int shadowGBuffer(vkx_ctx* ctx, const vkx_camera& cam);
int shadowFrame(vkx_ctx* ctx, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light);
int reflectionFrame(vkx_ctx* ctx, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light); // reflection.cu

DeviceScene deviceScene(const vkx_ctx* ctx);
DeviceProbes deviceProbes(const vkx_ctx* ctx);
int scheduleProbes(vkx_ctx* ctx, uint32_t probesPerUpdate, uint32_t* countOut); // schedule.cu
int finalGather(vkx_ctx* ctx, const vkx_camera& cam, const vkx_light& light, bool haveReflection); // gather.cu

// ---------------------------------------------------------------------------------------------------------------
// device helpers
#ifdef __CUDACC__
#include <cuda_fp16.h>

// Total order on floats as unsigned keys (-0 < +0): min/max through atomicMin/atomicMax on the key.
__host__ __device__ inline uint32_t okey(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float unkey(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// ---- atlas texel formats (integer code identical to oracle/packing.h)
template <int MB>
__device__ inline uint32_t packUF(float f) {
    const uint32_t maxCode = (31u << MB) - 1u;
    uint32_t b = __float_as_uint(f);
    // Fast path, 2^-14 <= f < 2^16 (a normal number of the target format, or overflow): re-bias the exponent, round to nearest even
    // by adding half - 1 + lsb below the kept bits (a carry out of the mantissa increments the exponent, as it should), clamp.
    // Same code as the general path below for every such f.
    if (b - (113u << 23) < (30u << 23)) {
        const uint32_t t = b - (112u << 23) + ((1u << (22 - MB)) - 1u) + ((b >> (23 - MB)) & 1u);
        return min(t >> (23 - MB), maxCode);
    }
    if (!(f > 0.0f)) return 0;
    int e = int(b >> 23) - 127;
    uint32_t m = b & 0x7FFFFFu;
    if (e > 15) return maxCode;
    uint32_t code;
    if (e >= -14) {
        const int sh = 23 - MB;
        uint32_t q = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
        code = (uint32_t(e + 15) << MB) + q;
        if (rem > half || (rem == half && (q & 1u))) code += 1;
    } else {
        int sh = (23 - MB) + (-14 - e);
        if (sh > 24) return 0;
        uint32_t full = m | 0x800000u;
        uint32_t q = full >> sh, rem = full & ((1u << sh) - 1u), half = 1u << (sh - 1);
        code = q;
        if (rem > half || (rem == half && (q & 1u))) code += 1;
    }
    return code > maxCode ? maxCode : code;
}
template <int MB>
__device__ inline float unpackUF(uint32_t c) {
    uint32_t e = c >> MB, m = c & ((1u << MB) - 1u);
    if (e == 0) return float(m) * __uint_as_float(uint32_t(127 - 14 - MB) << 23);
    if (e == 31) return m ? __uint_as_float(0x7FC00000u) : __uint_as_float(0x7F800000u);
    return __uint_as_float(((e + 112u) << 23) | (m << (23 - MB)));
}
__device__ inline uint32_t packR11G11B10(float r, float g, float b) { return packUF<6>(r) | (packUF<6>(g) << 11) | (packUF<5>(b) << 22); }
// Decode through binary16: an unsigned 5e6m / 5e5m float is a half with the low mantissa bits zero (exact, including denormals,
// inf and NaN), so one shift + one F2F per channel replaces the branchy integer decode. Same values as unpackUF<>.
__device__ inline float3 unpackR11G11B10(uint32_t p) {
    return make_float3(__half2float(__ushort_as_half((unsigned short)((p & 0x7FFu) << 4))), __half2float(__ushort_as_half((unsigned short)(((p >> 11) & 0x7FFu) << 4))),
                       __half2float(__ushort_as_half((unsigned short)((p >> 22) << 5))));
}
__device__ inline uint32_t packRG16F(float r, float g) {
    return uint32_t(__half_as_ushort(__float2half_rn(r))) | (uint32_t(__half_as_ushort(__float2half_rn(g))) << 16);
}
__device__ inline float2 unpackRG16F(uint32_t p) {
    return __half22float2(*reinterpret_cast<const __half2*>(&p));
}

#endif // __CUDACC__
