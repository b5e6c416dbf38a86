/*
 * vkx.h — C ABI of libvkexp_b200.so: the B200-native DDGI probe update and 1-spp sun-shadow pass.
 *
 * This is the drop-in boundary for the hot path of Senryoku/VulkanExp. The reference has no FFI layer: its
 * Editor calls two C++ classes directly (reference src/IrradianceProbes.hpp:16-129, src/Renderer.hpp:39-86) and
 * the shaders see one shared descriptor ABI (reference src/RaytracingDescriptors.hpp:8-77). Every entry point
 * below names the reference interface it replaces. All signatures are plain pointers and sizes; no C++ / torch
 * types. Every function returns 0 on success and a negative VKX_E_* code on failure (the reference throws from
 * VK_CHECK, src/vulkan/VkTools.hpp:39-47; the C++ facade in vulkanexp_b200/csrc/host re-throws these codes).
 *
 * There is no CPU fallback behind this ABI. If no CUDA device is usable vkx_create fails.
 *
 * POD layouts are byte-identical to the reference's (sizes are static_assert'ed in the implementation):
 *   vkx_vertex        64 B  src/vulkan/Vertex.hpp:8-15
 *   vkx_material      48 B  src/vulkan/Material.hpp:16-25
 *   vkx_offset_entry  12 B  src/Renderer.hpp:21-25
 *   vkx_grid_info     64 B  src/IrradianceProbes.hpp:49-60 (GLSL: src/shaders/ProbeGrid.glsl:4-15)
 *   vkx_light         32 B  src/Light.hpp:6-9
 *   vkx_camera       144 B  src/Editor.hpp:58-63
 */
#ifndef VKX_H
#define VKX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKX_ABI_VERSION 1

/* error codes */
#define VKX_OK 0
#define VKX_E_INVALID (-1)  /* bad argument / call order */
#define VKX_E_CUDA (-2)     /* a CUDA runtime call failed; vkx_last_error has the string */
#define VKX_E_NOMEM (-3)
#define VKX_E_UNSUPPORTED (-4) /* e.g. a texture larger than 16384 texels on a side */
#define VKX_E_NCCL (-5)

#define VKX_MAX_RAYS_PER_PROBE 256 /* IrradianceProbes::MaxRaysPerProbe, src/IrradianceProbes.hpp:42 */
#define VKX_INVALID_TEXTURE 0xFFFFFFFFu

/* instance masks, src/shaders/InstanceMasks.glsl:1-3 */
#define VKX_INSTANCE_STATIC 0x01u
#define VKX_INSTANCE_DYNAMIC 0x02u
#define VKX_INSTANCE_SKINNED 0x04u

typedef struct vkx_vertex {
    float pos[3];
    float color[3];
    float normal[3];
    float tangent[4];
    float texCoord[2];
    uint32_t padding;
} vkx_vertex;

typedef struct vkx_material {
    float metallicFactor;
    float roughnessFactor;
    float baseColorFactor[3];
    float emissiveFactor[3];
    uint32_t albedoTexture;
    uint32_t normalTexture;
    uint32_t metallicRoughnessTexture;
    uint32_t emissiveTexture;
} vkx_material;

typedef struct vkx_offset_entry {
    uint32_t materialIndex; /* the mesh's defaultMaterialIndex, src/Renderer.cpp:110-114 */
    uint32_t vertexOffset;  /* in vertices */
    uint32_t indexOffset;   /* in indices */
} vkx_offset_entry;

/* One texture of Scene's texture list: the decoded image (what stb_image hands to Image::upload, src/vulkan/Image.cpp:62-111), its
 * VkFormat (albedo / emissive textures are R8G8B8A8_SRGB, normal / metallic-roughness maps R8G8B8A8_UNORM: src/Scene.cpp:43,671,677)
 * and the glTF sampler description stored in the .scene file (src/Resources.cpp:88-93; 0 = the reference's default, 9729 / 10497). */
typedef struct vkx_texture {
    const uint8_t* pixels; /* width * height RGBA8 texels, row-major, tightly packed */
    uint32_t width, height;
    uint32_t srgb;         /* 1: VK_FORMAT_R8G8B8A8_SRGB, 0: VK_FORMAT_R8G8B8A8_UNORM */
    uint32_t magFilter;    /* glTF: 9728 NEAREST, 9729 LINEAR */
    uint32_t minFilter;    /* glTF: 9728, 9729, 9984..9987; mapped like glTFToVkFilter / glTFToVkSamplerMipmapMode (src/Resources.cpp:8-32) */
    uint32_t wrapS, wrapT; /* glTF: 10497 REPEAT, 33071 CLAMP_TO_EDGE, 33648 MIRRORED_REPEAT (src/Resources.cpp:34-43) */
} vkx_texture;

/* One TLAS instance (VkAccelerationStructureInstanceKHR as filled by Renderer::createTLAS, src/Renderer.cpp:532-551). */
typedef struct vkx_instance {
    float transform[12];  /* row-major 3x4 object-to-world (VkTransformMatrixKHR) */
    uint32_t meshEntry;   /* instanceCustomIndex = index into the offset table */
    uint32_t mask;        /* VKX_INSTANCE_* */
} vkx_instance;

typedef struct vkx_grid_info {
    float extentMin[3];
    float depthSharpness;
    float extentMax[3];
    float hysteresis;
    int32_t resolution[3];
    uint32_t raysPerProbe;
    uint32_t colorRes; /* must be 8 */
    uint32_t depthRes; /* must be 16 */
    float shadowBias;
    uint32_t padding;
} vkx_grid_info;

typedef struct vkx_light {
    float direction[4];
    float color[4];
} vkx_light;

typedef struct vkx_camera {
    float view[16]; /* column-major, as glm::mat4 */
    float proj[16];
    float origin[3];
    uint32_t frameIndex;
} vkx_camera;

/* Closest-hit record of the traversal kernels (parity primitive; the reference's RT hardware exposes the same
 * quantities as gl_HitTEXT, gl_InstanceID/gl_PrimitiveID, hitAttributeEXT and gl_HitKindEXT). */
typedef struct vkx_hit {
    float t;           /* < 0: miss */
    uint32_t instance; /* index into the uploaded instance list */
    uint32_t primitive;/* triangle index inside the instance's mesh; bit 31 set: back-facing hit */
    float u, v;        /* barycentrics of vertices 1 and 2 */
} vkx_hit;

typedef struct vkx_bvh_info {
    uint32_t numNodes;      /* 8-wide compressed nodes, 80 B each */
    uint32_t numTriangles;  /* flattened world-space triangles, 48 B each */
    uint32_t numBinaryNodes;/* inner nodes of the intermediate binary SAH tree */
    uint32_t depth;         /* depth of the wide tree (root = 1) */
    float sceneMin[3];
    float sceneMax[3];
    float sahCost;          /* sum over wide nodes of area(node)/area(root) (diagnostic) */
    float buildMs;          /* device time of the last build */
} vkx_bvh_info;

typedef struct vkx_ctx vkx_ctx;

/* ---- context ---------------------------------------------------------------------------------------------- */
/* Replaces Device creation + IrradianceProbes::init's device argument (src/IrradianceProbes.cpp:12-13). */
int vkx_create(int device, vkx_ctx** out);
void vkx_destroy(vkx_ctx* ctx); /* IrradianceProbes::destroy, src/IrradianceProbes.cpp:596-624 */
const char* vkx_last_error(vkx_ctx* ctx); /* ctx may be NULL: error of the last failed vkx_create on this thread */
int vkx_abi_version(void);

/* ---- geometry: Renderer::allocateMeshes + createAccelerationStructures + createTLAS ------------------------ */
/* Uploads the mesh arenas, the offset table (src/Renderer.cpp:97-126), the material SSBO and the instance list
 * (one per MeshRendererComponent, already sorted as Renderer::sortRenderers does, src/Renderer.cpp:512-523).
 * meshIndexCounts[m] = number of indices of mesh entry m (the reference keeps it in VkAccelerationStructure
 * BuildRangeInfo.primitiveCount, src/Renderer.cpp:294-300). Texture indices of the materials refer to the list given to
 * vkx_scene_textures (call it first; without it every material must be untextured). */
int vkx_scene_upload(vkx_ctx* ctx, const vkx_vertex* vertices, size_t numVertices, const uint32_t* indices,
                     size_t numIndices, const vkx_offset_entry* offsets, const uint32_t* meshIndexCounts,
                     size_t numMeshes, const vkx_material* materials, size_t numMaterials,
                     const vkx_instance* instances, size_t numInstances);

/* uploadTextures (src/Resources.cpp:46-95): copies the images to the device and generates their mip chains (Image::generateMipmaps,
 * src/vulkan/Image.cpp:195-275: floor(log2(max(w, h))) + 1 levels, each a LINEAR vkCmdBlitImage of the previous one). The images may
 * be freed after the call. Sampling follows "sampler spec v1" (DESIGN.md section 9f): the Vulkan texel-filtering equations with exact
 * fp32 weights, isotropic level of detail (the reference enables anisotropic filtering, whose footprint is implementation-defined).
 * Consumers: closest-hit shading of probe and reflection rays (textureGrad with the ray differentials of texDerivative,
 * src/shaders/closesthit.glsl:50-107,161-192) and the alpha cut-out of anyhit.rahit in the sun-shadow and reflection pipelines.
 * numTextures = 0 removes the list. Must precede the vkx_scene_upload whose materials use the textures; replacing the list by one
 * shorter than the uploaded materials need discards the uploaded scene (tracing fails until the next vkx_scene_upload + vkx_bvh_build). */
int vkx_scene_textures(vkx_ctx* ctx, const vkx_texture* textures, size_t numTextures);
/* Host-side image decoder (no context, no GPU): what STBImage does in uploadTextures (src/Resources.cpp:56-60), for the formats the
 * library reads itself: PNG (every colour type and bit depth incl. palettes and tRNS keys, non-interlaced, at most 16384 x 16384),
 * Netpbm P6 and P7. Always expands to RGBA8. rgba may be NULL to query the size. Returns VKX_E_UNSUPPORTED for files it cannot
 * decode (never throws across the boundary; VKX_E_NOMEM if the host allocation fails). */
int vkx_image_decode(const char* path, uint8_t* rgba, size_t rgbaBytes, uint32_t* width, uint32_t* height);
/* Parity primitives for the texture path: the generated mip chain (level-major RGBA8, numLevels = floor(log2(max(w, h))) + 1) and
 * n texture look-ups with explicit gradients (uv: 2 floats, grads: dudx, dvdx, dudy, dvdy per look-up; grads NULL = texture() of a
 * ray-tracing stage = base level) -> out: 4 floats per look-up. */
int vkx_texture_download(vkx_ctx* ctx, uint32_t texture, void* texels, size_t texelsBytes, uint32_t* numLevels);
int vkx_texture_sample(vkx_ctx* ctx, uint32_t texture, const float* uv, const float* grads, size_t n, float* out);

/* Dynamic instances: Renderer::updateAccelerationStructureInstances + updateTLAS (src/Renderer.cpp:671-742, called from
 * onHierarchicalChanges). New transforms / masks / ids for the instance list of vkx_scene_upload (same count, same meshes). The
 * reference refits its TLAS in place; here all instanced geometry lives in one world-space BVH, so the structure is marked stale
 * and the next vkx_bvh_build rebuilds it (the same deterministic build: the result equals a fresh upload with these transforms) or
 * vkx_bvh_refit re-fits it in place like the reference's TLAS update. */
int vkx_instances_update(vkx_ctx* ctx, const vkx_instance* instances, size_t numInstances);
/* Skinned meshes: vertexSkinning.comp (src/shaders/vertexSkinning.comp:37-60) as dispatched per SkinnedMeshRendererComponent by
 * Renderer::updateSkinnedVertexBuffer (src/Renderer.cpp:201-240), then Renderer::updateSkinnedBLAS (src/Renderer.cpp:644-669).
 * The caller lays the arenas out like Renderer::allocateSkinnedMeshes / updateSkinnedMeshOffsetTable (src/Renderer.cpp:133-164): a
 * bind-pose copy of the mesh's vertices appended to the vertex arena (the destination range), an offset-table entry {material,
 * that vertex offset, the mesh's index offset} and an instance with mask VKX_INSTANCE_SKINNED pointing at it.
 *   jointTransforms  numJoints column-major mat4 (the reference's jointPoses: inverse(parent global) * joint global * inverse bind)
 *   skinJoints       4 joint indices per vertex (uint16, as the shader reads them), skinWeights: 4 floats per vertex
 *   src/dstOffset    first vertex of the mesh and of the skinned copy (push constants srcOffset / dstOffset), size: vertices
 *   motionVectors    optional host array, 4 floats per vertex: new position - previous skinned position, w = 1
 * Writes the skinned positions into the destination range and, as the shader does, the skinned normals / tangents into the
 * *source* range. The BVH becomes stale: call vkx_bvh_build next (the reference rebuilds the skinned BLASes in place). */
int vkx_skin_vertices(vkx_ctx* ctx, const float* jointTransforms, size_t numJoints, const uint16_t* skinJoints,
                      const float* skinWeights, uint32_t srcOffset, uint32_t dstOffset, uint32_t size, float* motionVectors);
/* Reads vertices of the device arena back (parity checks of vkx_skin_vertices). */
int vkx_vertices_download(vkx_ctx* ctx, size_t firstVertex, size_t count, vkx_vertex* out);
/* Deterministic binned-SAH build of the 8-wide compressed BVH on the device (replaces the driver's BLAS/TLAS
 * build, src/Renderer.cpp:272-449,525-642). Topology is bit-identical to oracle/bvh.cpp. */
int vkx_bvh_build(vkx_ctx* ctx);
/* Topology-preserving refit: Renderer::updateTLAS (src/Renderer.cpp:735-742, vkCmdBuildAccelerationStructuresKHR in UPDATE mode
 * after updateAccelerationStructureInstances, :671-733) and the in-place skinned BLAS update (:644-669). After vkx_instances_update
 * or vkx_skin_vertices: the tree, the slot assignment and the triangle order of the last vkx_bvh_build stay; the world-space
 * triangles are recomputed and every node is re-quantised from the new bounds, leaves first (one launch per level). With unchanged
 * inputs the result is the built structure byte for byte; otherwise it is a conservative hierarchy over the same triangles (hits
 * equal a rebuild's up to grazing ties) whose quality decays with the motion, as the reference's refitted TLAS does: rebuild with
 * vkx_bvh_build when that matters. Bytes are identical to oracle/bvh.cpp::refit. */
int vkx_bvh_refit(vkx_ctx* ctx);
int vkx_bvh_info_get(vkx_ctx* ctx, vkx_bvh_info* out);
/* Copies the device BVH back: nodes (80 B each) and triangles (48 B each). Either pointer may be NULL. */
int vkx_bvh_download(vkx_ctx* ctx, void* nodes, size_t nodesBytes, void* triangles, size_t trianglesBytes);

/* Parity primitive: traces n rays given in host memory (origins/directions: 3 floats each) through the device
 * BVH. anyHit != 0: terminate-on-first-hit query (shadow rays), out[i].t = 1 if occluded else -1. */
int vkx_trace(vkx_ctx* ctx, const float* origins, const float* directions, size_t n, float tmin, float tmax,
              uint32_t cullMask, int anyHit, vkx_hit* out);

/* The same through a pipeline whose hit group has anyhit.rahit (the direct-light and reflection pipelines,
 * src/RenderPasses/DirectLightPipeline.cpp:43-51): candidates on materials with an albedo texture whose alpha at the hit is below
 * 0.01 are ignored (src/shaders/anyhit.rahit:24-48). */
int vkx_trace_alpha(vkx_ctx* ctx, const float* origins, const float* directions, size_t n, float tmin, float tmax,
                    uint32_t cullMask, int anyHit, vkx_hit* out);

/* ---- DDGI: IrradianceProbes ----------------------------------------------------------------------------------- */
/* IrradianceProbes::init (src/IrradianceProbes.cpp:12-104): allocates both atlases (work + sampled), the state
 * buffer and the per-chunk ray buffer; atlases and states are zero-initialised. */
int vkx_probes_init(vkx_ctx* ctx, const vkx_grid_info* grid);
/* IrradianceProbes::initProbes (src/IrradianceProbes.cpp:357-394) -> probesInit.rgen. orientation = the push
 * constant mat4 (column-major). */
int vkx_probes_classify(vkx_ctx* ctx, const float orientation[16]);
/* IrradianceProbes::update's recorded work (src/IrradianceProbes.cpp:486-576): trace + shade, blend irradiance and
 * depth, copy borders, publish. probeIndices = the to-update list produced by selectProbesToUpdate
 * (src/IrradianceProbes.cpp:396-424), host memory; NULL = every probe in linear order ("full-volume update").
 * Asynchronous on the context's stream unless sync != 0. */
int vkx_probes_update(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16],
                      const uint32_t* probeIndices, uint32_t count, int sync);
/* Readback of the *sampled* atlases (what FinalGather/closest-hit read), packed as the reference formats:
 * irradiance B10G11R11_UFLOAT_PACK32 [8*rz rows][8*rx*ry cols] u32, depth R16G16_SFLOAT same shape x2, state u32[P].
 * rays (optional) = the RGBA32F ray buffer of the last update [count][raysPerProbe] (rgb, depth). NULL skips. */
int vkx_probes_download(vkx_ctx* ctx, uint32_t* irradiance, uint32_t* depth, uint32_t* state, float* rays,
                        size_t raysCapacityBytes);
/* Asynchronous read-back of the sampled atlases as they are at the time of the call. The copies run on a copy stream; on a
 * single-GPU context they are started by the next update right before its primary traversal (or by vkx_probes_download_wait /
 * any other writer of the atlases, whichever comes first), so that they overlap one long kernel instead of the update's small
 * set-up launches; that update's publish waits for them. The host buffers must stay valid until vkx_probes_download_wait returns
 * and should be pinned (cudaHostAlloc / torch pin_memory). VKX_READBACK=eager queues the copies immediately (A/B). */
int vkx_probes_download_async(vkx_ctx* ctx, uint32_t* irradiance, uint32_t* depth, uint32_t* state);
/* Same for the z-slices [z0, z1) only (the contiguous rows one rank of a sharded run owns); pointers address the first copied row. */
int vkx_probes_download_slab_async(vkx_ctx* ctx, uint32_t z0, uint32_t z1, uint32_t* irradiance, uint32_t* depth, uint32_t* state);
int vkx_probes_download_wait(vkx_ctx* ctx);
/* On-device IrradianceProbes::selectProbesToUpdate (src/IrradianceProbes.cpp:396-424): the round-robin scan over the probe
 * states runs as a stream compaction on the GPU, so a frame neither reads the P state words back nor uploads a list; only the
 * list length returns to the host (the hysteresis ramp of update() needs it, :462-476). probesPerUpdate = the reference's
 * ProbesPerUpdate (0 = no limit). The list is identical to vkx_host_select_probes on the same states and counters.
 * vkx_probes_update_scheduled consumes the list of the preceding vkx_probes_schedule (one schedule per update). */
int vkx_probes_schedule(vkx_ctx* ctx, uint32_t probesPerUpdate, uint32_t* count);
int vkx_probes_update_scheduled(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16], int sync);
/* The scheduler's two counters {s_LoopIndex, _lastUpdateOffset} (both statics/members of the reference): set (or NULL), get (or NULL). */
int vkx_probes_scheduler_state(vkx_ctx* ctx, const uint32_t set[2], uint32_t get[2]);
/* The list the last vkx_probes_schedule produced (parity / debugging). indices may be NULL to query the count only. */
int vkx_probes_scheduled_list(vkx_ctx* ctx, uint32_t* indices, uint32_t capacity, uint32_t* count);
/* Checkpoint/resume of GI state. NULL skips an array. */
int vkx_probes_upload(vkx_ctx* ctx, const uint32_t* irradiance, const uint32_t* depth, const uint32_t* state);
/* Enables the parity side buffers (hit records, shadow flags, unpacked blend results) and forces single-chunk updates. */
int vkx_probes_debug(vkx_ctx* ctx, int enable);
/* Debug/parity: pre-pack fp32 blend results of the last update: irr [count][36][3], depth [count][196][2]. */
int vkx_probes_download_unpacked(vkx_ctx* ctx, float* irr, float* depth);
/* Hit records (vkx_hit) of the primary rays of the last update, [count][raysPerProbe]; and the shadow-ray
 * visibility bytes (0 = not traced, 1 = lit, 2 = shadowed). NULL skips. */
int vkx_probes_download_hits(vkx_ctx* ctx, vkx_hit* hits, uint8_t* shadow);
/* getComputeTimes/TraceTimes/UpdateTimes/BorderCopyTimes/CopyTimes (src/IrradianceProbes.hpp:68-72), last update:
 * ms[0] full, ms[1] trace+shade, ms[2] blend (+borders, fused), ms[3] border (0: fused), ms[4] publish. Syncs. */
int vkx_probes_timings(vkx_ctx* ctx, float ms[5]);
/* Device time of the four kernels of the first chunk of the last update: ms[0] k_trace_primary, ms[1] k_shade,
 * ms[2] k_trace_shadow, ms[3] k_blend; *probes = probes in that chunk, *shadowRays = shadow rays of the last chunk. Syncs. */
int vkx_probes_kernel_timings(vkx_ctx* ctx, float ms[4], uint32_t* probes, uint32_t* shadowRays);
/* Device pointers for zero-copy consumers (e.g. torch tensors over the sampled atlases). */
int vkx_probes_device_ptrs(vkx_ctx* ctx, void** irradiance, void** depth, void** state);

/* ---- multi-GPU: probe z-slabs + atlas all-gather (no reference equivalent) -------------------------- */
/* ncclUniqueId is 128 bytes; create it on rank 0 with vkx_comm_unique_id and broadcast it out of band. */
int vkx_comm_unique_id(void* id128);
int vkx_comm_init(vkx_ctx* ctx, int rank, int nranks, const void* id128);
/* Optional: blend fused with the atlas exchange over NVLink peer memory instead of the NCCL all-gather. After vkx_comm_init and
 * vkx_probes_init every rank exports a CUDA IPC handle of its atlas slab (two atlas sets + arrival flags), the host passes all
 * handles around (like the NCCL id) and every rank imports them. From then on vkx_probes_update_sharded lets k_blend store its
 * finished tiles straight into every rank's next atlas set, raises an arrival flag in peer memory, and the next reader of the
 * sampled atlases waits for all flags on the device - no collective, no host synchronisation. Results are identical. */
int vkx_comm_p2p_export(vkx_ctx* ctx, void* handle64);
int vkx_comm_p2p_import(vkx_ctx* ctx, const void* handles /* nranks x 64 bytes, rank order; NULL, 0 = back to NCCL */, int count);
/* How the peer-memory exchange moves the rows (every rank must choose the same; default 0 after an import):
 *   0  the blend kernel stores its tiles straight into every peer (the fused path described above; CUDA-core blend);
 *   1  the blend writes the rank's own rows locally (tensor-core blend) and copy engines push them to every peer's next set over
 *      NVLink (strided DMA copies, then the arrival flag as a 4-byte copy in the same stream): no SM takes part in the exchange,
 *      so it runs beside the next update's persistent traversal kernel, which leaves no room for a collective's thread blocks. */
int vkx_comm_p2p_mode(vkx_ctx* ctx, int copyEngines);
/* Full-volume update of this rank's z-slices (vkx_shard_slices) followed by the all-gathers of the atlas / state rows. */
int vkx_probes_update_sharded(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light,
                              const float orientation[16], int sync);
/* Partial update of a to-update list on several GPUs (the reference's ProbesPerUpdate / refresh-period scheduling,
 * src/IrradianceProbes.cpp:396-424, sharded): every rank passes the SAME list; rank r traces and blends the list positions
 * vkx_shard_range gives it, the updated probes' tiles (2 x 16x16 depth + 8x8 irradiance texels + state word, 1296 bytes per probe)
 * are packed, all-gathered and written into every rank's work and sampled atlases. Results equal vkx_probes_update with the same
 * list on one GPU. With one rank (or without vkx_comm_init) it is vkx_probes_update. */
int vkx_probes_update_sharded_list(vkx_ctx* ctx, const vkx_grid_info* grid, const vkx_light* light, const float orientation[16],
                                   const uint32_t* probeIndices, uint32_t count, int sync);
/* Orders the context's stream after a pending atlas exchange (the deferred all-gather of vkx_probes_update_sharded); returns at
 * once. An event recorded on vkx_stream() afterwards completes when the exchange has landed (bench.py times with it). */
int vkx_stream_wait_exchange(vkx_ctx* ctx);
/* The sharding arithmetic itself (host only, no context). Full-volume sharded update: every rank owns *numGroups groups of
 * *groupSlices consecutive z-slices, dealt round robin (two slices per group = whole 2x2x2 probe blocks, which spreads every region of
 * the volume over all ranks; an odd number of slices per rank gives one group = one slab); vkx_shard_slices returns group `group` of
 * rank `rank` as [*z0, *z1). The slices of all ranks in one group are consecutive atlas rows: one all-gather per group. To-update
 * lists: vkx_shard_range gives the positions [*first, *first + *n) of a `count`-long list that rank `rank` processes in
 * vkx_probes_update_sharded_list. VKX_E_INVALID for rz not divisible by nranks / bad rank / bad group. */
int vkx_shard_groups(uint32_t rz, int nranks, uint32_t* groupSlices, uint32_t* numGroups);
int vkx_shard_slices(uint32_t rz, int nranks, int rank, uint32_t group, uint32_t* z0, uint32_t* z1);
int vkx_shard_range(uint32_t count, int nranks, int rank, uint32_t* first, uint32_t* n);

/* ---- sun shadows: DirectLight pass ---------------------------------------------------------------------------- */
/* Blue-noise slices (RGBA32F = byte/255, src/vulkan/Image.cpp:62-69): [slices][h][w][4]. */
int vkx_shadow_set_noise(vkx_ctx* ctx, const float* rgba, uint32_t w, uint32_t h, uint32_t slices);
int vkx_shadow_init(vkx_ctx* ctx, uint32_t width, uint32_t height);
/* Fixture generator for the G-buffer the raster pass produces (src/shaders/GBuffer.frag:64-68): primary rays
 * through the same BVH. Fills the device-resident positionDepth / normalMetalness images. */
int vkx_gbuffer_generate(vkx_ctx* ctx, const vkx_camera* cam);
/* Or upload a host G-buffer (RGBA32F, [h][w][4]). */
int vkx_gbuffer_upload(vkx_ctx* ctx, const float* positionDepth, const float* normalMetalness);
int vkx_gbuffer_download(vkx_ctx* ctx, float* positionDepth, float* normalMetalness);
/* The two G-buffer targets only the composite reads: albedoRoughness and emissive (src/shaders/GBuffer.frag:66-67).
 * vkx_gbuffer_generate fills them too (albedo = interpolated vertex colour * baseColorFactor, untextured). */
int vkx_gbuffer_upload_material(vkx_ctx* ctx, const float* albedoRoughness, const float* emissive);
int vkx_gbuffer_download_material(vkx_ctx* ctx, float* albedoRoughness, float* emissive);
/* One frame of directLight.rgen -> directLightFilterX -> directLightFilterY (src/SwapchainManagement.cpp:409-438)
 * with the history ping-pong of src/Editor.cpp:287-316. */
int vkx_shadow_frame(vkx_ctx* ctx, const vkx_camera* cur, const vkx_camera* prev, const vkx_light* light, int sync);
/* Reflection pass: src/shaders/reflection.rgen:117-189 (one jittered reflection ray per pixel with roughness < 0.4
 * or metalness > 0.01, shaded by the same closest-hit / miss / shadow shaders as the probe rays) -> reflectionFilterX -> reflectionFilterY
 * with reprojected history (src/shaders/reflectionFilter.glsl, src/SwapchainManagement.cpp:401-455). Needs the G-buffer (including the
 * albedoRoughness target), the noise slices of vkx_shadow_set_noise and an initialised irradiance volume. Static scenes: motion vectors 0. */
int vkx_reflection_frame(vkx_ctx* ctx, const vkx_camera* cur, const vkx_camera* prev, const vkx_light* light, int sync);
/* stage: 0 = raw 1-spp (rgb, roughness), 1 = after filter X, 2 = final (rgb, depth). RGBA32F. */
int vkx_reflection_download(vkx_ctx* ctx, int stage, float* rgba);
/* Parity side buffers of the last frame: jittered directions [h][w][4], closest hits (t = -1: none), mask bytes
 * (0 no ray, 1 miss, 2 back face, 3 front face lit, 4 front face shadowed). hits / mask are only recorded while
 * vkx_probes_debug is enabled. */
int vkx_reflection_download_debug(vkx_ctx* ctx, float* dirs4, vkx_hit* hits, uint8_t* mask);
int vkx_reflection_reset_history(vkx_ctx* ctx);
/* ms[4] = full, trace + shade, filter X, filter Y of the last vkx_reflection_frame. */
int vkx_reflection_timings(vkx_ctx* ctx, float ms[4]);
/* Final composite of the frame, src/shaders/FinalGather.frag:38-77 (drawn by src/SwapchainManagement.cpp:466-474): sky on
 * empty pixels, else direct * (filtered shadow of the last vkx_shadow_frame) + specular * reflection + sampleProbes(sampled
 * atlases) * diffuse + emissive. reflection: a host RGBA32F [h][w][4] image, or NULL = the device-resident result of the last
 * vkx_reflection_frame (black if none was run).
 * Output: linear RGBA32F, device resident; vkx_final_gather_download copies it out and returns the kernel time. */
int vkx_final_gather(vkx_ctx* ctx, const vkx_camera* cam, const vkx_light* light, const float* reflection, int sync);
int vkx_final_gather_download(vkx_ctx* ctx, float* rgba, float* ms);
/* stage: 0 = raw 1-spp (directLight.rgen output), 1 = after filter X, 2 = final (after Y + temporal). RGBA32F. */
int vkx_shadow_download(vkx_ctx* ctx, int stage, float* rgba);
/* Parity side buffers of the last frame: jittered light directions [h][w][4] floats, mask bytes (0 not traced, 1 lit, 2 shadowed). */
int vkx_shadow_download_debug(vkx_ctx* ctx, float* dirs4, uint8_t* mask);
int vkx_shadow_reset_history(vkx_ctx* ctx);
/* ms[0] full, ms[1] trace, ms[2] filter X, ms[3] filter Y. */
int vkx_shadow_timings(vkx_ctx* ctx, float ms[4]);

/* ---- host logic of IrradianceProbes.cpp (pure CPU helpers, no device work) ------------------------------------------ */
/* The per-frame random orientation push constant: glm::sphericalRand(1) + genBasis -> mat4(transpose(mat3(X, Y, Z)))
 * (src/IrradianceProbes.cpp:347-355, 455-460). *rngState is the MSVC rand() state; the reference never seeds it: start at 1. */
void vkx_host_next_orientation(uint32_t* rngState, float orientation[16]);
/* Host-side scene code without a context or GPU: Scene::loadScene + Scene::update (src/Scene.cpp:818-961), Renderer::allocateMeshes
 * (arenas + offset table, src/Renderer.cpp:57-131) and Renderer::createTLAS's instance list (src/Renderer.cpp:512-575), i.e. exactly
 * the arrays vkx_scene_textures / vkx_scene_upload take. counts = {vertices, indices, meshes, materials, instances, textures};
 * vkx_host_scene_copy fills caller arrays of those sizes (any pointer may be NULL) and the scene bounds (IrradianceProbes::init's
 * extent, src/VulkanLifecycle.cpp:132-133); vkx_host_scene_texture returns texture i with pixels pointing into the scene object
 * (valid until vkx_host_scene_free); vkx_host_scene_save is Scene::save (src/Scene.cpp:710-816). */
typedef struct vkx_host_scene vkx_host_scene;
int vkx_host_scene_load(const char* path, vkx_host_scene** out);
void vkx_host_scene_free(vkx_host_scene* scene);
int vkx_host_scene_counts(const vkx_host_scene* scene, size_t counts[6]);
int vkx_host_scene_copy(const vkx_host_scene* scene, vkx_vertex* vertices, uint32_t* indices, vkx_offset_entry* offsets,
                        uint32_t* meshIndexCounts, vkx_material* materials, vkx_instance* instances, float boundsMinMax[6]);
int vkx_host_scene_texture(const vkx_host_scene* scene, size_t index, vkx_texture* desc);
int vkx_host_scene_save(const vkx_host_scene* scene, const char* path);
/* selectProbesToUpdate (src/IrradianceProbes.cpp:396-424). loopIndex / lastUpdateOffset are the function's statics. */
uint32_t vkx_host_select_probes(uint32_t* loopIndex, uint32_t* lastUpdateOffset, const uint32_t* state, uint32_t probeCount,
                                uint32_t probesPerUpdate, uint32_t* out);

/* Kernel launch counter since context creation (bench.py's gpu_launches claim). */
uint64_t vkx_launch_count(vkx_ctx* ctx);
/* The context's CUDA stream (cudaStream_t) so callers can order torch work after it. */
void* vkx_stream(vkx_ctx* ctx);
int vkx_sync(vkx_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VKX_H */
