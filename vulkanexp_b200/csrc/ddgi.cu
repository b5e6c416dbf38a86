// DDGI probe update on the device. Replaces the command buffer IrradianceProbes::update records (reference
// src/IrradianceProbes.cpp:486-576): traceProbes.rgen -> {closesthit_noreflection.rchit | miss.rmiss, shadow.rmiss}
// -> probesUpdateIrradiance.comp + probesUpdateDepth.comp -> probesCopyBorders.comp -> 2x vkCmdCopyImage.
//
// Kernel pipeline per chunk of probes (wavefront, rays grouped by what they need next):
//   k_trace_primary  closest-hit traversal of probe rays                              -> hit records
//   k_shade          miss: sky | back face: 0.8 t | front: material + 2x sampleProbes -> ray records, shadow queue
//   k_trace_shadow   any-hit traversal of the compacted shadow-ray queue, adds the direct term
//   k_blend          one CTA per probe: ray records staged in shared memory, 196 depth + 36 irradiance texels,
//                    hysteresis mix against the work atlas, state machine, border texels, packed tile stores
//   k_publish        work -> sampled for the updated probes (the reference copies both whole atlases)
#include "common.cuh"
#include "traverse.cuh"
#include "shade.cuh"

namespace {

struct TraceParams {
    vkx_grid_info grid;
    float tmin, tmax;
    uint32_t raysPerProbe, numRays; // numRays = chunk probes * raysPerProbe
};

__global__ void __launch_bounds__(128) k_trace_primary(DeviceScene sc, TraceParams tp, const uint32_t* __restrict__ probeIndices,
                                                       const float4* __restrict__ dirs, vkx_hit* __restrict__ hits) {
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= tp.numRays) return;
    const uint32_t slot = ri / tp.raysPerProbe, ray = ri - slot * tp.raysPerProbe;
    int ix, iy, iz; probeGridIndex(__ldg(probeIndices + slot), tp.grid, ix, iy, iz);
    const v3 o = probeWorldPos(ix, iy, iz, tp.grid);
    const float4 d = __ldg(dirs + ray);
    const Ray r = makeRay(o.x, o.y, o.z, d.x, d.y, d.z);
    HitRec h;
    traverse<false>(sc.nodes, sc.tris, r, tp.tmin, tp.tmax, VKX_INSTANCE_STATIC | VKX_INSTANCE_DYNAMIC, h);
    vkx_hit out; out.t = h.t; out.instance = h.inst; out.primitive = h.prim; out.u = h.u; out.v = h.v;
    hits[ri] = out;
}

struct ShadeParams {
    vkx_grid_info grid;
    vkx_light light;
    uint32_t raysPerProbe, numRays;
};

__global__ void __launch_bounds__(128) k_shade(DeviceScene sc, DeviceProbes pr, ShadeParams sp, const uint32_t* __restrict__ probeIndices,
                                               const float4* __restrict__ dirs, const vkx_hit* __restrict__ hits, float4* __restrict__ rays,
                                               float4* __restrict__ queue, uint32_t* __restrict__ queueCount, uint8_t* __restrict__ shadowFlags) {
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= sp.numRays) return;
    const uint32_t slot = ri / sp.raysPerProbe, ray = ri - slot * sp.raysPerProbe;
    int ix, iy, iz; probeGridIndex(__ldg(probeIndices + slot), sp.grid, ix, iy, iz);
    const v3 origin = probeWorldPos(ix, iy, iz, sp.grid);
    const float4 d4 = __ldg(dirs + ray);
    const v3 direction = mk3(d4.x, d4.y, d4.z);
    const vkx_hit h = hits[ri];
    const v3 lightDir = mk3(sp.light.direction[0], sp.light.direction[1], sp.light.direction[2]);
    const v3 lightColor = mk3(sp.light.color[0], sp.light.color[1], sp.light.color[2]);
    if (shadowFlags) shadowFlags[ri] = 0;
    if (h.t < 0.0f) { // miss.rmiss
        const v3 c = skyColor(origin, direction, lightDir, lightColor, sp.light.color[3]);
        rays[ri] = make_float4(c.x, c.y, c.z, -1.0f);
        return;
    }
    if (h.primitive & 0x80000000u) { rays[ri] = make_float4(0.f, 0.f, 0.f, h.t * 0.80f); return; } // closesthit.glsl:137-141
    // front face: closesthit.glsl:143-288 (NO_REFLECTION, untextured)
    const float u = h.u, v = h.v;
    const float bx = 1.0f - u - v, by = u, bz = v;
    const v3 position = direction * h.t + origin;
    const vkx_instance* inst = sc.instances + h.instance;
    const uint32_t meshEntry = __ldg(&inst->meshEntry);
    const vkx_offset_entry oe = sc.offsets[meshEntry];
    const uint32_t prim = h.primitive & 0x7FFFFFFFu;
    v3 n[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t vi = oe.vertexOffset + __ldg(sc.indices + oe.indexOffset + 3 * prim + c);
        const float* nn = sc.vertices[vi].normal;
        n[c] = mk3(__ldg(nn), __ldg(nn + 1), __ldg(nn + 2));
    }
    const vkx_material m = sc.materials[oe.materialIndex];
    const v3 tsn = norm3(n[0] * bx + n[1] * by + n[2] * bz);
    const float* W = sc.worldToObject + size_t(h.instance) * 9; // W[row][col]
    // vec3(tsn * worldToObject): component j = dot(tsn, column j)
    const v3 normal = norm3(mk3(dot3(tsn, mk3(W[0], W[3], W[6])), dot3(tsn, mk3(W[1], W[4], W[7])), dot3(tsn, mk3(W[2], W[5], W[8]))));
    const v3 albedo = mk3(m.baseColorFactor[0], m.baseColorFactor[1], m.baseColorFactor[2]);
    const float metalness = m.metallicFactor, roughness = m.roughnessFactor;
    v3 color = mk3(0.0f) + mk3(m.emissiveFactor[0], m.emissiveFactor[1], m.emissiveFactor[2]);
    const v3 f0 = mk3(0.04f);
    v3 diffuseColor = albedo * (1.0f - f0);
    diffuseColor = diffuseColor * (1.0f - metalness);
    const v3 specularColor = mix3(f0, albedo, metalness);
    const v3 reflectDir = reflect3(direction, normal);
    const v3 reflection = sampleProbes(pr, position, reflectDir, -direction);
    color = color + specularColor * reflection;
    const v3 indirectLight = sampleProbes(pr, position, normal, -direction);
    color = color + indirectLight * diffuseColor;
    rays[ri] = make_float4(color.x, color.y, color.z, h.t); // value if the sun is occluded
    // direct term, applied by k_trace_shadow if the shadow ray escapes
    v3 lit = color + pbrMetallicRoughness(normal, norm3(-direction), lightColor, lightDir, albedo, metalness, roughness);
    if (lightDir.y < 0.0f) lit = lit * (1.0f - clampS(-lightDir.y, 0.0f, 0.1f) / 0.1f);
    // warp-aggregated append to the shadow queue
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(int(active)) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(queueCount, uint32_t(__popc(active)));
    base = __shfl_sync(active, base, leader);
    const uint32_t qi = base + uint32_t(__popc(active & ((1u << lane) - 1u)));
    queue[2 * size_t(qi)] = make_float4(position.x, position.y, position.z, __uint_as_float(ri));
    queue[2 * size_t(qi) + 1] = make_float4(lit.x, lit.y, lit.z, 0.0f);
}

__global__ void __launch_bounds__(128) k_trace_shadow(DeviceScene sc, vkx_light light, const float4* __restrict__ queue, const uint32_t* __restrict__ queueCount,
                                                      float4* __restrict__ rays, uint8_t* __restrict__ shadowFlags) {
    const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= *queueCount) return;
    const float4 q0 = queue[2 * size_t(qi)], q1 = queue[2 * size_t(qi) + 1];
    const uint32_t ri = __float_as_uint(q0.w);
    const Ray r = makeRay(q0.x, q0.y, q0.z, light.direction[0], light.direction[1], light.direction[2]);
    HitRec h;
    const bool shadowed = traverse<true>(sc.nodes, sc.tris, r, 0.1f, 10000.0f, 0xFFu, h); // closesthit.glsl:270-281
    if (!shadowed) { float4 rec = rays[ri]; rec.x = q1.x; rec.y = q1.y; rec.z = q1.z; rays[ri] = rec; }
    if (shadowFlags) shadowFlags[ri] = shadowed ? 2 : 1;
}

// ------------------------------------------------------------------------------------------------ blend
__device__ __forceinline__ void borderSource(int T, int x, int y, int& sx, int& sy) { // probesCopyBorders.comp:21-220 as a formula
    const int L = T - 1;
    const bool bx = (x == 0 || x == L), by = (y == 0 || y == L);
    if (bx && by) { sx = x == 0 ? L - 1 : 1; sy = y == 0 ? L - 1 : 1; }
    else if (bx) { sx = x == 0 ? 1 : L - 1; sy = L - y; }
    else { sx = L - x; sy = y == 0 ? 1 : L - 1; }
}

struct BlendParams {
    vkx_grid_info grid;
    uint32_t raysPerProbe;
    float gridCellLen;   // length(probeGridCellSize)
    int intSharpness;    // > 0: depthSharpness is this small integer (pow by repeated multiplication)
};

// One CTA per updated probe. Threads 0..195: depth texels, 196..231: irradiance texels. Accumulation order over rays is
// sequential (i = 0..N-1) with separate multiply and add, like the oracle, so packed texels agree bit for bit whenever
// the ray records do.
__global__ void __launch_bounds__(256) k_blend(BlendParams bp, DeviceProbes pr, const uint32_t* __restrict__ probeIndices, const float4* __restrict__ rays,
                                               const float4* __restrict__ dirs, float* __restrict__ irrUnpacked, float* __restrict__ depUnpacked, uint32_t slotBase) {
    __shared__ float4 sRay[VKX_MAX_RAYS_PER_PROBE];
    __shared__ float4 sDir[VKX_MAX_RAYS_PER_PROBE];
    __shared__ uint32_t sIrr[64];
    __shared__ uint32_t sDep[256];
    __shared__ uint32_t sMaxChange;
    __shared__ uint32_t sOutOfRange;
    const uint32_t slot = blockIdx.x, tid = threadIdx.x, N = bp.raysPerProbe;
    const uint32_t linearIndex = __ldg(probeIndices + slot);
    int ix, iy, iz; probeGridIndex(linearIndex, bp.grid, ix, iy, iz);
    for (uint32_t i = tid; i < N; i += blockDim.x) { sRay[i] = rays[size_t(slot) * N + i]; sDir[i] = __ldg(dirs + i); }
    if (tid == 0) { sMaxChange = 0u; sOutOfRange = 0u; }
    __syncthreads();
    const float hysteresis = bp.grid.hysteresis;
    const float cellLen = bp.gridCellLen;
    const int tile = iy * bp.grid.resolution[0] + ix;
    if (tid < 196) { // ---- depth texel (probesUpdate.glsl, DEPTH)
        const int lx = int(tid % 14u), ly = int(tid / 14u);
        const v3 td = octDecode(0.142857f * (float(lx) - 6.5f), 0.142857f * (float(ly) - 6.5f));
        float r0 = 0.f, r1 = 0.f, rw = 0.f;
        const float sharp = bp.grid.depthSharpness;
        const int ip = bp.intSharpness;
        for (uint32_t i = 0; i < N; ++i) {
            const float4 rd = sRay[i]; const float4 dd = sDir[i];
            float depth = minS(cellLen, rd.w);
            if (depth < 0.0f) depth = cellLen;
            const float c = maxS(0.0f, td.x * dd.x + td.y * dd.y + td.z * dd.z);
            float weight;
            if (ip > 0) { // c^ip by square-and-multiply (exact-integer exponent fast path; <= 2 ulp from powf)
                float b = c; int e = ip; weight = 1.0f;
                while (e) { if (e & 1) weight = weight * b; b = b * b; e >>= 1; }
            } else weight = powf(c, sharp);
            r0 = r0 + weight * depth;
            r1 = r1 + weight * depth * depth;
            rw = rw + weight;
        }
        if (rw > 1e-3f) { r0 = r0 / rw; r1 = r1 / rw; }
        const size_t gi = size_t(16 * iz + 1 + ly) * pr.depW + size_t(16 * tile + 1 + lx);
        const float2 prev = unpackRG16F(pr.depWork[gi]);
        const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis);
        sDep[(ly + 1) * 16 + (lx + 1)] = packRG16F(o0, o1);
        if (depUnpacked) { float* up = depUnpacked + (size_t(slotBase + slot) * 196 + tid) * 2; up[0] = o0; up[1] = o1; }
    } else if (tid < 232) { // ---- irradiance texel (probesUpdate.glsl, IRRADIANCE)
        const uint32_t t = tid - 196u;
        const int lx = int(t % 6u), ly = int(t / 6u);
        const v3 td = octDecode(0.33333f * (float(lx) - 2.5f), 0.33333f * (float(ly) - 2.5f));
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, rw = 0.f; uint32_t outOfRange = 0;
        for (uint32_t i = 0; i < N; ++i) {
            const float4 rd = sRay[i]; const float4 dd = sDir[i];
            if (rd.w < 0.0f || rd.w > cellLen) ++outOfRange;
            const float weight = maxS(0.0f, td.x * dd.x + td.y * dd.y + td.z * dd.z);
            r0 = r0 + weight * rd.x; r1 = r1 + weight * rd.y; r2 = r2 + weight * rd.z; rw = rw + weight;
        }
        if (rw > 1e-3f) { r0 = r0 / rw; r1 = r1 / rw; r2 = r2 / rw; }
        const size_t gi = size_t(8 * iz + 1 + ly) * pr.irrW + size_t(8 * tile + 1 + lx);
        const float3 prev = unpackR11G11B10(pr.irrWork[gi]);
        const float maxChange = maxS(maxS(fabsf(r0 - prev.x), fabsf(r1 - prev.y)), fabsf(r2 - prev.z));
        const float o0 = mixf(r0, prev.x, hysteresis), o1 = mixf(r1, prev.y, hysteresis), o2 = mixf(r2, prev.z, hysteresis);
        sIrr[(ly + 1) * 8 + (lx + 1)] = packR11G11B10(o0, o1, o2);
        if (irrUnpacked) { float* up = irrUnpacked + (size_t(slotBase + slot) * 36 + t) * 3; up[0] = o0; up[1] = o1; up[2] = o2; }
        atomicMax(&sMaxChange, __float_as_uint(maxChange)); // probesUpdate.glsl:106-107 (non-negative floats order as uints)
        if (t == 0) sOutOfRange = outOfRange;
    }
    __syncthreads();
    if (tid == 0) { // state machine, probesUpdate.glsl:110-119 (decree A.5.3: full max over the 36 texels)
        uint32_t st = pr.stateWork[linearIndex];
        if (sOutOfRange >= N) st = 8;
        else {
            const float maxChange = __uint_as_float(sMaxChange);
            if (maxChange < 0.02f / float(st)) st = min(st + 1u, 8u);
            else if (maxChange > 0.04f / float(st)) st = max(st - 1u, 1u);
            else if (maxChange > 0.25f) st = 1;
        }
        pr.stateWork[linearIndex] = st;
    }
    // ---- borders (probesCopyBorders.comp) from the shared tiles
    if (tid < 60) { // depth border texels: 4 corners + 4 x 14
        int x, y;
        if (tid < 16) { x = int(tid); y = 0; } else if (tid < 32) { x = int(tid) - 16; y = 15; } else if (tid < 46) { x = 0; y = int(tid) - 32 + 1; } else { x = 15; y = int(tid) - 46 + 1; }
        int sx, sy; borderSource(16, x, y, sx, sy);
        sDep[y * 16 + x] = sDep[sy * 16 + sx];
    } else if (tid >= 64 && tid < 92) { // irradiance border texels: 28
        const int b = int(tid) - 64; int x, y;
        if (b < 8) { x = b; y = 0; } else if (b < 16) { x = b - 8; y = 7; } else if (b < 22) { x = 0; y = b - 16 + 1; } else { x = 7; y = b - 22 + 1; }
        int sx, sy; borderSource(8, x, y, sx, sy);
        sIrr[y * 8 + x] = sIrr[sy * 8 + sx];
    }
    __syncthreads();
    // ---- vectorised tile stores: depth 16 rows x 64 B (4 x uint4), irradiance 8 rows x 32 B (2 x uint4)
    if (tid < 64) {
        const int row = int(tid >> 2), q = int(tid & 3);
        uint4 v = *reinterpret_cast<const uint4*>(&sDep[row * 16 + q * 4]);
        *reinterpret_cast<uint4*>(pr.depWork + size_t(16 * iz + row) * pr.depW + size_t(16 * tile + q * 4)) = v;
    } else if (tid < 80) {
        const int k = int(tid) - 64; const int row = k >> 1, q = k & 1;
        uint4 v = *reinterpret_cast<const uint4*>(&sIrr[row * 8 + q * 4]);
        *reinterpret_cast<uint4*>(pr.irrWork + size_t(8 * iz + row) * pr.irrW + size_t(8 * tile + q * 4)) = v;
    }
}

// work -> sampled for the updated probes (tiles + state word). One warp per probe: 16 depth rows + 8 irradiance rows.
__global__ void k_publish(DeviceProbes pr, uint32_t* __restrict__ irrSampled, uint32_t* __restrict__ depSampled, uint32_t* __restrict__ stateSampled,
                          const uint32_t* __restrict__ probeIndices, uint32_t count) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= count) return;
    const uint32_t linearIndex = __ldg(probeIndices + warp);
    int ix, iy, iz; probeGridIndex(linearIndex, pr.grid, ix, iy, iz);
    const int tile = iy * pr.grid.resolution[0] + ix;
    // 64 uint4 of depth (2 per lane) + 16 uint4 of irradiance (lanes 0..15)
    for (int k = int(lane); k < 64; k += 32) {
        const int row = k >> 2, q = k & 3;
        const size_t off = size_t(16 * iz + row) * pr.depW + size_t(16 * tile + q * 4);
        *reinterpret_cast<uint4*>(depSampled + off) = *reinterpret_cast<const uint4*>(pr.depWork + off);
    }
    if (lane < 16) {
        const int row = int(lane >> 1), q = int(lane & 1);
        const size_t off = size_t(8 * iz + row) * pr.irrW + size_t(8 * tile + q * 4);
        *reinterpret_cast<uint4*>(irrSampled + off) = *reinterpret_cast<const uint4*>(pr.irrWork + off);
    }
    if (lane == 0) stateSampled[linearIndex] = pr.stateWork[linearIndex];
}

// ------------------------------------------------------------------------------------------------ classification
// probesInit.rgen:31-64 + backfaceTest.rchit + probeInitMiss.rmiss. One CTA of 128 threads per probe, 4 rays each.
__global__ void __launch_bounds__(128) k_classify(DeviceScene sc, vkx_grid_info grid, const float4* __restrict__ dirs512, uint32_t* __restrict__ stateWork,
                                                  uint32_t* __restrict__ stateSampled) {
    __shared__ uint32_t sBack, sAffect;
    const uint32_t li = blockIdx.x;
    if (threadIdx.x == 0) { sBack = 0; sAffect = 0; }
    __syncthreads();
    int ix, iy, iz; probeGridIndex(li, grid, ix, iy, iz);
    const v3 origin = probeWorldPos(ix, iy, iz, grid);
    const v3 cell = gridCellSize(grid);
    const float maxDistance = len3(cell);
    const float tmax = 1.5f * maxDistance;
    uint32_t back = 0, affect = 0;
    for (uint32_t i = threadIdx.x; i < 512u; i += blockDim.x) {
        const float4 d = __ldg(dirs512 + i);
        const Ray r = makeRay(origin.x, origin.y, origin.z, d.x, d.y, d.z);
        HitRec h;
        float depth = 3.402823466e+38f; bool isBack = false;
        if (traverse<false>(sc.nodes, sc.tris, r, 0.01f, tmax, VKX_INSTANCE_STATIC, h)) { depth = h.t; isBack = (h.prim & 0x80000000u) != 0u; }
        if (depth < maxDistance) {
            if (isBack) ++back;
            const v3 position = origin + depth * mk3(d.x, d.y, d.z);
            const v3 dist = abs3(position - origin);
            if (dist.x < cell.x && dist.y < cell.y && dist.z < cell.z) affect = 1;
        }
    }
    if (back) atomicAdd(&sBack, back);
    if (affect) atomicOr(&sAffect, 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t st;
        if (float(sBack) > 0.5f * float(grid.raysPerProbe)) st = 0;
        else st = sAffect ? 1u : 8u;
        stateWork[li] = st; stateSampled[li] = st;
    }
}

} // namespace

DeviceScene deviceScene(const vkx_ctx* ctx) {
    DeviceScene s;
    s.vertices = ctx->dVertices; s.indices = ctx->dIndices; s.offsets = ctx->dOffsets; s.materials = ctx->dMaterials; s.instances = ctx->dInstances;
    s.worldToObject = ctx->dWorldToObject; s.nodes = ctx->dNodes; s.tris = ctx->dTris;
    return s;
}

static DeviceProbes deviceProbes(const vkx_ctx* ctx) {
    DeviceProbes p;
    p.grid = ctx->grid; p.irrW = ctx->irrW; p.irrH = ctx->irrH; p.depW = ctx->depW; p.depH = ctx->depH; p.probeCount = ctx->probeCount;
    p.irrSampled = ctx->dIrrSampled; p.depSampled = ctx->dDepSampled; p.stateSampled = ctx->dStateSampled;
    p.irrWork = ctx->dIrrWork; p.depWork = ctx->dDepWork; p.stateWork = ctx->dStateWork;
    return p;
}

int ddgiClassify(vkx_ctx* ctx, const float* dirs512) {
    cudaStream_t st = ctx->stream;
    std::vector<float4> d(512);
    for (int i = 0; i < 512; ++i) d[i] = make_float4(dirs512[3 * i], dirs512[3 * i + 1], dirs512[3 * i + 2], 0.f);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dDirs, d.data(), 512 * sizeof(float4), cudaMemcpyHostToDevice, st));
    k_classify<<<ctx->probeCount, 128, 0, st>>>(deviceScene(ctx), ctx->grid, ctx->dDirs, ctx->dStateWork, ctx->dStateSampled);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return VKX_OK;
}

// Traces + blends `count` probes whose indices are already in ctx->dIndicesList[listOffset ...]. Directions are in ctx->dDirs.
int ddgiUpdate(vkx_ctx* ctx, const vkx_light& light, const uint32_t* /*unused*/, uint32_t count, uint32_t listOffset, bool /*unused*/) {
    cudaStream_t st = ctx->stream;
    const uint32_t N = ctx->grid.raysPerProbe;
    const DeviceScene sc = deviceScene(ctx);
    const DeviceProbes pr = deviceProbes(ctx);
    const float ex = ctx->grid.extentMax[0] - ctx->grid.extentMin[0], ey = ctx->grid.extentMax[1] - ctx->grid.extentMin[1], ez = ctx->grid.extentMax[2] - ctx->grid.extentMin[2];
    const float tmax = sqrtf(ex * ex + ey * ey + ez * ez); // traceProbes.rgen:33
    const float cx = ex / float(ctx->grid.resolution[0] - 1), cy = ey / float(ctx->grid.resolution[1] - 1), cz = ez / float(ctx->grid.resolution[2] - 1);
    BlendParams bp; bp.grid = ctx->grid; bp.raysPerProbe = N; bp.gridCellLen = sqrtf(cx * cx + cy * cy + cz * cz);
    const float sh = ctx->grid.depthSharpness;
    bp.intSharpness = (sh >= 1.0f && sh <= 64.0f && sh == floorf(sh)) ? int(sh) : 0;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
    for (uint32_t base = 0; base < count; base += ctx->chunkProbes) {
        const uint32_t n = std::min(ctx->chunkProbes, count - base);
        const uint32_t numRays = n * N;
        const uint32_t* idx = ctx->dIndicesList + listOffset + base;
        TraceParams tp; tp.grid = ctx->grid; tp.tmin = 0.01f; tp.tmax = tmax; tp.raysPerProbe = N; tp.numRays = numRays;
        ShadeParams sp; sp.grid = ctx->grid; sp.light = light; sp.raysPerProbe = N; sp.numRays = numRays;
        const bool timed = base == 0; // per-kernel events on the first chunk
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->dQueueCount, 0, 4, st));
        if (timed) { CUDA_TRY(ctx, cudaEventRecord(ctx->kev[0], st)); ctx->kevProbes = n; }
        k_trace_primary<<<divUp(numRays, 128), 128, 0, st>>>(sc, tp, idx, ctx->dDirs, ctx->dHits); LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[1], st));
        k_shade<<<divUp(numRays, 128), 128, 0, st>>>(sc, pr, sp, idx, ctx->dDirs, ctx->dHits, ctx->dRays, ctx->dShadowQueue, ctx->dQueueCount, ctx->debugBuffers ? ctx->dShadowFlags : nullptr); LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[2], st));
        k_trace_shadow<<<divUp(numRays, 128), 128, 0, st>>>(sc, light, ctx->dShadowQueue, ctx->dQueueCount, ctx->dRays, ctx->debugBuffers ? ctx->dShadowFlags : nullptr); LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[3], st));
        if (base + n >= count) CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
        k_blend<<<n, 256, 0, st>>>(bp, pr, idx, ctx->dRays, ctx->dDirs, ctx->debugBuffers ? ctx->dIrrUnpacked : nullptr, ctx->debugBuffers ? ctx->dDepUnpacked : nullptr, base); LAUNCH_CHECK(ctx);
        if (timed) CUDA_TRY(ctx, cudaEventRecord(ctx->kev[4], st));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
    ctx->lastCount = count; ctx->lastRays = count * N;
    return VKX_OK;
}

int ddgiPublish(vkx_ctx* ctx, uint32_t count) {
    cudaStream_t st = ctx->stream;
    const DeviceProbes pr = deviceProbes(ctx);
    if (count) { k_publish<<<divUp(size_t(count) * 32, 256), 256, 0, st>>>(pr, ctx->dIrrSampled, ctx->dDepSampled, ctx->dStateSampled, ctx->dIndicesList, count); LAUNCH_CHECK(ctx); }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    return VKX_OK;
}
