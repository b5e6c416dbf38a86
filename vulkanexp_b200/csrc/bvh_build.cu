// Device builder of the deterministic 8-wide compressed BVH ("BVH spec v1").
// Replaces the driver-side BLAS/TLAS build the reference requests in Renderer::createAccelerationStructures /
// createTLAS (reference src/Renderer.cpp:272-449, 525-642). Produces node/triangle arrays bit-identical to the
// sequential definition in oracle/bvh.cpp:
//   phase 0  flatten instances to world-space triangles (one thread per triangle)
//   phase 1  breadth-first binned-SAH binary tree: per level  bin (atomics on order-preserving keys; min/max and
//            integer counts are order independent) -> split (one thread per node, sequential SAH sweep) ->
//            stable partition (global exclusive scan of "goes left" flags) -> child bounds (atomics)
//   phase 2  breadth-first greedy collapse to 8-wide nodes + octant slot assignment + 8-bit quantisation
// Compiled with --fmad=false: float expressions are evaluated exactly as written.
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <cfloat>

namespace {

constexpr int NBINS = 16;
constexpr uint32_t LEAF_MAX = 3;
constexpr uint32_t LEAF_FLAG = 0x80000000u;

__host__ __device__ inline uint32_t leafRef(uint32_t first, uint32_t count) { return LEAF_FLAG | (count << 29) | first; }
__host__ __device__ inline bool isLeafRef(uint32_t r) { return (r & LEAF_FLAG) != 0; }

struct FlatTri { float v0[3], e1[3], e2[3]; uint32_t inst, prim, pad; };
static_assert(sizeof(FlatTri) == 48, "FlatTri");

__device__ inline float halfArea(const float* lo, const float* hi) {
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return (ex * ey + ey * ez) + ez * ex;
}
__device__ inline float kmin(float a, float b) { return okey(a) <= okey(b) ? a : b; }
__device__ inline float kmax(float a, float b) { return okey(a) >= okey(b) ? a : b; }

// ------------------------------------------------------------------------------------------------ phase 0
__global__ void k_flatten(const vkx_vertex* __restrict__ vertices, const uint32_t* __restrict__ indices, const vkx_offset_entry* __restrict__ offsets,
                          const vkx_instance* __restrict__ instances, const uint32_t* __restrict__ instTriBase, uint32_t numInstances, uint32_t T,
                          FlatTri* __restrict__ flat, float* __restrict__ lo, float* __restrict__ hi, float* __restrict__ cent, uint32_t* __restrict__ rootKeys) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T) return;
    // instance = last k with instTriBase[k] <= g
    uint32_t a = 0, b = numInstances;
    while (b - a > 1) { uint32_t m = (a + b) >> 1; if (instTriBase[m] <= g) a = m; else b = m; }
    const uint32_t k = a, j = g - instTriBase[k];
    const vkx_instance in = instances[k];
    const vkx_offset_entry oe = offsets[in.meshEntry];
    const float* M = in.transform;
    float det = M[0] * (M[5] * M[10] - M[6] * M[9]) - M[1] * (M[4] * M[10] - M[6] * M[8]) + M[2] * (M[4] * M[9] - M[5] * M[8]);
    uint32_t flip = det < 0.0f ? 0x80000000u : 0u;
    float w[3][3];
    for (int c = 0; c < 3; ++c) {
        uint32_t vi = oe.vertexOffset + indices[oe.indexOffset + 3 * j + c];
        const float* p = vertices[vi].pos;
        float x = p[0], y = p[1], z = p[2];
        for (int r = 0; r < 3; ++r) w[c][r] = ((M[4 * r + 0] * x + M[4 * r + 1] * y) + M[4 * r + 2] * z) + M[4 * r + 3];
    }
    FlatTri t;
    for (int ax = 0; ax < 3; ++ax) {
        t.v0[ax] = w[0][ax];
        t.e1[ax] = w[1][ax] - w[0][ax];
        t.e2[ax] = w[2][ax] - w[0][ax];
        float l = kmin(kmin(w[0][ax], w[1][ax]), w[2][ax]);
        float h = kmax(kmax(w[0][ax], w[1][ax]), w[2][ax]);
        lo[3 * size_t(g) + ax] = l;
        hi[3 * size_t(g) + ax] = h;
        float c = (l + h) * 0.5f;
        cent[3 * size_t(g) + ax] = c;
        atomicMin(&rootKeys[ax], okey(l));
        atomicMax(&rootKeys[3 + ax], okey(h));
        atomicMin(&rootKeys[6 + ax], okey(c));
        atomicMax(&rootKeys[9 + ax], okey(c));
    }
    t.inst = k | ((in.mask & 0xFFu) << 24);
    t.prim = j | flip;
    t.pad = 0;
    flat[g] = t;
}

__global__ void k_iota(uint32_t* p, uint32_t n) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
__global__ void k_fill_i32(int* p, int v, size_t n) { size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// ------------------------------------------------------------------------------------------------ phase 1
struct ActiveArrays { // structure of arrays over the active nodes of one level
    uint32_t* first; uint32_t* count; uint32_t* id;
    float* box;  // [n][6] lo xyz, hi xyz
    float* cbox; // [n][6]
};
// per active node: bins[a][axis][bin] -> count + 6 keys
struct BinArrays { uint32_t* cnt; uint32_t* keys; }; // cnt [n][3][16], keys [n][3][16][6]

__device__ inline int binOf(float c, float clo, float k) {
    int b = int((c - clo) * k);
    return b < NBINS - 1 ? b : NBINS - 1;
}

__global__ void k_init_bins(BinArrays bins, uint32_t numActive) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t n = size_t(numActive) * 3 * NBINS;
    if (i >= n) return;
    bins.cnt[i] = 0;
    uint32_t* k = bins.keys + i * 6;
    k[0] = k[1] = k[2] = 0xFFFFFFFFu; // min keys
    k[3] = k[4] = k[5] = 0u;          // max keys
}

// Atomics are pre-reduced inside the warp: primitives are stored node by node, so near the root all 32 lanes of a warp target the
// same bin / child entry and one lane issues the atomics for the group (min / max / count are order-independent: same result).
__global__ void k_bin(const uint32_t* __restrict__ prim, const int* __restrict__ nodeOf, uint32_t T, ActiveArrays act, BinArrays bins,
                      const float* __restrict__ lo, const float* __restrict__ hi, const float* __restrict__ cent) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int a = pos < T ? nodeOf[pos] : -1;
    const bool valid = a >= 0;
    uint32_t g = 0, kl[3] = {0, 0, 0}, kh[3] = {0, 0, 0};
    const float* cb = act.cbox + size_t(valid ? a : 0) * 6;
    if (valid) {
        g = prim[pos];
        for (int ax = 0; ax < 3; ++ax) { kl[ax] = okey(lo[3 * size_t(g) + ax]); kh[ax] = okey(hi[3 * size_t(g) + ax]); }
    }
    for (int ax = 0; ax < 3; ++ax) {
        uint32_t bi = 0xFFFFFFFFu;
        if (valid) {
            const float ext = cb[3 + ax] - cb[ax];
            if (ext > 0.0f) {
                const float k = float(NBINS) / ext;
                bi = uint32_t((size_t(a) * 3 + ax) * NBINS + binOf(cent[3 * size_t(g) + ax], cb[ax], k));
            }
        }
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, bi);
        const uint32_t m0 = __reduce_min_sync(grp, kl[0]), m1 = __reduce_min_sync(grp, kl[1]), m2 = __reduce_min_sync(grp, kl[2]);
        const uint32_t x0 = __reduce_max_sync(grp, kh[0]), x1 = __reduce_max_sync(grp, kh[1]), x2 = __reduce_max_sync(grp, kh[2]);
        if (bi != 0xFFFFFFFFu && lane == uint32_t(__ffs(int(grp))) - 1u) {
            atomicAdd(&bins.cnt[bi], uint32_t(__popc(grp)));
            uint32_t* keys = bins.keys + size_t(bi) * 6;
            atomicMin(&keys[0], m0); atomicMin(&keys[1], m1); atomicMin(&keys[2], m2);
            atomicMax(&keys[3], x0); atomicMax(&keys[4], x1); atomicMax(&keys[5], x2);
        }
    }
}

struct SplitArrays { int* axis; int* plane; uint32_t* nL; uint32_t* innerFlag; /* [2n] */ };

__global__ void k_split(ActiveArrays act, BinArrays bins, SplitArrays sp, uint32_t numActive) {
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= numActive) return;
    const float* cb = act.cbox + size_t(a) * 6;
    const uint32_t count = act.count[a];
    int bestAxis = -1, bestPlane = -1; float bestCost = INFINITY; uint32_t bestNL = 0;
    for (int ax = 0; ax < 3; ++ax) {
        float ext = cb[3 + ax] - cb[ax];
        if (!(ext > 0.0f)) continue;
        const uint32_t* cnt = bins.cnt + (size_t(a) * 3 + ax) * NBINS;
        const uint32_t* keys = bins.keys + (size_t(a) * 3 + ax) * NBINS * 6;
        // right-to-left suffix: area and count of bins s+1..15
        float areaR[NBINS - 1]; uint32_t nR[NBINS - 1];
        uint32_t k[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u}; uint32_t c = 0;
        for (int s = NBINS - 2; s >= 0; --s) {
            if (cnt[s + 1]) { const uint32_t* bk = keys + (s + 1) * 6; for (int i = 0; i < 3; ++i) { k[i] = min(k[i], bk[i]); k[3 + i] = max(k[3 + i], bk[3 + i]); } }
            c += cnt[s + 1]; nR[s] = c;
            if (c) { float l[3] = {unkey(k[0]), unkey(k[1]), unkey(k[2])}, h[3] = {unkey(k[3]), unkey(k[4]), unkey(k[5])}; areaR[s] = halfArea(l, h); } else areaR[s] = 0.0f;
        }
        uint32_t kk[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u}; c = 0;
        for (int s = 0; s < NBINS - 1; ++s) {
            if (cnt[s]) { const uint32_t* bk = keys + s * 6; for (int i = 0; i < 3; ++i) { kk[i] = min(kk[i], bk[i]); kk[3 + i] = max(kk[3 + i], bk[3 + i]); } }
            c += cnt[s];
            if (c == 0 || nR[s] == 0) continue;
            float l[3] = {unkey(kk[0]), unkey(kk[1]), unkey(kk[2])}, h[3] = {unkey(kk[3]), unkey(kk[4]), unkey(kk[5])};
            float cost = halfArea(l, h) * float(c) + areaR[s] * float(nR[s]);
            if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestPlane = s; bestNL = c; }
        }
    }
    if (bestAxis < 0) bestNL = count / 2;
    sp.axis[a] = bestAxis; sp.plane[a] = bestPlane; sp.nL[a] = bestNL;
    sp.innerFlag[2 * a + 0] = bestNL > LEAF_MAX ? 1u : 0u;
    sp.innerFlag[2 * a + 1] = (count - bestNL) > LEAF_MAX ? 1u : 0u;
}

__global__ void k_left_flags(const uint32_t* __restrict__ prim, const int* __restrict__ nodeOf, uint32_t T, ActiveArrays act, SplitArrays sp,
                             const float* __restrict__ cent, uint32_t* __restrict__ flag) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= T) return;
    int a = nodeOf[pos];
    uint32_t f = 0;
    if (a >= 0) {
        int ax = sp.axis[a];
        if (ax >= 0) {
            const float* cb = act.cbox + size_t(a) * 6;
            float k = float(NBINS) / (cb[3 + ax] - cb[ax]);
            f = binOf(cent[3 * size_t(prim[pos]) + ax], cb[ax], k) <= sp.plane[a] ? 1u : 0u;
        } else f = (pos - act.first[a]) < sp.nL[a] ? 1u : 0u;
    }
    flag[pos] = f;
}

// child bounds as keys: [2n][12] = box lo/hi, cbox lo/hi
__global__ void k_init_child_keys(uint32_t* childKeys, uint32_t numChildren) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= size_t(numChildren) * 12) return;
    uint32_t w = uint32_t(i % 12);
    childKeys[i] = (w % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

__global__ void k_scatter(const uint32_t* __restrict__ prim, uint32_t* __restrict__ prim2, const int* __restrict__ nodeOf, int* __restrict__ nodeOf2, uint32_t T,
                          ActiveArrays act, SplitArrays sp, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flagScan,
                          const uint32_t* __restrict__ innerScan, uint32_t* __restrict__ childKeys,
                          const float* __restrict__ lo, const float* __restrict__ hi, const float* __restrict__ cent) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int a = pos < T ? nodeOf[pos] : -1;
    uint32_t child = 0xFFFFFFFFu, key[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) key[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u; // neutral elements of min (0..2, 6..8) and max (3..5, 9..11)
    if (pos < T) {
        const uint32_t g = prim[pos];
        if (a < 0) { prim2[pos] = g; nodeOf2[pos] = -1; }
        else {
            const uint32_t first = act.first[a], nL = sp.nL[a];
            const uint32_t rankL = flagScan[pos] - flagScan[first];
            const uint32_t side = flag[pos] ? 0u : 1u;
            const uint32_t newpos = side == 0 ? first + rankL : first + nL + ((pos - first) - rankL);
            prim2[newpos] = g;
            child = 2 * uint32_t(a) + side;
            nodeOf2[newpos] = sp.innerFlag[child] ? int(innerScan[child]) : -1;
            for (int ax = 0; ax < 3; ++ax) {
                key[ax] = okey(lo[3 * size_t(g) + ax]); key[3 + ax] = okey(hi[3 * size_t(g) + ax]);
                key[6 + ax] = key[9 + ax] = okey(cent[3 * size_t(g) + ax]);
            }
        }
    }
    // warp-level pre-reduction per child (see k_bin)
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, child);
#pragma unroll
    for (int i = 0; i < 12; ++i) key[i] = (i % 6) < 3 ? __reduce_min_sync(grp, key[i]) : __reduce_max_sync(grp, key[i]);
    if (child != 0xFFFFFFFFu && lane == uint32_t(__ffs(int(grp))) - 1u) {
        uint32_t* k = childKeys + size_t(child) * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) { if ((i % 6) < 3) atomicMin(&k[i], key[i]); else atomicMax(&k[i], key[i]); }
    }
}

struct BinaryNodes { uint32_t* ref; float* box; }; // ref [2*id + side], box [(2*id + side)*6]

__global__ void k_finalize_children(ActiveArrays act, ActiveArrays next, SplitArrays sp, const uint32_t* __restrict__ innerScan,
                                    const uint32_t* __restrict__ childKeys, BinaryNodes bn, uint32_t numActive, uint32_t nextIdBase) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * numActive) return;
    const uint32_t a = c >> 1, side = c & 1u;
    const uint32_t nL = sp.nL[a];
    const uint32_t first = side == 0 ? act.first[a] : act.first[a] + nL;
    const uint32_t count = side == 0 ? nL : act.count[a] - nL;
    const uint32_t* k = childKeys + size_t(c) * 12;
    const uint32_t pid = act.id[a];
    float* bb = bn.box + (size_t(pid) * 2 + side) * 6;
    for (int i = 0; i < 6; ++i) bb[i] = unkey(k[i]);
    if (sp.innerFlag[c]) {
        const uint32_t r = innerScan[c];
        bn.ref[size_t(pid) * 2 + side] = nextIdBase + r;
        next.first[r] = first; next.count[r] = count; next.id[r] = nextIdBase + r;
        for (int i = 0; i < 6; ++i) { next.box[size_t(r) * 6 + i] = unkey(k[i]); next.cbox[size_t(r) * 6 + i] = unkey(k[6 + i]); }
    } else bn.ref[size_t(pid) * 2 + side] = leafRef(first, count);
}

// ------------------------------------------------------------------------------------------------ phase 2
struct WideLevel { uint32_t* ref; float* box; }; // per node of a wide level: binary ref, bounds [6]

struct Node80 {
    float p[3]; uint8_t e[3]; uint8_t imask; uint32_t childBase; uint32_t primBase; uint32_t valid; uint32_t pad; uint8_t qlo[3][8]; uint8_t qhi[3][8]; // BVH spec v2 (oracle/bvh.h)
};
static_assert(sizeof(Node80) == 80, "Node80");

// per wide node scratch written by k_collapse and consumed by k_emit: for each slot the binary ref + box of the entry
struct WideScratch { uint32_t* slotRef; /* [n][8], 0xFFFFFFFF = empty */ float* slotBox; /* [n][8][6] */ uint32_t* numInner; uint32_t* numTris; };

__global__ void k_collapse(WideLevel lvl, uint32_t numW, BinaryNodes bn, Node80* __restrict__ nodesOut /* level slice */, WideScratch ws) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= numW) return;
    uint32_t eref[8]; float ebox[8][6]; int n = 0;
    const uint32_t wref = lvl.ref[j];
    float wbox[6]; for (int i = 0; i < 6; ++i) wbox[i] = lvl.box[size_t(j) * 6 + i];
    if (isLeafRef(wref)) { eref[0] = wref; for (int i = 0; i < 6; ++i) ebox[0][i] = wbox[i]; n = 1; }
    else {
        for (int s = 0; s < 2; ++s) { eref[s] = bn.ref[size_t(wref) * 2 + s]; for (int i = 0; i < 6; ++i) ebox[s][i] = bn.box[(size_t(wref) * 2 + s) * 6 + i]; }
        n = 2;
    }
    while (n < 8) {
        int best = -1; float bestA = -INFINITY;
        for (int i = 0; i < n; ++i) if (!isLeafRef(eref[i])) { float a = halfArea(&ebox[i][0], &ebox[i][3]); if (a > bestA) { bestA = a; best = i; } }
        if (best < 0) break;
        const uint32_t b = eref[best];
        for (int i = n; i > best + 1; --i) { eref[i] = eref[i - 1]; for (int q = 0; q < 6; ++q) ebox[i][q] = ebox[i - 1][q]; }
        for (int s = 0; s < 2; ++s) { eref[best + s] = bn.ref[size_t(b) * 2 + s]; for (int q = 0; q < 6; ++q) ebox[best + s][q] = bn.box[(size_t(b) * 2 + s) * 6 + q]; }
        n++;
    }
    float nc[3]; for (int a = 0; a < 3; ++a) nc[a] = (wbox[a] + wbox[3 + a]) * 0.5f;
    float cost[8][8];
    for (int c = 0; c < n; ++c) {
        float off[3]; for (int a = 0; a < 3; ++a) off[a] = (ebox[c][a] + ebox[c][3 + a]) * 0.5f - nc[a];
        for (int s = 0; s < 8; ++s) {
            float sx = (s & 4) ? -1.0f : 1.0f, sy = (s & 2) ? -1.0f : 1.0f, sz = (s & 1) ? -1.0f : 1.0f;
            cost[c][s] = (sx * off[0] + sy * off[1]) + sz * off[2];
        }
    }
    int slotOf[8]; uint32_t cused = 0, sused = 0;
    for (int it = 0; it < n; ++it) {
        int bc = -1, bs = -1; float bv = -INFINITY;
        for (int c = 0; c < n; ++c) if (!(cused & (1u << c))) for (int s = 0; s < 8; ++s) if (!(sused & (1u << s))) { if (cost[c][s] > bv) { bv = cost[c][s]; bc = c; bs = s; } }
        if (bc < 0) {
            for (int c = 0; c < n && bc < 0; ++c) if (!(cused & (1u << c))) bc = c;
            for (int s = 0; s < 8 && bs < 0; ++s) if (!(sused & (1u << s))) bs = s;
        }
        cused |= 1u << bc; sused |= 1u << bs; slotOf[bc] = bs;
    }
    int entAt[8]; for (int s = 0; s < 8; ++s) entAt[s] = -1;
    for (int c = 0; c < n; ++c) entAt[slotOf[c]] = c;

    Node80 node; memset(&node, 0, sizeof(node));
    float cell[3], inv[3];
    for (int a = 0; a < 3; ++a) {
        node.p[a] = wbox[a];
        float ext = wbox[3 + a] - wbox[a];
        uint32_t bits = __float_as_uint(ext / 255.0f);
        uint32_t e = (bits >> 23) & 0xFFu;
        if (bits & 0x7FFFFFu) e += 1;
        e = min(max(e, 1u), 253u);
        if (ext * __uint_as_float((254u - e) << 23) > 255.0f) e = min(e + 1, 253u);
        node.e[a] = uint8_t(e);
        cell[a] = __uint_as_float(e << 23);
        inv[a] = __uint_as_float((254u - e) << 23);
    }
    uint32_t triOff = 0, nInner = 0;
    for (int s = 0; s < 8; ++s) {
        int c = entAt[s];
        ws.slotRef[size_t(j) * 8 + s] = c < 0 ? 0xFFFFFFFFu : eref[c];
        if (c < 0) continue;
        for (int q = 0; q < 6; ++q) ws.slotBox[(size_t(j) * 8 + s) * 6 + q] = ebox[c][q];
        for (int a = 0; a < 3; ++a) {
            float ql = floorf((ebox[c][a] - node.p[a]) * inv[a]);
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            if (ql > 0.0f && node.p[a] + ql * cell[a] > ebox[c][a]) ql -= 1.0f;
            float qh = ceilf((ebox[c][3 + a] - node.p[a]) * inv[a]);
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            if (qh < 255.0f && node.p[a] + qh * cell[a] < ebox[c][3 + a]) qh += 1.0f;
            node.qlo[a][s] = uint8_t(ql);
            node.qhi[a][s] = uint8_t(qh);
        }
        if (isLeafRef(eref[c])) {
            uint32_t cnt = (eref[c] >> 29) & 3u;
            node.valid |= ((1u << cnt) - 1u) << (3 * s);
            triOff += cnt;
        } else {
            node.imask |= uint8_t(1u << s);
            node.valid |= 1u << (24 + s);
            nInner++;
        }
    }
    nodesOut[j] = node;
    ws.numInner[j] = nInner;
    ws.numTris[j] = triOff;
}

__global__ void k_emit(uint32_t numW, Node80* __restrict__ nodesOut, WideScratch ws, const uint32_t* __restrict__ innerScan, const uint32_t* __restrict__ triScan,
                       uint32_t nextBase, uint32_t triBase, WideLevel next, const uint32_t* __restrict__ prim, const FlatTri* __restrict__ flat, FlatTri* __restrict__ trisOut) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= numW) return;
    uint32_t io = innerScan[j], to = triBase + triScan[j];
    nodesOut[j].childBase = nextBase + io;
    nodesOut[j].primBase = to;
    for (int s = 0; s < 8; ++s) {
        uint32_t ref = ws.slotRef[size_t(j) * 8 + s];
        if (ref == 0xFFFFFFFFu) continue;
        if (isLeafRef(ref)) {
            uint32_t cnt = (ref >> 29) & 3u, first = ref & 0x1FFFFFFFu;
            for (uint32_t i = 0; i < cnt; ++i) trisOut[to++] = flat[prim[first + i]];
        } else {
            next.ref[io] = ref;
            for (int q = 0; q < 6; ++q) next.box[size_t(io) * 6 + q] = ws.slotBox[(size_t(j) * 8 + s) * 6 + q];
            io++;
        }
    }
}

struct Scan {
    void* temp = nullptr; size_t tempBytes = 0;
    cudaError_t run(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st) {
        size_t need = 0;
        cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, int(n), st);
        if (e != cudaSuccess) return e;
        if (need > tempBytes) { if (temp) cudaFree(temp); e = cudaMalloc(&temp, need); if (e != cudaSuccess) return e; tempBytes = need; }
        return cub::DeviceScan::ExclusiveSum(temp, need, in, out, int(n), st);
    }
    ~Scan() { if (temp) cudaFree(temp); }
};

// Build scratch: a bump allocator over a few large device blocks (about 40 arrays are needed; one cudaMalloc / cudaFree pair per
// array cost more than all build kernels together: 265 k triangles built in 106-190 ms with per-array allocations, kernels 17 ms).
struct Pool { // frees everything on scope exit
    struct Block { char* base; size_t size, used; };
    std::vector<Block> blocks;
    size_t nextSize = size_t(64) << 20;
    template <typename Tp> cudaError_t alloc(Tp** p, size_t count) {
        const size_t bytes = (std::max<size_t>(count, 1) * sizeof(Tp) + 255) & ~size_t(255);
        if (blocks.empty() || blocks.back().used + bytes > blocks.back().size) {
            Block b; b.size = std::max(bytes, nextSize); b.used = 0; b.base = nullptr;
            cudaError_t e = cudaMalloc(&b.base, b.size);
            if (e != cudaSuccess) return e;
            blocks.push_back(b); nextSize = std::max<size_t>(b.size / 2, size_t(64) << 20);
        }
        Block& b = blocks.back();
        *p = reinterpret_cast<Tp*>(b.base + b.used); b.used += bytes;
        return cudaSuccess;
    }
    void reserve(size_t bytes) { nextSize = std::max(nextSize, bytes); }
    ~Pool() { for (Block& b : blocks) cudaFree(b.base); }
};

} // namespace

int bvhBuildDevice(vkx_ctx* ctx) {
    cudaStream_t st = ctx->stream;
    const uint32_t T = uint32_t(ctx->numFlatTris);
    if (ctx->dNodes) { cudaFree(ctx->dNodes); ctx->dNodes = nullptr; }
    if (ctx->dTris) { cudaFree(ctx->dTris); ctx->dTris = nullptr; }
    ctx->bvhBuilt = false; ctx->bvhTopology = false;
    memset(&ctx->bvh, 0, sizeof(ctx->bvh));
    cudaEvent_t e0, e1; CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
    CUDA_TRY(ctx, cudaEventRecord(e0, st));

    if (T == 0) {
        Node80 n; memset(&n, 0, sizeof(n)); n.e[0] = n.e[1] = n.e[2] = 1;
        CUDA_TRY(ctx, cudaMalloc(&ctx->dNodes, 80));
        CUDA_TRY(ctx, cudaMalloc(&ctx->dTris, 48));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dNodes, &n, 80, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        ctx->bvh.numNodes = 1; ctx->bvh.depth = 1; ctx->bvhBuilt = true; ctx->bvhTopology = true; ctx->hLevelBase = {0u, 1u};
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return VKX_OK;
    }
    if (T >= (1u << 28)) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "too many triangles (%u)", T); // bin indices (12 per primitive slot) are 32-bit

    Pool pool; Scan scan;
    pool.reserve(size_t(T) * 800 + (size_t(8) << 20)); // the arrays below: ~560 bytes per triangle for the binary phase + ~190 for the wide phase
    const unsigned B = 256;
    FlatTri* flat; float *lo, *hi, *cent; uint32_t* rootKeys;
    CUDA_TRY(ctx, pool.alloc(&flat, T)); CUDA_TRY(ctx, pool.alloc(&lo, 3 * size_t(T))); CUDA_TRY(ctx, pool.alloc(&hi, 3 * size_t(T))); CUDA_TRY(ctx, pool.alloc(&cent, 3 * size_t(T)));
    CUDA_TRY(ctx, pool.alloc(&rootKeys, 12));
    {
        uint32_t init[12] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
        CUDA_TRY(ctx, cudaMemcpyAsync(rootKeys, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    k_flatten<<<divUp(T, B), B, 0, st>>>(ctx->dVertices, ctx->dIndices, ctx->dOffsets, ctx->dInstances, ctx->dInstTriBase, uint32_t(ctx->numInstances), T, flat, lo, hi, cent, rootKeys);
    LAUNCH_CHECK(ctx);
    uint32_t hRoot[12];
    CUDA_TRY(ctx, cudaMemcpyAsync(hRoot, rootKeys, sizeof(hRoot), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float rootBox[6], rootC[6];
    for (int i = 0; i < 6; ++i) { rootBox[i] = unkey(hRoot[i]); rootC[i] = unkey(hRoot[6 + i]); }
    for (int a = 0; a < 3; ++a) { ctx->bvh.sceneMin[a] = rootBox[a]; ctx->bvh.sceneMax[a] = rootBox[3 + a]; }

    uint32_t *prim, *prim2; int *nodeOf, *nodeOf2; uint32_t *flag, *flagScan;
    CUDA_TRY(ctx, pool.alloc(&prim, T)); CUDA_TRY(ctx, pool.alloc(&prim2, T)); CUDA_TRY(ctx, pool.alloc(&nodeOf, T)); CUDA_TRY(ctx, pool.alloc(&nodeOf2, T));
    CUDA_TRY(ctx, pool.alloc(&flag, T)); CUDA_TRY(ctx, pool.alloc(&flagScan, T));
    k_iota<<<divUp(T, B), B, 0, st>>>(prim, T); LAUNCH_CHECK(ctx);

    // binary nodes: at most T-1 inner nodes
    BinaryNodes bn; CUDA_TRY(ctx, pool.alloc(&bn.ref, 2 * size_t(T))); CUDA_TRY(ctx, pool.alloc(&bn.box, 12 * size_t(T)));
    const size_t maxActive = size_t(T) / (LEAF_MAX + 1) + 1; // inner nodes of one level have > LEAF_MAX prims and are disjoint
    ActiveArrays act[2];
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(ctx, pool.alloc(&act[i].first, maxActive)); CUDA_TRY(ctx, pool.alloc(&act[i].count, maxActive)); CUDA_TRY(ctx, pool.alloc(&act[i].id, maxActive));
        CUDA_TRY(ctx, pool.alloc(&act[i].box, 6 * maxActive)); CUDA_TRY(ctx, pool.alloc(&act[i].cbox, 6 * maxActive));
    }
    BinArrays bins; CUDA_TRY(ctx, pool.alloc(&bins.cnt, maxActive * 3 * NBINS)); CUDA_TRY(ctx, pool.alloc(&bins.keys, maxActive * 3 * NBINS * 6));
    SplitArrays sp; CUDA_TRY(ctx, pool.alloc(&sp.axis, maxActive)); CUDA_TRY(ctx, pool.alloc(&sp.plane, maxActive)); CUDA_TRY(ctx, pool.alloc(&sp.nL, maxActive)); CUDA_TRY(ctx, pool.alloc(&sp.innerFlag, 2 * maxActive + 1));
    uint32_t* innerScan; CUDA_TRY(ctx, pool.alloc(&innerScan, 2 * maxActive + 1));
    uint32_t* childKeys; CUDA_TRY(ctx, pool.alloc(&childKeys, 2 * maxActive * 12));

    uint32_t rootRef, numBinary = 0, numActive = 0;
    if (T <= LEAF_MAX) { rootRef = leafRef(0, T); k_fill_i32<<<divUp(T, B), B, 0, st>>>(nodeOf, -1, T); LAUNCH_CHECK(ctx); }
    else {
        rootRef = 0; numBinary = 1; numActive = 1;
        uint32_t f = 0, c = T, id = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(act[0].first, &f, 4, cudaMemcpyHostToDevice, st)); CUDA_TRY(ctx, cudaMemcpyAsync(act[0].count, &c, 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(act[0].id, &id, 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(act[0].box, rootBox, 24, cudaMemcpyHostToDevice, st)); CUDA_TRY(ctx, cudaMemcpyAsync(act[0].cbox, rootC, 24, cudaMemcpyHostToDevice, st));
        k_fill_i32<<<divUp(T, B), B, 0, st>>>(nodeOf, 0, T); LAUNCH_CHECK(ctx);
    }
    int cur = 0, levels = 0;
    while (numActive > 0) {
        if (++levels > 4096) return vkx_fail(ctx, VKX_E_INVALID, "BVH build did not terminate");
        ActiveArrays& A = act[cur]; ActiveArrays& N = act[cur ^ 1];
        k_init_bins<<<divUp(size_t(numActive) * 3 * NBINS, B), B, 0, st>>>(bins, numActive); LAUNCH_CHECK(ctx);
        k_bin<<<divUp(T, B), B, 0, st>>>(prim, nodeOf, T, A, bins, lo, hi, cent); LAUNCH_CHECK(ctx);
        k_split<<<divUp(numActive, 128), 128, 0, st>>>(A, bins, sp, numActive); LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemsetAsync(sp.innerFlag + 2 * size_t(numActive), 0, 4, st));
        CUDA_TRY(ctx, scan.run(sp.innerFlag, innerScan, 2 * size_t(numActive) + 1, st)); ctx->launches += 2;
        k_left_flags<<<divUp(T, B), B, 0, st>>>(prim, nodeOf, T, A, sp, cent, flag); LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, scan.run(flag, flagScan, T, st)); ctx->launches += 2;
        k_init_child_keys<<<divUp(size_t(numActive) * 24, B), B, 0, st>>>(childKeys, 2 * numActive); LAUNCH_CHECK(ctx);
        k_scatter<<<divUp(T, B), B, 0, st>>>(prim, prim2, nodeOf, nodeOf2, T, A, sp, flag, flagScan, innerScan, childKeys, lo, hi, cent); LAUNCH_CHECK(ctx);
        k_finalize_children<<<divUp(2 * size_t(numActive), B), B, 0, st>>>(A, N, sp, innerScan, childKeys, bn, numActive, numBinary); LAUNCH_CHECK(ctx);
        uint32_t numNext = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(&numNext, innerScan + 2 * size_t(numActive), 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        numBinary += numNext; numActive = numNext;
        std::swap(prim, prim2); std::swap(nodeOf, nodeOf2); cur ^= 1;
    }
    ctx->bvh.numBinaryNodes = numBinary;

    // ---- phase 2: wide levels
    const size_t maxWide = size_t(numBinary) + 1;
    Node80* nodesTmp; FlatTri* trisOut;
    CUDA_TRY(ctx, pool.alloc(&nodesTmp, maxWide));
    CUDA_TRY(ctx, cudaMalloc(&trisOut, size_t(T) * sizeof(FlatTri)));
    ctx->dTris = reinterpret_cast<float4*>(trisOut);
    WideLevel wl[2];
    for (int i = 0; i < 2; ++i) { CUDA_TRY(ctx, pool.alloc(&wl[i].ref, maxWide)); CUDA_TRY(ctx, pool.alloc(&wl[i].box, 6 * maxWide)); }
    WideScratch ws; CUDA_TRY(ctx, pool.alloc(&ws.slotRef, 8 * maxWide)); CUDA_TRY(ctx, pool.alloc(&ws.slotBox, 48 * maxWide));
    CUDA_TRY(ctx, pool.alloc(&ws.numInner, maxWide + 1)); CUDA_TRY(ctx, pool.alloc(&ws.numTris, maxWide + 1));
    uint32_t *wInnerScan, *wTriScan; CUDA_TRY(ctx, pool.alloc(&wInnerScan, maxWide + 1)); CUDA_TRY(ctx, pool.alloc(&wTriScan, maxWide + 1));
    CUDA_TRY(ctx, cudaMemcpyAsync(wl[0].ref, &rootRef, 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(wl[0].box, rootBox, 24, cudaMemcpyHostToDevice, st));
    uint32_t numW = 1, levelBase = 0, triBase = 0, depth = 0; cur = 0;
    ctx->hLevelBase.clear();
    while (numW > 0) {
        depth++;
        ctx->hLevelBase.push_back(levelBase);
        if (size_t(levelBase) + numW > maxWide) return vkx_fail(ctx, VKX_E_INVALID, "wide node overflow");
        k_collapse<<<divUp(numW, 64), 64, 0, st>>>(wl[cur], numW, bn, nodesTmp + levelBase, ws); LAUNCH_CHECK(ctx);
        // numInner / numTris have one extra slot so the exclusive scan's last element is the total
        CUDA_TRY(ctx, cudaMemsetAsync(ws.numInner + numW, 0, 4, st)); CUDA_TRY(ctx, cudaMemsetAsync(ws.numTris + numW, 0, 4, st));
        CUDA_TRY(ctx, scan.run(ws.numInner, wInnerScan, size_t(numW) + 1, st)); CUDA_TRY(ctx, scan.run(ws.numTris, wTriScan, size_t(numW) + 1, st)); ctx->launches += 4;
        const uint32_t nextBase = levelBase + numW;
        k_emit<<<divUp(numW, 64), 64, 0, st>>>(numW, nodesTmp + levelBase, ws, wInnerScan, wTriScan, nextBase, triBase, wl[cur ^ 1], prim, flat, trisOut); LAUNCH_CHECK(ctx);
        uint32_t tot[2];
        CUDA_TRY(ctx, cudaMemcpyAsync(&tot[0], wInnerScan + numW, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(&tot[1], wTriScan + numW, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        levelBase = nextBase; triBase += tot[1]; numW = tot[0]; cur ^= 1;
    }
    if (triBase != T) return vkx_fail(ctx, VKX_E_INVALID, "BVH build lost triangles (%u of %u)", triBase, T);
    if (depth > 44) return vkx_fail(ctx, VKX_E_UNSUPPORTED, "BVH too deep (%u levels)", depth);
    CUDA_TRY(ctx, cudaMalloc(&ctx->dNodes, size_t(levelBase) * 80));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dNodes, nodesTmp, size_t(levelBase) * 80, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(ctx, cudaEventRecord(e1, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->hLevelBase.push_back(levelBase);
    ctx->bvh.numNodes = levelBase; ctx->bvh.numTriangles = T; ctx->bvh.depth = depth; ctx->bvh.buildMs = ms;
    ctx->bvhBuilt = true; ctx->bvhTopology = true;
    return VKX_OK;
}

// ------------------------------------------------------------------------------------------------ refit
// Topology-preserving refit = Renderer::updateAccelerationStructureInstances + updateTLAS (reference src/Renderer.cpp:671-742: new
// instance transforms, then vkCmdBuildAccelerationStructuresKHR in UPDATE mode) on the single-level wide BVH: tree, slot assignment
// and triangle order stay; the world-space triangles are recomputed from the current vertex arena and instance transforms (so it
// also serves skinned meshes, src/Renderer.cpp:644-669), and origin / exponents / child planes of every node are re-quantised from
// the new bounds, one launch per level from the leaves up. Defined by oracle/bvh.cpp::refit; nodes and triangles are byte-identical
// to it (tests/test_bvh_parity.py).
namespace {

__global__ void k_refit_tris(const vkx_vertex* __restrict__ vertices, const uint32_t* __restrict__ indices, const vkx_offset_entry* __restrict__ offsets,
                             const vkx_instance* __restrict__ instances, uint32_t T, FlatTri* __restrict__ tris, float* __restrict__ triBox) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T) return;
    const uint32_t k = tris[g].inst & 0x00FFFFFFu, j = tris[g].prim & 0x7FFFFFFFu;
    const vkx_instance in = instances[k];
    const vkx_offset_entry oe = offsets[in.meshEntry];
    const float* M = in.transform;
    const float det = M[0] * (M[5] * M[10] - M[6] * M[9]) - M[1] * (M[4] * M[10] - M[6] * M[8]) + M[2] * (M[4] * M[9] - M[5] * M[8]);
    const uint32_t flip = det < 0.0f ? 0x80000000u : 0u;
    float w[3][3];
    for (int c = 0; c < 3; ++c) {
        const uint32_t vi = oe.vertexOffset + indices[oe.indexOffset + 3 * j + c];
        const float* p = vertices[vi].pos;
        const float x = p[0], y = p[1], z = p[2];
        for (int r = 0; r < 3; ++r) w[c][r] = ((M[4 * r + 0] * x + M[4 * r + 1] * y) + M[4 * r + 2] * z) + M[4 * r + 3];
    }
    FlatTri t;
    for (int ax = 0; ax < 3; ++ax) {
        t.v0[ax] = w[0][ax];
        t.e1[ax] = w[1][ax] - w[0][ax];
        t.e2[ax] = w[2][ax] - w[0][ax];
        triBox[6 * size_t(g) + ax] = kmin(kmin(w[0][ax], w[1][ax]), w[2][ax]);
        triBox[6 * size_t(g) + 3 + ax] = kmax(kmax(w[0][ax], w[1][ax]), w[2][ax]);
    }
    t.inst = k | ((in.mask & 0xFFu) << 24);
    t.prim = j | flip;
    t.pad = 0;
    tris[g] = t;
}

__global__ void k_refit_level(uint32_t base, uint32_t count, Node80* __restrict__ nodes, const float* __restrict__ triBox, float* __restrict__ nodeBox) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const uint32_t n = base + j;
    Node80 node = nodes[n];
    float sb[8][6]; bool present[8];
    float nb[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int s = 0; s < 8; ++s) {
        present[s] = false;
        if (node.imask & (1u << s)) {
            const uint32_t c = node.childBase + uint32_t(__popc(node.imask & ((1u << s) - 1u)));
            for (int q = 0; q < 6; ++q) sb[s][q] = nodeBox[6 * size_t(c) + q];
            present[s] = true;
        } else {
            const uint32_t cnt = uint32_t(__popc((node.valid >> (3 * s)) & 7u));
            if (cnt) {
                const uint32_t first = node.primBase + uint32_t(__popc(node.valid & 0x00FFFFFFu & ((1u << (3 * s)) - 1u)));
                for (int a = 0; a < 3; ++a) { sb[s][a] = INFINITY; sb[s][3 + a] = -INFINITY; }
                for (uint32_t i = 0; i < cnt; ++i)
                    for (int a = 0; a < 3; ++a) { sb[s][a] = kmin(sb[s][a], triBox[6 * size_t(first + i) + a]); sb[s][3 + a] = kmax(sb[s][3 + a], triBox[6 * size_t(first + i) + 3 + a]); }
                present[s] = true;
            }
        }
        if (present[s]) for (int a = 0; a < 3; ++a) { nb[a] = kmin(nb[a], sb[s][a]); nb[3 + a] = kmax(nb[3 + a], sb[s][3 + a]); }
    }
    for (int q = 0; q < 6; ++q) nodeBox[6 * size_t(n) + q] = nb[q];
    float cell[3], inv[3];
    for (int a = 0; a < 3; ++a) { // spec section 3.4, the arithmetic of k_collapse
        node.p[a] = nb[a];
        const float ext = nb[3 + a] - nb[a];
        const uint32_t bits = __float_as_uint(ext / 255.0f);
        uint32_t e = (bits >> 23) & 0xFFu;
        if (bits & 0x7FFFFFu) e += 1;
        e = min(max(e, 1u), 253u);
        if (ext * __uint_as_float((254u - e) << 23) > 255.0f) e = min(e + 1, 253u);
        node.e[a] = uint8_t(e);
        cell[a] = __uint_as_float(e << 23);
        inv[a] = __uint_as_float((254u - e) << 23);
    }
    for (int s = 0; s < 8; ++s) {
        for (int a = 0; a < 3; ++a) { node.qlo[a][s] = 0; node.qhi[a][s] = 0; }
        if (!present[s]) continue;
        for (int a = 0; a < 3; ++a) {
            float ql = floorf((sb[s][a] - node.p[a]) * inv[a]);
            ql = fminf(fmaxf(ql, 0.0f), 255.0f);
            if (ql > 0.0f && node.p[a] + ql * cell[a] > sb[s][a]) ql -= 1.0f;
            float qh = ceilf((sb[s][3 + a] - node.p[a]) * inv[a]);
            qh = fminf(fmaxf(qh, 0.0f), 255.0f);
            if (qh < 255.0f && node.p[a] + qh * cell[a] < sb[s][3 + a]) qh += 1.0f;
            node.qlo[a][s] = uint8_t(ql);
            node.qhi[a][s] = uint8_t(qh);
        }
    }
    nodes[n] = node;
}

} // namespace

int bvhRefitDevice(vkx_ctx* ctx) {
    cudaStream_t st = ctx->stream;
    const uint32_t T = uint32_t(ctx->bvh.numTriangles), N = ctx->bvh.numNodes;
    if (T == 0) { ctx->bvhBuilt = true; return VKX_OK; }
    cudaEvent_t e0, e1; CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
    CUDA_TRY(ctx, cudaEventRecord(e0, st));
    const size_t need = (size_t(T) + N) * 6 * sizeof(float);
    if (need > ctx->refitScratchBytes) { // kept between refits: a per-frame operation
        if (ctx->dRefitScratch) cudaFree(ctx->dRefitScratch);
        ctx->dRefitScratch = nullptr; ctx->refitScratchBytes = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->dRefitScratch, need)); ctx->refitScratchBytes = need;
    }
    float* triBox = ctx->dRefitScratch; float* nodeBox = triBox + size_t(T) * 6;
    Node80* nodes = reinterpret_cast<Node80*>(ctx->dNodes);
    k_refit_tris<<<divUp(T, 256), 256, 0, st>>>(ctx->dVertices, ctx->dIndices, ctx->dOffsets, ctx->dInstances, T, reinterpret_cast<FlatTri*>(ctx->dTris), triBox); LAUNCH_CHECK(ctx);
    for (size_t l = ctx->hLevelBase.size() - 1; l-- > 0;) {
        const uint32_t base = ctx->hLevelBase[l], count = ctx->hLevelBase[l + 1] - base;
        k_refit_level<<<divUp(count, 64), 64, 0, st>>>(base, count, nodes, triBox, nodeBox); LAUNCH_CHECK(ctx);
    }
    float root[6];
    CUDA_TRY(ctx, cudaMemcpyAsync(root, nodeBox, sizeof(root), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaEventRecord(e1, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int a = 0; a < 3; ++a) { ctx->bvh.sceneMin[a] = root[a]; ctx->bvh.sceneMax[a] = root[3 + a]; }
    ctx->bvh.buildMs = ms;
    ctx->bvhBuilt = true;
    return VKX_OK;
}
