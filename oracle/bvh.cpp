// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
// See bvh.h. Sequential definition of BVH spec v1.
#include "bvh.h"
#include <cmath>
#include <cstring>
#include <limits>
#include <algorithm>

namespace obvh {

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// Total order on floats (-0 < +0), the same order the CUDA builder gets from atomicMin/Max on mapped keys.
static inline uint32_t okey(float f) { uint32_t b = f2u(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
static inline float omin(float a, float b) { return okey(a) <= okey(b) ? a : b; }
static inline float omax(float a, float b) { return okey(a) >= okey(b) ? a : b; }

static const float INF = std::numeric_limits<float>::infinity();

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = INF; hi[a] = -INF; } }
    void grow(const float* l, const float* h) { for (int a = 0; a < 3; ++a) { lo[a] = omin(lo[a], l[a]); hi[a] = omax(hi[a], h[a]); } }
    void grow(const Box& b) { grow(b.lo, b.hi); }
};
// Half surface area, fixed operation order: (ex*ey + ey*ez) + ez*ex
static inline float halfArea(const float* lo, const float* hi) {
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return (ex * ey + ey * ez) + ez * ex;
}

// One instanced triangle in world space (spec section 3.1): vertices through the instance's 3x4 transform without FMA, edges from
// the transformed vertices, bounds by the total order on floats.
static void flattenTriangle(const vkx_vertex* vertices, const uint32_t* indices, const vkx_offset_entry& oe, const vkx_instance& in, uint32_t k, uint32_t j, uint32_t flip,
                            Tri48& t, float lo[3], float hi[3]) {
    const float* M = in.transform;
    float w[3][3];
    for (int c = 0; c < 3; ++c) {
        uint32_t vi = oe.vertexOffset + indices[oe.indexOffset + 3 * j + c];
        const float* p = vertices[vi].pos;
        for (int r = 0; r < 3; ++r)
            w[c][r] = ((M[4 * r + 0] * p[0] + M[4 * r + 1] * p[1]) + M[4 * r + 2] * p[2]) + M[4 * r + 3];
    }
    for (int a = 0; a < 3; ++a) {
        t.v0[a] = w[0][a];
        t.e1[a] = w[1][a] - w[0][a];
        t.e2[a] = w[2][a] - w[0][a];
        lo[a] = omin(omin(w[0][a], w[1][a]), w[2][a]);
        hi[a] = omax(omax(w[0][a], w[1][a]), w[2][a]);
    }
    t.inst = k | ((in.mask & 0xFFu) << 24);
    t.prim = j | flip;
    t.pad = 0;
}
static inline uint32_t windingFlip(const float* M) {
    float det = M[0] * (M[5] * M[10] - M[6] * M[9]) - M[1] * (M[4] * M[10] - M[6] * M[8]) + M[2] * (M[4] * M[9] - M[5] * M[8]);
    return det < 0.0f ? 0x80000000u : 0u;
}

void flatten(const vkx_vertex* vertices, const uint32_t* indices, const vkx_offset_entry* offsets,
             const uint32_t* meshIndexCounts, const vkx_instance* instances, size_t numInstances,
             std::vector<Tri48>& out, std::vector<float>& lo, std::vector<float>& hi) {
    out.clear(); lo.clear(); hi.clear();
    for (size_t k = 0; k < numInstances; ++k) {
        const vkx_instance& in = instances[k];
        const vkx_offset_entry& oe = offsets[in.meshEntry];
        const float* M = in.transform;
        uint32_t flip = windingFlip(M);
        uint32_t ntri = meshIndexCounts[in.meshEntry] / 3;
        for (uint32_t j = 0; j < ntri; ++j) {
            Tri48 t; float l[3], h[3];
            flattenTriangle(vertices, indices, oe, in, uint32_t(k), j, flip, t, l, h);
            for (int a = 0; a < 3; ++a) { lo.push_back(l[a]); hi.push_back(h[a]); }
            out.push_back(t);
        }
    }
}

namespace {

constexpr int NBINS = 16;
constexpr uint32_t LEAF_MAX = 3;
constexpr uint32_t LEAF_FLAG = 0x80000000u;
inline uint32_t leafRef(uint32_t first, uint32_t count) { return LEAF_FLAG | (count << 29) | first; }
inline bool isLeaf(uint32_t ref) { return (ref & LEAF_FLAG) != 0; }
inline uint32_t leafFirst(uint32_t ref) { return ref & 0x1FFFFFFFu; }
inline uint32_t leafCount(uint32_t ref) { return (ref >> 29) & 3u; }

struct BNode { uint32_t ref[2]; Box box[2]; };
struct Active { uint32_t first, count, id; Box box, cbox; };

} // namespace

// Quantisation origin / exponents of a wide node from its box, and the 8-bit child planes of the occupied slots (spec section 3.4):
// shared by the build and the refit, which must produce the same bytes for the same boxes.
static void quantiseNode(Node80& node, const Box& nodeBox, const Box* slotBox, const bool* present) {
    float cell[3], inv[3];
    for (int a = 0; a < 3; ++a) {
        node.p[a] = nodeBox.lo[a];
        float ext = nodeBox.hi[a] - nodeBox.lo[a];
        uint32_t bits = f2u(ext / 255.0f);
        uint32_t e = (bits >> 23) & 0xFFu;
        if (bits & 0x7FFFFFu) e += 1;
        e = std::min(std::max(e, 1u), 253u);
        if (ext * u2f((254u - e) << 23) > 255.0f) e = std::min(e + 1, 253u);
        node.e[a] = uint8_t(e);
        cell[a] = u2f(e << 23);
        inv[a] = u2f((254u - e) << 23);
    }
    for (int s = 0; s < 8; ++s) {
        for (int a = 0; a < 3; ++a) { node.qlo[a][s] = 0; node.qhi[a][s] = 0; }
        if (!present[s]) continue;
        const Box& b = slotBox[s];
        for (int a = 0; a < 3; ++a) {
            float ql = std::floor((b.lo[a] - node.p[a]) * inv[a]);
            ql = std::min(std::max(ql, 0.0f), 255.0f);
            if (ql > 0.0f && node.p[a] + ql * cell[a] > b.lo[a]) ql -= 1.0f;
            float qh = std::ceil((b.hi[a] - node.p[a]) * inv[a]);
            qh = std::min(std::max(qh, 0.0f), 255.0f);
            if (qh < 255.0f && node.p[a] + qh * cell[a] < b.hi[a]) qh += 1.0f;
            node.qlo[a][s] = uint8_t(ql);
            node.qhi[a][s] = uint8_t(qh);
        }
    }
}

void build(const std::vector<Tri48>& flat, const std::vector<float>& lo, const std::vector<float>& hi, Bvh& out) {
    const uint32_t T = uint32_t(flat.size());
    out.nodes.clear(); out.tris.clear(); out.numBinaryNodes = 0; out.depth = 1;
    if (T == 0) {
        Node80 n; std::memset(&n, 0, sizeof(n)); n.e[0] = n.e[1] = n.e[2] = 1;
        out.nodes.push_back(n);
        return;
    }
    std::vector<float> cent(3 * size_t(T));
    for (size_t i = 0; i < 3 * size_t(T); ++i) cent[i] = (lo[i] + hi[i]) * 0.5f;
    std::vector<uint32_t> prim(T), prim2(T);
    for (uint32_t i = 0; i < T; ++i) prim[i] = i;

    Box rootBox, rootC; rootBox.reset(); rootC.reset();
    for (uint32_t i = 0; i < T; ++i) { rootBox.grow(&lo[3 * i], &hi[3 * i]); rootC.grow(&cent[3 * i], &cent[3 * i]); }
    for (int a = 0; a < 3; ++a) { out.sceneMin[a] = rootBox.lo[a]; out.sceneMax[a] = rootBox.hi[a]; }

    // ---------------- phase 1: breadth-first binned-SAH binary tree ----------------
    std::vector<BNode> bn;
    uint32_t rootRef;
    std::vector<Active> cur, next;
    if (T <= LEAF_MAX) rootRef = leafRef(0, T);
    else { rootRef = 0; bn.emplace_back(); cur.push_back(Active{0, T, 0, rootBox, rootC}); }

    std::vector<uint8_t> side;
    while (!cur.empty()) {
        next.clear();
        prim2 = prim;
        for (const Active& n : cur) {
            // binning
            int bestAxis = -1, bestPlane = -1; float bestCost = INF;
            for (int a = 0; a < 3; ++a) {
                float ext = n.cbox.hi[a] - n.cbox.lo[a];
                if (!(ext > 0.0f)) continue;
                float k = float(NBINS) / ext;
                uint32_t cnt[NBINS]; Box bb[NBINS];
                for (int b = 0; b < NBINS; ++b) { cnt[b] = 0; bb[b].reset(); }
                for (uint32_t i = 0; i < n.count; ++i) {
                    uint32_t g = prim[n.first + i];
                    int b = std::min(NBINS - 1, int((cent[3 * g + a] - n.cbox.lo[a]) * k));
                    cnt[b]++; bb[b].grow(&lo[3 * g], &hi[3 * g]);
                }
                Box L[NBINS - 1]; uint32_t nL[NBINS - 1];
                Box acc; acc.reset(); uint32_t c = 0;
                for (int s = 0; s < NBINS - 1; ++s) { if (cnt[s]) acc.grow(bb[s]); c += cnt[s]; L[s] = acc; nL[s] = c; }
                Box R[NBINS - 1]; uint32_t nR[NBINS - 1];
                acc.reset(); c = 0;
                for (int s = NBINS - 2; s >= 0; --s) { if (cnt[s + 1]) acc.grow(bb[s + 1]); c += cnt[s + 1]; R[s] = acc; nR[s] = c; }
                for (int s = 0; s < NBINS - 1; ++s) {
                    if (nL[s] == 0 || nR[s] == 0) continue;
                    float cost = halfArea(L[s].lo, L[s].hi) * float(nL[s]) + halfArea(R[s].lo, R[s].hi) * float(nR[s]);
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestPlane = s; }
                }
            }
            // stable partition
            side.assign(n.count, 0);
            uint32_t nl = 0;
            if (bestAxis >= 0) {
                float k = float(NBINS) / (n.cbox.hi[bestAxis] - n.cbox.lo[bestAxis]);
                for (uint32_t i = 0; i < n.count; ++i) {
                    uint32_t g = prim[n.first + i];
                    int b = std::min(NBINS - 1, int((cent[3 * g + bestAxis] - n.cbox.lo[bestAxis]) * k));
                    side[i] = b <= bestPlane ? 0 : 1;
                    nl += side[i] == 0;
                }
            } else { // all centroids coincide: median split in current order
                nl = n.count / 2;
                for (uint32_t i = 0; i < n.count; ++i) side[i] = i < nl ? 0 : 1;
            }
            uint32_t wl = n.first, wr = n.first + nl;
            for (uint32_t i = 0; i < n.count; ++i) { if (side[i] == 0) prim2[wl++] = prim[n.first + i]; else prim2[wr++] = prim[n.first + i]; }
            // children
            uint32_t cf[2] = {n.first, n.first + nl}, cc[2] = {nl, n.count - nl};
            for (int s = 0; s < 2; ++s) {
                Box b, cb; b.reset(); cb.reset();
                for (uint32_t i = 0; i < cc[s]; ++i) { uint32_t g = prim2[cf[s] + i]; b.grow(&lo[3 * g], &hi[3 * g]); cb.grow(&cent[3 * g], &cent[3 * g]); }
                bn[n.id].box[s] = b;
                if (cc[s] <= LEAF_MAX) bn[n.id].ref[s] = leafRef(cf[s], cc[s]);
                else {
                    uint32_t id = uint32_t(bn.size()); bn.emplace_back();
                    bn[n.id].ref[s] = id;
                    next.push_back(Active{cf[s], cc[s], id, b, cb});
                }
            }
        }
        prim.swap(prim2);
        cur.swap(next);
    }
    out.numBinaryNodes = uint32_t(bn.size());

    // ---------------- phase 2+3: greedy collapse to 8-wide, slot assignment, quantisation ----------------
    struct Entry { uint32_t ref; Box box; };
    struct WideWork { uint32_t bref; Box box; };
    std::vector<WideWork> wcur, wnext;
    wcur.push_back(WideWork{rootRef, rootBox});
    uint32_t levelBase = 0; out.depth = 0;
    while (!wcur.empty()) {
        out.depth++;
        wnext.clear();
        uint32_t nextBase = levelBase + uint32_t(wcur.size());
        for (const WideWork& w : wcur) {
            Entry ent[8]; int n = 0;
            if (isLeaf(w.bref)) { ent[n++] = Entry{w.bref, w.box}; } // only the root of a <=3-triangle scene
            else { ent[n++] = Entry{bn[w.bref].ref[0], bn[w.bref].box[0]}; ent[n++] = Entry{bn[w.bref].ref[1], bn[w.bref].box[1]}; }
            while (n < 8) {
                int best = -1; float bestA = -INF;
                for (int i = 0; i < n; ++i) if (!isLeaf(ent[i].ref)) { float a = halfArea(ent[i].box.lo, ent[i].box.hi); if (a > bestA) { bestA = a; best = i; } }
                if (best < 0) break;
                const BNode& b = bn[ent[best].ref];
                for (int i = n; i > best + 1; --i) ent[i] = ent[i - 1];
                ent[best] = Entry{b.ref[0], b.box[0]};
                ent[best + 1] = Entry{b.ref[1], b.box[1]};
                n++;
            }
            // slot assignment (octant order heuristic)
            float nc[3]; for (int a = 0; a < 3; ++a) nc[a] = (w.box.lo[a] + w.box.hi[a]) * 0.5f;
            float cost[8][8];
            for (int c = 0; c < n; ++c) {
                float off[3]; for (int a = 0; a < 3; ++a) off[a] = (ent[c].box.lo[a] + ent[c].box.hi[a]) * 0.5f - nc[a];
                for (int s = 0; s < 8; ++s) {
                    float sx = (s & 4) ? -1.0f : 1.0f, sy = (s & 2) ? -1.0f : 1.0f, sz = (s & 1) ? -1.0f : 1.0f;
                    cost[c][s] = (sx * off[0] + sy * off[1]) + sz * off[2];
                }
            }
            int slotOf[8]; bool cused[8] = {false, false, false, false, false, false, false, false}, sused[8] = {false, false, false, false, false, false, false, false};
            for (int it = 0; it < n; ++it) {
                int bc = -1, bs = -1; float bv = -INF;
                for (int c = 0; c < n; ++c) if (!cused[c]) for (int s = 0; s < 8; ++s) if (!sused[s]) { if (cost[c][s] > bv) { bv = cost[c][s]; bc = c; bs = s; } }
                if (bc < 0) { // all remaining costs are -inf/NaN: first free pair
                    for (int c = 0; c < n && bc < 0; ++c) if (!cused[c]) bc = c;
                    for (int s = 0; s < 8 && bs < 0; ++s) if (!sused[s]) bs = s;
                }
                cused[bc] = true; sused[bs] = true; slotOf[bc] = bs;
            }
            int entAt[8]; for (int s = 0; s < 8; ++s) entAt[s] = -1;
            for (int c = 0; c < n; ++c) entAt[slotOf[c]] = c;

            Node80 node; std::memset(&node, 0, sizeof(node));
            Box slotBox[8]; bool present[8];
            for (int s = 0; s < 8; ++s) { present[s] = entAt[s] >= 0; if (present[s]) slotBox[s] = ent[entAt[s]].box; }
            quantiseNode(node, w.box, slotBox, present);
            node.childBase = nextBase + uint32_t(wnext.size());
            node.primBase = uint32_t(out.tris.size());
            for (int s = 0; s < 8; ++s) {
                int c = entAt[s];
                if (c < 0) continue;
                const Entry& en = ent[c];
                if (isLeaf(en.ref)) {
                    uint32_t cnt = leafCount(en.ref), first = leafFirst(en.ref);
                    node.valid |= ((1u << cnt) - 1u) << (3 * s);
                    for (uint32_t i = 0; i < cnt; ++i) out.tris.push_back(flat[prim[first + i]]);
                } else {
                    node.imask |= uint8_t(1u << s);
                    node.valid |= 1u << (24 + s);
                    wnext.push_back(WideWork{en.ref, en.box});
                }
            }
            out.nodes.push_back(node);
        }
        levelBase = nextBase;
        wcur.swap(wnext);
    }
}

// Topology-preserving refit (the reference refits its TLAS in place after instance transforms change, src/Renderer.cpp:681-742): the
// tree, the slot assignment and the triangle order stay; the world-space triangles are recomputed from the current vertices and
// instance transforms, and every node's origin, exponents and child planes are re-quantised from the new bounds, children before
// parents (nodes are stored level by level, so descending index order is bottom-up). With unchanged transforms this reproduces the
// built structure byte for byte; with moved instances the result differs from a rebuild (the split decisions are the old ones) but
// is still a valid, conservative hierarchy over the same triangles.
void refit(const vkx_vertex* vertices, const uint32_t* indices, const vkx_offset_entry* offsets, const vkx_instance* instances, Bvh& bvh) {
    const size_t T = bvh.tris.size(), N = bvh.nodes.size();
    std::vector<Box> triBox(T), nodeBox(N);
    for (size_t i = 0; i < T; ++i) {
        Tri48& t = bvh.tris[i];
        const uint32_t k = t.inst & 0x00FFFFFFu, j = t.prim & 0x7FFFFFFFu;
        const vkx_instance& in = instances[k];
        flattenTriangle(vertices, indices, offsets[in.meshEntry], in, k, j, windingFlip(in.transform), t, triBox[i].lo, triBox[i].hi);
    }
    for (size_t n = N; n-- > 0;) {
        Node80& node = bvh.nodes[n];
        Box slotBox[8]; bool present[8]; Box nb; nb.reset();
        for (int s = 0; s < 8; ++s) {
            present[s] = false;
            if (node.imask & (1u << s)) {
                slotBox[s] = nodeBox[node.childBase + uint32_t(__builtin_popcount(node.imask & ((1u << s) - 1u)))];
                present[s] = true;
            } else {
                const uint32_t cnt = uint32_t(__builtin_popcount((node.valid >> (3 * s)) & 7u));
                if (cnt) {
                    const uint32_t first = node.primBase + uint32_t(__builtin_popcount(node.valid & 0x00FFFFFFu & ((1u << (3 * s)) - 1u)));
                    slotBox[s].reset();
                    for (uint32_t i = 0; i < cnt; ++i) slotBox[s].grow(triBox[first + i]);
                    present[s] = true;
                }
            }
            if (present[s]) nb.grow(slotBox[s]);
        }
        nodeBox[n] = nb;
        quantiseNode(node, nb, slotBox, present);
    }
    if (N) for (int a = 0; a < 3; ++a) { bvh.sceneMin[a] = nodeBox[0].lo[a]; bvh.sceneMax[a] = nodeBox[0].hi[a]; }
}

// ------------------------------------------------------------------------------------------------------------
// Traversal (spec section 3.4). Every float operation below is a single correctly-rounded IEEE op in a fixed
// order; the CUDA kernels use the matching __f*_rn intrinsics.
int gMaxStack = 0; // deepest traversal stack seen so far (diagnostic: sizes the device-side shared-memory stack; benign race)
namespace {

struct RayCtx {
    float o[3], d[3], idir[3];
    uint32_t oct;
};

inline void setupRay(RayCtx& r, const float o[3], const float d[3]) {
    r.oct = 0;
    for (int a = 0; a < 3; ++a) {
        r.o[a] = o[a];
        r.d[a] = d[a];
        float dd = std::fabs(d[a]) < 1e-20f ? std::copysign(1e-20f, d[a]) : d[a];
        r.idir[a] = 1.0f / dd;
        if (dd < 0.0f) r.oct |= (4u >> a);
    }
}

// Returns the 32-bit hit mask of one node: bits 24..31 inner children by slot, bits 0..23 triangles (3 per slot).
inline uint32_t intersectNode(const Node80& n, const RayCtx& r, float tmin, float tmax) {
    float ax[3], bx[3];
    for (int a = 0; a < 3; ++a) {
        ax[a] = u2f(uint32_t(n.e[a]) << 23) * r.idir[a];
        bx[a] = (n.p[a] - r.o[a]) * r.idir[a];
    }
    uint32_t mask = 0;
    for (int s = 0; s < 8; ++s) {
        float tl[3], th[3];
        for (int a = 0; a < 3; ++a) {
            bool neg = (r.oct & (4u >> a)) != 0;
            float qn = float(neg ? n.qhi[a][s] : n.qlo[a][s]);
            float qf = float(neg ? n.qlo[a][s] : n.qhi[a][s]);
            tl[a] = std::fmaf(qn, ax[a], bx[a]);
            th[a] = std::fmaf(qf, ax[a], bx[a]);
        }
        float tn = std::fmax(std::fmax(tl[0], tl[1]), std::fmax(tl[2], tmin));
        float tf = std::fmin(std::fmin(th[0], th[1]), std::fmin(th[2], tmax));
        if (tn <= tf) mask |= (7u << (3 * s)) | (1u << (24 + s));
    }
    return mask & n.valid; // empty slots and absent triangles have no bit in `valid`
}

// Moeller-Trumbore, fixed op order. Returns true if the triangle plane/edges are hit; t,u,v,det out.
// Spec v2.1: the barycentric acceptance tests compare the NUMERATORS (sign of det folded in by an exact sign flip) with |det|
// (0 <= un <= |det|, 0 <= vn, un + vn <= |det|), so rejected candidates - most of them - never pay the division; accepted ones get
// u, v, t = numerator * (1 / det) as before. (v2 tested the rounded quotients u <= 1, u + v <= 1: the two differ only where a
// quotient rounds across 1.)
inline float flipBy(float x, float s) { return u2f(f2u(x) ^ (f2u(s) & 0x80000000u)); } // x if s >= +0, -x if s carries the sign bit
inline bool intersectTri(const Tri48& tr, const RayCtx& r, float& t, float& u, float& v, float& det) {
    const float* d = r.d; const float* e1 = tr.e1; const float* e2 = tr.e2;
    float px = std::fmaf(d[1], e2[2], -(d[2] * e2[1]));
    float py = std::fmaf(d[2], e2[0], -(d[0] * e2[2]));
    float pz = std::fmaf(d[0], e2[1], -(d[1] * e2[0]));
    det = std::fmaf(e1[0], px, std::fmaf(e1[1], py, e1[2] * pz));
    if (det == 0.0f) return false;
    const float ad = std::fabs(det);
    float tx = r.o[0] - tr.v0[0], ty = r.o[1] - tr.v0[1], tz = r.o[2] - tr.v0[2];
    const float un = std::fmaf(tx, px, std::fmaf(ty, py, tz * pz));
    const float uns = flipBy(un, det);
    if (!(uns >= 0.0f && uns <= ad)) return false;
    float qx = std::fmaf(ty, e1[2], -(tz * e1[1]));
    float qy = std::fmaf(tz, e1[0], -(tx * e1[2]));
    float qz = std::fmaf(tx, e1[1], -(ty * e1[0]));
    const float vn = std::fmaf(d[0], qx, std::fmaf(d[1], qy, d[2] * qz));
    const float vns = flipBy(vn, det);
    if (!(vns >= 0.0f && uns + vns <= ad)) return false;
    const float inv = 1.0f / det;
    u = un * inv;
    v = vn * inv;
    t = std::fmaf(e2[0], qx, std::fmaf(e2[1], qy, e2[2] * qz)) * inv;
    return true;
}

template <bool ANY>
bool traverse(const Bvh& bvh, const float o[3], const float d[3], float tmin, float tmax, uint32_t cullMask,
              vkx_hit* hit, Counters* ctr, const AnyHitFilter* filter) {
    RayCtx r; setupRay(r, o, d);
    float tbest = tmax; bool found = false;
    uint32_t bestInst = 0xFFFFFFFFu, bestPrim = 0xFFFFFFFFu; float bu = 0, bv = 0; bool bback = false;
    struct G { uint32_t base, bits; };
    G stack[64]; int sp = 0;
    G g{0, 0x01000000u}; // the root: "slot 0" of a virtual parent whose first child is node 0
    if (ctr) ctr->rays++;
    for (;;) {
        uint32_t triBase = 0, triBits = 0, triValid = 0;
        if (g.bits & 0xFF000000u) {
            // next inner child: the pending slot s with the largest (s ^ octant) (children sit in the slot of their octant, so this
            // visits them front to back along the ray)
            uint32_t pending = g.bits >> 24, slot = 0, bestKey = 0;
            for (uint32_t s = 0; s < 8; ++s) if ((pending >> s) & 1u) { uint32_t key = s ^ r.oct; if (key >= bestKey) { bestKey = key; slot = s; } }
            g.bits &= ~(1u << (24 + slot));
            if (g.bits & 0xFF000000u) { stack[sp++] = g; if (sp > gMaxStack) gMaxStack = sp; }
            uint32_t rel = uint32_t(__builtin_popcount(g.bits & 0xFFu & ((1u << slot) - 1u)));
            const Node80& n = bvh.nodes[g.base + rel];
            if (ctr) ctr->nodes++;
            uint32_t m = intersectNode(n, r, tmin, tbest);
            g.base = n.childBase; g.bits = (m & 0xFF000000u) | n.imask;
            triBase = n.primBase; triBits = m & 0x00FFFFFFu; triValid = n.valid & 0x00FFFFFFu;
        }
        while (triBits) {
            uint32_t b = uint32_t(__builtin_ctz(triBits));
            triBits &= triBits - 1;
            const Tri48& tr = bvh.tris[triBase + uint32_t(__builtin_popcount(triValid & ((1u << b) - 1u)))];
            if (ctr) ctr->tris++;
            if (!((tr.inst >> 24) & cullMask)) continue;
            float t, u, v, det;
            if (!intersectTri(tr, r, t, u, v, det)) continue;
            if (!(t > tmin)) continue;
            uint32_t inst = tr.inst & 0x00FFFFFFu, prim = tr.prim & 0x7FFFFFFFu;
            bool closer = t < tbest || (found && t == tbest && (inst < bestInst || (inst == bestInst && prim < bestPrim)));
            if (!closer) continue;
            if (filter && filter->ignore(filter->user, inst, prim, u, v)) continue;
            if (ANY) return true;
            found = true; tbest = t; bestInst = inst; bestPrim = prim; bu = u; bv = v;
            bback = (det > 0.0f) == ((tr.prim & 0x80000000u) != 0); // front <=> det > 0 (unflipped)
        }
        if (!(g.bits & 0xFF000000u)) {
            if (sp == 0) break;
            g = stack[--sp];
        }
    }
    if (ANY) return false;
    if (hit) {
        if (found) { hit->t = tbest; hit->instance = bestInst; hit->primitive = bestPrim | (bback ? 0x80000000u : 0u); hit->u = bu; hit->v = bv; }
        else { hit->t = -1.0f; hit->instance = 0xFFFFFFFFu; hit->primitive = 0xFFFFFFFFu; hit->u = 0; hit->v = 0; }
    }
    return found;
}

} // namespace

bool traceClosest(const Bvh& bvh, const float o[3], const float d[3], float tmin, float tmax, uint32_t cullMask, vkx_hit& hit, Counters* ctr, const AnyHitFilter* filter) {
    return traverse<false>(bvh, o, d, tmin, tmax, cullMask, &hit, ctr, filter);
}
bool traceAny(const Bvh& bvh, const float o[3], const float d[3], float tmin, float tmax, uint32_t cullMask, Counters* ctr, const AnyHitFilter* filter) {
    return traverse<true>(bvh, o, d, tmin, tmax, cullMask, nullptr, ctr, filter);
}

} // namespace obvh
