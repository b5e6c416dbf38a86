// TEST INFRASTRUCTURE — C API of the CPU oracle (liboracle.so). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product path never does.
#include <chrono>
#include <cstring>
#include <new>
#include "bvh.h"
#include "ddgi.h"
#include "shadow.h"
#include "packing.h"
#ifdef _OPENMP
#include <omp.h>
#endif

struct orc_ctx {
    oddgi::Scene scene;
    oddgi::Probes probes;
    oshadow::State shadow;
};

extern "C" {

orc_ctx* orc_create() { return new (std::nothrow) orc_ctx(); }
void orc_destroy(orc_ctx* c) { delete c; }
// torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arms call this with the cores they may use.
int orc_max_stack(int reset) { int v = obvh::gMaxStack; if (reset) obvh::gMaxStack = 0; return v; }
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_scene_upload(orc_ctx* c, const vkx_vertex* v, size_t nv, const uint32_t* idx, size_t ni, const vkx_offset_entry* off,
                     const uint32_t* counts, size_t nm, const vkx_material* mat, size_t nmat, const vkx_instance* inst, size_t ninst) {
    oddgi::Scene& s = c->scene;
    s.vertices.assign(v, v + nv); s.indices.assign(idx, idx + ni); s.offsets.assign(off, off + nm);
    s.meshIndexCounts.assign(counts, counts + nm); s.materials.assign(mat, mat + nmat); s.instances.assign(inst, inst + ninst);
    return 0;
}
// Scene texture list (mip chains generated here, sampler spec v1); numTextures = 0 clears it.
int orc_scene_textures(orc_ctx* c, const vkx_texture* tex, size_t n) {
    c->scene.textures.clear();
    for (size_t i = 0; i < n; ++i) c->scene.textures.push_back(otex::makeTexture(tex[i]));
    return 0;
}
int orc_texture_download(orc_ctx* c, uint32_t index, void* texels, size_t bytes, uint32_t* levels) {
    if (index >= c->scene.textures.size()) return -1;
    const otex::Texture& t = c->scene.textures[index];
    if (levels) *levels = t.levels;
    size_t total = 0; for (auto& l : t.mip) total += l.size() * 4;
    if (texels) { if (bytes < total) return -1; char* o = static_cast<char*>(texels); for (auto& l : t.mip) { std::memcpy(o, l.data(), l.size() * 4); o += l.size() * 4; } }
    return 0;
}
int orc_texture_sample(orc_ctx* c, uint32_t index, const float* uv, const float* grads, size_t n, float* out) {
    if (index >= c->scene.textures.size()) return -1;
    const otex::Texture& t = c->scene.textures[index];
    for (size_t i = 0; i < n; ++i) {
        otex::RGBA r = grads ? otex::sampleGrad(t, uv[2 * i], uv[2 * i + 1], grads[4 * i], grads[4 * i + 1], grads[4 * i + 2], grads[4 * i + 3]) : otex::sampleBase(t, uv[2 * i], uv[2 * i + 1]);
        out[4 * i] = r.r; out[4 * i + 1] = r.g; out[4 * i + 2] = r.b; out[4 * i + 3] = r.a;
    }
    return 0;
}
// texDerivative for one hit (test hook): v = 3 vertices, o2w = row-major 3x4 instance transform
int orc_tex_derivative(const float position[3], const float rayOrigin[3], const float transform[12], const vkx_vertex* v, const float raydx[3], const float raydy[3], float out[4]) {
    ovm::mat3 m; for (int c = 0; c < 3; ++c) m[c] = ovm::V3(transform[c], transform[4 + c], transform[8 + c]);
    ovm::vec4 g = oddgi::texDerivative(ovm::V3(position[0], position[1], position[2]), ovm::V3(rayOrigin[0], rayOrigin[1], rayOrigin[2]), m, v[0], v[1], v[2],
                                       ovm::V3(raydx[0], raydx[1], raydx[2]), ovm::V3(raydy[0], raydy[1], raydy[2]));
    out[0] = g.x; out[1] = g.y; out[2] = g.z; out[3] = g.w;
    return 0;
}
int orc_skin_vertices(orc_ctx* c, const float* jointTransforms, size_t numJoints, const uint16_t* skinJoints, const float* skinWeights, uint32_t srcOffset, uint32_t dstOffset, uint32_t size, float* motionVectors) {
    (void)numJoints;
    oddgi::skinVertices(c->scene, jointTransforms, skinJoints, skinWeights, srcOffset, dstOffset, size, motionVectors);
    return 0;
}
int orc_vertices_download(orc_ctx* c, size_t first, size_t count, vkx_vertex* out) {
    if (first + count > c->scene.vertices.size()) return -1;
    std::memcpy(out, c->scene.vertices.data() + first, count * sizeof(vkx_vertex));
    return 0;
}
int orc_bvh_build(orc_ctx* c) { oddgi::sceneFinalize(c->scene); return 0; }
// New instance transforms / masks for the uploaded instance list (same meshes), then either orc_bvh_build (rebuild) or orc_bvh_refit.
int orc_instances_update(orc_ctx* c, const vkx_instance* inst, size_t n) {
    if (n != c->scene.instances.size()) return -1;
    c->scene.instances.assign(inst, inst + n);
    return 0;
}
int orc_bvh_refit(orc_ctx* c) { oddgi::sceneRefit(c->scene); return 0; }
int orc_bvh_info(orc_ctx* c, vkx_bvh_info* out) {
    const obvh::Bvh& b = c->scene.bvh;
    std::memset(out, 0, sizeof(*out));
    out->numNodes = uint32_t(b.nodes.size()); out->numTriangles = uint32_t(b.tris.size());
    out->numBinaryNodes = b.numBinaryNodes; out->depth = b.depth;
    for (int a = 0; a < 3; ++a) { out->sceneMin[a] = b.sceneMin[a]; out->sceneMax[a] = b.sceneMax[a]; }
    return 0;
}
int orc_bvh_download(orc_ctx* c, void* nodes, size_t nb, void* tris, size_t tb) {
    const obvh::Bvh& b = c->scene.bvh;
    if (nodes) { if (nb < b.nodes.size() * 80) return -1; std::memcpy(nodes, b.nodes.data(), b.nodes.size() * 80); }
    if (tris) { if (tb < b.tris.size() * 48) return -1; std::memcpy(tris, b.tris.data(), b.tris.size() * 48); }
    return 0;
}
int orc_trace(orc_ctx* c, const float* o, const float* d, size_t n, float tmin, float tmax, uint32_t mask, int any, vkx_hit* out, uint64_t* counters) {
    obvh::Counters total;
    // any: bit 0 terminate on first hit, bit 1 run anyhit.rahit on the candidates (alpha cut-outs)
    const obvh::AnyHitFilter filter = oddgi::anyHitFilter(c->scene);
    const obvh::AnyHitFilter* flt = ((any & 2) && !c->scene.textures.empty()) ? &filter : nullptr;
    any &= 1;
#pragma omp parallel
    {
        obvh::Counters ctr;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < int64_t(n); ++i) {
            if (any) {
                bool h = obvh::traceAny(c->scene.bvh, o + 3 * i, d + 3 * i, tmin, tmax, mask, &ctr, flt);
                out[i].t = h ? 1.0f : -1.0f; out[i].instance = 0xFFFFFFFFu; out[i].primitive = 0xFFFFFFFFu; out[i].u = out[i].v = 0.0f;
            } else obvh::traceClosest(c->scene.bvh, o + 3 * i, d + 3 * i, tmin, tmax, mask, out[i], &ctr, flt);
        }
#pragma omp critical
        { total.nodes += ctr.nodes; total.tris += ctr.tris; total.rays += ctr.rays; }
    }
    if (counters) { counters[0] = total.rays; counters[1] = total.nodes; counters[2] = total.tris; }
    return 0;
}

int orc_probes_init(orc_ctx* c, const vkx_grid_info* g) { oddgi::probesInit(c->probes, *g); return 0; }
int orc_probes_classify(orc_ctx* c, const float R[16]) { oddgi::classify(c->scene, c->probes, R); return 0; }
// returns elapsed seconds of the update in *seconds (trace+shade+blend+border+publish)
int orc_probes_update(orc_ctx* c, const vkx_grid_info* g, const vkx_light* l, const float R[16], const uint32_t* indices, uint32_t count, int threads, double* seconds) {
    auto t0 = std::chrono::steady_clock::now();
    oddgi::update(c->scene, c->probes, *g, *l, R, indices, count, threads);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
int orc_probes_download(orc_ctx* c, uint32_t* irr, uint32_t* dep, uint32_t* state, float* rays, size_t raysCap) {
    oddgi::Probes& p = c->probes;
    if (irr) std::memcpy(irr, p.irrSampled.data(), p.irrSampled.size() * 4);
    if (dep) std::memcpy(dep, p.depSampled.data(), p.depSampled.size() * 4);
    if (state) std::memcpy(state, p.state.data(), p.state.size() * 4);
    if (rays) { if (raysCap < p.rays.size() * 4) return -1; std::memcpy(rays, p.rays.data(), p.rays.size() * 4); }
    return 0;
}
int orc_probes_upload(orc_ctx* c, const uint32_t* irr, const uint32_t* dep, const uint32_t* state) {
    oddgi::Probes& p = c->probes;
    if (irr) { std::memcpy(p.irrSampled.data(), irr, p.irrSampled.size() * 4); p.irrWork = p.irrSampled; }
    if (dep) { std::memcpy(p.depSampled.data(), dep, p.depSampled.size() * 4); p.depWork = p.depSampled; }
    if (state) std::memcpy(p.state.data(), state, p.state.size() * 4);
    return 0;
}
int orc_probes_download_unpacked(orc_ctx* c, float* irr, float* dep) {
    oddgi::Probes& p = c->probes;
    if (irr) std::memcpy(irr, p.irrUnpacked.data(), p.irrUnpacked.size() * 4);
    if (dep) std::memcpy(dep, p.depUnpacked.data(), p.depUnpacked.size() * 4);
    return 0;
}
int orc_probes_download_hits(orc_ctx* c, vkx_hit* hits, uint8_t* shadow) {
    oddgi::Probes& p = c->probes;
    if (hits) std::memcpy(hits, p.hits.data(), p.hits.size() * sizeof(vkx_hit));
    if (shadow) std::memcpy(shadow, p.shadow.data(), p.shadow.size());
    return 0;
}
// counters of the last update: primary rays, nodes visited, triangles tested, front hits, shadow rays, nodes, triangles
int orc_probes_counters(orc_ctx* c, uint64_t out[7]) {
    out[0] = c->probes.counters.rays; out[1] = c->probes.counters.nodes; out[2] = c->probes.counters.tris; out[3] = c->probes.frontHits;
    out[4] = c->probes.shadowCounters.rays; out[5] = c->probes.shadowCounters.nodes; out[6] = c->probes.shadowCounters.tris;
    return 0;
}
int orc_ray_directions(const float R[16], uint32_t count, float n, float* out) {
    std::vector<float> d; oddgi::rayDirections(R, count, n, d); std::memcpy(out, d.data(), d.size() * 4); return 0;
}

// pure-function KATs
void orc_spherical_fibonacci(float i, float n, float out[3]) { ovm::vec3 v = oddgi::sphericalFibonacci(i, n); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void orc_oct_decode(float x, float y, float out[3]) { ovm::vec3 v = oddgi::octDecode(ovm::V2(x, y)); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void orc_sphere_to_oct_uv(const float d[3], float out[2]) { ovm::vec2 v = oddgi::spherePointToOctohedralUV(ovm::V3(d[0], d[1], d[2])); out[0] = v.x; out[1] = v.y; }
uint32_t orc_pack_r11g11b10(float r, float g, float b) { return opack::packR11G11B10(r, g, b); }
void orc_unpack_r11g11b10(uint32_t p, float out[3]) { opack::unpackR11G11B10(p, out); }
uint32_t orc_pack_rg16f(float r, float g) { return opack::packRG16F(r, g); }
void orc_unpack_rg16f(uint32_t p, float out[2]) { opack::unpackRG16F(p, out); }
void orc_sky(const float o[3], const float d[3], const vkx_light* l, float out[3]) {
    ovm::vec3 c = oddgi::sky(ovm::V3(o[0], o[1], o[2]), ovm::V3(d[0], d[1], d[2]), ovm::V3(l->direction[0], l->direction[1], l->direction[2]),
                             ovm::V3(l->color[0], l->color[1], l->color[2]), l->color[3], true);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void orc_sample_probes(orc_ctx* c, const float pos[3], const float n[3], const float toCam[3], float out[3]) {
    ovm::vec3 r = oddgi::sampleProbes(c->probes, ovm::V3(pos[0], pos[1], pos[2]), ovm::V3(n[0], n[1], n[2]), ovm::V3(toCam[0], toCam[1], toCam[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// ---- function-level batch entry points: the restated shader functions one by one, for the pin against oracle/_ref
// (the reference's own GLSL text compiled as C++, tests/test_glsl_ref_pin.py).
void orc_fetch_atlas(const void* user, int kind, float u, float v, float out4[4]) { // textureLod(colorTex | depthTex, uv, 0)
    const orc_ctx* c = static_cast<const orc_ctx*>(user);
    if (kind == 0) { ovm::vec3 r = oddgi::sampleIrradiance(c->probes, ovm::V2(u, v)); out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = 1.0f; }
    else { ovm::vec2 r = oddgi::sampleDepth(c->probes, ovm::V2(u, v)); out4[0] = r.x; out4[1] = r.y; out4[2] = 0.0f; out4[3] = 1.0f; }
}
void orc_fn_sky(const float* o, const float* d, const float* sun, const float* sunColor, float brightness, int showSun, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        ovm::vec3 c = oddgi::sky(ovm::V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), ovm::V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), ovm::V3(sun[0], sun[1], sun[2]), ovm::V3(sunColor[0], sunColor[1], sunColor[2]), brightness, showSun != 0);
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
void orc_fn_sample_probes(orc_ctx* c, const float* pos, const float* nrm, const float* toCam, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        ovm::vec3 r = oddgi::sampleProbes(c->probes, ovm::V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), ovm::V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), ovm::V3(toCam[3 * i], toCam[3 * i + 1], toCam[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
void orc_fn_pbr(const float* nrm, const float* view, const float* lightColor, const float* lightDir, const float* albedo, const float* metalRough, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        ovm::vec4 c = oddgi::pbrMetallicRoughness(ovm::V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), ovm::V3(view[3 * i], view[3 * i + 1], view[3 * i + 2]), ovm::V3(lightColor[0], lightColor[1], lightColor[2]),
                                                  ovm::V3(lightDir[0], lightDir[1], lightDir[2]), ovm::V4(albedo[4 * i], albedo[4 * i + 1], albedo[4 * i + 2], albedo[4 * i + 3]), metalRough[2 * i], metalRough[2 * i + 1]);
        out[4 * i] = c.x; out[4 * i + 1] = c.y; out[4 * i + 2] = c.z; out[4 * i + 3] = c.w;
    }
}
void orc_fn_spherical_fibonacci(const float* i_, float nn, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { ovm::vec3 v = oddgi::sphericalFibonacci(i_[i], nn); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
void orc_fn_oct_decode(const float* o, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { ovm::vec3 v = oddgi::octDecode(ovm::V2(o[2 * i], o[2 * i + 1])); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
void orc_fn_oct_encode(const float* d, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { ovm::vec2 v = oddgi::octEncode(ovm::V3(d[3 * i], d[3 * i + 1], d[3 * i + 2])); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
void orc_fn_sphere_to_oct_uv(const float* d, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { ovm::vec2 v = oddgi::spherePointToOctohedralUV(ovm::V3(d[3 * i], d[3 * i + 1], d[3 * i + 2])); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
void orc_fn_rotate_axis(const float* p, const float* axis, const float* angle, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) { ovm::vec3 v = oddgi::rotateAxis(ovm::V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), ovm::V3(axis[3 * i], axis[3 * i + 1], axis[3 * i + 2]), angle[i]); out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z; }
}
void orc_fn_gaussian(const float* stdDev, const float* dist, size_t n, float* out) { for (size_t i = 0; i < n; ++i) out[i] = oshadow::gaussian(stdDev[i], dist[i]); }
void orc_fn_gaussian_refl(const float* stdDev, const float* dist, size_t n, float* out) { for (size_t i = 0; i < n; ++i) out[i] = oshadow::rgaussian(stdDev[i], dist[i]); }
void orc_fn_probe_helpers(const vkx_grid_info* g, const uint32_t* index, size_t n, int* outI, float* outF) {
    for (size_t i = 0; i < n; ++i) oddgi::probeHelpers(*g, index[i], outI + 8 * i, outF + 6 * i);
}
void orc_border_source(int T, int x, int y, int out[2]);

// host logic (IrradianceProbes.cpp)
struct orc_host { oddgi::MsvcRand rng; oddgi::Scheduler sched; };
orc_host* orc_host_create() { return new orc_host(); }
void orc_host_destroy(orc_host* h) { delete h; }
void orc_host_next_orientation(orc_host* h, float out16[16], float Zout[3]) {
    float Z[3]; oddgi::sphericalRand(h->rng, Z); oddgi::orientationFromZ(Z, out16);
    if (Zout) { Zout[0] = Z[0]; Zout[1] = Z[1]; Zout[2] = Z[2]; }
}
void orc_orientation_from_z(const float Z[3], float out16[16]) { oddgi::orientationFromZ(Z, out16); }
int orc_host_rand(orc_host* h) { return h->rng.next(); }
uint32_t orc_host_select(orc_host* h, const uint32_t* state, uint32_t probeCount, uint32_t perUpdate, uint32_t* out) {
    return oddgi::selectProbesToUpdate(h->sched, state, probeCount, perUpdate, out);
}

// shadows
int orc_shadow_set_noise(orc_ctx* c, const float* rgba, uint32_t w, uint32_t h, uint32_t slices) {
    c->shadow.noise.assign(rgba, rgba + size_t(w) * h * slices * 4); c->shadow.noiseW = w; c->shadow.noiseH = h; c->shadow.noiseSlices = slices; return 0;
}
int orc_shadow_init(orc_ctx* c, uint32_t w, uint32_t h) { oshadow::init(c->shadow, w, h); return 0; }
int orc_gbuffer_generate(orc_ctx* c, const vkx_camera* cam) { oshadow::gbufferGenerate(c->scene, c->shadow, *cam); return 0; }
int orc_gbuffer_upload(orc_ctx* c, const float* pd, const float* nm) {
    std::memcpy(c->shadow.positionDepth.data(), pd, c->shadow.positionDepth.size() * 4);
    std::memcpy(c->shadow.normalMetalness.data(), nm, c->shadow.normalMetalness.size() * 4); return 0;
}
int orc_gbuffer_download(orc_ctx* c, float* pd, float* nm) {
    if (pd) std::memcpy(pd, c->shadow.positionDepth.data(), c->shadow.positionDepth.size() * 4);
    if (nm) std::memcpy(nm, c->shadow.normalMetalness.data(), c->shadow.normalMetalness.size() * 4);
    return 0;
}
int orc_gbuffer_upload_material(orc_ctx* c, const float* ar, const float* em) {
    std::memcpy(c->shadow.albedoRoughness.data(), ar, c->shadow.albedoRoughness.size() * 4);
    std::memcpy(c->shadow.emissive.data(), em, c->shadow.emissive.size() * 4); return 0;
}
int orc_gbuffer_download_material(orc_ctx* c, float* ar, float* em) {
    if (ar) std::memcpy(ar, c->shadow.albedoRoughness.data(), c->shadow.albedoRoughness.size() * 4);
    if (em) std::memcpy(em, c->shadow.emissive.data(), c->shadow.emissive.size() * 4);
    return 0;
}
int orc_reflection_frame(orc_ctx* c, const vkx_camera* cur, const vkx_camera* prev, const vkx_light* l, const float* dirOverride, double* seconds) {
    auto t0 = std::chrono::steady_clock::now();
    oshadow::reflectionFrame(c->scene, c->probes, c->shadow, *cur, *prev, *l, dirOverride);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
int orc_reflection_download(orc_ctx* c, int stage, float* rgba, float* dirs, vkx_hit* hits, uint8_t* mask) {
    const std::vector<float>& src = stage == 0 ? c->shadow.reflRaw : stage == 1 ? c->shadow.reflX : c->shadow.reflFinal;
    if (rgba) std::memcpy(rgba, src.data(), src.size() * 4);
    if (dirs) std::memcpy(dirs, c->shadow.reflDirs.data(), c->shadow.reflDirs.size() * 4);
    if (hits) std::memcpy(hits, c->shadow.reflHits.data(), c->shadow.reflHits.size() * sizeof(vkx_hit));
    if (mask) std::memcpy(mask, c->shadow.reflMask.data(), c->shadow.reflMask.size());
    return 0;
}
int orc_reflection_set_history(orc_ctx* c, const float* rgba) { std::memcpy(c->shadow.reflFinal.data(), rgba, c->shadow.reflFinal.size() * 4); return 0; }
int orc_final_gather(orc_ctx* c, const vkx_camera* cam, const vkx_light* l, const float* reflection, double* seconds) {
    auto t0 = std::chrono::steady_clock::now();
    oshadow::finalGather(c->scene, c->probes, c->shadow, *cam, *l, reflection);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
int orc_final_gather_download(orc_ctx* c, float* rgba) { std::memcpy(rgba, c->shadow.gathered.data(), c->shadow.gathered.size() * 4); return 0; }
int orc_shadow_frame(orc_ctx* c, const vkx_camera* cur, const vkx_camera* prev, const vkx_light* l, const float* dirOverride, double* seconds) {
    auto t0 = std::chrono::steady_clock::now();
    oshadow::frame(c->scene, c->shadow, *cur, *prev, *l, dirOverride);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}
int orc_shadow_download(orc_ctx* c, int stage, float* rgba, float* dirs, uint8_t* mask) {
    const std::vector<float>& src = stage == 0 ? c->shadow.raw : stage == 1 ? c->shadow.filteredX : c->shadow.final_;
    if (rgba) std::memcpy(rgba, src.data(), src.size() * 4);
    if (dirs) std::memcpy(dirs, c->shadow.dirs.data(), c->shadow.dirs.size() * 4);
    if (mask) std::memcpy(mask, c->shadow.mask.data(), c->shadow.mask.size());
    return 0;
}
int orc_shadow_set_history(orc_ctx* c, const float* rgba) { std::memcpy(c->shadow.final_.data(), rgba, c->shadow.final_.size() * 4); return 0; }
int orc_shadow_reset_history(orc_ctx* c) {
    std::fill(c->shadow.final_.begin(), c->shadow.final_.end(), 0.0f); std::fill(c->shadow.previous.begin(), c->shadow.previous.end(), 0.0f); return 0;
}

} // extern "C"

// border table restated as a formula (checked against the reference's tables in tests/test_oracle_kat.py)
extern "C" void orc_border_source(int T, int x, int y, int out[2]) {
    const int L = T - 1;
    bool bx = (x == 0 || x == L), by = (y == 0 || y == L);
    if (bx && by) { out[0] = x == 0 ? L - 1 : 1; out[1] = y == 0 ? L - 1 : 1; }
    else if (bx) { out[0] = x == 0 ? 1 : L - 1; out[1] = L - y; }
    else { out[0] = L - x; out[1] = y == 0 ? 1 : L - 1; }
}
