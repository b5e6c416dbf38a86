// TEST INFRASTRUCTURE — part of the CPU oracle. Never linked into, imported or called by the product path.
//
// CPU transliteration of the reference's 1-spp sun-shadow pass (citations relative to /root/reference):
//   src/shaders/directLight.rgen:42-99, src/shaders/shadow.rmiss, src/shaders/directLightFilter.glsl:53-142
//   (X and Y variants), host order of src/SwapchainManagement.cpp:409-438 and src/Editor.cpp:287-316.
// PARITY UNPINNED by the reference (no tests / golden images). Decrees: SURVEY A.5.4 (out-of-bounds image loads
// return 0, out-of-bounds stores are dropped), A.7. Untextured scenes: anyhit.rahit never ignores a hit.
#include "shadow.h"
#include <cmath>
#include <cstring>

using namespace ovm;

namespace oshadow {

static const float pi = 3.1415926538f;

static inline vec3 rotateAxis(vec3 p, vec3 axis, float angle) { // common.glsl:6-8
    return mix(dot(axis, p) * axis, p, std::cos(angle)) + cross(axis, p) * std::sin(angle);
}

static mat4 inverse4(const mat4& m) { // general 4x4 inverse (cofactors), fp32
    float a[16]; for (int c = 0; c < 4; ++c) { a[4 * c] = m[c].x; a[4 * c + 1] = m[c].y; a[4 * c + 2] = m[c].z; a[4 * c + 3] = m[c].w; }
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    float id = 1.0f / det;
    for (int i = 0; i < 16; ++i) inv[i] *= id;
    return mat4_from(inv);
}

void init(State& st, uint32_t w, uint32_t h) {
    st.w = w; st.h = h;
    size_t n = size_t(w) * h * 4;
    st.positionDepth.assign(n, 0.0f); st.normalMetalness.assign(n, 0.0f);
    st.albedoRoughness.assign(n, 0.0f); st.emissive.assign(n, 0.0f); st.gathered.assign(n, 0.0f);
    st.reflRaw.assign(n, 0.0f); st.reflX.assign(n, 0.0f); st.reflFinal.assign(n, 0.0f); st.reflPrevious.assign(n, 0.0f);
    st.reflDirs.assign(size_t(w) * h * 3, 0.0f); st.reflHits.assign(size_t(w) * h, vkx_hit{}); st.reflMask.assign(size_t(w) * h, 0);
    st.raw.assign(n, 0.0f); st.filteredX.assign(n, 0.0f); st.final_.assign(n, 0.0f); st.previous.assign(n, 0.0f);
    st.dirs.assign(size_t(w) * h * 3, 0.0f); st.mask.assign(size_t(w) * h, 0);
}

// Fixture: primary-ray G-buffer with the layout of GBuffer.frag:64-68 (positionDepth = world pos, |pos - camera|;
// normalMetalness = interpolated world normal, metalness). Ray set-up as raygen.rgen:27-33. Sky pixels stay 0.
void gbufferGenerate(const oddgi::Scene& s, State& st, const vkx_camera& cam) {
    mat4 iv = inverse4(mat4_from(cam.view)), ip = inverse4(mat4_from(cam.proj));
    vec4 origin4 = iv * V4(0, 0, 0, 1);
    vec3 origin = xyz(origin4);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < int64_t(st.h); ++y) for (uint32_t x = 0; x < st.w; ++x) {
        vec2 inUV = V2((float(x) + 0.5f) / float(st.w), (float(y) + 0.5f) / float(st.h));
        vec2 d = inUV * 2.0f - 1.0f;
        vec4 target = ip * V4(d.x, d.y, 1, 1);
        vec4 dir4 = iv * V4(normalize(xyz(target)), 0.0f);
        vec3 dir = xyz(dir4);
        size_t pi_ = (size_t(y) * st.w + x) * 4;
        vkx_hit h;
        float* pd = &st.positionDepth[pi_]; float* nm = &st.normalMetalness[pi_];
        float* ar = &st.albedoRoughness[pi_]; float* em = &st.emissive[pi_];
        pd[0] = pd[1] = pd[2] = pd[3] = 0.0f; nm[0] = nm[1] = nm[2] = nm[3] = 0.0f;
        ar[0] = ar[1] = ar[2] = ar[3] = 0.0f; em[0] = em[1] = em[2] = em[3] = 0.0f;
        // Fixture, not a transliteration of the raster pass: cut-outs follow anyhit.rahit (alpha < 0.01; GBuffer.frag:38 discards below 0.05)
        const obvh::AnyHitFilter filter = oddgi::anyHitFilter(s);
        if (!obvh::traceClosest(s.bvh, &origin.x, &dir.x, 0.001f, 100000.0f, 0xFFu, h, nullptr, s.textures.empty() ? nullptr : &filter)) continue;
        vec3 position = dir * h.t + origin;
        const vkx_instance& inst = s.instances[h.instance];
        const vkx_offset_entry& oe = s.offsets[inst.meshEntry];
        uint32_t prim = h.primitive & 0x7FFFFFFFu;
        vec3 n[3], col[3];
        for (int c = 0; c < 3; ++c) {
            const vkx_vertex& vx = s.vertices[oe.vertexOffset + s.indices[oe.indexOffset + 3 * prim + c]];
            n[c] = V3(vx.normal[0], vx.normal[1], vx.normal[2]); col[c] = V3(vx.color[0], vx.color[1], vx.color[2]);
        }
        vec3 on = normalize(n[0] * (1.0f - h.u - h.v) + n[1] * h.u + n[2] * h.v);
        vec3 vcolor = col[0] * (1.0f - h.u - h.v) + col[1] * h.u + col[2] * h.v; // the `color` varying of GBuffer.vert
        const mat3& W = s.worldToObject[h.instance];
        vec3 normal = normalize(V3(dot(on, W[0]), dot(on, W[1]), dot(on, W[2])));
        pd[0] = position.x; pd[1] = position.y; pd[2] = position.z; pd[3] = length(position - V3(cam.origin[0], cam.origin[1], cam.origin[2]));
        const vkx_material& mat = s.materials[oe.materialIndex];
        nm[0] = normal.x; nm[1] = normal.y; nm[2] = normal.z; nm[3] = mat.metallicFactor;
        // GBuffer.frag:35,52-53,59,66-67 (untextured): albedo = color * baseColorFactor
        ar[0] = vcolor.x * mat.baseColorFactor[0]; ar[1] = vcolor.y * mat.baseColorFactor[1]; ar[2] = vcolor.z * mat.baseColorFactor[2]; ar[3] = mat.roughnessFactor;
        em[0] = mat.emissiveFactor[0]; em[1] = mat.emissiveFactor[1]; em[2] = mat.emissiveFactor[2]; em[3] = 1.0f;
    }
}

static inline vec4 load(const std::vector<float>& img, uint32_t w, uint32_t h, int x, int y) { // robust image load: 0 out of bounds
    if (x < 0 || y < 0 || x >= int(w) || y >= int(h)) return V4(0, 0, 0, 0);
    const float* p = &img[(size_t(y) * w + size_t(x)) * 4];
    return V4(p[0], p[1], p[2], p[3]);
}

// texture(sampler2D(blueNoise[frameIndex % 64], linear/REPEAT), pixel / 64.0)  (directLight.rgen:74)
static vec4 sampleNoise(const State& st, uint32_t slice, float u, float v) {
    const float* tex = &st.noise[size_t(slice) * st.noiseW * st.noiseH * 4];
    auto setup = [](float uu, uint32_t size, int& i0, int& i1, float& f) {
        float x = uu * float(size) - 0.5f; float fl = std::floor(x); f = x - fl;
        int i = int(fl), isz = int(size); i0 = ((i % isz) + isz) % isz; i1 = (i0 + 1) % isz;
    };
    int x0, x1, y0, y1; float fx, fy;
    setup(u, st.noiseW, x0, x1, fx); setup(v, st.noiseH, y0, y1, fy);
    vec4 r;
    for (int c = 0; c < 4; ++c) {
        float t00 = tex[(size_t(y0) * st.noiseW + x0) * 4 + c], t10 = tex[(size_t(y0) * st.noiseW + x1) * 4 + c];
        float t01 = tex[(size_t(y1) * st.noiseW + x0) * 4 + c], t11 = tex[(size_t(y1) * st.noiseW + x1) * 4 + c];
        float top = t00 * (1.0f - fx) + t10 * fx, bot = t01 * (1.0f - fx) + t11 * fx;
        r[c] = top * (1.0f - fy) + bot * fy;
    }
    return r;
}

float gaussian(float stdDev, float dist) { // directLightFilter.glsl:29-31
    return (1.0f / (std::sqrt(2.0f * 3.14159f) * stdDev)) * std::exp(-(dist * dist) / (2.0f * stdDev * stdDev));
}

static const float maxDev = 7.0f;
static const int iMaxDev = 8;
static const float depthFactor = 1.0f / 0.5f;
static const float baseHysteresis = 0.94f;
static const float depthStdDev = 0.01f;
static const float historyDistanceThreshold = 0.05f;

template <int DIR>
static void filterPass(const State& st, const std::vector<float>& in, std::vector<float>& out, const vkx_camera* prevCam, const std::vector<float>* prevImg) {
    const int W = int(st.w), H = int(st.h);
    mat4 pview, pproj;
    if (DIR == 1) { pview = mat4_from(prevCam->view); pproj = mat4_from(prevCam->proj); }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        int coords[2] = {x, int(y)};
        int launchSize[2] = {W, H};
        vec4 positionDepth = load(st.positionDepth, st.w, st.h, x, int(y));
        vec3 position = xyz(positionDepth);
        float depth = positionDepth.w;
        float stdDev = 1.0f + std::max(1.0f, maxDev / (std::max(1.0f, depthFactor * depth)));
        int window = int(clampf(std::ceil(std::sqrt(-2.0f * stdDev * stdDev * std::log(0.01f * stdDev * std::sqrt(2.0f * 3.14159f)))), 1.0f, float(iMaxDev)));
        float totalFactor = 0.0f;
        int minOffset = -std::min(window, coords[DIR]);
        int maxOffset = std::min(window, launchSize[DIR] - coords[DIR]);
        vec4 fin = V4(0, 0, 0, 0);
        for (int i = minOffset; i <= maxOffset; ++i) {
            int ox = x + (DIR == 0 ? i : 0), oy = int(y) + (DIR == 1 ? i : 0);
            float factor = gaussian(stdDev, float(i));
            factor *= gaussian(depthStdDev, std::fabs(depth - load(st.positionDepth, st.w, st.h, ox, oy).w));
            totalFactor += factor;
            fin += factor * load(in, st.w, st.h, ox, oy);
        }
        if (totalFactor > 1e-2f) fin = fin / totalFactor; else fin = V4(0, 0, 0, 0);
        float* o = &out[(size_t(y) * st.w + size_t(x)) * 4];
        if (DIR == 0) { o[0] = fin.x; o[1] = fin.y; o[2] = fin.z; o[3] = depth; continue; }
        fin.x = clampf(fin.x, 0.0f, 1.0f);
        fin.y = fin.x * fin.x;
        vec4 prevCoords = pproj * (pview * V4(position, 1.0f));
        prevCoords.x /= prevCoords.w; prevCoords.y /= prevCoords.w;
        prevCoords.x = (0.5f * prevCoords.x + 0.5f) * float(W);
        prevCoords.y = (0.5f * prevCoords.y + 0.5f) * float(H);
        vec4 previousValue = V4(0, 0, 0, 0);
        float hysteresis = baseHysteresis;
        if (fin.z > 0.0f) hysteresis = fin.z == 1.0f ? 0.5f : 0.0f;
        if (prevCoords.x >= float(W) || prevCoords.x < 0.0f || prevCoords.y >= float(H) || prevCoords.y < 0.0f) hysteresis = 0.0f;
        else {
            previousValue = load(*prevImg, st.w, st.h, int(prevCoords.x), int(prevCoords.y));
            vec3 porigin = V3(prevCam->origin[0], prevCam->origin[1], prevCam->origin[2]);
            vec3 previousPosition = porigin + previousValue.w * normalize(position - porigin);
            float factor = clampf(length(position - previousPosition), 0.0f, historyDistanceThreshold) / historyDistanceThreshold;
            hysteresis *= 1.0f - clampf(factor, 0.0f, 1.0f);
            float variance = std::fabs(previousValue.x * previousValue.x - previousValue.y);
            if (variance < 0.25f && std::fabs(previousValue.x - fin.x) > 0.75f) { hysteresis = 0.0f; fin.z = 1.0f; }
            else fin.z = 0.0f;
        }
        o[0] = hysteresis * previousValue.x + (1.0f - hysteresis) * fin.x;
        o[1] = hysteresis * previousValue.y + (1.0f - hysteresis) * fin.y;
        o[2] = hysteresis * previousValue.z + (1.0f - hysteresis) * fin.z;
        o[3] = depth;
    }
}

void frame(const oddgi::Scene& s, State& st, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light, const float* dirOverride) {
    const int W = int(st.w), H = int(st.h);
    // history copy: previous <- last frame's final (src/Editor.cpp:287-316)
    st.previous = st.final_;
    vec3 L = V3(light.direction[0], light.direction[1], light.direction[2]);
    uint32_t slice = cur.frameIndex % st.noiseSlices;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < H; ++y) for (int x = 0; x < W; ++x) { // directLight.rgen:42-99
        size_t pix = size_t(y) * st.w + size_t(x);
        const float* pd = &st.positionDepth[pix * 4]; const float* nm = &st.normalMetalness[pix * 4];
        float* o = &st.raw[pix * 4];
        st.mask[pix] = 0;
        vec3 position = V3(pd[0], pd[1], pd[2]); float depth = pd[3];
        vec3 normal = V3(nm[0], nm[1], nm[2]);
        if (depth <= 0.0f) { o[0] = o[1] = o[2] = o[3] = -1.0f; continue; }
        const float* pv = &st.previous[pix * 4];
        float outColor = 0.0f;
        vec3 direction = normalize(L);
        float angle = 0.02f;
        vec4 noise = sampleNoise(st, slice, float(x) / 64.0f, float(y) / 64.0f);
        vec3 temp = rotateAxis(direction, normalize(cross(normal, direction)), 2.0f * (noise.x - 0.5f) * angle);
        direction = rotateAxis(temp, direction, 2.0f * pi * noise.y);
        if (dirOverride) direction = V3(dirOverride[3 * pix], dirOverride[3 * pix + 1], dirOverride[3 * pix + 2]);
        st.dirs[3 * pix] = direction.x; st.dirs[3 * pix + 1] = direction.y; st.dirs[3 * pix + 2] = direction.z;
        if (dot(direction, normal) > 0.0f) {
            // the direct-light pipeline's hit group is anyhit.rahit alone (src/RenderPasses/DirectLightPipeline.cpp:43-51)
            const obvh::AnyHitFilter filter = oddgi::anyHitFilter(s);
            bool isShadowed = obvh::traceAny(s.bvh, &position.x, &direction.x, 0.01f, 10000.0f, 0xFFu, nullptr, s.textures.empty() ? nullptr : &filter);
            st.mask[pix] = isShadowed ? 2 : 1;
            if (!isShadowed) {
                outColor = 1.0f;
                if (direction.y < 0.0f) outColor *= 1.0f - clampf(-direction.y, 0.0f, 0.1f) / 0.1f;
            }
            o[0] = outColor; o[1] = pv[1]; o[2] = pv[2]; o[3] = 1.0f;
        } else { o[0] = o[1] = o[2] = o[3] = 0.0f; }
    }
    filterPass<0>(st, st.raw, st.filteredX, nullptr, nullptr);
    filterPass<1>(st, st.filteredX, st.final_, &prev, &st.previous);
}

// ---------------------------------------------------------------- reflection pass
float rgaussian(float stdDev, float dist) { // reflectionFilter.glsl:37-39
    return (1.0f / (std::sqrt(2.0f * 3.14159f) * stdDev)) * std::exp(-(dist * dist) / (2.0f * stdDev * stdDev));
}

// reflectionFilter.glsl:58-147. The shared-memory cache of the shader is a plain image load here (out of bounds = 0, A.5.4).
template <int DIR>
static void reflectionFilterPass(const State& st, const std::vector<float>& in, std::vector<float>& out, const vkx_camera* cur, const vkx_camera* prevCam, const std::vector<float>* prevImg) {
    const float maxDev = 5.0f, depthFactor = 1.0f / 20.0f, baseHysteresis = 0.98f, depthStdDev = 0.1f;
    const int W = int(st.w), H = int(st.h);
    mat4 pview, pproj;
    if (DIR == 1) { pview = mat4_from(prevCam->view); pproj = mat4_from(prevCam->proj); }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        int coords[2] = {x, int(y)};
        int launchSize[2] = {W, H};
        vec4 positionDepth = load(st.positionDepth, st.w, st.h, x, int(y));
        vec3 position = xyz(positionDepth);
        float depth = positionDepth.w;
        vec4 center = load(in, st.w, st.h, x, int(y));
        float roughness = center.w;
        float* o = &out[(size_t(y) * st.w + size_t(x)) * 4];
        float stdDev = std::max(0.0f, maxDev * roughness / std::max(1.0f, depthFactor * depth));
        if (stdDev == 0.0f) { o[0] = center.x; o[1] = center.y; o[2] = center.z; o[3] = center.w; continue; } // :89-92
        float sqrDev = stdDev * stdDev;
        int window = int(clampf(std::ceil(std::sqrt(-2.0f * sqrDev * std::log(0.01f * stdDev * std::sqrt(2.0f * 3.14159f)))), 1.0f, maxDev));
        float totalFactor = 0.0f;
        int minOffset = -std::min(window, coords[DIR]);
        int maxOffset = std::min(window, launchSize[DIR] - coords[DIR]);
        vec4 fin = V4(0, 0, 0, 0);
        for (int i = minOffset; i <= maxOffset; ++i) {
            int ox = x + (DIR == 0 ? i : 0), oy = int(y) + (DIR == 1 ? i : 0);
            float factor = rgaussian(stdDev, float(i));
            factor *= rgaussian(depthStdDev, std::fabs(depth - load(st.positionDepth, st.w, st.h, ox, oy).w));
            totalFactor += factor;
            fin += factor * load(in, st.w, st.h, ox, oy);
        }
        if (totalFactor > 1e-2f) fin = fin / totalFactor; else fin = V4(0, 0, 0, 0);
        if (DIR == 0) { o[0] = fin.x; o[1] = fin.y; o[2] = fin.z; o[3] = roughness; continue; }
        float hysteresis = baseHysteresis;
        vec4 previousValue = V4(0, 0, 0, 0);
        vec3 corigin = V3(cur->origin[0], cur->origin[1], cur->origin[2]), porigin = V3(prevCam->origin[0], prevCam->origin[1], prevCam->origin[2]);
        float cameraMovement = length(corigin - porigin);
        hysteresis *= std::max(0.0f, 1.0f - cameraMovement);
        if (hysteresis > 0.0f) {
            // motion vectors are zero: static scenes (GBuffer.vert.glsl:46-48)
            vec4 prevCoords = pproj * (pview * V4(position, 1.0f));
            prevCoords.x /= prevCoords.w; prevCoords.y /= prevCoords.w;
            prevCoords.x = (0.5f * prevCoords.x + 0.5f) * float(W);
            prevCoords.y = (0.5f * prevCoords.y + 0.5f) * float(H);
            if (prevCoords.x > float(W) || prevCoords.x < 0.0f || prevCoords.y > float(H) || prevCoords.y < 0.0f) hysteresis = 0.0f; // note: > not >= (:128)
            else {
                previousValue = load(*prevImg, st.w, st.h, int(prevCoords.x), int(prevCoords.y));
                vec3 previousPosition = porigin + previousValue.w * normalize(position - porigin);
                float factor = length(position - previousPosition);
                hysteresis *= 1.0f - clampf(factor, 0.0f, 1.0f);
            }
        }
        o[0] = mix(fin.x, previousValue.x, hysteresis); o[1] = mix(fin.y, previousValue.y, hysteresis); o[2] = mix(fin.z, previousValue.z, hysteresis); o[3] = depth;
    }
}

void reflectionFrame(const oddgi::Scene& s, const oddgi::Probes& probes, State& st, const vkx_camera& cur, const vkx_camera& prev, const vkx_light& light, const float* dirOverride) {
    const int W = int(st.w), H = int(st.h);
    st.reflPrevious = st.reflFinal; // history copy (the editor copies last frame's filtered image)
    uint32_t slice = cur.frameIndex % st.noiseSlices;
    vec3 camOrigin = V3(cur.origin[0], cur.origin[1], cur.origin[2]);
    // ivec2(frameIndex / 64, frameIndex / 64 / 64) + pixel  (reflection.rgen:150)
    const int offx = int(cur.frameIndex / 64u), offy = int(cur.frameIndex / 64u / 64u);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < H; ++y) for (int x = 0; x < W; ++x) { // reflection.rgen:117-189
        size_t pix = size_t(y) * st.w + size_t(x);
        const float* pd = &st.positionDepth[pix * 4]; const float* nm = &st.normalMetalness[pix * 4]; const float* ar = &st.albedoRoughness[pix * 4];
        float* o = &st.reflRaw[pix * 4];
        vec3 position = V3(pd[0], pd[1], pd[2]); float depth = pd[3];
        vec3 normal = V3(nm[0], nm[1], nm[2]); float metalness = nm[3];
        float roughness = ar[3];
        st.reflMask[pix] = 0; st.reflHits[pix] = vkx_hit{}; st.reflHits[pix].t = -1.0f;
        st.reflDirs[3 * pix] = st.reflDirs[3 * pix + 1] = st.reflDirs[3 * pix + 2] = 0.0f;
        if (!(depth > 0.0f && (roughness < 0.4f || metalness > 0.01f))) { o[0] = o[1] = o[2] = o[3] = 0.0f; continue; }
        vec3 toOrigin = normalize(camOrigin - position);
        vec3 reflectDir = normalize(reflect(-toOrigin, normal));
        vec4 noise = sampleNoise(st, slice, float(offx + x) / 64.0f, float(offy + int(y)) / 64.0f);
        const float theta = roughness * (noise.x - 0.5f) * 2.0f * pi;
        const float phi = (noise.y - 0.5f) * 2.0f * pi;
        vec3 tangent;
        if (dot(reflectDir, normal) < 0.9f) tangent = normalize(cross(reflectDir, normal));
        else tangent = normalize(cross(reflectDir, V3(1, 0, 0)));
        vec3 direction = rotateAxis(reflectDir, tangent, theta);
        direction = rotateAxis(direction, reflectDir, phi);
        if (dirOverride) direction = V3(dirOverride[3 * pix], dirOverride[3 * pix + 1], dirOverride[3 * pix + 2]);
        st.reflDirs[3 * pix] = direction.x; st.reflDirs[3 * pix + 1] = direction.y; st.reflDirs[3 * pix + 2] = direction.z;
        vkx_hit hit; uint8_t shadowFlag = 0;
        // payload.raydx / raydy, reflection.rgen:169-170; the reflection pipeline's hit group has the any-hit shader
        vec3 raydx = V3(0.0f), raydy = V3(0.0f);
        if (!s.textures.empty()) { raydx = oddgi::rotateAxis(direction, normal, 0.001f); raydy = oddgi::rotateAxis(direction, cross(normal, direction), 0.001f); }
        vec4 c = oddgi::traceAndShade(s, probes, light, position, direction, 0.1f, 10000.0f, 0xFFu, hit, shadowFlag, raydx, raydy, true);
        if (c.w < 0.0f) { st.reflMask[pix] = 1; hit = vkx_hit{}; hit.t = -1.0f; }
        else st.reflMask[pix] = (hit.primitive & 0x80000000u) ? 2 : (shadowFlag == 2 ? 4 : 3);
        st.reflHits[pix] = hit;
        vec3 v = xyz(c);
        // colorCompression = reinhard_whitepoint(v, 1.0) (:95-110)
        const float max_value = 1.0f;
        vec3 comp = v * (V3(1, 1, 1) + (v / (max_value * max_value))) / (V3(1, 1, 1) + v);
        o[0] = comp.x; o[1] = comp.y; o[2] = comp.z; o[3] = roughness;
    }
    reflectionFilterPass<0>(st, st.reflRaw, st.reflX, nullptr, nullptr, nullptr);
    reflectionFilterPass<1>(st, st.reflX, st.reflFinal, &cur, &prev, &st.reflPrevious);
}

// FinalGather.frag:38-77. fragPosition of FullScreenQuad.vert interpolates to the pixel centre ((x + 0.5) / W, (y + 0.5) / H).
// inverse() is the cofactor expansion above (GLSL leaves its precision to the implementation).
void finalGather(const oddgi::Scene& s, const oddgi::Probes& probes, State& st, const vkx_camera& cam, const vkx_light& light, const float* reflection) {
    (void)s;
    const int W = int(st.w), H = int(st.h);
    mat4 iv = inverse4(mat4_from(cam.view)), ip = inverse4(mat4_from(cam.proj));
    vec3 origin = xyz(iv * V4(0, 0, 0, 1));
    vec3 Ldir = V3(light.direction[0], light.direction[1], light.direction[2]);
    vec3 Lcol = V3(light.color[0], light.color[1], light.color[2]);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        size_t pi_ = (size_t(y) * st.w + size_t(x)) * 4;
        vec3 color = V3(0, 0, 0);
        vec2 fragPosition = V2((float(x) + 0.5f) / float(W), (float(y) + 0.5f) / float(H));
        const float* pd = &st.positionDepth[pi_]; const float* nm = &st.normalMetalness[pi_];
        const float* ar = &st.albedoRoughness[pi_]; const float* em = &st.emissive[pi_];
        vec3 position = V3(pd[0], pd[1], pd[2]);
        float depth = pd[3];
        if (depth <= 0.0f) {
            vec2 d = 2.0f * fragPosition - 1.0f;
            vec4 t = ip * V4(d.x, d.y, 0.0f, 1.0f);
            vec4 dir4 = iv * V4(normalize(xyz(t)), 0.0f);
            color = oddgi::sky(origin, xyz(dir4), Ldir, Lcol, 1.0f, true);
        } else {
            vec3 normal = normalize(V3(nm[0], nm[1], nm[2]));
            float metalness = nm[3];
            vec4 albedo = V4(ar[0], ar[1], ar[2], 1.0f);
            float roughness = ar[3];
            vec3 refl = reflection ? V3(reflection[pi_], reflection[pi_ + 1], reflection[pi_ + 2]) : V3(0, 0, 0);
            vec3 view = normalize(origin - position);
            float direct = st.final_[pi_]; // subpassLoad(inputDirectLight).r
            color += direct * xyz(oddgi::pbrMetallicRoughness(normal, view, Lcol, Ldir, albedo, metalness, roughness));
            vec3 f0 = V3(0.004f, 0.004f, 0.004f);
            vec3 diffuseColor = xyz(albedo) * (V3(1, 1, 1) - f0);
            diffuseColor = diffuseColor * (1.0f - metalness);
            vec3 specularColor = mix(f0, xyz(albedo), metalness);
            color += specularColor * refl;
            vec3 indirectLight = oddgi::sampleProbes(probes, position, normal, view);
            color += indirectLight * diffuseColor;
            color += V3(em[0], em[1], em[2]);
        }
        float* o = &st.gathered[pi_];
        o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = 1.0f;
    }
}

} // namespace oshadow
