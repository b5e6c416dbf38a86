#include "Renderer.hpp"
#include <algorithm>

namespace vkx {

void check(vkx_ctx* ctx, int rc) {
    if (rc != VKX_OK) throw Error(rc, std::string("vkx error ") + std::to_string(rc) + ": " + vkx_last_error(ctx));
}

Device::Device(int index) {
    int rc = vkx_create(index, &_ctx);
    if (rc != VKX_OK) throw Error(rc, std::string("vkx_create: ") + vkx_last_error(nullptr));
}
Device::~Device() { vkx_destroy(_ctx); }

void Renderer::allocateMeshes() {
    Vertices.clear(); Indices.clear(); OffsetTable.clear(); MeshIndexCounts.clear();
    uint32_t vo = 0, io = 0;
    for (Mesh& m : _scene->getMeshes()) {
        if (!m.isValid()) continue; // invalid meshes are skipped (reference src/Renderer.cpp:103-105)
        m.indexIntoOffsetTable = uint32_t(OffsetTable.size());
        OffsetTable.push_back(OffsetEntry{m.defaultMaterialIndex, vo, io});
        MeshIndexCounts.push_back(uint32_t(m.indices.size()));
        Vertices.insert(Vertices.end(), m.vertices.begin(), m.vertices.end());
        Indices.insert(Indices.end(), m.indices.begin(), m.indices.end());
        vo += uint32_t(m.vertices.size()); io += uint32_t(m.indices.size());
    }
    allocateSkinnedMeshes();
}

void Renderer::allocateSkinnedMeshes() { // updateSkinnedMeshOffsetTable (reference src/Renderer.cpp:144-164) + the bind-pose copy (:812-831)
    for (SkinnedMeshRenderer& r : _scene->getSkinnedRenderers()) {
        const Mesh& mesh = _scene->getMeshes().at(r.meshIndex);
        if (!mesh.isValid()) throw Error(VKX_E_INVALID, "allocateSkinnedMeshes: skinned renderer uses an invalid mesh");
        if (r.joints.size() != 4 * mesh.vertices.size() || r.weights.size() != 4 * mesh.vertices.size()) throw Error(VKX_E_INVALID, "allocateSkinnedMeshes: 4 joints and 4 weights per vertex expected");
        r.indexIntoOffsetTable = uint32_t(OffsetTable.size());
        r.vertexOffset = uint32_t(Vertices.size());
        OffsetTable.push_back(OffsetEntry{r.materialIndex, r.vertexOffset, OffsetTable.at(mesh.indexIntoOffsetTable).indexOffset});
        MeshIndexCounts.push_back(uint32_t(mesh.indices.size()));
        Vertices.insert(Vertices.end(), mesh.vertices.begin(), mesh.vertices.end());
    }
}

bool Renderer::updateSkinnedVertexBuffer(const std::vector<std::vector<mat4>>& jointPoses) {
    const auto& skinned = _scene->getSkinnedRenderers();
    if (skinned.empty()) return false;
    if (jointPoses.size() != skinned.size()) throw Error(VKX_E_INVALID, "updateSkinnedVertexBuffer: one pose array per skinned renderer expected");
    vkx_ctx* ctx = _device->ctx();
    for (size_t i = 0; i < skinned.size(); ++i) {
        const SkinnedMeshRenderer& r = skinned[i];
        const Mesh& mesh = _scene->getMeshes().at(r.meshIndex);
        check(ctx, vkx_skin_vertices(ctx, &jointPoses[i].at(0).m[0][0], jointPoses[i].size(), r.joints.data(), r.weights.data(), OffsetTable.at(mesh.indexIntoOffsetTable).vertexOffset,
                                     r.vertexOffset, uint32_t(mesh.vertices.size()), nullptr));
    }
    return true;
}

bool Renderer::updateSkinnedBLAS() {
    if (_scene->getSkinnedRenderers().empty()) return false; // reference src/Renderer.cpp:647-648
    check(_device->ctx(), vkx_bvh_build(_device->ctx()));
    return true;
}

void Renderer::createTLAS() {
    // sortRenderers: by (materialIndex, meshIndex), entity order as the final key (reference src/Renderer.cpp:512-523)
    struct R { uint32_t material, mesh; int entity; };
    std::vector<R> rs;
    const auto& nodes = _scene->getNodes();
    for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i].hasMeshRenderer) rs.push_back(R{nodes[i].materialIndex, nodes[i].meshIndex, int(i)});
    std::stable_sort(rs.begin(), rs.end(), [](const R& a, const R& b) { return a.material == b.material ? a.mesh < b.mesh : a.material < b.material; });
    _instances.clear();
    for (const R& r : rs) {
        const Mesh& mesh = _scene->getMeshes().at(r.mesh);
        if (!mesh.isValid()) continue;
        vkx_instance inst{};
        const mat4& g = nodes[size_t(r.entity)].globalTransform; // column-major -> row-major 3x4 (VkTransformMatrixKHR)
        for (int row = 0; row < 3; ++row) for (int col = 0; col < 4; ++col) inst.transform[4 * row + col] = g.m[col][row];
        inst.meshEntry = mesh.indexIntoOffsetTable;
        inst.mask = VKX_INSTANCE_STATIC;
        _instances.push_back(inst);
    }
    for (const SkinnedMeshRenderer& r : _scene->getSkinnedRenderers()) { // skinned instances follow the static ones (reference src/Renderer.cpp:553-575)
        vkx_instance inst{};
        const mat4& g = nodes.at(size_t(r.node)).globalTransform;
        for (int row = 0; row < 3; ++row) for (int col = 0; col < 4; ++col) inst.transform[4 * row + col] = g.m[col][row];
        inst.meshEntry = r.indexIntoOffsetTable;
        inst.mask = VKX_INSTANCE_SKINNED;
        _instances.push_back(inst);
    }
}

void Renderer::createAccelerationStructures() {
    createTLAS();
    std::vector<vkx_material> mats;
    for (const auto& m : _scene->getMaterials()) mats.push_back(m.properties);
    vkx_ctx* ctx = _device->ctx();
    { // uploadTextures (reference src/Resources.cpp:46-95): images + sampler descriptions; mip chains are generated on the device
        std::vector<vkx_texture> tex;
        for (const auto& t : _scene->getTextures()) {
            vkx_texture d{};
            d.pixels = t.image.pixels.data(); d.width = t.image.width; d.height = t.image.height;
            d.srgb = t.format == 43u ? 1u : 0u; // VK_FORMAT_R8G8B8A8_SRGB, else R8G8B8A8_UNORM (src/Scene.cpp:43,671,677)
            d.magFilter = t.magFilter; d.minFilter = t.minFilter; d.wrapS = t.wrapS; d.wrapT = t.wrapT;
            tex.push_back(d);
        }
        check(ctx, vkx_scene_textures(ctx, tex.data(), tex.size()));
    }
    check(ctx, vkx_scene_upload(ctx, Vertices.data(), Vertices.size(), Indices.data(), Indices.size(), OffsetTable.data(), MeshIndexCounts.data(), OffsetTable.size(),
                                mats.data(), mats.size(), _instances.data(), _instances.size()));
    check(ctx, vkx_bvh_build(ctx));
}

void Renderer::updateAccelerationStructureInstances() {
    const size_t before = _instances.size();
    createTLAS(); // same order (sortRenderers is stable), transforms re-read from the scene graph
    if (_instances.size() != before) throw Error(VKX_E_INVALID, "updateAccelerationStructureInstances: the set of renderers changed; call createAccelerationStructures");
    check(_device->ctx(), vkx_instances_update(_device->ctx(), _instances.data(), _instances.size()));
}

// The reference refits its TLAS in place (VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR, src/Renderer.cpp:681-733): the default here
// too (vkx_bvh_refit keeps topology); RebuildOnUpdate = true runs the deterministic rebuild instead (no quality decay, ~5 ms at 265 k triangles).
void Renderer::updateTLAS() { check(_device->ctx(), RebuildOnUpdate ? vkx_bvh_build(_device->ctx()) : vkx_bvh_refit(_device->ctx())); }

vkx_bvh_info Renderer::getTLAS() const {
    vkx_bvh_info info{};
    check(_device->ctx(), vkx_bvh_info_get(_device->ctx(), &info));
    return info;
}

} // namespace vkx
