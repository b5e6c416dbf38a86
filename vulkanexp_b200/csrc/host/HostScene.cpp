// C ABI over the facade's host-side scene code (no GPU involved): the .scene loader / writer, Scene::update's transform propagation,
// Renderer::allocateMeshes (arenas + offset table) and Renderer::createTLAS (instance list), so that hosts in other languages get
// the arrays vkx_scene_textures / vkx_scene_upload take from the product's own loader (reference src/Scene.cpp:818-961,
// src/Renderer.cpp:57-131,512-575) instead of re-implementing it.
#include <cstring>
#include <new>
#include "Renderer.hpp"

struct vkx_host_scene {
    vkx::Scene scene;
    vkx::Renderer renderer;
    std::vector<vkx_material> materials;
};

extern "C" {

int vkx_host_scene_load(const char* path, vkx_host_scene** out) {
    if (!path || !out) return VKX_E_INVALID;
    *out = nullptr;
    vkx_host_scene* h = new (std::nothrow) vkx_host_scene();
    if (!h) return VKX_E_NOMEM;
    try {
        if (!h->scene.load(path)) { delete h; return VKX_E_INVALID; }
        h->scene.update();
        h->renderer.setScene(h->scene);
        h->renderer.allocateMeshes();
        h->renderer.createTLAS();
        for (const auto& m : h->scene.getMaterials()) h->materials.push_back(m.properties);
    } catch (const std::exception&) { delete h; return VKX_E_INVALID; }
    *out = h;
    return VKX_OK;
}

void vkx_host_scene_free(vkx_host_scene* h) { delete h; }

int vkx_host_scene_counts(const vkx_host_scene* h, size_t counts[6]) {
    if (!h || !counts) return VKX_E_INVALID;
    counts[0] = h->renderer.Vertices.size(); counts[1] = h->renderer.Indices.size(); counts[2] = h->renderer.OffsetTable.size();
    counts[3] = h->materials.size(); counts[4] = h->renderer.getInstances().size(); counts[5] = h->scene.getTextures().size();
    return VKX_OK;
}

int vkx_host_scene_copy(const vkx_host_scene* h, vkx_vertex* vertices, uint32_t* indices, vkx_offset_entry* offsets, uint32_t* meshIndexCounts, vkx_material* materials,
                        vkx_instance* instances, float boundsMinMax[6]) {
    if (!h) return VKX_E_INVALID;
    const vkx::Renderer& r = h->renderer;
    if (vertices && !r.Vertices.empty()) std::memcpy(vertices, r.Vertices.data(), r.Vertices.size() * sizeof(vkx_vertex));
    if (indices && !r.Indices.empty()) std::memcpy(indices, r.Indices.data(), r.Indices.size() * 4);
    if (offsets && !r.OffsetTable.empty()) std::memcpy(offsets, r.OffsetTable.data(), r.OffsetTable.size() * sizeof(vkx_offset_entry));
    if (meshIndexCounts && !r.MeshIndexCounts.empty()) std::memcpy(meshIndexCounts, r.MeshIndexCounts.data(), r.MeshIndexCounts.size() * 4);
    if (materials && !h->materials.empty()) std::memcpy(materials, h->materials.data(), h->materials.size() * sizeof(vkx_material));
    if (instances && !r.getInstances().empty()) std::memcpy(instances, r.getInstances().data(), r.getInstances().size() * sizeof(vkx_instance));
    if (boundsMinMax) {
        const vkx::Bounds& b = h->scene.getBounds();
        boundsMinMax[0] = b.min.x; boundsMinMax[1] = b.min.y; boundsMinMax[2] = b.min.z; boundsMinMax[3] = b.max.x; boundsMinMax[4] = b.max.y; boundsMinMax[5] = b.max.z;
    }
    return VKX_OK;
}

int vkx_host_scene_texture(const vkx_host_scene* h, size_t index, vkx_texture* desc) {
    if (!h || !desc || index >= h->scene.getTextures().size()) return VKX_E_INVALID;
    const vkx::TextureDesc& t = h->scene.getTextures()[index];
    desc->pixels = t.image.pixels.data(); desc->width = t.image.width; desc->height = t.image.height;
    desc->srgb = t.format == 43u ? 1u : 0u; desc->magFilter = t.magFilter; desc->minFilter = t.minFilter; desc->wrapS = t.wrapS; desc->wrapT = t.wrapT;
    return VKX_OK;
}

int vkx_host_scene_save(const vkx_host_scene* h, const char* path) {
    if (!h || !path) return VKX_E_INVALID;
    return h->scene.save(path) ? VKX_OK : VKX_E_INVALID;
}

} // extern "C"
