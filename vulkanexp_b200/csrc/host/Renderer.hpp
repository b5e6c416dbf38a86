// Geometry side of the reference's Renderer (reference src/Renderer.hpp:39-86), above the C ABI: mesh arenas + offset
// table (allocateMeshes), instance list (createTLAS) and the acceleration-structure build, which here is the device-side
// SAH wide-BVH build. Skinning / animation / TLAS refit are out of scope (component 2).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "Scene.hpp"

namespace vkx {

struct Error : std::runtime_error { // the reference throws std::runtime_error from VK_CHECK (src/vulkan/VkTools.hpp:39-47)
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
void check(vkx_ctx* ctx, int rc);

class Device { // stands in for the reference's Device: owns the vkx context of one GPU
  public:
    explicit Device(int index = 0);
    ~Device();
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    vkx_ctx* ctx() const { return _ctx; }
  private:
    vkx_ctx* _ctx = nullptr;
};

class Renderer {
  public:
    using OffsetEntry = vkx_offset_entry; // reference src/Renderer.hpp:21-25
    void setDevice(const Device& device) { _device = &device; }
    void setScene(Scene& scene) { _scene = &scene; }
    void allocateMeshes();                // arenas + offset table, reference src/Renderer.cpp:57-131
    void createAccelerationStructures();  // reference src/Renderer.cpp:272-449 (+ createTLAS)
    void createTLAS();                    // instance list, reference src/Renderer.cpp:525-642
    void destroyTLAS() { _instances.clear(); }
    // reference src/Renderer.cpp:671-742: new instance transforms after Scene::update(), then the structure update (here: the
    // deterministic rebuild of the world-space BVH; the reference refits its TLAS)
    void updateAccelerationStructureInstances();
    bool RebuildOnUpdate = false; // updateTLAS: false = topology-preserving refit like the reference, true = full rebuild
    void updateTLAS();
    void onHierarchicalChanges() { updateAccelerationStructureInstances(); updateTLAS(); }
    // Skinned meshes (reference src/Renderer.cpp:133-164, 201-240, 644-669): allocateMeshes appends a bind-pose copy of the vertices of
    // every Scene::getSkinnedRenderers() entry + its offset-table entry; createTLAS adds its instance (mask SKINNED);
    // updateSkinnedVertexBuffer runs the skinning kernel with one pose array per skinned renderer (jointPoses[r][j], the reference
    // computes them from the animated scene graph); updateSkinnedBLAS rebuilds the structure.
    void allocateSkinnedMeshes();
    bool updateSkinnedVertexBuffer(const std::vector<std::vector<mat4>>& jointPoses);
    bool updateSkinnedBLAS();
    vkx_bvh_info getTLAS() const;         // the reference returns the TLAS handle; here: the wide-BVH description

    std::vector<vkx_vertex> Vertices;     // public arenas, as in the reference
    std::vector<uint32_t> Indices;
    std::vector<OffsetEntry> OffsetTable;
    std::vector<uint32_t> MeshIndexCounts;
    const std::vector<vkx_instance>& getInstances() const { return _instances; }
    const Device& getDevice() const { return *_device; }
    Scene& getScene() const { return *_scene; }

  private:
    const Device* _device = nullptr;
    Scene* _scene = nullptr;
    std::vector<vkx_instance> _instances;
};

} // namespace vkx
