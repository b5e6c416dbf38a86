// Host-side scene: the reference's Scene class reduced to what the hot path consumes (reference src/Scene.hpp:87-118):
// the binary .scene container (loadScene / save), the node hierarchy with cached global transforms (update), bounds.
// glTF/OBJ importers, skinning and animation are out of scope (component 5).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../../include/vkx.h"
#include "Math.hpp"

namespace vkx {

struct Mesh {
    std::string name;
    uint32_t defaultMaterialIndex = 0;
    std::vector<vkx_vertex> vertices;
    std::vector<uint32_t> indices;
    Bounds bounds;
    uint32_t indexIntoOffsetTable = 0; // set by Renderer::allocateMeshes
    void computeBounds();              // reference src/vulkan/Mesh.cpp:96-106
    bool isValid() const { return !vertices.empty() && !indices.empty(); }
};

struct NodeComponent { // reference src/Scene.hpp:15-27
    std::string name;
    mat4 transform;
    mat4 globalTransform;
    int parent = -1;
    std::vector<int> children;
    bool hasMeshRenderer = false;
    uint32_t meshIndex = 0, materialIndex = 0; // MeshRendererComponent
};

struct MaterialDesc { std::string name; vkx_material properties; };

class Scene {
  public:
    bool load(const std::string& path) { return loadScene(path); }
    bool loadScene(const std::string& path);      // reference src/Scene.cpp:818-934
    bool save(const std::string& path) const;     // reference src/Scene.cpp:710-816
    bool update(float deltaTime = 0.f);           // transform propagation, reference src/Scene.cpp:936-961
    const Bounds& computeBounds();                // reference src/Scene.cpp:1077-1094
    const Bounds& getBounds() const { return _bounds; }
    std::vector<Mesh>& getMeshes() { return _meshes; }
    const std::vector<Mesh>& getMeshes() const { return _meshes; }
    std::vector<NodeComponent>& getNodes() { return _nodes; }
    const std::vector<NodeComponent>& getNodes() const { return _nodes; }
    std::vector<MaterialDesc>& getMaterials() { return _materials; }
    const std::vector<MaterialDesc>& getMaterials() const { return _materials; }
    int getRoot() const { return _root; }
    void markDirty() { _dirty = true; }

  private:
    std::vector<Mesh> _meshes;
    std::vector<NodeComponent> _nodes;
    std::vector<MaterialDesc> _materials;
    int _root = -1;
    bool _dirty = false;
    Bounds _bounds;
};

} // namespace vkx
