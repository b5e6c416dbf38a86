// Host-side scene: the reference's Scene class reduced to what the hot path consumes (reference src/Scene.hpp:87-118):
// the binary .scene container (loadScene / save), the node hierarchy with cached global transforms (update), bounds.
// and the texture list (decoded by Image.cpp). glTF/OBJ importers, skinning and animation are out of scope (component 5).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../../include/vkx.h"
#include "Image.hpp"
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

// One entry of the reference's global Textures list (reference src/Resources.hpp, filled by Scene::loadScene src/Scene.cpp:895-904):
// image source relative to the scene file, VkFormat (43 = R8G8B8A8_SRGB, 37 = R8G8B8A8_UNORM) and the glTF sampler description.
struct TextureDesc {
    std::string source;
    uint32_t format = 43;
    uint32_t magFilter = 0, minFilter = 0, wrapS = 0, wrapT = 0; // 0 = absent (the reference's defaults 9729 / 10497 apply)
    Image image;                                                  // decoded at load time (the reference decodes in uploadTextures)
};

// SkinnedMeshRendererComponent + the mesh's SkinVertexData (reference src/Scene.hpp:39-45, src/vulkan/Mesh.hpp:18-21), reduced to what
// vertexSkinning.comp consumes. The animation system that produces joint poses is out of scope: poses are handed to
// Renderer::updateSkinnedVertexBuffer.
struct SkinnedMeshRenderer {
    int node = 0;                     // entity whose global transform places the instance
    uint32_t meshIndex = 0, materialIndex = 0;
    std::vector<uint16_t> joints;     // 4 per vertex
    std::vector<float> weights;       // 4 per vertex
    uint32_t indexIntoOffsetTable = 0; // set by Renderer::allocateSkinnedMeshes
    uint32_t vertexOffset = 0;         // first vertex of the skinned copy in Renderer::Vertices
};

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
    std::vector<TextureDesc>& getTextures() { return _textures; }
    const std::vector<TextureDesc>& getTextures() const { return _textures; }
    std::vector<SkinnedMeshRenderer>& getSkinnedRenderers() { return _skinned; }
    const std::vector<SkinnedMeshRenderer>& getSkinnedRenderers() const { return _skinned; }
    int getRoot() const { return _root; }
    void markDirty() { _dirty = true; }

  private:
    std::vector<Mesh> _meshes;
    std::vector<NodeComponent> _nodes;
    std::vector<MaterialDesc> _materials;
    std::vector<TextureDesc> _textures;
    std::vector<SkinnedMeshRenderer> _skinned;
    int _root = -1;
    bool _dirty = false;
    Bounds _bounds;
};

} // namespace vkx
