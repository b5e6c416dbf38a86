#include "Scene.hpp"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <functional>
#include "Json.hpp"

namespace vkx {

namespace {
constexpr uint32_t kMagic = 0x4e454353, kChunkJson = 0x4E4F534A, kChunkBin = 0x004E4942;
struct Header { uint32_t magic, version, length; };
struct ChunkHeader { uint32_t length, type; };

uint32_t textureIndex(const Json& parent, const char* key) { // "index" is written as -1 for "none" (src/vulkan/Material.cpp:30-53)
    if (!parent.contains(key)) return VKX_INVALID_TEXTURE;
    const Json& t = parent[key];
    return uint32_t(t.get("index", -1));
}
} // namespace

void Mesh::computeBounds() {
    if (vertices.empty()) { bounds = Bounds{}; return; }
    vec3 lo{vertices[0].pos[0], vertices[0].pos[1], vertices[0].pos[2]}, hi = lo;
    for (const auto& v : vertices) { vec3 p{v.pos[0], v.pos[1], v.pos[2]}; lo = vkx::min(lo, p); hi = vkx::max(hi, p); }
    bounds = {lo, hi};
}

bool Scene::loadScene(const std::string& path) {
    _meshes.clear(); _nodes.clear(); _materials.clear(); _textures.clear(); _root = -1;
    std::ifstream file(path, std::ios::binary | std::ios::ate);
    if (!file) { std::fprintf(stderr, "Scene::loadScene error: Could not open file '%s'.\n", path.c_str()); return false; }
    std::streamsize size = file.tellg();
    file.seekg(0, std::ios::beg);
    std::vector<char> buf(static_cast<size_t>(size));
    if (!file.read(buf.data(), size) || size < std::streamsize(sizeof(Header) + sizeof(ChunkHeader))) return false;
    Header header; std::memcpy(&header, buf.data(), sizeof(header));
    if (header.magic != kMagic) { std::fprintf(stderr, "Scene::loadScene: '%s' is not a .scene file.\n", path.c_str()); return false; }
    ChunkHeader jc; std::memcpy(&jc, buf.data() + sizeof(Header), sizeof(jc));
    if (jc.type != kChunkJson) return false;
    size_t off = sizeof(Header) + sizeof(ChunkHeader);
    if (size_t(jc.length) > buf.size() - off) { std::fprintf(stderr, "Scene::loadScene: '%s' is truncated (JSON chunk of %u bytes, %zu left).\n", path.c_str(), jc.length, buf.size() - off); return false; }
    Json root;
    try { root = Json::parse(buf.data() + off, jc.length); } catch (const std::exception& e) { std::fprintf(stderr, "Scene::loadScene: %s\n", e.what()); return false; }
    off += jc.length;
    std::vector<std::pair<const char*, size_t>> chunks;
    while (off + sizeof(ChunkHeader) <= header.length && off + sizeof(ChunkHeader) <= buf.size()) {
        ChunkHeader ch; std::memcpy(&ch, buf.data() + off, sizeof(ch));
        if (ch.type != kChunkBin || off + sizeof(ch) + ch.length > buf.size()) return false;
        chunks.emplace_back(buf.data() + off + sizeof(ch), ch.length);
        off += sizeof(ch) + ch.length;
    }
    try {
        for (const Json& n : root["entities"].items()) {
            NodeComponent node;
            node.name = n["name"].asString();
            const Json& t = n["transform"];
            for (int i = 0; i < 16; ++i) node.transform.m[i / 4][i % 4] = t[i].asFloat();
            if (n.contains("children")) for (const Json& c : n["children"].items()) node.children.push_back(c.asInt());
            if (n.contains("meshRenderer")) { node.hasMeshRenderer = true; node.meshIndex = uint32_t(n["meshRenderer"]["meshIndex"].asInt()); node.materialIndex = uint32_t(n["meshRenderer"]["materialIndex"].asInt()); }
            _nodes.push_back(std::move(node));
        }
        // The hierarchy must be a forest: child indices in range, one parent each, no node its own ancestor (a file where a node lists
        // itself or an ancestor would send Scene::update / computeBounds into unbounded recursion).
        for (size_t i = 0; i < _nodes.size(); ++i)
            for (int c : _nodes[i].children) {
                if (c < 0 || size_t(c) >= _nodes.size() || size_t(c) == i) throw std::runtime_error("entity " + std::to_string(i) + ": child index " + std::to_string(c) + " is out of range or the node itself");
                if (_nodes[size_t(c)].parent != -1) throw std::runtime_error("entity " + std::to_string(c) + " has two parents");
                _nodes[size_t(c)].parent = int(i);
            }
        for (size_t i = 0; i < _nodes.size(); ++i) { // every parent chain ends at a root within |nodes| steps
            size_t steps = 0;
            for (int p = _nodes[i].parent; p != -1; p = _nodes[size_t(p)].parent) if (++steps > _nodes.size()) throw std::runtime_error("entity hierarchy contains a cycle");
        }
        if (root.contains("materials"))
            for (const Json& m : root["materials"].items()) { // parseMaterial, src/vulkan/Material.cpp:7-28
                MaterialDesc d; d.name = m.contains("name") ? m["name"].asString() : "NoName";
                vkx_material& p = d.properties;
                p.metallicFactor = 1.0f; p.roughnessFactor = 1.0f; p.baseColorFactor[0] = p.baseColorFactor[1] = p.baseColorFactor[2] = 1.0f;
                p.emissiveFactor[0] = p.emissiveFactor[1] = p.emissiveFactor[2] = 0.0f;
                p.albedoTexture = p.normalTexture = p.metallicRoughnessTexture = p.emissiveTexture = VKX_INVALID_TEXTURE;
                if (m.contains("pbrMetallicRoughness")) {
                    const Json& pbr = m["pbrMetallicRoughness"];
                    if (pbr.contains("baseColorFactor")) for (int i = 0; i < 3; ++i) p.baseColorFactor[i] = pbr["baseColorFactor"][size_t(i)].asFloat();
                    p.metallicFactor = pbr.get("metallicFactor", 1.0f);
                    p.roughnessFactor = pbr.get("roughnessFactor", 1.0f);
                    p.albedoTexture = textureIndex(pbr, "baseColorTexture");
                    p.metallicRoughnessTexture = textureIndex(pbr, "metallicRoughnessTexture");
                }
                if (m.contains("emissiveFactor")) for (int i = 0; i < 3; ++i) p.emissiveFactor[i] = m["emissiveFactor"][size_t(i)].asFloat();
                p.emissiveTexture = textureIndex(m, "emissiveTexture");
                p.normalTexture = textureIndex(m, "normalTexture");
                _materials.push_back(std::move(d));
            }
        for (const Json& m : root["meshes"].items()) {
            Mesh mesh; mesh.name = m["name"].asString(); mesh.defaultMaterialIndex = uint32_t(m.get("material", 0));
            const auto& vb = chunks.at(size_t(m["vertexArray"].asInt() - 1)); // chunk indices count the JSON chunk as 0
            const auto& ib = chunks.at(size_t(m["indexArray"].asInt() - 1));
            mesh.vertices.resize(vb.second / sizeof(vkx_vertex)); std::memcpy(mesh.vertices.data(), vb.first, mesh.vertices.size() * sizeof(vkx_vertex));
            mesh.indices.resize(ib.second / 4); std::memcpy(mesh.indices.data(), ib.first, mesh.indices.size() * 4);
            mesh.computeBounds();
            _meshes.push_back(std::move(mesh));
        }
        if (root.contains("textures")) { // src/Scene.cpp:895-904; sources are relative to the scene file
            const size_t slash = path.find_last_of("/\\");
            const std::string dir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
            for (const Json& t : root["textures"].items()) {
                TextureDesc d;
                d.source = t["source"].asString();
                d.format = uint32_t(t.get("format", 43));
                if (t.contains("sampler")) {
                    const Json& smp = t["sampler"];
                    d.magFilter = uint32_t(smp.get("magFilter", 0)); d.minFilter = uint32_t(smp.get("minFilter", 0));
                    d.wrapS = uint32_t(smp.get("wrapS", 0)); d.wrapT = uint32_t(smp.get("wrapT", 0));
                }
                std::string err;
                if (!d.image.load(dir + d.source, &err)) { // the reference substitutes a blank image for textures it cannot read (src/Scene.cpp:699)
                    std::fprintf(stderr, "Scene::loadScene: texture '%s': %s (replaced by a blank image).\n", d.source.c_str(), err.c_str());
                    d.image = Image::blank();
                }
                _textures.push_back(std::move(d));
            }
        }
    } catch (const std::exception& e) { std::fprintf(stderr, "Scene::loadScene: malformed scene '%s': %s\n", path.c_str(), e.what()); return false; }
    for (size_t i = 0; i < _nodes.size(); ++i) if (_nodes[i].parent < 0) { _root = int(i); break; } // src/Scene.cpp:924-928
    _dirty = true;
    computeBounds();
    return _root >= 0;
}

bool Scene::save(const std::string& path) const {
    Json root = Json::object();
    Json mats = Json::array();
    for (const auto& m : _materials) {
        Json j = Json::object(); j["name"] = m.name;
        Json pbr = Json::object();
        Json bc = Json::array(); for (int i = 0; i < 3; ++i) bc.push(Json(m.properties.baseColorFactor[i])); bc.push(Json(1.0f));
        pbr["baseColorFactor"] = bc; pbr["metallicFactor"] = Json(m.properties.metallicFactor); pbr["roughnessFactor"] = Json(m.properties.roughnessFactor);
        Json t1 = Json::object(); t1["index"] = Json(int(m.properties.albedoTexture)); pbr["baseColorTexture"] = t1;
        Json t2 = Json::object(); t2["index"] = Json(int(m.properties.metallicRoughnessTexture)); pbr["metallicRoughnessTexture"] = t2;
        j["pbrMetallicRoughness"] = pbr;
        Json t3 = Json::object(); t3["index"] = Json(int(m.properties.normalTexture)); j["normalTexture"] = t3;
        Json ef = Json::array(); for (int i = 0; i < 3; ++i) ef.push(Json(m.properties.emissiveFactor[i])); j["emissiveFactor"] = ef;
        Json t4 = Json::object(); t4["index"] = Json(int(m.properties.emissiveTexture)); j["emissiveTexture"] = t4;
        mats.push(j);
    }
    root["materials"] = mats;
    Json ents = Json::array();
    for (const auto& n : _nodes) {
        Json j = Json::object(); j["name"] = n.name;
        Json t = Json::array(); for (int i = 0; i < 16; ++i) t.push(Json(n.transform.m[i / 4][i % 4])); j["transform"] = t;
        j["parent"] = Json(-1);
        Json ch = Json::array(); for (int c : n.children) ch.push(Json(c)); j["children"] = ch;
        if (n.hasMeshRenderer) { Json mr = Json::object(); mr["meshIndex"] = Json(int(n.meshIndex)); mr["materialIndex"] = Json(int(n.materialIndex)); j["meshRenderer"] = mr; }
        ents.push(j);
    }
    root["entities"] = ents;
    Json meshes = Json::array();
    int chunk = 1;
    for (const auto& m : _meshes) {
        Json j = Json::object(); j["name"] = m.name; j["material"] = Json(int(m.defaultMaterialIndex)); j["vertexArray"] = Json(chunk); j["indexArray"] = Json(chunk + 1);
        chunk += 2; meshes.push(j);
    }
    root["meshes"] = meshes;
    Json texs = Json::array(); // src/Scene.cpp:776-784 (the image files themselves are not rewritten)
    for (const auto& t : _textures) {
        Json j = Json::object(); j["source"] = t.source; j["format"] = Json(int(t.format));
        Json smp = Json::object();
        if (t.magFilter) smp["magFilter"] = Json(int(t.magFilter));
        if (t.minFilter) smp["minFilter"] = Json(int(t.minFilter));
        if (t.wrapS) smp["wrapS"] = Json(int(t.wrapS));
        if (t.wrapT) smp["wrapT"] = Json(int(t.wrapT));
        j["sampler"] = smp;
        texs.push(j);
    }
    root["textures"] = texs;
    const std::string js = root.toString();
    uint32_t total = uint32_t(sizeof(Header) + sizeof(ChunkHeader) + js.size());
    for (const auto& m : _meshes) total += uint32_t(2 * sizeof(ChunkHeader) + m.vertices.size() * sizeof(vkx_vertex) + m.indices.size() * 4);
    std::ofstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "Scene::save error: Could not open '%s' file for writing.\n", path.c_str()); return false; }
    Header h{kMagic, 0, total}; f.write(reinterpret_cast<const char*>(&h), sizeof(h));
    ChunkHeader jc{uint32_t(js.size()), kChunkJson}; f.write(reinterpret_cast<const char*>(&jc), sizeof(jc)); f.write(js.data(), std::streamsize(js.size()));
    for (const auto& m : _meshes) {
        ChunkHeader vc{uint32_t(m.vertices.size() * sizeof(vkx_vertex)), kChunkBin}; f.write(reinterpret_cast<const char*>(&vc), sizeof(vc)); f.write(reinterpret_cast<const char*>(m.vertices.data()), vc.length);
        ChunkHeader ic{uint32_t(m.indices.size() * 4), kChunkBin}; f.write(reinterpret_cast<const char*>(&ic), sizeof(ic)); f.write(reinterpret_cast<const char*>(m.indices.data()), ic.length);
    }
    return bool(f);
}

bool Scene::update(float) {
    if (!_dirty || _root < 0) return false;
    // Propagation starts from the root's cached globalTransform; the root's own local transform is never folded in
    // (reference src/Scene.cpp:943-954).
    std::function<void(const mat4&, int)> rec = [&](const mat4& parentTransform, int parent) {
        for (int c : _nodes[size_t(parent)].children) {
            NodeComponent& child = _nodes[size_t(c)];
            child.globalTransform = parentTransform * child.transform;
            rec(child.globalTransform, c);
        }
    };
    rec(_nodes[size_t(_root)].globalTransform, _root);
    computeBounds();
    _dirty = false;
    return true;
}

const Bounds& Scene::computeBounds() {
    bool init = false;
    std::function<void(int, mat4)> visit = [&](int e, mat4 transform) { // visitNode, src/Scene.cpp:1096-1102
        const NodeComponent& node = _nodes[size_t(e)];
        transform = transform * node.transform;
        for (int c : node.children) visit(c, transform);
        if (node.hasMeshRenderer && node.meshIndex < _meshes.size()) {
            Bounds b = transform * _meshes[node.meshIndex].bounds;
            if (!init) { _bounds = b; init = true; } else _bounds += b;
        }
    };
    if (_root >= 0) visit(_root, mat4{});
    return _bounds;
}

} // namespace vkx
