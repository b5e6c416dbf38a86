"""Python reader/writer for the reference's binary `.scene` container and the host-side flattening that turns a
scene into the arrays the C ABI takes.

This is harness code (fixtures for tests and bench.py); the product's loader is the C++ one in
csrc/host/Scene.cpp. Both restate reference src/Scene.cpp:710-816 (save), :818-934 (loadScene), :936-961 (transform
propagation), :1077-1103 (bounds), src/Renderer.cpp:97-126 (offset table) and :512-551 (instance list).
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, field

import numpy as np

from .pods import INSTANCE_DTYPE, INSTANCE_STATIC, INVALID_TEXTURE, MATERIAL_DTYPE, OFFSET_DTYPE, VERTEX_DTYPE

MAGIC = 0x4E454353  # "SCEN"
CHUNK_JSON = 0x4E4F534A
CHUNK_BIN = 0x004E4942

IDENTITY16 = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]


@dataclass
class Mesh:
    name: str
    material: int
    vertices: np.ndarray  # VERTEX_DTYPE
    indices: np.ndarray  # uint32


@dataclass
class Entity:
    name: str
    transform: list = field(default_factory=lambda: list(IDENTITY16))  # column-major m[col][row]
    children: list = field(default_factory=list)
    mesh_renderer: tuple | None = None  # (meshIndex, materialIndex)


@dataclass
class SceneFile:
    materials: list = field(default_factory=list)  # glTF-style dicts
    entities: list = field(default_factory=list)
    meshes: list = field(default_factory=list)
    textures: list = field(default_factory=list)  # {"source": path, "format": VkFormat, "sampler": glTF sampler} (src/Scene.cpp:776-784)
    images: list = field(default_factory=list)  # decoded RGBA8 images [h, w, 4], parallel to textures (the reference decodes `source` with stb_image)


VK_FORMAT_R8G8B8A8_UNORM = 37
VK_FORMAT_R8G8B8A8_SRGB = 43


def write_pam(path, pixels):
    """Netpbm P7 (RGB_ALPHA, 8 bit): the image container of the synthetic scenes (no PNG encoder is assumed)."""
    px = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w, c = px.shape
    assert c == 4
    with open(path, "wb") as f:
        f.write(("P7\nWIDTH %d\nHEIGHT %d\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n" % (w, h)).encode("ascii"))
        f.write(px.tobytes())


def read_pam(path):
    data = open(path, "rb").read()
    end = data.index(b"ENDHDR\n") + 7
    hdr = dict(line.split(None, 1) for line in data[:end].decode("ascii").splitlines()[1:-1] if line and not line.startswith("#"))
    w, h, d = int(hdr["WIDTH"]), int(hdr["HEIGHT"]), int(hdr["DEPTH"])
    if d != 4 or int(hdr["MAXVAL"]) != 255:
        raise ValueError("%s: only 8-bit RGB_ALPHA PAM images are supported" % path)
    return np.frombuffer(data, dtype=np.uint8, count=w * h * 4, offset=end).reshape(h, w, 4).copy()


def add_skinned_instance(flat, mesh_entry, transform_rows=None, material=None):
    """A SkinnedMeshRendererComponent at the arena level (Renderer::allocateSkinnedMeshes / updateSkinnedMeshOffsetTable, reference
    src/Renderer.cpp:133-164): a bind-pose copy of the mesh's vertices appended to the vertex arena, an offset-table entry {material,
    that vertex offset, the mesh's index offset} and an instance with mask INSTANCE_SKINNED. Returns (flat', srcOffset, dstOffset, size)."""
    from .pods import INSTANCE_SKINNED

    out = dict(flat)
    src = int(flat["offsets"][mesh_entry]["vertexOffset"])
    size = int(flat["mesh_vertex_counts"][mesh_entry])
    dst = len(flat["vertices"])
    out["vertices"] = np.concatenate([flat["vertices"], flat["vertices"][src : src + size]])
    entry = np.zeros(1, dtype=OFFSET_DTYPE)
    entry[0] = (int(flat["offsets"][mesh_entry]["materialIndex"]) if material is None else material, dst, int(flat["offsets"][mesh_entry]["indexOffset"]))
    out["offsets"] = np.concatenate([flat["offsets"], entry])
    out["mesh_index_counts"] = np.concatenate([flat["mesh_index_counts"], flat["mesh_index_counts"][mesh_entry : mesh_entry + 1]])
    out["mesh_vertex_counts"] = np.concatenate([flat["mesh_vertex_counts"], np.array([size], dtype=np.uint32)])
    inst = np.zeros(1, dtype=INSTANCE_DTYPE)
    inst[0]["transform"] = np.asarray(transform_rows if transform_rows is not None else [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    inst[0]["meshEntry"] = len(flat["offsets"])
    inst[0]["mask"] = INSTANCE_SKINNED
    out["instances"] = np.concatenate([flat["instances"], inst])
    return out, src, dst, size


def write_image(path, pixels):
    """.png through Pillow when the source name asks for it (tests of the PNG decoder), Netpbm P7 otherwise."""
    if path.lower().endswith(".png"):
        from PIL import Image

        Image.fromarray(np.ascontiguousarray(pixels, dtype=np.uint8), "RGBA").save(path)
    else:
        write_pam(path, pixels)


def read_image(path):
    if path.lower().endswith(".png"):
        from PIL import Image

        return np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8).copy()
    return read_pam(path)


def texture_list(scene):
    """The scene's textures as the list vkx_scene_textures takes (pods.texture_array): image + VkFormat class + glTF sampler."""
    out = []
    for t, img in zip(scene.textures, scene.images):
        smp = t.get("sampler", {}) or {}
        out.append(
            {
                "pixels": img,
                "srgb": 1 if int(t.get("format", VK_FORMAT_R8G8B8A8_SRGB)) == VK_FORMAT_R8G8B8A8_SRGB else 0,
                "magFilter": int(smp.get("magFilter", 0)),
                "minFilter": int(smp.get("minFilter", 0)),
                "wrapS": int(smp.get("wrapS", 0)),
                "wrapT": int(smp.get("wrapT", 0)),
            }
        )
    return out


def _fmt(v):
    """Numbers as the reference's JSON writer prints them: std::to_string -> 6 decimals for floats (src/JSON.hpp:27-32)."""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return "%.6f" % float(v)
    if isinstance(v, str):
        return json.dumps(v)
    if isinstance(v, (list, tuple)):
        return "[" + ",".join(_fmt(x) for x in v) + "]"
    if isinstance(v, dict):
        return "{" + ",".join(json.dumps(k) + ":" + _fmt(x) for k, x in v.items()) + "}"
    raise TypeError(type(v))


def material_json(name, base_color=(1.0, 1.0, 1.0), metallic=0.0, roughness=1.0, emissive=(0.0, 0.0, 0.0)):
    return {
        "name": name,
        "pbrMetallicRoughness": {
            "baseColorFactor": [float(base_color[0]), float(base_color[1]), float(base_color[2]), 1.0],
            "metallicFactor": float(metallic),
            "roughnessFactor": float(roughness),
        },
        "emissiveFactor": [float(e) for e in emissive],
    }


def write_scene(path, scene: SceneFile):
    for t, img in zip(scene.textures, scene.images):  # texture sources are relative to the scene file (src/Scene.cpp:780,899)
        write_image(os.path.join(os.path.dirname(os.path.abspath(path)), t["source"]), img)
    root = {
        "materials": scene.materials,
        "entities": [],
        "meshes": [],
        "textures": scene.textures,
    }
    for e in scene.entities:
        ej = {"name": e.name, "transform": [float(x) for x in e.transform], "parent": -1, "children": [int(c) for c in e.children]}
        if e.mesh_renderer is not None:
            ej["meshRenderer"] = {"meshIndex": int(e.mesh_renderer[0]), "materialIndex": int(e.mesh_renderer[1])}
        root["entities"].append(ej)
    chunks = []
    for m in scene.meshes:
        offset = 1 + len(chunks)  # chunk indices count the JSON chunk as 0 (src/Scene.cpp:764-772)
        root["meshes"].append({"name": m.name, "material": int(m.material), "vertexArray": offset, "indexArray": offset + 1})
        chunks.append(np.ascontiguousarray(m.vertices, dtype=VERTEX_DTYPE).tobytes())
        chunks.append(np.ascontiguousarray(m.indices, dtype="<u4").tobytes())
    js = _fmt(root).encode("utf-8")
    total = 12 + 8 + len(js) + sum(8 + len(c) for c in chunks)
    with open(path, "wb") as f:
        f.write(struct.pack("<III", MAGIC, 0, total))
        f.write(struct.pack("<II", len(js), CHUNK_JSON))
        f.write(js)
        for c in chunks:
            f.write(struct.pack("<II", len(c), CHUNK_BIN))
            f.write(c)


def _to_f32(x):
    # Floats were parsed by std::from_chars<float> in the reference (src/JSON.cpp:183-186).
    return float(np.float32(x))


def read_scene(path) -> SceneFile:
    data = open(path, "rb").read()
    magic, version, length = struct.unpack_from("<III", data, 0)
    if magic != MAGIC:
        raise ValueError("not a .scene file (bad magic)")
    jlen, jtype = struct.unpack_from("<II", data, 12)
    if jtype != CHUNK_JSON:
        raise ValueError("first chunk is not JSON")
    root = json.loads(data[20 : 20 + jlen].decode("utf-8"))
    off = 20 + jlen
    buffers = []
    while off < length:
        clen, ctype = struct.unpack_from("<II", data, off)
        if ctype != CHUNK_BIN:
            raise ValueError("expected BIN chunk")
        buffers.append(data[off + 8 : off + 8 + clen])
        off += 8 + clen
    s = SceneFile(materials=root.get("materials", []), textures=root.get("textures", []))
    for e in root["entities"]:
        mr = e.get("meshRenderer")
        s.entities.append(
            Entity(
                name=e["name"],
                transform=[_to_f32(x) for x in e["transform"]],
                children=[int(c) for c in e.get("children", [])],
                mesh_renderer=(int(mr["meshIndex"]), int(mr["materialIndex"])) if mr else None,
            )
        )
    for m in root["meshes"]:
        v = np.frombuffer(buffers[m["vertexArray"] - 1], dtype=VERTEX_DTYPE).copy()
        i = np.frombuffer(buffers[m["indexArray"] - 1], dtype="<u4").copy()
        s.meshes.append(Mesh(m["name"], int(m.get("material", 0)), v, i))
    for t in s.textures:
        s.images.append(read_image(os.path.join(os.path.dirname(os.path.abspath(path)), t["source"])))
    return s


def _tex_index(obj, key):
    t = obj.get(key)
    if t is None:
        return INVALID_TEXTURE
    return int(t.get("index", -1)) & 0xFFFFFFFF


def materials_array(materials_json) -> np.ndarray:
    """parseMaterial, reference src/vulkan/Material.cpp:7-28."""
    out = np.zeros(len(materials_json), dtype=MATERIAL_DTYPE)
    for i, m in enumerate(materials_json):
        pbr = m.get("pbrMetallicRoughness", {})
        out[i]["baseColorFactor"] = [np.float32(x) for x in pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])[:3]] if "pbrMetallicRoughness" in m else [1, 1, 1]
        out[i]["metallicFactor"] = np.float32(pbr.get("metallicFactor", 1.0))
        out[i]["roughnessFactor"] = np.float32(pbr.get("roughnessFactor", 1.0))
        out[i]["emissiveFactor"] = [np.float32(x) for x in m.get("emissiveFactor", [0.0, 0.0, 0.0])]
        out[i]["albedoTexture"] = _tex_index(pbr, "baseColorTexture")
        out[i]["metallicRoughnessTexture"] = _tex_index(pbr, "metallicRoughnessTexture")
        out[i]["normalTexture"] = _tex_index(m, "normalTexture")
        out[i]["emissiveTexture"] = _tex_index(m, "emissiveTexture")
    return out


def _mat4_mul(a, b):
    """glm mat4 * mat4 in fp32 (column-major flat lists): col_j = ((A0*b0j + A1*b1j) + A2*b2j) + A3*b3j."""
    a = np.asarray(a, dtype=np.float32).reshape(4, 4)  # a[col][row]
    b = np.asarray(b, dtype=np.float32).reshape(4, 4)
    r = np.zeros((4, 4), dtype=np.float32)
    for j in range(4):
        acc = a[0] * b[j][0]
        acc = acc + a[1] * b[j][1]
        acc = acc + a[2] * b[j][2]
        acc = acc + a[3] * b[j][3]
        r[j] = acc
    return r.reshape(16)


def _mat4_mul_vec(m, v):
    """glm mat4 * vec4: (m0*v0 + m1*v1) + (m2*v2 + m3*v3), fp32."""
    m = np.asarray(m, dtype=np.float32).reshape(4, 4)
    v = np.asarray(v, dtype=np.float32)
    return (m[0] * v[0] + m[1] * v[1]) + (m[2] * v[2] + m[3] * v[3])


def flatten(scene: SceneFile) -> dict:
    """Scene -> the arrays vkx_scene_upload takes, plus the scene bounds used as the probe-grid extents."""
    n = len(scene.entities)
    parent = [-1] * n
    for i, e in enumerate(scene.entities):
        for c in e.children:
            parent[c] = i
    root = next(i for i in range(n) if parent[i] < 0)  # first entity without a parent (src/Scene.cpp:924-928)
    # Scene::update: globalTransform = parent.globalTransform * child.transform, starting from the root's cached
    # (identity) globalTransform (src/Scene.cpp:943-954; quirk A.10.2: the root's own local transform is not folded in).
    glob = [np.array(IDENTITY16, dtype=np.float32) for _ in range(n)]
    # visitNode accumulation (includes the root's local transform), used for the bounds (src/Scene.cpp:1096-1102)
    vis = [None] * n
    stack = [(root, np.array(IDENTITY16, dtype=np.float32), np.array(IDENTITY16, dtype=np.float32))]
    while stack:
        i, pg, pv = stack.pop()
        e = scene.entities[i]
        if i != root:
            glob[i] = _mat4_mul(pg, e.transform)
        vis[i] = _mat4_mul(pv, e.transform)
        for c in e.children:
            stack.append((c, glob[i], vis[i]))

    # offset table: tight packing in mesh order (src/Renderer.cpp:97-126)
    offsets = np.zeros(len(scene.meshes), dtype=OFFSET_DTYPE)
    counts = np.zeros(len(scene.meshes), dtype=np.uint32)
    vo = io = 0
    for mi, m in enumerate(scene.meshes):
        offsets[mi] = (m.material, vo, io)
        counts[mi] = len(m.indices)
        vo += len(m.vertices)
        io += len(m.indices)
    vertices = np.concatenate([np.asarray(m.vertices, dtype=VERTEX_DTYPE) for m in scene.meshes]) if scene.meshes else np.zeros(0, VERTEX_DTYPE)
    indices = np.concatenate([np.asarray(m.indices, dtype=np.uint32) for m in scene.meshes]) if scene.meshes else np.zeros(0, np.uint32)

    # instance list: one per MeshRendererComponent, stable-sorted by (materialIndex, meshIndex) with the entity
    # array index as final key (src/Renderer.cpp:512-551; SURVEY appendix B decree)
    rend = [(e.mesh_renderer[1], e.mesh_renderer[0], i) for i, e in enumerate(scene.entities) if e.mesh_renderer is not None]
    rend.sort()
    instances = np.zeros(len(rend), dtype=INSTANCE_DTYPE)
    for k, (_, mesh_index, ei) in enumerate(rend):
        g = glob[ei].reshape(4, 4)  # g[col][row]
        rows = np.zeros(12, dtype=np.float32)
        for r in range(3):
            for c in range(4):
                rows[4 * r + c] = g[c][r]
        instances[k]["transform"] = rows
        instances[k]["meshEntry"] = mesh_index
        instances[k]["mask"] = INSTANCE_STATIC

    # bounds: union of (T * meshBounds) with the two-corner transform (src/Bounds.hpp:38-45)
    bmin = bmax = None
    for i, e in enumerate(scene.entities):
        if e.mesh_renderer is None:
            continue
        m = scene.meshes[e.mesh_renderer[0]]
        lo = m.vertices["pos"].min(axis=0)
        hi = m.vertices["pos"].max(axis=0)
        a = _mat4_mul_vec(vis[i], [lo[0], lo[1], lo[2], 1.0])[:3]
        b = _mat4_mul_vec(vis[i], [hi[0], hi[1], hi[2], 1.0])[:3]
        l2, h2 = np.minimum(a, b), np.maximum(a, b)
        bmin = l2 if bmin is None else np.minimum(bmin, l2)
        bmax = h2 if bmax is None else np.maximum(bmax, h2)
    return {
        "vertices": vertices,
        "indices": indices,
        "offsets": offsets,
        "mesh_index_counts": counts,
        "mesh_vertex_counts": np.array([len(m.vertices) for m in scene.meshes], dtype=np.uint32),
        "materials": materials_array(scene.materials),
        "instances": instances,
        "bounds_min": np.asarray(bmin, dtype=np.float32),
        "bounds_max": np.asarray(bmax, dtype=np.float32),
        **({"textures": texture_list(scene)} if scene.textures else {}),
    }
