"""POD layouts shared by the ctypes bindings (product library and test oracle).

Byte-identical to include/vkx.h, which in turn mirrors the reference's structs:
Vertex (reference src/vulkan/Vertex.hpp:8-15), Material::Properties (src/vulkan/Material.hpp:16-25),
OffsetEntry (src/Renderer.hpp:21-25), GridInfo (src/IrradianceProbes.hpp:49-60), LightBuffer (src/Light.hpp:6-9),
CameraBuffer (src/Editor.hpp:58-63).
"""
import ctypes as C

import numpy as np

VERTEX_DTYPE = np.dtype(
    [("pos", "<f4", 3), ("color", "<f4", 3), ("normal", "<f4", 3), ("tangent", "<f4", 4), ("texCoord", "<f4", 2), ("padding", "<u4")]
)
MATERIAL_DTYPE = np.dtype(
    [
        ("metallicFactor", "<f4"),
        ("roughnessFactor", "<f4"),
        ("baseColorFactor", "<f4", 3),
        ("emissiveFactor", "<f4", 3),
        ("albedoTexture", "<u4"),
        ("normalTexture", "<u4"),
        ("metallicRoughnessTexture", "<u4"),
        ("emissiveTexture", "<u4"),
    ]
)
OFFSET_DTYPE = np.dtype([("materialIndex", "<u4"), ("vertexOffset", "<u4"), ("indexOffset", "<u4")])
INSTANCE_DTYPE = np.dtype([("transform", "<f4", 12), ("meshEntry", "<u4"), ("mask", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("instance", "<u4"), ("primitive", "<u4"), ("u", "<f4"), ("v", "<f4")])
NODE_DTYPE = np.dtype(
    [
        ("p", "<f4", 3),
        ("e", "u1", 3),
        ("imask", "u1"),
        ("childBase", "<u4"),
        ("primBase", "<u4"),
        ("valid", "<u4"),  # BVH spec v2: bits 24..31 imask, bits [3s, 3s + count) leaf slot s
        ("pad", "<u4"),
        ("qlo", "u1", (3, 8)),
        ("qhi", "u1", (3, 8)),
    ]
)
TRI_DTYPE = np.dtype([("v0", "<f4", 3), ("e1", "<f4", 3), ("e2", "<f4", 3), ("inst", "<u4"), ("prim", "<u4"), ("pad", "<u4")])

assert VERTEX_DTYPE.itemsize == 64 and MATERIAL_DTYPE.itemsize == 48 and OFFSET_DTYPE.itemsize == 12
assert INSTANCE_DTYPE.itemsize == 56 and HIT_DTYPE.itemsize == 20 and NODE_DTYPE.itemsize == 80 and TRI_DTYPE.itemsize == 48

INVALID_TEXTURE = 0xFFFFFFFF
INSTANCE_STATIC, INSTANCE_DYNAMIC, INSTANCE_SKINNED = 1, 2, 4
MAX_RAYS_PER_PROBE = 256


class GridInfo(C.Structure):
    _fields_ = [
        ("extentMin", C.c_float * 3),
        ("depthSharpness", C.c_float),
        ("extentMax", C.c_float * 3),
        ("hysteresis", C.c_float),
        ("resolution", C.c_int32 * 3),
        ("raysPerProbe", C.c_uint32),
        ("colorRes", C.c_uint32),
        ("depthRes", C.c_uint32),
        ("shadowBias", C.c_float),
        ("padding", C.c_uint32),
    ]

    @staticmethod
    def make(extent_min, extent_max, resolution=(32, 16, 32), rays_per_probe=192, depth_sharpness=12.0, hysteresis=0.0, shadow_bias=0.3):
        g = GridInfo()
        g.extentMin[:] = [float(np.float32(x)) for x in extent_min]
        g.extentMax[:] = [float(np.float32(x)) for x in extent_max]
        g.depthSharpness = depth_sharpness
        g.hysteresis = hysteresis
        g.resolution[:] = list(resolution)
        g.raysPerProbe = rays_per_probe
        g.colorRes, g.depthRes = 8, 16
        g.shadowBias = shadow_bias
        g.padding = 0
        return g

    @property
    def probe_count(self):
        return self.resolution[0] * self.resolution[1] * self.resolution[2]

    def atlas_shapes(self):
        """(irradiance (H, W), depth (H, W)) in texels; reference src/IrradianceProbes.cpp:40-82."""
        rx, ry, rz = self.resolution
        return (8 * rz, 8 * rx * ry), (16 * rz, 16 * rx * ry)


class Light(C.Structure):
    _fields_ = [("direction", C.c_float * 4), ("color", C.c_float * 4)]

    @staticmethod
    def default():
        """LightBuffer defaults, reference src/Light.hpp:7-8."""
        l = Light()
        d = np.array([0.2, 2.0, 0.2], dtype=np.float32)
        d = d * np.float32(1.0) / np.sqrt(np.float32(d.dot(d)))
        l.direction[:] = [float(d[0]), float(d[1]), float(d[2]), 1.0]
        l.color[:] = [10.0, 10.0, 10.0, 10.0]
        return l


class Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("origin", C.c_float * 3), ("frameIndex", C.c_uint32)]


class Texture(C.Structure):
    """vkx_texture (include/vkx.h): one decoded RGBA8 image + VkFormat class + glTF sampler description."""

    _fields_ = [
        ("pixels", C.c_void_p),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("srgb", C.c_uint32),
        ("magFilter", C.c_uint32),
        ("minFilter", C.c_uint32),
        ("wrapS", C.c_uint32),
        ("wrapT", C.c_uint32),
    ]


def texture_array(textures):
    """list of dicts {pixels: uint8 [h, w, 4], srgb, magFilter, minFilter, wrapS, wrapT} -> (ctypes array, keep-alive list)."""
    arr = (Texture * max(1, len(textures)))()
    keep = []
    for i, t in enumerate(textures):
        px = np.ascontiguousarray(t["pixels"], dtype=np.uint8)
        assert px.ndim == 3 and px.shape[2] == 4, "texture pixels must be [height, width, 4] uint8"
        keep.append(px)
        arr[i].pixels = px.ctypes.data
        arr[i].height, arr[i].width = px.shape[0], px.shape[1]
        arr[i].srgb = int(t.get("srgb", 1))
        arr[i].magFilter = int(t.get("magFilter", 0))
        arr[i].minFilter = int(t.get("minFilter", 0))
        arr[i].wrapS = int(t.get("wrapS", 0))
        arr[i].wrapT = int(t.get("wrapT", 0))
    return arr, keep


def mip_level_count(width, height):
    """floor(log2(max(w, h))) + 1 (reference src/vulkan/Image.cpp:29-31)."""
    return int(max(width, height)).bit_length()


def mip_chain_texels(width, height):
    return sum(max(1, width >> l) * max(1, height >> l) for l in range(mip_level_count(width, height)))


class BvhInfo(C.Structure):
    _fields_ = [
        ("numNodes", C.c_uint32),
        ("numTriangles", C.c_uint32),
        ("numBinaryNodes", C.c_uint32),
        ("depth", C.c_uint32),
        ("sceneMin", C.c_float * 3),
        ("sceneMax", C.c_float * 3),
        ("sahCost", C.c_float),
        ("buildMs", C.c_float),
    ]


assert C.sizeof(GridInfo) == 64 and C.sizeof(Light) == 32 and C.sizeof(Camera) == 144


def look_at(eye, center, up):
    """glm::lookAt (right-handed), column-major flat 16 floats (reference src/Camera.cpp:67)."""
    eye, center, up = (np.asarray(v, dtype=np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.identity(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s.dot(eye), -u.dot(eye), f.dot(eye)
    return m.T.reshape(16).astype(np.float32)  # row-major math matrix -> column-major storage


def perspective(fovy_rad, aspect, near, far):
    """glm::perspective (RH, -1..1 depth) with proj[1][1] *= -1 as Editor::updateUniformBuffer does (src/Editor.cpp:512)."""
    t = np.tan(fovy_rad / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = -1.0 / t
    m[2, 2] = -(far + near) / (far - near)
    m[3, 2] = -1.0
    m[2, 3] = -(2.0 * far * near) / (far - near)
    return m.T.reshape(16).astype(np.float32)


def make_camera(eye, center, up=(0, 1, 0), fov_deg=60.0, aspect=16.0 / 9.0, near=0.1, far=4000.0, frame_index=0):
    """CameraBuffer as the reference fills it (fov 60, near 0.1, far 4000: src/Camera.hpp:79-82)."""
    c = Camera()
    c.view[:] = look_at(eye, center, up).tolist()
    c.proj[:] = perspective(np.radians(fov_deg), aspect, near, far).tolist()
    c.origin[:] = [float(x) for x in eye]
    c.frameIndex = frame_index
    return c
