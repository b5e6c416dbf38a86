"""CPU tests of the oracle's texture path ("sampler spec v1", oracle/texture.h): mip generation (Image::generateMipmaps), the
sampler state of Resources.cpp:88-124, textureGrad level selection, anyhit.rahit and texDerivative (closesthit.glsl:50-107).
The reference ships no fixtures for its fixed-function sampler, so these pin the restatement to the Vulkan equations evaluated
independently in numpy and to size-independent properties."""
import ctypes as C

import numpy as np
import pytest

from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200.pods import GridInfo, Light, VERTEX_DTYPE, mip_level_count

f32 = np.float32


def _srgb_decode(c):
    c = np.asarray(c, dtype=np.float64) / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def _bilinear_np(level, s, t, wrap=("repeat", "repeat"), decode=lambda x: x.astype(f32) / f32(255)):
    """Vulkan LINEAR filter in fp32, one look-up, on an [h, w, 4] uint8 level."""
    h, w = level.shape[:2]

    def wr(i, n, mode):
        if mode == "clamp":
            return min(max(i, 0), n - 1)
        if mode == "repeat":
            return i % n
        m = i % (2 * n) - n
        m = m if m >= 0 else -(1 + m)
        return (n - 1) - m

    u, v = f32(s) * f32(w) - f32(0.5), f32(t) * f32(h) - f32(0.5)
    i0, j0 = int(np.floor(u)), int(np.floor(v))
    a, b = f32(u - f32(i0)), f32(v - f32(j0))
    tex = lambda i, j: decode(level[wr(j, h, wrap[1]), wr(i, w, wrap[0])])
    lerp = lambda x, y, k: (x * (f32(1) - k) + y * k).astype(f32)
    return lerp(lerp(tex(i0, j0), tex(i0 + 1, j0), a), lerp(tex(i0, j0 + 1), tex(i0 + 1, j0 + 1), a), b)


@pytest.fixture()
def orc(oracle_lib):
    return oracle_lib.Oracle()


def test_mip_chain_shapes_and_box_filter(orc):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    odd = rng.integers(0, 256, (12, 20, 4), dtype=np.uint8)
    orc.scene_textures([{"pixels": img, "srgb": 0}, {"pixels": odd, "srgb": 0}])
    m = orc.texture_download(0)
    assert [l.shape[:2] for l in m] == [(8, 16), (4, 8), (2, 4), (1, 2), (1, 1)] and len(m) == mip_level_count(16, 8)
    assert m[0].tobytes() == img.tobytes()
    prev = img
    for l in m[1:4]:  # even sizes: a LINEAR blit to half size is the 2x2 box average (then round-to-nearest 8 bit)
        box = (prev.astype(np.float32) / 255).reshape(prev.shape[0] // 2, 2, prev.shape[1] // 2, 2, 4)
        top = (box[:, 0, :, 0] * 0.5 + box[:, 0, :, 1] * 0.5)
        bot = (box[:, 1, :, 0] * 0.5 + box[:, 1, :, 1] * 0.5)
        want = np.floor((top * 0.5 + bot * 0.5) * 255 + 0.5).astype(np.uint8)
        assert np.array_equal(l, want)
        prev = l
    assert np.array_equal(m[4][0, 0], np.floor((prev[0, 0].astype(np.float32) / 255 * 0.5 + prev[0, 1].astype(np.float32) / 255 * 0.5) * 255 + 0.5).astype(np.uint8))  # 1x2 -> 1x1: the height stays 1
    mo = orc.texture_download(1)
    assert [l.shape[:2] for l in mo] == [(12, 20), (6, 10), (3, 5), (1, 2), (1, 1)]
    # odd source size (3x5 -> 1x2): blit equations, scale 2.5 horizontally, 3 vertically (rows 1 only: v = 1.0 exactly)
    src = mo[2].astype(np.float32) / 255
    for i in range(2):
        u = (i + 0.5) * 2.5 - 0.5
        i0, a = int(np.floor(u)), np.float32(u - np.floor(u))
        row = src[1]
        want = np.floor((row[i0] * (1 - a) + row[min(i0 + 1, 4)] * a) * 255 + 0.5).astype(np.uint8)
        assert np.array_equal(mo[3][0, i], want)


def test_srgb_levels_are_filtered_in_linear_space_and_codes_round_trip(orc):
    codes = np.arange(256, dtype=np.uint8)
    img = np.zeros((2, 512, 4), dtype=np.uint8)  # column pairs of one code each: level 1 must return the code unchanged
    img[:, :, 0] = np.repeat(codes, 2)[None, :]
    img[:, :, 1] = 255 - np.repeat(codes, 2)[None, :]
    img[:, :, 2] = 7
    img[:, :, 3] = np.repeat(codes, 2)[None, :]
    orc.scene_textures([{"pixels": img, "srgb": 1}])
    m = orc.texture_download(0)
    assert np.array_equal(m[1][0, :, 0], codes) and np.array_equal(m[1][0, :, 1], 255 - codes) and np.array_equal(m[1][0, :, 3], codes)
    # level 2 averages neighbouring codes in linear light: the result is the code whose decode is nearest in sRGB space
    lin = _srgb_decode(codes)
    avg = 0.5 * lin[0::2] + 0.5 * lin[1::2]
    enc = np.where(avg <= 0.0031308, avg * 12.92, 1.055 * avg ** (1 / 2.4) - 0.055) * 255
    assert np.abs(m[2][0, :, 0].astype(np.float64) - enc).max() <= 0.5 + 1e-3
    # alpha is linear even in sRGB formats
    assert np.array_equal(m[2][0, :, 3], np.floor((codes[0::2].astype(np.float32) / 255 * 0.5 + codes[1::2].astype(np.float32) / 255 * 0.5) * 255 + 0.5).astype(np.uint8))
    # decode of a texel centre = the exact transfer function
    got = orc.texture_sample(0, np.array([[(2 * 200 + 0.5) / 512, 0.25]], dtype=np.float32))[0]
    assert abs(got[0] - _srgb_decode(200)) < 1e-7 and abs(got[3] - 200 / 255) < 1e-7


@pytest.mark.parametrize("wrap_s,wrap_t,names", [(10497, 10497, ("repeat", "repeat")), (33071, 33648, ("clamp", "mirror")), (33648, 33071, ("mirror", "clamp"))])
def test_linear_filter_and_wrap_modes_match_vulkan_equations(orc, wrap_s, wrap_t, names):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (6, 10, 4), dtype=np.uint8)
    orc.scene_textures([{"pixels": img, "srgb": 0, "wrapS": wrap_s, "wrapT": wrap_t}])
    uv = rng.uniform(-2.5, 3.5, (200, 2)).astype(np.float32)
    uv[:4] = [[0.0, 0.0], [1.0, 1.0], [-1.0, 2.0], [0.05, 0.95]]
    got = orc.texture_sample(0, uv)
    for k in range(len(uv)):
        want = _bilinear_np(img, uv[k, 0], uv[k, 1], names)
        assert np.array_equal(got[k], want), (k, uv[k], got[k], want)


def test_nearest_filter_and_mirror(orc):
    img = np.zeros((1, 4, 4), dtype=np.uint8)
    img[0, :, 0] = [10, 20, 30, 40]
    img[..., 3] = 255
    orc.scene_textures([{"pixels": img, "srgb": 0, "magFilter": 9728, "minFilter": 9728, "wrapS": 33648}])
    u = np.array([0.1, 0.3, 0.6, 0.9, 1.1, 1.3, 1.9, 2.1, -0.1, -0.3], dtype=np.float32)
    got = orc.texture_sample(0, np.stack([u, np.full_like(u, 0.5)], axis=1))[:, 0] * 255
    assert np.allclose(got, [10, 20, 30, 40, 40, 30, 10, 10, 10, 20], atol=1e-4)


def test_texture_grad_selects_levels_as_the_spec_says(orc):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (32, 32, 4), dtype=np.uint8)
    orc.scene_textures([{"pixels": img, "srgb": 0}, {"pixels": img, "srgb": 0, "minFilter": 9984}])
    mips = orc.texture_download(0)
    uv = rng.uniform(0, 1, (50, 2)).astype(np.float32)

    def grads(rho_x, rho_y):  # rho in texels per pixel along u (x derivative) and v (y derivative)
        return np.tile(np.array([[rho_x / 32, 0, 0, rho_y / 32]], dtype=np.float32), (len(uv), 1))

    base = orc.texture_sample(0, uv)
    assert np.array_equal(orc.texture_sample(0, uv, grads(1.0, 0.5)), base), "lambda = 0: magnification at the base level"
    assert np.array_equal(orc.texture_sample(0, uv, grads(0.0, 0.0)), base), "zero footprint: base level (decree T3)"
    assert np.array_equal(orc.texture_sample(0, uv, grads(np.nan, 1.0)), base), "NaN derivative is dropped by max(): base level (decree T3)"
    for lvl in (1, 2, 3):  # isotropic: the larger axis decides (decree T2)
        got = orc.texture_sample(0, uv, grads(2.0**lvl, 1.0))
        want = np.stack([_bilinear_np(mips[lvl], u, v) for u, v in uv])
        assert np.array_equal(got, want), lvl
    got = orc.texture_sample(0, uv, grads(1.0, 2.0**1.5))  # trilinear: delta = 0.5 between levels 1 and 2
    l1 = np.stack([_bilinear_np(mips[1], u, v) for u, v in uv])
    l2 = np.stack([_bilinear_np(mips[2], u, v) for u, v in uv])
    assert np.abs(got - (0.5 * l1 + 0.5 * l2)).max() < 2e-6
    got = orc.texture_sample(0, uv, grads(1e6, 1.0))  # beyond the chain: last level (1x1)
    assert np.allclose(got, mips[-1][0, 0].astype(np.float32) / 255, atol=1e-7)
    # minFilter 9984 -> (VK_FILTER_NEAREST, MIPMAP_MODE_NEAREST) (Resources.cpp:8-32): level = ceil(d + 0.5) - 1, nearest texel
    def nearest(level, u, v):
        h, w = level.shape[:2]
        return level[int(np.floor(f32(v) * f32(h))) % h, int(np.floor(f32(u) * f32(w))) % w].astype(f32) / f32(255)

    for rho, lvl in ((2.0**0.4, 0), (2.0**0.6, 1), (2.0**1.5, 1), (2.0**1.51, 2)):
        got = orc.texture_sample(1, uv, grads(rho, 0.0))
        want = np.stack([nearest(mips[lvl], u, v) for u, v in uv])
        assert np.array_equal(got, want), (rho, lvl)


def _quad_scene(alpha_img, base=(1.0, 1.0, 1.0)):
    """One 2x2 m quad in the plane y = 0 (normal +y), uv = (x + 1, z + 1) / 2, albedo texture 0."""
    v = np.zeros(4, dtype=VERTEX_DTYPE)
    v["pos"] = [(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)]
    v["normal"] = (0, 1, 0)
    v["tangent"] = (1, 0, 0, 1)
    v["color"] = 1
    v["texCoord"] = [(0, 0), (1, 0), (1, 1), (0, 1)]
    mat = scene_format.material_json("quad", base, 0.0, 1.0)
    mat["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
    s = scene_format.SceneFile(materials=[mat])
    s.entities.append(scene_format.Entity("Root"))
    s.meshes.append(scene_format.Mesh("Quad", 0, v, np.array([0, 2, 1, 0, 3, 2], dtype=np.uint32)))  # front face up
    s.entities.append(scene_format.Entity("Quad", mesh_renderer=(0, 0)))
    s.entities[0].children.append(1)
    s.textures = [{"source": "q.pam", "format": scene_format.VK_FORMAT_R8G8B8A8_SRGB, "sampler": {"magFilter": 9728, "minFilter": 9728}}]
    s.images = [alpha_img]
    return scene_format.flatten(s)


def test_any_hit_cut_out(orc):
    img = np.full((4, 4, 4), 255, dtype=np.uint8)
    img[1, 2, 3] = 0      # a hole at texel (x 2, y 1)
    img[3, 0, 3] = 2      # alpha 2/255 < 0.01: also a hole
    img[0, 0, 3] = 3      # alpha 3/255 > 0.01: opaque
    flat = _quad_scene(img)
    orc.scene_upload(flat)
    orc.bvh_build()
    centres = np.array([[(x + 0.5) / 4 * 2 - 1, 1.0, (y + 0.5) / 4 * 2 - 1] for y in range(4) for x in range(4)], dtype=np.float32)
    down = np.tile(np.array([[0, -1, 0]], dtype=np.float32), (16, 1))
    plain = orc.trace(centres, down, 0.01, 10.0)
    assert (plain["t"] > 0).all(), "without the any-hit shader (probe pipeline) every ray hits the quad"
    cut = orc.trace(centres, down, 0.01, 10.0, alpha_test=True)
    holes = (cut["t"] < 0).reshape(4, 4)
    want = np.zeros((4, 4), dtype=bool)
    want[1, 2] = want[3, 0] = True
    assert np.array_equal(holes, want)
    occl = orc.trace(centres, down, 0.01, 10.0, any_hit=True, alpha_test=True)
    assert np.array_equal((occl["t"] < 0).reshape(4, 4), want)


def test_tex_derivative_against_the_differential_rays(oracle_lib):
    """The quad above, seen from (0.2, 3, -0.1) along a tilted ray; the footprint follows from intersecting the two differential
    rays with the plane. dudx / dudy are the geometric derivatives; dvdx / dvdy carry the reference's sign slip (closesthit.glsl:98,101)."""
    lib = oracle_lib.lib()
    v = np.zeros(3, dtype=VERTEX_DTYPE)
    v["pos"] = [(-1, 0, -1), (1, 0, 1), (1, 0, -1)]
    v["texCoord"] = [(0, 0), (1, 1), (1, 0)]
    origin = np.array([0.2, 3.0, -0.1], dtype=np.float32)
    d = np.array([0.1, -1.0, 0.2], dtype=np.float32)
    d /= np.linalg.norm(d)
    t = -origin[1] / d[1]
    pos = (origin + d * t).astype(np.float32)

    def rot(p, axis, ang):
        axis = axis / np.linalg.norm(axis)
        return (np.dot(axis, p) * axis * (1 - np.cos(ang)) + p * np.cos(ang) + np.cross(axis, p) * np.sin(ang)).astype(np.float32)

    rdx, rdy = rot(d, np.cross(d, [1.0, 0, 0]), 0.001), rot(d, np.cross(d, [0, 1.0, 0]), 0.001)
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.orc_tex_derivative(p(pos), p(origin), p(ident), p(v), p(rdx), p(rdy), p(out))
    fx = origin + rdx * (-origin[1] / rdx[1]) - pos  # footprint vectors in the plane
    fy = origin + rdy * (-origin[1] / rdy[1]) - pos
    # uv = ((x + 1) / 2, (z + 1) / 2): du = dx / 2, dv = dz / 2
    assert abs(out[0] - fx[0] / 2) < 1e-6 and abs(out[2] - fy[0] / 2) < 1e-6
    # :98 computes (-a10 * dpdx[dim0] - a00 * dpdx[dim1]) / det, i.e. minus the geometric dv/dx here (a10 = 0)
    assert abs(out[1] + fx[2] / 2) < 1e-6 and abs(out[3] + fy[2] / 2) < 1e-6


def test_white_textures_equal_untextured_materials_and_textures_matter(oracle_lib):
    """Multiplying by a 1x1 white texel is exact, so a scene whose texture slots all point to it shades bit-identically to the
    untextured scene; the procedural textures change the result."""
    light = Light.default()
    host = oracle_lib.HostLogic()
    R, _ = host.next_orientation()

    def run(make):
        s = make()
        flat = scene_format.flatten(s)
        o = oracle_lib.Oracle()
        o.scene_upload(flat)
        o.bvh_build()
        grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (4, 3, 4), 32)
        o.probes_init(grid)
        o.probes_update(grid, light, R, None)
        return o.probes_download(rays=True)

    def white():
        s = synth.make_open_court()
        for m in s.meshes:
            synth.planar_uvs(m)
        s.images = [np.full((1, 1, 4), 255, dtype=np.uint8)]
        s.textures = [{"source": "w.pam", "format": scene_format.VK_FORMAT_R8G8B8A8_SRGB, "sampler": {}}]
        for m in s.materials:
            m["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
            m["pbrMetallicRoughness"]["metallicRoughnessTexture"] = {"index": 0}
            m["emissiveTexture"] = {"index": 0}
        return s

    plain = run(synth.make_open_court)
    wh = run(white)
    for a, b in zip(plain, wh):
        assert a.tobytes() == b.tobytes()
    tex = run(synth.make_textured_court)
    assert np.array_equal(plain[3][..., 3], tex[3][..., 3]), "the probe pipeline has no any-hit shader: ray depths do not change"
    assert np.abs(plain[3][..., :3] - tex[3][..., :3]).max() > 1e-2


def test_scene_file_round_trip_keeps_textures(tmp_path):
    s = synth.make_textured_court()
    path = tmp_path / "t.scene"
    scene_format.write_scene(str(path), s)
    r = scene_format.read_scene(str(path))
    assert len(r.images) == 5 and all(a.tobytes() == b.tobytes() for a, b in zip(r.images, s.images))
    fa, fb = scene_format.flatten(s), scene_format.flatten(r)
    assert fa["materials"].tobytes() == fb["materials"].tobytes()
    assert [t["srgb"] for t in fb["textures"]] == [1, 0, 0, 1, 1]
    assert [t["minFilter"] for t in fb["textures"]] == [0, 9987, 9984, 9986, 9729]


def test_image_decoder_of_the_library(tmp_path):
    """vkx_image_decode (host only): PNG in every 8-bit colour type PIL writes, P6 and P7, against Pillow's own RGBA conversion."""
    Image = pytest.importorskip("PIL.Image")
    from vulkanexp_b200._lib import VkxError, image_decode

    rng = np.random.default_rng(5)
    rgba = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    for mode in ("RGBA", "RGB", "L", "LA", "P"):
        im = Image.fromarray(rgba, "RGBA").convert(mode) if mode != "P" else Image.fromarray(rgba[..., :3].copy(), "RGB").quantize(64)
        path = str(tmp_path / ("img_%s.png" % mode))
        im.save(path)
        want = np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8)
        got = image_decode(path)
        assert got.shape == want.shape and np.array_equal(got, want), mode
    big = rng.integers(0, 256, (300, 500, 4), dtype=np.uint8)  # several IDAT chunks, every filter type
    path = str(tmp_path / "big.png")
    Image.fromarray(big, "RGBA").save(path, optimize=True)
    assert np.array_equal(image_decode(path), big)
    scene_format.write_pam(str(tmp_path / "a.pam"), rgba)
    assert np.array_equal(image_decode(str(tmp_path / "a.pam")), rgba)
    with open(tmp_path / "b.ppm", "wb") as f:
        f.write(b"P6\n# comment\n53 37\n255\n" + rgba[..., :3].tobytes())
    got = image_decode(str(tmp_path / "b.ppm"))
    assert np.array_equal(got[..., :3], rgba[..., :3]) and (got[..., 3] == 255).all()
    with open(tmp_path / "c.jpg", "wb") as f:
        f.write(b"\xff\xd8\xff\xe0" + bytes(64))
    with pytest.raises(VkxError):
        image_decode(str(tmp_path / "c.jpg"))


def test_image_decoder_matches_the_reference_stb_image():
    """tests/golden/stb_pin.json holds the SHA-256 of stbi_load(path, .., 4) — the call of src/STBImage.hpp:25, from the reference's
    vendored ext/stb_image.h — for the files under tests/golden/img (tools/gen_golden.py): the library's decoders must return the
    same bytes."""
    import hashlib
    import json
    import os

    from vulkanexp_b200._lib import image_decode

    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(gold_dir, "stb_pin.json")))
    assert len(gold) >= 8
    for name, want in gold.items():
        px = image_decode(os.path.join(gold_dir, "img", name))
        assert px.shape == (want["height"], want["width"], 4), name
        assert hashlib.sha256(px.tobytes()).hexdigest() == want["sha256"], name
