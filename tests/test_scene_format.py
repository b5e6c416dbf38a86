"""CPU-only: the binary .scene container (reference README.md:75-90, src/Scene.cpp:710-934) and host-side flattening."""
import os
import struct

import numpy as np
import pytest

from vulkanexp_b200 import scene_format, synth


def test_scene_round_trip(tmp_path):
    s = synth.make_open_court()
    path = os.path.join(tmp_path, "court.scene")
    scene_format.write_scene(path, s)
    raw = open(path, "rb").read()
    magic, version, length = struct.unpack_from("<III", raw, 0)
    assert magic == 0x4E454353 and version == 0 and length == len(raw)
    jlen, jtype = struct.unpack_from("<II", raw, 12)
    assert jtype == 0x4E4F534A
    clen, ctype = struct.unpack_from("<II", raw, 20 + jlen)
    assert ctype == 0x004E4942 and clen == s.meshes[0].vertices.nbytes
    s2 = scene_format.read_scene(path)
    assert len(s2.meshes) == len(s.meshes) and len(s2.entities) == len(s.entities)
    for a, b in zip(s.meshes, s2.meshes):
        assert a.vertices.tobytes() == b.vertices.tobytes() and a.indices.tobytes() == b.indices.tobytes() and a.material == b.material
    fa, fb = scene_format.flatten(s), scene_format.flatten(s2)
    for k in fa:
        assert np.asarray(fa[k]).tobytes() == np.asarray(fb[k]).tobytes(), k


def test_flatten_orders_instances_like_sort_renderers():
    s = synth.make_open_court()
    f = scene_format.flatten(s)
    keys = [(int(f["offsets"][e]["materialIndex"]), int(e)) for e in f["instances"]["meshEntry"]]
    assert keys == sorted(keys)
    assert len(f["instances"]) == sum(1 for e in s.entities if e.mesh_renderer is not None)
    assert (f["instances"]["mask"] == 1).all()
    # offset table packs tightly in mesh order
    vo = io = 0
    for mi, m in enumerate(s.meshes):
        assert tuple(f["offsets"][mi]) == (m.material, vo, io)
        vo += len(m.vertices); io += len(m.indices)
    assert np.all(f["bounds_min"] <= f["bounds_max"])


def test_generators_are_deterministic_and_sized():
    a, b = synth.make_cfg1(), synth.make_cfg1()
    assert scene_format.flatten(a)["vertices"].tobytes() == scene_format.flatten(b)["vertices"].tobytes()
    assert 4000 < synth.count_triangles(a) < 6000
    assert abs(synth.count_triangles(synth.make_cfg2()) - 262_144) < 0.03 * 262_144


@pytest.mark.parametrize("maker", ["court", "tcourt", "cfg1"])
def test_facade_loader_equals_python_harness(tmp_path, maker):
    """The product's own .scene loader + flattening (csrc/host: Scene::loadScene, Scene::update, Renderer::allocateMeshes / createTLAS,
    the image decoders), reached through vkx_host_scene_* without a GPU, must produce the arrays the Python harness produces."""
    from vulkanexp_b200._lib import host_scene_load, host_scene_resave

    s = {"court": synth.make_open_court, "tcourt": synth.make_textured_court, "cfg1": synth.make_cfg1}[maker]()
    if maker == "tcourt":
        pytest.importorskip("PIL")
        s.textures[1]["source"] = "tex_normal.png"  # one image through the PNG decoder
    path = str(tmp_path / "a.scene")
    scene_format.write_scene(path, s)
    want = scene_format.flatten(scene_format.read_scene(path))
    got = host_scene_load(path)
    for key in ("vertices", "indices", "offsets", "mesh_index_counts", "materials", "instances"):
        assert got[key].tobytes() == want[key].tobytes(), key
    assert np.array_equal(got["bounds_min"], want["bounds_min"]) and np.array_equal(got["bounds_max"], want["bounds_max"])
    assert ("textures" in got) == ("textures" in want)
    for a, b in zip(got.get("textures", []), want.get("textures", [])):
        assert np.array_equal(a["pixels"], b["pixels"])
        assert all(int(a[k]) == int(b[k]) for k in ("srgb", "magFilter", "minFilter", "wrapS", "wrapT"))
    # Scene::save -> the Python reader: same scene again (texture files are referenced, not rewritten)
    out = str(tmp_path / "b.scene")
    host_scene_resave(path, out)
    again = scene_format.flatten(scene_format.read_scene(out))
    for key in ("vertices", "indices", "offsets", "mesh_index_counts", "materials", "instances"):
        assert again[key].tobytes() == want[key].tobytes(), "after Scene::save: " + key
    assert [t["minFilter"] for t in again.get("textures", [])] == [t["minFilter"] for t in want.get("textures", [])]


def test_facade_loader_reports_missing_files(tmp_path):
    from vulkanexp_b200._lib import VkxError, host_scene_load

    with pytest.raises(VkxError):
        host_scene_load(str(tmp_path / "nope.scene"))
    s = synth.make_textured_court()
    path = str(tmp_path / "a.scene")
    scene_format.write_scene(path, s)
    os.remove(str(tmp_path / "tex_grate.pam"))  # an unreadable texture becomes the blank image (reference src/Scene.cpp:699), with a warning
    got = host_scene_load(path)
    assert got["textures"][4]["pixels"].shape == (1, 1, 4) and (got["textures"][4]["pixels"] == 255).all()


def _write_raw_scene(path, entities, json_len_lie=None):
    import json

    doc = json.dumps({"materials": [], "entities": entities, "meshes": [], "textures": []}).encode()
    jlen = len(doc) if json_len_lie is None else json_len_lie
    total = 12 + 8 + len(doc)
    open(path, "wb").write(struct.pack("<III", 0x4E454353, 0, total) + struct.pack("<II", jlen, 0x4E4F534A) + doc)


def test_loader_rejects_truncated_and_cyclic_files(tmp_path):
    """Corrupt input must fail with an error code, not read past the buffer or recurse without end."""
    from vulkanexp_b200._lib import VkxError, host_scene_load

    ident = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]
    ok = [{"name": "root", "transform": ident, "parent": -1, "children": [1]}, {"name": "a", "transform": ident, "parent": -1, "children": []}]
    p = os.path.join(tmp_path, "ok.scene"); _write_raw_scene(p, ok); host_scene_load(p)
    p = os.path.join(tmp_path, "lie.scene"); _write_raw_scene(p, ok, json_len_lie=1 << 20)  # JSON chunk claims 1 MiB
    with pytest.raises(VkxError):
        host_scene_load(p)
    for name, ents in (("self", [{"name": "root", "transform": ident, "parent": -1, "children": [1]}, {"name": "a", "transform": ident, "parent": -1, "children": [1]}]),
                       ("cycle", [{"name": "root", "transform": ident, "parent": -1, "children": [1]}, {"name": "a", "transform": ident, "parent": -1, "children": [2]}, {"name": "b", "transform": ident, "parent": -1, "children": [1]}]),
                       ("range", [{"name": "root", "transform": ident, "parent": -1, "children": [7]}])):
        p = os.path.join(tmp_path, name + ".scene"); _write_raw_scene(p, ents)
        with pytest.raises(VkxError):
            host_scene_load(p)


def test_png_decoder_refuses_huge_headers(tmp_path):
    import zlib

    from vulkanexp_b200._lib import VkxError, image_decode

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    p = os.path.join(tmp_path, "huge.png")
    open(p, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 0xFFFFFFF0, 0xFFFFFFF0, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 16)) + chunk(b"IEND", b""))
    with pytest.raises(VkxError):
        image_decode(p)
    p = os.path.join(tmp_path, "thin.png")  # plausible size, but 8 bytes of IDAT cannot hold 4096 x 4096 RGBA
    open(p, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 4096, 4096, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 8)) + chunk(b"IEND", b""))
    with pytest.raises(VkxError):
        image_decode(p)
