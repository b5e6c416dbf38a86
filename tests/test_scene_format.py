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
