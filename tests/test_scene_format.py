"""CPU-only: the binary .scene container (reference README.md:75-90, src/Scene.cpp:710-934) and host-side flattening."""
import os
import struct

import numpy as np

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
