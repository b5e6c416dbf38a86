"""Skinned meshes: vkx_skin_vertices (CUDA) vs the oracle's vertexSkinning.comp restatement, and the rebuild that follows."""
import numpy as np
import pytest

from conftest import get_scene
from test_skinning_oracle import skinning_inputs
from vulkanexp_b200 import scene_format
from vulkanexp_b200._lib import Context, VkxError
from vulkanexp_b200.pods import GridInfo, Light

pytestmark = pytest.mark.gpu


def test_skinned_vertices_rebuild_and_update_match_oracle(oracle_lib):
    flat, src, dst, size = scene_format.add_skinned_instance(get_scene("court"), 2, transform_rows=[1, 0, 0, 0.5, 0, 1, 0, 2.0, 0, 0, 1, -0.5])
    o, g = oracle_lib.Oracle(), Context(0)
    for c in (o, g):
        c.scene_upload(flat); c.bvh_build()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (6, 5, 6), 64)
    g.probes_debug(True)
    o.probes_init(grid); g.probes_init(grid)
    host = oracle_lib.HostLogic()
    light = Light.default()
    rng = np.random.default_rng(21)
    n = 40000
    origins = rng.uniform(flat["bounds_min"], flat["bounds_max"], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    for pose in range(3):
        jt, sj, sw = skinning_inputs(size, joints=6, seed=30 + pose)
        mo = o.skin_vertices(jt, sj, sw, src, dst, motion=True)
        mg = g.skin_vertices(jt, sj, sw, src, dst, motion=True)
        assert mo.tobytes() == mg.tobytes(), "pose %d: motion vectors differ" % pose
        assert o.vertices_download(0, len(flat["vertices"])).tobytes() == g.vertices_download(0, len(flat["vertices"])).tobytes(), "pose %d: vertex arena differs" % pose
        with pytest.raises(VkxError):  # the structure is stale until rebuilt
            g.trace(origins[:4], d[:4], 0.01, 100.0)
        o.bvh_build(); g.bvh_build()
        no, to = o.bvh_download(); ng, tg = g.bvh_download()
        assert no.tobytes() == ng.tobytes() and to.tobytes() == tg.tobytes(), "pose %d: rebuilt BVH differs" % pose
        assert o.trace(origins, d, 0.01, 100.0).tobytes() == g.trace(origins, d, 0.01, 100.0).tobytes()
        R, _ = host.next_orientation()
        o.probes_update(grid, light, R, None); g.probes_update(grid, light, R, None)
        ho, so = o.probes_download_hits(); hg, sg = g.probes_download_hits()
        assert ho.tobytes() == hg.tobytes() and np.array_equal(so, sg)
        skinned_id = len(flat["instances"]) - 1
        assert not (hg["instance"] == skinned_id).any(), "probe rays must not see skinned instances (cull mask static | dynamic)"
        io, do, sto, _ = o.probes_download(); g.probes_upload(io, do, sto)
        grid.hysteresis = 0.5
    # shadow rays (mask 0xFF) do see the skinned ball
    hit = g.trace(origins, d, 0.01, 100.0)
    assert (hit["instance"] == skinned_id).any()


def test_skinned_refit_matches_oracle(oracle_lib):
    """Renderer::updateSkinnedBLAS updates the skinned BLASes in place (reference src/Renderer.cpp:644-669): after vkx_skin_vertices the
    topology-preserving vkx_bvh_refit gives the oracle's refitted structure byte for byte over three poses, and traced hits agree bit
    for bit with the oracle's and (up to grazing ties) with a rebuilt structure's."""
    flat, src, dst, size = scene_format.add_skinned_instance(get_scene("court"), 2, transform_rows=[1, 0, 0, 0.5, 0, 1, 0, 2.0, 0, 0, 1, -0.5])
    o, g, r = oracle_lib.Oracle(), Context(0), Context(0)
    for c in (o, g, r):
        c.scene_upload(flat); c.bvh_build()
    rng = np.random.default_rng(22)
    n = 40000
    origins = rng.uniform(flat["bounds_min"], flat["bounds_max"], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    for pose in range(3):
        jt, sj, sw = skinning_inputs(size, joints=6, seed=40 + pose)
        o.skin_vertices(jt, sj, sw, src, dst); g.skin_vertices(jt, sj, sw, src, dst); r.skin_vertices(jt, sj, sw, src, dst)
        o.bvh_refit(); g.bvh_refit(); r.bvh_build()
        no, to = o.bvh_download(); ng, tg = g.bvh_download()
        assert no.tobytes() == ng.tobytes() and to.tobytes() == tg.tobytes(), "pose %d: refitted BVH differs from the oracle" % pose
        hg = g.trace(origins, d, 0.01, 100.0)
        assert o.trace(origins, d, 0.01, 100.0).tobytes() == hg.tobytes()
        hr = r.trace(origins, d, 0.01, 100.0)
        same = (hg["t"] == hr["t"]) & (hg["instance"] == hr["instance"]) & (hg["primitive"] == hr["primitive"])
        assert same.mean() > 0.9999, same.mean()
    assert (hg["instance"] == len(flat["instances"]) - 1).any(), "rays with mask 0xFF see the skinned instance"


def test_skin_argument_checks():
    flat, src, dst, size = scene_format.add_skinned_instance(get_scene("court"), 2)
    g = Context(0); g.scene_upload(flat)
    jt, sj, sw = skinning_inputs(size, joints=4)
    with pytest.raises(VkxError):
        g.skin_vertices(jt, sj, sw, src, len(flat["vertices"]) - 3)  # destination range out of bounds
    with pytest.raises(VkxError):
        g.skin_vertices(jt, sj, sw, src, src + 1)  # overlapping ranges
    bad = sj.copy(); bad[5, 2] = 4
    with pytest.raises(VkxError):
        g.skin_vertices(jt, bad, sw, src, dst)  # joint index out of range
    g.skin_vertices(jt, sj, sw, src, dst)
