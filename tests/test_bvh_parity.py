"""GPU builder / traversal vs the oracle: BVH topology and hit records must be bit-exact (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import get_scene, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene", ["tiny", "court", "cfg1", "cfg2"])
def test_bvh_topology_bit_exact(oracle_lib, scene):
    o, g, flat = make_pair(oracle_lib, scene)
    oi, gi = o.bvh_info(), g.bvh_info()
    assert (oi.numNodes, oi.numTriangles, oi.numBinaryNodes, oi.depth) == (gi.numNodes, gi.numTriangles, gi.numBinaryNodes, gi.depth)
    assert list(oi.sceneMin) == list(gi.sceneMin) and list(oi.sceneMax) == list(gi.sceneMax)
    on, ot = o.bvh_download()
    gn, gt = g.bvh_download()
    assert on.tobytes() == gn.tobytes(), "wide nodes differ"
    assert ot.tobytes() == gt.tobytes(), "triangle order / data differs"


def _random_rays(flat, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = flat["bounds_min"], flat["bounds_max"]
    org = rng.uniform(lo - 0.5, hi + 0.5, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # axis-aligned and zero-component directions exercise the zero-fix path
    d[: n // 50] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=n // 50)] * rng.choice([-1.0, 1.0], size=(n // 50, 1)).astype(np.float32)
    return org, d


@pytest.mark.parametrize("scene", ["court", "cfg1", "cfg2"])
def test_closest_hit_bit_exact(oracle_lib, scene):
    o, g, flat = make_pair(oracle_lib, scene)
    org, d = _random_rays(flat, 200_000, 7)
    ho = o.trace(org, d, 0.01, 1000.0, 0x3)
    hg = g.trace(org, d, 0.01, 1000.0, 0x3)
    assert (ho["t"] > 0).mean() > 0.3
    assert ho.tobytes() == hg.tobytes()


@pytest.mark.parametrize("scene", ["court", "cfg2"])
def test_any_hit_bit_exact(oracle_lib, scene):
    o, g, flat = make_pair(oracle_lib, scene)
    org, d = _random_rays(flat, 200_000, 11)
    ho = o.trace(org, d, 0.1, 10000.0, 0xFF, any_hit=True)
    hg = g.trace(org, d, 0.1, 10000.0, 0xFF, any_hit=True)
    assert np.array_equal(ho["t"], hg["t"])


def test_cull_mask_and_empty_scene(oracle_lib):
    from vulkanexp_b200._lib import Context
    from vulkanexp_b200.pods import INSTANCE_DTYPE, MATERIAL_DTYPE, OFFSET_DTYPE, VERTEX_DTYPE

    flat = dict(get_scene("tiny"))
    inst = flat["instances"].copy()
    inst["mask"][::2] = 4  # skinned: invisible to probe rays (mask 0x3), visible to shadow rays (0xFF)
    flat["instances"] = inst
    o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build()
    g = Context(0); g.scene_upload(flat); g.bvh_build()
    org, d = _random_rays(flat, 20_000, 3)
    for mask in (0x3, 0xFF, 0x4):
        assert o.trace(org, d, 0.01, 100.0, mask).tobytes() == g.trace(org, d, 0.01, 100.0, mask).tobytes()
    # empty scene: every ray misses
    empty = {"vertices": np.zeros(0, VERTEX_DTYPE), "indices": np.zeros(0, np.uint32), "offsets": np.zeros(0, OFFSET_DTYPE),
             "mesh_index_counts": np.zeros(0, np.uint32), "materials": np.zeros(0, MATERIAL_DTYPE), "instances": np.zeros(0, INSTANCE_DTYPE)}
    g2 = Context(0); g2.scene_upload(empty); g2.bvh_build()
    assert g2.bvh_info().numNodes == 1
    assert (g2.trace(org[:100], d[:100], 0.01, 100.0)["t"] < 0).all()


def test_instances_update_equals_fresh_upload(oracle_lib):
    """Dynamic instances (Renderer::updateAccelerationStructureInstances + updateTLAS): new transforms + rebuild must give exactly the
    structure a fresh upload with those transforms gives - on the GPU and in the oracle - and the next probe update must agree."""
    import copy
    from conftest import get_scene
    from vulkanexp_b200._lib import Context, VkxError
    from vulkanexp_b200.pods import GridInfo, Light

    flat = get_scene("cfg1")
    moved = dict(flat)
    inst = flat["instances"].copy()
    assert len(inst) > 3
    rng = np.random.default_rng(3)
    for k in range(1, len(inst)):  # translate + rotate about y every instance but the room
        ang = float(rng.uniform(0, 2 * np.pi)); c, s = np.float32(np.cos(ang)), np.float32(np.sin(ang))
        M = inst[k]["transform"].reshape(3, 4).copy()  # row-major 3x4 (VkTransformMatrixKHR)
        R = np.eye(3, dtype=np.float32); R[0, 0] = c; R[0, 2] = s; R[2, 0] = -s; R[2, 2] = c
        M = (R @ M).astype(np.float32)
        M[0, 3] += np.float32(rng.uniform(-0.8, 0.8)); M[1, 3] += np.float32(rng.uniform(0.0, 0.5)); M[2, 3] += np.float32(rng.uniform(-0.8, 0.8))
        inst[k]["transform"] = M.reshape(-1)
    moved["instances"] = inst
    a = Context(0); a.scene_upload(flat); a.bvh_build()
    before = a.bvh_download()[0].tobytes()
    a.instances_update(inst)
    with pytest.raises(VkxError):
        a.trace(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32), 0.0, 10.0)  # stale structure: must be rebuilt first
    a.bvh_build()
    b = Context(0); b.scene_upload(moved); b.bvh_build()
    o = oracle_lib.Oracle(); o.scene_upload(moved); o.bvh_build()
    na, ta = a.bvh_download(); nb, tb = b.bvh_download(); no, to = o.bvh_download()
    assert na.tobytes() != before, "the move must change the structure"
    assert na.tobytes() == nb.tobytes() == no.tobytes() and ta.tobytes() == tb.tobytes() == to.tobytes()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (6, 6, 6), 32)
    R, _ = oracle_lib.HostLogic().next_orientation()
    outs = []
    for g in (a, b):
        g.probes_init(grid); g.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32)); g.probes_update(grid, Light.default(), R)
        outs.append(g.probes_download())
    assert all(np.array_equal(x, y) for x, y in zip(outs[0][:3], outs[1][:3]))
    bad = inst.copy(); bad[1]["meshEntry"] = bad[0]["meshEntry"]
    if bad[1]["meshEntry"] != inst[1]["meshEntry"]:
        with pytest.raises(VkxError):
            a.instances_update(bad)


def test_refit_equals_oracle_refit(oracle_lib):
    """Topology-preserving refit (Renderer::updateAccelerationStructureInstances + updateTLAS, reference src/Renderer.cpp:671-742):
    vkx_bvh_refit after vkx_instances_update gives the oracle's refitted structure byte for byte (nodes and triangles), unchanged
    transforms reproduce the built structure, traced hits equal the oracle's on the refitted tree bit for bit and a rebuilt tree's
    up to grazing ties, and the next probe update on the refitted structure agrees with the oracle's."""
    from conftest import get_scene
    from test_oracle_kat import _moved_instances
    from vulkanexp_b200._lib import Context, VkxError
    from vulkanexp_b200.pods import GridInfo, Light

    for name in ("cfg1", "court"):
        flat = get_scene(name)
        g = Context(0); g.scene_upload(flat)
        with pytest.raises(VkxError):
            g.bvh_refit()  # nothing built yet
        g.bvh_build()
        o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build()
        n0, t0 = g.bvh_download()
        g.bvh_refit()
        n1, t1 = g.bvh_download()
        assert n0.tobytes() == n1.tobytes() and t0.tobytes() == t1.tobytes(), "refit with unchanged transforms must be the identity"
        if len(flat["instances"]) < 2:
            continue
        for seed in (3, 4):  # two successive moves: refit of a refitted tree
            inst = _moved_instances(flat, seed)
            g.instances_update(inst); g.bvh_refit()
            o.instances_update(inst); o.bvh_refit()
            ng, tg = g.bvh_download(); no, to = o.bvh_download()
            assert ng.tobytes() == no.tobytes(), "%s seed %d: refitted nodes differ from the oracle" % (name, seed)
            assert tg.tobytes() == to.tobytes(), "%s seed %d: refitted triangles differ from the oracle" % (name, seed)
            assert ng.tobytes() != n0.tobytes()
            io, ig = o.bvh_info(), g.bvh_info()
            assert np.array_equal(np.array(io.sceneMin[:]), np.array(ig.sceneMin[:])) and np.array_equal(np.array(io.sceneMax[:]), np.array(ig.sceneMax[:]))
        rng = np.random.default_rng(11)
        lo, hi = np.array(flat["bounds_min"]), np.array(flat["bounds_max"])
        org = rng.uniform(lo, hi, size=(50000, 3)).astype(np.float32)
        d = rng.normal(size=(50000, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        hg = g.trace(org, d, 0.001, 1000.0); ho = o.trace(org, d, 0.001, 1000.0)
        assert hg.tobytes() == ho.tobytes(), "hits on the refitted structure differ from the oracle"
        moved = dict(flat); moved["instances"] = inst
        r = Context(0); r.scene_upload(moved); r.bvh_build()
        hr = r.trace(org, d, 0.001, 1000.0)
        same = (hg["t"] == hr["t"]) & (hg["instance"] == hr["instance"]) & (hg["primitive"] == hr["primitive"])
        assert same.mean() > 0.9999, same.mean()
        grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (6, 6, 6), 32)
        R, _ = oracle_lib.HostLogic().next_orientation()
        g.probes_debug(True); g.probes_init(grid); o.probes_init(grid)
        ones = np.ones(grid.probe_count, dtype=np.uint32)
        g.probes_upload(state=ones); o.probes_upload(state=ones)
        g.probes_update(grid, Light.default(), R); o.probes_update(grid, Light.default(), R, None)
        hg2, _ = g.probes_download_hits(); ho2, _ = o.probes_download_hits()
        assert hg2.tobytes() == ho2.tobytes(), "probe rays on the refitted structure differ from the oracle"
