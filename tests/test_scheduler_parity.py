"""On-device selectProbesToUpdate (vkx_probes_schedule) vs the oracle's restatement of IrradianceProbes.cpp:396-424."""
import numpy as np
import pytest

from conftest import get_scene, make_pair
from vulkanexp_b200.pods import GridInfo, Light

pytestmark = pytest.mark.gpu


def _grid(flat, res, rays=16):
    return GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, rays)


@pytest.mark.parametrize("res,per_update", [((8, 6, 7), 0), ((8, 6, 7), 37), ((8, 6, 7), 336), ((8, 6, 7), 5000), ((5, 3, 3), 1), ((64, 32, 64), 0), ((64, 32, 64), 20000)])
def test_device_scheduler_equals_oracle(oracle_lib, res, per_update):
    from vulkanexp_b200._lib import Context
    flat = get_scene("tiny")
    g = Context(0)
    grid = _grid(flat, res)
    g.probes_init(grid)
    host = oracle_lib.HostLogic()  # counters start at (0, 0) on both sides
    rng = np.random.default_rng(res[0] * 1000 + per_update)
    P = grid.probe_count
    for rnd in range(10):
        if rnd == 3:
            state = np.zeros(P, dtype=np.uint32)            # nothing to update: the scan still advances by P
        elif rnd == 4:
            state = np.ones(P, dtype=np.uint32)             # everything, every frame
        else:
            state = rng.integers(0, 9, size=P, dtype=np.uint32)
        g.probes_upload(state=state)
        want = host.select(state, per_update)
        n = g.probes_schedule(per_update)
        got = g.probes_scheduled_list()
        assert n == len(want) == len(got), "round %d: count %d vs oracle %d" % (rnd, n, len(want))
        assert np.array_equal(got, want), "round %d: list differs" % rnd
    loop, off = g.probes_scheduler_state()
    assert off < P and loop >= 1


def test_scheduled_update_equals_host_list_update(oracle_lib):
    """schedule + update_scheduled on one context == state read-back + host selection + list update on another (bitwise)."""
    o, g1, flat = make_pair(oracle_lib, "court")
    from vulkanexp_b200._lib import Context
    g2 = Context(0); g2.scene_upload(flat); g2.bvh_build()
    grid = _grid(flat, (8, 6, 8), 64); grid.hysteresis = 0.7
    hostR = oracle_lib.HostLogic()
    sched = oracle_lib.HostLogic()
    light = Light.default()
    R, _ = hostR.next_orientation()
    for g in (g1, g2):
        g.probes_init(grid); g.probes_classify(R)
    for frame in range(8):
        R, _ = hostR.next_orientation()
        n = g1.probes_schedule(150)
        g1.probes_update_scheduled(grid, light, R)
        _, _, st2, _ = g2.probes_download()
        lst = sched.select(st2, 150)
        assert n == len(lst)
        if len(lst):
            g2.probes_update(grid, light, R, lst)
        a = g1.probes_download(); b = g2.probes_download()
        for x, y, name in zip(a[:3], b[:3], ("irradiance", "depth", "state")):
            assert np.array_equal(x, y), "frame %d: %s differs" % (frame, name)
    assert len(set(a[2].tolist())) > 1, "probe states should have diverged by now"


def test_update_scheduled_needs_a_schedule(oracle_lib):
    from vulkanexp_b200._lib import VkxError
    o, g, flat = make_pair(oracle_lib, "tiny")
    grid = _grid(flat, (4, 3, 5))
    g.probes_init(grid)
    R, _ = oracle_lib.HostLogic().next_orientation()
    with pytest.raises(VkxError):
        g.probes_update_scheduled(grid, Light.default(), R)
    g.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    assert g.probes_schedule(0) == grid.probe_count
    g.probes_update_scheduled(grid, Light.default(), R)
    with pytest.raises(VkxError):  # consumed
        g.probes_update_scheduled(grid, Light.default(), R)
