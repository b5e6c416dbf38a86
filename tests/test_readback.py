"""Asynchronous read-back (vkx_probes_download_async / _slab_async / _wait): a request returns the sampled atlases as they were when it
was made, although the copy itself is queued later - alongside the next update's traversal (csrc/api.cu::flushCopyRequests)."""
import numpy as np
import pytest

from conftest import make_pair
from vulkanexp_b200.pods import GridInfo, Light

pytestmark = pytest.mark.gpu


def _bufs(grid):
    (ih, iw), (dh, dw) = grid.atlas_shapes()
    return (np.full((ih, iw), 0xDEADBEEF, dtype=np.uint32), np.full((dh, dw), 0xDEADBEEF, dtype=np.uint32), np.full(grid.probe_count, 0xDEADBEEF, dtype=np.uint32))


def _same(got, want):
    return all(np.array_equal(a, b) for a, b in zip(got, want[:3]))


def test_requested_readback_returns_the_atlases_of_the_request(oracle_lib):
    o, g, flat = make_pair(oracle_lib, "court")
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (8, 6, 8), 64); grid.hysteresis = 0.7
    host = oracle_lib.HostLogic()
    light = Light.default()
    R, _ = host.next_orientation()
    g.probes_init(grid); g.probes_classify(R)
    outs = [_bufs(grid) for _ in range(3)]
    for f in range(3):  # request -> next update (the copy runs alongside it) -> request ...
        R, _ = host.next_orientation()
        g.probes_update(grid, light, R, None, sync=False)
        g.probes_download_async(outs[f])
    g.probes_download_wait()
    last = g.probes_download()
    assert _same(outs[2], last), "the last request does not hold the final atlases"
    assert not _same(outs[0], last) and not _same(outs[1], last), "three updates left identical atlases: the test cannot tell the frames apart"
    # replay on a second context with blocking read-backs: every request must equal the state right after its own update
    o2, g2, _ = make_pair(oracle_lib, "court")
    host2 = oracle_lib.HostLogic()
    R, _ = host2.next_orientation()
    g2.probes_init(grid); g2.probes_classify(R)
    for f in range(3):
        R, _ = host2.next_orientation()
        g2.probes_update(grid, light, R, None)
        assert _same(outs[f], g2.probes_download()), "request %d returned another frame's atlases" % f


def test_readback_is_ordered_before_an_upload_and_a_slab_request(oracle_lib):
    o, g, flat = make_pair(oracle_lib, "court")
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (8, 6, 8), 64)
    host = oracle_lib.HostLogic()
    R, _ = host.next_orientation()
    g.probes_init(grid); g.probes_classify(R)
    R, _ = host.next_orientation()
    g.probes_update(grid, Light.default(), R, None)
    before = g.probes_download()
    out = _bufs(grid)
    g.probes_download_async(out)                           # requested ...
    zeros = tuple(np.zeros_like(a) for a in before[:3])
    g.probes_upload(irr=zeros[0], dep=zeros[1], state=zeros[2])  # ... and then overwritten: the request still sees the old bytes
    g.probes_download_wait()
    assert _same(out, before)
    assert _same(g.probes_download(), zeros)
    # slab request: rows of z-slices [2, 5)
    g.probes_upload(irr=before[0], dep=before[1], state=before[2])
    (ih, iw), (dh, dw) = grid.atlas_shapes()
    plane = grid.resolution[0] * grid.resolution[1]
    slab = (np.zeros((8 * 3, iw), np.uint32), np.zeros((16 * 3, dw), np.uint32), np.zeros(3 * plane, np.uint32))
    g.probes_download_slab_async(2, 5, slab)
    g.probes_download_wait()
    assert np.array_equal(slab[0], before[0][16:40]) and np.array_equal(slab[1], before[1][32:80]) and np.array_equal(slab[2], before[2][2 * plane:5 * plane])
