"""GPU-box check: the textured sponza-scale atrium (synth.make_cfg2(textured=True)) through the CUDA path and the CPU oracle,
8x4x8 probes x 64 rays, two frames with resynchronisation. Prints one line; exits non-zero on a parity failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.pods import GridInfo, Light

flat = scene_format.flatten(synth.make_cfg2(textured=True))
o, g = pyoracle.Oracle(), Context(0)
for c in (o, g):
    c.scene_upload(flat); c.bvh_build()
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (8, 4, 8), 64)
g.probes_debug(True)
o.probes_init(grid); g.probes_init(grid)
host = pyoracle.HostLogic(); light = Light.default()
worst = 0.0
for frame in range(2):
    R, _ = host.next_orientation()
    o.probes_update(grid, light, R, None); g.probes_update(grid, light, R, None)
    ho, so = o.probes_download_hits(); hg, sg = g.probes_download_hits()
    assert ho.tobytes() == hg.tobytes() and np.array_equal(so, sg), "hit records differ"
    ro, rg = o.probes_download(rays=True)[3], g.probes_download(rays=True)[3]
    e = np.abs(ro[..., :3].astype(np.float64) - rg[..., :3]) / np.maximum(np.maximum(np.abs(ro[..., :3]), np.abs(rg[..., :3])), 1e-3)
    worst = max(worst, float(e.max()))
    uo, do = o.probes_download_unpacked(); ug, dg = g.probes_download_unpacked()
    for a, b in ((uo, ug), (do, dg)):
        worst = max(worst, float((np.abs(a.astype(np.float64) - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-3)).max()))
    io, dpo, sto, _ = o.probes_download(); g.probes_upload(io, dpo, sto)
    grid.hysteresis = 0.5
assert worst < 1e-3, "rel err %g" % worst
print("textured cfg2 parity ok: %d triangles, %d textures, hits bit-exact, max rel err %.2e" % (g.bvh_info().numTriangles, len(flat["textures"]), worst))
