"""GPU diagnostic: runs the cfg2 (or another scene's) update against the oracle and prints the rays with the largest radiance
error together with what they hit, so a tolerance failure can be traced to a term. Usage: diag_parity.py [scene] [rx ry rz rays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.pods import GridInfo, Light

scene = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
res = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (32, 16, 32)
rays = int(sys.argv[5]) if len(sys.argv) > 5 else 256
flat = scene_format.flatten({"cfg2": synth.make_cfg2, "cfg4": synth.make_cfg4, "court": synth.make_open_court, "cfg1": synth.make_cfg1}[scene]())
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, rays)
o = pyoracle.Oracle(); o.scene_upload(flat); o.bvh_build(); o.probes_init(grid)
g = Context(0); g.scene_upload(flat); g.bvh_build(); g.probes_debug(True); g.probes_init(grid)
ones = np.ones(grid.probe_count, dtype=np.uint32)
o.probes_upload(state=ones); g.probes_upload(state=ones)
host = pyoracle.HostLogic(); host.next_orientation()
light = Light.default()
idx = None
if scene == "cfg4":  # one z-slab, as tests/test_ddgi_parity.py::test_cfg4_slab_update_parity
    plane = res[0] * res[1]; idx = np.arange(30 * plane, 32 * plane, dtype=np.uint32)
for frame in range(3):
    R, _ = host.next_orientation()
    grid.hysteresis = 0.0 if frame == 0 else (0.8 if scene == "cfg4" else 0.7)
    o.probes_update(grid, light, R, idx); g.probes_update(grid, light, R, idx)
    ho, so = o.probes_download_hits(); hg, sg = g.probes_download_hits()
    io, do, sto, ro = o.probes_download(rays=True); ig, dg, stg, rg = g.probes_download(rays=True)
    a, b = ro[..., :3].astype(np.float64), rg[..., :3].astype(np.float64)
    e = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-3)
    worst = np.argsort(e.max(axis=-1).ravel())[::-1][:8]
    print("frame %d: hits equal %s, shadow equal %s, max rel err %.3e, rays > 1e-3: %d, > 3e-4: %d of %d" % (frame, ho.tobytes() == hg.tobytes(), np.array_equal(so, sg), e.max(), int((e.max(axis=-1) > 1e-3).sum()), int((e.max(axis=-1) > 3e-4).sum()), e.shape[0] * e.shape[1]))
    for w in worst:
        p, r = divmod(int(w), rays)
        h = ho.reshape(-1)[w]
        print("   probe %d ray %d: err %.3e oracle %s gpu %s depth %.4f shadow %d hit t=%.4f inst=%d prim=%d" % (p, r, e.reshape(-1, 3)[w].max(), ro.reshape(-1, 4)[w, :3], rg.reshape(-1, 4)[w, :3], ro.reshape(-1, 4)[w, 3], so.reshape(-1)[w], h["t"], h["instance"], h["primitive"] & 0x7FFFFFFF))
    uio, udo = o.probes_download_unpacked(); uig, udg = g.probes_download_unpacked()
    rel = lambda x, y: float((np.abs(x.astype(np.float64) - y) / np.maximum(np.maximum(np.abs(x), np.abs(y)), 1e-3)).max())
    print("   texels fp32: irradiance %.3e depth %.3e; packed flips irr %.2e dep %.2e state %.2e" % (rel(uio, uig), rel(udo, udg), float((io != ig).mean()), float((do != dg).mean()), float((sto != stg).mean())))
    g.probes_upload(io, do, sto)
