"""Dev helper: time the cfg2 DDGI update on the GPU (not the bench contract; see bench.py)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.pods import GridInfo, Light
from oracle import pyoracle

t = time.time(); s = synth.make_cfg2(); flat = scene_format.flatten(s); print('scene', time.time() - t, 'tris', synth.count_triangles(s))
g = Context(0); g.scene_upload(flat); t = time.time(); g.bvh_build(); print('build wall', time.time() - t, 'ms dev', g.bvh_info().buildMs, 'nodes', g.bvh_info().numNodes, 'depth', g.bvh_info().depth)
grid = GridInfo.make(flat['bounds_min'], flat['bounds_max'], (32, 16, 32), 256)
g.probes_init(grid)
host = pyoracle.HostLogic()
R, _ = host.next_orientation()
t = time.time(); g.probes_classify(R); print('classify wall', time.time() - t)
st = g.probes_download()[2]; print('states', np.bincount(st, minlength=9))
g.probes_upload(state=np.ones_like(st))
light = Light.default()
for f in range(8):
    R, _ = host.next_orientation()
    grid.hysteresis = min(0.98, 0.2 * f)
    g.probes_update(grid, light, R, None)
    tm = g.probes_timings()
    print(f, tm, 'Grays/s', grid.probe_count * 256 / tm['full'] / 1e6)
