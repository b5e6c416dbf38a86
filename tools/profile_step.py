"""Profiling driver (run under ncu on the GPU box): cfg2 scene, a few full-volume DDGI updates, no torch import."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
TEXTURED = os.environ.get("VKX_CFG2_TEXTURED", "0") == "1"  # the same atrium with the procedural texture set on every material
flat = scene_format.flatten(synth.make_cfg2(textured=TEXTURED))
g = Context(0); g.scene_upload(flat); g.bvh_build()
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (32, 16, 32), 256)
g.probes_init(grid)
g.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
gen = OrientationGenerator(); gen.next()
light = Light.default()
hist = []
for f in range(steps):
    grid.hysteresis = min(0.98, 0.25 * f)
    g.probes_update(grid, light, gen.next(), None)
    t, k = g.probes_timings(), g.probes_kernel_timings()
    print(f, t, k)
    hist.append({**k, "full": t["full"]})
if len(hist) > 2:  # last line: medians over the steps after the first two (tools/pick_defer.py reads it)
    print("median", "textured" if TEXTURED else "untextured", {name: float(np.median([h[name] for h in hist[2:]])) for name in hist[0]})
