"""How evenly does a z-slab split of the cfg4 volume divide the work? Runs, on ONE GPU, the partial update of each of the N slabs a
sharded run would give to its N ranks (same kernels, same probes) and prints the device time of each: max / mean is the load
imbalance a strong-scaling run pays.  usage: python tools/slab_balance.py [N] [interleave_pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flat = scene_format.flatten(synth.make_cfg4())
g = Context(0); g.scene_upload(flat); g.bvh_build()
res = (64, 32, 64)
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, 256, hysteresis=0.5)
g.probes_init(grid); g.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
gen = OrientationGenerator(); gen.next(); light = Light.default()
for w in range(2):
    g.probes_update(grid, light, gen.next(), None)
plane = res[0] * res[1]
for mode in ("slabs", "interleaved pairs of z-slices"):
    times = []
    for r in range(N):
        if mode == "slabs":
            zs = range(r * res[2] // N, (r + 1) * res[2] // N)
        else:
            zs = [z for z in range(res[2]) if (z // 2) % N == r]
        idx = np.concatenate([np.arange(z * plane, (z + 1) * plane, dtype=np.uint32) for z in zs])
        R = gen.next()
        g.probes_update(grid, light, R, idx)  # warm
        g.probes_update(grid, light, R, idx)
        t = g.probes_timings(); k = g.probes_kernel_timings()
        times.append(t["full"])
        print(mode, "rank", r, "probes", len(idx), "full %.3f ms" % t["full"], {n: round(v, 3) for n, v in k.items() if n in ("trace_primary", "shade", "trace_shadow", "blend")}, flush=True)
    print("==", mode, "max %.3f mean %.3f imbalance %.3f" % (max(times), np.mean(times), max(times) / np.mean(times)), flush=True)
