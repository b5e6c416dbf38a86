import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
for name, mk in (("cfg2", synth.make_cfg2), ("cfg4", synth.make_cfg4)):
    flat = scene_format.flatten(mk())
    g = Context(0); g.scene_upload(flat)
    for i in range(3):
        t0 = time.perf_counter(); g.bvh_build(); dt = (time.perf_counter() - t0) * 1e3
        info = g.bvh_info()
        print(name, "build", i, "wall %.1f ms, event %.1f ms, %d tris, %d nodes, depth %d" % (dt, info.buildMs, info.numTriangles, info.numNodes, info.depth), flush=True)
    for i in range(3):  # topology-preserving refit (vkx_bvh_refit) of the same structure
        g.instances_update(flat["instances"])
        t0 = time.perf_counter(); g.bvh_refit(); dt = (time.perf_counter() - t0) * 1e3
        print(name, "refit", i, "wall %.2f ms, event %.2f ms" % (dt, g.bvh_info().buildMs), flush=True)
