"""Runs the BASELINE.json configs on one GPU (synthetic scenes of the stated scale) and prints one JSON line per config:
BVH build time, full-volume DDGI update time, rays/s. usage: run_configs.py [cfg1 cfg2 cfg4 cfg5 ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light

CFG = {"cfg1": (synth.make_cfg1, (8, 8, 8), 64), "cfg2": (synth.make_cfg2, (32, 16, 32), 256), "cfg4": (synth.make_cfg4, (64, 32, 64), 256),
       "cfg5": (synth.make_cfg5, (128, 128, 128), 256)}
for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]):
    make, res, rays = CFG[name]
    t0 = time.time(); scene = make(); flat = scene_format.flatten(scene); t_scene = time.time() - t0
    g = Context(0); g.scene_upload(flat)
    t0 = time.time(); g.bvh_build(); t_build = time.time() - t0
    info = g.bvh_info()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, rays)
    g.probes_init(grid)
    gen = OrientationGenerator()
    t0 = time.time(); g.probes_classify(gen.next()); t_classify = time.time() - t0
    st = g.probes_download()[2]
    g.probes_upload(state=np.ones_like(st))
    light = Light.default(); ms = []
    for f in range(4):
        grid.hysteresis = min(0.98, 0.3 * f)
        g.probes_update(grid, light, gen.next(), None)
        ms.append(g.probes_timings()["full"])
    irr = g.probes_download()[0]
    print(json.dumps({"config": name, "triangles": int(info.numTriangles), "instances": int(len(flat["instances"])), "bvh_nodes": int(info.numNodes), "bvh_depth": int(info.depth),
                      "bvh_build_ms_device": round(info.buildMs, 2), "bvh_build_s_wall": round(t_build, 3), "scene_gen_s": round(t_scene, 2), "probes": grid.probe_count, "rays_per_probe": rays,
                      "classified_states": np.bincount(st, minlength=9).tolist(), "classify_s": round(t_classify, 3), "update_ms": [round(x, 3) for x in ms],
                      "grays_per_s": round(grid.probe_count * rays / (min(ms[1:]) * 1e-3) / 1e9, 3), "atlas_nonzero": float((irr != 0).mean())}), flush=True)
    g.close()
