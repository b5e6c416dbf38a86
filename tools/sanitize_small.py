"""Small end-to-end pass over every kernel family (run under compute-sanitizer on the GPU box): BVH build, classify, list / full /
scheduled updates with an odd ray count, shadow + reflection + composite frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light, make_camera

flat = scene_format.flatten(synth.make_open_court(columns=2, col_segments=6, col_stacks=1))
g = Context(0); g.scene_upload(flat); g.bvh_build()
for rays in (17, 64):
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (5, 3, 4), rays, hysteresis=0.5)
    g.probes_debug(rays == 17)
    g.probes_init(grid)
    gen = OrientationGenerator(); light = Light.default()
    g.probes_classify(gen.next())
    g.probes_update(grid, light, gen.next())
    g.probes_update(grid, light, gen.next(), np.array([3, 7, 11, 59, 0], dtype=np.uint32))
    n = g.probes_schedule(13); g.probes_update_scheduled(grid, light, gen.next())
    g.probes_download()
W, H = 96, 54
g.shadow_set_noise(synth.blue_noise_like(4, 64)); g.shadow_init(W, H)
cam = make_camera((-5.0, 2.5, 4.5), (0.0, 3.0, 0.0), aspect=W / H, frame_index=1)
g.gbuffer_generate(cam)
ar, em = g.gbuffer_download_material(); ar[..., 3] = 0.2; g.gbuffer_upload_material(ar, em)
for f in range(2):
    g.shadow_frame(cam, cam, light); g.reflection_frame(cam, cam, light); g.final_gather(cam, light)
img, _ = g.final_gather_download()
print("sanitize pass ok", float(img.mean()), g.launch_count())
