"""Small end-to-end pass over every kernel family (run under compute-sanitizer on the GPU box): BVH build, classify, list / full /
scheduled updates with an odd ray count, shadow + reflection + composite frame; then the textured scene (mip blits, decoded arena, textured
closest-hit shading, cut-outs in the shadow / reflection traversals, texture look-up entry point) and a skinned instance (skinning kernel +
rebuild). The deferred-leaf traversal runs by default for the shadow rays; VKX_PT_DEFER=16 covers it for the primary rays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light, make_camera

flat = scene_format.flatten(synth.make_open_court(columns=2, col_segments=6, col_stacks=1))
g = Context(0); g.scene_upload(flat); g.bvh_build()
g.instances_update(flat["instances"]); g.bvh_refit()  # topology-preserving refit kernels
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
# textured scene + skinned instance
tflat, src, dst, size = scene_format.add_skinned_instance(scene_format.flatten(synth.make_textured_court()), 2)
t = Context(0); t.scene_upload(tflat); t.bvh_build()
t.texture_sample(0, np.random.default_rng(0).uniform(-1, 2, (64, 2)).astype(np.float32), np.full((64, 4), 0.01, dtype=np.float32)); t.texture_download(3)
jt = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (2, 1)); jt[1, 12] = 0.5
t.skin_vertices(jt, np.tile(np.array([[0, 1, 0, 1]], dtype=np.uint16), (size, 1)), np.full((size, 4), 0.25, dtype=np.float32), src, dst, motion=True); t.bvh_build()
grid = GridInfo.make(tflat["bounds_min"], tflat["bounds_max"], (4, 3, 4), 32, hysteresis=0.5)
t.probes_init(grid); t.probes_classify(gen.next()); t.probes_update(grid, light, gen.next())
t.shadow_set_noise(synth.blue_noise_like(4, 64)); t.shadow_init(W, H); t.gbuffer_generate(cam)
ar, em = t.gbuffer_download_material(); ar[..., 3] = 0.2; t.gbuffer_upload_material(ar, em)
t.shadow_frame(cam, cam, light); t.reflection_frame(cam, cam, light); t.final_gather(cam, light)
t.trace(np.zeros((8, 3), dtype=np.float32) + [0, 6.5, 0], np.tile(np.array([[0.1, -1, 0.1]], dtype=np.float32), (8, 1)), 0.01, 100.0, alpha_test=True)
print("sanitize pass ok", float(img.mean()), g.launch_count(), t.launch_count())
