"""Sun-shadow pass timing (BASELINE.json configs[2]): synthetic dungeon-like scene, 3840x2160, 1 spp + X/Y depth-aware
Gaussian + temporal accumulation, camera path of 16 frames. Prints one JSON line (secondary metric; bench.py is the contract)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light, make_camera

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ALPHA = os.environ.get("VKX_CFG3_ALPHA", "0") == "1"  # cfg3 with the cut-out (alpha-textured) grates: shadow rays run the any-hit test
s = synth.make_cfg3(alpha_grates=ALPHA); flat = scene_format.flatten(s)
g = Context(0); g.scene_upload(flat); g.bvh_build()
info = g.bvh_info()
g.shadow_set_noise(synth.blue_noise_like(64, 64)); g.shadow_init(W, H)
light = Light.default()
# a lit irradiance volume for the composite (FinalGather): 32 x 8 x 32 probes x 256 rays, 6 updates
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (32, 8, 32), 256, hysteresis=0.7)
g.probes_init(grid); gen = OrientationGenerator(); g.probes_classify(gen.next())
for _ in range(6):
    g.probes_update(grid, light, gen.next())
cams = [make_camera((-20.0 + 1.2 * f, 2.2, -18.0 + 0.9 * f), (0.0 + 0.5 * f, 1.5, 0.0), aspect=W / H, frame_index=f) for f in range(16)]
prev = cams[0]; ms = []; gb = []; img = np.zeros((H, W, 4), dtype=np.float32)
for f, cam in enumerate(cams):
    t0 = time.perf_counter(); g.gbuffer_generate(cam); gb.append((time.perf_counter() - t0) * 1e3)
    g.shadow_frame(cam, prev, light)
    t = g.shadow_timings()
    # the dungeon's materials are rough dielectrics; polish the floor (normal.y > 0.9 -> roughness 0.15) so the reflection pass has work
    ar, em = g.gbuffer_download_material(); nmm = g.gbuffer_download()[1]
    ar[..., 3] = np.where(nmm[..., 1] > 0.9, np.float32(0.15), ar[..., 3])
    g.gbuffer_upload_material(ar, em)
    g.reflection_frame(cam, prev, light)
    rt = g.reflection_timings()
    t.update({"reflection": rt["full"], "reflection_trace_shade": rt["trace_shade"], "reflection_filter_x": rt["filter_x"], "reflection_filter_y": rt["filter_y"]})
    g.final_gather(cam, light)
    t["gather"] = g.final_gather_download(out=img)[1]
    ms.append(t); prev = cam
pd, _ = g.gbuffer_download()
steady = ms[4:]
avg = {k: float(np.mean([m[k] for m in steady])) for k in steady[0]}
px = W * H
print(json.dumps({"metric": "sun_shadow_pass_ms", "alpha_grates": ALPHA, "value": avg["full"], "unit": "ms", "width": W, "height": H, "triangles": int(info.numTriangles), "stages_ms": avg,
                  "gbuffer_fixture_ms": float(np.mean(gb[4:])), "geometry_pixels": float((pd[..., 3] > 0).mean()),
                  "filter_bytes_per_px": 48 + 64, "filter_gbs": px * (48 + 64) / ((avg["filter_x"] + avg["filter_y"]) * 1e-3) / 1e9,
                  "shadow_rays_per_s": px * float((pd[..., 3] > 0).mean()) / (avg["trace"] * 1e-3),
                  "reflection_ms": avg["reflection"], "reflecting_pixels": float((g.reflection_download(0)[..., 3] > 0).mean()), "final_gather_ms": avg["gather"], "final_gather_bytes_per_px": 5 * 16 + 16, "final_gather_gbs": px * 96 / (avg["gather"] * 1e-3) / 1e9,
                  "composite_mean_rgb": [float(v) for v in img[..., :3].mean(axis=(0, 1))]}))
if len(sys.argv) > 3:  # optional: write the last composite as a PPM (Reinhard + gamma 2.2) for a look at the picture
    rgb = img[..., :3] / (1.0 + img[..., :3]); rgb = (np.clip(rgb, 0, 1) ** (1 / 2.2) * 255 + 0.5).astype(np.uint8)
    with open(sys.argv[3], "wb") as f:
        f.write(b"P6 %d %d 255\n" % (W, H)); f.write(rgb.tobytes())
