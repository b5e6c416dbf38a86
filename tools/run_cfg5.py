"""BASELINE.json configs[4]: synthetic instanced stress scene (~10 M triangles), 128^3 probes x 256 rays, DDGI + 4K sun shadows on
N GPUs (launch with torchrun, one rank per GPU; N = 1 works too). Prints one JSON line (rank 0):

  * BVH build time (device / wall), triangles, nodes, depth
  * sharded full-volume update: ms per step over one CUDA-event interval of K back-to-back steps closed after the last all-gather
    (max over ranks), whole-job and per-GPU Grays/s, all-gather bytes per step, peak device memory
  * `sharded_equals_single`: the gathered 128^3 atlases of two sharded frames at a reduced ray count compared word for word with
    a single-GPU run of the same frames on rank 0
  * the 4K shadow pass (1 spp + X/Y depth-aware Gaussian + temporal accumulation) on the same scene, rank 0

usage: torchrun --nproc-per-node N tools/run_cfg5.py [--res 128 128 128] [--rays 256] [--steps 3] [--check-rays 16] [--scene cfg5|cfg4]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light, make_camera

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, nargs=3, default=[128, 128, 128])
ap.add_argument("--rays", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--check-rays", type=int, default=16)
ap.add_argument("--scene", default="cfg5")
ap.add_argument("--no-shadows", action="store_true")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rmax(x):
    if world > 1:
        t = torch.tensor([x], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
    return x


t0 = time.time()
flat = scene_format.flatten({"cfg5": synth.make_cfg5, "cfg4": synth.make_cfg4}[args.scene]())
t_scene = time.time() - t0
free0, total_mem = torch.cuda.mem_get_info()
g = Context(local); g.scene_upload(flat)
t0 = time.time(); g.bvh_build(); t_build = time.time() - t0
info = g.bvh_info()
light = Light.default()
gen = OrientationGenerator(); gen.next()
Rs = [gen.next() for _ in range(args.warmup + args.steps + 4)]
res = tuple(args.res)
if world > 1:
    uid = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(rank, world, uid[0])


def init_volume(ctx, rays):
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, rays, hysteresis=0.0)
    ctx.probes_init(grid)
    ctx.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    return grid


def update(ctx, grid, R, h, sharded, sync=False):
    grid.hysteresis = h
    if sharded:
        ctx.probes_update_sharded(grid, light, R, sync=sync)
    else:
        ctx.probes_update(grid, light, R, None, sync=sync)


# ---- 1. equality at the full 128^3 volume with a reduced ray count: sharded + gathered == single GPU, word for word
equal = None
if world > 1 and args.check_rays > 0:
    grid = init_volume(g, args.check_rays)
    for f, h in enumerate((0.0, 0.6)):
        update(g, grid, Rs[f], h, True)
    got = g.probes_download()
    if rank == 0:
        ref = Context(local); ref.scene_upload(flat); ref.bvh_build()
        rgrid = init_volume(ref, args.check_rays)
        for f, h in enumerate((0.0, 0.6)):
            update(ref, rgrid, Rs[f], h, False)
        want = ref.probes_download()
        equal = bool(all(np.array_equal(a, b) for a, b in zip(got[:3], want[:3])))
        ref.close(); del ref, want
    del got
    barrier()

# ---- 2. the named workload: 256 rays per probe
grid = init_volume(g, args.rays)
sharded = world > 1
stream = torch.cuda.ExternalStream(g.stream(), device=dev)
h = 0.0
for w in range(args.warmup):
    update(g, grid, Rs[w], h, sharded); h = min(0.98, h + 0.4)
g.sync(); barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for s in range(args.steps):
    update(g, grid, Rs[args.warmup + s], h, sharded)
g.stream_wait_exchange()
b.record(stream)
g.sync(); barrier()
ms = rmax(a.elapsed_time(b) / args.steps)
free1, _ = torch.cuda.mem_get_info()
peak_used = rmax(float(total_mem - free1))
irr, dep, st, _ = g.probes_download()
(ih, iw), (dh, dw) = grid.atlas_shapes()
gather_bytes = (ih * iw + dh * dw + grid.probe_count) * 4
line = None
if rank == 0:
    rays_total = grid.probe_count * args.rays
    line = {"config": args.scene, "n_gpus": world, "triangles": int(info.numTriangles), "instances": int(len(flat["instances"])), "bvh_nodes": int(info.numNodes), "bvh_depth": int(info.depth),
            "bvh_build_ms_device": round(info.buildMs, 2), "bvh_build_s_wall": round(t_build, 3), "scene_gen_s": round(t_scene, 2),
            "probes": grid.probe_count, "resolution": list(res), "rays_per_probe": args.rays, "steps": args.steps, "warmup": args.warmup,
            "update_ms": ms, "grays_per_s": rays_total / (ms * 1e-3) / 1e9, "grays_per_s_per_gpu": rays_total / (ms * 1e-3) / 1e9 / world,
            "timing": "one CUDA-event interval over the steps on the library's stream, closed after the last all-gather; max over ranks",
            "allgather_bytes_per_step": int(gather_bytes) if sharded else 0, "allgather_bytes_sent_per_rank": int(gather_bytes // world) if sharded else 0,
            "peak_device_memory_gb": peak_used / 2**30, "device_memory_total_gb": total_mem / 2**30,
            "sharded_equals_single": equal, "check_rays_per_probe": args.check_rays if world > 1 else None,
            "atlas_nonzero": float((irr != 0).mean()), "atlas_checksum": int(irr.astype(np.uint64).sum() % (1 << 32)), "depth_checksum": int(dep.astype(np.uint64).sum() % (1 << 32))}
del irr, dep, st

# ---- 3. 4K sun shadows on the same scene (rank 0; the pass is specified for one GPU)
if rank == 0 and not args.no_shadows:
    W, H = 3840, 2160
    g.shadow_set_noise(synth.reference_blue_noise(64)); g.shadow_init(W, H)
    lo, hi = np.array(flat["bounds_min"]), np.array(flat["bounds_max"])
    c, ext = (lo + hi) / 2, hi - lo
    cams = [make_camera((c[0] - 0.3 * ext[0] + 0.01 * ext[0] * f, hi[1] * 0.6 + 10.0, c[2] - 0.3 * ext[2] + 0.008 * ext[2] * f), (c[0] + 0.02 * ext[0] * f, lo[1], c[2]), aspect=W / H, frame_index=f) for f in range(10)]
    prev, sms = cams[0], []
    for cam in cams:
        g.gbuffer_generate(cam); g.shadow_frame(cam, prev, light); sms.append(g.shadow_timings()); prev = cam
    steady = sms[4:]
    pd, _ = g.gbuffer_download()
    line["shadow_pass_4k"] = {k: float(np.mean([m[k] for m in steady])) for k in steady[0]}
    line["shadow_pass_4k"]["geometry_pixels"] = float((pd[..., 3] > 0).mean())
if rank == 0:
    print(json.dumps(line), flush=True)
if world > 1:
    barrier()
    dist.destroy_process_group()
