"""Strong-scaling run of BASELINE.json configs[3]: nature-like scene (~2 M instanced triangles), 64x32x64 probes x 256 rays, the
fixed volume sharded over N GPUs (z-slabs + NCCL all-gather). Launch with torchrun like bench.py; N = 1 runs the unsharded update.
Prints one JSON line (secondary measurement; bench.py is the contract)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--res", type=int, nargs=3, default=[64, 32, 64])
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
flat = scene_format.flatten(synth.make_cfg4())
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], tuple(args.res), 256, hysteresis=0.0)
light = Light.default()
ctx = Context(local); ctx.scene_upload(flat); ctx.bvh_build(); ctx.probes_init(grid)
ctx.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
if world > 1:
    uid = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    exchange = os.environ.get("VKX_EXCHANGE", "nccl")
    if exchange == "p2p":
        exchange = "p2p" if ctx.comm_p2p_enable(dist) else "nccl"  # VKX_EXCHANGE=p2p: k_blend stores into every rank's next atlas set over NVLink (default: deferred NCCL all-gather)
else:
    exchange = "none"
gen = OrientationGenerator(); gen.next()
Rs = [gen.next() for _ in range(args.warmup + args.steps)]
stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def step(i, h):
    grid.hysteresis = h
    if world > 1:
        ctx.probes_update_sharded(grid, light, Rs[i], sync=False)
    else:
        ctx.probes_update(grid, light, Rs[i], None, sync=False)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); ctx.sync()


h = 0.0
for w in range(args.warmup):
    step(w, h); h = min(0.98, h + 0.25)
barrier()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
for s in range(args.steps):
    with torch.cuda.stream(stream):
        flush.fill_(s & 0xFF)
    ev[s][0].record(stream); step(args.warmup + s, h); ev[s][1].record(stream)
    ctx.sync()
barrier()
ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
irr, dep, st, _ = ctx.probes_download()
if rank == 0:
    info = ctx.bvh_info()
    print(json.dumps({"metric": "ddgi_full_volume_update_ms", "value": ms, "unit": "ms", "n_gpus": world, "scaling": "strong", "steps": args.steps, "warmup": args.warmup,
                      "probe_rays_per_sec": grid.probe_count * 256 / (ms * 1e-3), "exchange": exchange, "workload": "nature-like scene, %d triangles, %dx%dx%d probes x 256 rays" % (info.numTriangles, *args.res),
                      "atlas_checksum": int(irr.astype(np.uint64).sum() % (1 << 32)), "depth_checksum": int(dep.astype(np.uint64).sum() % (1 << 32))}), flush=True)
if world > 1:
    dist.destroy_process_group()
