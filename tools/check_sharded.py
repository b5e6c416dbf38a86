"""Multi-GPU check (run with torchrun, one rank per GPU): the sharded update + NCCL all-gather must give every rank
atlases identical to a single-GPU update. Prints one line per rank and exits non-zero on mismatch. VKX_P2P=1 / VKX_P2P=ce select the
peer-memory exchange (1: fused peer stores from the CUDA-core blend - run it with VKX_BLEND=simt so that the single-GPU side uses the same kernel; ce: copy-engine pushes, tensor-core blend on both sides)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200._lib import Context
from vulkanexp_b200.host_logic import OrientationGenerator
from vulkanexp_b200.pods import GridInfo, Light

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
flat = scene_format.flatten(synth.make_open_court())
res = tuple(int(x) for x in os.environ.get("VKX_CHECK_RES", "8,6,%d" % (8 * world)).split(","))
RAYS = int(os.environ.get("VKX_CHECK_RAYS", "64"))
grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, RAYS)
light = Light.default()
ctx = Context(local); ctx.scene_upload(flat); ctx.bvh_build(); ctx.probes_init(grid)
ones = np.ones(grid.probe_count, dtype=np.uint32)
ctx.probes_upload(state=ones)
uid = [Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
P2P = os.environ.get("VKX_P2P", "0") in ("1", "ce")  # 1: blend fused with peer stores (compare with VKX_BLEND=simt); ce: copy-engine pushes
if P2P:  # atlas exchange over NVLink peer memory instead of the NCCL all-gather
    assert ctx.comm_p2p_enable(dist), "peer atlases could not be mapped"
    if os.environ["VKX_P2P"] == "ce":
        ctx.comm_p2p_mode(1)
    print("rank %d: peer-memory exchange enabled (%s)" % (rank, "copy engines" if os.environ["VKX_P2P"] == "ce" else "fused peer stores"), flush=True)
ref = Context(local); ref.scene_upload(flat); ref.bvh_build(); ref.probes_init(grid); ref.probes_upload(state=ones)
gen = OrientationGenerator()
ok = True
for frame in range(4):
    R = gen.next()
    grid.hysteresis = min(0.9, 0.3 * frame)
    ctx.probes_update_sharded(grid, light, R)
    ref.probes_update(grid, light, R, None)
    a, b = ctx.probes_download(), ref.probes_download()
    same = all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
    ok &= same
    print("rank %d frame %d: sharded == single-GPU: %s (sharded %.3f ms, single %.3f ms)" % (rank, frame, same, ctx.probes_timings()["full"], ref.probes_timings()["full"]), flush=True)
# own-slab read-back queued right behind a sharded update (it reads the rank's work atlases and does not wait for the all-gather):
# must equal the same rows of the single-GPU atlases, frame after frame, with the next update already queued
(ih, iw), (dh, dw) = grid.atlas_shapes()
from vulkanexp_b200._lib import shard_slices
slices = shard_slices(res[2], world, rank)
bufs = [[(np.zeros((8 * (z1 - z0), iw), np.uint32), np.zeros((16 * (z1 - z0), dw), np.uint32), np.zeros((z1 - z0) * res[0] * res[1], np.uint32)) for (z0, z1) in slices] for _ in range(2)]
for frame in range(4, 8):
    R = gen.next()
    ctx.probes_update_sharded(grid, light, R, sync=False)
    for (z0, z1), buf in zip(slices, bufs[frame & 1]):
        ctx.probes_download_slab_async(z0, z1, buf)
    ref.probes_update(grid, light, R, None)
    ctx.probes_download_wait()
    b = ref.probes_download()
    same = True
    for (z0, z1), got in zip(slices, bufs[frame & 1]):
        same &= bool(np.array_equal(got[0], b[0][8 * z0:8 * z1]) and np.array_equal(got[1], b[1][16 * z0:16 * z1]) and np.array_equal(got[2], b[2][z0 * res[0] * res[1]:z1 * res[0] * res[1]]))
    ok &= same
    print("rank %d frame %d: async own-slice read-back == single-GPU rows: %s" % (rank, frame, same), flush=True)
a, b = ctx.probes_download(), ref.probes_download()
same = all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
ok &= same
print("rank %d: full atlases after the async frames equal: %s" % (rank, same), flush=True)
# list updates (ProbesPerUpdate scheduling) on all ranks: every rank passes the same list, traces its share, the tiles are exchanged
# as packed records; must equal the single-GPU list update; alternate with full-volume sharded frames (pending deferred gather)
if not P2P:
    rng = np.random.default_rng(5)
    for frame in range(8, 14):
        R = gen.next()
        if frame % 3 == 2:
            ctx.probes_update_sharded(grid, light, R, sync=False); ref.probes_update(grid, light, R, None)
        else:
            lst = rng.permutation(grid.probe_count).astype(np.uint32)[: int(rng.integers(1, grid.probe_count))]
            ctx.probes_update_sharded_list(grid, light, R, lst, sync=False); ref.probes_update(grid, light, R, lst)
        a, b = ctx.probes_download(), ref.probes_download()
        same = all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
        ok &= same
        print("rank %d frame %d: sharded %s update == single-GPU: %s" % (rank, frame, "full" if frame % 3 == 2 else "list", same), flush=True)
t = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(int(t.item() != 0))
