"""CPU-only, world_size 2 over gloo: the host-side logic of the sharded update (slice plan, gather layout, rendezvous of
the 128-byte communicator id) with the oracle standing in for the per-rank device work. Checks that the gathered atlases
equal a single-rank update bit for bit, which is the property the NCCL path relies on."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_plan_partitions_the_volume():
    from vulkanexp_b200.sharding import chunk_plan, rank_slices

    for rz in (2, 8, 32, 64, 128, 12):
        for n in (1, 2, 4, 8):
            if rz % n:
                with pytest.raises(ValueError):
                    chunk_plan(rz, n)
                continue
            s, K = chunk_plan(rz, n, 2048)
            assert s * n * K == rz
            owned = sorted(z for r in range(n) for (a, b) in rank_slices(rz, n, r, 2048) for z in range(a, b))
            assert owned == list(range(rz))
            for k in range(K):  # within a chunk, ranks own consecutive equal blocks -> plain all-gather layout
                blocks = [rank_slices(rz, n, r, 2048)[k] for r in range(n)]
                assert all(blocks[r][1] == blocks[r + 1][0] for r in range(n - 1)) and len({b - a for a, b in blocks}) == 1


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    from vulkanexp_b200 import scene_format, synth
    from vulkanexp_b200.host_logic import OrientationGenerator
    from vulkanexp_b200.pods import GridInfo, Light
    from vulkanexp_b200.sharding import chunk_plan, rank_slices

    # rendezvous of an opaque 128-byte id, as bench.py does for ncclUniqueId
    uid = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ids = [None] * world
    dist.all_gather_object(ids, uid[0])
    assert len(set(ids)) == 1 and len(ids[0]) == 128

    flat = scene_format.flatten(synth.make_open_court(columns=2, col_segments=6, col_stacks=1))
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (4, 3, 8), 24, hysteresis=0.0)
    light = Light.default()
    o = pyoracle.Oracle(); o.scene_upload(flat); o.bvh_build(); o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    gen = OrientationGenerator()
    rx, ry, rz = grid.resolution
    plane = rx * ry
    s, K = chunk_plan(rz, world, 4096)  # small plane value -> 2 chunks in this tiny test
    for frame in range(3):
        R = gen.next()
        grid.hysteresis = 0.4 * frame
        # every rank holds the previous frame's full sampled atlases; trace + blend own slices only
        idx = np.array([p for (z0, z1) in rank_slices(rz, world, rank, 4096) for p in range(z0 * plane, z1 * plane)], dtype=np.uint32)
        o.probes_update(grid, light, R, idx, 1)
        irr, dep, st, _ = o.probes_download()
        nirr, ndep, nst = np.zeros_like(irr), np.zeros_like(dep), np.zeros_like(st)
        for k in range(K):  # per chunk: all-gather of contiguous row blocks
            z0 = k * s * world
            for arr, out, rows in ((irr, nirr, 8), (dep, ndep, 16)):
                mine = torch.from_numpy(arr[rows * (z0 + rank * s) : rows * (z0 + (rank + 1) * s)].astype(np.int64))
                parts = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                out[rows * z0 : rows * (z0 + s * world)] = torch.cat(parts).numpy().astype(np.uint32)
            mine = torch.from_numpy(st[(z0 + rank * s) * plane : (z0 + (rank + 1) * s) * plane].astype(np.int64))
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            nst[z0 * plane : (z0 + s * world) * plane] = torch.cat(parts).numpy().astype(np.uint32)
        o.probes_upload(nirr, ndep, nst)  # publish = swap to the gathered atlases
    irr, dep, st, _ = o.probes_download()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), irr=irr, dep=dep, st=st)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_update_equals_single_rank(tmp_path, oracle_lib):
    from vulkanexp_b200 import scene_format, synth
    from vulkanexp_b200.host_logic import OrientationGenerator
    from vulkanexp_b200.pods import GridInfo, Light

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    flat = scene_format.flatten(synth.make_open_court(columns=2, col_segments=6, col_stacks=1))
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (4, 3, 8), 24, hysteresis=0.0)
    o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build(); o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    gen = OrientationGenerator()
    for frame in range(3):
        grid.hysteresis = 0.4 * frame
        o.probes_update(grid, Light.default(), gen.next(), None, 1)
    irr, dep, st, _ = o.probes_download()
    for r in range(2):
        z = np.load(os.path.join(tmp_path, "rank%d.npz" % r))
        assert np.array_equal(z["irr"], irr) and np.array_equal(z["dep"], dep) and np.array_equal(z["st"], st)
