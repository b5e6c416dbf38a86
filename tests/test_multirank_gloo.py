"""CPU-only, world_size 2 over gloo: the host-side logic of the two sharded updates, driven by the sharding arithmetic the
library itself uses (vkx_shard_groups / vkx_shard_slices / vkx_shard_range, exported host-only through the C ABI: csrc/api.cu), with the oracle
standing in for the per-rank device work.
  * full-volume update: z-slab per rank, all-gather of contiguous atlas rows and state words;
  * list update (ProbesPerUpdate scheduling): list positions per rank, all-gather of packed 1296-byte tile records, scatter into
    the atlases in list order (the layout k_pack_tiles / k_unpack_tiles use).
Both must reproduce a single-rank update bit for bit, which is the property the NCCL path relies on."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES, RAYS = (4, 3, 8), 24


def test_slice_arithmetic_partitions_the_volume():
    from vulkanexp_b200._lib import VkxError, shard_groups, shard_slices

    for rz in (2, 8, 32, 64, 128, 12, 3):
        for n in (1, 2, 3, 4, 8):
            if rz % n:
                with pytest.raises(VkxError):
                    shard_groups(rz, n)
                continue
            s, groups = shard_groups(rz, n)
            assert s * groups * n == rz and (s == 2 or groups == 1)        # pairs of slices (whole 2x2x2 probe blocks), else one slab
            per_rank = [shard_slices(rz, n, r) for r in range(n)]
            assert all(len(p) == groups and all(b - a == s for a, b in p) for p in per_rank)  # equal counts: ncclAllGather takes one count
            for g in range(groups):  # the ranks' slices of one group are consecutive, in rank order: a plain all-gather layout
                assert per_rank[0][g][0] == g * s * n
                assert all(per_rank[r][g][1] == per_rank[r + 1][g][0] for r in range(n - 1))
            covered = sorted(z for p in per_rank for a, b in p for z in range(a, b))
            assert covered == list(range(rz))                                # every slice exactly once
    with pytest.raises(VkxError):
        shard_slices(8, 2, 2)


def test_list_range_arithmetic_covers_every_position_once():
    from vulkanexp_b200._lib import shard_range

    for count in (0, 1, 5, 7, 8, 100, 16384, 131071):
        for n in (1, 2, 3, 4, 8):
            per = (count + n - 1) // n
            seen = []
            for r in range(n):
                f, c = shard_range(count, n, r)
                assert c <= per and (c == per or f + c == count)  # only the tail is short
                assert f == min(count, r * per)                    # rank r's records start at r * per in the gathered buffer
                seen.extend(range(f, f + c))
            assert seen == list(range(count))


def _tiles(irr, dep, st, grid, probe):
    rx, ry, _ = grid.resolution
    ix, iy, iz = probe % rx, (probe % (rx * ry)) // rx, probe // (rx * ry)
    t = iy * rx + ix
    return irr[8 * iz:8 * iz + 8, 8 * t:8 * t + 8], dep[16 * iz:16 * iz + 16, 16 * t:16 * t + 16]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    from vulkanexp_b200 import scene_format, synth
    from vulkanexp_b200._lib import shard_groups, shard_range, shard_slices
    from vulkanexp_b200.pods import GridInfo, Light

    # rendezvous of an opaque 128-byte id, as bench.py does for ncclUniqueId
    uid = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ids = [None] * world
    dist.all_gather_object(ids, uid[0])
    assert len(set(ids)) == 1 and len(ids[0]) == 128

    flat = scene_format.flatten(synth.make_open_court(columns=2, col_segments=6, col_stacks=1))
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], RES, RAYS, hysteresis=0.0)
    light = Light.default()
    o = pyoracle.Oracle(); o.scene_upload(flat); o.bvh_build(); o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    host = pyoracle.HostLogic()
    rx, ry, rz = grid.resolution
    plane = rx * ry

    def gather(t):
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return torch.cat(parts)

    for frame in range(4):
        R, _ = host.next_orientation()
        grid.hysteresis = 0.3 * frame
        if frame % 2 == 0:  # ---- full-volume update: own z-slices, one all-gather of contiguous rows per slice group
            mine = shard_slices(rz, world, rank)
            idx = np.concatenate([np.arange(z0 * plane, z1 * plane, dtype=np.uint32) for z0, z1 in mine])
            o.probes_update(grid, light, R, idx, 1)
            irr, dep, st, _ = o.probes_download()
            nirr, ndep, nst = irr.copy(), dep.copy(), st.copy()
            gs, groups = shard_groups(rz, world)
            for g, (z0, z1) in enumerate(mine):  # group g: rows of the slices [g gs world, (g + 1) gs world), rank order
                lo, hi = g * gs * world, (g + 1) * gs * world
                nirr[8 * lo:8 * hi] = gather(torch.from_numpy(irr[8 * z0:8 * z1].astype(np.int64))).numpy().astype(np.uint32)
                ndep[16 * lo:16 * hi] = gather(torch.from_numpy(dep[16 * z0:16 * z1].astype(np.int64))).numpy().astype(np.uint32)
                nst[lo * plane:hi * plane] = gather(torch.from_numpy(st[z0 * plane:z1 * plane].astype(np.int64))).numpy().astype(np.uint32)
            o.probes_upload(nirr, ndep, nst)  # publish = swap to the gathered atlases
        else:  # ---- list update: own list positions, all-gather of packed tile records, scatter in list order
            lst = np.random.default_rng(100 + frame).permutation(grid.probe_count).astype(np.uint32)[: 37 + frame]
            count = len(lst)
            per = (count + world - 1) // world
            first, mine = shard_range(count, world, rank)
            o.probes_update(grid, light, R, lst[first:first + mine], 1)
            irr, dep, st, _ = o.probes_download()
            rec = np.zeros((per, 81 * 4), dtype=np.int64)  # 64 uint4 depth + 16 uint4 irradiance + (state, index, 0, 0)
            for j, p in enumerate(lst[first:first + mine]):
                ti, td = _tiles(irr, dep, st, grid, int(p))
                rec[j, :256] = td.reshape(-1); rec[j, 256:320] = ti.reshape(-1); rec[j, 320] = st[p]; rec[j, 321] = p
            allrec = gather(torch.from_numpy(rec)).numpy()
            for s, p in enumerate(lst):  # list position s = rank s // per, item s % per = record s
                ti, td = _tiles(irr, dep, st, grid, int(p))
                td[...] = allrec[s, :256].reshape(16, 16); ti[...] = allrec[s, 256:320].reshape(8, 8); st[p] = allrec[s, 320]
                assert allrec[s, 321] == p
            o.probes_upload(irr, dep, st)
    irr, dep, st, _ = o.probes_download()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), irr=irr, dep=dep, st=st)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_updates_equal_single_rank(tmp_path, oracle_lib):
    from vulkanexp_b200 import scene_format, synth
    from vulkanexp_b200.pods import GridInfo, Light

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    flat = scene_format.flatten(synth.make_open_court(columns=2, col_segments=6, col_stacks=1))
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], RES, RAYS, hysteresis=0.0)
    o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build(); o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    host = oracle_lib.HostLogic()
    for frame in range(4):
        R, _ = host.next_orientation()
        grid.hysteresis = 0.3 * frame
        lst = None if frame % 2 == 0 else np.random.default_rng(100 + frame).permutation(grid.probe_count).astype(np.uint32)[: 37 + frame]
        o.probes_update(grid, Light.default(), R, lst, 1)
    irr, dep, st, _ = o.probes_download()
    for r in range(2):
        z = np.load(os.path.join(tmp_path, "rank%d.npz" % r))
        assert np.array_equal(z["irr"], irr) and np.array_equal(z["dep"], dep) and np.array_equal(z["st"], st)
