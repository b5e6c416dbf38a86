#!/usr/bin/env python
"""Benchmark of the DDGI probe update (BASELINE.json metric: probe rays/s and full-volume update ms at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            the CUDA path (libvkexp_b200.so through its C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the CPU transliteration (oracle/) on the host cores

A step is one full-volume DDGI update: every probe of the volume traces raysPerProbe rays (closest hit, shading with sun
shadow ray and two sampleProbes look-ups, or sky on a miss), blends them into both atlases with hysteresis, writes the
border texels and publishes.

  N = 1   BASELINE.json configs[1]: 32x16x32 probes x 256 rays on the synthetic Sponza-scale atrium (the named
          data/sponza_test.scene is not in the reference checkout). Per-step CUDA events, 256 MiB L2 flush between steps.
  N > 1   BASELINE.json configs[3] (--scaling strong, the default): the nature-like scene (2.24 M instanced triangles), a fixed
          64x32x64 volume x 256 rays cut into N z-slabs, NCCL all-gather of the atlas slabs. The K steps are timed in ONE event
          interval without host synchronisation between steps, ending after the last all-gather has landed, so the exchange is
          inside the measurement (it is designed to overlap the next step's primary traversal). The line also carries the same
          workload on one GPU (rank 0, same run) and `sharded_equals_single`: the gathered atlases of a sharded run compared
          word for word with a single-GPU run of the same volume.
          --scaling weak: 32x16x(32 N) probes over the atrium (round-1 behaviour).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vulkanexp_b200 import scene_format, synth  # noqa: E402
from vulkanexp_b200.pods import GridInfo, Light  # noqa: E402

RAYS = 256
NODE_BYTES, TRI_BYTES, HIT_BYTES = 80, 48, 20
CFG2_RES, CFG4_RES = (32, 16, 32), (64, 32, 64)


def workload(n, scaling):
    """(name, scene maker, grid resolution) of the run."""
    if n == 1 or scaling == "weak":
        res = (CFG2_RES[0], CFG2_RES[1], CFG2_RES[2] * n)
        return "DDGI full-volume update, synthetic sponza-scale atrium (265k triangles), %dx%dx%d probes x %d rays" % (*res, RAYS), synth.make_cfg2, res
    return "DDGI full-volume update, synthetic nature-like scene (2.24M instanced triangles), %dx%dx%d probes x %d rays, z-slices dealt over %d ranks" % (*CFG4_RES, RAYS, n), synth.make_cfg4, CFG4_RES


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def orientations(k):
    """The reference's per-frame random rotations: MSVC-LCG replay of glm::sphericalRand + genBasis. Host-side harness input (the
    C ABI takes the matrix); the product arm computes them with the C++ facade's generator so that it stays oracle-free."""
    from vulkanexp_b200.host_logic import OrientationGenerator

    gen = OrientationGenerator()
    gen.next()  # the first draw is consumed by initProbes (reference src/IrradianceProbes.cpp:361)
    return [gen.next() for _ in range(k)]


def oracle_orientations(k):
    """Same sequence from the oracle's own host logic (bit-identical, tests/golden/glm_pin.json): the reference arm loads no
    product library."""
    from oracle import pyoracle

    host = pyoracle.HostLogic()
    host.next_orientation()
    return [host.next_orientation()[0] for _ in range(k)]


def cpu_baseline(flat, grid, light, sample_probes):
    """Times the oracle (CPU transliteration) once on a bounded sample of the same workload with every host core this process may
    use; also returns its traversal counters (the per-ray node / triangle means of the roofline)."""
    from oracle import pyoracle

    cores = pyoracle.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1
    o = pyoracle.Oracle()
    o.scene_upload(flat)
    o.bvh_build()
    o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    idx = np.linspace(0, grid.probe_count - 1, sample_probes).astype(np.uint32)
    R = oracle_orientations(3)
    o.probes_update(grid, light, R[0], idx, 0)  # warm-up: fills the atlases so sampleProbes does real work
    sec = [o.probes_update(grid, light, R[1 + k], idx, 0) for k in range(2)]
    c = o.probes_counters()
    rays = len(idx) * grid.raysPerProbe
    return {
        "value": rays / min(sec), "unit": "probe rays/s", "cores": cores, "kind": "port",
        "sample": "%d of %d probes (evenly spaced) x %d rays, best of 2 updates after one warm-up update, %.2f s per update; the reference itself (Vulkan RT shaders, Windows) cannot run here, so this is its CPU transliteration (oracle/), pinned function by function against the reference's GLSL (oracle/_ref)" % (len(idx), grid.probe_count, grid.raysPerProbe, min(sec)),
    }, c, rays


_RESULT_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries may write to file descriptor 1 (NCCL prints its version line there when
    NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the rest of the run and the result line goes to a
    duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _RESULT_OUT


def _emit(line):
    out = _claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path = its transliteration (oracle/, kind "port": the reference is Vulkan RT
    shaders under Windows and cannot be built here), with every host core, on this arm's workload; each step a bounded sample of
    the volume. Loads nothing from vulkanexp_b200/ but the scene generator and the POD definitions."""
    if rank != 0:
        return
    from oracle import pyoracle

    cores = pyoracle.use_all_cores()
    name, maker, res = workload(args.gpus, args.scaling)
    flat = scene_format.flatten(maker())
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, RAYS, hysteresis=0.9)
    light = Light.default()
    o = pyoracle.Oracle()
    o.scene_upload(flat)
    o.bvh_build()
    o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    sample = min(grid.probe_count, args.ref_sample)
    idx = np.linspace(0, grid.probe_count - 1, sample).astype(np.uint32)
    Rs = oracle_orientations(args.warmup + args.steps)
    for w in range(args.warmup):
        o.probes_update(grid, light, Rs[w], idx, 0)
    total = 0.0
    for s in range(args.steps):
        total += o.probes_update(grid, light, Rs[args.warmup + s], idx, 0)
    rays = sample * RAYS
    value = rays * args.steps / total
    line = {
        "impl": "reference", "metric": "ddgi_probe_rays_per_sec", "value": value, "unit": "probe rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak" if (args.gpus == 1 or args.scaling == "weak") else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "sample": "%d of %d probes per step" % (sample, grid.probe_count)},
        "cpu_baseline": {"value": value, "unit": "probe rays/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d probes (evenly spaced) x %d rays per step; the reference itself (Vulkan RT shaders, Windows) cannot run here, so this is its CPU transliteration (oracle/)" % (sample, grid.probe_count, RAYS)},
        "e2e": {"value": value, "unit": "probe rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "full_volume_update_ms_extrapolated": 1e3 * total / args.steps * grid.probe_count / sample,
    }
    _emit(line)


def secondary_shadow_pass(device, frames=12):
    """BASELINE.json configs[2]: 1-spp sun shadows + depth-aware Gaussian + temporal accumulation at 3840x2160 on the dungeon-like
    scene with cut-out grates, jittered with the reference's blue-noise slices; mean device time of the steady frames."""
    from vulkanexp_b200._lib import Context
    from vulkanexp_b200.pods import make_camera

    W, H = 3840, 2160
    flat = scene_format.flatten(synth.make_cfg3(alpha_grates=True))
    g = Context(device)
    g.scene_upload(flat); g.bvh_build()
    g.shadow_set_noise(synth.reference_blue_noise(64)); g.shadow_init(W, H)
    light = Light.default()
    cams = [make_camera((-20.0 + 1.2 * f, 2.2, -18.0 + 0.9 * f), (0.0 + 0.5 * f, 1.5, 0.0), aspect=W / H, frame_index=f) for f in range(frames)]
    prev, ms = cams[0], []
    for cam in cams:
        g.gbuffer_generate(cam)
        g.shadow_frame(cam, prev, light)
        ms.append(g.shadow_timings()); prev = cam
    steady = ms[4:]
    avg = {k: float(np.mean([m[k] for m in steady])) for k in steady[0]}
    pd, _ = g.gbuffer_download()
    px = W * H
    out = {"shadow_pass_ms": avg["full"], "stages_ms": avg, "width": W, "height": H, "triangles": int(g.bvh_info().numTriangles), "noise": "reference data/BlueNoise/64_64/LDR_RGBA_0..63.png (tests/golden fixture)",
           "alpha_grates": True, "geometry_pixels": float((pd[..., 3] > 0).mean()), "shadow_rays_per_s": px * float((pd[..., 3] > 0).mean()) / (avg["trace"] * 1e-3),
           "filter_bytes_per_px": 112, "filter_gbs": px * 112 / ((avg["filter_x"] + avg["filter_y"]) * 1e-3) / 1e9}
    g.close()
    return out


def cfg5_leg(local, rank, world, torch, dist, light, steps=3, check_rays=16):
    """BASELINE configs[4] inside the 8-GPU run: the synthetic instanced stress scene (~10 M triangles), 128^3 probes x 256 rays,
    sharded full-volume updates with the NVLink all-gather inside one timed interval, the gathered atlases of two frames at a reduced
    ray count compared word for word with a single-GPU run on rank 0, and the 4K sun-shadow pass on the same scene. Collective: every
    rank calls it; a rank that fails before the first collective makes all ranks return an error record instead of hanging."""
    from vulkanexp_b200._lib import Context
    from vulkanexp_b200.pods import make_camera

    dev = torch.device("cuda", local)
    ok, err, c5, flat5, info = 1, None, None, None, None
    t0 = time.perf_counter()
    try:
        flat5 = scene_format.flatten(synth.make_cfg5())
        c5 = Context(local); c5.scene_upload(flat5); c5.bvh_build(); info = c5.bvh_info()
    except Exception as e:
        ok, err = 0, repr(e)
    flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return {"error": err or "another rank failed to set the scene up"}
    uid = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c5.comm_init(rank, world, uid[0])
    res = (128, 128, 128)
    Rs = orientations(steps + 8)

    def init_volume(ctx, rays):
        g = GridInfo.make(flat5["bounds_min"], flat5["bounds_max"], res, rays, hysteresis=0.0)
        ctx.probes_init(g); ctx.probes_upload(state=np.ones(g.probe_count, dtype=np.uint32))
        return g

    def sync_all():
        dist.barrier(); torch.cuda.synchronize()

    # sharded + gathered == single GPU at the full 128^3 volume (reduced ray count keeps the single-GPU side short)
    equal = None
    g16 = init_volume(c5, check_rays)
    for f, h in enumerate((0.0, 0.6)):
        g16.hysteresis = h; c5.probes_update_sharded(g16, light, Rs[f], sync=False)
    got = c5.probes_download()
    if rank == 0:
        ref = Context(local); ref.scene_upload(flat5); ref.bvh_build()
        r16 = init_volume(ref, check_rays)
        for f, h in enumerate((0.0, 0.6)):
            r16.hysteresis = h; ref.probes_update(r16, light, Rs[f], None, sync=False)
        want = ref.probes_download()
        equal = bool(all(np.array_equal(a, b) for a, b in zip(got[:3], want[:3])))
        ref.close(); del ref, want
    del got
    sync_all()
    # the named workload
    grid5 = init_volume(c5, RAYS)
    stream = torch.cuda.ExternalStream(c5.stream(), device=dev)
    h = 0.0
    for w in range(2):
        grid5.hysteresis = h; c5.probes_update_sharded(grid5, light, Rs[w], sync=False); h = min(0.98, h + 0.4)
    c5.sync(); sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for s in range(steps):
        grid5.hysteresis = h; c5.probes_update_sharded(grid5, light, Rs[2 + s], sync=False)
    c5.stream_wait_exchange()
    b.record(stream)
    c5.sync(); sync_all()
    t = torch.tensor([a.elapsed_time(b) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    (ih, iw), (dh, dw) = grid5.atlas_shapes()
    out = None
    if rank == 0:
        rays_total = grid5.probe_count * RAYS
        out = {"workload": "BASELINE configs[4]: synthetic instanced stress scene, 128x128x128 probes x 256 rays, probe z-slices dealt round robin over %d ranks" % world, "triangles": int(info.numTriangles), "instances": int(len(flat5["instances"])),
               "bvh_nodes": int(info.numNodes), "bvh_build_ms": round(float(info.buildMs), 2), "update_ms": ms, "value": rays_total / (ms * 1e-3), "unit": "probe rays/s", "steps": steps,
               "timing": "one CUDA-event interval over the steps, closed after the last all-gather; max over ranks", "allgather_bytes_per_update": int((ih * iw + dh * dw + grid5.probe_count) * 4),
               "sharded_equals_single": equal, "check_rays_per_probe": check_rays}
        try:  # 4K shadows on the same scene (rank 0: the pass is specified for one GPU)
            W, H = 3840, 2160
            c5.shadow_set_noise(synth.reference_blue_noise(64)); c5.shadow_init(W, H)
            lo, hi = np.array(flat5["bounds_min"]), np.array(flat5["bounds_max"])
            c, ext = (lo + hi) / 2, hi - lo
            cams = [make_camera((c[0] - 0.3 * ext[0] + 0.01 * ext[0] * f, hi[1] * 0.6 + 10.0, c[2] - 0.3 * ext[2] + 0.008 * ext[2] * f), (c[0] + 0.02 * ext[0] * f, lo[1], c[2]), aspect=W / H, frame_index=f) for f in range(10)]
            prev, sms = cams[0], []
            for cam in cams:
                c5.gbuffer_generate(cam); c5.shadow_frame(cam, prev, light); sms.append(c5.shadow_timings()); prev = cam
            out["shadow_pass_4k_ms"] = float(np.mean([m["full"] for m in sms[4:]]))
        except Exception as e:
            out["shadow_pass_4k_ms"] = repr(e)
        out["wall_s"] = round(time.perf_counter() - t0, 1)
    sync_all()
    c5.close()
    return out


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vkx", choices=["vkx", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: strong = BASELINE configs[3] (fixed 64x32x64 volume on the nature-like scene), weak = 32x16x32N on the atrium")
    ap.add_argument("--ref-sample", type=int, default=4096, help="probes per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="probes of the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg3 shadow pass and the cfg4 single-GPU line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "vkx" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from vulkanexp_b200._lib import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = max(world, 1)
    scaling = "weak" if (n == 1 or args.scaling == "weak") else "strong"
    name, maker, res = workload(n, args.scaling)
    flat = scene_format.flatten(maker())
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, RAYS, hysteresis=0.0)
    light = Light.default()
    dev = torch.device("cuda", local)

    def make_ctx():
        c = Context(local)
        c.scene_upload(flat)
        c.bvh_build()
        c.probes_init(grid)
        c.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))  # every probe active: the named workload traces all of them
        return c

    ctx = make_ctx()
    if world > 1:
        uid = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
        exchange = os.environ.get("VKX_EXCHANGE", "ce")  # ce (default) | nccl | p2p; ce / p2p fall back to nccl when the peers' atlases cannot be mapped
        if exchange == "p2p":  # VKX_EXCHANGE=p2p: blend fused with the atlas exchange over NVLink peer memory (measured slower than the deferred all-gather at 8 GPUs)
            exchange = "p2p" if ctx.comm_p2p_enable(dist) else "nccl"
        elif exchange == "ce":  # VKX_EXCHANGE=ce: rows pushed into the peers' atlases by copy engines (no SM takes part: runs beside the next step's traversal)
            if ctx.comm_p2p_enable(dist):
                ctx.comm_p2p_mode(1)
            else:
                exchange = "nccl"
    else:
        exchange = "none"
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    K, Wm = args.steps, args.warmup
    Rs = orientations(Wm + 3 * K + args.e2e_steps + 8)
    probes_per_rank = grid.probe_count // n
    rays_per_step_total = grid.probe_count * RAYS

    def step(c, i, hyst, sharded):
        grid.hysteresis = hyst
        if sharded:
            c.probes_update_sharded(grid, light, Rs[i], sync=False)
        else:
            c.probes_update(grid, light, Rs[i], None, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- sharded == single GPU, word for word (before anything is timed): two sharded frames from a clean state against the same
    # two frames of an unsharded context on rank 0
    sharded_equals_single = None
    single_ctx = None
    if world > 1:
        for f, h in enumerate((0.0, 0.6)):
            step(ctx, f, h, True)
        got = ctx.probes_download()
        if rank == 0:
            single_ctx = make_ctx()
            for f, h in enumerate((0.0, 0.6)):
                step(single_ctx, f, h, False)
            want = single_ctx.probes_download()
            sharded_equals_single = bool(all(np.array_equal(a, b) for a, b in zip(got[:3], want[:3])))
        barrier()

    hyst = 0.0
    for w in range(Wm):
        step(ctx, w, hyst, world > 1); hyst = min(0.98, hyst + 0.25)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    launches0 = ctx.launch_count()
    if world == 1:
        # per-step events, L2 flushed between steps (untimed); no host synchronisation inside the loop
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        barrier()
        for s in range(K):
            with torch.cuda.stream(stream):
                flush.fill_(s & 0xFF)
            starts[s].record(stream)
            step(ctx, Wm + s, hyst, False)
            ends[s].record(stream)
        barrier()
        total_ms = float(sum(a.elapsed_time(b) for a, b in zip(starts, ends)))
        l2_note = "256 MiB buffer written between timed steps (flush, untimed)"
    else:
        # ONE interval over the K steps: the all-gather of step s runs on the communication stream while step s+1 traces, and the
        # interval ends only after the last gather has landed (vkx_stream_wait_exchange orders the stream behind it)
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_start.record(stream)
        for s in range(K):
            step(ctx, Wm + s, hyst, True)
        ctx.stream_wait_exchange()
        t_end.record(stream)
        barrier()
        total_ms = float(t_start.elapsed_time(t_end))
        scratch_mb = probes_per_rank * RAYS * (16 + 20 + 32 + 20) / 2**20
        l2_note = "no flush: steps run back to back in one timed interval so that the atlas all-gather is inside it; per-step streamed scratch (ray, hit, queue records: %.0f MiB per rank) exceeds the 126 MB L2" % scratch_mb
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / K
    value = rays_per_step_total / (ms_per_step * 1e-3)

    # ---- per-rank update time of the last timed step (the library's own events around one update): how uneven the ranks are
    rank_update_ms = None
    if world > 1:
        tm = ctx.probes_timings()
        in_loop = {"setup_ms": tm["trace"], "first_chunk_kernels_ms": tm["blend"], "after_first_chunk_ms": tm["publish"], "first_chunk": {k: v for k, v in ctx.probes_kernel_timings().items() if k in ("trace_primary", "shade", "trace_shadow", "blend")}}
        mine = torch.tensor([tm["full"]], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_update_ms = [round(float(t.item()), 4) for t in allr]

    # ---- per-kernel device times: a separate, untimed pass (vkx_probes_kernel_timings synchronises)
    kt = {"trace_primary": 0.0, "shade": 0.0, "trace_shadow": 0.0, "blend": 0.0}
    shadow_rays = 0
    for s in range(K):
        with torch.cuda.stream(stream):
            flush.fill_(s & 0xFF)
        step(ctx, Wm + K + s, hyst, world > 1)
        k = ctx.probes_kernel_timings()
        for nm in kt:
            kt[nm] += k[nm]
        shadow_rays = k["shadow_rays"]
    barrier()
    rank_kernel_sum_ms = None
    if world > 1:  # the same sum on every rank (first chunk of its update, run alone): how evenly the dealt slices load the ranks
        mine = torch.tensor([sum(kt.values()) / K], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_kernel_sum_ms = [round(float(t.item()), 4) for t in allr]

    # ---- back-to-back time at N = 1 too (no flush, one interval): what a renderer that updates every frame sees
    b2b_ms = None
    if world == 1:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for s in range(K):
            step(ctx, Wm + 2 * K + s, hyst, False)
        b.record(stream)
        barrier()
        b2b_ms = a.elapsed_time(b) / K

    # ---- end to end through the C ABI with host buffers: every step copies the to-update list + parameters host->device and reads
    # both atlases and the state words back into pinned host memory. The read-back of step s is queued asynchronously
    # (vkx_probes_download_slab_async) and overlaps the tracing of step s+1; all copies have landed before the clock stops.
    # With N ranks every rank reads back the z-slices it traced (vkx_shard_slices), so the job as a whole reads the volume back exactly once per step.
    (ih, iw), (dh, dw) = grid.atlas_shapes()
    ih, dh, nst = ih // n, dh // n, grid.probe_count // n
    from vulkanexp_b200._lib import shard_slices
    my_slices = shard_slices(grid.resolution[2], n, rank) if n > 1 else [(0, grid.resolution[2])]  # the z-slices this rank traced
    plane = grid.resolution[0] * grid.resolution[1]
    outs = []
    for _ in range(2):
        pin = (torch.empty((ih, iw), dtype=torch.int32).pin_memory(), torch.empty((dh, dw), dtype=torch.int32).pin_memory(), torch.empty(nst, dtype=torch.int32).pin_memory())
        outs.append((pin, tuple(t.numpy().view(np.uint32) for t in pin)))
    pin_irr, pin_dep, pin_st = outs[0][0]
    pin_idx = torch.arange(grid.probe_count, dtype=torch.int32).pin_memory()
    idx_np = pin_idx.numpy().view(np.uint32)
    e2e_steps = max(args.e2e_steps, K)
    base = Wm + 3 * K
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        grid.hysteresis = hyst
        if world > 1:
            ctx.probes_update_sharded(grid, light, Rs[base + (s % (args.e2e_steps + 8))], sync=False)
        else:
            ctx.probes_update(grid, light, Rs[base + (s % (args.e2e_steps + 8))], idx_np, sync=False)
        oi, od, ost = outs[s & 1][1]
        r0 = 0
        for (z0, z1) in my_slices:  # one read-back per slice group, packed one after the other in the pinned buffers
            nz = z1 - z0
            ctx.probes_download_slab_async(z0, z1, (oi[8 * r0:8 * (r0 + nz)], od[16 * r0:16 * (r0 + nz)], ost[r0 * plane:(r0 + nz) * plane]))
            r0 += nz
    ctx.probes_download_wait()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    checksum = int(outs[(e2e_steps - 1) & 1][1][0].astype(np.uint64).sum() % (1 << 32))  # the host really holds the last step's atlas
    h2d = ((grid.probe_count * 4 if world == 1 else 0) + 64 + 32 + 64 + RAYS * 16) * n
    d2h = (pin_irr.numel() * 4 + pin_dep.numel() * 4 + pin_st.numel() * 4) * n  # whole job

    # ---- the same workload on ONE GPU (rank 0), measured like the sharded run (one interval, no flush): the denominator of the
    # strong-scaling efficiency, in the same line
    single = None
    if world > 1 and scaling == "strong" and rank == 0 and single_ctx is not None:
        sstream = torch.cuda.ExternalStream(single_ctx.stream(), device=dev)
        for w in range(3):
            step(single_ctx, w, 0.5, False)
        single_ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = max(3, min(K, 10))
        a.record(sstream)
        for s in range(ks):
            step(single_ctx, Wm + s, hyst, False)
        b.record(sstream)
        single_ctx.sync(); torch.cuda.synchronize()
        sms = a.elapsed_time(b) / ks
        single = {"ms_per_step": sms, "value": rays_per_step_total / (sms * 1e-3), "steps": ks, "efficiency_vs_this": (rays_per_step_total / (ms_per_step * 1e-3)) / (n * rays_per_step_total / (sms * 1e-3))}
    if world > 1:
        barrier()
    cfg5 = None
    if world == 8 and scaling == "strong" and not args.no_secondary and os.environ.get("VKX_BENCH_CFG5", "1") != "0":
        cfg5 = cfg5_leg(local, rank, world, torch, dist, light)

    if rank == 0:
        peak, peak_src = peaks()
        line = {
            "metric": "ddgi_probe_rays_per_sec", "value": value, "unit": "probe rays/s", "n_gpus": n, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "probes_per_gpu": probes_per_rank, "rays_per_probe": RAYS, "l2": l2_note,
                       "timing": "per-step CUDA events on the library's stream, summed" if world == 1 else "one CUDA-event interval over all steps on the library's stream, closed after the last all-gather",
                       "parallelism": ("probe z-slices dealt round robin over %d ranks, %s" % (n, "blend fused with the atlas exchange over NVLink peer memory (P2P stores + device-side flags)" if exchange == "p2p" else "atlas rows pushed to every peer by copy engines over NVLink (IPC-mapped atlases, arrival flags), beside the next step's primary traversal" if exchange == "ce" else "NCCL all-gather of atlas slabs, deferred behind the next step's primary traversal")) if n > 1 else "single GPU"},
            "full_volume_update_ms": ms_per_step, "grays_per_sec_per_gpu": value / n / 1e9,
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": rays_per_step_total / (e2e_ms * 1e-3), "unit": "probe rays/s", "ms_per_step": e2e_ms, "steps": e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "vkx_probes_update from host buffers + asynchronous read-back of both atlases and the state words into pinned host memory every step (each rank its own z-slab; read-back of step s overlaps step s+1); byte counts are whole-job", "host_atlas_checksum": checksum},
            "kernel_ms": {k: v / K for k, v in kt.items()},
            **({"rank_update_ms": rank_update_ms, "rank_kernel_sum_ms": rank_kernel_sum_ms, "last_timed_step_on_rank0": in_loop} if rank_update_ms else {}),
        }
        if b2b_ms is not None:
            line["ms_per_step_back_to_back"] = b2b_ms
        if world > 1:
            line["sharded_equals_single"] = sharded_equals_single
            if single:
                line["single_gpu_same_workload"] = single
            if cfg5 is not None:
                line["secondary"] = {"cfg5": cfg5}
        if world == 1 and not args.no_secondary:
            try:
                line["secondary"] = secondary_shadow_pass(local)
            except Exception as e:  # the secondary measurement must never cost the headline
                line["secondary"] = {"error": repr(e)}
            try:  # BASELINE configs[3] on one GPU: the strong-scaling denominator next to the headline
                flat4 = scene_format.flatten(synth.make_cfg4())
                grid4 = GridInfo.make(flat4["bounds_min"], flat4["bounds_max"], CFG4_RES, RAYS, hysteresis=0.5)
                c4 = Context(local); c4.scene_upload(flat4); c4.bvh_build(); c4.probes_init(grid4)
                c4.probes_upload(state=np.ones(grid4.probe_count, dtype=np.uint32))
                s4 = torch.cuda.ExternalStream(c4.stream(), device=dev)
                for w in range(3):
                    c4.probes_update(grid4, light, Rs[w], None, sync=False)
                c4.sync()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s4)
                for s in range(5):
                    c4.probes_update(grid4, light, Rs[3 + s], None, sync=False)
                b.record(s4)
                c4.sync(); torch.cuda.synchronize()
                ms4 = a.elapsed_time(b) / 5
                line["secondary"]["cfg4_single_gpu"] = {"workload": workload(2, "strong")[0].split(", z-slices dealt")[0], "ms_per_step": ms4, "value": grid4.probe_count * RAYS / (ms4 * 1e-3), "unit": "probe rays/s", "steps": 5}
                c4.close()
            except Exception as e:
                line.setdefault("secondary", {})["cfg4_single_gpu"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            cb, ctr, sample_rays = cpu_baseline(flat, grid, light, min(args.cpu_sample, grid.probe_count))
            line["cpu_baseline"] = cb
            # Algorithmic bytes per unit, SURVEY 8(d):
            #   trace:  mean 80-byte nodes + 48-byte triangles fetched per ray (the oracle's traversal counters on the sample; identical
            #           traversal order on the device) + the ray's output record
            #   shade:  hit record + ray record + per front hit: offsets / indices / 3 vertices / material + 16 probe taps x 8 texels x 4 B
            #           (SURVEY's f_front term: 12 + 12 + 192 + 48 + 512 = 776 B) + the shadow-queue entry
            #   blend:  256 ray records + both tiles read and written + borders (6304 B per probe at 256 rays)
            nodes_p, tris_p = ctr["nodes"] / ctr["rays"], ctr["tris"] / ctr["rays"]
            nodes_s, tris_s = ctr["shadow_nodes"] / max(1, ctr["shadow_rays"]), ctr["shadow_tris"] / max(1, ctr["shadow_rays"])
            front = ctr["front"] / sample_rays
            per_unit = {
                "trace_primary": nodes_p * NODE_BYTES + tris_p * TRI_BYTES + HIT_BYTES + 4,
                "trace_shadow": nodes_s * NODE_BYTES + tris_s * TRI_BYTES + 32 + 16,
                "shade": HIT_BYTES + 16 + front * (776 + 32),
                "blend": (RAYS * 16 + (36 + 196) * 4 * 2 + (28 + 60) * 4) / RAYS,
            }
            dom = max(kt, key=kt.get)
            units = shadow_rays if dom == "trace_shadow" else probes_per_rank * RAYS
            achieved = per_unit[dom] * units / (kt[dom] / K * 1e-3) / 1e9
            prof = {}
            try:  # per-kernel ncu figures of the committed capture: DRAM bytes per launch, issue-slot and lane utilisation
                prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            except Exception:
                pass
            # whole step against SURVEY's B_ray (primary + shadow traversal + shading term), and the compulsory HBM floor
            b_ray = nodes_p * NODE_BYTES + tris_p * TRI_BYTES + front * (nodes_s * NODE_BYTES + tris_s * TRI_BYTES + 776) + 16
            step_gbs = b_ray * probes_per_rank * RAYS / (ms_per_step * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": "k_" + dom + (" (k_shade_miss + k_shade_front)" if dom == "shade" else ""), "achieved": achieved, "peak": peak, "unit": "GB/s",
                                "frac": achieved / peak, "traffic": prof.get("k_" + dom), "peak_source": peak_src, "bytes_per_unit": per_unit[dom], "units_per_launch": int(units),
                                "nodes_per_primary_ray": nodes_p, "tris_per_primary_ray": tris_p, "nodes_per_shadow_ray": nodes_s, "tris_per_shadow_ray": tris_s,
                                "front_hit_fraction": front,
                                "whole_step": {"bytes_per_primary_ray": b_ray, "achieved": step_gbs, "frac": step_gbs / peak, "what": "SURVEY 8(d) B_ray (node + triangle fetches of the primary and the shadow ray, vertices / material / 16 probe taps per front hit, ray record) x rays / update time"},
                                "compulsory_hbm_floor_ms": (2208.0 * probes_per_rank + 36e6) / (peak * 1e9) * 1e3,
                                "operative_bound": dict({"kind": "SM issue slots and warp lane utilisation: the algorithmic bytes above are L1/L2 traffic (the BVH and the atlases live in the 126 MB L2; ncu DRAM throughput 1-6 % of peak), see profiles/"}, **({"kernels": {k: prof.get("_issue", {}).get(k) for k in (["k_shade_front", "k_shade_miss"] if dom == "shade" else ["k_blend_tc"] if dom == "blend" else ["k_" + dom]) if prof.get("_issue", {}).get(k)}, "rays_per_launch": int(units)}))}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
