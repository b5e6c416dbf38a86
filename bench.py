#!/usr/bin/env python
"""Benchmark of the DDGI probe update (BASELINE.json metric: probe rays/s and full-volume update ms).

  python bench.py --gpus N --steps K --warmup W            the CUDA path (libvkexp_b200.so through its C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the CPU transliteration (oracle/) on the host cores

A step is one full-volume DDGI update: every probe of the volume traces raysPerProbe rays (closest hit, shading with
sun shadow ray and two sampleProbes look-ups, or sky on a miss), blends them into both atlases with hysteresis, writes
the border texels and publishes. N = 1 runs BASELINE.json configs[1]: 32x16x32 probes x 256 rays on the synthetic
Sponza-scale scene (the named data/sponza_test.scene is not in the reference checkout). N > 1 is weak scaling: the
volume becomes 32x16x(32 N) probes over the same scene, every rank traces and blends 16384 probes (z-slices
interleaved per chunk) and all-gathers its atlas slices over NCCL.
"""
import argparse
import ctypes as C
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

RES = (32, 16, 32)
RAYS = 256
NODE_BYTES, TRI_BYTES, HIT_BYTES = 80, 48, 20


def workload_name(n):
    return "DDGI full-volume update, synthetic sponza-scale atrium (265k triangles), %dx%dx%d probes x %d rays" % (RES[0], RES[1], RES[2] * n, RAYS)


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
    """The reference's per-frame random rotations: MSVC-LCG replay of glm::sphericalRand + genBasis. Host-side harness
    input (the C ABI takes the matrix); computed with the C++ facade's generator so the product path stays oracle-free."""
    from vulkanexp_b200.host_logic import OrientationGenerator

    gen = OrientationGenerator()
    gen.next()  # the first draw is consumed by initProbes (reference src/IrradianceProbes.cpp:361)
    return [gen.next() for _ in range(k)]


def cpu_baseline(flat, grid, light, sample_probes, threads=0):
    """Times the oracle (CPU transliteration) on a bounded sample of the same workload; also returns its traversal counters."""
    from oracle import pyoracle

    o = pyoracle.Oracle()
    o.scene_upload(flat)
    o.bvh_build()
    o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    idx = np.linspace(0, grid.probe_count - 1, sample_probes).astype(np.uint32)
    reps = 4
    R = orientations(1 + reps)
    o.probes_update(grid, light, R[0], idx, threads)  # warm-up: fills the atlases so sampleProbes does real work
    sec = sum(o.probes_update(grid, light, R[1 + k], idx, threads) for k in range(reps)) / reps
    c = o.probes_counters()
    rays = len(idx) * grid.raysPerProbe
    return {
        "value": rays / sec,
        "unit": "probe rays/s",
        "cores": pyoracle.lib().orc_max_threads() if threads <= 0 else threads,
        "kind": "port",
        "sample": "%d of %d probes (evenly spaced) x %d rays, mean of %d updates after one warm-up update, %.2f s per update" % (len(idx), grid.probe_count, grid.raysPerProbe, reps, sec),
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
    if rank != 0:
        return
    flat = scene_format.flatten(synth.make_cfg2())
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (RES[0], RES[1], RES[2] * args.gpus), RAYS, hysteresis=0.9)
    light = Light.default()
    from oracle import pyoracle

    o = pyoracle.Oracle()
    o.scene_upload(flat)
    o.bvh_build()
    o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    sample = min(grid.probe_count, args.ref_sample)
    idx = np.linspace(0, grid.probe_count - 1, sample).astype(np.uint32)
    Rs = orientations(args.warmup + args.steps)
    for w in range(args.warmup):
        o.probes_update(grid, light, Rs[w], idx, 0)
    total = 0.0
    for s in range(args.steps):
        total += o.probes_update(grid, light, Rs[args.warmup + s], idx, 0)
    rays = sample * RAYS
    value = rays * args.steps / total
    cores = pyoracle.lib().orc_max_threads()
    line = {
        "impl": "reference", "metric": "ddgi_probe_rays_per_sec", "value": value, "unit": "probe rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample": "%d of %d probes per step" % (sample, grid.probe_count)},
        "cpu_baseline": {"value": value, "unit": "probe rays/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d probes (evenly spaced) x %d rays per step; the reference itself (Vulkan RT shaders, Windows) cannot run here, so this is its CPU transliteration (oracle/)" % (sample, grid.probe_count, RAYS)},
        "e2e": {"value": value, "unit": "probe rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "full_volume_update_ms_extrapolated": 1e3 * total / args.steps * grid.probe_count / sample,
    }
    _emit(line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vkx", choices=["vkx", "reference"])
    ap.add_argument("--ref-sample", type=int, default=4096, help="probes per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=16384, help="probes of the cpu_baseline leg (default: the whole 32x16x32 volume)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    flat = scene_format.flatten(synth.make_cfg2())
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (RES[0], RES[1], RES[2] * n), RAYS, hysteresis=0.0)
    light = Light.default()
    ctx = Context(local)
    ctx.scene_upload(flat)
    ctx.bvh_build()
    ctx.probes_init(grid)
    ctx.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))  # every probe active: the named workload traces all of them
    if world > 1:
        uid = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
        exchange = os.environ.get("VKX_EXCHANGE", "nccl")
        if exchange == "p2p":  # VKX_EXCHANGE=p2p: blend fused with the atlas exchange over NVLink peer memory (measured slower than the deferred all-gather at 8 GPUs)
            exchange = "p2p" if ctx.comm_p2p_enable(dist) else "nccl"
    else:
        exchange = "none"
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    Rs = orientations(args.warmup + args.steps + args.steps)
    probes_per_rank = grid.probe_count // n
    rays_per_step_total = grid.probe_count * RAYS

    def step(i, hyst):
        grid.hysteresis = hyst
        if world > 1:
            ctx.probes_update_sharded(grid, light, Rs[i], sync=False)
        else:
            ctx.probes_update(grid, light, Rs[i], None, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    hyst = 0.0
    for w in range(args.warmup):
        step(w, hyst); hyst = min(0.98, hyst + 0.25)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    launches0 = ctx.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kt = {"trace_primary": 0.0, "shade": 0.0, "trace_shadow": 0.0, "blend": 0.0}
    shadow_rays = 0
    barrier()
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(s & 0xFF)  # evict L2 between timed iterations (not timed)
        starts[s].record(stream)
        step(args.warmup + s, hyst)
        ends[s].record(stream)
        k = ctx.probes_kernel_timings()  # syncs; per-kernel CUDA events recorded inside the library on its launch stream
        for name in kt:
            kt[name] += k[name]
        shadow_rays = k["shadow_rays"]
    barrier()
    launches = ctx.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
    total_ms = float(sum(step_ms))
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = rays_per_step_total / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers: every step copies the to-update list + parameters host->device and reads
    # both atlases and the state words back into pinned host memory. The read-back of step s is queued asynchronously
    # (vkx_probes_download_async) and overlaps the tracing of step s+1; all copies have landed before the clock stops.
    # With N ranks every rank reads back the z-slab it traced, so the job as a whole reads the volume back exactly once per step.
    (ih, iw), (dh, dw) = grid.atlas_shapes()
    ih, dh, nst = ih // n, dh // n, grid.probe_count // n
    z0, z1 = rank * (grid.resolution[2] // n), (rank + 1) * (grid.resolution[2] // n)
    outs = []
    for _ in range(2):
        pin = (torch.empty((ih, iw), dtype=torch.int32).pin_memory(), torch.empty((dh, dw), dtype=torch.int32).pin_memory(), torch.empty(nst, dtype=torch.int32).pin_memory())
        outs.append((pin, tuple(t.numpy().view(np.uint32) for t in pin)))
    pin_irr, pin_dep, pin_st = outs[0][0]
    pin_idx = torch.arange(grid.probe_count, dtype=torch.int32).pin_memory()
    idx_np = pin_idx.numpy().view(np.uint32)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        grid.hysteresis = hyst
        if world > 1:
            ctx.probes_update_sharded(grid, light, Rs[args.warmup + args.steps + s], sync=False)
        else:
            ctx.probes_update(grid, light, Rs[args.warmup + args.steps + s], idx_np, sync=False)
        ctx.probes_download_slab_async(z0, z1, outs[s & 1][1])
    ctx.probes_download_wait()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    checksum = int(outs[(args.steps - 1) & 1][1][0].astype(np.uint64).sum() % (1 << 32))  # the host really holds the last step's atlas
    h2d = (grid.probe_count * 4 if world == 1 else 0) + 64 + 32 + 64 + RAYS * 16
    d2h = (pin_irr.numel() * 4 + pin_dep.numel() * 4 + pin_st.numel() * 4) * n  # whole job
    h2d *= n

    if rank == 0:
        peak, peak_src = peaks()
        line = {
            "metric": "ddgi_probe_rays_per_sec", "value": value, "unit": "probe rays/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n), "probes_per_gpu": probes_per_rank, "rays_per_probe": RAYS, "l2": "256 MiB buffer written between timed steps (flush, untimed)",
                       "parallelism": ("probe z-slabs x%d, %s" % (n, "blend fused with the atlas exchange over NVLink peer memory (P2P stores + device-side flags)" if exchange == "p2p" else "NCCL all-gather of atlas slabs")) if n > 1 else "single GPU"},
            "full_volume_update_ms": ms_per_step, "grays_per_sec_per_gpu": value / n / 1e9,
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": rays_per_step_total / (e2e_ms * 1e-3), "unit": "probe rays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "vkx_probes_update from host buffers + asynchronous read-back of both atlases and the state words into pinned host memory every step (each rank its own z-slab; read-back of step s overlaps step s+1); byte counts are whole-job", "host_atlas_checksum": checksum},
            "kernel_ms": {k: v / args.steps for k, v in kt.items()},
        }
        if not args.no_cpu_baseline:
            cb, ctr, sample_rays = cpu_baseline(flat, grid, light, min(args.cpu_sample, grid.probe_count))
            line["cpu_baseline"] = cb
            # roofline of the kernel with the largest measured time. Algorithmic bytes per unit:
            #   trace kernels: mean 80-byte nodes + 48-byte triangles fetched per ray (the oracle's traversal counters on the
            #                  sample; identical traversal order on the device) + the ray's output record
            #   shade:         hit record + ray record + per front hit: offsets/indices/3 normals/material/inverse matrix + 16 probe
            #                  taps x 8 texels x 4 B + shadow-queue entry
            #   blend:         256 ray records + both tiles read and written + borders (6304 B per probe at 256 rays)
            nodes_p, tris_p = ctr["nodes"] / ctr["rays"], ctr["tris"] / ctr["rays"]
            nodes_s, tris_s = ctr["shadow_nodes"] / max(1, ctr["shadow_rays"]), ctr["shadow_tris"] / max(1, ctr["shadow_rays"])
            front = ctr["front"] / sample_rays
            per_unit = {
                "trace_primary": nodes_p * NODE_BYTES + tris_p * TRI_BYTES + HIT_BYTES + 4,
                "trace_shadow": nodes_s * NODE_BYTES + tris_s * TRI_BYTES + 32 + 16,
                "shade": HIT_BYTES + 16 + front * (12 + 12 + 36 + 48 + 36 + 16 * 8 * 4 + 32),
                "blend": (RAYS * 16 + (36 + 196) * 4 * 2 + (28 + 60) * 4) / RAYS,
            }
            dom = max(kt, key=kt.get)
            units = shadow_rays if dom == "trace_shadow" else probes_per_rank * RAYS
            achieved = per_unit[dom] * units / (kt[dom] / args.steps * 1e-3) / 1e9
            traffic = None
            try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture of this kernel
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_" + dom)
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "kernel": "k_" + dom + (" (k_shade_miss + k_shade_front)" if dom == "shade" else ""), "achieved": achieved, "peak": peak, "unit": "GB/s",
                                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "bytes_per_unit": per_unit[dom], "units_per_launch": int(units),
                                "nodes_per_primary_ray": nodes_p, "tris_per_primary_ray": tris_p, "nodes_per_shadow_ray": nodes_s, "tris_per_shadow_ray": tris_s,
                                "front_hit_fraction": front,
                                "note": "every kernel of this path is instruction-issue bound at this scene size (DRAM < 6 % of peak in ncu); the HBM fraction is reported as required, issue-slot utilisation is in profiles/"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
