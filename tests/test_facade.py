"""The C++ facade (csrc/host: Scene, Renderer, IrradianceProbes with the reference's method names) must produce exactly
what the C ABI produces when driven by the same host logic."""
import os
import subprocess

import numpy as np
import pytest

from vulkanexp_b200 import scene_format, synth
from vulkanexp_b200.pods import GridInfo, Light

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _demo():
    """Path of the facade demo executable (a build artefact, not tracked): built on demand if build() has not run in this tree."""
    path = os.path.join(ROOT, "vulkanexp_b200", "vkx_facade_demo")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "vulkanexp_b200", "csrc")])
    return path


@pytest.mark.parametrize("textured,skinned", [(False, False), (True, False), (False, True)])
def test_facade_equals_c_abi(tmp_path, textured, skinned):
    from vulkanexp_b200._lib import Context
    from vulkanexp_b200.host_logic import OrientationGenerator, ProbeScheduler

    s = synth.make_textured_court() if textured else synth.make_open_court()
    if textured:  # two of the images go through the facade's PNG decoder, the others through its Netpbm reader
        pytest.importorskip("PIL")
        s.textures[0]["source"] = "tex_albedo.png"
        s.textures[4]["source"] = "tex_grate.png"
    path = os.path.join(tmp_path, "court.scene")
    scene_format.write_scene(path, s)
    res, rays, frames = (7, 5, 6), 48, 14
    out = os.path.join(tmp_path, "facade.bin")
    skin_mesh = 2  # the ball mesh, as a skinned renderer on the root node (facade: Renderer::updateSkinnedVertexBuffer + updateSkinnedBLAS per frame)
    extra = ["0", "host", str(skin_mesh)] if skinned else []
    r = subprocess.run([_demo(), path, str(res[0]), str(res[1]), str(res[2]), str(rays), str(frames), out] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "facade ok" in r.stdout
    # the same sequence through the C ABI, flattening done by the Python harness
    flat = scene_format.flatten(scene_format.read_scene(path))
    if skinned:
        flat, src, dst, size = scene_format.add_skinned_instance(flat, skin_mesh)
        sj = ((np.arange(size)[:, None] + np.arange(4)[None, :]) % 3).astype(np.uint16)
        sw = np.tile(np.array([[0.4, 0.3, 0.2, 0.1]], dtype=np.float32), (size, 1))

        def pose(frame):
            js = np.zeros((3, 4, 4), dtype=np.float32)  # [joint][col][row]
            for j in range(3):
                sc = np.float32(1.0) + np.float32(j) / np.float32(8.0)
                js[j, 0, 0] = js[j, 1, 1] = js[j, 2, 2] = sc
                js[j, 3, 3] = 1.0
                js[j, 3, 0] = np.float32(frame) * np.float32(j + 1) / np.float32(32.0)
                js[j, 3, 1] = np.float32(frame) / np.float32(64.0)
            return js.reshape(3, 16)
    g = Context(0); g.scene_upload(flat); g.bvh_build()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, rays)
    g.probes_init(grid)
    gen, sched = OrientationGenerator(), ProbeScheduler()
    g.probes_classify(gen.next())
    light = Light.default()
    target, updated, device_h = np.float32(0.98), 0, np.float32(0.0)
    h = np.float32(0.0)
    for f in range(frames):
        if skinned:
            g.skin_vertices(pose(f), sj, sw, src, dst); g.bvh_build()
        st = g.probes_download()[2]
        idx = sched.select(st)
        R = gen.next()
        if updated >= grid.probe_count:  # hysteresis ramp, reference src/IrradianceProbes.cpp:462-476
            if abs(h - target) > 0.05:
                device_h = h
                h = np.float32(h + np.float32(0.1) * np.float32(target - h))
            elif h != target:
                h = target
                device_h = h
            updated -= grid.probe_count
        updated += len(idx)
        if len(idx):
            grid.hysteresis = float(device_h)
            g.probes_update(grid, light, R, idx)
    irr, dep, st, _ = g.probes_download()
    raw = np.fromfile(out, dtype=np.uint32)
    a, b = irr.size, irr.size + dep.size
    assert np.array_equal(raw[:a], irr.reshape(-1)), "irradiance atlas differs between facade and C ABI"
    assert np.array_equal(raw[a:b], dep.reshape(-1)), "depth atlas differs"
    assert np.array_equal(raw[b:], st), "probe states differ"


def test_facade_device_scheduler_equals_host_scheduler(tmp_path):
    """IrradianceProbes::DeviceScheduler (vkx_probes_schedule) must not change a single bit of the result."""
    s = synth.make_open_court()
    path = os.path.join(tmp_path, "court.scene")
    scene_format.write_scene(path, s)
    outs = []
    for mode in ("host", "device"):
        out = os.path.join(tmp_path, mode + ".bin")
        r = subprocess.run([_demo(), path, "7", "5", "6", "48", "16", out, "60", mode], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(np.fromfile(out, dtype=np.uint32))
    assert outs[0].size and np.array_equal(outs[0], outs[1])
