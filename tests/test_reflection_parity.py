"""Reflection pass (reflection.rgen + reflectionFilterX/Y): CUDA path vs the CPU oracle."""
import numpy as np
import pytest

from conftest import make_pair, rel_err
from vulkanexp_b200 import synth
from vulkanexp_b200.pods import GridInfo, Light, make_camera

pytestmark = pytest.mark.gpu

W, H = 256, 144
TOL = 1e-3


def _prepare(oracle_lib, scene, res):
    o, g, flat = make_pair(oracle_lib, scene)
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, 64, hysteresis=0.5)
    g.probes_debug(True)  # hit / mask side buffers
    o.probes_init(grid); g.probes_init(grid)
    host = oracle_lib.HostLogic()
    R, _ = host.next_orientation()
    g.probes_classify(R)
    light = Light.default()
    for _ in range(4):
        R, _ = host.next_orientation()
        g.probes_update(grid, light, R)
    irr, dep, st, _ = g.probes_download()
    o.probes_upload(irr, dep, st)
    noise = synth.blue_noise_like(8, 64)
    for c in (o, g):
        c.shadow_set_noise(noise); c.shadow_init(W, H)
    return o, g, light


def _gbuffer(o, g, cam, rng):
    """GPU G-buffer on both sides; roughness / metalness are varied per 8x8 tile so that most pixels reflect with a wide
    range of lobe widths (the synthetic scenes only have a few shiny objects)."""
    g.gbuffer_generate(cam)
    pd, nm = g.gbuffer_download()
    ar, em = g.gbuffer_download_material()
    ty, tx = np.arange(H)[:, None] // 8, np.arange(W)[None, :] // 8
    tile = rng.random((H // 8 + 1, W // 8 + 1), dtype=np.float32)
    geo = pd[..., 3] > 0
    ar[..., 3] = np.where(geo, np.where(tile[ty, tx] < 0.7, tile[ty, tx] * 0.5, ar[..., 3]), 0.0).astype(np.float32)
    nm[..., 3] = np.where(geo & (tile[ty, tx] > 0.9), 1.0, nm[..., 3]).astype(np.float32)
    for c in (o, g):
        c.gbuffer_upload(pd, nm); c.gbuffer_upload_material(ar, em)
    return pd, nm, ar


@pytest.mark.parametrize("scene,res,eye,target", [("court", (8, 6, 8), (-5.0, 2.5, 4.5), (0.0, 3.0, 0.0)), ("cfg1", (8, 8, 8), (-3.0, 2.0, 3.5), (0.5, 1.5, 0.0)),
                                                    ("tcourt", (8, 6, 8), (-5.0, 2.5, 4.5), (0.0, 3.0, 0.0))])
def test_reflection_frames_match_oracle(oracle_lib, scene, res, eye, target):
    o, g, light = _prepare(oracle_lib, scene, res)
    rng = np.random.default_rng(11)
    # frames 0-1: camera at rest (history fully used), then a small move (partial history), then a jump > 1 m (history dropped)
    eyes = [eye, eye, (eye[0] + 0.05, eye[1], eye[2] - 0.03), (eye[0] + 1.5, eye[1] + 0.2, eye[2] - 0.8)]
    cams = [make_camera(e, target, aspect=W / H, frame_index=f + 62) for f, e in enumerate(eyes)]  # frame 64: the noise offset becomes (1, 0)
    prev = cams[0]
    for f, cam in enumerate(cams):
        pd, nm, ar = _gbuffer(o, g, cam, rng)
        g.reflection_frame(cam, prev, light)
        dirs_g, hits_g, mask_g = g.reflection_download_debug()
        # (1) the oracle's own jittered directions agree with the GPU's to fp32 rounding
        o.reflection_frame(cam, prev, light)
        _, dirs_o, _, mask_free = o.reflection_download(0)
        assert (mask_g > 0).mean() > 0.3, "most pixels should trace a reflection ray"
        assert np.array_equal(mask_free > 0, mask_g > 0), "which pixels reflect is decided by the G-buffer alone"
        assert np.abs(dirs_o - dirs_g).max() < 5e-6
        assert (mask_free != mask_g).mean() < 2e-3, "unconstrained classifications may only differ on grazing rays"
        # (2) same directions on both sides: hit records and classification bit-exact, images within tolerance
        o.reflection_set_history(g_prev if f else np.zeros((H, W, 4), dtype=np.float32))
        o.reflection_frame(cam, prev, light, dir_override=dirs_g)
        raw_o, _, hits_o, mask_o = o.reflection_download(0)
        assert np.array_equal(mask_o, mask_g), "frame %d: miss / back / front / shadow classification differs" % f
        assert hits_o.tobytes() == hits_g.tobytes(), "frame %d: reflection hit records differ" % f
        raw_g = g.reflection_download(0)
        e0 = rel_err(raw_o, raw_g).max()
        e1 = rel_err(o.reflection_download(1)[0], g.reflection_download(1)).max()
        fin_g = g.reflection_download(2)
        e2 = rel_err(o.reflection_download(2)[0], fin_g).max()
        print("frame %d: %.2f of pixels reflect (miss %.2f, back %.3f, lit %.2f, shadowed %.2f); max rel err raw %.2e, X %.2e, final %.2e"
              % (f, (mask_g > 0).mean(), (mask_g == 1).mean(), (mask_g == 2).mean(), (mask_g == 3).mean(), (mask_g == 4).mean(), e0, e1, e2))
        assert e0 < TOL and e1 < TOL and e2 < TOL
        assert np.array_equal(raw_g[..., 3], ar[..., 3] * (mask_g > 0)), "raw alpha carries the roughness"
        g_prev = fin_g
        prev = cam
    t = g.reflection_timings()
    assert t["full"] > 0


def test_composite_uses_device_reflection(oracle_lib):
    o, g, light = _prepare(oracle_lib, "court", (8, 6, 8))
    rng = np.random.default_rng(5)
    cam = make_camera((-5.0, 2.5, 4.5), (0.0, 3.0, 0.0), aspect=W / H, frame_index=3)
    _gbuffer(o, g, cam, rng)
    g.shadow_frame(cam, cam, light)
    g.final_gather(cam, light)
    without, _ = g.final_gather_download()
    g.reflection_frame(cam, cam, light)
    refl = g.reflection_download(2)
    g.final_gather(cam, light)            # NULL reflection -> the device-resident result of vkx_reflection_frame
    with_dev, _ = g.final_gather_download()
    g.final_gather(cam, light, refl)      # the same image handed over from the host
    with_host, _ = g.final_gather_download()
    assert np.array_equal(with_dev, with_host)
    assert (with_dev[..., :3] >= without[..., :3]).all() and (with_dev[..., :3] > without[..., :3]).mean() > 0.2
    o.shadow_set_history(g.shadow_download(2))
    img_o, _ = o.final_gather(cam, light, refl)
    assert rel_err(img_o[..., :3], with_dev[..., :3]).max() < TOL
