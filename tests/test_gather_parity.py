"""Final composite (FinalGather.frag): CUDA path vs the CPU oracle on identical G-buffer, shadow image and atlases."""
import numpy as np
import pytest

from conftest import make_pair, rel_err
from vulkanexp_b200 import synth
from vulkanexp_b200.pods import GridInfo, Light, make_camera

pytestmark = pytest.mark.gpu

W, H = 320, 180
TOL = 1e-3


def _prepare(oracle_lib, scene, res):
    o, g, flat = make_pair(oracle_lib, scene)
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], res, 64, hysteresis=0.5)
    o.probes_init(grid); g.probes_init(grid)
    host = oracle_lib.HostLogic()
    R, _ = host.next_orientation()
    g.probes_classify(R)
    light = Light.default()
    for _ in range(4):  # a lit irradiance volume
        R, _ = host.next_orientation()
        g.probes_update(grid, light, R)
    irr, dep, st, _ = g.probes_download()
    o.probes_upload(irr, dep, st)  # both sides sample the same atlases
    noise = synth.blue_noise_like(8, 64)
    for c in (o, g):
        c.shadow_set_noise(noise); c.shadow_init(W, H)
    return o, g, flat, grid, light


def _frame(o, g, cam, prev, light):
    g.gbuffer_generate(cam)
    pd, nm = g.gbuffer_download()
    ar, em = g.gbuffer_download_material()
    o.gbuffer_upload(pd, nm); o.gbuffer_upload_material(ar, em)
    g.shadow_frame(cam, prev, light)
    o.shadow_set_history(g.shadow_download(2))  # identical direct-light input
    return pd, ar, em


def test_gbuffer_material_targets_match_oracle(oracle_lib):
    o, g, flat = make_pair(oracle_lib, "cfg1")
    o.shadow_init(W, H); g.shadow_init(W, H)
    cam = make_camera((-3.0, 2.0, 3.5), (0.0, 1.0, 0.0), aspect=W / H, frame_index=0)
    o.gbuffer_generate(cam); g.gbuffer_generate(cam)
    ao, eo = o.gbuffer_download_material()
    ag, eg = g.gbuffer_download_material()
    assert (ao[..., :3].sum(axis=-1) > 0).mean() > 0.3
    assert rel_err(ao, ag).max() < 1e-5
    assert np.array_equal(eo, eg)


@pytest.mark.parametrize("scene,res,eye,target", [("court", (8, 6, 8), (-5.0, 2.5, 4.5), (0.0, 6.0, 0.0)), ("cfg1", (8, 8, 8), (-3.0, 2.0, 3.5), (0.5, 1.5, 0.0))])
def test_final_gather_matches_oracle(oracle_lib, scene, res, eye, target):
    o, g, flat, grid, light = _prepare(oracle_lib, scene, res)
    cams = [make_camera((eye[0] + 0.3 * f, eye[1], eye[2] - 0.2 * f), target, aspect=W / H, frame_index=f) for f in range(3)]
    prev = cams[0]
    rng = np.random.default_rng(7)
    for f, cam in enumerate(cams):
        pd, ar, em = _frame(o, g, cam, prev, light)
        refl = rng.random((H, W, 4), dtype=np.float32) if f == 2 else None
        g.final_gather(cam, light, refl)
        img_g, ms = g.final_gather_download()
        img_o, _ = o.final_gather(cam, light, refl)
        geo = pd[..., 3] > 0
        assert np.isfinite(img_g).all()
        assert np.array_equal(img_g[..., 3], np.ones((H, W), dtype=np.float32))
        e = rel_err(img_o[..., :3], img_g[..., :3])
        print("frame %d: geometry %.2f of pixels, max rel err geometry %.2e, sky %.2e, kernel %.3f ms" %
              (f, geo.mean(), e[geo].max() if geo.any() else 0.0, e[~geo].max() if (~geo).any() else 0.0, ms))
        assert e.max() < TOL
        assert img_g[geo][:, :3].max() > 0.05, "composite of lit geometry must not be black"
        prev = cam


def test_final_gather_requires_setup(oracle_lib):
    from vulkanexp_b200._lib import Context, VkxError
    g = Context(0)
    cam = make_camera((0, 1, 5), (0, 1, 0), aspect=1.0, frame_index=0)
    with pytest.raises(VkxError):
        g.final_gather(cam, Light.default())
