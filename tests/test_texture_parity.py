"""Texture path ("sampler spec v1"): CUDA (through the C ABI) vs the CPU oracle — generated mip chains, look-ups with and without
gradients, the any-hit cut-out inside traversal, textured closest-hit shading of the probe update and the sun-shadow pass through
a cut-out canopy. The textured scene also runs through test_ddgi_parity / test_reflection_parity ("tcourt")."""
import numpy as np
import pytest

from conftest import get_scene, make_pair, rel_err
from vulkanexp_b200 import synth
from vulkanexp_b200._lib import Context, VkxError
from vulkanexp_b200.pods import GridInfo, Light, make_camera

pytestmark = pytest.mark.gpu


def _random_textures(rng):
    mk = lambda h, w: rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    return [
        {"pixels": mk(64, 64), "srgb": 1},
        {"pixels": mk(12, 20), "srgb": 0, "wrapS": 33071, "wrapT": 33648},
        {"pixels": mk(37, 5), "srgb": 1, "minFilter": 9986, "wrapS": 33648, "wrapT": 33071},
        {"pixels": mk(1, 1), "srgb": 0},
        {"pixels": mk(16, 16), "srgb": 0, "magFilter": 9728, "minFilter": 9984},
        {"pixels": mk(3, 129), "srgb": 1, "minFilter": 9985},
    ]


def test_mip_chains_are_bit_identical(oracle_lib):
    rng = np.random.default_rng(11)
    tex = _random_textures(rng) + get_scene("tcourt")["textures"]
    o, g = oracle_lib.Oracle(), Context(0)
    o.scene_textures(tex); g.scene_textures(tex)
    for i in range(len(tex)):
        mo, mg = o.texture_download(i), g.texture_download(i)
        assert len(mo) == len(mg)
        for l, (a, b) in enumerate(zip(mo, mg)):
            assert a.tobytes() == b.tobytes(), "texture %d level %d differs" % (i, l)


def test_lookups_match_oracle(oracle_lib):
    rng = np.random.default_rng(12)
    tex = _random_textures(rng)
    o, g = oracle_lib.Oracle(), Context(0)
    o.scene_textures(tex); g.scene_textures(tex)
    n = 4000
    uv = rng.uniform(-3, 4, (n, 2)).astype(np.float32)
    uv[:6] = [[0, 0], [1, 1], [0.5, 0.5], [-1, 2], [1e-8, 1 - 1e-8], [123.25, -77.75]]
    grads = (rng.normal(size=(n, 4)) * np.exp(rng.uniform(-9, 1, (n, 1)))).astype(np.float32)
    grads[:40] = 0.0
    grads[40:60, 0] = np.nan
    for i in range(len(tex)):
        assert np.array_equal(o.texture_sample(i, uv), g.texture_sample(i, uv)), "texture %d: base-level look-ups must be bit-exact (they decide the cut-outs)" % i
        a, b = o.texture_sample(i, uv, grads), g.texture_sample(i, uv, grads)
        assert np.isfinite(b).all()
        # log2 differs by an ulp between glibc and CUDA: the trilinear weight moves by ~1e-6
        assert np.abs(a - b).max() < 2e-5, "texture %d: textureGrad max abs err %g" % (i, np.abs(a - b).max())


def test_cut_outs_in_traversal_bit_exact(oracle_lib):
    o, g, flat = make_pair(oracle_lib, "tcourt")
    rng = np.random.default_rng(13)
    n = 60000
    lo, hi = flat["bounds_min"], flat["bounds_max"]
    origins = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    origins[: n // 2, 1] = rng.uniform(5.0, 6.9, n // 2)  # many rays start near the canopy
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for any_hit in (False, True):
        ho = o.trace(origins, d, 0.01, 100.0, any_hit=any_hit, alpha_test=True)
        hg = g.trace(origins, d, 0.01, 100.0, any_hit=any_hit, alpha_test=True)
        assert ho.tobytes() == hg.tobytes(), "any_hit=%s: hit records differ with the cut-out test on" % any_hit
        plain = g.trace(origins, d, 0.01, 100.0, any_hit=any_hit)
        assert plain.tobytes() == o.trace(origins, d, 0.01, 100.0, any_hit=any_hit).tobytes()
        changed = (plain["t"] != hg["t"]).mean()
        assert changed > 1e-3, "the cut-outs must change some rays (%.2e)" % changed


def test_texture_errors():
    g = Context(0)
    flat = get_scene("tcourt")
    with pytest.raises(VkxError):  # materials use textures that were not provided
        g.scene_upload({k: v for k, v in flat.items() if k != "textures"})
    g.scene_upload(flat); g.bvh_build()
    ray = (np.array([[0.0, 3.0, 0.0]], dtype=np.float32), np.array([[0.0, -1.0, 0.0]], dtype=np.float32))
    assert g.trace(*ray, 0.01, 100.0)["t"][0] > 0
    with pytest.raises(VkxError):
        g.texture_sample(99, np.zeros((1, 2), dtype=np.float32))
    g.scene_textures([])  # dropping textures the uploaded materials use discards that scene
    with pytest.raises(VkxError):
        g.trace(*ray, 0.01, 100.0)
    g.scene_upload(get_scene("court")); g.bvh_build()
    assert g.trace(*ray, 0.01, 100.0)["t"][0] > 0


def test_shadow_pass_through_cut_out_canopy(oracle_lib):
    W, H = 320, 180
    o, g, flat = make_pair(oracle_lib, "tcourt")
    pg = Context(0)  # same geometry, opaque canopy
    pg.scene_upload(get_scene("court")); pg.bvh_build()
    noise = synth.blue_noise_like(4, 64)
    for c in (o, g, pg):
        c.shadow_set_noise(noise); c.shadow_init(W, H)
    light = Light.default()
    prev = None
    for f in range(3):
        cam = make_camera((-5.0 + 0.5 * f, 2.0, 4.5 - 0.3 * f), (0.0, 1.0, 0.0), aspect=W / H, frame_index=f)
        prev = prev or cam
        o.gbuffer_generate(cam); g.gbuffer_generate(cam)
        pd_o, nm_o = o.gbuffer_download()
        pd, nm = g.gbuffer_download()
        assert pd_o.tobytes() == pd.tobytes(), "G-buffer fixture (primary rays see through the cut-outs) differs"
        o.gbuffer_upload(pd, nm)
        g.shadow_frame(cam, prev, light)
        dirs, mask_g = g.shadow_download_debug()
        o.shadow_frame(cam, prev, light, dir_override=dirs)
        raw_o, _, mask_o = o.shadow_download(0)
        assert np.array_equal(mask_o, mask_g), "frame %d: shadow mask differs" % f
        for stage, tol in ((1, 1e-3),):
            assert np.abs(o.shadow_download(stage)[0].astype(np.float64) - g.shadow_download(stage)).max() < tol
        e = np.abs(o.shadow_download(2)[0].astype(np.float64) - g.shadow_download(2)).max(axis=-1)
        assert float((e > 1e-3).mean()) < 1e-3
        o.shadow_set_history(g.shadow_download(2))
        # the same frame with an opaque canopy: light passes through the holes only in the textured scene
        pg.gbuffer_upload(pd, nm)
        pg.shadow_frame(cam, prev, light)
        _, mask_plain = pg.shadow_download_debug()
        lit_more = ((mask_g == 1) & (mask_plain == 2)).sum()
        assert lit_more > 50 and ((mask_g == 2) & (mask_plain == 1)).sum() == 0
        prev = cam


def test_textured_update_differs_from_untextured_only_in_radiance(oracle_lib):
    """Probe pipeline: no any-hit shader, so the rays hit what they hit in the untextured scene (same hit distances; instance ids differ
    because the instance list is sorted by material); radiance changes; parity as usual."""
    o, g, flat = make_pair(oracle_lib, "tcourt")
    gp = Context(0)
    gp.scene_upload(get_scene("court")); gp.bvh_build()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (6, 5, 6), 128)
    for c in (o, g, gp):
        if c is not o:
            c.probes_debug(True)
        c.probes_init(grid)
    host = oracle_lib.HostLogic()
    light = Light.default()
    for frame in range(2):
        R, _ = host.next_orientation()
        for c in (o, g, gp):
            c.probes_update(grid, light, R, None)
        grid.hysteresis = 0.6
        io, do, sto, _ = o.probes_download()  # the next frame reads these atlases: continue from the oracle's (a 1-ulp difference can
        g.probes_upload(io, do, sto)          # flip a packed code, which is 2^-6 relative; same protocol as test_ddgi_parity)
    hg, sg = g.probes_download_hits()
    hp, sp = gp.probes_download_hits()
    ho, so = o.probes_download_hits()
    assert hg.tobytes() == ho.tobytes() and np.array_equal(sg, so)
    assert np.array_equal(hg["t"], hp["t"]), "hit distances must not depend on the textures"
    assert np.array_equal(sg, sp), "the nested shadow rays of the probe pipeline do not see the cut-outs either"
    ro = o.probes_download(rays=True)[3]
    rg = g.probes_download(rays=True)[3]
    rp = gp.probes_download(rays=True)[3]
    assert np.array_equal(ro[..., 3], rg[..., 3])
    e = rel_err(ro[..., :3], rg[..., :3])
    print("textured update: ray radiance max rel err %.2e" % e.max())
    assert e.max() < 1e-3
    assert np.abs(rg[..., :3] - rp[..., :3]).max() > 1e-2
    uio, udo = o.probes_download_unpacked()
    uig, udg = g.probes_download_unpacked()
    assert rel_err(uio, uig).max() < 1e-3 and rel_err(udo, udg).max() < 1e-3
