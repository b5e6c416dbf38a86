"""CPU-only: pins the oracle against everything the reference offers as a fixed point for this path.

The reference has no tests or golden vectors. The fixtures under tests/golden/ were generated inside the
build container from the reference checkout by tools/gen_golden.py: the constant border tables of
probesCopyBorders.comp and glm::sphericalRand / genBasis evaluated by the reference's vendored GLM.
"""
import json
import os

import ctypes as C
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_border_formula_equals_reference_tables(oracle_lib):
    t = json.load(open(os.path.join(GOLD, "border_tables.json")))
    lib = oracle_lib.lib()
    for T, dst, src in ((8, t["irradianceCopiesDst"], t["irradianceCopiesSrc"]), (16, t["depthCopiesDst"], t["depthCopiesSrc"])):
        assert len(dst) == 4 * (T - 2) + 4 and len(set(map(tuple, dst))) == len(dst)
        for (dx, dy), (sx, sy) in zip(dst, src):
            out = (C.c_int * 2)()
            lib.orc_border_source(C.c_int(T), C.c_int(dx), C.c_int(dy), out)
            assert (out[0], out[1]) == (sx, sy), (T, dx, dy)
        border = {(x, y) for x in range(T) for y in range(T) if x in (0, T - 1) or y in (0, T - 1)}
        assert border == set(map(tuple, dst))


def test_glm_spherical_rand_and_basis_bit_exact(oracle_lib):
    gold = json.load(open(os.path.join(GOLD, "glm_pin.json")))
    h = oracle_lib.HostLogic()
    for e in gold:
        R, Z = h.next_orientation()
        assert Z.view(np.uint32).tolist() == e["Z"]
        assert R.view(np.uint32).tolist() == e["M"]
        M = R.reshape(4, 4)[:3, :3]  # rows of M (column-major storage of the transpose) = X, Y, Z
        assert np.allclose(M.astype(np.float64) @ M.astype(np.float64).T, np.eye(3), atol=1e-5)


def test_msvc_rand_sequence(oracle_lib):
    h = oracle_lib.HostLogic()
    assert [h.rand() for _ in range(5)] == [41, 18467, 6334, 26500, 19169]  # the well-known MSVC rand() prefix for seed 1


def test_pack_unpack_round_trip(oracle_lib):
    lib = oracle_lib.lib()
    rng = np.random.default_rng(0)
    out = (C.c_float * 3)()
    for code in list(range(0, 2048, 7)) + [2047 - 64, 1, 63, 64]:
        if (code >> 6) == 31:
            continue
        lib.orc_unpack_r11g11b10(C.c_uint32(code | (code << 11) | ((code >> 1) << 22)), out)
        assert lib.orc_pack_r11g11b10(out[0], out[1], out[2]) == code | (code << 11) | ((code >> 1) << 22)
    # RTNE / saturation / negative and NaN
    assert lib.orc_pack_r11g11b10(-1.0, float("nan"), 0.0) == 0
    assert lib.orc_pack_r11g11b10(1e9, 65024.0, 64512.0) == (0x7BF | (0x7BF << 11) | (0x3DF << 22))
    assert lib.orc_pack_r11g11b10(1.0, 1.0 + 1.0 / 128, 1.0 + 3.0 / 128) & 0x3FFFFF == (0x3C0 | ((0x3C0) << 11))  # ties to even
    vals = np.concatenate([rng.uniform(-70000, 70000, 2000), rng.uniform(-1e-4, 1e-4, 2000), [0.0, -0.0, 65504.0, 65520.0, 1e-8, 6e-8]]).astype(np.float32)
    ref = vals.astype(np.float16)
    o2 = (C.c_float * 2)()
    for v, r in zip(vals, ref):
        p = lib.orc_pack_rg16f(float(v), float(v))
        assert (p & 0xFFFF) == int(r.view(np.uint16)), v
        lib.orc_unpack_rg16f(C.c_uint32(p), o2)
        assert np.float32(o2[0]) == np.float32(r) or (np.isnan(o2[0]) and np.isnan(r))


def test_spherical_fibonacci_and_oct_maps(oracle_lib):
    lib = oracle_lib.lib()
    v = (C.c_float * 3)()
    pts = []
    for i in range(256):
        lib.orc_spherical_fibonacci(float(i), 256.0, v)
        pts.append([v[0], v[1], v[2]])
    pts = np.array(pts)
    assert np.allclose(np.linalg.norm(pts, axis=1), 1.0, atol=1e-5)
    assert abs(pts[0][2] - (1 - 1 / 256)) < 1e-6 and (np.diff(pts[:, 2]) < 0).all()  # spiral from +z to -z
    assert np.linalg.norm(pts.mean(axis=0)) < 0.02  # near-uniform
    uv = (C.c_float * 2)()
    for p in pts:
        d = (C.c_float * 3)(*p)
        lib.orc_sphere_to_oct_uv(d, uv)
        assert 0.0 <= uv[0] <= 1.0 and 0.0 <= uv[1] <= 1.0
        lib.orc_oct_decode(C.c_float(2 * uv[0] - 1), C.c_float(2 * uv[1] - 1), v)
        assert np.allclose([v[0], v[1], v[2]], p, atol=2e-5)  # octDecode inverts spherePointToOctohedralUV


def test_scheduler_matches_reference_semantics(oracle_lib):
    """selectProbesToUpdate (reference src/IrradianceProbes.cpp:396-424) restated in numpy."""
    rng = np.random.default_rng(5)
    state = rng.choice([0, 1, 2, 3, 8], size=257).astype(np.uint32)
    h = oracle_lib.HostLogic()
    loop, last = 0, 0
    for per_update in (0, 0, 50, 50, 50, 0, 7):
        got = h.select(state, per_update)
        exp, idx, checked = [], last, 0
        while checked < len(state) and (per_update == 0 or len(exp) < per_update):
            if state[idx] != 0 and (idx + loop) % state[idx] == 0:
                exp.append(idx)
            idx += 1
            if idx >= len(state):
                idx, loop = 0, loop + 1
            checked += 1
        last = idx
        assert got.tolist() == exp


def test_oracle_bvh_matches_brute_force(oracle_lib, scene_getter):
    flat = scene_getter("court")
    o = oracle_lib.Oracle()
    o.scene_upload(flat)
    o.bvh_build()
    nodes, tris = o.bvh_download()
    info = o.bvh_info()
    assert info.numTriangles == int(flat["mesh_index_counts"][flat["instances"]["meshEntry"]].sum() // 3)
    # every triangle referenced exactly once, child boxes conservative
    assert sorted(zip((tris["inst"] & 0xFFFFFF).tolist(), (tris["prim"] & 0x7FFFFFFF).tolist())) == sorted(
        (k, j) for k, e in enumerate(flat["instances"]["meshEntry"]) for j in range(int(flat["mesh_index_counts"][e]) // 3))
    rng = np.random.default_rng(1)
    n = 1500
    org = rng.uniform(flat["bounds_min"] + 0.1, flat["bounds_max"] - 0.1, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    hits = o.trace(org, d, 0.01, 1000.0)
    any_ = o.trace(org, d, 0.01, 1000.0, any_hit=True)
    v0, e1, e2 = (tris[k].astype(np.float64) for k in ("v0", "e1", "e2"))
    for i in range(n):
        p = np.cross(d[i].astype(np.float64), e2); det = (e1 * p).sum(1)
        with np.errstate(all="ignore"):
            inv = 1 / det; tv = org[i].astype(np.float64) - v0; u = (tv * p).sum(1) * inv
            q = np.cross(tv, e1); v = (q * d[i]).sum(1) * inv; t = (e2 * q).sum(1) * inv
        ok = (det != 0) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 0.01) & (t < 1000)
        tb = t[ok].min() if ok.any() else -1.0
        assert abs(tb - hits["t"][i]) <= 1e-3 * max(1.0, abs(tb)), i
        assert (any_["t"][i] > 0) == (tb > 0)
        if tb > 0 and (ok & (np.abs(t - tb) < 1e-4)).sum() == 1:  # coincident two-sided geometry ties are resolved by id
            k = int(np.argmin(np.where(ok, t, np.inf)))
            back = det[k] < 0
            assert bool(hits["primitive"][i] >> 31) == bool(back)


def test_oracle_update_properties(oracle_lib, scene_getter):
    """Size-independent properties of one DDGI update: border texels mirror the interior, hysteresis 1 is idempotent,
    a zero-radiance world converges to zero, depth moments satisfy E[d^2] >= E[d]^2."""
    from vulkanexp_b200.pods import GridInfo, Light

    flat = scene_getter("tiny")
    o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (4, 3, 5), 48)
    o.probes_init(grid)
    o.probes_upload(state=np.ones(grid.probe_count, dtype=np.uint32))
    host = oracle_lib.HostLogic()
    light = Light.default()
    for _ in range(3):
        R, _ = host.next_orientation()
        o.probes_update(grid, light, R)
    irr, dep, st, _ = o.probes_download()
    lib = oracle_lib.lib()
    out = (C.c_int * 2)()
    rx, ry, rz = grid.resolution
    for T, img in ((8, irr), (16, dep)):
        for p in range(grid.probe_count):
            ix, iy, iz = p % rx, (p % (rx * ry)) // rx, p // (rx * ry)
            tile = img[T * iz : T * iz + T, T * (iy * rx + ix) : T * (iy * rx + ix) + T]
            for x in range(T):
                for y in range(T):
                    if x in (0, T - 1) or y in (0, T - 1):
                        lib.orc_border_source(C.c_int(T), C.c_int(x), C.c_int(y), out)
                        assert tile[y, x] == tile[out[1], out[0]]
    d = dep.view(np.float16).reshape(dep.shape[0], dep.shape[1], 2).astype(np.float64)
    assert (d[..., 1] + 1e-2 * np.maximum(1.0, d[..., 1]) >= d[..., 0] ** 2).all()
    # hysteresis = 1: the update must leave both atlases unchanged
    grid.hysteresis = 1.0
    R, _ = host.next_orientation()
    o.probes_update(grid, light, R)
    irr2, dep2, _, _ = o.probes_download()
    assert np.array_equal(irr, irr2) and np.array_equal(dep, dep2)


def test_oracle_final_gather_properties(oracle_lib, scene_getter):
    """FinalGather.frag restatement: with all probes off and an unlit shadow image a geometry pixel is emissive only; the
    composite is linear in the reflection input with slope mix(0.004, albedo, metalness); sky pixels ignore every input."""
    from vulkanexp_b200.pods import GridInfo, Light, make_camera

    W, H = 96, 54
    flat = scene_getter("court")
    o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (4, 3, 4), 32)
    o.probes_init(grid)  # state 0 everywhere: sampleProbes returns 0
    o.shadow_init(W, H)
    cam = make_camera((-5.0, 2.5, 4.5), (0.0, 8.0, 0.0), aspect=W / H, frame_index=0)
    o.gbuffer_generate(cam)
    pd, nm = o.gbuffer_download()
    ar, em = o.gbuffer_download_material()
    geo = pd[..., 3] > 0
    assert 0.2 < geo.mean() < 1.0, "the view must contain both geometry and sky"
    assert np.array_equal(em[geo][:, 3], np.ones(int(geo.sum()), dtype=np.float32)) and not em[~geo].any()
    light = Light.default()
    base, _ = o.final_gather(cam, light)
    assert np.array_equal(base[geo][:, :3], em[geo][:, :3]), "no light, no probes: emissive only"
    assert base[~geo][:, :3].min() >= 0.0 and base[~geo][:, :3].max() > 0.0, "sky pixels are lit by the atmosphere"
    refl = np.full((H, W, 4), 0.5, dtype=np.float32)
    withr, _ = o.final_gather(cam, light, refl)
    assert np.array_equal(withr[~geo], base[~geo])
    spec = 0.004 * (1.0 - nm[..., 3:4]) + ar[..., :3] * nm[..., 3:4]
    assert np.allclose((withr - base)[geo][:, :3], 0.5 * spec[geo], rtol=1e-5, atol=1e-7)
    # direct term: a fully lit shadow image adds the PBR direct lighting, never darkens
    o.shadow_set_history(np.ones((H, W, 4), dtype=np.float32))
    lit, _ = o.final_gather(cam, light)
    assert (lit[geo][:, :3] >= base[geo][:, :3]).all() and lit[geo][:, :3].max() > base[geo][:, :3].max()


def test_oracle_reflection_properties(oracle_lib, scene_getter):
    """reflection.rgen / reflectionFilter.glsl restatement: only pixels with roughness < 0.4 or metalness > 0.01 trace a ray;
    a mirror (roughness 0) reflects exactly about the normal and skips both filters; a still camera blends 98 % history."""
    from vulkanexp_b200 import synth
    from vulkanexp_b200.pods import GridInfo, Light, make_camera

    W, H = 96, 54
    flat = scene_getter("court")
    o = oracle_lib.Oracle(); o.scene_upload(flat); o.bvh_build()
    grid = GridInfo.make(flat["bounds_min"], flat["bounds_max"], (4, 3, 4), 32)
    o.probes_init(grid)
    o.shadow_set_noise(synth.blue_noise_like(4, 64)); o.shadow_init(W, H)
    cam = make_camera((-5.0, 2.5, 4.5), (0.0, 3.0, 0.0), aspect=W / H, frame_index=0)
    o.gbuffer_generate(cam)
    pd, nm = o.gbuffer_download()
    ar, em = o.gbuffer_download_material()
    geo = pd[..., 3] > 0
    light = Light.default()
    o.reflection_frame(cam, cam, light)
    raw, dirs, hits, mask = o.reflection_download(0)
    want = geo & ((ar[..., 3] < 0.4) | (nm[..., 3] > 0.01))
    assert np.array_equal(mask > 0, want) and 0 < want.mean() < 1
    assert not raw[~want].any() and np.array_equal(hits["t"][~want], np.full(int((~want).sum()), -1.0, dtype=np.float32))
    # mirrors everywhere
    ar2 = ar.copy(); ar2[..., 3] = 0.0
    o.gbuffer_upload_material(ar2, em)
    o.reflection_frame(cam, cam, light)
    raw, dirs, hits, mask = o.reflection_download(0)
    n = nm[..., :3].astype(np.float64); p = pd[..., :3].astype(np.float64)
    to_origin = np.array([-5.0, 2.5, 4.5]) - p; to_origin /= np.linalg.norm(to_origin, axis=-1, keepdims=True) + 1e-30
    mirror = -to_origin - 2.0 * (n * -to_origin).sum(-1, keepdims=True) * n
    mirror /= np.linalg.norm(mirror, axis=-1, keepdims=True) + 1e-30
    assert np.abs(dirs[geo] - mirror[geo]).max() < 1e-5
    assert np.array_equal(mask > 0, geo)
    x_img = o.reflection_download(1)[0]; fin = o.reflection_download(2)[0]
    assert np.array_equal(x_img, raw) and np.array_equal(fin, raw), "roughness 0: both filters pass the pixel through"
    assert (raw[geo][:, :3] >= 0).all() and raw[geo][:, :3].max() > 0
    # rough metal: still camera -> 0.98 history weight
    ar3 = ar.copy(); ar3[..., 3] = 0.3
    nm3 = nm.copy(); nm3[..., 3] = 1.0
    o.gbuffer_upload(pd, nm3); o.gbuffer_upload_material(ar3, em)
    hist = np.zeros((H, W, 4), dtype=np.float32); hist[..., :3] = 7.0; hist[..., 3] = pd[..., 3]
    o.reflection_set_history(hist)
    o.reflection_frame(cam, cam, light)
    x_img = o.reflection_download(1)[0]; fin = o.reflection_download(2)[0]
    inner = geo.copy(); inner[:6] = inner[-6:] = False; inner[:, :6] = inner[:, -6:] = False
    assert np.array_equal(fin[..., 3][geo], pd[..., 3][geo]), "final alpha carries the depth"
    assert (fin[inner][:, :3] > 0.9 * 7.0 * 0.98).mean() > 0.9, "a still camera keeps 98 % of a matching history"


def _moved_instances(flat, seed=3):
    inst = flat["instances"].copy()
    rng = np.random.default_rng(seed)
    for k in range(1, len(inst)):  # translate + rotate about y every instance but the room
        ang = float(rng.uniform(0, 2 * np.pi)); c, s = np.float32(np.cos(ang)), np.float32(np.sin(ang))
        M = inst[k]["transform"].reshape(3, 4).copy()
        R = np.eye(3, dtype=np.float32); R[0, 0] = c; R[0, 2] = s; R[2, 0] = -s; R[2, 2] = c
        M = (R @ M).astype(np.float32)
        M[0, 3] += np.float32(rng.uniform(-0.8, 0.8)); M[1, 3] += np.float32(rng.uniform(0.0, 0.5)); M[2, 3] += np.float32(rng.uniform(-0.8, 0.8))
        inst[k]["transform"] = M.reshape(-1)
    return inst


def test_refit_restatement_properties():
    """Topology-preserving refit (Renderer::updateTLAS, reference src/Renderer.cpp:681-742, on the wide BVH): with unchanged transforms
    it reproduces the built structure byte for byte; after moving instances it keeps topology (child / triangle layout words) and
    still finds every hit a rebuilt structure finds (same triangles, conservative boxes)."""
    from conftest import get_scene
    from oracle import pyoracle

    flat = get_scene("cfg1")
    o = pyoracle.Oracle(); o.scene_upload(flat); o.bvh_build()
    n0, t0 = o.bvh_download()
    o.bvh_refit()
    n1, t1 = o.bvh_download()
    assert n0.tobytes() == n1.tobytes() and t0.tobytes() == t1.tobytes(), "refit with unchanged transforms must be the identity"
    inst = _moved_instances(flat)
    o.instances_update(inst); o.bvh_refit()
    n2, t2 = o.bvh_download()
    assert n2.tobytes() != n0.tobytes()
    for f in ("imask", "childBase", "primBase", "valid"):
        assert np.array_equal(n2[f], n0[f]), f
    assert np.array_equal(t2["inst"] & 0xFFFFFF, t0["inst"] & 0xFFFFFF) and np.array_equal(t2["prim"] & 0x7FFFFFFF, t0["prim"] & 0x7FFFFFFF)
    moved = dict(flat); moved["instances"] = inst
    r = pyoracle.Oracle(); r.scene_upload(moved); r.bvh_build()
    rng = np.random.default_rng(9)
    lo, hi = np.array(flat["bounds_min"]), np.array(flat["bounds_max"])
    org = rng.uniform(lo, hi, size=(20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ha = o.trace(org, d, 0.001, 1000.0)
    hb = r.trace(org, d, 0.001, 1000.0)
    same = (ha["t"] == hb["t"]) & (ha["instance"] == hb["instance"]) & (ha["primitive"] == hb["primitive"])
    assert same.mean() > 0.9999, "refit and rebuilt structures must find the same hits (up to grazing ties): %g" % same.mean()
    assert (ha["t"] > 0).mean() > 0.5
