"""TEST INFRASTRUCTURE. Function-by-function pin of oracle/ against oracle/_ref/libglslref.so, the reference's own GLSL text
(src/shaders/{common,irradiance,sky,pbrMetallicRoughness}.glsl + gaussian() of the two filters) compiled as C++ against the
reference's vendored GLM (oracle/ref_glsl/). Shared by tests/test_glsl_ref_pin.py (live sweep when /root/reference is present,
committed fixture otherwise) and tools/gen_golden_glsl.py (writes the fixture)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import pyoracle
from vulkanexp_b200.pods import GridInfo

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libglslref.so")
REFERENCE = os.environ.get("VKX_REFERENCE", "/root/reference")


def build_ref():
    """Compiles oracle/_ref/libglslref.so from the reference's sources where they lie; returns the path or None if the
    reference checkout is not available (GPU box)."""
    if os.path.isdir(os.path.join(REFERENCE, "src", "shaders")) and os.path.isdir(os.path.join(REFERENCE, "ext", "glm")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "ref_glsl"), "REF=" + REFERENCE])
    return REF_SO if os.path.exists(REF_SO) else None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def unit(rng, n):
    v = rng.normal(size=(n, 3))
    return _f(v / np.linalg.norm(v, axis=1, keepdims=True))


def make_inputs(seed, n):
    """Seeded inputs for every pinned function. Ranges cover what the path feeds them (plus edge cases appended by hand)."""
    rng = np.random.default_rng(seed)
    grid = GridInfo.make((-14.5, -0.25, -9.0), (15.25, 12.5, 8.75), (12, 7, 9), 64, hysteresis=0.9)
    (ih, iw), (dh, dw) = grid.atlas_shapes()
    l = pyoracle.lib()
    # atlases: packed texels of plausible values (irradiance 0..8, depth mean 0..6 with variance), a few zero texels
    irr = rng.gamma(1.5, 0.6, size=(ih, iw, 3)).astype(np.float32)
    irr[rng.random((ih, iw)) < 0.03] = 0.0
    pk = np.vectorize(lambda r, g, b: l.orc_pack_r11g11b10(float(r), float(g), float(b)), otypes=[np.uint32])
    irr_packed = pk(irr[..., 0], irr[..., 1], irr[..., 2]).astype(np.uint32)
    mean = rng.uniform(0.05, 6.0, size=(dh, dw)).astype(np.float32)
    m2 = (mean * mean * rng.uniform(1.0, 1.6, size=(dh, dw))).astype(np.float32)
    dep_packed = (mean.astype(np.float16).view(np.uint16).astype(np.uint32) | (m2.astype(np.float16).view(np.uint16).astype(np.uint32) << 16)).astype(np.uint32)
    state = rng.integers(0, 9, size=grid.probe_count).astype(np.uint32)
    state[rng.random(grid.probe_count) < 0.15] = 0
    lo, hi = np.array(grid.extentMin[:]), np.array(grid.extentMax[:])
    pos = _f(rng.uniform(lo - 1.5, hi + 1.5, size=(n, 3)))
    pos[: n // 50] = _f(lo + (hi - lo) * rng.integers(0, 2, size=(n // 50, 3)))  # exactly on the volume's corners / faces
    dirs = unit(rng, n)
    dirs[:6] = _f([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]])
    dirs[6:12] = _f([[0.6, 0.8, 0], [0, 0.6, -0.8], [-0.8, 0, -0.6], [0, -0.6, -0.8], [0.70710677, 0, -0.70710677], [0, 0, -1]])
    inp = {
        "grid": grid, "irr": irr_packed, "dep": dep_packed, "state": state,
        "pos": pos, "nrm": dirs, "view": unit(rng, n),
        "sky_o": _f(np.concatenate([rng.uniform(-60, 60, size=(n - 8, 3)) * [1, 0.3, 1] + [0, 8, 0], [[0, -100.0, 0], [0, -99.9995, 0], [5, -150, 3], [0, 2400, 0], [0, 2500, 0], [0, 9000, 0], [0, -100.01, 0], [1e4, 50, -1e4]]])),
        "sky_d": unit(rng, n), "sun": _f([0.2, 2.0, 0.2]), "sun_color": _f([10, 10, 10]), "brightness": 1.0,
        "albedo": _f(rng.uniform(0, 1, size=(n, 4))), "metal_rough": _f(rng.uniform(0, 1, size=(n, 2))), "light_dir": _f([0.2, 2.0, 0.2]), "light_color": _f([10, 10, 10]),
        "fib_i": _f(rng.integers(0, 512, size=n)), "fib_n": 256.0,
        "oct": _f(np.concatenate([rng.uniform(-1, 1, size=(n - 4, 2)), [[0, 0], [1, 1], [-1, 0.5], [0.5, -0.5]]])),
        "angle": _f(np.concatenate([rng.uniform(-3.2, 3.2, size=n - 2), [0.001, 0.0]])),
        "std": _f(rng.uniform(0.01, 9.0, size=n)), "dist": _f(np.concatenate([rng.integers(-8, 9, size=n // 2), rng.uniform(0, 0.2, size=n - n // 2)])),
        "index": rng.integers(0, grid.probe_count, size=n).astype(np.uint32),
    }
    inp["metal_rough"][: n // 20] = _f(rng.integers(0, 2, size=(n // 20, 2)))
    return inp


class _RefGrid(C.Structure):
    _fields_ = [("extentMin", C.c_float * 3), ("depthSharpness", C.c_float), ("extentMax", C.c_float * 3), ("hysteresis", C.c_float), ("resolution", C.c_int * 3),
                ("raysPerProbe", C.c_uint), ("colorRes", C.c_uint), ("depthRes", C.c_uint), ("shadowBias", C.c_float), ("pad", C.c_uint)]


FUNCTIONS = ("sky", "sky_nosun", "sample_probes", "pbr", "spherical_fibonacci", "oct_decode", "oct_encode", "sphere_to_oct_uv", "rotate_axis", "gaussian", "gaussian_refl", "probe_helpers_i", "probe_helpers_f")


def evaluate(inp, side, ref_lib=None):
    """side = 'oracle' (liboracle.so) or 'ref' (libglslref.so). Returns {function: float32/int32 array}."""
    l = pyoracle.lib()
    n = len(inp["pos"])
    grid = inp["grid"]
    o = pyoracle.Oracle()
    o.probes_init(grid)
    o.probes_upload(inp["irr"], inp["dep"], inp["state"])
    out = {}
    z3, z2, z4, z1 = (lambda: np.zeros((n, 3), np.float32)), (lambda: np.zeros((n, 2), np.float32)), (lambda: np.zeros((n, 4), np.float32)), (lambda: np.zeros(n, np.float32))
    if side == "oracle":
        L, pre = l, "orc_fn_"
        gptr = C.byref(grid)
    else:
        L, pre = ref_lib, "ref_"
        rg = _RefGrid.from_buffer_copy(bytes(grid))
        gptr = C.byref(rg)
        l.orc_fetch_atlas.restype = None
        L.ref_set_fetch(C.cast(l.orc_fetch_atlas, C.c_void_p), o.h)
        keep_state = np.ascontiguousarray(inp["state"])
        L.ref_set_probes(_p(keep_state))
    f = lambda name: getattr(L, pre + name)
    for key, show in (("sky", 1), ("sky_nosun", 0)):
        r = z3(); f("sky")(_p(inp["sky_o"]), _p(inp["sky_d"]), _p(inp["sun"]), _p(inp["sun_color"]), C.c_float(inp["brightness"]), C.c_int(show), C.c_size_t(n), _p(r)); out[key] = r
    r = z3()
    if side == "oracle":
        f("sample_probes")(o.h, _p(inp["pos"]), _p(inp["nrm"]), _p(inp["view"]), C.c_size_t(n), _p(r))
    else:
        f("sample_probes")(gptr, _p(inp["pos"]), _p(inp["nrm"]), _p(inp["view"]), C.c_size_t(n), _p(r))
    out["sample_probes"] = r
    r = z4(); f("pbr")(_p(inp["nrm"]), _p(inp["view"]), _p(inp["light_color"]), _p(inp["light_dir"]), _p(inp["albedo"]), _p(inp["metal_rough"]), C.c_size_t(n), _p(r)); out["pbr"] = r
    r = z3(); f("spherical_fibonacci")(_p(inp["fib_i"]), C.c_float(inp["fib_n"]), C.c_size_t(n), _p(r)); out["spherical_fibonacci"] = r
    r = z3(); f("oct_decode")(_p(inp["oct"]), C.c_size_t(n), _p(r)); out["oct_decode"] = r
    r = z2(); f("oct_encode")(_p(inp["nrm"]), C.c_size_t(n), _p(r)); out["oct_encode"] = r
    r = z2(); f("sphere_to_oct_uv")(_p(inp["nrm"]), C.c_size_t(n), _p(r)); out["sphere_to_oct_uv"] = r
    r = z3(); f("rotate_axis")(_p(inp["nrm"]), _p(inp["view"]), _p(inp["angle"]), C.c_size_t(n), _p(r)); out["rotate_axis"] = r
    r = z1(); f("gaussian")(_p(inp["std"]), _p(inp["dist"]), C.c_size_t(n), _p(r)); out["gaussian"] = r
    r = z1(); f("gaussian_refl")(_p(inp["std"]), _p(inp["dist"]), C.c_size_t(n), _p(r)); out["gaussian_refl"] = r
    ri, rf = np.zeros((n, 8), np.int32), np.zeros((n, 6), np.float32)
    f("probe_helpers")(gptr, _p(inp["index"]), C.c_size_t(n), _p(ri), _p(rf)); out["probe_helpers_i"] = ri; out["probe_helpers_f"] = rf
    o.close()
    return out


def ulp_diff(a, b):
    """Largest distance in units in the last place between two float32 arrays (NaN == NaN; a NaN against a number is inf)."""
    a = np.ascontiguousarray(a, np.float32).ravel(); b = np.ascontiguousarray(b, np.float32).ravel()
    na, nb = np.isnan(a), np.isnan(b)
    if (na != nb).any():
        return float("inf")
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia); ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    d[na] = 0
    return float(d.max()) if d.size else 0.0
