#!/usr/bin/env python
"""Writes tests/golden/glsl_ref_pin.npz: outputs of the REFERENCE's shader functions (oracle/_ref/libglslref.so, built from
/root/reference/src/shaders by oracle/ref_glsl/Makefile) on the seeded inputs of oracle/glsl_pin.py. The fixture travels to
machines without the reference checkout; tests/test_glsl_ref_pin.py compares the oracle against it bit for bit."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import glsl_pin  # noqa: E402

SEED, N = 20261017, 1536

if __name__ == "__main__":
    so = glsl_pin.build_ref()
    if so is None:
        raise SystemExit("reference checkout not found: cannot generate the fixture")
    out = glsl_pin.evaluate(glsl_pin.make_inputs(SEED, N), "ref", C.CDLL(so))
    path = os.path.join(ROOT, "tests", "golden", "glsl_ref_pin.npz")
    np.savez_compressed(path, seed=SEED, n=N, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
