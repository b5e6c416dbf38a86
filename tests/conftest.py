import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


_SCENES = {}


def get_scene(name):
    """Flattened synthetic scene (cached per session)."""
    from vulkanexp_b200 import scene_format, synth

    if name not in _SCENES:
        if name == "court":
            s = synth.make_open_court()
        elif name == "cfg1":
            s = synth.make_cfg1()
        elif name == "cfg2":
            s = synth.make_cfg2()
        elif name == "cfg3":
            s = synth.make_cfg3()
        elif name == "tcourt":  # every texture slot in use + a cut-out canopy (sampler spec v1)
            s = synth.make_textured_court()
        elif name == "tiny":
            s = synth.make_open_court(columns=2, col_segments=6, col_stacks=1)
        else:
            raise KeyError(name)
        _SCENES[name] = scene_format.flatten(s)
    return _SCENES[name]


@pytest.fixture(scope="session")
def scene_getter():
    return get_scene


def make_pair(oracle_lib, scene_name):
    """(oracle, gpu context) with the same scene uploaded and both BVHs built."""
    from vulkanexp_b200._lib import Context

    flat = get_scene(scene_name)
    o = oracle_lib.Oracle()
    o.scene_upload(flat)
    o.bvh_build()
    g = Context(0)
    g.scene_upload(flat)
    g.bvh_build()
    return o, g, flat


def rel_err(a, b, floor=1e-3):
    """Relative error with an absolute floor (values below `floor` compare absolutely)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
