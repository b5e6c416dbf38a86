"""Python view of the host logic exported by the library (csrc/host/HostLogic.cpp): the reference's per-frame random
orientation and probe scheduler (reference src/IrradianceProbes.cpp:347-355, 396-424, 455-460)."""
import ctypes as C

import numpy as np

from . import _lib


class OrientationGenerator:
    """MSVC-rand replay of glm::sphericalRand + genBasis; next() returns the 16-float column-major push constant."""

    def __init__(self, seed=1):
        self.state = C.c_uint32(seed)
        self.l = _lib.load()
        self.l.vkx_host_next_orientation.restype = None

    def next(self):
        out = np.zeros(16, dtype=np.float32)
        self.l.vkx_host_next_orientation(C.byref(self.state), out.ctypes.data_as(C.c_void_p))
        return out


class ProbeScheduler:
    """selectProbesToUpdate with its two function-local statics."""

    def __init__(self):
        self.loop_index = C.c_uint32(0)
        self.last_offset = C.c_uint32(0)
        self.l = _lib.load()
        self.l.vkx_host_select_probes.restype = C.c_uint32

    def select(self, state, probes_per_update=0):
        state = np.ascontiguousarray(state, dtype=np.uint32)
        out = np.zeros(len(state), dtype=np.uint32)
        n = self.l.vkx_host_select_probes(C.byref(self.loop_index), C.byref(self.last_offset), state.ctypes.data_as(C.c_void_p), C.c_uint32(len(state)), C.c_uint32(probes_per_update), out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()
