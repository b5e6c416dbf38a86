"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE: import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from vulkanexp_b200.pods import BvhInfo, Camera, GridInfo, HIT_DTYPE, Light, NODE_DTYPE, TRI_DTYPE, VERTEX_DTYPE, mip_chain_texels, texture_array

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_host_create.restype = C.c_void_p
        _LIB.orc_pack_r11g11b10.restype = C.c_uint32
        _LIB.orc_pack_r11g11b10.argtypes = [C.c_float] * 3
        _LIB.orc_pack_rg16f.restype = C.c_uint32
        _LIB.orc_pack_rg16f.argtypes = [C.c_float] * 2
        _LIB.orc_spherical_fibonacci.argtypes = [C.c_float, C.c_float, C.c_void_p]
        _LIB.orc_oct_decode.argtypes = [C.c_float, C.c_float, C.c_void_p]
        _LIB.orc_host_select.restype = C.c_uint32
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def use_all_cores():
    """All cores this process may run on (torchrun sets OMP_NUM_THREADS=1 in the environment of every rank): returns the count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_threads(C.c_int(n))
    return lib().orc_max_threads()


class Oracle:
    def __init__(self):
        self.l = lib()
        self.h = C.c_void_p(self.l.orc_create())
        self.grid = None
        self.count = 0

    def close(self):
        if self.h:
            self.l.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def scene_textures(self, textures):
        arr, keep = texture_array(textures)
        self.l.orc_scene_textures(self.h, arr, C.c_size_t(len(textures)))
        self._tex_dims = [(k.shape[1], k.shape[0]) for k in keep]

    def texture_download(self, index):
        w, h = self._tex_dims[index]
        buf = np.zeros(mip_chain_texels(w, h) * 4, dtype=np.uint8)
        levels = C.c_uint32(0)
        assert self.l.orc_texture_download(self.h, C.c_uint32(index), _p(buf), C.c_size_t(buf.nbytes), C.byref(levels)) == 0
        out, off = [], 0
        for l in range(levels.value):
            lw, lh = max(1, w >> l), max(1, h >> l)
            out.append(buf[off : off + lw * lh * 4].reshape(lh, lw, 4))
            off += lw * lh * 4
        return out

    def texture_sample(self, index, uv, grads=None):
        uv = np.ascontiguousarray(uv, dtype=np.float32)
        g = np.ascontiguousarray(grads, dtype=np.float32) if grads is not None else None
        out = np.zeros((len(uv), 4), dtype=np.float32)
        assert self.l.orc_texture_sample(self.h, C.c_uint32(index), _p(uv), _p(g), C.c_size_t(len(uv)), _p(out)) == 0
        return out

    def scene_upload(self, flat):
        if "textures" in flat:
            self.scene_textures(flat["textures"])
        v, i, o, c, m, inst = (np.ascontiguousarray(flat[k]) for k in ("vertices", "indices", "offsets", "mesh_index_counts", "materials", "instances"))
        self._keep = (v, i, o, c, m, inst)
        self.l.orc_scene_upload(self.h, _p(v), C.c_size_t(len(v)), _p(i), C.c_size_t(len(i)), _p(o), _p(c), C.c_size_t(len(o)), _p(m), C.c_size_t(len(m)), _p(inst), C.c_size_t(len(inst)))

    def skin_vertices(self, joint_transforms, skin_joints, skin_weights, src_offset, dst_offset, motion=False):
        """vertexSkinning.comp: joint_transforms [J, 16] column-major, skin_joints uint16 [n, 4], skin_weights float32 [n, 4]."""
        jt = np.ascontiguousarray(joint_transforms, dtype=np.float32).reshape(-1, 16)
        sj = np.ascontiguousarray(skin_joints, dtype=np.uint16).reshape(-1, 4)
        sw = np.ascontiguousarray(skin_weights, dtype=np.float32).reshape(-1, 4)
        assert len(sj) == len(sw)
        mv = np.zeros((len(sj), 4), dtype=np.float32) if motion else None
        assert 0 == (self.l.orc_skin_vertices(self.h, _p(jt), C.c_size_t(len(jt)), _p(sj), _p(sw), C.c_uint32(src_offset), C.c_uint32(dst_offset), C.c_uint32(len(sj)), _p(mv)))
        return mv

    def vertices_download(self, first, count):
        out = np.zeros(count, dtype=VERTEX_DTYPE)
        assert 0 == (self.l.orc_vertices_download(self.h, C.c_size_t(first), C.c_size_t(count), _p(out)))
        return out

    def bvh_build(self):
        self.l.orc_bvh_build(self.h)

    def instances_update(self, instances):
        inst = np.ascontiguousarray(instances)
        assert self.l.orc_instances_update(self.h, _p(inst), C.c_size_t(len(inst))) == 0

    def bvh_refit(self):
        self.l.orc_bvh_refit(self.h)

    def bvh_info(self):
        info = BvhInfo()
        self.l.orc_bvh_info(self.h, C.byref(info))
        return info

    def bvh_download(self):
        info = self.bvh_info()
        nodes = np.zeros(info.numNodes, dtype=NODE_DTYPE)
        tris = np.zeros(info.numTriangles, dtype=TRI_DTYPE)
        self.l.orc_bvh_download(self.h, _p(nodes), C.c_size_t(nodes.nbytes), _p(tris), C.c_size_t(tris.nbytes))
        return nodes, tris

    def trace(self, origins, dirs, tmin, tmax, mask=0xFF, any_hit=False, alpha_test=False):
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        out = np.zeros(len(o), dtype=HIT_DTYPE)
        ctr = np.zeros(3, dtype=np.uint64)
        self.l.orc_trace(self.h, _p(o), _p(d), C.c_size_t(len(o)), C.c_float(tmin), C.c_float(tmax), C.c_uint32(mask), C.c_int(int(any_hit) | (2 if alpha_test else 0)), _p(out), _p(ctr))
        self.trace_counters = ctr
        return out

    def probes_init(self, grid: GridInfo):
        self.grid = grid
        self.l.orc_probes_init(self.h, C.byref(grid))

    def probes_classify(self, R):
        R = np.ascontiguousarray(R, dtype=np.float32)
        self.l.orc_probes_classify(self.h, _p(R))

    def probes_update(self, grid, light, R, indices=None, threads=0):
        R = np.ascontiguousarray(R, dtype=np.float32)
        self.grid = grid
        if indices is not None:
            indices = np.ascontiguousarray(indices, dtype=np.uint32)
            self.count = len(indices)
        else:
            self.count = grid.probe_count
        sec = C.c_double(0)
        self.l.orc_probes_update(self.h, C.byref(grid), C.byref(light), _p(R), _p(indices), C.c_uint32(self.count), C.c_int(threads), C.byref(sec))
        return sec.value

    def probes_download(self, rays=False):
        (ih, iw), (dh, dw) = self.grid.atlas_shapes()
        irr = np.zeros((ih, iw), dtype=np.uint32)
        dep = np.zeros((dh, dw), dtype=np.uint32)
        st = np.zeros(self.grid.probe_count, dtype=np.uint32)
        r = np.zeros((self.count, self.grid.raysPerProbe, 4), dtype=np.float32) if rays else None
        rc = self.l.orc_probes_download(self.h, _p(irr), _p(dep), _p(st), _p(r), C.c_size_t(r.nbytes if rays else 0))
        assert rc == 0
        return irr, dep, st, r

    def probes_upload(self, irr=None, dep=None, state=None):
        a = [np.ascontiguousarray(x, dtype=np.uint32) if x is not None else None for x in (irr, dep, state)]
        self.l.orc_probes_upload(self.h, _p(a[0]), _p(a[1]), _p(a[2]))

    def probes_download_unpacked(self):
        irr = np.zeros((self.count, 36, 3), dtype=np.float32)
        dep = np.zeros((self.count, 196, 2), dtype=np.float32)
        self.l.orc_probes_download_unpacked(self.h, _p(irr), _p(dep))
        return irr, dep

    def probes_download_hits(self):
        hits = np.zeros((self.count, self.grid.raysPerProbe), dtype=HIT_DTYPE)
        sh = np.zeros((self.count, self.grid.raysPerProbe), dtype=np.uint8)
        self.l.orc_probes_download_hits(self.h, _p(hits), _p(sh))
        return hits, sh

    def probes_counters(self):
        c = np.zeros(7, dtype=np.uint64)
        self.l.orc_probes_counters(self.h, _p(c))
        return {"rays": int(c[0]), "nodes": int(c[1]), "tris": int(c[2]), "front": int(c[3]), "shadow_rays": int(c[4]), "shadow_nodes": int(c[5]), "shadow_tris": int(c[6])}

    # shadows
    def shadow_set_noise(self, noise):
        n = np.ascontiguousarray(noise, dtype=np.float32)
        self.l.orc_shadow_set_noise(self.h, _p(n), C.c_uint32(n.shape[2]), C.c_uint32(n.shape[1]), C.c_uint32(n.shape[0]))

    def shadow_init(self, w, h):
        self.sw, self.sh = w, h
        self.l.orc_shadow_init(self.h, C.c_uint32(w), C.c_uint32(h))

    def gbuffer_generate(self, cam: Camera):
        self.l.orc_gbuffer_generate(self.h, C.byref(cam))

    def gbuffer_upload(self, pd, nm):
        pd = np.ascontiguousarray(pd, dtype=np.float32)
        nm = np.ascontiguousarray(nm, dtype=np.float32)
        self.l.orc_gbuffer_upload(self.h, _p(pd), _p(nm))

    def gbuffer_download(self):
        pd = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        nm = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self.l.orc_gbuffer_download(self.h, _p(pd), _p(nm))
        return pd, nm

    def gbuffer_upload_material(self, ar, em):
        ar = np.ascontiguousarray(ar, dtype=np.float32); em = np.ascontiguousarray(em, dtype=np.float32)
        self.l.orc_gbuffer_upload_material(self.h, _p(ar), _p(em))

    def gbuffer_download_material(self):
        ar = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        em = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self.l.orc_gbuffer_download_material(self.h, _p(ar), _p(em))
        return ar, em

    def final_gather(self, cam, light, reflection=None):
        r = np.ascontiguousarray(reflection, dtype=np.float32) if reflection is not None else None
        sec = C.c_double(0)
        self.l.orc_final_gather(self.h, C.byref(cam), C.byref(light), _p(r) if r is not None else None, C.byref(sec))
        img = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self.l.orc_final_gather_download(self.h, _p(img))
        return img, sec.value

    def reflection_frame(self, cur, prev, light, dir_override=None):
        d = np.ascontiguousarray(dir_override, dtype=np.float32) if dir_override is not None else None
        sec = C.c_double(0)
        self.l.orc_reflection_frame(self.h, C.byref(cur), C.byref(prev), C.byref(light), _p(d), C.byref(sec))
        return sec.value

    def reflection_download(self, stage=2):
        img = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        dirs = np.zeros((self.sh, self.sw, 3), dtype=np.float32)
        hits = np.zeros((self.sh, self.sw), dtype=HIT_DTYPE)
        mask = np.zeros((self.sh, self.sw), dtype=np.uint8)
        self.l.orc_reflection_download(self.h, C.c_int(stage), _p(img), _p(dirs), _p(hits), _p(mask))
        return img, dirs, hits, mask

    def reflection_set_history(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self.l.orc_reflection_set_history(self.h, _p(img))

    def shadow_frame(self, cur, prev, light, dir_override=None):
        d = np.ascontiguousarray(dir_override, dtype=np.float32) if dir_override is not None else None
        sec = C.c_double(0)
        self.l.orc_shadow_frame(self.h, C.byref(cur), C.byref(prev), C.byref(light), _p(d), C.byref(sec))
        return sec.value

    def shadow_download(self, stage=2):
        img = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        dirs = np.zeros((self.sh, self.sw, 3), dtype=np.float32)
        mask = np.zeros((self.sh, self.sw), dtype=np.uint8)
        self.l.orc_shadow_download(self.h, C.c_int(stage), _p(img), _p(dirs), _p(mask))
        return img, dirs, mask

    def shadow_set_history(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self.l.orc_shadow_set_history(self.h, _p(img))

    def shadow_reset_history(self):
        self.l.orc_shadow_reset_history(self.h)


class HostLogic:
    """MSVC-rand replay of glm::sphericalRand + genBasis and selectProbesToUpdate (reference src/IrradianceProbes.cpp)."""

    def __init__(self):
        self.l = lib()
        self.h = C.c_void_p(self.l.orc_host_create())

    def __del__(self):
        if self.h:
            self.l.orc_host_destroy(self.h)
            self.h = None

    def next_orientation(self):
        R = np.zeros(16, dtype=np.float32)
        Z = np.zeros(3, dtype=np.float32)
        self.l.orc_host_next_orientation(self.h, _p(R), _p(Z))
        return R, Z

    def rand(self):
        return self.l.orc_host_rand(self.h)

    def select(self, state, per_update=0):
        state = np.ascontiguousarray(state, dtype=np.uint32)
        out = np.zeros(len(state), dtype=np.uint32)
        n = self.l.orc_host_select(self.h, _p(state), C.c_uint32(len(state)), C.c_uint32(per_update), _p(out))
        return out[:n].copy()


def orientation_from_z(Z):
    R = np.zeros(16, dtype=np.float32)
    Z = np.ascontiguousarray(Z, dtype=np.float32)
    lib().orc_orientation_from_z(_p(Z), _p(R))
    return R


def ray_directions(R, count, n):
    R = np.ascontiguousarray(R, dtype=np.float32)
    out = np.zeros((count, 3), dtype=np.float32)
    lib().orc_ray_directions(_p(R), C.c_uint32(count), C.c_float(n), _p(out))
    return out
