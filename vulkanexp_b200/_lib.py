"""ctypes binding of libvkexp_b200.so (the C ABI declared in include/vkx.h).

The library is the product: CUDA kernels for sm_100a behind a C ABI. This module only marshals numpy arrays into
it. If the shared library is missing the import of `load()` fails loudly; there is no Python or CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from .pods import BvhInfo, Camera, GridInfo, HIT_DTYPE, Light, NODE_DTYPE, TRI_DTYPE, VERTEX_DTYPE, mip_chain_texels, texture_array

_HERE = os.path.dirname(os.path.abspath(__file__))
# VKX_LIB_PATH: an alternative build of the same library (A/B timing of compile-time variants, tools/gpu_session.sh)
LIB_PATH = os.environ.get("VKX_LIB_PATH") or os.path.join(_HERE, "libvkexp_b200.so")
_LIB = None


class VkxError(RuntimeError):
    """Raised for every non-zero return code (the reference throws std::runtime_error from VK_CHECK)."""

    def __init__(self, code, message):
        super().__init__("vkx error %d: %s" % (code, message))
        self.code = code


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C vulkanexp_b200/csrc`). There is no CPU fallback." % LIB_PATH
            )
        lib = C.CDLL(LIB_PATH)
        lib.vkx_last_error.restype = C.c_char_p
        lib.vkx_last_error.argtypes = [C.c_void_p]
        lib.vkx_launch_count.restype = C.c_uint64
        lib.vkx_launch_count.argtypes = [C.c_void_p]
        lib.vkx_stream.restype = C.c_void_p
        lib.vkx_stream.argtypes = [C.c_void_p]
        lib.vkx_destroy.restype = None
        lib.vkx_destroy.argtypes = [C.c_void_p]
        lib.vkx_host_scene_free.restype = None
        lib.vkx_host_scene_free.argtypes = [C.c_void_p]
        lib.vkx_host_scene_counts.argtypes = [C.c_void_p, C.c_void_p]
        lib.vkx_host_scene_copy.argtypes = [C.c_void_p] * 8
        lib.vkx_host_scene_texture.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.vkx_host_scene_save.argtypes = [C.c_void_p, C.c_char_p]
        lib.vkx_host_scene_load.argtypes = [C.c_char_p, C.c_void_p]
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def image_decode(path):
    """vkx_image_decode: PNG / P6 / P7 file -> uint8 [h, w, 4] (host only, no GPU needed)."""
    l = load()
    w, h = C.c_uint32(0), C.c_uint32(0)
    rc = l.vkx_image_decode(os.fsencode(path), None, C.c_size_t(0), C.byref(w), C.byref(h))
    if rc != 0:
        raise VkxError(rc, "vkx_image_decode(%s) failed" % path)
    out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
    rc = l.vkx_image_decode(os.fsencode(path), _p(out), C.c_size_t(out.nbytes), C.byref(w), C.byref(h))
    if rc != 0:
        raise VkxError(rc, "vkx_image_decode(%s) failed" % path)
    return out


def shard_groups(rz, nranks):
    """(slices per group, groups per rank) of a full-volume sharded update (vkx_shard_groups)."""
    s, k = C.c_uint32(), C.c_uint32()
    rc = load().vkx_shard_groups(C.c_uint32(rz), C.c_int(nranks), C.byref(s), C.byref(k))
    if rc != 0:
        raise VkxError(rc, "vkx_shard_groups(%d, %d)" % (rz, nranks))
    return s.value, k.value


def shard_slices(rz, nranks, rank):
    """The z-slice ranges [(z0, z1), ...] of `rank` in a full-volume sharded update (vkx_shard_slices: the arithmetic
    vkx_probes_update_sharded uses), in z order."""
    _, groups = shard_groups(rz, nranks)
    out = []
    for g in range(groups):
        z0, z1 = C.c_uint32(), C.c_uint32()
        rc = load().vkx_shard_slices(C.c_uint32(rz), C.c_int(nranks), C.c_int(rank), C.c_uint32(g), C.byref(z0), C.byref(z1))
        if rc != 0:
            raise VkxError(rc, "vkx_shard_slices(%d, %d, %d, %d)" % (rz, nranks, rank, g))
        out.append((z0.value, z1.value))
    return out


def shard_range(count, nranks, rank):
    """List positions [first, first + n) of `rank` in vkx_probes_update_sharded_list."""
    f, n = C.c_uint32(0), C.c_uint32(0)
    rc = load().vkx_shard_range(C.c_uint32(count), C.c_int(nranks), C.c_int(rank), C.byref(f), C.byref(n))
    if rc != 0:
        raise VkxError(rc, "vkx_shard_range(%d, %d, %d)" % (count, nranks, rank))
    return f.value, n.value


def host_scene_load(path):
    """The product's own .scene loader + flattening (vkx_host_scene_*; host only): returns the dict Context.scene_upload takes."""
    from .pods import INSTANCE_DTYPE, MATERIAL_DTYPE, OFFSET_DTYPE, Texture

    l = load()
    h = C.c_void_p()
    rc = l.vkx_host_scene_load(os.fsencode(path), C.byref(h))
    if rc != 0:
        raise VkxError(rc, "vkx_host_scene_load(%s) failed" % path)
    try:
        counts = (C.c_size_t * 6)()
        l.vkx_host_scene_counts(h, counts)
        nv, ni, nm, nmat, ninst, ntex = (int(c) for c in counts)
        flat = {
            "vertices": np.zeros(nv, dtype=VERTEX_DTYPE), "indices": np.zeros(ni, dtype=np.uint32), "offsets": np.zeros(nm, dtype=OFFSET_DTYPE),
            "mesh_index_counts": np.zeros(nm, dtype=np.uint32), "materials": np.zeros(nmat, dtype=MATERIAL_DTYPE), "instances": np.zeros(ninst, dtype=INSTANCE_DTYPE),
        }
        bounds = np.zeros(6, dtype=np.float32)
        rc = l.vkx_host_scene_copy(h, _p(flat["vertices"]), _p(flat["indices"]), _p(flat["offsets"]), _p(flat["mesh_index_counts"]), _p(flat["materials"]), _p(flat["instances"]), _p(bounds))
        if rc != 0:
            raise VkxError(rc, "vkx_host_scene_copy failed")
        flat["bounds_min"], flat["bounds_max"] = bounds[:3].copy(), bounds[3:].copy()
        if ntex:
            flat["textures"] = []
            for i in range(ntex):
                d = Texture()
                l.vkx_host_scene_texture(h, C.c_size_t(i), C.byref(d))
                px = np.ctypeslib.as_array(C.cast(d.pixels, C.POINTER(C.c_uint8)), shape=(d.height, d.width, 4)).copy()
                flat["textures"].append({"pixels": px, "srgb": d.srgb, "magFilter": d.magFilter, "minFilter": d.minFilter, "wrapS": d.wrapS, "wrapT": d.wrapT})
        return flat
    finally:
        l.vkx_host_scene_free(h)


def host_scene_resave(path_in, path_out):
    """Scene::loadScene followed by Scene::save through the facade (host only)."""
    l = load()
    h = C.c_void_p()
    rc = l.vkx_host_scene_load(os.fsencode(path_in), C.byref(h))
    if rc != 0:
        raise VkxError(rc, "vkx_host_scene_load(%s) failed" % path_in)
    try:
        rc = l.vkx_host_scene_save(h, os.fsencode(path_out))
        if rc != 0:
            raise VkxError(rc, "vkx_host_scene_save(%s) failed" % path_out)
    finally:
        l.vkx_host_scene_free(h)


class Context:
    """One per GPU (vkx_ctx). Method names follow the C ABI; see include/vkx.h for the reference call each replaces."""

    def __init__(self, device=0):
        self.l = load()
        h = C.c_void_p()
        rc = self.l.vkx_create(C.c_int(device), C.byref(h))
        if rc != 0:
            raise VkxError(rc, self.l.vkx_last_error(None).decode())
        self.h = h
        self.grid = None
        self.count = 0
        self.sw = self.sh = 0

    def close(self):
        if getattr(self, "h", None):
            self.l.vkx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VkxError(rc, self.l.vkx_last_error(self.h).decode())

    # ---- geometry
    def scene_textures(self, textures):
        """textures: list of dicts, see pods.texture_array; [] removes the list."""
        arr, keep = texture_array(textures)
        self._check(self.l.vkx_scene_textures(self.h, arr, C.c_size_t(len(textures))))
        self._tex_dims = [(k.shape[1], k.shape[0]) for k in keep]

    def texture_download(self, index):
        """-> list of levels, each uint8 [h_l, w_l, 4]"""
        w, h = self._tex_dims[index]
        buf = np.zeros(mip_chain_texels(w, h) * 4, dtype=np.uint8)
        levels = C.c_uint32(0)
        self._check(self.l.vkx_texture_download(self.h, C.c_uint32(index), _p(buf), C.c_size_t(buf.nbytes), C.byref(levels)))
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
        self._check(self.l.vkx_texture_sample(self.h, C.c_uint32(index), _p(uv), _p(g), C.c_size_t(len(uv)), _p(out)))
        return out

    def scene_upload(self, flat):
        if "textures" in flat:
            self.scene_textures(flat["textures"])
        v, i, o, c, m, inst = (np.ascontiguousarray(flat[k]) for k in ("vertices", "indices", "offsets", "mesh_index_counts", "materials", "instances"))
        self._check(
            self.l.vkx_scene_upload(self.h, _p(v), C.c_size_t(len(v)), _p(i), C.c_size_t(len(i)), _p(o), _p(c), C.c_size_t(len(o)), _p(m), C.c_size_t(len(m)), _p(inst), C.c_size_t(len(inst)))
        )

    def instances_update(self, instances):
        inst = np.ascontiguousarray(instances)
        self._check(self.l.vkx_instances_update(self.h, _p(inst), C.c_size_t(len(inst))))

    def skin_vertices(self, joint_transforms, skin_joints, skin_weights, src_offset, dst_offset, motion=False):
        """vertexSkinning.comp: joint_transforms [J, 16] column-major, skin_joints uint16 [n, 4], skin_weights float32 [n, 4]."""
        jt = np.ascontiguousarray(joint_transforms, dtype=np.float32).reshape(-1, 16)
        sj = np.ascontiguousarray(skin_joints, dtype=np.uint16).reshape(-1, 4)
        sw = np.ascontiguousarray(skin_weights, dtype=np.float32).reshape(-1, 4)
        assert len(sj) == len(sw)
        mv = np.zeros((len(sj), 4), dtype=np.float32) if motion else None
        self._check(self.l.vkx_skin_vertices(self.h, _p(jt), C.c_size_t(len(jt)), _p(sj), _p(sw), C.c_uint32(src_offset), C.c_uint32(dst_offset), C.c_uint32(len(sj)), _p(mv)))
        return mv

    def vertices_download(self, first, count):
        out = np.zeros(count, dtype=VERTEX_DTYPE)
        self._check(self.l.vkx_vertices_download(self.h, C.c_size_t(first), C.c_size_t(count), _p(out)))
        return out

    def bvh_build(self):
        self._check(self.l.vkx_bvh_build(self.h))

    def bvh_refit(self):
        self._check(self.l.vkx_bvh_refit(self.h))

    def bvh_info(self):
        info = BvhInfo()
        self._check(self.l.vkx_bvh_info_get(self.h, C.byref(info)))
        return info

    def bvh_download(self):
        info = self.bvh_info()
        nodes = np.zeros(info.numNodes, dtype=NODE_DTYPE)
        tris = np.zeros(info.numTriangles, dtype=TRI_DTYPE)
        self._check(self.l.vkx_bvh_download(self.h, _p(nodes), C.c_size_t(nodes.nbytes), _p(tris), C.c_size_t(tris.nbytes)))
        return nodes, tris

    def trace(self, origins, dirs, tmin, tmax, mask=0xFF, any_hit=False, alpha_test=False):
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        out = np.zeros(len(o), dtype=HIT_DTYPE)
        fn = self.l.vkx_trace_alpha if alpha_test else self.l.vkx_trace
        self._check(fn(self.h, _p(o), _p(d), C.c_size_t(len(o)), C.c_float(tmin), C.c_float(tmax), C.c_uint32(mask), C.c_int(int(any_hit)), _p(out)))
        return out

    # ---- DDGI
    def probes_init(self, grid: GridInfo):
        self.grid = grid
        self._check(self.l.vkx_probes_init(self.h, C.byref(grid)))

    def probes_debug(self, enable=True):
        self._check(self.l.vkx_probes_debug(self.h, C.c_int(int(enable))))

    def probes_classify(self, R):
        R = np.ascontiguousarray(R, dtype=np.float32)
        self._check(self.l.vkx_probes_classify(self.h, _p(R)))

    def probes_update(self, grid, light, R, indices=None, sync=True):
        R = np.ascontiguousarray(R, dtype=np.float32)
        self.grid = grid
        if indices is not None:
            indices = np.ascontiguousarray(indices, dtype=np.uint32)
            self.count = len(indices)
        else:
            self.count = grid.probe_count
        self._check(self.l.vkx_probes_update(self.h, C.byref(grid), C.byref(light), _p(R), _p(indices), C.c_uint32(self.count), C.c_int(int(sync))))

    def probes_update_sharded(self, grid, light, R, sync=True):
        R = np.ascontiguousarray(R, dtype=np.float32)
        self.grid = grid
        self._check(self.l.vkx_probes_update_sharded(self.h, C.byref(grid), C.byref(light), _p(R), C.c_int(int(sync))))

    def probes_update_sharded_list(self, grid, light, R, indices, sync=True):
        """Partial update of a to-update list on several GPUs: every rank passes the same list (vkx_probes_update_sharded_list)."""
        R = np.ascontiguousarray(R, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        self.grid = grid
        self.count = len(indices)
        self._check(self.l.vkx_probes_update_sharded_list(self.h, C.byref(grid), C.byref(light), _p(R), _p(indices), C.c_uint32(len(indices)), C.c_int(int(sync))))

    def stream_wait_exchange(self):
        """Orders the context's stream after a pending atlas exchange; an event recorded afterwards covers the all-gather."""
        self._check(self.l.vkx_stream_wait_exchange(self.h))

    def probes_download(self, rays=False, out=None):
        (ih, iw), (dh, dw) = self.grid.atlas_shapes()
        if out is None:
            irr = np.zeros((ih, iw), dtype=np.uint32)
            dep = np.zeros((dh, dw), dtype=np.uint32)
            st = np.zeros(self.grid.probe_count, dtype=np.uint32)
        else:
            irr, dep, st = out
        r = np.zeros((self.count, self.grid.raysPerProbe, 4), dtype=np.float32) if rays else None
        self._check(self.l.vkx_probes_download(self.h, _p(irr), _p(dep), _p(st), _p(r), C.c_size_t(r.nbytes if rays else 0)))
        return irr, dep, st, r

    def probes_download_async(self, out):
        irr, dep, st = out
        self._check(self.l.vkx_probes_download_async(self.h, _p(irr), _p(dep), _p(st)))

    def probes_download_slab_async(self, z0, z1, out):
        irr, dep, st = out
        self._check(self.l.vkx_probes_download_slab_async(self.h, C.c_uint32(z0), C.c_uint32(z1), _p(irr), _p(dep), _p(st)))

    def probes_download_wait(self):
        self._check(self.l.vkx_probes_download_wait(self.h))

    def probes_schedule(self, probes_per_update=0):
        n = C.c_uint32(0)
        self._check(self.l.vkx_probes_schedule(self.h, C.c_uint32(probes_per_update), C.byref(n)))
        return int(n.value)

    def probes_scheduled_list(self):
        n = C.c_uint32(0)
        self._check(self.l.vkx_probes_scheduled_list(self.h, None, C.c_uint32(0), C.byref(n)))
        out = np.zeros(int(n.value), dtype=np.uint32)
        if n.value:
            self._check(self.l.vkx_probes_scheduled_list(self.h, _p(out), C.c_uint32(len(out)), C.byref(n)))
        return out

    def probes_update_scheduled(self, grid, light, R, sync=True):
        R = np.ascontiguousarray(R, dtype=np.float32)
        self.grid = grid
        self._check(self.l.vkx_probes_update_scheduled(self.h, C.byref(grid), C.byref(light), _p(R), C.c_int(int(sync))))

    def probes_scheduler_state(self, set=None):
        get = (C.c_uint32 * 2)()
        st = (C.c_uint32 * 2)(*set) if set is not None else None
        self._check(self.l.vkx_probes_scheduler_state(self.h, st, get))
        return int(get[0]), int(get[1])

    def probes_upload(self, irr=None, dep=None, state=None):
        a = [np.ascontiguousarray(x, dtype=np.uint32) if x is not None else None for x in (irr, dep, state)]
        self._check(self.l.vkx_probes_upload(self.h, _p(a[0]), _p(a[1]), _p(a[2])))

    def probes_download_unpacked(self):
        irr = np.zeros((self.count, 36, 3), dtype=np.float32)
        dep = np.zeros((self.count, 196, 2), dtype=np.float32)
        self._check(self.l.vkx_probes_download_unpacked(self.h, _p(irr), _p(dep)))
        return irr, dep

    def probes_download_hits(self):
        hits = np.zeros((self.count, self.grid.raysPerProbe), dtype=HIT_DTYPE)
        sh = np.zeros((self.count, self.grid.raysPerProbe), dtype=np.uint8)
        self._check(self.l.vkx_probes_download_hits(self.h, _p(hits), _p(sh)))
        return hits, sh

    def probes_timings(self):
        ms = (C.c_float * 5)()
        self._check(self.l.vkx_probes_timings(self.h, ms))
        return {"full": ms[0], "trace": ms[1], "blend": ms[2], "border": ms[3], "publish": ms[4]}

    def probes_kernel_timings(self):
        ms = (C.c_float * 4)()
        probes, shadow = C.c_uint32(0), C.c_uint32(0)
        self._check(self.l.vkx_probes_kernel_timings(self.h, ms, C.byref(probes), C.byref(shadow)))
        return {"trace_primary": ms[0], "shade": ms[1], "trace_shadow": ms[2], "blend": ms[3], "probes": probes.value, "shadow_rays": shadow.value}

    def probes_device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.l.vkx_probes_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_ubyte * 128)()
        rc = load().vkx_comm_unique_id(buf)
        if rc != 0:
            raise VkxError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init(self, rank, nranks, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self.l.vkx_comm_init(self.h, C.c_int(rank), C.c_int(nranks), buf))

    def comm_p2p_export(self):
        buf = (C.c_ubyte * 64)()
        self._check(self.l.vkx_comm_p2p_export(self.h, buf))
        return bytes(buf)

    def comm_p2p_import(self, handles):
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.l.vkx_comm_p2p_import(self.h, buf, C.c_int(len(handles))))

    def comm_p2p_mode(self, copy_engines):
        self._check(self.l.vkx_comm_p2p_mode(self.h, C.c_int(int(copy_engines))))

    def comm_p2p_enable(self, dist):
        """Exchange the IPC handles through torch.distributed and map every peer's atlas slab."""
        try:
            mine = self.comm_p2p_export()
        except VkxError:
            mine = None
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, mine)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                self.comm_p2p_import(handles)
            except VkxError:
                ok = False
        flags = [None] * dist.get_world_size()
        dist.all_gather_object(flags, ok)
        if not all(flags):  # every rank or none: a rank without peer access sends everybody back to the all-gather
            self._check(self.l.vkx_comm_p2p_import(self.h, None, C.c_int(0)))
            return False
        return True

    # ---- shadows
    def shadow_set_noise(self, noise):
        n = np.ascontiguousarray(noise, dtype=np.float32)
        self._check(self.l.vkx_shadow_set_noise(self.h, _p(n), C.c_uint32(n.shape[2]), C.c_uint32(n.shape[1]), C.c_uint32(n.shape[0])))

    def shadow_init(self, w, h):
        self.sw, self.sh = w, h
        self._check(self.l.vkx_shadow_init(self.h, C.c_uint32(w), C.c_uint32(h)))

    def gbuffer_generate(self, cam: Camera):
        self._check(self.l.vkx_gbuffer_generate(self.h, C.byref(cam)))

    def gbuffer_upload(self, pd, nm):
        pd = np.ascontiguousarray(pd, dtype=np.float32)
        nm = np.ascontiguousarray(nm, dtype=np.float32)
        self._check(self.l.vkx_gbuffer_upload(self.h, _p(pd), _p(nm)))

    def gbuffer_download(self):
        pd = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        nm = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self._check(self.l.vkx_gbuffer_download(self.h, _p(pd), _p(nm)))
        return pd, nm

    def gbuffer_upload_material(self, ar, em):
        ar = np.ascontiguousarray(ar, dtype=np.float32); em = np.ascontiguousarray(em, dtype=np.float32)
        self._check(self.l.vkx_gbuffer_upload_material(self.h, _p(ar), _p(em)))

    def gbuffer_download_material(self):
        ar = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        em = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self._check(self.l.vkx_gbuffer_download_material(self.h, _p(ar), _p(em)))
        return ar, em

    def final_gather(self, cam, light, reflection=None, sync=True):
        r = np.ascontiguousarray(reflection, dtype=np.float32) if reflection is not None else None
        self._check(self.l.vkx_final_gather(self.h, C.byref(cam), C.byref(light), _p(r) if r is not None else None, C.c_int(int(sync))))

    def final_gather_download(self, out=None):
        img = out if out is not None else np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        ms = C.c_float(0)
        self._check(self.l.vkx_final_gather_download(self.h, _p(img), C.byref(ms)))
        return img, float(ms.value)

    def reflection_frame(self, cur, prev, light, sync=True):
        self._check(self.l.vkx_reflection_frame(self.h, C.byref(cur), C.byref(prev), C.byref(light), C.c_int(int(sync))))

    def reflection_download(self, stage=2, out=None):
        img = out if out is not None else np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self._check(self.l.vkx_reflection_download(self.h, C.c_int(stage), _p(img)))
        return img

    def reflection_download_debug(self):
        from .pods import HIT_DTYPE
        dirs = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        hits = np.zeros((self.sh, self.sw), dtype=HIT_DTYPE)
        mask = np.zeros((self.sh, self.sw), dtype=np.uint8)
        self._check(self.l.vkx_reflection_download_debug(self.h, _p(dirs), _p(hits), _p(mask)))
        return dirs[..., :3].copy(), hits, mask

    def reflection_reset_history(self):
        self._check(self.l.vkx_reflection_reset_history(self.h))

    def reflection_timings(self):
        ms = (C.c_float * 4)()
        self._check(self.l.vkx_reflection_timings(self.h, ms))
        return {"full": ms[0], "trace_shade": ms[1], "filter_x": ms[2], "filter_y": ms[3]}

    def shadow_frame(self, cur, prev, light, sync=True):
        self._check(self.l.vkx_shadow_frame(self.h, C.byref(cur), C.byref(prev), C.byref(light), C.c_int(int(sync))))

    def shadow_download(self, stage=2, out=None):
        img = out if out is not None else np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        self._check(self.l.vkx_shadow_download(self.h, C.c_int(stage), _p(img)))
        return img

    def shadow_download_debug(self):
        dirs = np.zeros((self.sh, self.sw, 4), dtype=np.float32)
        mask = np.zeros((self.sh, self.sw), dtype=np.uint8)
        self._check(self.l.vkx_shadow_download_debug(self.h, _p(dirs), _p(mask)))
        return dirs[..., :3].copy(), mask

    def shadow_reset_history(self):
        self._check(self.l.vkx_shadow_reset_history(self.h))

    def shadow_timings(self):
        ms = (C.c_float * 4)()
        self._check(self.l.vkx_shadow_timings(self.h, ms))
        return {"full": ms[0], "trace": ms[1], "filter_x": ms[2], "filter_y": ms[3]}

    # ---- misc
    def launch_count(self):
        return int(self.l.vkx_launch_count(self.h))

    def stream(self):
        return self.l.vkx_stream(self.h)

    def sync(self):
        self._check(self.l.vkx_sync(self.h))
