"""CPU tests of the oracle's vertexSkinning.comp restatement (oracle/ddgi.cpp::skinVertices; reference
src/shaders/vertexSkinning.comp:37-60, src/Renderer.cpp:133-164,201-240): an independent numpy evaluation in fp32 and
size-independent properties. The reference ships no fixtures for this path."""
import numpy as np
import pytest

from conftest import get_scene
from vulkanexp_b200 import scene_format
from vulkanexp_b200.pods import INSTANCE_SKINNED, INSTANCE_STATIC

f32 = np.float32


def _rigid(angle, axis, t):
    axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = t
    return M.T.reshape(16).astype(np.float32)  # column-major


def skinning_inputs(size, joints=5, seed=3):
    rng = np.random.default_rng(seed)
    jt = np.stack([_rigid(rng.uniform(-0.6, 0.6), rng.normal(size=3), rng.uniform(-0.5, 0.5, 3)) for _ in range(joints)])
    sj = rng.integers(0, joints, (size, 4)).astype(np.uint16)
    sw = rng.uniform(0, 1, (size, 4)).astype(np.float32)
    sw = (sw / sw.sum(axis=1, keepdims=True)).astype(np.float32)
    return jt, sj, sw


def _numpy_skin(verts, jt, sj, sw, src, dst, size):
    """fp32, operation for operation (component-wise sums left to right; mat4 * vec4 as (m0 x + m1 y) + (m2 z + m3 w))."""
    v = verts.copy()
    J = jt.reshape(-1, 4, 4)  # [joint][col][row]
    M = ((sw[:, 0, None, None] * J[sj[:, 0]] + sw[:, 1, None, None] * J[sj[:, 1]]) + sw[:, 2, None, None] * J[sj[:, 2]]) + sw[:, 3, None, None] * J[sj[:, 3]]
    p = verts["pos"][src : src + size]
    new = ((M[:, 0, :3] * p[:, 0:1] + M[:, 1, :3] * p[:, 1:2]) + (M[:, 2, :3] * p[:, 2:3] + M[:, 3, :3] * f32(1))).astype(f32)
    mot = (new - verts["pos"][dst : dst + size]).astype(f32)
    n, t = verts["normal"][dst : dst + size], verts["tangent"][dst : dst + size, :3]
    nn = ((M[:, 0, :3] * n[:, 0:1] + M[:, 1, :3] * n[:, 1:2]) + M[:, 2, :3] * n[:, 2:3]).astype(f32)
    tt = ((M[:, 0, :3] * t[:, 0:1] + M[:, 1, :3] * t[:, 1:2]) + M[:, 2, :3] * t[:, 2:3]).astype(f32)
    v["pos"][dst : dst + size] = new
    v["normal"][src : src + size] = nn
    v["tangent"][src : src + size, :3] = tt
    return v, mot


def test_skinning_matches_numpy_and_keeps_the_shader_quirk(oracle_lib):
    flat, src, dst, size = scene_format.add_skinned_instance(get_scene("court"), 2)  # the ball mesh
    assert flat["instances"][-1]["mask"] == INSTANCE_SKINNED and flat["offsets"][-1]["vertexOffset"] == dst
    o = oracle_lib.Oracle()
    o.scene_upload(flat)
    jt, sj, sw = skinning_inputs(size)
    mv = o.skin_vertices(jt, sj, sw, src, dst, motion=True)
    got = o.vertices_download(0, len(flat["vertices"]))
    want, mot = _numpy_skin(flat["vertices"], jt, sj, sw, src, dst, size)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(mv[:, :3], mot) and (mv[:, 3] == 1).all()
    # the quirk: the skinned copy keeps bind-pose normals, the source mesh receives the skinned ones (vertexSkinning.comp:52-57)
    assert np.array_equal(got["normal"][dst : dst + size], flat["vertices"]["normal"][src : src + size])
    assert not np.array_equal(got["normal"][src : src + size], flat["vertices"]["normal"][src : src + size])
    assert np.array_equal(got["pos"][src : src + size], flat["vertices"]["pos"][src : src + size])
    # a second pose: motion vectors are relative to the previous skinned positions
    jt2, _, _ = skinning_inputs(size, seed=4)
    mv2 = o.skin_vertices(jt2, sj, sw, src, dst, motion=True)
    again = o.vertices_download(dst, size)
    assert np.allclose(mv2[:, :3], again["pos"] - got["pos"][dst : dst + size], atol=1e-6)


def test_identity_pose_and_rigid_pose_properties(oracle_lib):
    base = get_scene("court")
    flat, src, dst, size = scene_format.add_skinned_instance(base, 2)
    o = oracle_lib.Oracle()
    o.scene_upload(flat)
    ident = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (3, 1))
    _, sj, sw = skinning_inputs(size, joints=3)
    one_hot = np.zeros_like(sw); one_hot[:, 0] = 1
    mv = o.skin_vertices(ident, sj, one_hot, src, dst, motion=True)
    assert o.vertices_download(0, len(flat["vertices"])).tobytes() == flat["vertices"].tobytes() and not mv[:, :3].any()
    # Every vertex bound to one joint with a rigid transform T == an ordinary instance of the mesh placed with T: same rays, same hits
    # up to rounding (skinning rounds T * p once, the instance path rounds M * p in a different association).
    T = _rigid(0.4, (0.2, 1.0, 0.1), (1.5, 2.5, -1.0))
    o.skin_vertices(np.tile(T.reshape(1, 16), (3, 1)), sj, one_hot, src, dst)
    o.bvh_build()
    rows = T.reshape(4, 4).T[:3].reshape(12)
    flat2 = dict(base)
    inst = flat["instances"][-1:].copy()
    inst[0]["transform"] = rows; inst[0]["meshEntry"] = 2
    flat2["instances"] = np.concatenate([base["instances"], inst])
    o2 = oracle_lib.Oracle(); o2.scene_upload(flat2); o2.bvh_build()
    rng = np.random.default_rng(8)
    n = 20000
    centre = np.array([1.5, 3.7, -1.0], dtype=np.float32)  # around the posed ball
    origins = (centre + rng.normal(size=(n, 3)) * 3.0).astype(np.float32)
    d = centre + rng.normal(size=(n, 3)) * 0.8 - origins
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a, b = o.trace(origins, d, 0.01, 100.0), o2.trace(origins, d, 0.01, 100.0)
    hit = (a["t"] > 0) & (b["t"] > 0)
    assert (hit.mean() > 0.9) and ((a["t"] > 0) == (b["t"] > 0)).mean() > 0.999
    assert np.abs(a["t"][hit] - b["t"][hit]).max() < 1e-3 and (a["primitive"][hit] == b["primitive"][hit]).mean() > 0.99
    # probe rays (mask static | dynamic) do not see skinned instances (traceProbes.rgen:43), shadow rays (0xFF) do
    masked = o.trace(origins, d, 0.01, 100.0, mask=INSTANCE_STATIC | 2)
    inst_ids = a["instance"][hit]
    skinned_id = len(flat["instances"]) - 1
    assert (inst_ids == skinned_id).mean() > 0.1
    assert not (masked["instance"][masked["t"] > 0] == skinned_id).any()
