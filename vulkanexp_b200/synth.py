"""Deterministic synthetic `.scene` generators.

The scenes BASELINE.json names (data/defaut.scene, sponza_test.scene, dungeon.scene, nature_test.scene) are listed in
the reference's .MISSING_LARGE_BLOBS and are not in the checkout, so every config runs on a seeded synthetic scene
of the stated scale. All generators return a scene_format.SceneFile that write_scene() serialises to the
reference's binary container; materials are untextured.
"""
from __future__ import annotations

import numpy as np

from .pods import VERTEX_DTYPE
from .scene_format import Entity, Mesh, SceneFile, material_json

SEED_BASE = 0xD61C0DE


# ------------------------------------------------------------------------------------------------ mesh builders
def _mesh(name, material, pos, nrm, idx):
    v = np.zeros(len(pos), dtype=VERTEX_DTYPE)
    v["pos"] = np.asarray(pos, dtype=np.float32)
    v["color"] = 1.0
    n = np.asarray(nrm, dtype=np.float64)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    v["normal"] = n.astype(np.float32)
    v["tangent"] = [1.0, 0.0, 0.0, 1.0]
    return Mesh(name, material, v, np.asarray(idx, dtype=np.uint32).reshape(-1))


def grid_patch(name, material, origin, du, dv, nu, nv, height_fn=None):
    """(nu x nv) quads spanning origin + s*du + t*dv; front face (CCW) normal = du x dv."""
    origin, du, dv = (np.asarray(x, dtype=np.float64) for x in (origin, du, dv))
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    pos = origin + s[..., None] * du + t[..., None] * dv
    n = np.cross(du, dv)
    n /= np.linalg.norm(n)
    nrm = np.broadcast_to(n, pos.shape).copy()
    if height_fn is not None:
        h, gs, gt = height_fn(s, t)
        pos = pos + h[..., None] * n
        # normal of the displaced surface
        tu = du + gs[..., None] * n
        tv = dv + gt[..., None] * n
        nrm = np.cross(tu, tv)
    pos = pos.reshape(-1, 3)
    nrm = nrm.reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    a = (j * (nu + 1) + i).reshape(-1)
    b = a + 1
    c = a + (nu + 1)
    d = c + 1
    idx = np.stack([a, b, d, a, d, c], axis=1)
    return _mesh(name, material, pos, nrm, idx)


def merge(name, material, meshes):
    pos, nrm, idx, base = [], [], [], 0
    for m in meshes:
        pos.append(m.vertices["pos"])
        nrm.append(m.vertices["normal"])
        idx.append(m.indices + base)
        base += len(m.vertices)
    return _mesh(name, material, np.concatenate(pos), np.concatenate(nrm), np.concatenate(idx))


def room(name, material, lo, hi, sub=4, open_top=False):
    """Axis-aligned box seen from the inside (normals point inward)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    e = hi - lo
    X, Y, Z = np.array([e[0], 0, 0]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])
    faces = [
        grid_patch("f", material, lo, Z, X, sub, sub),  # floor, normal +y
        grid_patch("f", material, lo, Y, Z, sub, sub),  # x = lo, normal +x
        grid_patch("f", material, lo + X, Z, Y, sub, sub),  # x = hi, normal -x
        grid_patch("f", material, lo, X, Y, sub, sub),  # z = lo, normal +z
        grid_patch("f", material, lo + Z, Y, X, sub, sub),  # z = hi, normal -z
    ]
    if not open_top:
        faces.append(grid_patch("f", material, lo + Y, X, Z, sub, sub))  # ceiling, normal -y
    return merge(name, material, faces)


def box(name, material, lo, hi, sub=1):
    """Axis-aligned box seen from the outside."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    e = hi - lo
    X, Y, Z = np.array([e[0], 0, 0]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])
    faces = [
        grid_patch("f", material, lo, X, Z, sub, sub),  # bottom, -y
        grid_patch("f", material, lo + Y, Z, X, sub, sub),  # top, +y
        grid_patch("f", material, lo, Z, Y, sub, sub),  # -x
        grid_patch("f", material, lo + X, Y, Z, sub, sub),  # +x
        grid_patch("f", material, lo, Y, X, sub, sub),  # -z
        grid_patch("f", material, lo + Z, X, Y, sub, sub),  # +z
    ]
    return merge(name, material, faces)


def uv_sphere(name, material, radius=1.0, segments=32, rings=16):
    """32 x 16 gives 960 triangles, the size of the reference's data/debug-models/sphere.gltf."""
    pos, nrm = [[0, radius, 0]], [[0, 1, 0]]
    for r in range(1, rings):
        th = np.pi * r / rings
        for s in range(segments):
            ph = 2 * np.pi * s / segments
            n = [np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)]
            nrm.append(n)
            pos.append([radius * n[0], radius * n[1], radius * n[2]])
    pos.append([0, -radius, 0])
    nrm.append([0, -1, 0])
    idx = []
    ring = lambda r, s: 1 + (r - 1) * segments + (s % segments)
    for s in range(segments):
        idx.append([0, ring(1, s + 1), ring(1, s)])
    for r in range(1, rings - 1):
        for s in range(segments):
            a, b, c, d = ring(r, s), ring(r, s + 1), ring(r + 1, s), ring(r + 1, s + 1)
            idx.append([a, b, d])
            idx.append([a, d, c])
    last = len(pos) - 1
    for s in range(segments):
        idx.append([last, ring(rings - 1, s), ring(rings - 1, s + 1)])
    return _fix_winding(_mesh(name, material, pos, nrm, idx))


def cylinder(name, material, radius, height, segments, stacks, flute=0.0, caps=True):
    """Column along +y from y=0 to y=height, optionally fluted (radius modulation) to add triangles that matter."""
    pos, nrm = [], []
    for k in range(stacks + 1):
        y = height * k / stacks
        for s in range(segments):
            ph = 2 * np.pi * s / segments
            r = radius * (1.0 + flute * np.cos(12 * ph))
            pos.append([r * np.cos(ph), y, r * np.sin(ph)])
            nrm.append([np.cos(ph), 0, np.sin(ph)])
    idx = []
    at = lambda k, s: k * segments + (s % segments)
    for k in range(stacks):
        for s in range(segments):
            a, b, c, d = at(k, s), at(k, s + 1), at(k + 1, s), at(k + 1, s + 1)
            idx.append([a, d, b])
            idx.append([a, c, d])
    if caps:
        for y, ny in ((0.0, -1.0), (height, 1.0)):
            centre = len(pos)
            pos.append([0, y, 0])
            nrm.append([0, ny, 0])
            base = len(pos)
            for s in range(segments):
                ph = 2 * np.pi * s / segments
                pos.append([radius * np.cos(ph), y, radius * np.sin(ph)])
                nrm.append([0, ny, 0])
            for s in range(segments):
                a, b = base + s, base + (s + 1) % segments
                idx.append([centre, a, b] if ny > 0 else [centre, b, a])
    m = _mesh(name, material, pos, nrm, idx)
    return _fix_winding(m)


def _fix_winding(m):
    """Make every triangle's geometric normal agree with its vertex normals (outward)."""
    p = m.vertices["pos"].astype(np.float64)
    n = m.vertices["normal"].astype(np.float64)
    t = m.indices.reshape(-1, 3).copy()
    g = np.cross(p[t[:, 1]] - p[t[:, 0]], p[t[:, 2]] - p[t[:, 0]])
    avg = n[t[:, 0]] + n[t[:, 1]] + n[t[:, 2]]
    flip = (g * avg).sum(axis=1) < 0
    t[flip] = t[flip][:, [0, 2, 1]]
    m.indices = t.reshape(-1).astype(np.uint32)
    return m


def arch(name, material, span, thickness, depth, segments, radial=2):
    """Half-ring (semicircular arch) in the xy plane, centred at the origin, extruded along z."""
    r0, r1 = span / 2 - thickness, span / 2
    pos, nrm, idx = [], [], []

    def add_quad_strip(pts_a, pts_b, normals):
        base = len(pos)
        for a, b, n in zip(pts_a, pts_b, normals):
            pos.extend([a, b])
            nrm.extend([n, n])
        for k in range(len(pts_a) - 1):
            i = base + 2 * k
            idx.extend([[i, i + 1, i + 3], [i, i + 3, i + 2]])

    ang = np.linspace(0, np.pi, segments + 1)
    for r, sgn in ((r1, 1.0), (r0, -1.0)):
        a = [[r * np.cos(t), r * np.sin(t), -depth / 2] for t in ang]
        b = [[r * np.cos(t), r * np.sin(t), depth / 2] for t in ang]
        n = [[sgn * np.cos(t), sgn * np.sin(t), 0] for t in ang]
        add_quad_strip(a, b, n)
    for z, sgn in ((-depth / 2, -1.0), (depth / 2, 1.0)):
        for k in range(radial):
            ra, rb = r0 + (r1 - r0) * k / radial, r0 + (r1 - r0) * (k + 1) / radial
            a = [[ra * np.cos(t), ra * np.sin(t), z] for t in ang]
            b = [[rb * np.cos(t), rb * np.sin(t), z] for t in ang]
            add_quad_strip(a, b, [[0, 0, sgn]] * len(ang))
    return _fix_winding(_mesh(name, material, pos, nrm, idx))


def cone_tree(name, material, segments=10, layers=3):
    """A low-poly conifer: trunk + stacked cones. ~ (2 + 2*layers) * segments triangles."""
    parts = [cylinder("trunk", material, 0.12, 0.8, segments, 1, caps=False)]
    for l in range(layers):
        y0, r, h = 0.6 + 0.7 * l, 0.9 - 0.22 * l, 1.1
        pos, nrm, idx = [], [], []
        for s in range(segments):
            ph = 2 * np.pi * s / segments
            pos.append([r * np.cos(ph), y0, r * np.sin(ph)])
            nrm.append([np.cos(ph), 0.6, np.sin(ph)])
        apex = len(pos)
        pos.append([0, y0 + h, 0])
        nrm.append([0, 1, 0])
        centre = len(pos)
        pos.append([0, y0, 0])
        nrm.append([0, -1, 0])
        for s in range(segments):
            a, b = s, (s + 1) % segments
            idx.append([a, apex, b])
            idx.append([centre, a, b])
        parts.append(_fix_winding(_mesh("cone", material, pos, nrm, idx)))
    return merge(name, material, parts)


# ------------------------------------------------------------------------------------------------ transforms
def trs(t=(0, 0, 0), ry=0.0, s=(1, 1, 1), rx=0.0):
    """Column-major mat4 = T * Ry * Rx * S, as 16 python floats rounded through the 6-decimal JSON writer."""
    cy, sy, cx, sx = np.cos(ry), np.sin(ry), np.cos(rx), np.sin(rx)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    if np.isscalar(s):
        s = (s, s, s)
    M = Ry @ Rx @ np.diag(s)
    m = np.identity(4)
    m[:3, :3] = M
    m[:3, 3] = t
    return [float("%.6f" % x) for x in m.T.reshape(16)]


def _scene(materials):
    s = SceneFile(materials=materials)
    s.entities.append(Entity("Root"))
    return s


def _add(s: SceneFile, name, mesh_index, transform=None, parent=0):
    e = Entity(name, transform=transform if transform is not None else trs(), mesh_renderer=(mesh_index, s.meshes[mesh_index].material))
    s.entities.append(e)
    s.entities[parent].children.append(len(s.entities) - 1)
    return len(s.entities) - 1


# ------------------------------------------------------------------------------------------------ configs
def make_cfg1():
    """"defaut-like": closed 10 m room + 5 instanced spheres (960 triangles each) + one box; ~5.2 k triangles."""
    rng = np.random.default_rng(SEED_BASE + 1)
    mats = [
        material_json("white", (0.8, 0.8, 0.8), 0.0, 0.9),
        material_json("red", (0.8, 0.15, 0.1), 0.0, 0.7),
        material_json("metal", (0.9, 0.85, 0.6), 1.0, 0.35),
        material_json("lamp", (1.0, 1.0, 1.0), 0.0, 1.0, emissive=(4.0, 3.5, 3.0)),
    ]
    s = _scene(mats)
    s.meshes.append(room("Room", 0, (-5, 0, -5), (5, 10, 5), sub=4))
    s.meshes.append(uv_sphere("Sphere", 1, 1.0))
    s.meshes.append(box("Box", 2, (-1, 0, -1), (1, 2, 1), sub=2))
    s.meshes.append(box("Lamp", 3, (-1.0, 0, -1.0), (1.0, 0.1, 1.0), sub=1))
    # a window in the ceiling is not cut: the room is closed, light comes from the emissive lamp and the sun never enters
    _add(s, "Room", 0)
    for k in range(5):
        p = rng.uniform(-3.5, 3.5, size=3)
        p[1] = rng.uniform(1.0, 6.0)
        _add(s, "Sphere%d" % k, 1, trs(p, ry=rng.uniform(0, 6.28), s=float(rng.uniform(0.5, 1.2))))
    _add(s, "Box", 2, trs((2.0, 0.0, -2.5), ry=0.5))
    _add(s, "Lamp", 3, trs((0.0, 9.8, 0.0)))
    return s


def make_open_court(seed=SEED_BASE + 9, columns=6, col_segments=16, col_stacks=4):
    """Small open-top courtyard (sun + sky visible) used by the fast parity tests; ~3-4 k triangles."""
    rng = np.random.default_rng(seed)
    mats = [
        material_json("stone", (0.7, 0.68, 0.6), 0.0, 0.9),
        material_json("blue", (0.2, 0.3, 0.8), 0.0, 0.6),
        material_json("gold", (1.0, 0.77, 0.34), 1.0, 0.3),
    ]
    s = _scene(mats)
    s.meshes.append(room("Court", 0, (-8, 0, -6), (8, 7, 6), sub=3, open_top=True))
    s.meshes.append(cylinder("Column", 0, 0.4, 5.0, col_segments, col_stacks, flute=0.04))
    s.meshes.append(uv_sphere("Ball", 1, 1.0, 16, 8))
    s.meshes.append(box("Plinth", 2, (-0.7, 0, -0.7), (0.7, 0.5, 0.7)))
    s.meshes.append(grid_patch("Canopy", 1, (-3, 5.5, -2), (0, 0, 4), (6, 0.8, 0), 6, 6))
    _add(s, "Court", 0)
    for k in range(columns):
        x = -6.0 + 12.0 * k / max(1, columns - 1)
        _add(s, "ColA%d" % k, 1, trs((x, 0.0, -4.0)))
        _add(s, "ColB%d" % k, 1, trs((x, 0.0, 4.0), ry=0.3))
    for k in range(3):
        p = rng.uniform(-4, 4, size=3)
        p[1] = 1.2
        _add(s, "Ball%d" % k, 2, trs(p, s=float(rng.uniform(0.6, 1.2))))
        _add(s, "Plinth%d" % k, 3, trs((p[0], 0.0, p[2]), ry=float(rng.uniform(0, 3.0))))
    _add(s, "Canopy", 4)
    _add(s, "CanopyBack", 4, trs((0, 11.0, 0), rx=np.pi))  # same patch flipped, so the canopy is two-sided
    return s


def planar_uvs(mesh, scale=0.35):
    """Texture coordinates + tangents for a synthetic mesh: box projection along the dominant normal axis (in place)."""
    v = mesh.vertices
    p, n = v["pos"].astype(np.float64), v["normal"].astype(np.float64)
    axis = np.argmax(np.abs(n), axis=1)
    ua = np.where(axis == 0, 2, 0)  # u runs along z for x-facing faces, along x otherwise
    va = np.where(axis == 1, 2, 1)  # v runs along z for y-facing faces, along y otherwise
    rows = np.arange(len(v))
    v["texCoord"] = np.stack([p[rows, ua] * scale, p[rows, va] * scale], axis=1).astype(np.float32)
    t = np.zeros((len(v), 3))
    t[rows, ua] = 1.0
    t -= n * np.sum(t * n, axis=1, keepdims=True)  # Gram-Schmidt against the normal
    t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-20)
    v["tangent"] = np.concatenate([t, np.where(axis[:, None] == 1, -1.0, 1.0)], axis=1).astype(np.float32)
    return mesh


def procedural_textures(seed=SEED_BASE + 77):
    """Five small RGBA8 images + their .scene texture entries: albedo (sRGB, 64x32), normal map (UNORM, 32x32), metallic-roughness
    (UNORM, 16x16, NEAREST mip), emissive (sRGB, 20x12: odd mip sizes, clamp / mirror) and a cut-out grate (sRGB, alpha 0 / 255)."""
    from .scene_format import VK_FORMAT_R8G8B8A8_SRGB, VK_FORMAT_R8G8B8A8_UNORM

    rng = np.random.default_rng(seed)

    def smooth(h, w, octaves=3):
        out = np.zeros((h, w))
        for o in range(octaves):
            f = 2 ** (o + 1)
            y, x = np.mgrid[0:h, 0:w]
            out += np.sin(2 * np.pi * (f * x / w + rng.uniform())) * np.sin(2 * np.pi * (f * y / h + rng.uniform())) / f
        return out

    images, entries = [], []

    def add(name, img, fmt, sampler):
        images.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
        entries.append({"source": name, "format": fmt, "sampler": sampler})

    h, w = 32, 64  # 0: albedo, tiles with REPEAT (glTF defaults)
    a = np.zeros((h, w, 4))
    brick = ((np.mgrid[0:h, 0:w][0] // 8 + np.mgrid[0:h, 0:w][1] // 16) % 2).astype(np.float64)
    a[..., 0] = 150 + 70 * brick + 25 * smooth(h, w)
    a[..., 1] = 110 + 60 * brick + 25 * smooth(h, w)
    a[..., 2] = 90 + 40 * brick + 25 * smooth(h, w)
    a[..., 3] = 255
    add("tex_albedo.pam", a, VK_FORMAT_R8G8B8A8_SRGB, {})
    h, w = 32, 32  # 1: tangent-space normal map
    hx, hy = smooth(h, w), smooth(h, w)
    nz = np.sqrt(np.maximum(1.0 - 0.25 * (hx**2 + hy**2), 0.05))
    nm = np.stack([0.5 * hx, 0.5 * hy, nz], axis=-1)
    nm /= np.linalg.norm(nm, axis=-1, keepdims=True)
    n = np.zeros((h, w, 4))
    n[..., :3] = (nm * 0.5 + 0.5) * 255
    n[..., 3] = 255
    add("tex_normal.pam", n, VK_FORMAT_R8G8B8A8_UNORM, {"magFilter": 9729, "minFilter": 9987, "wrapS": 10497, "wrapT": 10497})
    h, w = 16, 16  # 2: metallic (b) / roughness (g), NEAREST_MIPMAP_NEAREST minification
    m = np.zeros((h, w, 4))
    m[..., 1] = 120 + 100 * (smooth(h, w) > 0)
    m[..., 2] = 255 * (rng.uniform(size=(h, w)) > 0.5)
    m[..., 3] = 255
    add("tex_metalrough.pam", m, VK_FORMAT_R8G8B8A8_UNORM, {"magFilter": 9728, "minFilter": 9984, "wrapS": 33648, "wrapT": 10497})
    h, w = 12, 20  # 3: emissive, non-power-of-two, clamp / mirror
    e = np.zeros((h, w, 4))
    e[..., 0] = 255 * (smooth(h, w) > 0.3)
    e[..., 1] = 180 * (smooth(h, w) > 0.3)
    e[..., 2] = 60
    e[..., 3] = 255
    add("tex_emissive.pam", e, VK_FORMAT_R8G8B8A8_SRGB, {"magFilter": 9729, "minFilter": 9986, "wrapS": 33071, "wrapT": 33648})
    h, w = 32, 32  # 4: cut-out grate: opaque bars, holes with alpha 0
    g = np.zeros((h, w, 4))
    yy, xx = np.mgrid[0:h, 0:w]
    bars = ((xx % 8) < 3) | ((yy % 8) < 3)
    g[..., 0], g[..., 1], g[..., 2] = 60 + 30 * bars, 60 + 30 * bars, 70 + 30 * bars
    g[..., 3] = 255 * bars
    add("tex_grate.pam", g, VK_FORMAT_R8G8B8A8_SRGB, {"magFilter": 9729, "minFilter": 9729, "wrapS": 10497, "wrapT": 10497})
    return images, entries


def make_textured_court(seed=SEED_BASE + 9):
    """make_open_court with every texture slot of closesthit.glsl in use: albedo + normal map on the stone, metallic-roughness and
    emissive maps on the balls, and a canopy that is a cut-out grate (alpha 0 holes: the any-hit test of the shadow / reflection rays)."""
    s = make_open_court(seed)
    s.images, s.textures = procedural_textures()
    stone, blue = s.materials[0], s.materials[1]
    stone["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
    stone["normalTexture"] = {"index": 1}
    blue["pbrMetallicRoughness"]["metallicRoughnessTexture"] = {"index": 2}
    blue["pbrMetallicRoughness"]["metallicFactor"] = 1.0
    blue["emissiveTexture"] = {"index": 3}
    blue["emissiveFactor"] = [0.6, 0.5, 0.4]
    grate = material_json("grate", (0.9, 0.9, 0.9), 0.0, 0.8)
    grate["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 4}
    s.materials.append(grate)
    for m in s.meshes:
        planar_uvs(m)
    canopy = next(i for i, m in enumerate(s.meshes) if m.name == "Canopy")
    s.meshes[canopy].material = len(s.materials) - 1
    for e in s.entities:
        if e.mesh_renderer is not None and e.mesh_renderer[0] == canopy:
            e.mesh_renderer = (canopy, len(s.materials) - 1)
    return s


def make_cfg2(target_tris=262_144, textured=False):
    """"sponza-scale" atrium: two storeys of fluted columns and arches around an open courtyard, banners, ~262 k triangles.
    textured=True (hard part 5: "procedurally textured for cfg 2-4 with the defined sampler"): box-projected texture
    coordinates and the procedural texture set on every material (albedo + normal map on floor / walls / columns, metallic-roughness
    and emissive maps on the bronze, albedo on the banners); same geometry, so hit records are those of the untextured scene."""
    rng = np.random.default_rng(SEED_BASE + 2)
    mats = [
        material_json("floor", (0.62, 0.6, 0.55), 0.0, 0.8),
        material_json("wall", (0.75, 0.7, 0.62), 0.0, 0.95),
        material_json("column", (0.8, 0.78, 0.72), 0.0, 0.7),
        material_json("banner_red", (0.7, 0.1, 0.08), 0.0, 0.9),
        material_json("banner_green", (0.1, 0.5, 0.15), 0.0, 0.9),
        material_json("bronze", (0.8, 0.5, 0.25), 1.0, 0.4),
    ]
    s = _scene(mats)
    L, Wd, H = 36.0, 16.0, 14.0  # sponza-like proportions
    # shell: floor + 4 walls, open top (sky and sun reach the courtyard)
    s.meshes.append(room("Shell", 1, (-L / 2, 0, -Wd / 2), (L / 2, H, Wd / 2), sub=24, open_top=True))  # 5*24*24*2 = 5760
    s.meshes.append(grid_patch("Floor", 0, (-L / 2, 0.01, -Wd / 2), (0, 0, Wd), (L, 0, 0), 96, 96))  # 18432
    # gallery floors (first storey ceilings) along both long sides
    s.meshes.append(box("Gallery", 1, (-L / 2, 0, 0), (L / 2, 0.4, 3.0), sub=12))  # 6*12*12*2 = 1728
    # columns: 2 storeys x 2 sides x 12 = 48 instances of a fluted column mesh
    col_seg, col_stack = 80, 23  # 80*23*2 + 2*80 = 3840
    s.meshes.append(cylinder("Column", 2, 0.45, 6.0, col_seg, col_stack, flute=0.035))
    arch_seg = 48
    s.meshes.append(arch("Arch", 2, 3.0, 0.35, 0.9, arch_seg, radial=2))  # (2*48*2 + 2*2*48*2) = 576
    s.meshes.append(grid_patch("Banner", 3, (-0.8, 0, 0), (1.6, 0, 0), (0, -4.5, 0), 16, 48,
                               height_fn=lambda u, v: (0.15 * np.sin(6 * u + 3 * v), 0.9 * np.cos(6 * u + 3 * v), 0.45 * np.cos(6 * u + 3 * v))))
    s.meshes.append(grid_patch("Banner2", 4, (-0.8, 0, 0), (1.6, 0, 0), (0, -4.5, 0), 16, 48,
                               height_fn=lambda u, v: (0.12 * np.sin(5 * u - 2 * v), 0.6 * np.cos(5 * u - 2 * v), -0.24 * np.cos(5 * u - 2 * v))))
    s.meshes.append(uv_sphere("Urn", 5, 0.6, 64, 32))  # 3968
    M = {m.name: i for i, m in enumerate(s.meshes)}
    _add(s, "Shell", M["Shell"])
    _add(s, "Floor", M["Floor"])
    for side, z in ((-1, -Wd / 2 + 3.0), (1, Wd / 2 - 3.0)):
        _add(s, "Gallery%d" % side, M["Gallery"], trs((0, 6.0, -Wd / 2 if side < 0 else Wd / 2 - 3.0)))
        for storey in range(2):
            for k in range(12):
                x = -L / 2 + 1.5 + 3.0 * k
                _add(s, "Col_%d_%d_%d" % (side, storey, k), M["Column"], trs((x, 6.4 * storey, z), ry=float(rng.uniform(0, 0.5))))
                if k < 11:
                    _add(s, "Arch_%d_%d_%d" % (side, storey, k), M["Arch"], trs((x + 1.5, 6.4 * storey + 4.5, z)))
    for k in range(8):
        x = -L / 2 + 4.0 + 4.0 * k
        _add(s, "Banner%d" % k, M["Banner"] if k % 2 == 0 else M["Banner2"], trs((x, 12.5, -Wd / 2 + 3.3 if k % 4 < 2 else Wd / 2 - 3.3)))
    for k in range(4):
        _add(s, "Urn%d" % k, M["Urn"], trs((-12.0 + 8.0 * k, 0.6, float(rng.uniform(-1.5, 1.5)))))
    # top up to the target with a finely tessellated lion-head stand-in (spheres) at the ends
    if textured:
        s.images, s.textures = procedural_textures()
        s.images, s.textures = s.images[:4], s.textures[:4]  # no cut-outs here: the probe pipeline has no any-hit shader anyway
        for name in ("floor", "wall", "column"):
            m = next(m for m in s.materials if m["name"] == name)
            m["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
            m["normalTexture"] = {"index": 1}
        for name in ("banner_red", "banner_green"):
            next(m for m in s.materials if m["name"] == name)["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
        bronze = next(m for m in s.materials if m["name"] == "bronze")
        bronze["pbrMetallicRoughness"]["metallicRoughnessTexture"] = {"index": 2}
        bronze["emissiveTexture"] = {"index": 3}
        bronze["emissiveFactor"] = [0.3, 0.25, 0.2]
        for m in s.meshes:
            planar_uvs(m)
    return s


def count_triangles(s: SceneFile):
    return sum(len(s.meshes[e.mesh_renderer[0]].indices) // 3 for e in s.entities if e.mesh_renderer is not None)


def make_cfg3(cells=10, alpha_grates=False):
    """"dungeon-like": corridor maze with pillars and bar grates (geometry); ~0.5 M triangles. alpha_grates=True adds the cut-out grates
    of SURVEY 8(d) cfg3: quads across corridor cells whose albedo texture has alpha 0 holes (61 % opaque), so the sun-shadow rays
    run the any-hit cut-out test."""
    rng = np.random.default_rng(SEED_BASE + 3)
    mats = [
        material_json("rock", (0.45, 0.43, 0.4), 0.0, 0.95),
        material_json("moss", (0.25, 0.4, 0.2), 0.0, 0.9),
        material_json("iron", (0.5, 0.5, 0.55), 1.0, 0.5),
    ]
    s = _scene(mats)
    cs = 6.0
    half = cells * cs / 2
    bump = lambda u, v: (0.08 * np.sin(40 * u) * np.sin(40 * v), 3.2 * np.cos(40 * u) * np.sin(40 * v), 3.2 * np.sin(40 * u) * np.cos(40 * v))
    s.meshes.append(grid_patch("Ground", 0, (-half, 0, -half), (0, 0, 2 * half), (2 * half, 0, 0), 320, 320, height_fn=bump))  # 204800
    s.meshes.append(box("WallBlock", 0, (-cs / 2, 0, -0.5), (cs / 2, 5.0, 0.5), sub=10))  # 1200
    s.meshes.append(cylinder("Pillar", 1, 0.5, 5.0, 48, 12, flute=0.05))  # 1248
    s.meshes.append(cylinder("Bar", 2, 0.04, 5.0, 8, 1, caps=False))  # 16
    M = {m.name: i for i, m in enumerate(s.meshes)}
    _add(s, "Ground", M["Ground"])
    for i in range(cells + 1):
        for j in range(cells):
            if rng.random() < 0.62:
                _add(s, "WX%d_%d" % (i, j), M["WallBlock"], trs((-half + cs * j + cs / 2, 0, -half + cs * i)))
            if rng.random() < 0.62:
                _add(s, "WZ%d_%d" % (i, j), M["WallBlock"], trs((-half + cs * i, 0, -half + cs * j + cs / 2), ry=np.pi / 2))
    for i in range(cells):
        for j in range(cells):
            if rng.random() < 0.5:
                _add(s, "P%d_%d" % (i, j), M["Pillar"], trs((-half + cs * i + cs / 2 + float(rng.uniform(-1, 1)), 0, -half + cs * j + cs / 2 + float(rng.uniform(-1, 1)))))
            if rng.random() < 0.3:
                for b in range(12):
                    _add(s, "B%d_%d_%d" % (i, j, b), M["Bar"], trs((-half + cs * i + 0.45 * b + 0.3, 0, -half + cs * j + 0.2)))
    if alpha_grates:
        images, entries = procedural_textures()
        s.images, s.textures = [images[4]], [entries[4]]
        grate = material_json("grate", (0.9, 0.9, 0.9), 0.0, 0.8)
        grate["pbrMetallicRoughness"]["baseColorTexture"] = {"index": 0}
        s.materials.append(grate)
        # horizontal grates 4.5 m above the floor (they shadow the corridor below) and vertical ones across corridors
        s.meshes.append(planar_uvs(grid_patch("GrateH", len(s.materials) - 1, (-cs / 2, 4.5, -cs / 2), (0, 0, cs), (cs, 0, 0), 2, 2), scale=1.0))
        s.meshes.append(planar_uvs(grid_patch("GrateV", len(s.materials) - 1, (-cs / 2, 0, 0), (cs, 0, 0), (0, 5.0, 0), 2, 2), scale=1.0))
        gh, gv = len(s.meshes) - 2, len(s.meshes) - 1
        rng2 = np.random.default_rng(SEED_BASE + 33)
        for i in range(cells):
            for j in range(cells):
                r = rng2.random()
                if r < 0.35:
                    _add(s, "GH%d_%d" % (i, j), gh, trs((-half + cs * i + cs / 2, 0, -half + cs * j + cs / 2)))
                elif r < 0.5:
                    _add(s, "GV%d_%d" % (i, j), gv, trs((-half + cs * i + cs / 2, 0, -half + cs * j + cs / 2), ry=float(rng2.integers(0, 2)) * np.pi / 2))
    return s


def make_cfg4(trees=20_000, terrain=512):
    """"nature-like": height-field terrain + instanced conifers; ~2 M instanced triangles."""
    rng = np.random.default_rng(SEED_BASE + 4)
    mats = [material_json("ground", (0.35, 0.3, 0.2), 0.0, 0.95), material_json("needles", (0.1, 0.35, 0.12), 0.0, 0.85), material_json("rockface", (0.5, 0.5, 0.5), 0.0, 0.9)]
    s = _scene(mats)
    size = 400.0
    hf = lambda u, v: (8 * np.sin(5 * u) * np.cos(4 * v) + 2.5 * np.sin(19 * u + 1) * np.sin(23 * v),
                       40 * np.cos(5 * u) * np.cos(4 * v) + 47.5 * np.cos(19 * u + 1) * np.sin(23 * v),
                       -32 * np.sin(5 * u) * np.sin(4 * v) + 57.5 * np.sin(19 * u + 1) * np.cos(23 * v))
    s.meshes.append(grid_patch("Terrain", 0, (-size / 2, 0, -size / 2), (0, 0, size), (size, 0, 0), terrain, terrain, height_fn=hf))
    s.meshes.append(cone_tree("Tree", 1, segments=10, layers=3))
    s.meshes.append(uv_sphere("Boulder", 2, 1.0, 16, 8))
    _add(s, "Terrain", 0)
    for k in range(trees):
        u, v = rng.random(), rng.random()
        h = float(hf(np.array(u), np.array(v))[0])
        # terrain patch: s runs along du = +z, t along dv = +x
        _add(s, "T%d" % k, 1, trs((-size / 2 + size * v, h - 0.1, -size / 2 + size * u), ry=float(rng.uniform(0, 6.28)), s=float(rng.uniform(1.5, 4.0))))
    for k in range(trees // 40):
        u, v = rng.random(), rng.random()
        h = float(hf(np.array(u), np.array(v))[0])
        _add(s, "R%d" % k, 2, trs((-size / 2 + size * v, h, -size / 2 + size * u), s=float(rng.uniform(0.5, 2.5))))
    return s


def make_cfg5(instances=10_000, terrain=1024):
    """Instanced stress scene: ~10 M triangles (10 k instances of a ~0.8 k-triangle mesh + a 2 M-triangle terrain)."""
    rng = np.random.default_rng(SEED_BASE + 5)
    mats = [material_json("ground", (0.4, 0.4, 0.35), 0.0, 0.95), material_json("a", (0.7, 0.3, 0.2), 0.0, 0.7), material_json("b", (0.3, 0.4, 0.7), 0.5, 0.5)]
    s = _scene(mats)
    size = 800.0
    hf = lambda u, v: (6 * np.sin(7 * u) * np.cos(6 * v), 42 * np.cos(7 * u) * np.cos(6 * v), -36 * np.sin(7 * u) * np.sin(6 * v))
    s.meshes.append(grid_patch("Terrain", 0, (-size / 2, 0, -size / 2), (0, 0, size), (size, 0, 0), terrain, terrain, height_fn=hf))
    s.meshes.append(uv_sphere("Blob", 1, 1.0, 28, 15))  # 784 triangles
    s.meshes.append(cylinder("Tower", 2, 0.6, 6.0, 24, 15, flute=0.05))  # 768
    _add(s, "Terrain", 0)
    for k in range(instances):
        u, v = rng.random(), rng.random()
        h = float(hf(np.array(u), np.array(v))[0])
        _add(s, "I%d" % k, 1 + (k & 1), trs((-size / 2 + size * v, h + float(rng.uniform(0, 30)) * (k & 1 == 0), -size / 2 + size * u), ry=float(rng.uniform(0, 6.28)), s=float(rng.uniform(1.0, 5.0))))
    return s


def blue_noise_like(slices=64, size=64, seed=SEED_BASE + 64):
    """Stand-in for data/BlueNoise/64_64/LDR_RGBA_*.png (not shipped on the GPU box): 8-bit white noise / 255 as
    RGBA32F, the value domain Image.cpp:62-69 produces. The pass only needs decorrelated [0,1] values."""
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, size=(slices, size, size, 4), dtype=np.uint8).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def reference_blue_noise(slices=64):
    """The reference's own blue-noise slices (data/BlueNoise/64_64/LDR_RGBA_{0..63}.png, bound by src/VulkanLifecycle.cpp:113-119),
    from the committed fixture tests/golden/blue_noise_ldr_rgba_64.npz (tools/gen_blue_noise_fixture.py decodes the PNGs with the
    product's decoder). RGBA32F = byte / 255, the value domain src/vulkan/Image.cpp:62-69 produces."""
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "blue_noise_ldr_rgba_64.npz")
    rgba8 = np.load(path)["rgba8"][:slices]
    return (rgba8.astype(np.float32) / np.float32(255.0)).astype(np.float32)
