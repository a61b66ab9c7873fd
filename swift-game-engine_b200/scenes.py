"""Scene recipes and synthetic workloads (numpy only; no device code, no oracle).

Host-side harness shared by tests/ and bench.py: it builds the *inputs* (mesh
parts with their model matrices, query batches, character states) that are then
fed, bit-identically, to both the CUDA library and the CPU oracle.

Scene recipe citations (reference = /root/reference/Game):
  ground plane 80x80 @ y=-3 ......... DemoScene.swift:87,101-131 ; ProceduralMeshes.swift:169-181
  ornate mirror placement ........... DemoScene.swift:293-377 (rot X 270 deg, x8, +(-10,1,4), layer 1<<4)
  matrix -> TRS decomposition ....... DemoScene.swift:718-735
  TRS -> modelMatrix = T*(R*S) ...... Components.swift:26-44
All matrix/quaternion math is float32; it runs once on the host and its result
is shared by oracle and GPU, so last-ulp agreement with Apple's simd is not
needed.
"""
import os

import numpy as np

f32 = np.float32
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

LAYER_DEFAULT = 1 << 0
LAYER_MIRROR = 1 << 4
GROUND_Y = -3.0

# ---------------------------------------------------------------- quaternion / matrix helpers (ix,iy,iz,r)


def quat_angle_axis(angle, axis):
    axis = np.asarray(axis, f32)
    h = f32(angle) * f32(0.5)
    s, c = f32(np.sin(h)), f32(np.cos(h))
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, c], f32)


def quat_mul(a, b):
    ax, ay, az, aw = [f32(v) for v in a]
    bx, by, bz, bw = [f32(v) for v in b]
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], f32)


def quat_to_mat3(q):
    x, y, z, w = [f32(v) for v in q]
    one, two = f32(1), f32(2)
    return np.array([[one - two * (y * y + z * z), two * (x * y - z * w), two * (x * z + y * w)],
                     [two * (x * y + z * w), one - two * (x * x + z * z), two * (y * z - x * w)],
                     [two * (x * z - y * w), two * (y * z + x * w), one - two * (x * x + y * y)]], f32)


def quat_from_mat3(m):
    """Standard trace-based extraction (simd_quatf(matrix_float3x3)); m[row][col]."""
    m = np.asarray(m, f32)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = f32(np.sqrt(tr + f32(1))) * f32(2)
        return np.array([(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, f32(0.25) * s], f32)
    if m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = f32(np.sqrt(f32(1) + m[0, 0] - m[1, 1] - m[2, 2])) * f32(2)
        return np.array([f32(0.25) * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s], f32)
    if m[1, 1] > m[2, 2]:
        s = f32(np.sqrt(f32(1) + m[1, 1] - m[0, 0] - m[2, 2])) * f32(2)
        return np.array([(m[0, 1] + m[1, 0]) / s, f32(0.25) * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s], f32)
    s = f32(np.sqrt(f32(1) + m[2, 2] - m[0, 0] - m[1, 1])) * f32(2)
    return np.array([(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, f32(0.25) * s, (m[1, 0] - m[0, 1]) / s], f32)


def trs_model(translation=(0, 0, 0), rotation=(0, 0, 0, 1), scale=(1, 1, 1)):
    """TransformComponent.modelMatrix (Components.swift:26-44) -> 16 floats, column-major."""
    r = quat_to_mat3(rotation)
    s = np.asarray(scale, f32)
    m = np.zeros((4, 4), f32)  # m[row][col]
    m[:3, :3] = r * s[None, :]
    m[:3, 3] = np.asarray(translation, f32)
    m[3, 3] = 1
    return np.ascontiguousarray(m.T).reshape(16).astype(f32)  # column-major flat


def transform_from_matrix(colmajor16):
    """DemoScene.transformFromMatrix (DemoScene.swift:718-735): (translation, quat, scale)."""
    m = np.asarray(colmajor16, f32).reshape(4, 4).T  # m[row][col]
    t = m[:3, 3].copy()
    cols = [m[:3, i].copy() for i in range(3)]
    sc = np.array([f32(np.sqrt(np.dot(c, c))) for c in cols], f32)
    ident = np.eye(3, dtype=f32)
    rot = np.stack([cols[i] / sc[i] if sc[i] > 0 else ident[:, i] for i in range(3)], axis=1)
    return t, quat_from_mat3(rot), sc


def rowmajor_to_colmajor(values16):
    """StaticMeshLoader.matrixFromArrayRowMajor (StaticMeshLoader.swift:127-134)."""
    return np.ascontiguousarray(np.asarray(values16, f32).reshape(4, 4).T).reshape(16)


# ---------------------------------------------------------------- meshes


def plane_mesh(size=80.0):
    """ProceduralMeshes.plane (ProceduralMeshes.swift:169-181)."""
    s = f32(size) * f32(0.5)
    v = np.array([[-s, 0, s], [s, 0, s], [s, 0, -s], [-s, 0, -s]], f32)
    return v, np.array([0, 1, 2, 0, 2, 3], np.uint32)


def box_mesh(size=1.0):
    """Axis-aligned cube, 12 triangles, outward winding."""
    s = f32(size) * f32(0.5)
    v = np.array([[x, y, z] for x in (-s, s) for y in (-s, s) for z in (-s, s)], f32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    idx = []
    for a, b, c, d in quads:
        idx += [a, b, c, a, c, d]
    return v, np.array(idx, np.uint32)


def part(positions, indices, model=None, layer=LAYER_DEFAULT, mu_s=0.8, mu_k=0.6, flatten_ground=False,
         is_dynamic=False, entity_id=0):
    return {"positions": np.ascontiguousarray(positions, f32).reshape(-1, 3),
            "indices": np.ascontiguousarray(indices, np.uint32).reshape(-1),
            "model": trs_model() if model is None else np.asarray(model, f32).reshape(16),
            "layer": layer, "mu_s": mu_s, "mu_k": mu_k, "flatten_ground": flatten_ground,
            "is_dynamic": is_dynamic, "entity_id": entity_id}


def load_mirror_fixture():
    """tests/golden/ornate_mirror.npz (made by tests/golden/make_mesh_fixture.py from the reference asset)."""
    z = np.load(os.path.join(_GOLDEN, "ornate_mirror.npz"))
    hulls = [(z[f"hull{i}_positions"], z[f"hull{i}_indices"].astype(np.uint32)) for i in range(int(z["n_hulls"]))]
    return {"name": str(z["name"]), "transform": rowmajor_to_colmajor(z["transform_rowmajor"]),
            "positions": z["positions"], "indices": z["indices"].astype(np.uint32), "hulls": hulls}


def load_asset_fixture(name):
    """tests/golden/{semla,cheese}.npz: stand-ins for the reference's missing Semla.static.json / 17-Cheese.static.json,
    regenerated from the FBX sources by tools/fbx_to_static_mesh.py (render mesh exact up to vertex order and a
    few quad diagonals; collision hulls are scipy convex-hull stand-ins)."""
    z = np.load(os.path.join(_GOLDEN, name + ".npz"))
    hulls = [(z[f"hull{i}_positions"], z[f"hull{i}_indices"].astype(np.uint32)) for i in range(int(z["n_hulls"]))]
    return {"name": str(z["name"]), "transform": rowmajor_to_colmajor(z["transform_rowmajor"]),
            "positions": z["positions"], "indices": z["indices"].astype(np.uint32), "hulls": hulls}


LAYER_SEMLA = 1 << 3


def semla_model(asset_transform):
    """Demo placement of Semla (DemoScene.swift:217-221): part transform + (18, 0, 10)."""
    t, q, s = transform_from_matrix(asset_transform)
    return trs_model(t + np.array([18, 0, 10], f32), q, s)


def cheese_model(asset_transform):
    """Demo placement of 17-Cheese (DemoScene.swift:167-168): the part transform as is."""
    t, q, s = transform_from_matrix(asset_transform)
    return trs_model(t, q, s)


def semla_scene(use_hulls=False, with_ground=True):
    """C2 world: Semla at its demo placement (layer 1<<3) + the ground plane."""
    a = load_asset_fixture("semla")
    parts = [ground_part(0)] if with_ground else []
    geoms = a["hulls"] if use_hulls else [(a["positions"], a["indices"])]
    for k, (v, i) in enumerate(geoms):
        parts.append(part(v, i, semla_model(a["transform"]), layer=LAYER_SEMLA, mu_s=0.6, mu_k=0.5, entity_id=len(parts)))
    return parts


def merged_scene(mirror_dynamic=True):
    """C5 world: 17-Cheese + Semla + ornate_mirror render meshes merged at their demo placements (~136k triangles)
    + the ground plane; the mirror is in the dynamic set so it can spin and be refitted every step."""
    parts = [ground_part(0)]
    c, sm, m = load_asset_fixture("cheese"), load_asset_fixture("semla"), load_mirror_fixture()
    parts.append(part(c["positions"], c["indices"], cheese_model(c["transform"]), mu_s=0.6, mu_k=0.5, entity_id=1))
    parts.append(part(sm["positions"], sm["indices"], semla_model(sm["transform"]), layer=LAYER_SEMLA, mu_s=0.6, mu_k=0.5,
                      entity_id=2))
    parts.append(part(m["positions"], m["indices"], mirror_model(m["transform"]), layer=LAYER_MIRROR, mu_s=0.6, mu_k=0.5,
                      is_dynamic=mirror_dynamic, entity_id=3))
    return parts


def ground_part(entity_id=0):
    v, i = plane_mesh(80.0)
    return part(v, i, trs_model(translation=(0, GROUND_Y, 0)), layer=LAYER_DEFAULT, mu_s=0.9, mu_k=0.8,
                entity_id=entity_id)


def mirror_model(asset_transform):
    """Demo placement of every ornate-mirror part (DemoScene.swift:330-336)."""
    t, q, s = transform_from_matrix(asset_transform)
    upright = quat_angle_axis(np.pi * 0.5, (1, 0, 0))
    flip = quat_angle_axis(np.pi, (1, 0, 0))
    q = quat_mul(q, quat_mul(upright, flip))
    s = s * f32(8.0)
    t = t + np.array([-10, 1, 4], f32)
    return trs_model(t, q, s)


def mirror_scene(use_hulls=True, mirror_dynamic=False, with_ground=True):
    """C3 world: the 80x80 ground plane + the ornate mirror at its demo placement.
    use_hulls=True  -> the 2 collision hulls (what the demo really collides with, DemoScene.swift:358-371)
    use_hulls=False -> the 14,246-triangle render mesh through the same API (collisionMesh ?? mesh, CQ:344)."""
    a = load_mirror_fixture()
    model = mirror_model(a["transform"])
    parts = []
    eid = 0
    if with_ground:
        parts.append(ground_part(eid))
        eid += 1
    geoms = a["hulls"] if use_hulls else [(a["positions"], a["indices"])]
    for v, i in geoms:
        parts.append(part(v, i, model, layer=LAYER_MIRROR, mu_s=0.6, mu_k=0.5, is_dynamic=mirror_dynamic,
                          entity_id=eid))
        eid += 1
    return parts


# ---------------------------------------------------------------- procedural terrain (C4)


def _hash2(ix, iz, seed):
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + iz.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         + np.uint64(seed)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(31)
    h = (h * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(29)
    return ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(f32)


def _value_noise(x, z, seed):
    ix, iz = np.floor(x).astype(np.int64), np.floor(z).astype(np.int64)
    fx, fz = (x - ix).astype(f32), (z - iz).astype(f32)
    sx, sz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)
    a, b = _hash2(ix, iz, seed), _hash2(ix + 1, iz, seed)
    c, d = _hash2(ix, iz + 1, seed), _hash2(ix + 1, iz + 1, seed)
    return (a + (b - a) * sx) * (1 - sz) + (c + (d - c) * sx) * sz


def terrain_height(x, z, seed=0xC0111DE4, amplitude=12.0, wavelength=256.0, octaves=4):
    x, z = np.asarray(x, np.float64), np.asarray(z, np.float64)
    h = np.zeros(x.shape, f32)
    amp, freq = 1.0, 1.0 / wavelength
    norm = 0.0
    for o in range(octaves):
        h += f32(amp) * (_value_noise(x * freq, z * freq, seed + o) - f32(0.5))
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return (h * f32(2.0 * amplitude / norm)).astype(f32)


def terrain_mesh(cells=2236, cell=2.0, seed=0xC0111DE4, amplitude=12.0):
    """Regular heightfield, 2 triangles per cell split along (i,j)-(i+1,j+1).
    Returns (positions (V,3) f32, indices (T*3,) u32, extent)."""
    n = cells + 1
    half = cells * cell * 0.5
    xs = (np.arange(n, dtype=np.float64) * cell - half)
    X, Z = np.meshgrid(xs, xs, indexing="xy")  # row j = z, col i = x
    Y = terrain_height(X, Z, seed, amplitude)
    pos = np.stack([X.astype(f32), Y, Z.astype(f32)], axis=-1).reshape(-1, 3)
    j, i = np.meshgrid(np.arange(cells, dtype=np.uint32), np.arange(cells, dtype=np.uint32), indexing="ij")
    v00 = j * n + i
    v10 = v00 + 1
    v01 = v00 + n
    v11 = v01 + 1
    tris = np.stack([v00, v11, v10, v00, v01, v11], axis=-1).reshape(-1)  # +Y facing
    return pos, tris.astype(np.uint32), half


def terrain_scene(cells=2236, cell=2.0, seed=0xC0111DE4):
    pos, idx, half = terrain_mesh(cells, cell, seed)
    return [part(pos, idx, entity_id=0)], half


def platform_record(positions, model, prev_translation, translation):
    """cq_platform of a kinematic platform: world AABB of its mesh under the current model matrix
    (meshWorldAABB, Systems.swift:627-642) and positionF - prevPositionF."""
    m = np.asarray(model, f32).reshape(4, 4).T
    p = np.asarray(positions, f32).reshape(-1, 3)
    w = ((m[:3, 0][None, :] * p[:, 0:1] + m[:3, 1][None, :] * p[:, 1:2]) + m[:3, 2][None, :] * p[:, 2:3]) + m[:3, 3][None, :]
    rec = np.zeros(1, np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("delta", "<f4", 3)]))
    rec["aabb_min"], rec["aabb_max"] = w.min(0), w.max(0)
    rec["delta"] = np.asarray(translation, f32) - np.asarray(prev_translation, f32)
    return rec


# ---------------------------------------------------------------- dtypes of the query records (include/cq.h)

RAY = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("max_distance", "<f4"), ("mask", "<u4")])
CAST = np.dtype([("from", "<f4", 3), ("delta", "<f4", 3), ("radius", "<f4"), ("half_height", "<f4"),
                 ("mask", "<u4"), ("min_normal_y", "<f4")])
CAPSULE = np.dtype([("from", "<f4", 3), ("radius", "<f4"), ("half_height", "<f4"), ("mask", "<u4")])


def scene_aabb(parts):
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for p in parts:
        m = np.asarray(p["model"], np.float64).reshape(4, 4).T
        w = p["positions"].astype(np.float64) @ m[:3, :3].T + m[:3, 3]
        lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
    return lo, hi


def _unit_vectors(rng, n):
    g = rng.standard_normal((n, 3))
    return (g / np.linalg.norm(g, axis=1, keepdims=True)).astype(f32)


def gen_casts(n, lo, hi, seed, radius=1.5, half_height=1.0, len_range=(0.05, 2.0), expand=3.0, mask=0xFFFFFFFF,
              min_normal_y=0.5):
    """C2-style sweeps: from ~ U(AABB expanded), direction uniform on the sphere, |delta| ~ U[len_range]."""
    rng = np.random.default_rng(seed)
    q = np.zeros(n, CAST)
    q["from"] = rng.uniform(np.asarray(lo) - expand, np.asarray(hi) + expand, (n, 3)).astype(f32)
    q["delta"] = _unit_vectors(rng, n) * rng.uniform(len_range[0], len_range[1], (n, 1)).astype(f32)
    q["radius"], q["half_height"], q["mask"], q["min_normal_y"] = radius, half_height, mask, min_normal_y
    return q


def gen_rays(n, lo, hi, seed, max_distance=100.0, expand=5.0, y_range=None, mask=0xFFFFFFFF):
    """C5-style rays: origins ~ U(AABB expanded), unit directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    r = np.zeros(n, RAY)
    lo2, hi2 = np.asarray(lo, np.float64) - expand, np.asarray(hi, np.float64) + expand
    if y_range is not None:
        lo2[1], hi2[1] = y_range
    r["origin"] = rng.uniform(lo2, hi2, (n, 3)).astype(f32)
    r["direction"] = _unit_vectors(rng, n)
    r["max_distance"], r["mask"] = max_distance, mask
    return r


def gen_capsules(n, lo, hi, seed, radius=1.5, half_height=1.0, expand=1.0, mask=0xFFFFFFFF):
    rng = np.random.default_rng(seed)
    c = np.zeros(n, CAPSULE)
    c["from"] = rng.uniform(np.asarray(lo) - expand, np.asarray(hi) + expand, (n, 3)).astype(f32)
    c["radius"], c["half_height"], c["mask"] = radius, half_height, mask
    return c


def gen_c3_characters(n, seed=0xC0111DE3, radius=1.5, half_height=1.0):
    """C3 characters around the mirror: x~U[-14,-6], z~U[0.5,8.5],
    y = -3 + r + hh + U[0,0.5]; horizontal velocity, heading U[0,2pi), speed U[0,12.5].
    Returns (positions (n,3) f32, velocities (n,3) f32)."""
    rng = np.random.default_rng(seed)
    pos = np.empty((n, 3), f32)
    pos[:, 0] = rng.uniform(-14.0, -6.0, n)
    pos[:, 2] = rng.uniform(0.5, 8.5, n)
    pos[:, 1] = GROUND_Y + radius + half_height + rng.uniform(0.0, 0.5, n)
    heading = rng.uniform(0.0, 2 * np.pi, n)
    speed = rng.uniform(0.0, 12.5, n)
    vel = np.stack([np.cos(heading) * speed, np.zeros(n), np.sin(heading) * speed], axis=1).astype(f32)
    return pos, vel


def gen_c4_casts(n, half, seed=0xC0111DE4, radius=0.4, half_height=0.5, terrain_seed=0xC0111DE4, border=4.0):
    """C4 sweeps over the terrain: walkers just above the surface, blocking mode."""
    rng = np.random.default_rng(seed)
    q = np.zeros(n, CAST)
    x = rng.uniform(-half + border, half - border, n)
    z = rng.uniform(-half + border, half - border, n)
    y = terrain_height(x, z, terrain_seed) + f32(radius + half_height) + rng.uniform(0.02, 0.5, n).astype(f32)
    q["from"] = np.stack([x.astype(f32), y.astype(f32), z.astype(f32)], axis=1)
    heading = rng.uniform(0.0, 2 * np.pi, n)
    step = rng.uniform(0.05, 0.5, n)
    q["delta"] = np.stack([np.cos(heading) * step, -rng.uniform(0.0, 0.3, n), np.sin(heading) * step],
                          axis=1).astype(f32)
    q["radius"], q["half_height"], q["mask"], q["min_normal_y"] = radius, half_height, 0xFFFFFFFF, 0.5
    return q


# ---------------------------------------------------------------- C1: single character, 600 frames


def wedge_mesh(width=8.0, depth=10.0, height=4.0):
    """A ramp like the demo's "Test Ramp" (DemoScene.swift:588-590; ProceduralMeshes.ramp): low edge at +z,
    high edge at -z, centred on the origin.  8 triangles: slope, bottom, back wall, two sides."""
    w, d, h = f32(width) * f32(0.5), f32(depth) * f32(0.5), f32(height) * f32(0.5)
    v = np.array([[-w, -h, d], [w, -h, d], [-w, -h, -d], [w, -h, -d], [-w, h, -d], [w, h, -d]], f32)
    idx = [0, 1, 5, 0, 5, 4,      # slope
           0, 2, 3, 0, 3, 1,      # bottom
           2, 4, 5, 2, 5, 3,      # back wall
           0, 4, 2,               # left side
           1, 3, 5]               # right side
    return v, np.array(idx, np.uint32)


def c1_scene():
    """Static demo world restated from DemoScene.swift: ground plane 80x80 @ y=-3, wall
    box 6 @ (0,0,-10), ramp 8x10x4 @ (8,-1,0) flattenGround muS .35, step box 2 @ (-6,-2,4), and the ornate
    mirror hulls at their demo placement.  (Dome, 17-Cheese and Semla are left out: their assets are
    missing blobs / not needed to exercise every controller branch.)"""
    parts = [ground_part(0)]
    bv, bi = box_mesh(6.0)
    parts.append(part(bv, bi, trs_model((0, 0, -10)), entity_id=1))
    rv, ri = wedge_mesh(8.0, 10.0, 4.0)
    parts.append(part(rv, ri, trs_model((8, GROUND_Y + 2.0, 0)), mu_s=0.35, mu_k=0.25, flatten_ground=True,
                      entity_id=2))
    sv, si = box_mesh(2.0)
    parts.append(part(sv, si, trs_model((-6, -2, 4)), entity_id=3))
    a = load_mirror_fixture()
    model = mirror_model(a["transform"])
    for k, (v, i) in enumerate(a["hulls"]):
        parts.append(part(v, i, model, layer=LAYER_MIRROR, mu_s=0.6, mu_k=0.5, entity_id=4 + k))
    return parts


def c1_apply_intent(states, frame, n_frames=600, speed=12.5, accel=20.0, decel=36.0, dt=1.0 / 60.0):
    """PhysicsIntentSystem rule for a character-controller body (Systems.swift:228-233 + approachVecD :419-426):
    horizontal velocity approaches the desired velocity by at most accel*dt; scripted heading
    yaw = 2*pi*frame/n_frames at `speed`.  Operates in place on a STATE record array (float64 velocity)."""
    yaw = 2.0 * np.pi * frame / n_frames
    v = states["velocity"]
    spd = np.broadcast_to(np.asarray(speed, np.float64), (len(v),))
    tgt = np.stack([np.float32(np.cos(yaw) * spd).astype(np.float64), np.zeros(len(v)),
                    np.float32(np.sin(yaw) * spd).astype(np.float64)], axis=1)
    cur = np.stack([v[:, 0], np.zeros(len(v)), v[:, 2]], axis=1)
    acc = np.where(np.linalg.norm(tgt, axis=1) >= np.linalg.norm(cur, axis=1), np.float32(accel), np.float32(decel))
    max_delta = acc.astype(np.float64) * np.float64(np.float32(dt))
    delta = tgt - cur
    ln = np.sqrt((delta[:, 0] * delta[:, 0] + delta[:, 1] * delta[:, 1]) + delta[:, 2] * delta[:, 2])
    reach = (ln <= max_delta) | (ln < 0.00001)
    safe = np.where(ln > 0, ln, 1.0)
    nxt = np.where(reach[:, None], tgt, cur + delta / safe[:, None] * max_delta[:, None])
    v[:, 0] = nxt[:, 0]
    v[:, 2] = nxt[:, 2]


C1_STARTS = np.array([[0, 7.5, 0], [-18.5, 1.5, -5.05], [-13.16, 1.5, -3.16], [20.7, 1.5, -12.7]], f32)
C1_SPEEDS = np.array([12.5, 6.0, 4.5, 8.0])


def c1_run(step_fn, states, n_frames=600, speeds=12.5):
    """Drive `n_frames` fixed steps: intent -> step_fn(states) (gravity + move-and-slide).  Returns the
    per-frame record (position f64x3, grounded, grounded_near, ground_triangle_index, ground_normal)."""
    rec = np.zeros((n_frames, len(states)), dtype=[("position", "<f8", 3), ("grounded", "u1"), ("grounded_near", "u1"),
                                                   ("ground_triangle_index", "<i4"), ("ground_normal", "<f4", 3)])
    for f in range(n_frames):
        c1_apply_intent(states, f, n_frames, speeds)
        step_fn(states)
        for k in rec.dtype.names:
            rec[k][f] = states[k]
    return rec
