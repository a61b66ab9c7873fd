"""ctypes binding of the CPU oracle (oracle/libcq_oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the product
package.  Record layouts are declared here independently of the product's.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcq_oracle.so")

RAY = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("max_distance", "<f4"), ("mask", "<u4")])
RAY_HIT = np.dtype([("distance", "<f4"), ("position", "<f4", 3), ("normal", "<f4", 3), ("triangle_index", "<i4")])
CAST = np.dtype([("from", "<f4", 3), ("delta", "<f4", 3), ("radius", "<f4"), ("half_height", "<f4"),
                 ("mask", "<u4"), ("min_normal_y", "<f4")])
CAST_HIT = np.dtype([("toi", "<f4"), ("position", "<f4", 3), ("normal", "<f4", 3),
                     ("triangle_normal", "<f4", 3), ("triangle_index", "<i4")])
CAPSULE = np.dtype([("from", "<f4", 3), ("radius", "<f4"), ("half_height", "<f4"), ("mask", "<u4")])
OVERLAP_HIT = np.dtype([("depth", "<f4"), ("position", "<f4", 3), ("normal", "<f4", 3),
                        ("triangle_normal", "<f4", 3), ("triangle_index", "<i4")])
PARAMS = np.dtype([("radius", "<f4"), ("half_height", "<f4"), ("skin_width", "<f4"), ("ground_snap_skin", "<f4"),
                   ("snap_distance", "<f4"), ("fall_probe_distance", "<f4"), ("ground_snap_max_speed", "<f4"),
                   ("ground_snap_max_toi", "<f4"), ("ground_snap_max_step", "<f4"),
                   ("ground_sweep_max_step", "<f4"), ("max_slide_iterations", "<i4"), ("min_ground_dot", "<f4"),
                   ("collision_mask", "<u4")])
STATE = np.dtype([("position", "<f8", 3), ("velocity", "<f8", 3), ("ground_normal", "<f4", 3),
                  ("ground_distance", "<f4"), ("side_contact_normal", "<f4", 3), ("ground_triangle_index", "<i4"),
                  ("ground_transition_frames", "<i4"), ("side_contact_frames", "<i4"), ("manifold_frames", "<i4"),
                  ("manifold_count", "<i4"), ("manifold_triangles", "<i4", 4), ("manifold_normals", "<f4", (4, 3)),
                  ("grounded", "u1"), ("grounded_near", "u1"), ("ground_sliding", "u1"), ("_pad", "u1", 5)])
assert STATE.itemsize == 168 and CAST.itemsize == 40 and CAST_HIT.itemsize == 44 and RAY.itemsize == 32
assert RAY_HIT.itemsize == 32 and CAPSULE.itemsize == 24 and PARAMS.itemsize == 52

PLATFORM = np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("delta", "<f4", 3)])
ORDER_REFERENCE, ORDER_CANONICAL = 0, 1


class _Part(C.Structure):
    _fields_ = [("positions_xyz", C.c_void_p), ("indices", C.c_void_p), ("n_verts", C.c_int32),
                ("n_indices", C.c_int32), ("model", C.c_float * 16), ("layer", C.c_uint32), ("mu_s", C.c_float),
                ("mu_k", C.c_float), ("flatten_ground", C.c_uint8), ("is_dynamic", C.c_uint8), ("_pad", C.c_uint16),
                ("entity_id", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("candidates", "sweep_tests", "sweep_iterations", "max_iterations",
                                         "distance_evals", "nodes_visited", "ties", "overflows")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "cq_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_world_create.restype = C.c_void_p
        L.orc_world_create.argtypes = [C.c_void_p, C.c_int32]
        L.orc_world_create_ex.restype = C.c_void_p
        L.orc_world_create_ex.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_world_triangle_material.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_world_destroy.argtypes = [C.c_void_p]
        L.orc_world_counts.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_world_read_soup.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
        L.orc_world_update_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_world_check_bvh.restype = C.c_int32
        L.orc_world_check_bvh.argtypes = [C.c_void_p, C.c_int32]
        L.orc_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_capsule_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p]
        L.orc_capsule_overlap.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                          C.c_void_p]
        L.orc_capsule_overlap_all.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_move_and_slide.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_void_p,
                                         C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_move_and_slide_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_void_p,
                                            C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_segment_triangle_distance.restype = C.c_float
        L.orc_segment_triangle_distance.argtypes = [C.c_void_p, C.c_float] + [C.c_void_p] * 5
        L.orc_segment_triangle_distance_batch.argtypes = [C.c_int32] + [C.c_void_p] * 6
        L.orc_ray_triangle_batch.argtypes = [C.c_int32] + [C.c_void_p] * 5
        L.orc_capsule_capsule_sweep_batch.argtypes = [C.c_int32] + [C.c_void_p] * 8
        L.orc_agent_separation.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                                           C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_closest_point_on_triangle.restype = C.c_float
        L.orc_closest_point_on_triangle.argtypes = [C.c_void_p] * 5
        L.orc_segment_segment_distance_sq.restype = C.c_float
        L.orc_segment_segment_distance_sq.argtypes = [C.c_void_p] * 6
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def default_params(**over):
    """CharacterControllerComponent defaults (Components.swift:380-404)."""
    p = np.zeros(1, PARAMS)
    p["radius"], p["half_height"], p["skin_width"], p["ground_snap_skin"] = 1.5, 1.0, 0.3, 0.05
    p["snap_distance"], p["fall_probe_distance"], p["ground_snap_max_speed"] = 0.8, 200.0, 5.0
    p["ground_snap_max_toi"], p["ground_snap_max_step"], p["ground_sweep_max_step"] = 0.1, 0.1, 0.1
    p["max_slide_iterations"], p["min_ground_dot"], p["collision_mask"] = 4, 0.5, 0xFFFFFFFF
    for k, v in over.items():
        p[k] = v
    return p


def init_states(positions, velocities=None):
    """PhysicsBodyComponent.init + CharacterControllerComponent.init state defaults."""
    positions = np.asarray(positions, np.float32).reshape(-1, 3)
    n = positions.shape[0]
    s = np.zeros(n, STATE)
    s["position"] = positions.astype(np.float64)
    if velocities is not None:
        s["velocity"] = np.asarray(velocities, np.float32).reshape(-1, 3).astype(np.float64)
    s["ground_normal"] = (0, 1, 0)
    s["ground_distance"] = np.finfo(np.float32).max
    s["ground_triangle_index"] = -1
    return s


_SURFACE_MATERIAL = np.dtype([("mu_s", "<f4"), ("mu_k", "<f4"), ("flatten_ground", "u1"), ("_pad", "u1", (3,))])


class _TriangleMaterials(C.Structure):
    _fields_ = [("entity_id", C.c_uint32), ("n", C.c_int32), ("materials", C.c_void_p)]


class OracleWorld:
    """parts: list of dicts {positions (V,3) f32, indices (I,) u32, model (16,) f32 column-major, layer,
    mu_s, mu_k, flatten_ground, is_dynamic, entity_id}."""

    def __init__(self, parts):
        self._keep = []
        arr = (_Part * len(parts))()
        for i, p in enumerate(parts):
            pos = np.ascontiguousarray(p["positions"], np.float32).reshape(-1, 3)
            idx = np.ascontiguousarray(p["indices"], np.uint32).reshape(-1)
            self._keep += [pos, idx]
            arr[i].positions_xyz = pos.ctypes.data
            arr[i].indices = idx.ctypes.data
            arr[i].n_verts = pos.shape[0]
            arr[i].n_indices = idx.shape[0]
            m = np.asarray(p["model"], np.float32).reshape(16)
            for k in range(16):
                arr[i].model[k] = float(m[k])
            arr[i].layer = int(p.get("layer", 1))
            arr[i].mu_s = float(p.get("mu_s", 0.8))
            arr[i].mu_k = float(p.get("mu_k", 0.6))
            arr[i].flatten_ground = int(bool(p.get("flatten_ground", False)))
            arr[i].is_dynamic = int(bool(p.get("is_dynamic", False)))
            arr[i].entity_id = int(p.get("entity_id", i))
        per_tri = []
        for i, p in enumerate(parts):  # StaticMeshComponent.triangleMaterials: (T, 3) rows of mu_s, mu_k, flatten_ground
            if p.get("triangle_materials") is not None:
                rows = np.asarray(p["triangle_materials"], np.float32).reshape(-1, 3)
                mats = np.zeros(len(rows), _SURFACE_MATERIAL)
                mats["mu_s"], mats["mu_k"], mats["flatten_ground"] = rows[:, 0], rows[:, 1], rows[:, 2] != 0
                per_tri.append((int(p.get("entity_id", i)), mats))
        tm = (_TriangleMaterials * max(len(per_tri), 1))()
        for k, (eid, mats) in enumerate(per_tri):
            self._keep.append(mats)
            tm[k].entity_id, tm[k].n, tm[k].materials = eid, len(mats), mats.ctypes.data
        self._h = lib().orc_world_create_ex(C.byref(arr), len(parts), C.byref(tm), len(per_tri))

    def close(self):
        if self._h:
            lib().orc_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def triangle_material(self, triangle_index):
        out = np.zeros(3, np.float32)
        lib().orc_world_triangle_material(self._h, int(triangle_index), out.ctypes.data)
        return {"mu_s": float(out[0]), "mu_k": float(out[1]), "flatten_ground": bool(out[2])}

    def counts(self, which=0):
        out = np.zeros(3, np.int32)
        lib().orc_world_counts(self._h, which, _ptr(out))
        return {"vertices": int(out[0]), "triangles": int(out[1]), "nodes": int(out[2])}

    def read_soup(self, which=0):
        c = self.counts(which)
        pos = np.zeros((c["vertices"], 3), np.float32)
        idx = np.zeros((c["triangles"], 3), np.uint32)
        aabb = np.zeros((c["triangles"], 6), np.float32)
        lay = np.zeros(c["triangles"], np.uint32)
        par = np.zeros(c["triangles"], np.int32)
        lib().orc_world_read_soup(self._h, which, _ptr(pos), _ptr(idx), _ptr(aabb), _ptr(lay), _ptr(par))
        return {"positions": pos, "indices": idx, "aabbs": aabb, "layers": lay, "parts": par}

    def update_transforms(self, entity_ids, models):
        ids = np.ascontiguousarray(entity_ids, np.uint32)
        m = np.ascontiguousarray(models, np.float32).reshape(-1, 16)
        lib().orc_world_update_transforms(self._h, _ptr(ids), _ptr(m), len(ids))

    def check_bvh(self, which=0):
        return bool(lib().orc_world_check_bvh(self._h, which))

    def raycast(self, rays, order=ORDER_CANONICAL, n_threads=1, stats=None):
        rays = np.ascontiguousarray(rays, RAY)
        out = np.zeros(len(rays), RAY_HIT)
        lib().orc_raycast(self._h, _ptr(rays), len(rays), _ptr(out), order, n_threads,
                          C.byref(stats) if stats is not None else None)
        return out

    def capsule_cast(self, q, mode=0, order=ORDER_CANONICAL, n_threads=1, stats=None):
        q = np.ascontiguousarray(q, CAST)
        out = np.zeros(len(q), CAST_HIT)
        lib().orc_capsule_cast(self._h, _ptr(q), len(q), mode, _ptr(out), order, n_threads,
                               C.byref(stats) if stats is not None else None)
        return out

    def capsule_overlap(self, q, order=ORDER_CANONICAL, n_threads=1, stats=None):
        q = np.ascontiguousarray(q, CAPSULE)
        out = np.zeros(len(q), OVERLAP_HIT)
        lib().orc_capsule_overlap(self._h, _ptr(q), len(q), _ptr(out), order, n_threads,
                                  C.byref(stats) if stats is not None else None)
        return out

    def capsule_overlap_all(self, q, max_hits=8, order=ORDER_CANONICAL, n_threads=1, stats=None):
        q = np.ascontiguousarray(q, CAPSULE)
        out = np.zeros((len(q), max_hits), OVERLAP_HIT)
        counts = np.zeros(len(q), np.int32)
        overflow = np.zeros(len(q), np.uint8)
        lib().orc_capsule_overlap_all(self._h, _ptr(q), len(q), max_hits, _ptr(out), _ptr(counts), _ptr(overflow),
                                      order, n_threads, C.byref(stats) if stats is not None else None)
        return out, counts, overflow

    def move_and_slide(self, states, params, dt=1.0 / 60.0, gravity=(0.0, -98.0, 0.0), flags=1,
                       order=ORDER_CANONICAL, n_threads=1, stats=None, platforms=None):
        """In place on `states` (a STATE array); returns it."""
        assert states.dtype == STATE and states.flags["C_CONTIGUOUS"]
        params = np.ascontiguousarray(params, PARAMS)
        g = np.asarray(gravity, np.float32)
        pl = np.ascontiguousarray(platforms if platforms is not None else np.zeros(0, PLATFORM), PLATFORM)
        lib().orc_move_and_slide_ex(self._h, _ptr(states), len(states), _ptr(params), np.float32(dt), _ptr(g), flags,
                                    order, n_threads, C.byref(stats) if stats is not None else None, _ptr(pl), len(pl))
        return states

    def agent_separation(self, states, params, mass_weight=None, iterations=2, separation_margin=0.2, height_margin=0.1,
                         use_query=True, order=ORDER_CANONICAL, n_threads=1):
        """AgentSeparationSystem.fixedUpdate over the batch, in place; returns the number of resolved pairs."""
        assert states.dtype == STATE and states.flags["C_CONTIGUOUS"]
        params = np.ascontiguousarray(params, PARAMS)
        mw = None if mass_weight is None else np.ascontiguousarray(mass_weight, np.float32)
        assert mw is None or len(mw) == len(states)
        pairs = C.c_int64(0)
        lib().orc_agent_separation(self._h, _ptr(states), len(states), _ptr(params), None if mw is None else _ptr(mw),
                                   iterations, separation_margin, height_margin, 1 if use_query else 0, order, n_threads,
                                   C.byref(pairs))
        return pairs.value


def capsule_capsule_sweep_batch(frm, delta, other_pos, other_delta, dims):
    """capsuleCapsuleSweep over n packed cases; dims (n,4) = radius, halfHeight, otherRadius, otherHalfHeight."""
    a = [np.ascontiguousarray(x, np.float32) for x in (frm, delta, other_pos, other_delta, dims)]
    n = len(a[0])
    hit, toi, normal = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    lib().orc_capsule_capsule_sweep_batch(n, *[_ptr(x) for x in a], _ptr(hit), _ptr(toi), _ptr(normal))
    return hit, toi, normal


def segment_triangle_distance(center, half_height, v0, v1, v2):
    a = [np.asarray(x, np.float32) for x in (center, v0, v1, v2)]
    seg, tri = np.zeros(3, np.float32), np.zeros(3, np.float32)
    d = lib().orc_segment_triangle_distance(_ptr(a[0]), np.float32(half_height), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]),
                                            _ptr(seg), _ptr(tri))
    return float(d), seg, tri


def closest_point_on_triangle(p, a, b, c):
    v = [np.asarray(x, np.float32) for x in (p, a, b, c)]
    out = np.zeros(3, np.float32)
    d = lib().orc_closest_point_on_triangle(_ptr(v[0]), _ptr(v[1]), _ptr(v[2]), _ptr(v[3]), _ptr(out))
    return float(d), out


def segment_segment_distance_sq(p1, q1, p2, q2):
    v = [np.asarray(x, np.float32) for x in (p1, q1, p2, q2)]
    c1, c2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    d = lib().orc_segment_segment_distance_sq(_ptr(v[0]), _ptr(v[1]), _ptr(v[2]), _ptr(v[3]), _ptr(c1), _ptr(c2))
    return float(d), c1, c2


def segment_triangle_distance_batch(centers, hh, tris):
    centers = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
    n = len(centers)
    hh = np.ascontiguousarray(np.broadcast_to(np.asarray(hh, np.float32), (n,)))
    tris = np.ascontiguousarray(tris, np.float32).reshape(n, 9)
    dist, seg, tri = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    lib().orc_segment_triangle_distance_batch(n, _ptr(centers), _ptr(hh), _ptr(tris), _ptr(dist), _ptr(seg), _ptr(tri))
    return dist, seg, tri


def ray_triangle_batch(origins, dirs, tris):
    origins = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    n = len(origins)
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(n, 3)
    tris = np.ascontiguousarray(tris, np.float32).reshape(n, 9)
    t, hit = np.zeros(n, np.float32), np.zeros(n, np.int32)
    lib().orc_ray_triangle_batch(n, _ptr(origins), _ptr(dirs), _ptr(tris), _ptr(t), _ptr(hit))
    return t, hit
