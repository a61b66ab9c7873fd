"""The DEVICE narrow phase (swift-game-engine_b200/csrc/cq_math.cuh, branch-free formulation) compiled for the
host and compared BIT-EXACTLY with the oracle's literal restatement of CollisionQuery.swift:1396-1601 on a few
million random and adversarial (touching / piercing / degenerate) inputs.  Runs without a GPU."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostmath", "hostmath.cu")
LIB = os.path.join(HERE, "hostmath", "libcq_hostmath.so")
HDR = os.path.join(os.path.dirname(HERE), "swift-game-engine_b200", "csrc", "cq_math.cuh")


@pytest.fixture(scope="module")
def hm():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    if not os.path.exists(LIB) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(LIB):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC",
                               "-Xcompiler", "-ffp-contract=off", "-shared", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.hm_segment_triangle_distance_batch.argtypes = [C.c_int] + [C.c_void_p] * 6
    L.hm_ray_triangle_batch.argtypes = [C.c_int] + [C.c_void_p] * 5
    L.hm_capsule_capsule_sweep_batch.argtypes = [C.c_int] + [C.c_void_p] * 8
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cases(rng, n):
    scale = 10.0 ** rng.uniform(-2.5, 1.5, (n, 1)).astype(np.float32)
    tris = (rng.standard_normal((n, 9)).astype(np.float32) * scale)
    off = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    tris = (tris.reshape(n, 3, 3) + off[:, None, :]).reshape(n, 9)
    hh = rng.choice(np.float32([0.0, 1e-4, 0.5, 1.0, 3.0]), n)
    kind = rng.integers(0, 5, n)
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    on_tri = (tris.reshape(n, 3, 3) * w[:, :, None]).sum(1)
    centers = np.where((kind == 0)[:, None], on_tri + rng.standard_normal((n, 3)).astype(np.float32) * scale * 3, on_tri)
    centers = np.where((kind == 1)[:, None], on_tri + np.float32([0, 1, 0]) * (hh[:, None] + np.float32(1.5)), centers)  # resting
    centers = np.where((kind == 2)[:, None], tris[:, 0:3] + rng.standard_normal((n, 3)).astype(np.float32) * 1e-3, centers)
    centers = np.where((kind == 3)[:, None], on_tri + rng.standard_normal((n, 3)).astype(np.float32) * 0.3, centers)  # piercing
    # kind 4: exactly on the triangle
    deg = rng.random(n) < 0.05  # degenerate edges (two equal vertices / tiny edge)
    tris[deg, 3:6] = tris[deg, 0:3] + (rng.standard_normal((deg.sum(), 3)) * 1e-4).astype(np.float32)
    same = rng.random(n) < 0.01
    tris[same, 6:9] = tris[same, 3:6]
    return centers.astype(np.float32), hh.astype(np.float32), np.ascontiguousarray(tris)


def test_segment_triangle_distance_bit_exact(hm, orc):
    rng = np.random.default_rng(1234)
    n = 1_500_000
    centers, hh, tris = _cases(rng, n)
    od, oseg, otri = orc.segment_triangle_distance_batch(centers, hh, tris)
    d, seg, tri = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    hm.hm_segment_triangle_distance_batch(n, _p(centers), _p(hh), _p(tris), _p(d), _p(seg), _p(tri))
    fin = np.isfinite(od)
    assert fin.mean() > 0.99
    assert np.array_equal(d[fin], od[fin])
    assert np.array_equal(np.isnan(d), np.isnan(od))
    assert np.array_equal(seg[fin], oseg[fin]) and np.array_equal(tri[fin], otri[fin])
    assert (od == 0).sum() > 1000 and (od > 0).sum() > 1000  # both the pierced and the separated paths were hit


def test_segment_triangle_distance_bit_exact_from_denormals_to_1e9(hm, orc):
    """Domain of the bit-exactness claim: scenes of every magnitude from float32 denormals (1e-44) to 1e9 — distances and
    both contact points equal bit for bit (no flush-to-zero on either side: nvcc -ftz=false, and the same holds for the
    host build checked here).  Beyond ~1e10 the squared quantities of the distance function overflow float32 (cross
    products squared reach 1e40): the distance still agrees, but the branch-free formulation — which evaluates every
    Voronoi region and selects — may then pick another, equally meaningless contact point; worlds that large are outside
    the contract (a 10 M-triangle terrain spans 4.5e3 m)."""
    rng = np.random.default_rng(3)
    n = 600_000
    ex = rng.uniform(-44, 19, (n, 1))
    scale = (10.0 ** ex).astype(np.float32)
    tris = np.ascontiguousarray(rng.standard_normal((n, 9)).astype(np.float32) * scale)
    centers = np.ascontiguousarray(rng.standard_normal((n, 3)).astype(np.float32) * scale)
    hh = (rng.uniform(0, 2, n).astype(np.float32) * scale[:, 0]).astype(np.float32)
    od, oseg, otri = orc.segment_triangle_distance_batch(centers, hh, tris)
    d, seg, tri = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    hm.hm_segment_triangle_distance_batch(n, _p(centers), _p(hh), _p(tris), _p(d), _p(seg), _p(tri))

    def same(a, b):
        return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same(d, od).all()  # the distance agrees at every magnitude, overflow included
    inside = ex[:, 0] <= 9.0
    assert inside.sum() > 400_000 and (ex[:, 0] < -38).sum() > 30_000  # the denormal range is well represented
    assert same(seg, oseg)[inside].all() and same(tri, otri)[inside].all()
    assert np.isfinite(od[inside]).all()


def test_ray_triangle_bit_exact(hm, orc):
    rng = np.random.default_rng(99)
    n = 500_000
    _, _, tris = _cases(rng, n)
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    target = (tris.reshape(n, 3, 3) * w[:, :, None]).sum(1)
    origins = target + rng.standard_normal((n, 3)).astype(np.float32) * 5
    dirs = (target - origins) * rng.uniform(0.2, 2.0, (n, 1)).astype(np.float32)
    dirs[: n // 4] = rng.standard_normal((n // 4, 3)).astype(np.float32)
    ot, ohit = orc.ray_triangle_batch(origins, dirs, tris)
    t, hit = np.zeros(n, np.float32), np.zeros(n, np.int32)
    hm.hm_ray_triangle_batch(n, _p(np.ascontiguousarray(origins)), _p(np.ascontiguousarray(dirs)), _p(tris), _p(t), _p(hit))
    assert np.array_equal(hit, ohit) and np.array_equal(t[hit == 1], ot[ohit == 1])
    assert 0.3 < hit.mean() < 0.95


def test_capsule_capsule_sweep_bit_exact(hm, orc):
    """capsule_pair_sweep (device source) vs the oracle's capsuleCapsuleSweep (Systems.swift:1505-1590): approaching,
    receding, resting (|relative motion| < 1e-6), vertically stacked and purely vertical relative motion."""
    rng = np.random.default_rng(77)
    n = 1_000_000
    dims = np.stack([rng.choice(np.float32([0.3, 1.5]), n), rng.choice(np.float32([0.0, 0.6, 1.0]), n),
                     rng.choice(np.float32([0.3, 1.5]), n), rng.choice(np.float32([0.0, 0.6, 1.0]), n)], axis=1).astype(np.float32)
    frm = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    other = (frm + rng.standard_normal((n, 3)) * [3, 3, 3]).astype(np.float32)
    delta = (rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-3, 0.7, (n, 1))).astype(np.float32)
    odelta = (rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-3, 0.7, (n, 1))).astype(np.float32)
    kind = rng.integers(0, 6, n)
    odelta[kind == 0] = delta[kind == 0]                      # no relative motion -> overlap test only
    delta[kind == 1, 1] = 0                                   # horizontal relative motion (|vy| < 1e-6 branch)
    odelta[kind == 1, 1] = 0
    other[kind == 2] = frm[kind == 2] + np.float32([0, 1, 0]) * rng.uniform(0, 6, ((kind == 2).sum(), 1)).astype(np.float32)
    aim = kind == 3                                           # aimed straight at the other capsule
    delta[aim] = ((other[aim] - frm[aim]) * rng.uniform(0.2, 1.5, (aim.sum(), 1))).astype(np.float32)
    odelta[kind == 4] = 0                                     # static obstacle
    delta[kind == 5, 0] = 0                                   # purely vertical relative motion
    delta[kind == 5, 2] = 0
    odelta[kind == 5, 0] = 0
    odelta[kind == 5, 2] = 0
    oh, ot, on = orc.capsule_capsule_sweep_batch(frm, delta, other, odelta, dims)
    hit, toi, normal = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    hm.hm_capsule_capsule_sweep_batch(n, _p(frm), _p(delta), _p(other), _p(odelta), _p(dims), _p(hit), _p(toi), _p(normal))
    assert 0.1 < oh.mean() < 0.9
    assert np.array_equal(oh, hit)
    assert ot.tobytes() == toi.tobytes() and on.tobytes() == normal.tobytes()


# ---------------------------------------------------------------- host half of the mesh upload (csrc/cq_assemble.h)
ASM_SRC = os.path.join(HERE, "hostmath", "assemble.cpp")
ASM_LIB = os.path.join(HERE, "hostmath", "libcq_assemble.so")
ASM_HDR = os.path.join(os.path.dirname(HERE), "swift-game-engine_b200", "csrc", "cq_assemble.h")


@pytest.fixture(scope="module")
def asm():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    if not os.path.exists(ASM_LIB) or max(os.path.getmtime(ASM_SRC), os.path.getmtime(ASM_HDR)) > os.path.getmtime(ASM_LIB):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", ASM_LIB, ASM_SRC])
    L = C.CDLL(ASM_LIB)
    L.asm_plan.restype = C.c_void_p
    L.asm_plan.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
    L.asm_execute.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
    L.asm_placement.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.asm_rows_of.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.asm_free.argtypes = [C.c_void_p]
    return L


def _mesh_parts(cq, parts, keep):
    arr = (cq.MeshPart * max(len(parts), 1))()
    for i, p in enumerate(parts):
        pos = np.ascontiguousarray(p["positions"], np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(p["indices"], np.uint32).reshape(-1)
        keep += [pos, idx]
        arr[i].positions_xyz, arr[i].indices = pos.ctypes.data, idx.ctypes.data
        arr[i].n_verts, arr[i].n_indices = pos.shape[0], idx.shape[0]
        arr[i].layer, arr[i].is_dynamic = int(p["layer"]), int(bool(p["is_dynamic"]))
    return arr


def _upload(L, cq, parts, stage_limit):
    """Plan (cq_assemble.h) -> copies executed by the harness -> the two expansion kernels of cq_build.cu restated in
    numpy (k_expand_verts: xyz + part tag; k_expand_tris: part-local -> set-global indices, layer and part per triangle)."""
    keep = []
    arr = _mesh_parts(cq, parts, keep)
    sizes, err = np.zeros(8, np.int64), np.zeros(3, np.int64)
    h = L.asm_plan(C.byref(arr), len(parts), stage_limit, _p(sizes), _p(err))
    out, direct_total = [], 0
    for s in range(2):
        nv, nt, nrows, _ = (int(x) for x in sizes[4 * s:4 * s + 4])
        raw_pos, raw_idx = np.full(3 * nv, np.nan, np.float32), np.full(3 * nt, 0xDEADBEEF, np.uint32)
        rows, starts, of_set = np.zeros((nrows, 8), np.int32), np.zeros(nrows + 1, np.int32), np.zeros(max(nrows, 1), np.int32)
        covered, direct = np.zeros(nv + nt, np.int32), np.zeros(1, np.int32)
        L.asm_execute(h, s, _p(raw_pos), _p(raw_idx), _p(rows), _p(starts), _p(of_set), _p(covered), _p(direct))
        assert (covered == 1).all(), "every vertex and triangle is uploaded exactly once"
        direct_total += int(direct[0])
        # part_row_of against numpy's upper bound, on every key (small sets) or around every boundary (large ones)
        for by_tri, col, total in ((0, 0, nv), (1, 2, nt)):
            if total == 0:
                continue
            keys = np.arange(total, dtype=np.int32) if total <= 200_000 else np.unique(np.clip(
                np.concatenate([rows[:, col] + d for d in (-1, 0, 1)] + [np.int32([0, total - 1])]), 0, total - 1)).astype(np.int32)
            got = np.zeros(len(keys), np.int32)
            L.asm_rows_of(h, s, _p(keys), len(keys), by_tri, _p(got))
            want = np.searchsorted(rows[:, col], keys, side="right") - 1
            assert np.array_equal(got, want)
            assert (rows[got, col + 1] > 0).all()  # never an empty part
        vrow = np.searchsorted(rows[:, 0], np.arange(nv), side="right") - 1
        trow = np.searchsorted(rows[:, 2], np.arange(nt), side="right") - 1
        pos4 = np.concatenate([raw_pos.reshape(nv, 3), rows[vrow, 5].astype(np.int32).reshape(nv, 1).view(np.float32)], axis=1)
        local = raw_idx.reshape(nt, 3)
        bad = (local >= rows[trow, 1].astype(np.uint32)[:, None]).any(axis=1)
        idx = (local + rows[trow, 0].astype(np.uint32)[:, None]).astype(np.uint32)
        out.append({"pos4": pos4.astype(np.float32), "idx": idx, "layer": np.ascontiguousarray(rows[trow, 4]).view(np.uint32),
                    "part": rows[trow, 5].astype(np.int32), "starts": starts, "of_set": of_set[:nrows],
                    "bad_tri": int(np.flatnonzero(bad)[0]) if bad.any() else -1})
    for i, p in enumerate(parts):
        plc = np.zeros(3, np.int32)
        L.asm_placement(h, i, _p(plc))
        assert plc[0] == int(bool(p["is_dynamic"])) and plc[2] - plc[1] == len(np.asarray(p["positions"]).reshape(-1, 3))
    L.asm_free(h)
    return out, err, direct_total


def _assemble_reference(parts):
    """partitionEntities + the per-entity loops of TriangleMeshSet.rebuild (CollisionQuery.swift:331-417, 886-900) in numpy."""
    out = []
    for s in range(2):
        pos4, idx, layer, part, starts, of_set = [], [], [], [], [], []
        nv = nt = 0
        for i, p in enumerate(parts):
            if int(bool(p["is_dynamic"])) != s:
                continue
            pos = np.asarray(p["positions"], np.float32).reshape(-1, 3)
            ind = np.asarray(p["indices"], np.uint32).reshape(-1)
            k = len(ind) // 3
            tag = np.full((len(pos), 1), i, np.int32).view(np.float32)
            pos4.append(np.concatenate([pos, tag], axis=1))
            idx.append(ind[:3 * k].reshape(k, 3) + np.uint32(nv))
            layer.append(np.full(k, p["layer"], np.uint32))
            part.append(np.full(k, i, np.int32))
            starts.append(nt)
            of_set.append(i)
            nv, nt = nv + len(pos), nt + k
        starts.append(nt)
        cat = lambda a, shape, dt: np.concatenate(a) if a else np.zeros(shape, dt)  # noqa: E731
        out.append({"pos4": cat(pos4, (0, 4), np.float32), "idx": cat(idx, (0, 3), np.uint32), "layer": cat(layer, 0, np.uint32),
                    "part": cat(part, 0, np.int32), "starts": np.int32(starts), "of_set": np.int32(of_set)})
    return out


def _random_parts(rng, n_parts, max_verts, max_tris):
    parts = []
    for i in range(n_parts):
        nv = int(rng.integers(0, max_verts + 1))
        nt = int(rng.integers(0, max_tris + 1)) if nv > 0 else 0
        extra = int(rng.integers(0, 3))  # a trailing partial triple is ignored (`while tri + 2 < count`)
        parts.append({"positions": rng.standard_normal((nv, 3)).astype(np.float32),
                      "indices": rng.integers(0, max(nv, 1), 3 * nt + (extra if nv > 0 else 0)).astype(np.uint32),
                      "layer": int(rng.integers(1, 2 ** 32)), "is_dynamic": bool(rng.random() < 0.4)})
    return parts


def test_mesh_upload_plan_reproduces_reference_layout(asm, cq):
    """csrc/cq_assemble.h + the expansion kernels' arithmetic: whatever mix of staged and direct copies the plan chooses,
    the set's input arrays equal the reference's per-entity loops — vertices, set-global indices, layer / part per triangle,
    part ranges — for empty worlds, empty parts, trailing partial triples, many small parts and parts above the staging
    limit."""
    rng = np.random.default_rng(99)
    cases = [[], _random_parts(rng, 1, 50, 80), _random_parts(rng, 40, 300, 500), _random_parts(rng, 300, 20, 30)]
    big = _random_parts(rng, 6, 10, 10)
    big[1] = {"positions": rng.standard_normal((70_001, 3)).astype(np.float32),
              "indices": rng.integers(0, 70_001, 3 * 60_007 + 2).astype(np.uint32), "layer": 16, "is_dynamic": False}
    big[3] = {"positions": rng.standard_normal((30_000, 3)).astype(np.float32),
              "indices": rng.integers(0, 30_000, 3 * 2_000).astype(np.uint32), "layer": 2, "is_dynamic": True}
    cases.append(big)
    for parts in cases:
        want = _assemble_reference(parts)
        for limit in (1, 64, 16384, 1 << 30):  # everything direct ... everything staged
            got, err, direct = _upload(asm, cq, parts, limit)
            assert err[0] == 0
            for s in range(2):
                assert got[s]["bad_tri"] == -1
                for k in want[s]:
                    assert got[s][k].tobytes() == want[s][k].tobytes(), (len(parts), limit, s, k)
            if limit == 1 << 30:
                assert direct == 0
    # the default limit: the two large parts go straight from the caller's arrays (positions and indices of part 1,
    # positions of part 3), the small ones are staged
    _, _, direct = _upload(asm, cq, big, 16384)
    assert direct == 3


def test_mesh_upload_plan_rejects_bad_parts(asm, cq):
    rng = np.random.default_rng(5)
    parts = _random_parts(rng, 6, 40, 60)
    for p in parts:
        p["is_dynamic"] = False
    parts[2] = {"positions": rng.standard_normal((60_000, 3)).astype(np.float32),
                "indices": rng.integers(0, 60_000, 3 * 70_000).astype(np.uint32), "layer": 1, "is_dynamic": False}
    parts[4] = {"positions": rng.standard_normal((10, 3)).astype(np.float32),
                "indices": rng.integers(0, 10, 30).astype(np.uint32), "layer": 1, "is_dynamic": False}
    parts[2]["indices"][3 * 65_000 + 1] = 60_000   # first offender, far into a directly copied part
    parts[4]["indices"][7] = 10
    got, err, _ = _upload(asm, cq, parts, 16384)
    assert err[0] == 0                              # the index range is the expansion kernel's check
    static_tri_base = sum(len(p["indices"]) // 3 for p in parts[:2])
    assert got[0]["bad_tri"] == static_tri_base + 65_000
    # invalid arrays are the plan's check: negative counts, missing arrays
    keep = []
    sizes, err = np.zeros(8, np.int64), np.zeros(3, np.int64)
    arr = _mesh_parts(cq, parts, keep)
    arr[3].n_verts = -1
    asm.asm_free(asm.asm_plan(C.byref(arr), len(parts), 16384, _p(sizes), _p(err)))
    assert tuple(err) == (-1, 3, 1)
    arr = _mesh_parts(cq, parts, keep)
    arr[5].n_indices, arr[5].indices = 3, None
    asm.asm_free(asm.asm_plan(C.byref(arr), len(parts), 16384, _p(sizes), _p(err)))
    assert tuple(err) == (-1, 5, 1)
