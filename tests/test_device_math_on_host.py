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
