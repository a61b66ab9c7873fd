"""The product's host rebuild of the reference BVH's visiting order (swift-game-engine_b200/csrc/cq_reftree.h) against
the independently written Python transliteration of the reference's build (tests/independent_bvh.py, straight from
Game/CollisionQuery.swift:577-670): same triOrder, same leaves, hence the same visiting rank for every triangle —
whatever the number of worker threads.  Runs without a GPU."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import independent_bvh as ibvh  # noqa: E402

SRC = os.path.join(HERE, "hostmath", "reftree.cpp")
LIB = os.path.join(HERE, "hostmath", "libcq_reftree.so")
HDR = os.path.join(os.path.dirname(HERE), "swift-game-engine_b200", "csrc", "cq_reftree.h")


@pytest.fixture(scope="module")
def rt():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    if not os.path.exists(LIB) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(LIB):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread", "-ffp-contract=off",
                               "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.reftree_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build(L, lo, hi, threads, want_nodes=False):
    n = len(lo)
    order, rank, counts = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(2, np.int32)
    nodes = np.zeros((2 * max(n, 1), 5), np.int32) if want_nodes else None
    L.reftree_build(_p(lo), _p(hi), n, threads, _p(order), _p(rank), _p(counts), _p(nodes) if want_nodes else None,
                    2 * max(n, 1))
    return order, rank, counts, (nodes[:counts[0]] if want_nodes else None)


def reference_visit_order(bvh):
    """The order in which the reference's walks reach triangles when nothing is culled: pop, push left, push right."""
    out, stack = [], [bvh.root] if bvh.root >= 0 else []
    while stack:
        nd = bvh.nodes[stack.pop()]
        if nd["left"] < 0:
            out += [bvh.tri_order[i] for i in range(nd["start"], nd["start"] + nd["count"])]
        else:
            stack += [nd["left"], nd["right"]]
    return out


def boxes(rng, n, kind):
    c = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    if kind == "flat":  # a terrain-like slab: one axis nearly constant
        c[:, 1] = (rng.standard_normal(n) * 0.01).astype(np.float32)
    if kind == "dupes":  # many identical centroids: the sorted-median fallback (CollisionQuery.swift:637-653)
        c = c[rng.integers(0, max(n // 40, 1), n)]
    if kind == "grid":  # exact ties on the pivot
        c = np.round(c / 4) * 4
    if kind == "same":
        c[:] = c[0]
    e = rng.uniform(0.0, 2.0, (n, 3)).astype(np.float32)
    return np.ascontiguousarray(c - e), np.ascontiguousarray(c + e)


@pytest.mark.parametrize("kind,n", [("random", 1), ("random", 4), ("random", 5), ("random", 9), ("random", 3000), ("flat", 2500),
                                     ("dupes", 3000), ("grid", 3000), ("same", 300)])
def test_ref_tree_equals_independent_transliteration(rt, kind, n):
    rng = np.random.default_rng(n * 7 + len(kind))
    lo, hi = boxes(rng, n, kind)
    bvh = ibvh.ReferenceBVH([(tuple(lo[i]), tuple(hi[i])) for i in range(n)])
    order, rank, counts, nodes = build(rt, lo, hi, 1, want_nodes=True)
    assert order.tolist() == bvh.tri_order
    assert counts[0] == len(bvh.nodes)
    visit = reference_visit_order(bvh)
    assert len(visit) == n and np.argsort(rank).tolist() == visit
    # the same set of leaves (the numbering of nodes is the builder's own)
    mine = sorted((int(s), int(c)) for l, r, s, c, p in nodes if l < 0)
    theirs = sorted((nd["start"], nd["count"]) for nd in bvh.nodes if nd["left"] < 0)
    assert mine == theirs and counts[1] == len(theirs)
    # structure: every child points back at its parent, internal nodes cover their children
    for k, (l, r, s, c, p) in enumerate(nodes):
        if l >= 0:
            assert nodes[l][4] == k and nodes[r][4] == k


def test_ref_tree_is_thread_count_invariant(rt):
    rng = np.random.default_rng(99)
    n = 400_000  # large enough for the breadth-first top + parallel subtrees path
    lo, hi = boxes(rng, n, "flat")
    o1, r1, c1, _ = build(rt, lo, hi, 1)
    for threads in (2, 5, 8):
        o, r, c, _ = build(rt, lo, hi, threads)
        assert np.array_equal(o, o1) and np.array_equal(r, r1) and np.array_equal(c, c1)
    assert sorted(r1.tolist()) == list(range(n))
