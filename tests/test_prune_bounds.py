"""The exact-safe bounds the query kernels use to skip work (csrc/cq_pool.cuh: sweep_reach_box, sweep_cannot_matter, the
look-ahead prune of pool_eval), restated in numpy and checked against the oracle's own sweep of single triangles: whenever a
bound says "this triangle cannot matter for a sweep whose best toi so far is bestT", the reference's sweepCapsuleTriangle
(CollisionQuery.swift:1285-1394, through the oracle) must indeed report no hit at or before bestT.  CPU only: this pins the
ARGUMENT; that the kernels implement it is what the GPU parity tests check (every C2 sweep byte-equal to the oracle)."""
import numpy as np
import pytest


def margin_of(frm, L):
    return np.float32(1e-3) + (np.abs(frm).sum(axis=-1) + L) * np.float32(8e-6)


def box_axis_distance_sq(frm, hh, tlo, thi):
    """squared distance from the capsule's axis at t = 0 (vertical segment, centre frm, half height hh) to a box"""
    dx = np.maximum(np.maximum(tlo[..., 0] - frm[..., 0], frm[..., 0] - thi[..., 0]), 0)
    dz = np.maximum(np.maximum(tlo[..., 2] - frm[..., 2], frm[..., 2] - thi[..., 2]), 0)
    dy = np.maximum(np.maximum(tlo[..., 1] - (frm[..., 1] + hh), (frm[..., 1] - hh) - thi[..., 1]), 0)
    return dx * dx + dy * dy + dz * dz


@pytest.fixture(scope="module")
def sweeps_and_single_triangle_tois(orc, scenes):
    rng = np.random.default_rng(2024)
    n_tri, n_sweep = 48, 3000
    tris = []
    for _ in range(n_tri):
        c = rng.uniform(-3, 3, 3)
        tris.append((c + rng.normal(0, rng.uniform(0.05, 2.0), (3, 3))).astype(np.float32))
    q = scenes.gen_casts(n_sweep, [-4, -4, -4], [4, 4, 4], seed=99, radius=0.6, half_height=0.5, len_range=(0.05, 4.0), expand=1.0)
    toi = np.full((n_tri, n_sweep), np.inf, np.float32)
    for k, t in enumerate(tris):
        w = orc.OracleWorld([scenes.part(t, np.array([0, 1, 2], np.uint32), entity_id=0)])
        if w.counts(0)["triangles"] == 0:
            continue  # degenerate sample
        h = w.capsule_cast(q, 0, orc.ORDER_REFERENCE)
        hit = h["triangle_index"] >= 0
        toi[k, hit] = h["toi"][hit]
        w.close()
    return np.stack(tris), q, toi


def test_box_lower_bound_never_drops_a_triangle_that_could_win(sweeps_and_single_triangle_tois):
    """sweep_cannot_matter (b): axis-to-box distance > radius + bestT + margin  =>  no contact at or before bestT."""
    tris, q, toi = sweeps_and_single_triangle_tois
    frm, r, hh = q["from"], q["radius"], q["half_height"]
    L = np.linalg.norm(q["delta"], axis=1).astype(np.float32)
    tlo, thi = tris.min(axis=1), tris.max(axis=1)
    dropped_total = 0
    for best_scale in (0.0, 0.1, 0.35, 0.7, 1.0):  # "best toi so far" anywhere between 0 and the whole sweep
        bestT = (L * np.float32(best_scale)).astype(np.float32)
        reach = r + bestT + margin_of(frm, L)
        d2 = box_axis_distance_sq(frm[None, :, :], hh[None, :], tlo[:, None, :], thi[:, None, :])
        dropped = d2 > (reach * reach * np.float32(1.0001))[None, :]
        dropped_total += int(dropped.sum())
        wrong = dropped & (toi <= bestT[None, :])
        assert not wrong.any(), (best_scale, int(wrong.sum()))
    assert dropped_total > 10000  # the bound really bites on this sample


def test_reach_box_contains_every_triangle_that_could_win(sweeps_and_single_triangle_tois):
    """sweep_reach_box: a triangle with a contact at or before bestT overlaps the box of the capsule swept to bestT."""
    tris, q, toi = sweeps_and_single_triangle_tois
    frm, r, hh = q["from"], q["radius"], q["half_height"]
    L = np.linalg.norm(q["delta"], axis=1).astype(np.float32)
    d = q["delta"] / np.maximum(L, np.float32(1e-30))[:, None]
    tlo, thi = tris.min(axis=1), tris.max(axis=1)
    for best_scale in (0.0, 0.2, 0.6, 1.0):
        bestT = (L * np.float32(best_scale)).astype(np.float32)
        m = margin_of(frm, L)
        b = frm + d * (bestT + m)[:, None]
        ext = np.stack([r + m, r + hh + m, r + m], 1)
        qlo, qhi = np.minimum(frm, b) - ext, np.maximum(frm, b) + ext
        disjoint = ((tlo[:, None, :] > qhi[None, :, :]) | (thi[:, None, :] < qlo[None, :, :])).any(axis=2)
        wrong = disjoint & (toi <= bestT[None, :])
        assert not wrong.any(), (best_scale, int(wrong.sum()))


def test_a_best_toi_of_zero_cannot_be_beaten(sweeps_and_single_triangle_tois):
    """sweep_cannot_matter (a) rests on toi >= 0 for every contact the reference can report."""
    _, _, toi = sweeps_and_single_triangle_tois
    assert (toi >= 0).all() and (toi == 0).any()  # capsules that start in contact report exactly 0


def test_look_ahead_step_beyond_best_means_no_contact_before_best(orc, sweeps_and_single_triangle_tois):
    """pool_eval's look-ahead prune: an evaluation at t0 that reports distance `dist` (no contact) and whose conservative
    step lands beyond bestT + margin (t0 + (dist - r) > bestT + margin) proves there is no contact in [t0, bestT]."""
    tris, q, toi = sweeps_and_single_triangle_tois
    rng = np.random.default_rng(5)
    frm, r, hh = q["from"], q["radius"], q["half_height"]
    L = np.linalg.norm(q["delta"], axis=1).astype(np.float32)
    d = q["delta"] / np.maximum(L, np.float32(1e-30))[:, None]
    pruned = 0
    for k in range(len(tris)):
        t0 = (L * rng.uniform(0, 1, len(L))).astype(np.float32)
        centers = (frm + d * t0[:, None]).astype(np.float32)
        dist = orc.segment_triangle_distance_batch(centers, hh, np.broadcast_to(tris[k], (len(L), 3, 3)).copy())[0]
        for best_scale in (0.3, 0.6, 1.0):
            bestT = np.maximum(t0, L * np.float32(best_scale)).astype(np.float32)
            prune = (dist > r + np.float32(1e-5)) & (t0 + (dist - r) > bestT + margin_of(frm, L))
            pruned += int(prune.sum())
            wrong = prune & (toi[k] >= t0) & (toi[k] <= bestT)
            assert not wrong.any(), (k, best_scale, int(wrong.sum()))
    assert pruned > 10000
