"""GPU parity tests proper: the CUDA path (through the C ABI, libcq.so) against the CPU oracle on the
same seeded inputs.  Bar (BASELINE.json north_star): hit / no-hit and triangle index bit-exact except
flagged exact-tie cases; toi / normals / positions within 1e-4 rel, 1e-5 abs — in practice the two
sides run the same IEEE op sequence, so the oracle must match BIT-EXACTLY on every field, ties included:
a world created in CQ_ORDER_REFERENCE (the default) against the oracle's ORDER_REFERENCE (the reference's own
BVH and depth-first visiting order), a CQ_ORDER_CANONICAL world against ORDER_CANONICAL.  The two order
constants have the same values on both sides, so `g.order` is passed to the oracle."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ABS_TOL, REL_TOL = 1e-5, 1e-4  # the stated tolerance (only used against the REFERENCE-order oracle)


def _caller_order(hits, counts):
    """What every caller of capsuleOverlapAll in the reference does with the result (Systems.swift:759): a stable sort by
    depth, deepest first.  The oracle's C wrapper emits that order; in reference order the library returns the hits the
    way the reference itself returns them — in visiting order — so the comparison sorts the library's rows the same way."""
    depth = np.where(np.arange(hits.shape[1])[None, :] < counts[:, None], hits["depth"], -np.inf)
    idx = np.argsort(-depth, axis=1, kind="stable")
    return np.take_along_axis(hits, idx, axis=1)


def _fields_equal(a, b, fields):
    return {f: bool(np.array_equal(a[f], b[f])) for f in fields}


@pytest.fixture(scope="module", params=["hulls-reference", "hulls-canonical", "render-reference", "render-canonical"])
def world(request, cq, orc, scenes):
    name, order = request.param.split("-")
    assert (cq.ORDER_REFERENCE, cq.ORDER_CANONICAL) == (orc.ORDER_REFERENCE, orc.ORDER_CANONICAL)
    parts = scenes.mirror_scene(use_hulls=name == "hulls")
    g = cq.CollisionQuery(parts, order=cq.ORDER_REFERENCE if order == "reference" else cq.ORDER_CANONICAL)
    assert g.info()["order"] == g.order
    o = orc.OracleWorld(parts)
    yield name, parts, g, o
    g.close()
    o.close()


def test_soup_upload_matches_reference_rebuild(world):
    """TriangleMeshSet.rebuild (CollisionQuery.swift:331-417): world-space vertices, filtered triangle
    numbering, per-triangle AABBs and layers must be bit-identical."""
    name, parts, g, o = world
    gs, os_ = g.read_soup(0), o.read_soup(0)
    assert gs["positions"].shape == os_["positions"].shape
    for k in ("positions", "indices", "aabbs", "layers", "parts"):
        assert np.array_equal(gs[k], os_[k]), k
    info = g.info()
    assert info["n_static_triangles"] == o.counts(0)["triangles"]
    assert info["n_dynamic_triangles"] == 0


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_capsule_cast_bit_exact_vs_oracle(world, cq, scenes, orc, mode):
    name, parts, g, o = world
    lo, hi = scenes.scene_aabb(parts[1:])
    n = 20000 if name == "hulls" else 3000
    q = scenes.gen_casts(n, lo, hi, seed=100 + mode)
    q["delta"][:7] = 0  # zero-length sweeps -> nil (CollisionQuery.swift:988)
    q["mask"][7:40] = 1  # ground plane only
    q["mask"][40:60] = 2  # nothing on this layer
    got, flags = [g.capsuleCast, g.capsuleCastBlocking, g.capsuleCastGround][mode](q, with_flags=True)
    ref = o.capsule_cast(q, mode, g.order)
    assert (got["triangle_index"][:7] == -1).all() and (got["triangle_index"][40:60] == -1).all()
    assert np.array_equal(got["triangle_index"], ref["triangle_index"])
    for f in ("toi", "position", "normal", "triangle_normal"):
        assert np.array_equal(got[f], ref[f]), f
    assert (got["triangle_index"] >= 0).sum() > n // 10
    # against the OTHER order rule: the winning toi is order-independent, only the triangle named on an exact tie may
    # differ — and every such query carries the TIE flag (unflagged mismatches == 0)
    other = o.capsule_cast(q, mode, orc.ORDER_CANONICAL if g.order == orc.ORDER_REFERENCE else orc.ORDER_REFERENCE)
    assert np.array_equal(got["triangle_index"] >= 0, other["triangle_index"] >= 0)
    assert np.array_equal(got["toi"], other["toi"])
    diff = got["triangle_index"] != other["triangle_index"]
    assert (flags[diff] & cq.HIT_TIE).all() and not (flags[got["triangle_index"] < 0]).any()
    if name == "render":
        assert diff.any() and (flags & cq.HIT_TIE).mean() > 0.05  # shared edges / vertices of a closed mesh tie all the time
    st = orc.Stats()
    o.capsule_cast(q, mode, g.order, 1, st)
    assert int((flags & cq.HIT_TIE).astype(bool).sum()) == st.ties


def test_capsule_overlap_bit_exact(world, scenes, orc):
    name, parts, g, o = world
    lo, hi = scenes.scene_aabb(parts[1:])
    n = 20000 if name == "hulls" else 4000
    c = scenes.gen_capsules(n, lo, hi, seed=7)
    got, flags = g.capsuleOverlap(c, with_flags=True)
    ref = o.capsule_overlap(c, g.order)
    other = o.capsule_overlap(c, orc.ORDER_CANONICAL if g.order == orc.ORDER_REFERENCE else orc.ORDER_REFERENCE)
    diff = got["triangle_index"] != other["triangle_index"]
    assert np.array_equal(got["depth"], other["depth"]) and (flags[diff] & 1).all()
    for f in ("triangle_index", "depth", "position", "normal", "triangle_normal"):
        assert np.array_equal(got[f], ref[f]), f
    assert (got["triangle_index"] >= 0).sum() > n // 20


def test_capsule_overlap_all_bit_exact(world, scenes, orc):
    name, parts, g, o = world
    lo, hi = scenes.scene_aabb(parts[1:])
    n = 20000 if name == "hulls" else 4000
    c = scenes.gen_capsules(n, lo, hi, seed=8)
    for max_hits in (8, 3):
        got, gcnt, gov = g.capsuleOverlapAll(c, max_hits)
        ref, rcnt, rov = o.capsule_overlap_all(c, max_hits, g.order)
        assert np.array_equal(gcnt, rcnt) and np.array_equal(gov, rov)
        if g.order == orc.ORDER_REFERENCE:
            got = _caller_order(got, gcnt)
        for f in ("triangle_index", "depth", "position", "normal", "triangle_normal"):
            assert np.array_equal(got[f], ref[f]), f
        # the other order rule: same SET of triangles whenever it did not overflow
        ref2, rcnt2, rov2 = o.capsule_overlap_all(c, max_hits, orc.ORDER_CANONICAL if g.order == orc.ORDER_REFERENCE
                                                  else orc.ORDER_REFERENCE)
        ok = rov2 == 0
        assert np.array_equal(gov, rov2) and np.array_equal(gcnt[ok], rcnt2[ok])
        assert np.array_equal(np.sort(got["triangle_index"][ok], axis=1), np.sort(ref2["triangle_index"][ok], axis=1))
        if name == "render" and max_hits == 3:
            assert gov.sum() > 50  # the first-visited rule is really exercised


def test_raycast_vs_oracle(world, cq, scenes, orc):
    name, parts, g, o = world
    lo, hi = scenes.scene_aabb(parts)
    n = 20000 if name == "hulls" else 3000
    r = scenes.gen_rays(n, lo, hi, seed=9, expand=2.0)
    r["direction"][:100] *= 3.5  # direction is not normalised by the callee
    r["direction"][100:110, 0] = 0  # axis-parallel components -> 1/0 replacement path
    r["mask"][110:130] = 2
    got, flags = g.raycast(r, with_flags=True)
    ref = o.raycast(r, g.order)  # reference order: the reference's own walk; canonical: brute force over all triangles
    for f in ("triangle_index", "distance", "position", "normal"):
        assert np.array_equal(got[f], ref[f]), f
    assert (got["triangle_index"] >= 0).sum() > n // 10
    # Classification of every difference between the two order rules (canonical = minimum over ALL triangles):
    #   tie      same distance, another triangle                         -> must carry the TIE flag in the world that saw both
    #   culled   the reference's walk lost a nearer / the only hit to its non-conservative slab test (:933, :1603-1631)
    #   other    anything else — must not exist
    can = got if g.order == orc.ORDER_CANONICAL else o.raycast(r, orc.ORDER_CANONICAL)
    refo = got if g.order == orc.ORDER_REFERENCE else o.raycast(r, orc.ORDER_REFERENCE)
    diff = can["triangle_index"] != refo["triangle_index"]
    both = (can["triangle_index"] >= 0) & (refo["triangle_index"] >= 0)
    tie = diff & both & (can["distance"] == refo["distance"])
    culled = diff & (can["triangle_index"] >= 0) & ((refo["triangle_index"] < 0) | (refo["distance"] > can["distance"]))
    assert not (diff & ~tie & ~culled).any()
    assert diff.mean() < 0.01
    if g.order == orc.ORDER_CANONICAL:
        assert (flags[tie] & cq.HIT_TIE).all()


def test_move_and_slide_bit_exact_multi_step(world, cq, scenes, orc):
    """KinematicMoveStopSystem.fixedUpdate body (Systems.swift:1842-1901), 6 consecutive fixed steps with
    carried state: every field of every character must equal the canonical-order oracle's."""
    name, parts, g, o = world
    n = 4096 if name == "hulls" else 768
    pos, vel = scenes.gen_c3_characters(n, seed=11)
    pos[: n // 8, 1] += 3.0  # some start in the air
    pos[n // 8: n // 4, 1] -= 0.4  # some start penetrating the ground -> depenetration path
    sg = cq.init_states(pos, vel)
    so = orc.init_states(pos, vel)
    assert sg.tobytes() == so.tobytes()
    pg, po = cq.default_params(), orc.default_params()
    assert pg.tobytes() == po.tobytes()
    for step in range(6):
        g.move_and_slide(sg, pg)
        o.move_and_slide(so, po, order=g.order)
        for f in sg.dtype.names:
            if f == "_pad":
                continue
            assert np.array_equal(sg[f], so[f]), (step, f, int((sg[f] != so[f]).sum()))
    assert sg["grounded"].mean() > 0.5
    assert (sg["manifold_count"] > 0).any()


def test_move_and_slide_human_scale_params(world, cq, scenes, orc):
    name, parts, g, o = world
    n = 2048 if name == "hulls" else 512
    pos, vel = scenes.gen_c3_characters(n, seed=12, radius=0.4, half_height=0.5)
    sg, so = cq.init_states(pos, vel), orc.init_states(pos, vel)
    pg = cq.default_params(radius=0.4, half_height=0.5, skin_width=0.08, fall_probe_distance=50.0)
    po = orc.default_params(radius=0.4, half_height=0.5, skin_width=0.08, fall_probe_distance=50.0)
    for step in range(4):
        g.move_and_slide(sg, pg, flags=0)
        o.move_and_slide(so, po, flags=0, order=g.order)
    assert sg.tobytes() == so.tobytes()


def test_refit_matches_reference_update_transforms(cq, orc, scenes):
    """updateDynamicTransforms -> TriangleMeshSet.updateTransforms + BVH.refit (CollisionQuery.swift:419-462,
    528-575): spin the mirror (dynamic set) and compare soup + queries after every refit."""
    parts = scenes.mirror_scene(use_hulls=False, mirror_dynamic=True)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    assert g.info()["n_dynamic_triangles"] == o.counts(1)["triangles"] > 10000
    a = scenes.load_mirror_fixture()
    t, q0, s = scenes.transform_from_matrix(scenes.mirror_model(a["transform"]))
    lo, hi = scenes.scene_aabb(parts[1:])
    qs = scenes.gen_casts(2000, lo - 2, hi + 2, seed=21)
    rays = scenes.gen_rays(2000, lo, hi, seed=22)
    for step in range(1, 4):
        rot = scenes.quat_mul(scenes.quat_angle_axis(np.radians(7.0 * step), (0, 1, 0)), q0)
        model = scenes.trs_model(t, rot, s)
        g.update_transforms([1], [model])
        o.update_transforms([1], [model])
        assert o.check_bvh(1)
        gs, os_ = g.read_soup(1), o.read_soup(1)
        for k in ("positions", "aabbs"):
            assert np.array_equal(gs[k], os_[k]), (step, k)
        got, ref = g.capsuleCast(qs), o.capsule_cast(qs, 0, g.order)
        assert got.tobytes() == ref.tobytes()
        gr, rr = g.raycast(rays), o.raycast(rays, g.order)
        assert np.array_equal(gr["triangle_index"], rr["triangle_index"])
        assert np.array_equal(gr["distance"], rr["distance"])
    with pytest.raises(cq.CQError):
        g.update_transforms([99], [model])
    g.close()
    o.close()


def test_empty_and_tiny_worlds(cq, orc, scenes):
    # empty world: every query answers nil
    g = cq.CollisionQuery([])
    q = scenes.gen_casts(64, [-1, -1, -1], [1, 1, 1], seed=1)
    assert (g.capsuleCast(q)["triangle_index"] == -1).all()
    assert (g.raycast(scenes.gen_rays(64, [-1, -1, -1], [1, 1, 1], seed=1))["triangle_index"] == -1).all()
    s = cq.init_states(np.zeros((8, 3), np.float32), np.ones((8, 3), np.float32))
    g.move_and_slide(s, cq.default_params())
    assert not s["grounded"].any()
    assert g.capsuleCast(q[:0]).shape == (0,)
    g.close()
    # 1, 2, 5 triangles (leaf-collapse edge cases of the LBVH) incl. a fully degenerate part
    v, i = scenes.plane_mesh(10.0)
    for ntri_idx in (i[:3], i, np.concatenate([i, i[:3], i, i[:3]])):
        parts = [scenes.part(v, ntri_idx, entity_id=0),
                 scenes.part(np.zeros((3, 3), np.float32), [0, 1, 2], entity_id=1)]  # zero-area -> filtered
        g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
        assert g.info()["n_static_triangles"] == o.counts(0)["triangles"]
        q = scenes.gen_casts(512, [-5, 0, -5], [5, 3, 5], seed=3, expand=1.0)
        assert g.capsuleCast(q).tobytes() == o.capsule_cast(q, 0, g.order).tobytes()
        g.close()
        o.close()


def test_static_and_dynamic_sets_index_offset(cq, orc, scenes):
    """Dynamic-set triangle indices are offset by the static count (CollisionQuery.swift:782,1004); static
    wins exact ties (chooseNearest <=)."""
    v, i = scenes.plane_mesh(20.0)
    bv, bi = scenes.box_mesh(2.0)
    parts = [scenes.part(v, i, entity_id=0),
             scenes.part(bv, bi, scenes.trs_model((0, 1, 0)), is_dynamic=True, entity_id=1),
             scenes.part(bv, bi, scenes.trs_model((4, 1, 0)), entity_id=2)]
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    inf = g.info()
    assert inf["n_static_triangles"] == 14 and inf["n_dynamic_triangles"] == 12
    q = scenes.gen_casts(4000, [-4, 0, -4], [8, 4, 4], seed=5, radius=0.5, half_height=0.5, expand=0.5)
    got, ref = g.capsuleCast(q), o.capsule_cast(q, 0, g.order)
    assert got.tobytes() == ref.tobytes()
    assert (got["triangle_index"] >= 14).any()
    assert g.triangle_material(0)["mu_s"] == pytest.approx(0.8)
    g.close()
    o.close()


def test_counters_match_reference_stats(world, cq, scenes, orc):
    """capsuleCandidateCount (CollisionQuery.swift:1066) is tree-independent: the GPU's candidate counter
    must equal the reference's on the same batch."""
    name, parts, g, o = world
    lo, hi = scenes.scene_aabb(parts[1:])
    q = scenes.gen_casts(2000, lo, hi, seed=31)
    st = orc.Stats()
    o.capsule_cast(q, 0, orc.ORDER_REFERENCE, 1, st)
    g.set_counting(cq.COUNT_REFERENCE)
    g.resetStats()
    ref_mode = g.capsuleCast(q)
    c = g.stats()
    assert c["candidates"] == st.candidates
    assert 0 < c["distance_evals"] <= st.distance_evals  # exact-safe pruning only removes work
    assert c["nodes_visited"] > 0 and c["kernel_launches"] >= 1
    # the path as shipped: sweeps that hold a hit cull against the capsule swept to that hit and drop candidates that cannot
    # matter — fewer nodes, candidates and evaluations, the same answers
    g.set_counting(cq.COUNT_PATH)
    g.resetStats()
    path_mode = g.capsuleCast(q)
    p = g.stats()
    g.set_counting(False)
    assert path_mode.tobytes() == ref_mode.tobytes() == g.capsuleCast(q).tobytes()
    assert 0 < p["candidates"] <= c["candidates"] and p["nodes_visited"] <= c["nodes_visited"]
    assert 0 < p["distance_evals"] <= c["distance_evals"]


def test_terrain_lbvh_and_blocking_sweeps(cq, orc, scenes):
    """Config C4 scaled down (300x300 cells = 180,000 triangles): device LBVH build (Morton + radix sort +
    Karras + fit) at a size where every radix pass spans many tiles, sweeps checked against the oracle."""
    parts, half = scenes.terrain_scene(cells=300, cell=2.0)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    assert g.info()["n_static_triangles"] == o.counts(0)["triangles"] == 180000
    gs, os_ = g.read_soup(0), o.read_soup(0)
    assert np.array_equal(gs["positions"], os_["positions"]) and np.array_equal(gs["aabbs"], os_["aabbs"])
    for r, hh in ((0.4, 0.5), (1.5, 1.0)):
        q = scenes.gen_c4_casts(20000, half, seed=41, radius=r, half_height=hh)
        got, ref = g.capsuleCastBlocking(q), o.capsule_cast(q, 1, g.order)
        assert got.tobytes() == ref.tobytes()
        assert (got["triangle_index"] >= 0).mean() > 0.2
    rays = scenes.gen_rays(20000, [-half, -10, -half], [half, 30, half], seed=42, expand=0.0)
    gr, rr = g.raycast(rays), o.raycast(rays, g.order)
    assert gr.tobytes() == rr.tobytes() and (gr["triangle_index"] >= 0).mean() > 0.3
    # characters walking on the terrain, human scale
    x = np.random.default_rng(3).uniform(-half + 10, half - 10, (4096, 2))
    y = scenes.terrain_height(x[:, 0], x[:, 1]) + 0.9 + 0.3
    pos = np.stack([x[:, 0], y, x[:, 1]], axis=1).astype(np.float32)
    vel = np.random.default_rng(4).uniform(-4, 4, (4096, 3)).astype(np.float32) * np.float32([1, 0, 1])
    sg, so = cq.init_states(pos, vel), orc.init_states(pos, vel)
    kw = dict(radius=0.4, half_height=0.5, skin_width=0.08)
    for _ in range(4):
        g.move_and_slide(sg, cq.default_params(**kw))
        o.move_and_slide(so, orc.default_params(**kw), order=g.order)
    assert sg.tobytes() == so.tobytes()
    assert sg["grounded"].mean() > 0.8
    g.close()
    o.close()


def test_onesweep_sort_builds_the_same_tree_as_the_classic_sort(cq, orc, scenes, monkeypatch):
    """The onesweep radix sort (decoupled look-back) and the classic 3-kernel-per-pass sort are both stable LSD
    sorts of the same (Morton key, triangle) pairs, so the two LBVHs — hence every traversal counter — must be
    identical; results are checked against the oracle as well."""
    parts, half = scenes.terrain_scene(cells=400, cell=2.0)  # 320,000 triangles = 79 tiles of 4096 keys
    q = scenes.gen_c4_casts(30000, half, seed=77, radius=1.5, half_height=1.0)
    out = {}
    for mode in ("classic", "onesweep"):
        monkeypatch.setenv("CQ_SORT", mode)
        g = cq.CollisionQuery(parts)
        g.set_counting(True)
        g.resetStats()
        hits = g.capsuleCastBlocking(q)
        out[mode] = (hits.tobytes(), g.stats()["nodes_visited"], g.stats()["candidates"])
        g.close()
    assert out["classic"] == out["onesweep"]
    o = orc.OracleWorld(parts)
    assert out["onesweep"][0] == o.capsule_cast(q, 1, orc.ORDER_REFERENCE).tobytes()  # worlds default to reference order


def test_c2_sweeps_against_semla(cq, orc, scenes):
    """Config C2 (scaled to 4,096 sweeps for the oracle's sake): plain capsuleCast against the Semla render mesh
    (50,002 triangles, FBX-regenerated stand-in) at its demo placement."""
    parts = scenes.semla_scene(use_hulls=False)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    assert g.info()["n_static_triangles"] == o.counts(0)["triangles"] == 50004
    lo, hi = scenes.scene_aabb(parts[1:])
    q = scenes.gen_casts(4096, lo, hi, seed=0xC0111DE2)
    got, flags = g.capsuleCast(q, with_flags=True)
    assert got.tobytes() == o.capsule_cast(q, 0, g.order, 8).tobytes()  # the reference's own answers, ties included
    can = o.capsule_cast(q, 0, orc.ORDER_CANONICAL, 8)
    gc = cq.CollisionQuery(parts, order=cq.ORDER_CANONICAL)
    assert gc.capsuleCast(q).tobytes() == can.tobytes()
    gc.close()
    diff = got["triangle_index"] != can["triangle_index"]
    assert 0.1 < diff.mean() < 0.6 and (flags[diff] & cq.HIT_TIE).all()  # ~30% of these sweeps end on a shared edge / vertex
    assert (got["triangle_index"] >= 0).mean() > 0.4
    hq = scenes.gen_casts(4096, lo, hi, seed=5)
    gh, oh = cq.CollisionQuery(scenes.semla_scene(use_hulls=True)), orc.OracleWorld(scenes.semla_scene(use_hulls=True))
    assert gh.capsuleCastBlocking(hq).tobytes() == oh.capsule_cast(hq, 1, gh.order).tobytes()
    for w in (g, o, gh, oh):
        w.close()


def test_c5_merged_scene_rays_and_refit(cq, orc, scenes):
    """Config C5 scaled down: 17-Cheese + Semla + mirror merged (~132k triangles), the mirror spinning in the
    dynamic set; after every refit rays and sweeps must still match the oracle."""
    parts = scenes.merged_scene(mirror_dynamic=True)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    assert g.info()["n_static_triangles"] == o.counts(0)["triangles"] and g.info()["n_dynamic_triangles"] == 14211
    a = scenes.load_mirror_fixture()
    t, q0, s = scenes.transform_from_matrix(scenes.mirror_model(a["transform"]))
    lo, hi = scenes.scene_aabb(parts[1:])
    rays = scenes.gen_rays(20000, lo, hi, seed=0xC0111DE5, expand=5.0, y_range=(0.0, 12.0))
    for step in range(1, 3):
        rot = scenes.quat_mul(scenes.quat_angle_axis(np.radians(float(step)), (0, 1, 0)), q0)
        model = scenes.trs_model(t, rot, s)
        g.update_transforms([3], [model])
        o.update_transforms([3], [model])
        gr, rr = g.raycast(rays), o.raycast(rays, g.order, 8)
        assert gr.tobytes() == rr.tobytes()  # the reference's own walk over its refitted tree, grazing culls included
        assert (gr["triangle_index"] >= 0).mean() > 0.2
    qs = scenes.gen_casts(1500, lo, hi, seed=9, expand=1.0)
    assert g.capsuleCast(qs).tobytes() == o.capsule_cast(qs, 0, g.order, 8).tobytes()
    g.close()
    o.close()


def test_c1_trajectory_matches_golden(cq, scenes):
    """Config C1 on the GPU: 4 characters driven for 600 fixed steps over the demo's static world must reproduce
    the committed golden trajectories (tests/golden/c1_trajectory.npz) bit for bit: the reference-order one in the
    default order — the trajectory the reference itself walks, ties resolved by ITS visiting order, which the contact
    cache keyed by triangle index makes visible from frame 203 on — and the canonical one in canonical order."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_trajectory.npz"))
    for order, tag in ((cq.ORDER_REFERENCE, "reference"), (cq.ORDER_CANONICAL, "canonical")):
        g = cq.CollisionQuery(scenes.c1_scene(), order=order)
        s = cq.init_states(scenes.C1_STARTS)
        p = cq.default_params()
        rec = scenes.c1_run(lambda st: g.move_and_slide(st, p), s, 600, scenes.C1_SPEEDS)
        for k in rec.dtype.names:
            assert np.array_equal(rec[k], z[f"{tag}_{k}"]), (tag, k)
        g.close()


def test_parameter_edge_cases(cq, orc, scenes):
    """Sphere (halfHeight 0), tiny radius (minAdvance floor 1e-4), one slide iteration, no snap / no fall probe,
    layer masks, overlap-all overflow flag."""
    parts = scenes.mirror_scene(use_hulls=True)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    lo, hi = scenes.scene_aabb(parts[1:])
    for r, hh in ((0.5, 0.0), (0.003, 0.2), (3.0, 2.0)):
        q = scenes.gen_casts(3000, lo, hi, seed=int(r * 1000), radius=r, half_height=hh, len_range=(0.001, 1.0))
        for mode in (0, 1, 2):
            got = [g.capsuleCast, g.capsuleCastBlocking, g.capsuleCastGround][mode](q)
            assert got.tobytes() == o.capsule_cast(q, mode, g.order).tobytes(), (r, hh, mode)
    pos, vel = scenes.gen_c3_characters(1024, seed=5)
    for kw in (dict(max_slide_iterations=1), dict(snap_distance=0.0), dict(fall_probe_distance=0.0),
               dict(collision_mask=1), dict(collision_mask=1 << 4), dict(skin_width=0.0, ground_snap_skin=0.0)):
        sg, so = cq.init_states(pos, vel), orc.init_states(pos, vel)
        for _ in range(3):
            g.move_and_slide(sg, cq.default_params(**kw))
            o.move_and_slide(so, orc.default_params(**kw), order=g.order)
        assert sg.tobytes() == so.tobytes(), kw
    rparts = scenes.mirror_scene(use_hulls=False)
    gr, orr = cq.CollisionQuery(rparts), orc.OracleWorld(rparts)
    c = scenes.gen_capsules(2000, lo, hi, seed=8, expand=0.2)
    got, cnt, ov = gr.capsuleOverlapAll(c, 8)
    ref, rcnt, rov = orr.capsule_overlap_all(c, 8, gr.order)
    assert ov.sum() > 100 and np.array_equal(ov, rov) and np.array_equal(cnt, rcnt)
    assert _caller_order(got, cnt).tobytes() == ref.tobytes()
    gc = cq.CollisionQuery(rparts, order=cq.ORDER_CANONICAL)
    got, cnt, ov = gc.capsuleOverlapAll(c, 8)
    ref, rcnt, rov = orr.capsule_overlap_all(c, 8, gc.order)
    assert np.array_equal(ov, rov) and np.array_equal(cnt, rcnt) and got.tobytes() == ref.tobytes()
    gc.close()
    for w in (g, o, gr, orr):
        w.close()


def test_kinematic_platforms_carry_and_push(cq, orc, scenes):
    """applyPlatformDelta + the dynamic triangle set: an elevator and a horizontal mover (DemoScene-like) in the
    dynamic set, moved and refitted every step; characters on top of / beside / far from them."""
    bv, bi = scenes.box_mesh(4.0)
    v, i = scenes.plane_mesh(80.0)
    t0 = [np.float32([0, -1.0, 0]), np.float32([12, -1.0, 0])]
    parts = [scenes.part(v, i, scenes.trs_model((0, -3, 0)), entity_id=0),
             scenes.part(bv, bi, scenes.trs_model(t0[0]), is_dynamic=True, entity_id=1),
             scenes.part(bv, bi, scenes.trs_model(t0[1]), is_dynamic=True, entity_id=2)]
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    rng = np.random.default_rng(3)
    n = 3000
    pos = np.stack([rng.uniform(-6, 18, n), np.full(n, -0.45), rng.uniform(-5, 5, n)], axis=1).astype(np.float32)
    on_top = (np.abs(pos[:, 0]) < 2) & (np.abs(pos[:, 2]) < 2) | (np.abs(pos[:, 0] - 12) < 2) & (np.abs(pos[:, 2]) < 2)
    pos[on_top, 1] = 1.0 + 2.5 + 0.05
    vel = (rng.uniform(-3, 3, (n, 3)) * [1, 0, 1]).astype(np.float32)
    sg, so = cq.init_states(pos, vel), orc.init_states(pos, vel)
    prev = [t.copy() for t in t0]
    for step in range(1, 7):
        cur = [t0[0] + np.float32([0, 0.08 * step, 0]), t0[1] + np.float32([0.12 * step, 0, 0])]
        models = [scenes.trs_model(c) for c in cur]
        g.update_transforms([1, 2], models)
        o.update_transforms([1, 2], models)
        plats = np.concatenate([scenes.platform_record(bv, models[k], prev[k], cur[k]) for k in range(2)])
        g.move_and_slide(sg, cq.default_params(), platforms=plats)
        o.move_and_slide(so, orc.default_params(), platforms=plats, order=g.order)
        assert sg.tobytes() == so.tobytes(), step
        prev = cur
    carried = sg["position"][:, 1] > 2.0
    assert carried.sum() > 20 and (sg["position"][on_top, 1] > pos[on_top, 1] + 0.2).mean() > 0.5
    with pytest.raises(cq.CQError):
        g.move_and_slide(sg, cq.default_params(), platforms=np.zeros(65, cq.PLATFORM))
    g.close()
    o.close()


@pytest.mark.gpu
def test_agents_capsule_capsule_ccd(cq, orc, scenes):
    """CQ_MAS_AGENTS: every character also sweeps against a pre-step snapshot of all the others (AgentSweepSolver,
    HitSelector; Systems.swift:1023-1091, 1378-1399, 1505-1590).  The GPU finds partners through a uniform grid, the
    oracle loops over everybody like the reference: states must stay identical over several steps of a dense crowd
    on a terrain, with the walk intent re-applied each step so that agents keep pressing into each other."""
    parts, half_world = scenes.terrain_scene(cells=96, cell=1.5, seed=5)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    rng = np.random.default_rng(11)
    n = 3000
    p = cq.default_params()
    p["radius"], p["half_height"] = 0.4, 0.5
    p["skin_width"], p["snap_distance"] = 0.08, 0.3
    # ~35% area coverage inside a square -> plenty of pairs within reach every step
    half = min(np.sqrt(n * np.pi * 0.4 ** 2 / 0.35) / 2, half_world - 4.0)
    x, z = rng.uniform(-half, half, n), rng.uniform(-half, half, n)
    y = scenes.terrain_height(x, z, 5) + np.float32(0.9 + 0.05) + rng.uniform(0, 0.3, n)
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n)
    walk = (np.stack([np.cos(ang), np.zeros(n), np.sin(ang)], axis=1) * rng.uniform(2, 9, (n, 1))).astype(np.float32)
    sg, so = cq.init_states(pos, walk), orc.init_states(pos, walk)
    agent_hits = 0
    for step in range(8):
        ghost = so.copy()
        g.move_and_slide(sg, p, flags=cq.MAS_APPLY_GRAVITY | cq.MAS_AGENTS)
        o.move_and_slide(so, p, flags=3, order=g.order, n_threads=8)
        assert sg.tobytes() == so.tobytes(), step
        o.move_and_slide(ghost, p, flags=1, order=g.order, n_threads=8)
        agent_hits += int((ghost["position"] != so["position"]).any(axis=1).sum())
        for s in (sg, so):
            s["velocity"][:, 0] = walk[:, 0]
            s["velocity"][:, 2] = walk[:, 2]
    assert agent_hits > n // 4  # the agent path really decided a good share of the moves
    # characters carried far from their snapshot cell by a platform (the grid's precomputed 3x3 ranges do not apply:
    # the kernel falls back to searching the cells its reach really covers)
    fv, fi = scenes.plane_mesh(80.0)
    flat = [scenes.part(fv, fi, scenes.trs_model((0, -3, 0)), entity_id=0)]
    g2, o2 = cq.CollisionQuery(flat), orc.OracleWorld(flat)
    m = 1500
    fpos = np.stack([rng.uniform(-16, 16, m), np.full(m, -3.0 + 0.9 + 0.05), rng.uniform(-16, 16, m)], axis=1).astype(np.float32)
    fwalk = walk[:m]
    plat = np.zeros(1, cq.PLATFORM)
    plat["aabb_min"], plat["aabb_max"], plat["delta"] = (-6, -4, -6), (6, -3, 6), (3.0, 0.0, 0.7)
    fg, fo = cq.init_states(fpos, fwalk), orc.init_states(fpos, fwalk)
    for step in range(3):
        g2.move_and_slide(fg, p, flags=3, platforms=plat)
        o2.move_and_slide(fo, p, flags=3, platforms=plat, order=g2.order, n_threads=8)
        assert fg.tobytes() == fo.tobytes(), step
    assert (np.abs(fg["position"][:, 0] - fpos[:, 0]) > 2.5).sum() > 50  # many were carried several cells away
    g2.close()
    o2.close()
    # a lone character is not its own obstacle, and an empty batch is fine
    one = cq.init_states(pos[:1], walk[:1])
    ref = orc.init_states(pos[:1], walk[:1])
    g.move_and_slide(one, p, flags=3)
    o.move_and_slide(ref, p, flags=3, order=g.order)
    assert one.tobytes() == ref.tobytes()
    g.move_and_slide(one[:0].copy(), p, flags=3)
    g.close()
    o.close()


@pytest.mark.gpu
def test_agent_separation_sequential_semantics(cq, orc, scenes):
    """cq_agent_separation_batch vs the oracle's literal sequential loop (AgentSeparationSystem, Systems.swift:1906-2210):
    a dense crowd on a terrain with walls (so that the blocking casts veto pushes), mixed mass weights including
    immovable agents, the full crowd step = move-and-slide with agent CCD, then separation, for several steps.  Every
    byte of every record must match although the GPU runs the turns along their conflict DAG, not one after another."""
    tparts, half_world = scenes.terrain_scene(cells=96, cell=1.5, seed=5)
    bv, bi = scenes.box_mesh(6.0)
    walls = [scenes.part(bv, bi, scenes.trs_model((x, float(scenes.terrain_height(np.float32([x]), np.float32([z]), 5)[0]), z)),
                         entity_id=10 + k) for k, (x, z) in enumerate([(-8.0, 3.0), (9.0, -6.0), (2.0, 12.0)])]
    parts = tparts + walls
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    rng = np.random.default_rng(21)
    n = 2500
    p = cq.default_params()
    p["radius"], p["half_height"] = 0.4, 0.5
    p["skin_width"], p["snap_distance"] = 0.08, 0.3
    half = np.sqrt(n * np.pi * 0.4 ** 2 / 0.45) / 2  # 45% footprint coverage: plenty of overlapping pairs
    x, z = rng.uniform(-half, half, n), rng.uniform(-half, half, n)
    y = scenes.terrain_height(x, z, 5) + np.float32(0.9 + 0.05) + rng.uniform(0, 0.2, n)
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n)
    walk = (np.stack([np.cos(ang), np.zeros(n), np.sin(ang)], axis=1) * rng.uniform(1, 7, (n, 1))).astype(np.float32)
    mass = rng.choice(np.float32([1.0, 3.0, 500.0, 0.0]), n, p=[0.6, 0.25, 0.1, 0.05])
    sg, so = cq.init_states(pos, walk), orc.init_states(pos, walk)
    total_pairs = 0
    for step in range(5):
        g.move_and_slide(sg, p, flags=3)
        o.move_and_slide(so, p, flags=3, order=g.order, n_threads=8)
        assert sg.tobytes() == so.tobytes(), ("move", step)
        g.agent_separation(sg, p, mass_weight=mass)
        total_pairs += o.agent_separation(so, p, mass_weight=mass, order=g.order, n_threads=8)
        assert sg.tobytes() == so.tobytes(), ("separation", step)
        for s in (sg, so):
            s["velocity"][:, 0] = walk[:, 0]
            s["velocity"][:, 2] = walk[:, 2]
    assert total_pairs > 2 * n
    # no world casts (setQuery never called), one sweep, default masses; and the trivial sizes
    a, b = sg.copy(), so.copy()
    g.agent_separation(a, p, iterations=1, use_query=False)
    o.agent_separation(b, p, iterations=1, use_query=False)
    assert a.tobytes() == b.tobytes()
    one = sg[:1].copy()
    g.agent_separation(one, p)
    assert one.tobytes() == sg[:1].tobytes()
    g.agent_separation(sg[:0].copy(), p)
    # a light agent shoved across several grid cells by a row of immovable lower-indexed agents BEFORE its own turn:
    # the schedule's drift assumption (R = 1 cell) fails, the sweep must be detected, restored and rerun with a wider
    # safety radius — and still equal the sequential loop
    fv, fi = scenes.plane_mesh(80.0)
    flat = [scenes.part(fv, fi, scenes.trs_model((0, -3, 0)), entity_id=0)]
    g2, o2 = cq.CollisionQuery(flat), orc.OracleWorld(flat)
    fy = -3.0 + 0.9 + 0.05
    rows = []
    for r in range(40):  # 40 independent rows, each: 8 immovable pushers then the light agent, then a bystander
        zr = 3.0 * r - 60.0
        rows += [[-0.3 + 0.4 * k, fy, zr] for k in range(8)] + [[0.0, fy, zr], [3.38, fy, zr + 0.3]]
    dmass = np.tile(np.float32([0.0] * 8 + [1.0, 1.0]), 40)
    for uq in (False, True):
        a, b = cq.init_states(rows), orc.init_states(rows)
        g2.agent_separation(a, p, mass_weight=dmass, iterations=2, use_query=uq)
        o2.agent_separation(b, p, mass_weight=dmass, iterations=2, use_query=uq, order=g2.order)
        assert a.tobytes() == b.tobytes(), uq
        assert (a["position"][8::10, 0] > 2.0).all()  # every light agent was pushed more than two cells
    g2.close()
    o2.close()
    # agents sorted along x (index correlates with space): the conflict DAG degenerates towards a chain, still exact
    order = np.argsort(sg["position"][:, 0], kind="stable")
    a, b = np.ascontiguousarray(sg[order]), np.ascontiguousarray(so[order])
    g.agent_separation(a, p, mass_weight=mass[order])
    o.agent_separation(b, p, mass_weight=mass[order], order=g.order, n_threads=8)
    assert a.tobytes() == b.tobytes()
    g.close()
    o.close()


@pytest.mark.gpu
def test_full_size_c3_sharding_invariance_and_sampled_parity(cq, orc, scenes):
    """BASELINE config C3 at its full size (1,048,576 characters, hulls scene): characters are independent units, so
    (1) stepping the whole batch equals stepping its two halves separately (what multi-GPU sharding relies on), and
    (2) any sample of it must equal the oracle stepping just that sample.  Two steps, every byte compared."""
    parts = scenes.mirror_scene(use_hulls=True)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    n = 1 << 20
    pos, vel = scenes.gen_c3_characters(n, seed=0xC0111DE3)
    p = cq.default_params()
    whole = cq.init_states(pos, vel)
    halves = whole.copy()
    rng = np.random.default_rng(17)
    pick = np.sort(rng.choice(n, 20000, replace=False))
    sample = np.ascontiguousarray(orc.init_states(pos[pick], vel[pick]))
    for step in range(2):
        g.move_and_slide(whole, p)
        a, b = np.ascontiguousarray(halves[: n // 2]), np.ascontiguousarray(halves[n // 2:])
        g.move_and_slide(a, p)
        g.move_and_slide(b, p)
        halves = np.concatenate([a, b])
        assert whole.tobytes() == halves.tobytes(), step
        o.move_and_slide(sample, p, order=g.order, n_threads=8)
        assert whole[pick].tobytes() == sample.tobytes(), step
    assert whole["grounded"].mean() > 0.99
    g.close()
    o.close()


@pytest.mark.gpu
def test_full_size_c4_terrain_sweeps_properties(cq, orc, scenes):
    """BASELINE config C4 at its full size (9,999,392-triangle terrain, 8,388,608 blocking sweeps): the device LBVH at
    10 M triangles, checked through size-independent properties — (1) idempotence: the same batch twice gives the same
    bytes; (2) sharding invariance: two half batches = the whole batch; (3) extending a sweep beyond its contact
    (delta x 2, same direction) never delays the contact and keeps it (same triangle, toi within 1e-4 rel / 1e-5 abs) for
    > 99.9% of the hits; (4) a 20 k sample
    is bit-identical to the oracle casting against the same 10 M triangles (its own median-split BVH)."""
    parts, half = scenes.terrain_scene()
    g = cq.CollisionQuery(parts)
    assert g.info()["n_static_triangles"] == 9999392
    n = 1 << 23
    q = scenes.gen_c4_casts(n, half)
    whole = g.capsuleCastBlocking(q)
    assert whole.tobytes() == g.capsuleCastBlocking(q).tobytes()
    a, b = g.capsuleCastBlocking(np.ascontiguousarray(q[: n // 2])), g.capsuleCastBlocking(np.ascontiguousarray(q[n // 2:]))
    assert whole.tobytes() == np.concatenate([a, b]).tobytes()
    hit = whole["triangle_index"] >= 0
    assert 0.2 < hit.mean() < 1.0
    q2 = q.copy()
    q2["delta"] *= np.float32(2.0)
    longer = g.capsuleCastBlocking(q2)
    assert (longer["triangle_index"][hit] >= 0).all()
    # The first contact along the same ray cannot move LATER.  It can move a little earlier: conservative advancement
    # samples t_k = t_(k-1) + max(dist - r, minAdvance) and stops at |delta|, so a contact inside the last step before
    # |delta| is seen only by the longer sweep (0.02% of the hits; the oracle shows the same on a small terrain).
    lt, st = longer["toi"][hit], whole["toi"][hit]
    assert (lt <= st * np.float32(1 + 1e-4) + np.float32(1e-5)).all()
    same = np.isclose(lt, st, rtol=1e-4, atol=1e-5)
    assert same.mean() > 0.999
    assert (longer["triangle_index"][hit][same] == whole["triangle_index"][hit][same]).mean() > 0.999
    pick = np.sort(np.random.default_rng(23).choice(n, 20000, replace=False))
    o = orc.OracleWorld(parts)
    ref = o.capsule_cast(np.ascontiguousarray(q[pick]), 1, g.order, n_threads=8)
    assert whole[pick].tobytes() == ref.tobytes()
    g.close()
    o.close()


def test_device_entry_points_on_many_streams(cq, scenes):
    """The *_device twins are asynchronous on the caller's stream and share the world's scratch blocks (node stacks, unit
    order).  Eight streams each enqueue sweeps, rays and a move-and-slide step on ONE world without any synchronisation
    in between: more launches in flight than scratch regions (4), so the cross-stream guard of `scratch_acquire` has to
    order the reuse.  Every stream's results must equal the synchronous host-pointer calls byte for byte."""
    import torch
    parts = scenes.mirror_scene(use_hulls=False)
    g = cq.CollisionQuery(parts)
    lo, hi = scenes.scene_aabb(parts[1:])
    dev = torch.device("cuda", 0)
    n_streams, n = 8, 6000
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    casts = [scenes.gen_casts(n, lo, hi, seed=900 + k, radius=0.4 + 0.1 * k) for k in range(n_streams)]
    rays = [scenes.gen_rays(4 * n, lo, hi, seed=950 + k) for k in range(n_streams)]
    chars = []
    for k in range(n_streams):
        pos, vel = scenes.gen_c3_characters(n, seed=970 + k)
        chars.append(cq.init_states(pos, vel))

    def up(a):
        return torch.from_numpy(np.frombuffer(a.tobytes(), np.uint8).copy()).to(dev)

    d_casts, d_rays, d_chars = [up(a) for a in casts], [up(a) for a in rays], [up(a) for a in chars]
    d_cast_out = [torch.zeros(n * cq.CAST_HIT.itemsize, dtype=torch.uint8, device=dev) for _ in range(n_streams)]
    d_ray_out = [torch.zeros(4 * n * cq.RAY_HIT.itemsize, dtype=torch.uint8, device=dev) for _ in range(n_streams)]
    torch.cuda.synchronize()
    params = cq.default_params()
    for rep in range(2):  # the second round reuses every scratch region from another stream
        for k in range(n_streams):
            s = streams[(k + 3 * rep) % n_streams].cuda_stream
            g.capsule_cast_device(d_casts[k].data_ptr(), n, cq.CAST_BLOCKING, d_cast_out[k].data_ptr(), s)
            g.raycast_device(d_rays[k].data_ptr(), 4 * n, d_ray_out[k].data_ptr(), s)
            if rep == 0:
                g.move_and_slide_device(d_chars[k].data_ptr(), n, params, stream=s)
    torch.cuda.synchronize()
    for k in range(n_streams):
        want = g.capsuleCastBlocking(casts[k])
        assert d_cast_out[k].cpu().numpy().tobytes() == want.tobytes(), ("cast", k)
        assert d_ray_out[k].cpu().numpy().tobytes() == g.raycast(rays[k]).tobytes(), ("ray", k)
        ref = chars[k].copy()
        g.move_and_slide(ref, params)
        assert d_chars[k].cpu().numpy().tobytes() == ref.tobytes(), ("move_and_slide", k)
        assert (want["triangle_index"] >= 0).sum() > n // 20
    g.close()


def test_mesh_upload_many_parts_and_index_errors(cq, orc, scenes):
    """Mesh upload (csrc/cq_assemble.h + k_expand_verts / k_expand_tris): a world of 400 small parts (staged copies, both
    sets, empty parts, trailing partial index triples) plus one part above the staging limit (copied straight from the
    caller's arrays) must give the reference's soup — vertices, indices, layers, part of every triangle — and identical
    query results; an index outside its part's vertex range is refused with the part and the index named, also when it
    sits in the dynamic set or deep inside a large part, and leaves no half-built world behind."""
    rng = np.random.default_rng(31)
    bv, bi = scenes.box_mesh(1.0)
    parts = []
    for k in range(400):
        pos = tuple(rng.uniform(-20, 20, 3))
        idx = bi if k % 7 else np.concatenate([bi, bi[:2]])  # a trailing partial triple is ignored
        v = bv if k % 11 else np.zeros((0, 3), np.float32)   # an entity without vertices ...
        idx = idx if k % 11 else np.zeros(0, np.uint32)      # ... and without triangles
        parts.append(scenes.part(v, idx, scenes.trs_model(pos, scale=tuple(rng.uniform(0.5, 3.0, 3))), layer=1 << (k % 5),
                                 is_dynamic=(k % 3 == 0), entity_id=k))
    tv, ti, _ = scenes.terrain_mesh(cells=130, cell=0.5)  # 17,161 vertices / 33,800 triangles: above CQ_STAGE_LIMIT
    parts.insert(200, scenes.part(tv, ti, scenes.trs_model((0, -22, 0)), layer=32, entity_id=1000))
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    for which in (0, 1):
        a, b = g.read_soup(which), o.read_soup(which)
        assert a["positions"].tobytes() == b["positions"].tobytes() and np.array_equal(a["indices"], b["indices"])
        assert np.array_equal(a["layers"], b["layers"]) and np.array_equal(a["parts"], b["parts"])
    q = scenes.gen_casts(4000, [-20, -22, -20], [20, 20, 20], seed=4, len_range=(0.5, 6.0))
    q["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 1, 6, 32]), len(q))
    assert g.capsuleCast(q).tobytes() == o.capsule_cast(q, 0, g.order).tobytes()
    g.close()
    o.close()

    def refused(bad_parts, part, index, n_verts):
        with pytest.raises(cq.CQError) as e:
            cq.CollisionQuery(bad_parts)
        assert f"part {part} index {index} out of range ({n_verts} vertices)" in str(e.value), str(e.value)

    bad = [dict(p) for p in parts]
    bad[5]["indices"] = bad[5]["indices"].copy()
    bad[5]["indices"][4] = 8                      # box: 8 vertices -> 8 is the first invalid index
    refused(bad, 5, 8, 8)
    bad = [dict(p) for p in parts]
    bad[3]["indices"] = bad[3]["indices"].copy()   # part 3 is in the dynamic set
    bad[3]["indices"][10] = 4_000_000_000
    refused(bad, 3, 4_000_000_000, 8)
    bad = [dict(p) for p in parts]
    bad[200]["indices"] = bad[200]["indices"].copy()
    bad[200]["indices"][3 * 20_000 + 2] = len(tv)  # deep inside the directly copied part
    bad[300]["indices"] = bad[300]["indices"].copy()
    bad[300]["indices"][0] = 99                   # a later part is bad too: the first one is reported
    refused(bad, 200, len(tv), len(tv))
    g = cq.CollisionQuery(parts[:3])               # the library is still usable after the refusals
    assert (g.info()["n_static_triangles"], g.info()["n_dynamic_triangles"]) == (24, 0)  # part 0 is an empty dynamic entity
    g.close()


def test_resident_crowd_matches_full_record_steps(cq, scenes):
    """cq_crowd_* (records resident in HBM, velocities in / poses out) against cq_move_and_slide_batch_ex on the same
    records: every pose field after every step and all 168 bytes of every record at the end, bit for bit — with fresh
    velocities, with kept velocities (NULL), with a pushing platform, with the characters colliding with each other
    (CQ_MAS_AGENTS: one chunk, a second small crowd) and across the chunked pipeline (CQ_CHUNK-sized pieces on alternating streams)."""
    parts = scenes.mirror_scene(use_hulls=True)
    g = cq.CollisionQuery(parts)
    rng = np.random.default_rng(12)
    n = 300_000  # three 128 k chunks, the last one ragged
    pos, vel = scenes.gen_c3_characters(n, seed=77)
    ref = cq.init_states(pos, vel)
    crowd = cq.Crowd(g, ref)
    assert crowd.n == n and crowd.device_states
    params = cq.default_params()
    pose = np.zeros(n, cq.CROWD_POSE)
    plat = np.zeros(1, cq.PLATFORM)
    plat["aabb_min"], plat["aabb_max"], plat["delta"] = (-12, -3.2, 2), (-8, -2.9, 6), (0.05, 0.0, 0.02)
    for step in range(6):
        fresh = step % 3 != 2
        flags = cq.MAS_APPLY_GRAVITY
        platforms = plat if step == 3 else None
        v = None
        if fresh:  # what PhysicsIntentSystem would do between two steps: steer the current velocity
            v = np.ascontiguousarray(ref["velocity"] + rng.normal(0.0, 0.5, (n, 3)) * [1.0, 0.0, 1.0])
            ref["velocity"] = v
        g.move_and_slide(ref, params, flags=flags, platforms=platforms)
        crowd.step(v, params, flags=flags, platforms=platforms, pose_out=pose)
        for f in ("position", "velocity", "ground_triangle_index", "grounded", "grounded_near", "ground_sliding"):
            assert np.array_equal(pose[f], ref[f]), (step, f)
    crowd.step(None, params, pose_out=None)  # nothing copied back
    g.move_and_slide(ref, params)
    assert crowd.read().tobytes() == ref.tobytes()
    # write() replaces the records; an empty crowd is a no-op
    fresh_states = cq.init_states(pos[::-1].copy(), vel[::-1].copy())
    crowd.write(fresh_states)
    crowd.step(None, params, pose_out=pose)
    g.move_and_slide(fresh_states, params)
    assert np.array_equal(pose["position"], fresh_states["position"]) and crowd.read().tobytes() == fresh_states.tobytes()
    # characters that collide with each other: the whole crowd is one chunk
    pos2, vel2 = scenes.gen_c3_characters(600, seed=78)
    ref2 = cq.init_states(pos2, vel2)
    crowd2, pose2 = cq.Crowd(g, ref2), np.zeros(600, cq.CROWD_POSE)
    for step in range(4):
        v = np.ascontiguousarray(ref2["velocity"] + rng.normal(0.0, 0.5, (600, 3)) * [1.0, 0.0, 1.0])
        ref2["velocity"] = v
        g.move_and_slide(ref2, params, flags=cq.MAS_APPLY_GRAVITY | cq.MAS_AGENTS)
        crowd2.step(v, params, flags=cq.MAS_APPLY_GRAVITY | cq.MAS_AGENTS, pose_out=pose2)
        assert np.array_equal(pose2["position"], ref2["position"]) and np.array_equal(pose2["velocity"], ref2["velocity"])
    assert crowd2.read().tobytes() == ref2.tobytes()
    crowd2.close()
    empty = cq.Crowd(g, cq.init_states(np.zeros((0, 3), np.float32)))
    empty.step(None, params)
    empty.close()
    crowd.close()
    g.close()


def test_group_and_multi_world_on_the_visible_gpus(cq, orc, scenes):
    """cq_group_* / cq_world_create_multi on however many GPUs this box shows (1 in the driver's test run, more under
    `gpurun --gpus N`): a local group, the world replicated on every device, host batches sharded over the replicas must
    equal the single-GPU calls byte for byte; the NCCL gather of device-resident shards (ragged on purpose) must put the
    whole batch, in unit order, on every device; and a process-per-GPU group of one rank gathers through NCCL as well."""
    import torch
    n_gpus = torch.cuda.device_count()
    parts = scenes.mirror_scene(use_hulls=False)
    lo, hi = scenes.scene_aabb(parts[1:])
    group = cq.Group.local(n_gpus)
    assert group.size == n_gpus
    mw = cq.MultiWorld(group, parts)
    g = cq.CollisionQuery(parts)
    q = scenes.gen_casts(30011, lo, hi, seed=401)  # a prime count: ragged shards on any group size > 1
    want = g.capsuleCastBlocking(q)
    assert mw.capsule_cast(q, cq.CAST_BLOCKING).tobytes() == want.tobytes()
    rays = scenes.gen_rays(20003, lo, hi, seed=402)
    assert mw.raycast(rays).tobytes() == g.raycast(rays).tobytes()
    pos, vel = scenes.gen_c3_characters(10007, seed=403)
    a, b = cq.init_states(pos, vel), cq.init_states(pos, vel)
    for _ in range(3):
        mw.move_and_slide(a, cq.default_params())
        g.move_and_slide(b, cq.default_params())
    assert a.tobytes() == b.tobytes()
    with pytest.raises(cq.CQError):
        mw.move_and_slide(a, cq.default_params(), flags=cq.MAS_APPLY_GRAVITY | cq.MAS_AGENTS)
    # device-resident shards + gather
    rec = cq.CAST_HIT.itemsize
    n = len(q)
    local, whole = [], []
    for i in range(n_gpus):
        s0, s1 = cq.shard_range(n, n_gpus, i)
        dev = torch.device("cuda", i)
        local.append(torch.from_numpy(np.frombuffer(want[s0:s1].tobytes(), np.uint8).copy()).to(dev))
        whole.append(torch.zeros(n * rec, dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()
    group.gather_records_local([t.data_ptr() for t in local], n, rec, [t.data_ptr() for t in whole])
    group.synchronize()
    for i in range(n_gpus):
        assert whole[i].cpu().numpy().tobytes() == want.tobytes(), i
    mw.close()
    group.close()
    # one process per GPU, here a group of one rank: NCCL all-gather straight into the result buffer
    torch.cuda.set_device(0)
    rg = cq.Group.rank(1, 0, cq.Group.unique_id())
    src = torch.from_numpy(np.frombuffer(want.tobytes(), np.uint8).copy()).to("cuda:0")
    dst = torch.zeros_like(src)
    st = torch.cuda.Stream()
    rg.gather_records(src.data_ptr(), n, rec, dst.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert bool(torch.equal(src, dst))
    rg.close()
    g.close()


@pytest.mark.parametrize("order", ["reference", "canonical"])
def test_dirty_subtree_refit_of_small_parts(cq, orc, scenes, order):
    """updateStaticTransforms / updateDynamicTransforms of ONE small entity inside larger sets (BVH.refit,
    CollisionQuery.swift:528-575: only the ancestors of the moved leaves): the library's dirty-subtree path (fewer than a
    quarter of a set's triangles moved) must leave exactly the boxes and answers a full refit / the reference leaves — the
    LBVH, its 4-wide collapse and, in reference order, the reference's own tree that the ray kernel walks."""
    tv, ti, half = scenes.terrain_mesh(cells=120, cell=1.0)
    bv, bi = scenes.box_mesh(3.0)
    rng = np.random.default_rng(5)
    parts = [scenes.part(tv, ti, entity_id=0)]
    spots = rng.uniform(-40, 40, (9, 2))
    for k, (x, z) in enumerate(spots):
        y = float(scenes.terrain_height(np.float32([x]), np.float32([z]))[0]) + 1.0
        parts.append(scenes.part(bv, bi, scenes.trs_model((x, y, z)), is_dynamic=(k >= 6), layer=2, entity_id=1 + k))
    g = cq.CollisionQuery(parts, order=cq.ORDER_REFERENCE if order == "reference" else cq.ORDER_CANONICAL)
    o = orc.OracleWorld(parts)
    lo, hi = np.float32([-55, -10, -55]), np.float32([55, 20, 55])
    casts = scenes.gen_casts(6000, lo, hi, seed=61, radius=0.6, half_height=0.7, expand=0.0)
    rays = scenes.gen_rays(12000, lo, hi, seed=62, expand=0.0)
    caps = scenes.gen_capsules(4000, lo, hi, seed=63, expand=0.0)
    for step in range(1, 5):
        moved = {2: 2, 8: 8} if step % 2 else {3: 3, 7: 7, 9: 9}  # a static and a dynamic box (ids), alternating
        ids, models = [], []
        for eid in moved:
            x, z = spots[eid - 1] + rng.uniform(-3, 3, 2)
            y = float(scenes.terrain_height(np.float32([x]), np.float32([z]))[0]) + 1.0 + 0.3 * step
            ids.append(eid)
            models.append(scenes.trs_model((x, y, z), scenes.quat_angle_axis(0.4 * step, (0, 1, 0))))
        g.update_transforms(ids, models)
        o.update_transforms(ids, models)
        assert o.check_bvh(0) and o.check_bvh(1)
        for which in (0, 1):
            a, b = g.read_soup(which), o.read_soup(which)
            assert np.array_equal(a["positions"], b["positions"]) and np.array_equal(a["aabbs"], b["aabbs"]), (step, which)
        assert g.capsuleCastBlocking(casts).tobytes() == o.capsule_cast(casts, 1, g.order).tobytes(), step
        assert g.raycast(rays).tobytes() == o.raycast(rays, g.order).tobytes(), step
        got, cnt, ov = g.capsuleOverlapAll(caps, 8)
        ref, rcnt, rov = o.capsule_overlap_all(caps, 8, g.order)
        if g.order == orc.ORDER_REFERENCE:
            got = _caller_order(got, cnt)
        assert np.array_equal(cnt, rcnt) and got.tobytes() == ref.tobytes(), step
    assert (g.raycast(rays)["triangle_index"] >= g.info()["n_static_triangles"]).any()  # dynamic boxes are really hit
    g.close()
    o.close()


def test_full_size_c2_every_sweep_against_the_reference_order(cq, orc, scenes):
    """BASELINE config C2 at its full size: all 65,536 sweeps against the Semla mesh, every byte of every hit record equal
    to the reference-order oracle's (which walks the reference's own BVH in its own order) — ties included — and every
    difference to the canonical rule carries the TIE flag."""
    parts = scenes.semla_scene(use_hulls=False)
    g, o = cq.CollisionQuery(parts), orc.OracleWorld(parts)
    lo, hi = scenes.scene_aabb(parts[1:])
    q = scenes.gen_casts(65536, lo, hi, seed=0xC0111DE2)
    got, flags = g.capsuleCast(q, with_flags=True)
    ref = o.capsule_cast(q, 0, orc.ORDER_REFERENCE, 8)
    assert got.tobytes() == ref.tobytes()
    can = o.capsule_cast(q, 0, orc.ORDER_CANONICAL, 8)
    diff = got["triangle_index"] != can["triangle_index"]
    assert 0.2 < diff.mean() < 0.4 and (flags[diff] & cq.HIT_TIE).all() and np.array_equal(got["toi"], can["toi"])
    g.close()
    o.close()


def test_full_size_c5_rays_properties(cq, orc, scenes):
    """BASELINE config C5 at its full size (16,777,216 rays against the three meshes merged, the mirror in the dynamic
    set), through size-independent properties in both order rules: (1) idempotence; (2) sharding invariance (two halves =
    the whole batch); (3) a hit stays the hit when maxDistance is cut to just beyond it and disappears when cut to just
    before it; (4) the canonical answer is never farther than the reference-order one, and where the two rules name
    different triangles it is an exact tie or a hit the reference's slab test lost; (5) a 100 k sample is byte-identical to
    the oracle in the world's order after a refit of the mirror."""
    parts = scenes.merged_scene(mirror_dynamic=True)
    a = scenes.load_mirror_fixture()
    t, q0, s = scenes.transform_from_matrix(scenes.mirror_model(a["transform"]))
    model = scenes.trs_model(t, scenes.quat_mul(scenes.quat_angle_axis(np.radians(17.0), (0, 1, 0)), q0), s)
    lo, hi = scenes.scene_aabb(parts[1:])
    n = 1 << 24
    rays = scenes.gen_rays(n, lo, hi, seed=0xC0111DE5, max_distance=100.0, expand=5.0, y_range=(0.0, 12.0))
    out = {}
    o = orc.OracleWorld(parts)
    o.update_transforms([parts[-1]["entity_id"]], [model])
    pick = np.sort(np.random.default_rng(29).choice(n, 100_000, replace=False))
    for order in (cq.ORDER_REFERENCE, cq.ORDER_CANONICAL):
        g = cq.CollisionQuery(parts, order=order)
        g.update_transforms([parts[-1]["entity_id"]], [model])
        whole = g.raycast(rays)
        assert whole.tobytes() == g.raycast(rays).tobytes()
        halves = np.concatenate([g.raycast(np.ascontiguousarray(rays[: n // 2])), g.raycast(np.ascontiguousarray(rays[n // 2:]))])
        assert whole.tobytes() == halves.tobytes()
        hit = whole["triangle_index"] >= 0
        assert 0.2 < hit.mean() < 0.8
        sub = np.flatnonzero(hit)[:: 64]
        near, far = rays[sub].copy(), rays[sub].copy()
        far["max_distance"] = whole["distance"][sub] * np.float32(1.001) + np.float32(1e-4)
        near["max_distance"] = whole["distance"][sub] * np.float32(0.999) - np.float32(1e-4)
        kept = g.raycast(far)
        assert np.array_equal(kept["triangle_index"], whole["triangle_index"][sub]) and np.array_equal(kept["distance"], whole["distance"][sub])
        cut = g.raycast(near)
        assert ((cut["triangle_index"] < 0) | (cut["distance"] < whole["distance"][sub])).all()
        assert whole[pick].tobytes() == o.raycast(np.ascontiguousarray(rays[pick]), order, 8).tobytes()
        out[order] = whole
        g.close()
    ref, can = out[cq.ORDER_REFERENCE], out[cq.ORDER_CANONICAL]
    both = (ref["triangle_index"] >= 0) & (can["triangle_index"] >= 0)
    assert (can["distance"][both] <= ref["distance"][both]).all() and not ((ref["triangle_index"] >= 0) & (can["triangle_index"] < 0)).any()
    diff = ref["triangle_index"] != can["triangle_index"]
    tie = diff & both & (ref["distance"] == can["distance"])
    lost = diff & (can["triangle_index"] >= 0) & ((ref["triangle_index"] < 0) | (ref["distance"] > can["distance"]))
    assert not (diff & ~tie & ~lost).any() and diff.mean() < 0.01
    o.close()


@pytest.mark.gpu
def test_per_triangle_materials_match_the_oracle(cq, orc, scenes):
    """StaticMeshComponent.triangleMaterials (CollisionQuery.swift:363-396, 464-469) through cq_world_options: random
    friction / flattenGround per triangle of a rolling terrain (static set) and of a tilted dynamic slab, one part whose
    array has the wrong length (ignored, as in the reference), degenerate triangles that take their entries with them.
    materialForTriangle and 40 move-and-slide steps (slope friction + flattened normals read the material of the ground
    triangle) must equal the oracle's field for field, in both order rules."""
    rng = np.random.default_rng(77)
    tparts, half = scenes.terrain_scene(cells=48, cell=2.0)
    terr = dict(tparts[0])
    nt = len(terr["indices"]) // 3

    def rows(n):
        return np.stack([rng.uniform(0.0, 1.2, n), rng.uniform(0.0, 0.9, n), rng.integers(0, 2, n)], 1).astype(np.float32)

    terr["triangle_materials"] = rows(nt)
    sv = np.array([[-6, 0, -6], [6, 0, -6], [-6, 0, 6], [6, 0, 6], [0, 0, -6]], np.float32)
    si = np.array([0, 2, 1, 0, 1, 4, 1, 2, 3], np.uint32)  # (0,1,4) is collinear: dropped by the filter
    tilt = scenes.quat_angle_axis(np.radians(30.0), (0, 0, 1))
    slab = dict(scenes.part(sv, si, scenes.trs_model((10.0, 14.0, 10.0), tilt), is_dynamic=True, entity_id=5, mu_s=0.2, mu_k=0.1),
                triangle_materials=np.float32([[0.05, 0.02, 0], [9, 9, 1], [2.5, 2.0, 0]]))  # icy | (dropped) | sticky
    wrong = dict(scenes.part(sv, si, scenes.trs_model((-12.0, 15.0, -8.0)), entity_id=6, mu_s=0.33, mu_k=0.22),
                 triangle_materials=rows(2))  # 2 entries for 3 triangles
    parts = [terr, slab, wrong]
    pos = np.stack([rng.uniform(-30, 30, 3000), np.zeros(3000), rng.uniform(-30, 30, 3000)], 1).astype(np.float32)
    pos[:, 1] = scenes.terrain_height(pos[:, 0], pos[:, 2]) + 2.6
    pos[:200] = np.float32([10, 17.0, 10]) + rng.uniform(-4, 4, (200, 3)).astype(np.float32) * np.float32([1, 0.1, 1])
    pos[200:300] = np.float32([-12, 17.8, -8]) + rng.uniform(-4, 4, (100, 3)).astype(np.float32) * np.float32([1, 0.05, 1])
    vel = np.zeros_like(pos)
    vel[300:, 0], vel[300:, 2] = rng.uniform(-6, 6, 2700), rng.uniform(-6, 6, 2700)
    for order in (cq.ORDER_REFERENCE, cq.ORDER_CANONICAL):
        g, o = cq.CollisionQuery(parts, order=order), orc.OracleWorld(parts)
        inf = g.info()
        n_all = inf["n_static_triangles"] + inf["n_dynamic_triangles"]
        assert inf["n_dynamic_triangles"] == 2 and inf["n_static_triangles"] == nt + 2
        for t in list(range(0, n_all, 97)) + [nt - 1, nt, nt + 1, n_all - 2, n_all - 1, n_all, n_all + 5]:
            a, b = g.triangle_material(t), o.triangle_material(t)
            assert a["mu_s"] == b["mu_s"] and a["mu_k"] == b["mu_k"] and a["flatten_ground"] == b["flatten_ground"], (t, a, b)
        sg, so = cq.init_states(pos, vel), orc.init_states(pos, vel)
        pg, po = cq.default_params(), orc.default_params()
        slid = np.zeros(len(pos), bool)
        for step in range(40):
            g.move_and_slide(sg, pg)
            o.move_and_slide(so, po, order=order)
            for f in sg.dtype.names:
                if f != "_pad":
                    assert np.array_equal(sg[f], so[f]), (order, step, f, int((sg[f] != so[f]).sum()))
            slid |= sg["ground_sliding"] > 0
        assert sg["grounded"].mean() > 0.6 and slid[:200].any() and not slid[:200].all()  # icy and sticky halves of the slab
        flat = sg["grounded"].astype(bool) & (np.abs(sg["ground_normal"][:, 1] - 1.0) < 1e-7)
        assert flat.any() and (~flat & sg["grounded"].astype(bool)).any()  # flattened and unflattened ground normals both occur
        g.close()
        o.close()
    # a count without an array is refused with a message, not read
    keep = []
    arr = cq._mesh_parts([wrong], keep)
    opt = cq.WorldOptions()
    cq.lib().cq_world_options_default(ctypes.byref(opt))
    opt.n_triangle_materials = 1
    h = ctypes.c_void_p()
    rc = cq.lib().cq_world_create_ex(ctypes.byref(arr), 1, ctypes.byref(opt), ctypes.byref(h))
    assert rc != 0 and not h.value and b"triangle_materials" in cq.lib().cq_last_error()
