"""Pins the CPU oracle (oracle/cq_oracle.cpp) with analytic known answers.  The reference has no tests or
golden vectors for this path (GameTests/GameTests.swift is an empty template), so these closed-form cases,
the brute-force cross-check and the committed golden trajectory are what pin the restatement."""
import numpy as np
import pytest

F32_MAX = np.finfo(np.float32).max


def big_floor(scenes, y=0.0, size=200.0, eid=0, **kw):
    v, i = scenes.plane_mesh(size)
    return scenes.part(v, i, scenes.trs_model((0, y, 0)), entity_id=eid, **kw)


def wall_x(scenes, x=5.0, eid=1):
    """A large two-triangle wall in the plane x = const, facing -x."""
    v = np.array([[x, -50, -50], [x, -50, 50], [x, 50, 50], [x, 50, -50]], np.float32)
    return scenes.part(v, [0, 1, 2, 0, 2, 3], entity_id=eid)


# ---------------------------------------------------------------- primitives (CollisionQuery.swift:1396-1601)
def test_closest_point_on_triangle_regions(orc):
    a, b, c = (0, 0, 0), (2, 0, 0), (0, 2, 0)
    cases = [((0.5, 0.5, 3.0), (0.5, 0.5, 0), 9.0),  # face
             ((-1, -1, 0), (0, 0, 0), 2.0),  # vertex A
             ((3, -1, 0), (2, 0, 0), 2.0),  # vertex B
             ((-1, 3, 0), (0, 2, 0), 2.0),  # vertex C
             ((1, -2, 0), (1, 0, 0), 4.0),  # edge AB
             ((-2, 1, 0), (0, 1, 0), 4.0),  # edge AC
             ((2, 2, 0), (1, 1, 0), 2.0)]  # edge BC
    for p, want, d2 in cases:
        d, pt = orc.closest_point_on_triangle(p, a, b, c)
        assert np.allclose(pt, want, atol=1e-6) and d == pytest.approx(d2, rel=1e-6)


def test_segment_segment_distance(orc):
    d, c1, c2 = orc.segment_segment_distance_sq((0, 1, 0), (0, -1, 0), (1, 0, -1), (1, 0, 1))  # crossing at distance 1
    assert d == pytest.approx(1.0) and np.allclose(c1, (0, 0, 0)) and np.allclose(c2, (1, 0, 0))
    d, c1, c2 = orc.segment_segment_distance_sq((0, 1, 0), (0, -1, 0), (3, 5, 0), (3, 2, 0))  # endpoint-endpoint
    assert d == pytest.approx(9.0 + 1.0) and np.allclose(c1, (0, 1, 0)) and np.allclose(c2, (3, 2, 0))
    d, _, _ = orc.segment_segment_distance_sq((0, 0, 0), (0, 0, 0), (1, 0, 0), (1, 0, 0))  # both degenerate
    assert d == pytest.approx(1.0)
    d, c1, c2 = orc.segment_segment_distance_sq((0, 0, 0), (0, 0, 0), (2, -1, 0), (2, 1, 0))  # first degenerate
    assert d == pytest.approx(4.0) and np.allclose(c2, (2, 0, 0))


def test_segment_triangle_distance(orc):
    v0, v1, v2 = (-10, 0, -10), (10, 0, -10), (0, 0, 10)
    d, seg, tri = orc.segment_triangle_distance((0, 3.0, 0), 1.0, v0, v1, v2)  # axis above the face
    assert d == pytest.approx(2.0) and np.allclose(seg, (0, 2, 0)) and np.allclose(tri, (0, 0, 0))
    d, seg, tri = orc.segment_triangle_distance((0, 0.5, 0), 1.0, v0, v1, v2)  # axis pierces the face -> 0
    assert d == 0.0 and np.allclose(seg, tri) and np.allclose(tri, (0, 0, 0), atol=1e-6)
    d, seg, tri = orc.segment_triangle_distance((13, 0.0, -10), 1.0, v0, v1, v2)  # beside vertex v1
    assert d == pytest.approx(3.0) and np.allclose(tri, (10, 0, -10))


# ---------------------------------------------------------------- casts (CollisionQuery.swift:980-1359)
def test_vertical_drop_toi_is_gap_minus_capsule(orc, scenes):
    """Capsule dropped onto a floor: contact when bottom sphere touches, toi = y0 - floor - r - hh.
    refineTOI returns the upper end of 10 bisections of [lastSafeT, t], so toi is within span/1024 above."""
    w = orc.OracleWorld([big_floor(scenes, y=-3.0)])
    r, hh = 1.5, 1.0
    for y0 in (2.0, 7.5, 40.0):
        q = np.zeros(1, orc.CAST)
        q["from"], q["delta"] = (0.3, y0, -0.2), (0, -(y0 + 10), 0)
        q["radius"], q["half_height"], q["mask"], q["min_normal_y"] = r, hh, 0xFFFFFFFF, 0.5
        want = y0 - (-3.0) - r - hh
        for mode in (0, 1, 2):
            h = w.capsule_cast(q, mode, orc.ORDER_REFERENCE)[0]
            assert h["triangle_index"] in (0, 1)
            assert want - 1e-5 <= h["toi"] <= want + want / 1024 + 2e-5
            assert np.allclose(h["normal"], (0, 1, 0), atol=1e-5) and np.allclose(h["triangle_normal"], (0, 1, 0))
            assert h["position"][1] == pytest.approx(-3.0, abs=1e-5)  # position = point ON the triangle (:1342)


def test_horizontal_sweep_into_wall(orc, scenes):
    w = orc.OracleWorld([wall_x(scenes, 5.0, eid=0)])
    q = np.zeros(1, orc.CAST)
    q["from"], q["delta"], q["radius"], q["half_height"], q["mask"] = (0, 0, 0), (10, 0, 0), 0.5, 1.0, 0xFFFFFFFF
    h = w.capsule_cast(q, 0, orc.ORDER_REFERENCE)[0]
    want = 5.0 - 0.5
    assert want - 1e-5 <= h["toi"] <= want + want / 1024 + 2e-5
    assert np.allclose(h["normal"], (-1, 0, 0), atol=1e-5)
    assert np.allclose(h["triangle_normal"], (-1, 0, 0), atol=1e-6)  # flipped to agree with the contact normal
    # moving away: blocking mode rejects (dot(delta, normal) >= 0), plain mode has nothing ahead
    q["delta"] = (-3, 0, 0)
    assert w.capsule_cast(q, 1, orc.ORDER_REFERENCE)[0]["triangle_index"] == -1
    # sweep shorter than the gap -> nil ; zero-length sweep -> nil (CollisionQuery.swift:988)
    q["delta"] = (4.0, 0, 0)
    assert w.capsule_cast(q, 0, orc.ORDER_REFERENCE)[0]["triangle_index"] == -1
    q["delta"] = (0, 0, 0)
    assert w.capsule_cast(q, 0, orc.ORDER_REFERENCE)[0]["triangle_index"] == -1


def test_ground_mode_filters_steep_triangles_and_layer_mask(orc, scenes):
    parts = [wall_x(scenes, 2.0, eid=0), big_floor(scenes, y=-3.0, eid=1, layer=4)]
    w = orc.OracleWorld(parts)
    q = np.zeros(1, orc.CAST)
    q["from"], q["delta"], q["radius"], q["half_height"] = (1.0, 0, 0), (0, -5, 0), 1.5, 1.0
    q["mask"], q["min_normal_y"] = 0xFFFFFFFF, 0.5
    # the wall (triangles 0,1) is already overlapping at t=0 but is not ground; ground mode must skip it
    h = w.capsule_cast(q, 2, orc.ORDER_REFERENCE)[0]
    assert h["triangle_index"] in (2, 3) and h["toi"] == pytest.approx(0.5, abs=2e-3)
    assert w.capsule_cast(q, 0, orc.ORDER_REFERENCE)[0]["triangle_index"] in (0, 1)
    q["mask"] = 1  # floor is on layer 4 -> invisible
    assert w.capsule_cast(q, 2, orc.ORDER_REFERENCE)[0]["triangle_index"] == -1


# ---------------------------------------------------------------- overlap (CollisionQuery.swift:1119-1283)
def test_overlap_depth_is_radius_minus_distance(orc, scenes):
    w = orc.OracleWorld([big_floor(scenes, y=0.0)])
    c = np.zeros(3, orc.CAPSULE)
    c["from"] = [(0.25, 2.0, 0.5), (0.25, 3.0, 0.5), (0.25, 0.5, 0.5)]
    c["radius"], c["half_height"], c["mask"] = 1.5, 1.0, 0xFFFFFFFF
    h = w.capsule_overlap(c, orc.ORDER_REFERENCE)
    assert h["depth"][0] == pytest.approx(0.5, abs=1e-6) and np.allclose(h["normal"][0], (0, 1, 0))
    assert h["triangle_index"][1] == -1  # bottom of the capsule is 0.5 above the floor
    assert h["depth"][2] == pytest.approx(1.5) and np.allclose(h["normal"][2], (0, 1, 0))  # axis pierces: dist 0 -> tri normal
    hits, counts, ov = w.capsule_overlap_all(c, 8, orc.ORDER_REFERENCE)
    assert counts.tolist() == [1, 0, 2] or counts.tolist() == [2, 0, 2]
    assert not ov.any()


# ---------------------------------------------------------------- raycast (CollisionQuery.swift:916-978, 1575-1631)
def test_raycast_unit_triangle(orc, scenes):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    w = orc.OracleWorld([scenes.part(v, [0, 1, 2], entity_id=0)])
    r = np.zeros(6, orc.RAY)
    r["origin"] = [(0.25, 2, 0.25), (0.25, -2, 0.25), (0.75, 2, 0.75), (0.25, 2, 0.25), (0.25, 2, 0.25), (0.5, 2, 0.5)]
    r["direction"] = [(0, -1, 0), (0, 1, 0), (0, -1, 0), (0, -4, 0), (0, 1, 0), (0, -1, 0)]
    r["max_distance"], r["mask"] = 100.0, 0xFFFFFFFF
    r["max_distance"][3] = 0.4  # direction is not normalised: t is in units of |direction| -> hit at t = 0.5 > 0.4
    for order in (orc.ORDER_REFERENCE, orc.ORDER_CANONICAL):
        h = w.raycast(r, order)
        assert h["triangle_index"].tolist() == [0, 0, -1, -1, -1, 0]
        assert h["distance"][0] == pytest.approx(2.0) and np.allclose(h["position"][0], (0.25, 0, 0.25))
        assert np.allclose(h["normal"][0], (0, 1, 0)) and np.allclose(h["normal"][1], (0, -1, 0))  # opposes the ray
        assert h["distance"][5] == pytest.approx(2.0)  # on the hypotenuse u+v == 1 is still a hit


# ---------------------------------------------------------------- BVH (CollisionQuery.swift:496-707)
def test_reference_bvh_equals_brute_force_except_ties(orc, scenes):
    """Candidate sets are tree-independent; the only order-dependent outcome is WHICH triangle is named when
    two accepted candidates have bit-identical toi."""
    parts = scenes.mirror_scene(use_hulls=False)
    w = orc.OracleWorld(parts)
    assert w.check_bvh(0) and w.counts(0)["triangles"] == 14213  # 14,211 kept of 14,246 + 2 ground
    lo, hi = scenes.scene_aabb(parts[1:])
    q = scenes.gen_casts(1500, lo, hi, seed=5)
    a = w.capsule_cast(q, 0, orc.ORDER_REFERENCE)
    b = w.capsule_cast(q, 0, orc.ORDER_CANONICAL)
    assert np.array_equal(a["toi"], b["toi"])
    assert np.array_equal(a["triangle_index"] >= 0, b["triangle_index"] >= 0)
    same = a["triangle_index"] == b["triangle_index"]
    assert same.mean() > 0.5
    assert (b["triangle_index"][~same] < a["triangle_index"][~same]).all()  # canonical = smallest index of the tie
    r = scenes.gen_rays(1500, lo, hi, seed=6, expand=2.0)
    ra, rb = w.raycast(r, orc.ORDER_REFERENCE), w.raycast(r, orc.ORDER_CANONICAL)
    assert (ra["triangle_index"] != rb["triangle_index"]).mean() < 0.01


def test_degenerate_filter_is_absolute_in_world_units(orc, scenes):
    """|e1 x e2|^2 <= 1e-10 is tested AFTER the model transform (CollisionQuery.swift:385): the same mesh keeps
    3,515 triangles unscaled and 14,211 at the demo's x8 scale."""
    a = scenes.load_mirror_fixture()
    w1 = orc.OracleWorld([scenes.part(a["positions"], a["indices"], a["transform"])])
    assert w1.counts(0)["triangles"] == 3515
    w8 = orc.OracleWorld([scenes.part(a["positions"], a["indices"], scenes.mirror_model(a["transform"]))])
    assert w8.counts(0)["triangles"] == 14211


def test_refit_keeps_bvh_consistent(orc, scenes):
    parts = scenes.mirror_scene(use_hulls=True, mirror_dynamic=True)
    w = orc.OracleWorld(parts)
    a = scenes.load_mirror_fixture()
    t, q0, s = scenes.transform_from_matrix(scenes.mirror_model(a["transform"]))
    before = w.read_soup(1)["aabbs"].copy()
    model = scenes.trs_model(t + np.float32([1, 0, 0]), q0, s)
    w.update_transforms([1, 2], [model, model])
    after = w.read_soup(1)["aabbs"]
    assert w.check_bvh(1)
    assert np.allclose(after[:, 0], before[:, 0] + 1.0, atol=1e-5)


# ---------------------------------------------------------------- move-and-slide (Systems.swift:734-1821)
def test_character_lands_and_rests_on_floor(orc, scenes):
    w = orc.OracleWorld([big_floor(scenes, y=-3.0)])
    p = orc.default_params()
    s = orc.init_states([[0, 7.5, 0]])
    for _ in range(120):
        w.move_and_slide(s, p, order=orc.ORDER_REFERENCE)
    assert s["grounded"][0] == 1 and s["grounded_near"][0] == 1
    # resting height = floor + r + hh + groundSnapSkin
    assert s["position"][0][1] == pytest.approx(-3.0 + 1.5 + 1.0 + 0.05, abs=2e-3)
    assert abs(s["velocity"][0][1]) < 1e-6 and np.allclose(s["ground_normal"][0], (0, 1, 0), atol=1e-6)
    assert s["ground_triangle_index"][0] in (0, 1) and s["ground_distance"][0] == pytest.approx(0.05, abs=2e-3)


def test_character_walks_and_is_stopped_by_wall(orc, scenes):
    w = orc.OracleWorld([big_floor(scenes, y=-3.0, eid=0), wall_x(scenes, 6.0, eid=1)])
    p = orc.default_params()
    s = orc.init_states([[0, -0.45, 0]], [[12.5, 0, 0]])
    xs = []
    for _ in range(90):
        s["velocity"][0][0] = 12.5  # keep pushing (PhysicsIntentSystem would)
        w.move_and_slide(s, p, order=orc.ORDER_REFERENCE)
        xs.append(s["position"][0][0])
    assert xs[5] == pytest.approx(6 * 12.5 / 60, rel=1e-3)  # free walk: v*dt per step
    # stopped short of the wall by radius + skinWidth (contact skin 0.3 for side hits)
    assert 6.0 - 1.5 - 0.3 - 0.02 <= xs[-1] <= 6.0 - 1.5 + 1e-3
    assert s["grounded"][0] == 1 and s["side_contact_frames"][0] > 0 and s["manifold_count"][0] >= 1
    # (velocity is NOT zeroed while "sticky" at the wall: that branch returns before adjustVelocity, SYS:1320-1323)


def test_character_pushed_out_of_floor(orc, scenes):
    w = orc.OracleWorld([big_floor(scenes, y=-3.0)])
    s = orc.init_states([[0, -1.0, 0]])  # bottom at -3.5: 0.5 below the floor
    w.move_and_slide(s, orc.default_params(), order=orc.ORDER_REFERENCE)
    assert s["position"][0][1] >= -0.5 - 1e-4 and s["grounded"][0] == 1


def test_empty_world_free_fall(orc, scenes):
    w = orc.OracleWorld([])
    s = orc.init_states([[0, 0, 0]], [[1, 0, 0]])
    w.move_and_slide(s, orc.default_params(), order=orc.ORDER_REFERENCE)
    dt = np.float32(1.0 / 60.0)
    vy = -98.0 * float(dt)
    assert s["velocity"][0][1] == pytest.approx(vy) and s["position"][0][1] == pytest.approx(vy * float(dt), rel=1e-5)
    assert s["grounded"][0] == 0 and s["ground_distance"][0] == F32_MAX


def test_platform_carry_and_side_push(orc, scenes):
    """PlatformCarry.computeDelta (Systems.swift:644-732): a character standing on an elevator is carried by the
    platform's motion; a platform moving into a character's side pushes it horizontally; a platform that does not
    move (|delta|^2 < 1e-8) or is out of reach does nothing."""
    bv, bi = scenes.box_mesh(4.0)
    prev, cur = np.float32([0, 0, 0]), np.float32([0.05, 0.1, 0])  # elevator top at y = 2 -> 2.1
    model = scenes.trs_model(cur)
    parts = [big_floor(scenes, y=-3.0, eid=0), scenes.part(bv, bi, model, is_dynamic=True, entity_id=1)]
    w = orc.OracleWorld(parts)
    plat = scenes.platform_record(bv, model, prev, cur)
    p = orc.default_params()
    rest = 2.1 + 1.5 + 1.0 + 0.05  # standing on the moved top with groundSnapSkin
    s = orc.init_states([[0.0, rest - 0.1, 0.0], [30.0, -0.45, 0.0], [-(2.0 + 1.5 + 0.2) + 0.05, -0.45, 0.0]])
    before = s["position"].copy()
    w.move_and_slide(s, p, platforms=plat, order=orc.ORDER_REFERENCE)
    assert s["position"][0][0] == pytest.approx(before[0][0] + 0.05, abs=1e-5)  # carried sideways with the platform
    # carried up by 0.1 too, then one step of gravity (-98/60^2 = -0.027) before the ground probe catches it
    assert rest - 0.05 <= s["position"][0][1] <= rest + 0.01 and s["grounded"][0] == 1
    assert np.allclose(s["position"][1][[0, 2]], before[1][[0, 2]], atol=1e-6)  # far away: not carried
    s2 = orc.init_states([[-(2.0 + 1.5 + 0.2) + 0.05, -0.45, 0.0]])
    mover = plat.copy()
    mover["delta"] = (-0.1, 0, 0)  # moving toward -x, into the character standing left of it
    w.move_and_slide(s2, p, platforms=mover, order=orc.ORDER_REFERENCE)
    assert s2["position"][0][0] < before[2][0] - 0.05
    still = plat.copy()
    still["delta"] = (5e-5, 0, 0)
    s3 = orc.init_states([[0.0, rest, 0.0]])
    w.move_and_slide(s3, p, platforms=still, order=orc.ORDER_REFERENCE)
    assert s3["position"][0][0] == pytest.approx(0.0, abs=1e-6)


def test_capsule_capsule_sweep_known_answers(orc):
    """capsuleCapsuleSweep (Systems.swift:1505-1590) on cases with closed-form answers."""
    dims = np.float32([[1.5, 1.0, 1.5, 1.0]])
    # head-on along x, other static, 10 m apart: contact when the centres are rSum = 3 apart -> toi 7 of a 10 m move
    hit, toi, n = orc.capsule_capsule_sweep_batch([[0, 0, 0]], [[10, 0, 0]], [[10, 0, 0]], [[0, 0, 0]], dims)
    assert hit[0] == 1 and toi[0] == pytest.approx(7.0, rel=1e-6) and np.allclose(n[0], [-1, 0, 0])
    # both moving toward each other at the same speed: they meet halfway, toi measured along the mover's own path
    hit, toi, n = orc.capsule_capsule_sweep_batch([[0, 0, 0]], [[5, 0, 0]], [[10, 0, 0]], [[-5, 0, 0]], dims)
    assert hit[0] == 1 and toi[0] == pytest.approx(3.5, rel=1e-6)
    # moving away / passing at a lateral distance > rSum: no hit
    hit, _, _ = orc.capsule_capsule_sweep_batch([[0, 0, 0], [0, 0, 3.01]], [[-10, 0, 0], [20, 0, 0]], [[10, 0, 0]] * 2,
                                                [[0, 0, 0]] * 2, np.repeat(dims, 2, 0))
    assert hit.tolist() == [0, 0]
    # landing on top of another capsule: cap spheres meet when the centres are hSum + rSum = 5 apart vertically
    hit, toi, n = orc.capsule_capsule_sweep_batch([[0, 9, 0]], [[0, -6, 0]], [[0, 0, 0]], [[0, 0, 0]], dims)
    assert hit[0] == 1 and toi[0] == pytest.approx(4.0, rel=1e-6) and np.allclose(n[0], [0, 1, 0])
    # no relative motion: overlap test only, toi 0 with the separating normal
    hit, toi, n = orc.capsule_capsule_sweep_batch([[0, 0, 0], [0, 0, 0]], [[1, 0, 0]] * 2, [[2, 0, 0], [4, 0, 0]],
                                                  [[1, 0, 0]] * 2, np.repeat(dims, 2, 0))
    assert hit.tolist() == [1, 0] and toi[0] == 0.0 and np.allclose(n[0], [-1, 0, 0])


def test_agents_block_each_other(orc, scenes):
    """flags & 2 (every character a solid agent): two walkers heading for each other stop at rSum instead of passing
    through; a third one far away is unaffected; without the flag they walk through each other."""
    parts = [big_floor(scenes, y=-3.0, eid=0)]
    w = orc.OracleWorld(parts)
    p = orc.default_params()
    y = -3.0 + 2.5 + 0.05
    pos = [[-4.0, y, 0.0], [4.0, y, 0.0], [40.0, y, 0.0]]
    vel = [[6.0, 0, 0], [-6.0, 0, 0], [6.0, 0, 0]]
    solid, ghost = orc.init_states(pos, vel), orc.init_states(pos, vel)
    for _ in range(90):
        w.move_and_slide(solid, p, flags=3, order=orc.ORDER_REFERENCE)
        w.move_and_slide(ghost, p, flags=1, order=orc.ORDER_REFERENCE)
        solid["velocity"][:, 0] = ghost["velocity"][:, 0] = [6.0, -6.0, 6.0]  # the walk intent is re-applied every step
    assert solid["position"][0][0] == pytest.approx(-1.5, abs=0.02) and solid["position"][1][0] == pytest.approx(1.5, abs=0.02)
    assert solid["position"][2][0] == pytest.approx(ghost["position"][2][0]) and ghost["position"][2][0] > 48.0
    assert ghost["position"][0][0] > 4.0 and ghost["position"][1][0] < -4.0  # passed through each other
    assert solid["grounded"].tolist() == [1, 1, 1]


def test_agent_separation_known_answers(orc, scenes):
    """AgentSeparationSystem (Systems.swift:1906-2210) on cases with closed-form answers: minDist = rA + rB +
    min(separationMargin 0.2, skinWidth 0.3) = 3.2 for default capsules; the correction is split by inverse mass; closing
    velocity along the pair normal is removed the same way; agents separated in height do not interact; velocity is
    rounded through Float by the write-back; a wall behind one agent hands the whole correction to the other."""
    w = orc.OracleWorld([big_floor(scenes, y=-3.0, eid=0)])
    p = orc.default_params()
    y = -3.0 + 2.5 + 0.05
    s = orc.init_states([[0.0, y, 0.0], [2.6, y, 0.0]], [[1.0, 0, 0.5], [-1.0, 0, 0.5]])
    s["grounded"] = s["grounded_near"] = 1
    n_pairs = w.agent_separation(s, p, iterations=1, use_query=False)
    assert n_pairs == 1
    assert s["position"][:, 0] == pytest.approx([-0.3, 2.9], abs=1e-6)  # penetration 0.6 split evenly
    assert s["velocity"][:, 0] == pytest.approx([0.0, 0.0], abs=1e-6)    # closing speed 2 along x removed, half each
    assert s["velocity"][:, 2] == pytest.approx([0.5, 0.5])              # tangential part untouched
    # mass weights 500 : 1 -> the light one takes 500/501 of the correction
    s = orc.init_states([[0.0, y, 0.0], [2.6, y, 0.0]])
    w.agent_separation(s, p, mass_weight=[500.0, 1.0], iterations=1, use_query=False)
    assert s["position"][0][0] == pytest.approx(-0.6 / 501, abs=1e-6) and s["position"][1][0] == pytest.approx(2.6 + 0.6 * 500 / 501, abs=1e-5)
    # massWeight <= 0 -> invWeight 0: immovable; two immovable agents are skipped
    s = orc.init_states([[0.0, y, 0.0], [2.6, y, 0.0]])
    w.agent_separation(s, p, mass_weight=[0.0, 1.0], iterations=1, use_query=False)
    assert s["position"][0][0] == 0.0 and s["position"][1][0] == pytest.approx(3.2, abs=1e-6)
    s = orc.init_states([[0.0, y, 0.0], [2.6, y, 0.0]])
    assert w.agent_separation(s, p, mass_weight=[0.0, -1.0], iterations=1, use_query=False) == 0
    # separated in height by more than heightMargin: no interaction (hh = 1: [y-1, y+1] against [y+2.2-1, ...])
    s = orc.init_states([[0.0, y, 0.0], [1.0, y + 2.2, 0.0]])
    assert w.agent_separation(s, p, iterations=2, use_query=False) == 0
    # sequential (Gauss-Seidel) order: pair (0,1) moves 1 to 3.1 BEFORE pair (1,2) is measured, so (1,2) sees a
    # penetration of 0.2, not 0.1 — a simultaneous (Jacobi) resolution would end at [-0.1, 3.05, 6.15]
    s = orc.init_states([[0.0, y, 0.0], [3.0, y, 0.0], [6.1, y, 0.0]])
    w.agent_separation(s, p, iterations=1, use_query=False)
    assert s["position"][:, 0] == pytest.approx([-0.1, 3.0, 6.2], abs=1e-5)
    # velocity goes through Float even without any contact (linearVelocityF -> d3), position too
    s = orc.init_states([[0.1, y, 0.0], [50.0, y, 0.0]], [[0.1, 0, 0], [0, 0, 0]])
    s["velocity"][0][0] = 0.1  # not representable in float32
    w.agent_separation(s, p, use_query=False)
    assert s["velocity"][0][0] == float(np.float32(0.1)) and s["position"][0][0] == float(np.float32(0.1))
    # fewer than two agents: the system returns before touching anything
    one = orc.init_states([[0.1, y, 0.0]], [[0.1, 0, 0]])
    one["velocity"][0][0] = 0.1
    w.agent_separation(one, p)
    assert one["velocity"][0][0] == 0.1
    # a wall right behind agent 1: its share of the push is blocked (toi <= skinWidth, side normal) -> agent 0 takes it all
    bv, bi = scenes.box_mesh(8.0)
    wall = scenes.part(bv, bi, scenes.trs_model((2.6 + 1.5 + 0.05 + 4.0, 0.0, 0.0)), entity_id=1)
    w2 = orc.OracleWorld([big_floor(scenes, y=-3.0, eid=0), wall])
    s = orc.init_states([[0.0, y, 0.0], [2.6, y, 0.0]])
    s["grounded"] = s["grounded_near"] = 1
    w2.agent_separation(s, p, iterations=1, use_query=True, order=orc.ORDER_REFERENCE)
    assert s["position"][1][0] == pytest.approx(2.6, abs=1e-6) and s["position"][0][0] == pytest.approx(-0.6, abs=1e-5)
    assert s["grounded"].tolist() == [1, 1]


def test_agent_layer_size_independent_properties(orc, scenes):
    """Properties that hold at any crowd size (used at full bench sizes as well): (1) a capsule-capsule sweep that hits
    at toi keeps that toi when the mover's path is extended beyond the contact (same direction, other static);
    (2) without world casts and with equal masses every separation correction is equal and opposite, so the crowd's
    centroid in XZ does not move, heights never change, and no pair ends up closer than it started beyond float noise."""
    rng = np.random.default_rng(9)
    n = 20000
    dims = np.tile(np.float32([0.4, 0.5, 0.4, 0.5]), (n, 1))
    frm = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    other = (frm + rng.standard_normal((n, 3)) * 2).astype(np.float32)
    delta = ((other - frm) * rng.uniform(0.5, 1.2, (n, 1))).astype(np.float32)
    zero = np.zeros((n, 3), np.float32)
    h1, t1, _ = orc.capsule_capsule_sweep_batch(frm, delta, other, zero, dims)
    h2, t2, _ = orc.capsule_capsule_sweep_batch(frm, delta * np.float32(2), other, zero, dims)
    both = (h1 == 1) & (t1 > 1e-3)
    assert both.sum() > n // 4 and (h2[both] == 1).all()
    assert np.allclose(t1[both], t2[both], rtol=2e-4, atol=2e-5)
    w = orc.OracleWorld([big_floor(scenes, y=-3.0, eid=0)])
    p = orc.default_params()
    p["radius"], p["half_height"], p["skin_width"] = 0.4, 0.5, 0.08
    m = 4000
    half = np.sqrt(m * np.pi * 0.16 / 0.5) / 2
    pos = np.stack([rng.uniform(-half, half, m), np.full(m, -3.0 + 0.95), rng.uniform(-half, half, m)], axis=1).astype(np.float32)
    s = orc.init_states(pos)
    pairs = w.agent_separation(s, p, iterations=2, use_query=False)
    assert pairs > m
    after = s["position"]
    assert np.allclose(after[:, [0, 2]].mean(0), pos[:, [0, 2]].astype(np.float64).mean(0), atol=1e-4)
    assert np.array_equal(after[:, 1], pos[:, 1].astype(np.float64))
    moved = np.linalg.norm(after[:, [0, 2]] - pos[:, [0, 2]], axis=1)
    assert 0 < moved.max() < 2.0 and (moved > 0).mean() > 0.5


def test_oracle_matches_an_independent_transliteration(orc, scenes):
    """The C++ oracle against tests/independent_narrow_phase.py — a second restatement written separately, straight
    from CollisionQuery.swift:1285-1573, in scalar float32 Python.  Bit-for-bit: segmentTriangleDistance on random /
    touching / piercing / degenerate-edge cases, and whole conservative-advancement sweeps (toi, contact point, both
    normals) through the oracle's public capsuleCast against one triangle each."""
    import independent_narrow_phase as ind
    F = np.float32
    rng = np.random.default_rng(2024)
    n = 60000
    scale = 10.0 ** rng.uniform(-1.5, 1.0, (n, 1))
    tris = (rng.standard_normal((n, 9)) * scale).astype(np.float32)
    wts = rng.dirichlet((1, 1, 1), n)
    on_tri = (tris.reshape(n, 3, 3) * wts[:, :, None]).sum(1)
    kind = rng.integers(0, 4, n)
    centers = on_tri + rng.standard_normal((n, 3)) * scale * np.where(kind == 0, 3.0, np.where(kind == 1, 0.3, 1e-3))[:, None]
    centers[kind == 3] = on_tri[kind == 3] + [0, 1.5, 0]
    hh = rng.choice(np.float32([0.0, 1e-4, 0.5, 1.0]), n)
    deg = rng.random(n) < 0.03  # an edge shorter than 1e-3: the e <= 1e-6 branches of the segment-segment test
    tris[deg, 3:6] = tris[deg, 0:3] + (rng.standard_normal((deg.sum(), 3)) * 3e-4).astype(np.float32)
    centers = centers.astype(np.float32)
    od, oseg, otri = orc.segment_triangle_distance_batch(centers, hh, tris)
    pierced = 0
    for i in range(n):
        t = tris[i]
        d, s, q = ind.segment_triangle_distance(tuple(centers[i]), hh[i], tuple(t[0:3]), tuple(t[3:6]), tuple(t[6:9]))
        assert F(d).tobytes() == od[i].tobytes(), i
        assert np.float32(s).tobytes() == oseg[i].tobytes() and np.float32(q).tobytes() == otri[i].tobytes(), i
        pierced += d == 0
    assert 100 < pierced < n // 2
    # sweeps: 6000 triangles 200 m apart in one world, one capsule cast aimed at (or past) each
    m = 6000
    base = np.stack([200.0 * np.arange(m), np.zeros(m), np.zeros(m)], axis=1)
    tv = (rng.standard_normal((m, 3, 3)) * rng.uniform(0.5, 4.0, (m, 1, 1)) + base[:, None, :]).astype(np.float32)
    parts = [scenes.part(tv.reshape(-1, 3), np.arange(3 * m, dtype=np.uint32), entity_id=0)]
    w = orc.OracleWorld(parts)
    assert w.counts(0)["triangles"] == m  # nothing filtered: triangle index == its position
    q = np.zeros(m, scenes.CAST)
    wts = rng.dirichlet((1, 1, 1), m)
    target = (tv * wts[:, :, None]).sum(1) + rng.standard_normal((m, 3)) * 0.5
    start = target + rng.standard_normal((m, 3)) * rng.uniform(1.5, 6.0, (m, 1))
    q["from"] = start.astype(np.float32)
    q["delta"] = ((target - start) * rng.uniform(0.6, 1.6, (m, 1))).astype(np.float32)
    q["radius"] = rng.choice(np.float32([0.4, 1.5]), m)
    q["half_height"] = rng.choice(np.float32([0.0, 0.5, 1.0]), m)
    q["mask"] = 0xFFFFFFFF
    ref = w.capsule_cast(q, 0, orc.ORDER_REFERENCE)
    hits = 0
    for i in range(m):
        d = tuple(q["delta"][i])
        length = np.sqrt(ind.dot(d, d))
        direction = (d[0] / length, d[1] / length, d[2] / length)
        got = ind.sweep_capsule_triangle(tuple(q["from"][i]), direction, length, q["radius"][i], q["half_height"][i],
                                         tuple(tv[i, 0]), tuple(tv[i, 1]), tuple(tv[i, 2]))
        if got is None:
            assert ref["triangle_index"][i] == -1, i
            continue
        hits += 1
        toi, pos, nrm, tri_n, _ = got
        assert ref["triangle_index"][i] == i
        assert F(toi).tobytes() == ref["toi"][i].tobytes(), i
        assert np.float32(pos).tobytes() == ref["position"][i].tobytes(), i
        assert np.float32(nrm).tobytes() == ref["normal"][i].tobytes(), i
        assert np.float32(tri_n).tobytes() == ref["triangle_normal"][i].tobytes(), i
    assert m // 4 < hits < m


def test_oracle_capsule_pair_sweep_matches_independent_transliteration(orc):
    """capsuleCapsuleSweep (Systems.swift:1417-1590): the C++ oracle against the separately written Python version, bit for
    bit, on approaching / receding / resting / stacked / vertical-only relative motion."""
    import warnings
    import independent_narrow_phase as ind
    rng = np.random.default_rng(4242)
    n = 40000
    dims = np.stack([rng.choice(np.float32([0.3, 1.5]), n), rng.choice(np.float32([0.0, 0.6, 1.0]), n),
                     rng.choice(np.float32([0.3, 1.5]), n), rng.choice(np.float32([0.0, 0.6, 1.0]), n)], axis=1).astype(np.float32)
    frm = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    other = (frm + rng.standard_normal((n, 3)) * 3).astype(np.float32)
    delta = (rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-3, 0.7, (n, 1))).astype(np.float32)
    odelta = (rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-3, 0.7, (n, 1))).astype(np.float32)
    kind = rng.integers(0, 6, n)
    odelta[kind == 0] = delta[kind == 0]
    delta[kind == 1, 1] = odelta[kind == 1, 1] = 0
    other[kind == 2] = frm[kind == 2] + np.float32([0, 1, 0]) * rng.uniform(0, 6, ((kind == 2).sum(), 1)).astype(np.float32)
    aim = kind == 3
    delta[aim] = ((other[aim] - frm[aim]) * rng.uniform(0.2, 1.5, (aim.sum(), 1))).astype(np.float32)
    odelta[kind == 4] = 0
    for a in (delta, odelta):
        a[kind == 5, 0] = a[kind == 5, 2] = 0
    oh, ot, on = orc.capsule_capsule_sweep_batch(frm, delta, other, odelta, dims)
    hits = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # float32 overflow / divide warnings of numpy scalars: inf and NaN are part of the contract
        for i in range(n):
            got = ind.capsule_capsule_sweep(tuple(frm[i]), tuple(delta[i]), dims[i, 0], dims[i, 1], tuple(other[i]),
                                            tuple(odelta[i]), dims[i, 2], dims[i, 3])
            assert (got is not None) == bool(oh[i]), i
            if got is not None:
                hits += 1
                assert np.float32(got[0]).tobytes() == ot[i].tobytes(), i
                assert np.float32(got[1]).tobytes() == on[i].tobytes(), i
    assert n // 10 < hits < n


def _state_equal(a, b):
    """Every field of two STATE records; manifold arrays only up to manifold_count (slots beyond it are dead storage)."""
    for name in a.dtype.names:
        if name in ("manifold_triangles", "manifold_normals", "_pad"):
            continue
        if a[name].tobytes() != b[name].tobytes():
            return name
    cnt = int(a["manifold_count"])
    if a["manifold_triangles"][:cnt].tobytes() != b["manifold_triangles"][:cnt].tobytes():
        return "manifold_triangles"
    if a["manifold_normals"][:cnt].tobytes() != b["manifold_normals"][:cnt].tobytes():
        return "manifold_normals"
    return None


@pytest.mark.parametrize("scene", ["hulls", "ramps"])
def test_oracle_controller_matches_independent_transliteration(orc, scenes, scene):
    """KinematicMoveStopSystem's per-character step (Systems.swift:603-1021, 1037-1051, 1102-1205, 1229-1375,
    1613-1901): the C++ oracle's move_and_slide against tests/independent_controller.py, a second transliteration of
    the controller logic written separately in scalar Python (Float = float32, Double = float64) that only borrows
    the oracle's three collision queries.  Every field of every character record after every step, bit for bit:
    walkers pressed into the mirror's hulls (side contacts, manifold cache, slide clipping, platform pushes) and
    human-scale walkers on ramps of 11-39 degrees (offset probes, normal smoothing, slope friction: stick / slide,
    flattenGround, ground transitions) next to a wall."""
    import independent_controller as ctl
    rng = np.random.default_rng(31)
    dt, gravity = 1.0 / 60.0, (0.0, -98.0, 0.0)
    if scene == "hulls":
        parts = scenes.mirror_scene(use_hulls=True)
        p = orc.default_params()
        pos, vel = scenes.gen_c3_characters(64, seed=5)
        plat = np.zeros(1, orc.PLATFORM)
        plat["aabb_min"], plat["aabb_max"], plat["delta"] = (-13, -4, 2), (-11, -3, 4), (0.03, 0.0, 0.02)
        steps = 80
    else:
        # ramps of 11, 22, 31 and 39 degrees (one flattenGround, one slippery) on a floor, plus a wall to run into
        fv, fi = scenes.plane_mesh(80.0)
        parts = [scenes.part(fv, fi, scenes.trs_model((0, -3, 0)), entity_id=0)]
        ramps = []
        for k, (h, mu_s, mu_k, flat) in enumerate([(2.0, 0.8, 0.6, False), (4.0, 0.3, 0.2, False), (6.0, 0.35, 0.25, True),
                                                    (8.0, 0.6, 0.5, False)]):
            rv, ri = scenes.wedge_mesh(8.0, 10.0, h)
            cx = -18.0 + 12.0 * k
            parts.append(scenes.part(rv, ri, scenes.trs_model((cx, -3.0 + h / 2, 0)), mu_s=mu_s, mu_k=mu_k,
                                     flatten_ground=flat, entity_id=1 + k))
            ramps.append((cx, h))
        bv, bi = scenes.box_mesh(6.0)
        parts.append(scenes.part(bv, bi, scenes.trs_model((0, 0, 12)), entity_id=9))
        p = orc.default_params(radius=0.4, half_height=0.5, skin_width=0.08)
        pos, vel = [], []
        for i in range(64):
            cx, h = ramps[i % 4]
            z = rng.uniform(-4.0, 4.5)
            y_slope = -3.0 + h * (5.0 - z) / 10.0  # ramp surface: low edge at z = +5, high edge at z = -5
            pos.append([cx + rng.uniform(-3, 3), y_slope + 0.9 + rng.uniform(0.05, 0.4), z])
            ang = rng.uniform(0, 2 * np.pi)
            vel.append([np.cos(ang) * rng.uniform(0, 5), 0.0, np.sin(ang) * rng.uniform(0, 5)])
        pos, vel = np.float32(pos), np.float32(vel)
        plat = np.zeros(0, orc.PLATFORM)
        steps = 90
    w = orc.OracleWorld(parts)
    q = ctl.Queries(orc, w, parts, orc.ORDER_CANONICAL)
    so, sp = orc.init_states(pos, vel), orc.init_states(pos, vel)
    walk = vel.copy()
    seen = {"side": 0, "slide": 0, "transition": 0, "manifold": 0}
    for step in range(steps):
        w.move_and_slide(so, p, dt, gravity, 1, orc.ORDER_CANONICAL, platforms=plat)
        for i in range(len(sp)):
            ctl.fixed_step(sp[i], p, q, dt, gravity, True, plat)
        for i in range(len(sp)):
            bad = _state_equal(so[i], sp[i])
            assert bad is None, (scene, step, i, bad, so[i][bad], sp[i][bad])
        seen["side"] += int((so["side_contact_frames"] > 0).sum())
        seen["slide"] += int(so["ground_sliding"].sum())
        seen["transition"] += int((so["ground_transition_frames"] > 0).sum())
        seen["manifold"] += int((so["manifold_count"] > 1).sum())
        if step % 10 == 9:  # walk intent re-applied: keeps the walkers pressing into walls / climbing slopes
            for s in (so, sp):
                s["velocity"][:, 0], s["velocity"][:, 2] = walk[:, 0], walk[:, 2]
    assert so["grounded"].mean() > 0.6
    if scene == "hulls":
        assert seen["side"] > 50 and seen["manifold"] > 0
    else:
        assert seen["slide"] > 20 and seen["transition"] > 0


def test_oracle_crowd_step_matches_independent_transliteration(orc, scenes):
    """The crowd step — move-and-slide with agent CCD (AgentSweepSolver + HitSelector + the .agentHit case of
    SlideResolver, Systems.swift:1053-1091, 1378-1399) followed by AgentSeparationSystem (:1906-2210: grid, sequential
    pair resolver with its blocking-cast vetoes, post slide + snap, Float write-back) — the C++ oracle against the
    separately written Python version, every field of every record after each of the two phases of every step."""
    import independent_controller as ctl
    rng = np.random.default_rng(77)
    dt, gravity = 1.0 / 60.0, (0.0, -98.0, 0.0)
    fv, fi = scenes.plane_mesh(80.0)
    bv, bi = scenes.box_mesh(6.0)
    rv, ri = scenes.wedge_mesh(8.0, 10.0, 3.0)
    parts = [scenes.part(fv, fi, scenes.trs_model((0, -3, 0)), entity_id=0),
             scenes.part(bv, bi, scenes.trs_model((5.5, 0, 0)), entity_id=1),
             scenes.part(rv, ri, scenes.trs_model((-6, -1.5, 0)), mu_s=0.4, mu_k=0.3, entity_id=2)]
    w = orc.OracleWorld(parts)
    q = ctl.Queries(orc, w, parts, orc.ORDER_CANONICAL)
    p = orc.default_params(radius=0.4, half_height=0.5, skin_width=0.08)
    n = 70
    pos = np.stack([rng.uniform(-3, 2.4, n), np.full(n, -3.0 + 0.95) + rng.uniform(0, 0.2, n), rng.uniform(-3, 3, n)],
                   axis=1).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n)
    walk = (np.stack([np.cos(ang), np.zeros(n), np.sin(ang)], axis=1) * rng.uniform(1, 6, (n, 1))).astype(np.float32)
    walk[: n // 3] = np.float32([4.0, 0, 0]) * rng.uniform(0.5, 1.5, (n // 3, 1)).astype(np.float32)  # a third walks into the wall
    mass = rng.choice(np.float32([1.0, 3.0, 500.0, 0.0]), n, p=[0.55, 0.25, 0.15, 0.05])
    so, sp = orc.init_states(pos, walk), orc.init_states(pos, walk)
    pairs = 0
    for step in range(25):
        agents = ctl.collect_agent_states(sp, dt, gravity)
        w.move_and_slide(so, p, dt, gravity, 3, orc.ORDER_CANONICAL)
        for i in range(n):
            ctl.fixed_step(sp[i], p, q, dt, gravity, True, (), agents, i)
        for i in range(n):
            bad = _state_equal(so[i], sp[i])
            assert bad is None, ("move", step, i, bad, so[i][bad], sp[i][bad])
        pairs += w.agent_separation(so, p, mass_weight=mass, order=orc.ORDER_CANONICAL)
        ctl.agent_separation(sp, p, q, mass_weight=mass)
        for i in range(n):
            bad = _state_equal(so[i], sp[i])
            assert bad is None, ("separation", step, i, bad, so[i][bad], sp[i][bad])
        for s in (so, sp):
            s["velocity"][:, 0], s["velocity"][:, 2] = walk[:, 0], walk[:, 2]
    assert pairs > 10 * n


def test_oracle_query_layer_matches_independent_brute_force(orc, scenes):
    """capsuleCast / capsuleCastBlocking / capsuleCastGround / capsuleOverlapAll (CollisionQuery.swift:787-882, 980-1283:
    swept box, layer mask, triangle-AABB test, blocking and ground filters, strict-< best, chooseNearest over the static
    and dynamic sets): the oracle in canonical order against a brute-force Python version written separately, on the
    mirror's hulls + floor with a dynamic box and mixed layers — bit for bit, including which triangle wins."""
    import independent_narrow_phase as ind
    rng = np.random.default_rng(123)
    parts = scenes.mirror_scene(use_hulls=True)
    bv, bi = scenes.box_mesh(3.0)
    parts.append(scenes.part(bv, bi, scenes.trs_model((-9.0, -1.5, 6.5)), layer=4, is_dynamic=True, entity_id=77))
    w = orc.OracleWorld(parts)
    tris = []
    for which in (0, 1):
        soup = w.read_soup(which)
        P, I, L = soup["positions"], soup["indices"], soup["layers"]
        tris += [(tuple(P[a]), tuple(P[b]), tuple(P[c]), int(l)) for (a, b, c), l in zip(I, L)]
    n_static = w.counts(0)["triangles"]
    n = 1500
    q = scenes.gen_casts(n, [-15, -3.5, -1], [-5, 4, 10], seed=8, len_range=(0.05, 4.0), expand=1.0)
    q["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 1, 4, 0xFFFFFFFB]), n)
    q["min_normal_y"] = 0.5
    for mode, blocking, min_y in ((0, False, None), (1, True, None), (2, False, np.float32(0.5))):
        ref = w.capsule_cast(q, mode, orc.ORDER_CANONICAL)
        hits = 0
        for i in range(n):
            got = ind.capsule_cast(tris, n_static, tuple(q["from"][i]), tuple(q["delta"][i]), q["radius"][i],
                                   q["half_height"][i], q["mask"][i], blocking, min_y)
            if got is None:
                assert ref["triangle_index"][i] == -1, (mode, i)
                continue
            hits += 1
            assert ref["triangle_index"][i] == got[4], (mode, i)
            assert np.float32(got[0]).tobytes() == ref["toi"][i].tobytes(), (mode, i)
            for field, val in (("position", got[1]), ("normal", got[2]), ("triangle_normal", got[3])):
                assert np.float32(val).tobytes() == ref[field][i].tobytes(), (mode, i, field)
        assert hits > n // 10, mode
    caps = scenes.gen_capsules(n, [-15, -3.5, -1], [-5, 2, 10], seed=9)
    caps["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 1, 4]), n)
    out, counts, _ = w.capsule_overlap_all(caps, 8, orc.ORDER_CANONICAL)
    overlapping = 0
    for i in range(n):
        got = ind.capsule_overlap_all(tris, tuple(caps["from"][i]), caps["radius"][i], caps["half_height"][i], caps["mask"][i], 8)
        assert counts[i] == len(got), i
        overlapping += len(got) > 0
        for k, h in enumerate(got):
            r = out[i][k]
            assert r["triangle_index"] == h[4] and np.float32(h[0]).tobytes() == r["depth"].tobytes(), (i, k)
            for field, val in (("position", h[1]), ("normal", h[2]), ("triangle_normal", h[3])):
                assert np.float32(val).tobytes() == r[field].tobytes(), (i, k, field)
    assert overlapping > n // 20


def test_oracle_soup_and_raycast_match_independent_brute_force(orc, scenes):
    """TriangleMeshSet.rebuild (CollisionQuery.swift:331-417: world-space vertices, absolute area filter, numbering of the
    survivors) and raycast (:768-785, 916-978, 1575-1601): the oracle's soup and its canonical-order ray results against
    the separately written Python versions — positions and indices equal bit for bit, then every ray's winner and
    distance.  The scene mixes a scaled + rotated part, a part whose small triangles fall under the area filter, a
    dynamic part and per-ray masks."""
    import independent_narrow_phase as ind
    rng = np.random.default_rng(77)
    bv, bi = scenes.box_mesh(2.0)
    pv, pi = scenes.plane_mesh(30.0)
    q = scenes.quat_angle_axis(0.7, (0.3, 1.0, -0.2))
    tiny = np.concatenate([bv * 1e-3, bv + np.float32([4, 0, 0])]).astype(np.float32)  # first copy: area² << 1e-10
    tiny_idx = np.concatenate([bi, bi + len(bv)]).astype(np.uint32)
    parts = [scenes.part(pv, pi, scenes.trs_model((0, -2, 0)), layer=1, entity_id=1),
             scenes.part(bv, bi, scenes.trs_model((1.5, 0.2, -1.0), q, (1.5, 0.75, 2.0)), layer=2, entity_id=2),
             scenes.part(tiny, tiny_idx, scenes.trs_model((-3, 0, 2)), layer=4, entity_id=3),
             scenes.part(bv, bi, scenes.trs_model((0, 1.0, 4.0), q), layer=8, is_dynamic=True, entity_id=4)]
    w = orc.OracleWorld(parts)
    tris = []
    for which, dyn in ((0, False), (1, True)):
        soup = w.read_soup(which)
        mine_p, mine_t = ind.build_soup([p for p in parts if bool(p["is_dynamic"]) == dyn])
        assert np.float32(mine_p).tobytes() == soup["positions"].tobytes(), which
        assert [t[:3] for t in mine_t] == [tuple(int(x) for x in r) for r in soup["indices"]], which
        assert [t[3] for t in mine_t] == [int(x) for x in soup["layers"]], which
        tris += [(mine_p[a], mine_p[b], mine_p[c], l) for a, b, c, l in mine_t]
    n_static = w.counts(0)["triangles"]
    assert n_static == 2 + 12 + 12 and w.counts(1)["triangles"] == 12  # the 12 millimetre-box triangles are dropped
    n = 3000
    rays = scenes.gen_rays(n, [-6, -2, -6], [6, 3, 6], seed=5, max_distance=25.0, expand=2.0)
    rays["direction"] *= rng.uniform(0.25, 3.0, (n, 1)).astype(np.float32)  # callee does not normalise (:768)
    rays["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 1, 2, 8, 0xFFFFFFF7]), n)
    ref = w.raycast(rays, orc.ORDER_CANONICAL)
    hits = 0
    for i in range(n):
        got = ind.raycast(tris, n_static, tuple(rays["origin"][i]), tuple(rays["direction"][i]), rays["max_distance"][i],
                          rays["mask"][i])
        if got is None:
            assert ref["triangle_index"][i] == -1, i
            continue
        hits += 1
        assert ref["triangle_index"][i] == got[1], i
        assert np.float32(got[0]).tobytes() == ref["distance"][i].tobytes(), i
    assert hits > n // 4


def test_oracle_reference_order_matches_independent_bvh_and_walks(orc, scenes):
    """ORDER_REFERENCE — the tree- and order-dependent half of the reference's answers (which triangle wins an exact toi
    tie, which maxHits overlaps are kept, which grazing ray hits its slab test culls) — against tests/independent_bvh.py:
    the reference's median-split BVH (CollisionQuery.swift:577-670, preorder numbering, sort fallback), refit (:528-575),
    and the three stack walks (right child first) transliterated separately.  Same node count, then every query's winner
    and value bit for bit, before and after a refit of a dynamic and a static part (the refitted tree keeps its topology,
    so a rebuild would answer differently)."""
    import independent_bvh as ib
    rng = np.random.default_rng(2024)
    parts = scenes.c1_scene()
    bv, bi = scenes.box_mesh(2.5)
    q0 = scenes.quat_angle_axis(0.4, (0.2, 1.0, 0.1))
    parts.append(scenes.part(bv, bi, scenes.trs_model((3.0, -1.5, 5.0), q0), layer=4, is_dynamic=True, entity_id=900))
    parts.append(scenes.part(bv, bi, scenes.trs_model((-4.0, -2.0, -2.0)), layer=8, is_dynamic=True, entity_id=901))
    for k, p in enumerate(parts):
        p["entity_id"] = p.get("entity_id", k)
    w = orc.OracleWorld(parts)
    sets = [ib.TriangleSet([p for p in parts if not p["is_dynamic"]]), ib.TriangleSet([p for p in parts if p["is_dynamic"]])]
    lo, hi = scenes.scene_aabb([p for p in parts[1:]])

    def compare(tag, n_cast, n_cap, n_ray, seed):
        for which in (0, 1):
            c = w.counts(which)
            assert c["triangles"] == len(sets[which].triangles) and c["nodes"] == len(sets[which].bvh.nodes), (tag, which)
        q = scenes.gen_casts(n_cast, lo, hi, seed=seed, len_range=(0.05, 3.0), expand=1.0)
        q["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 0xFFFFFFFF, 1, 12, 0xFFFFFFF3]), n_cast)
        q["min_normal_y"] = 0.5
        ties = 0
        for mode, blocking, min_y in ((0, False, None), (1, True, None), (2, False, np.float32(0.5))):
            ref = w.capsule_cast(q, mode, orc.ORDER_REFERENCE)
            canon = w.capsule_cast(q, mode, orc.ORDER_CANONICAL)
            hits = 0
            for i in range(n_cast):
                got = ib.capsule_cast(sets[0], sets[1], tuple(q["from"][i]), tuple(q["delta"][i]), q["radius"][i],
                                      q["half_height"][i], q["mask"][i], blocking, min_y)
                if got is None:
                    assert ref["triangle_index"][i] == -1, (tag, mode, i)
                    continue
                hits += 1
                assert ref["triangle_index"][i] == got[4], (tag, mode, i)
                assert np.float32(got[0]).tobytes() == ref["toi"][i].tobytes(), (tag, mode, i)
                for field, val in (("position", got[1]), ("normal", got[2]), ("triangle_normal", got[3])):
                    assert np.float32(val).tobytes() == ref[field][i].tobytes(), (tag, mode, i, field)
            assert hits > n_cast // 10, (tag, mode)
            ties += int((ref["triangle_index"] != canon["triangle_index"]).sum())
        caps = scenes.gen_capsules(n_cap, lo, hi, seed=seed + 1, expand=0.5)
        caps["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 0xFFFFFFFF, 1, 12]), n_cap)
        for max_hits in (2, 8):
            out, counts, overflow = w.capsule_overlap_all(caps, max_hits, orc.ORDER_REFERENCE)
            assert max_hits != 2 or overflow.sum() > 5, tag  # more overlaps than kept: the visiting order picks the survivors
            some = 0
            for i in range(n_cap):
                got = ib.capsule_overlap_all(sets[0], sets[1], tuple(caps["from"][i]), caps["radius"][i], caps["half_height"][i],
                                             caps["mask"][i], max_hits)
                assert counts[i] == len(got), (tag, max_hits, i)
                some += len(got) > 1
                # WHICH hits are kept is decided by the visiting order; the oracle then emits them the way the reference's
                # caller sorts them (Systems.swift:759, depth descending, stable)
                got = sorted(got, key=lambda h: -float(h[0]))
                for k, h in enumerate(got):
                    r = out[i][k]
                    assert r["triangle_index"] == h[4] and np.float32(h[0]).tobytes() == r["depth"].tobytes(), (tag, max_hits, i, k)
                    assert np.float32(h[1]).tobytes() == r["position"].tobytes() and np.float32(h[2]).tobytes() == r["normal"].tobytes()
            assert some > n_cap // 20, (tag, max_hits)
        single = w.capsule_overlap(caps, orc.ORDER_REFERENCE)  # capsuleOverlap: the deepest, first visited of equals
        deep = 0
        for i in range(n_cap):
            got = ib.capsule_overlap(sets[0], sets[1], tuple(caps["from"][i]), caps["radius"][i], caps["half_height"][i],
                                     caps["mask"][i])
            if got is None:
                assert single["triangle_index"][i] == -1, (tag, i)
                continue
            deep += 1
            assert single["triangle_index"][i] == got[4] and np.float32(got[0]).tobytes() == single["depth"][i].tobytes(), (tag, i)
            for field, val in (("position", got[1]), ("normal", got[2]), ("triangle_normal", got[3])):
                assert np.float32(val).tobytes() == single[field][i].tobytes(), (tag, i, field)
        assert deep > n_cap // 10, tag
        rays = scenes.gen_rays(n_ray, lo, hi, seed=seed + 2, max_distance=60.0, expand=2.0)
        rays["direction"][: n_ray // 8, rng.integers(0, 3)] = 0  # axis-parallel rays: the FLT_MAX inverse of rayAABB
        rays["mask"] = rng.choice(np.uint32([0xFFFFFFFF, 0xFFFFFFFF, 1, 12]), n_ray)
        ref = w.raycast(rays, orc.ORDER_REFERENCE)
        hits = 0
        for i in range(n_ray):
            got = ib.raycast(sets[0], sets[1], tuple(rays["origin"][i]), tuple(rays["direction"][i]), rays["max_distance"][i],
                             rays["mask"][i])
            if got is None:
                assert ref["triangle_index"][i] == -1, (tag, i)
                continue
            hits += 1
            assert ref["triangle_index"][i] == got[1] and np.float32(got[0]).tobytes() == ref["distance"][i].tobytes(), (tag, i)
        assert hits > n_ray // 4, tag
        return ties

    ties = compare("built", 500, 250, 1200, 11)
    # refit: the first dynamic box turns and moves, the wall box (static set) slides sideways
    dyn_parts = [p for p in parts if p["is_dynamic"]]
    sta_parts = [p for p in parts if not p["is_dynamic"]]
    m_dyn = scenes.trs_model((2.0, -1.0, 3.5), scenes.quat_angle_axis(1.1, (0.0, 1.0, 0.3)))
    base_t, base_q, base_s = scenes.transform_from_matrix(sta_parts[1]["model"])
    m_sta = scenes.trs_model((base_t[0] + 2.5, base_t[1], base_t[2] + 1.0), base_q, base_s)
    w.update_transforms([dyn_parts[0]["entity_id"], sta_parts[1]["entity_id"]], [m_dyn, m_sta])
    sets[1].update_transform(0, m_dyn)
    sets[0].update_transform(1, m_sta)
    ties += compare("refitted", 400, 200, 800, 21)
    assert ties > 20  # the scene does produce exact-tie cases, so the visiting order was really exercised


def test_per_triangle_materials_follow_the_reference_rules(orc, scenes):
    """StaticMeshComponent.triangleMaterials (CollisionQuery.swift:363-396): used when there is exactly one entry per
    triangle of the part (counted BEFORE the degenerate filter), ignored otherwise; a triangle the filter drops takes its
    entry with it; materialForTriangle indexes the filtered soup, static set first, default material out of range."""
    v = np.array([[0, 0, 0], [4, 0, 0], [0, 0, 4], [4, 0, 4], [8, 0, 0]], np.float32)
    idx = np.array([0, 2, 1, 0, 1, 4, 1, 2, 3], np.uint32)  # the middle triangle (0,1,4) is collinear: the filter drops it
    rows = np.array([[0.9, 0.7, 0], [0.5, 0.5, 1], [0.1, 0.05, 1]], np.float32)
    w = orc.OracleWorld([dict(scenes.part(v, idx, entity_id=7), triangle_materials=rows),
                         dict(scenes.part(v + np.float32([0, 5, 0]), idx[[0, 1, 2, 6, 7, 8]], entity_id=8, mu_s=0.3, mu_k=0.2,
                                          is_dynamic=True), triangle_materials=rows)])  # 3 entries for 2 triangles: ignored
    assert w.counts(0)["triangles"] == 2 and w.counts(1)["triangles"] == 2
    m = [w.triangle_material(t) for t in range(5)]
    assert m[0]["mu_s"] == pytest.approx(0.9) and not m[0]["flatten_ground"]
    assert m[1]["mu_s"] == pytest.approx(0.1) and m[1]["mu_k"] == pytest.approx(0.05) and m[1]["flatten_ground"]
    assert m[2]["mu_s"] == pytest.approx(0.3) and m[3]["mu_k"] == pytest.approx(0.2)  # the part's own material
    assert m[4]["mu_s"] == pytest.approx(0.8) and m[4]["mu_k"] == pytest.approx(0.6)  # SurfaceMaterial.default


def test_per_triangle_friction_decides_who_slides(orc, scenes):
    """SlopeFriction.apply (Systems.swift:965-1021) reads the material of the ground triangle: on one 27-degree slope a
    character standing on a sticky triangle stays, one standing on an icy triangle slides downhill."""
    s, L = np.float32(0.5), 40.0  # y = 0.5 * x
    v = np.array([[-L, -L * s, -L], [L, L * s, -L], [-L, -L * s, 0], [L, L * s, 0], [-L, -L * s, L], [L, L * s, L]], np.float32)
    idx = np.array([0, 2, 1, 1, 2, 3, 2, 4, 3, 3, 4, 5], np.uint32)  # z < 0: triangles 0,1; z > 0: triangles 2,3
    rows = np.array([[5.0, 5.0, 0], [5.0, 5.0, 0], [0.0, 0.0, 0], [0.0, 0.0, 0]], np.float32)
    w = orc.OracleWorld([dict(scenes.part(v, idx, entity_id=0), triangle_materials=rows)])
    p = orc.default_params()
    st = orc.init_states([[0, 3.2, -10], [0, 3.2, 10]])
    for _ in range(40):
        w.move_and_slide(st, p, order=orc.ORDER_REFERENCE)
    assert st["grounded"].all() and st["ground_triangle_index"].tolist() == [1, 2]
    sticky, icy = st["position"][0], st["position"][1]
    assert abs(sticky[0]) < 0.05 and icy[0] < -5.0, (sticky, icy)
