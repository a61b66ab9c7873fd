"""A SECOND, independent restatement of the reference's character-controller logic around the collision queries
(Game/Systems.swift:603-1021, 1037-1051, 1102-1205, 1229-1375, 1613-1901), written straight from the Swift in scalar
Python: Float = numpy.float32 (every operation rounded to single), Double = Python float.

TEST INFRASTRUCTURE ONLY.  The collision QUERIES themselves (capsuleCastBlocking / capsuleCastGround /
capsuleOverlapAll) are taken from the oracle through its public binding — their narrow phase is cross-checked
separately (independent_narrow_phase.py) — so what this file re-derives is everything the controller does between
two queries: gravity rule, contact-cache decay, platform carry, velocity gate, depenetration, the slide loop with
SlideResolver.resolveHit, ground probe / snap, slope friction and the write-back.  A step of the C++ oracle must
reproduce it bit for bit (tests/test_oracle_known_answers.py).
"""
import numpy as np

import independent_narrow_phase as ind

F = np.float32
ZERO, ONE = F(0), F(1)
smax, smin, add, sub, mul, neg, dot, cross, length_squared, normalize = (
    ind.smax, ind.smin, ind.add, ind.sub, ind.mul, ind.neg, ind.dot, ind.cross, ind.length_squared, ind.normalize)
MAX_MANIFOLD, MANIFOLD_FRAMES = 4, 8
FLT_MAX = np.finfo(np.float32).max


def length(a):
    return np.sqrt(dot(a, a))


def div(a, s):
    return (a[0] / s, a[1] / s, a[2] / s)


def d3(a):  # Float3 -> Double3
    return (float(a[0]), float(a[1]), float(a[2]))


def f3(a):  # Double3 -> Float3 (per component rounding)
    return (F(a[0]), F(a[1]), F(a[2]))


def ddot(a, b):  # simd_dot on Double3
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def dsub_scaled(v, n, s):  # v - n * s on Double3
    return (v[0] - n[0] * s, v[1] - n[1] * s, v[2] - n[2] * s)


class Controller:
    """CharacterControllerComponent + PhysicsBodyComponent of one character (Components.swift:353-431, 549-598)."""

    def __init__(self, rec, params):
        self.position = tuple(float(x) for x in rec["position"])
        self.velocity = tuple(float(x) for x in rec["velocity"])
        self.ground_normal = tuple(F(x) for x in rec["ground_normal"])
        self.ground_distance = F(rec["ground_distance"])
        self.side_contact_normal = tuple(F(x) for x in rec["side_contact_normal"])
        self.ground_triangle_index = int(rec["ground_triangle_index"])
        self.ground_transition_frames = int(rec["ground_transition_frames"])
        self.side_contact_frames = int(rec["side_contact_frames"])
        self.manifold_frames = int(rec["manifold_frames"])
        cnt = int(rec["manifold_count"])
        self.manifold_triangles = [int(t) for t in rec["manifold_triangles"][:cnt]]
        self.manifold_normals = [tuple(F(x) for x in n) for n in rec["manifold_normals"][:cnt]]
        self.grounded, self.grounded_near, self.ground_sliding = bool(rec["grounded"]), bool(rec["grounded_near"]), bool(rec["ground_sliding"])
        p = params.reshape(-1)[0]
        for name in ("radius", "half_height", "skin_width", "ground_snap_skin", "snap_distance", "fall_probe_distance",
                     "ground_snap_max_speed", "ground_snap_max_toi", "ground_snap_max_step", "ground_sweep_max_step",
                     "min_ground_dot"):
            setattr(self, name, F(p[name]))
        self.max_slide_iterations = int(p["max_slide_iterations"])
        self.collision_mask = int(p["collision_mask"])

    def store(self, rec):
        rec["position"], rec["velocity"] = self.position, self.velocity
        rec["ground_normal"], rec["ground_distance"] = self.ground_normal, self.ground_distance
        rec["side_contact_normal"] = self.side_contact_normal
        rec["ground_triangle_index"], rec["ground_transition_frames"] = self.ground_triangle_index, self.ground_transition_frames
        rec["side_contact_frames"], rec["manifold_frames"] = self.side_contact_frames, self.manifold_frames
        rec["manifold_count"] = len(self.manifold_triangles)
        for i, (t, n) in enumerate(zip(self.manifold_triangles, self.manifold_normals)):
            rec["manifold_triangles"][i] = t
            rec["manifold_normals"][i] = n
        rec["grounded"], rec["grounded_near"], rec["ground_sliding"] = self.grounded, self.grounded_near, self.ground_sliding


class Queries:
    """The three CollisionQuery calls the controller makes, through the oracle's public binding."""

    def __init__(self, oracle_module, world, parts, order):
        self.orc, self.w, self.order = oracle_module, world, order
        tri_part = np.concatenate([world.read_soup(s)["parts"] for s in (0, 1)])
        self.material = [(F(parts[p]["mu_s"]), F(parts[p]["mu_k"]), bool(parts[p]["flatten_ground"])) for p in tri_part]

    def _cast(self, c, frm, delta, mode, min_normal_y=0.0):
        q = np.zeros(1, self.orc.CAST)
        q["from"], q["delta"], q["radius"], q["half_height"] = frm, delta, c.radius, c.half_height
        q["mask"], q["min_normal_y"] = c.collision_mask, min_normal_y
        h = self.w.capsule_cast(q, mode, self.order)[0]
        if h["triangle_index"] < 0:
            return None
        return {"toi": F(h["toi"]), "position": tuple(h["position"]), "normal": tuple(h["normal"]),
                "triangle_normal": tuple(h["triangle_normal"]), "triangle_index": int(h["triangle_index"]),
                "material": self.material[int(h["triangle_index"])]}

    def blocking(self, c, frm, delta):
        return self._cast(c, frm, delta, 1)

    def ground(self, c, frm, delta):
        return self._cast(c, frm, delta, 2, c.min_ground_dot)

    def overlap_all(self, c, frm):
        q = np.zeros(1, self.orc.CAPSULE)
        q["from"], q["radius"], q["half_height"], q["mask"] = frm, c.radius, c.half_height, c.collision_mask
        out, counts, _ = self.w.capsule_overlap_all(q, 8, self.order)
        return [{"depth": F(h["depth"]), "normal": tuple(h["normal"]), "triangle_index": int(h["triangle_index"])}
                for h in out[0][: counts[0]]]


# ---------------------------------------------------------------- contact cache (:1102-1205)
def cache_decay(c):  # :1105-1116
    if c.side_contact_frames > 0:
        c.side_contact_frames -= 1
    if c.manifold_frames > 0:
        c.manifold_frames -= 1
        if c.manifold_frames == 0:
            c.manifold_triangles, c.manifold_normals = [], []
            c.manifold_frames = 0
            c.side_contact_normal = (ZERO, ZERO, ZERO)


def cached_normal(c, tri):  # :1169-1175
    for i, idx in enumerate(c.manifold_triangles):
        if idx == tri:
            return c.manifold_normals[i]
    return None


def manifold_update(c, tri, normal):  # :1177-1205
    n = normal
    if length_squared(n) < F(1e-8):
        return
    c.manifold_frames = MANIFOLD_FRAMES
    if tri in c.manifold_triangles:
        i = c.manifold_triangles.index(tri)
        cached = c.manifold_normals[i]
        if dot(cached, n) < ZERO:
            n = neg(n)
        blend = F(0.25)
        combined = normalize(add(mul(cached, ONE - blend), mul(n, blend)))
        c.manifold_normals[i] = combined
        c.side_contact_normal = combined
        return
    if len(c.manifold_triangles) >= MAX_MANIFOLD:
        c.manifold_triangles.pop()
        c.manifold_normals.pop()
    c.manifold_triangles.insert(0, tri)
    c.manifold_normals.insert(0, normalize(n))
    c.side_contact_normal = c.manifold_normals[0]


def cache_record(c, tri, normal, is_side):  # DefaultContactCachePolicy.record :1122-1133
    manifold_update(c, tri, normal)
    if is_side:
        c.side_contact_normal = normalize(normal)
        c.side_contact_frames = 3


# ---------------------------------------------------------------- platform carry (:644-732)
def platform_delta(position, c, platforms):
    if len(platforms) == 0:
        return (ZERO, ZERO, ZERO)
    capsule_half = c.half_height + c.radius
    base_y = position[1] - capsule_half
    cap_min = (position[0] - c.radius, position[1] - capsule_half, position[2] - c.radius)
    cap_max = (position[0] + c.radius, position[1] + capsule_half, position[2] + c.radius)
    side_tol = smax(c.skin_width, c.ground_snap_skin)
    best_carry, push = (ZERO, ZERO, ZERO), (ZERO, ZERO, ZERO)
    for pl in platforms:
        p_delta = tuple(F(x) for x in pl["delta"])
        if length_squared(p_delta) < F(1e-8):
            continue
        amin, amax = tuple(F(x) for x in pl["aabb_min"]), tuple(F(x) for x in pl["aabb_max"])
        emin, emax = tuple(x - side_tol for x in amin), tuple(x + side_tol for x in amax)
        if not all(cap_min[k] <= emax[k] and cap_max[k] >= emin[k] for k in range(3)):
            continue
        within_xz = (position[0] >= amin[0] - c.radius and position[0] <= amax[0] + c.radius and
                     position[2] >= amin[2] - c.radius and position[2] <= amax[2] + c.radius)
        top_y = amax[1]
        top_tol = c.snap_distance + smax(c.skin_width, c.ground_snap_skin) + F(0.05)
        if within_xz and base_y >= top_y - top_tol and base_y <= top_y + top_tol:
            if length_squared(p_delta) > length_squared(best_carry):
                best_carry = p_delta
            continue
        y_min, y_max = amin[1] - capsule_half, amax[1] + capsule_half
        if position[1] >= y_min and position[1] <= y_max:
            outside_x = position[0] < amin[0] - c.radius or position[0] > amax[0] + c.radius
            outside_z = position[2] < amin[2] - c.radius or position[2] > amax[2] + c.radius
            if not outside_x and not outside_z:
                continue
            cx = smax(amin[0], smin(position[0], amax[0]))
            cz = smax(amin[2], smin(position[2], amax[2]))
            dx, dz = position[0] - cx, position[2] - cz
            side_dist_sq = dx * dx + dz * dz
            side_push_tol = c.radius + side_tol
            if side_dist_sq <= side_push_tol * side_push_tol:
                dir_len = np.sqrt(smax(side_dist_sq, ZERO))
                if dir_len > F(1e-5):
                    direction = (dx / dir_len, ZERO, dz / dir_len)
                    flat = (p_delta[0], ZERO, p_delta[2])
                    if dot(flat, direction) > ZERO:
                        push = add(push, flat)
    if length_squared(best_carry) > F(1e-8):
        return best_carry
    if length_squared(push) > F(1e-8):
        return push
    return (ZERO, ZERO, ZERO)


# ---------------------------------------------------------------- depenetration (:734-808)
def depenetrate(c, position, q):
    slop = smax(c.skin_width * F(0.5), F(0.001))
    did, normal_sum, weight = False, (ZERO, ZERO, ZERO), ZERO
    for _ in range(4):
        hits = q.overlap_all(c, position)
        if not hits:
            break
        hits = sorted(hits, key=lambda h: -float(h["depth"]))
        deepest = hits[0]
        side = deepest["normal"][1] < c.min_ground_dot
        use = 1 if side else min(2, len(hits))
        max_depth = deepest["depth"]
        frame = (ZERO, ZERO, ZERO)
        for h in hits[:use]:
            max_depth = smax(max_depth, h["depth"])
            n = h["normal"]
            cached = cached_normal(c, h["triangle_index"])
            if cached is not None:
                n = cached
            frame = add(frame, mul(n, h["depth"]))
            cache_record(c, h["triangle_index"], n, h["normal"][1] < c.min_ground_dot)
        flen = length(frame)
        dn = div(frame, flen) if flen > F(1e-6) else frame
        push = smax(max_depth, ZERO) if side else smax(max_depth + slop, ZERO)
        if side:
            push = smin(push, c.skin_width)
        if push <= F(1e-6):
            break
        position = add(position, mul(dn, push))
        dnd = d3(dn)
        v_into = ddot(c.velocity, dnd)
        if v_into < 0:
            c.velocity = dsub_scaled(c.velocity, dnd, v_into)
        did = True
        normal_sum = add(normal_sum, mul(dn, max_depth))
        weight = weight + max_depth
    if not did:
        return position, None
    if weight > F(1e-6):
        return position, normalize(div(normal_sum, weight))
    return position, normalize(normal_sum)


# ---------------------------------------------------------------- SlideResolver.resolveHit (:1207-1375)
KINEMATIC_MOVE = dict(horizontal_ground_pass=False, adjust_velocity=True, ground_snap_skin_for_static=True, tri_normal_ground_like=True)
AGENT_SEPARATION = dict(horizontal_ground_pass=True, adjust_velocity=False, ground_snap_skin_for_static=False, tri_normal_ground_like=False)


def resolve_hit(c, position, remaining, seg_len, hit, was_grounded, was_grounded_near, cached_side, opt=KINEMATIC_MOVE):
    """hit: a static hit (dict with triangle_normal) or an agent hit (dict with "agent": True).  Returns (position,
    remaining, shouldBreak)."""
    is_static = not hit.get("agent", False)
    if opt["horizontal_ground_pass"] and is_static and abs(remaining[1]) < F(1e-5) and hit["normal"][1] >= c.min_ground_dot:
        return add(position, remaining), (ZERO, ZERO, ZERO), True
    hit_toi, slide_n = hit["toi"], hit["normal"]
    ground_like, tri_n = False, (ZERO, ZERO, ZERO)
    if is_static:
        ground_like = hit["triangle_normal"][1] >= c.min_ground_dot
        skin = c.ground_snap_skin if (opt["ground_snap_skin_for_static"] and ground_like) else c.skin_width
        tri_n = hit["triangle_normal"]
    else:
        skin = ZERO
    if is_static and slide_n[1] < c.min_ground_dot and c.side_contact_frames > 0:
        if cached_side is not None:
            cn = cached_side
            if dot(cn, slide_n) < ZERO:
                cn = neg(cn)
            slide_n = cn
        else:
            cached = c.side_contact_normal
            cl = length_squared(cached)
            if cl > F(1e-6):
                cn = div(cached, np.sqrt(cl))
                dc = dot(cn, slide_n)
                if abs(dc) > F(0.5):
                    slide_n = cn if dc >= ZERO else neg(cn)
    if slide_n[1] < c.min_ground_dot:
        if is_static and ground_like and opt["tri_normal_ground_like"]:
            slide_n = tri_n
        if slide_n[1] < c.min_ground_dot:
            slide_n = (slide_n[0], ZERO, slide_n[2])
            nl = length(slide_n)
            if nl > F(1e-5):
                slide_n = div(slide_n, nl)
            else:
                return add(position, remaining), (ZERO, ZERO, ZERO), True
    into = dot(remaining, slide_n)
    into_eps = F(1e-4) * seg_len
    eff = smin(skin, hit_toi * F(0.5)) if (hit_toi <= skin and into < -into_eps) else skin
    sticky = skin * F(0.1)
    if hit_toi <= sticky and into < -into_eps:
        return position, sub(remaining, mul(slide_n, into)), False
    if into >= -into_eps:
        if was_grounded_near and is_static and not ground_like and remaining[1] < ZERO:
            remaining = (remaining[0], ZERO, remaining[2])
        return add(position, remaining), (ZERO, ZERO, ZERO), True
    if (hit_toi <= eff and abs(into) <= into_eps) or into >= ZERO:
        return add(position, remaining), (ZERO, ZERO, ZERO), True
    move = smax(hit_toi - eff, ZERO)
    if slide_n[1] >= c.min_ground_dot and remaining[1] < ZERO and move > c.ground_sweep_max_step:
        move = c.ground_sweep_max_step
    direction = div(remaining, seg_len)
    position = add(position, mul(direction, move))
    left = sub(remaining, mul(direction, move))
    left = sub(left, mul(slide_n, dot(left, slide_n)))
    if was_grounded and was_grounded_near and left[1] < ZERO:
        left = (left[0], ZERO, left[2])
    residual = dot(left, slide_n)
    if abs(residual) < F(1e-5):
        left = sub(left, mul(slide_n, residual))
    if length_squared(left) < F(1e-8):
        return position, (ZERO, ZERO, ZERO), True
    if opt["adjust_velocity"]:
        snd = d3(slide_n)
        v_into = ddot(c.velocity, snd)
        if v_into < 0:
            c.velocity = dsub_scaled(c.velocity, snd, v_into)
    return position, left, False


# ---------------------------------------------------------------- agents (:1053-1091, 1378-1399) + slide loop (:1658-1765)
def agent_best_hit(position, remaining, remaining_len, base_move_len, dt, self_index, c, agents):
    """agents: list of (position, velocity) Float3 snapshots, one per character (all solid, controller radius)."""
    best = None
    time_scale = smin(remaining_len / base_move_len, ONE) if base_move_len > F(1e-6) else ONE
    segment_dt = dt * time_scale
    for k, (o_pos, o_vel) in enumerate(agents):
        if k == self_index:
            continue
        hit = ind.capsule_capsule_sweep(position, remaining, c.radius, c.half_height, o_pos, mul(o_vel, segment_dt),
                                        c.radius, c.half_height)
        if hit is not None and (best is None or hit[0] < best["toi"]):
            best = {"toi": hit[0], "normal": hit[1], "agent": True}
    return best


def select_best_hit(c, static_hit, agent_hit):  # HitSelector.selectBestHit
    if static_hit is not None and agent_hit is not None:
        skin = c.ground_snap_skin if static_hit["normal"][1] >= c.min_ground_dot else c.skin_width
        static_stop, agent_stop = smax(static_hit["toi"] - skin, ZERO), smax(agent_hit["toi"], ZERO)
        return static_hit if static_stop <= agent_stop else agent_hit
    return static_hit if static_hit is not None else agent_hit


def kinematic_sweep(c, position, remaining, was_grounded, was_grounded_near, q, dt=None, agents=None, self_index=-1):
    base_move_len = length(mul(f3(c.velocity), dt)) if agents is not None else ZERO
    last = None
    for _ in range(c.max_slide_iterations):
        seg_len = length(remaining)
        if seg_len < F(1e-6):
            break
        static_hit = q.blocking(c, position, remaining)
        if static_hit is not None and static_hit["normal"][1] < c.min_ground_dot and c.side_contact_frames > 0:
            cached = cached_normal(c, static_hit["triangle_index"])
            if cached is not None:
                if dot(cached, static_hit["normal"]) < ZERO:
                    cached = neg(cached)
                static_hit["normal"] = cached
        agent_hit = None
        if agents is not None:
            agent_hit = agent_best_hit(position, remaining, seg_len, base_move_len, dt, self_index, c, agents)
        hit = select_best_hit(c, static_hit, agent_hit)
        if hit is None:
            position = add(position, remaining)
            remaining = (ZERO, ZERO, ZERO)
            break
        hit_normal = hit["normal"]
        is_static = not hit.get("agent", False)
        cached_side = None
        if is_static and hit_normal[1] < c.min_ground_dot and c.side_contact_frames > 0:
            cached_side = cached_normal(c, hit["triangle_index"])
        position, remaining, should_break = resolve_hit(c, position, remaining, seg_len, hit, was_grounded,
                                                        was_grounded_near, cached_side)
        if is_static and hit_normal[1] < c.min_ground_dot:
            cache_record(c, hit["triangle_index"], hit_normal, True)
        if last is not None:
            dn = dot(last, hit_normal)
            if abs(dn) < F(0.98):
                axis = cross(last, hit_normal)
                al = length(axis)
                if al > F(1e-5):
                    an = div(axis, al)
                    remaining = mul(an, dot(remaining, an))
        last = hit_normal
        if should_break:
            break
    return position, remaining


# ---------------------------------------------------------------- ground probe / snap / slope friction (:826-1021, 1766-1800)
def ground_contact(c, position, q, was_grounded_near, gravity, dt):
    state = {"grounded": False, "grounded_near": False, "normal": (ZERO, ONE, ZERO), "material": (F(0.8), F(0.6), False),
             "triangle_index": -1, "distance": FLT_MAX}
    down = (ZERO, F(-1), ZERO)
    snap_delta = mul(down, c.snap_distance)
    center = q.ground(c, position, snap_delta) if c.snap_distance > ZERO else None
    if c.fall_probe_distance > ZERO:
        fall = q.ground(c, position, mul(down, c.fall_probe_distance))
        if fall is not None:
            state["distance"] = fall["toi"]
    can_snap, near_ground, probe_hit = False, False, None
    if center is not None and center["toi"] <= c.snap_distance:
        probe_hit = center
        bottom_y = (position[1] - c.half_height) - c.radius
        valid = center["position"][1] <= bottom_y + smax(c.skin_width, c.ground_snap_skin)
        near_ground = center["toi"] <= smax(c.ground_snap_skin, c.skin_width)
        state["grounded_near"] = near_ground
        state["distance"] = center["toi"]
        gate_vel = c.velocity[1] <= 0
        gate_speed = ddot(c.velocity, d3(center["normal"])) >= -float(c.ground_snap_max_speed)
        gate_toi = center["toi"] <= c.ground_snap_max_toi
        can_snap = valid and gate_vel and (near_ground or gate_speed or gate_toi)
        if was_grounded_near and center["toi"] <= c.snap_distance:
            can_snap = valid
        if valid and (near_ground or can_snap):
            state["grounded"], state["material"], state["triangle_index"] = True, center["material"], center["triangle_index"]
            nsum = center["triangle_normal"]
            if center["triangle_normal"][1] < F(0.98) and (was_grounded_near or near_ground):
                off = c.radius * F(0.6)
                tol = smax(smax(c.ground_snap_skin, c.skin_width), F(0.05))
                for ox, oz in ((off, ZERO), (-off, ZERO), (ZERO, off), (ZERO, -off)):
                    h = q.ground(c, add(position, (ox, ZERO, oz)), snap_delta)
                    if h is not None and h["toi"] <= center["toi"] + tol:
                        if dot(h["triangle_normal"], center["triangle_normal"]) > F(0.98):
                            nsum = add(nsum, h["triangle_normal"])
            nl = length(nsum)
            state["normal"] = div(nsum, nl) if nl > F(1e-6) else center["triangle_normal"]
        if state["grounded"] and was_grounded_near:
            if dot(c.ground_normal, state["normal"]) > F(0.9):
                blend = F(0.2)
                state["normal"] = normalize(add(mul(c.ground_normal, ONE - blend), mul(state["normal"], blend)))
        if state["grounded"] and state["material"][2]:
            state["normal"] = (ZERO, ONE, ZERO)
    # GroundSnap.apply (:945-963)
    if can_snap and probe_hit is not None:
        move = smax(probe_hit["toi"] - c.ground_snap_skin, ZERO)
        if near_ground and move > c.ground_snap_max_step:
            move = c.ground_snap_max_step
        position = add(position, mul(down, move))
        nd = d3(probe_hit["normal"])
        v_into = ddot(c.velocity, nd)
        if v_into < 0:
            c.velocity = dsub_scaled(c.velocity, nd, v_into)
    if state["grounded"]:  # :1787-1792
        if state["triangle_index"] != c.ground_triangle_index and state["normal"][1] - c.ground_normal[1] > F(0.02):
            c.ground_transition_frames = 3
    # SlopeFriction.apply (:965-1021)
    if not state["grounded"]:
        c.ground_sliding = False
        return position, state
    normal = normalize(state["normal"])
    if normal[1] > F(0.98):
        c.ground_transition_frames = 0
        c.ground_sliding = False
        return position, state
    if c.ground_transition_frames > 0:
        c.ground_transition_frames -= 1
        c.ground_sliding = False
        return position, state
    g_n = dot(gravity, normal)
    g_tan = sub(gravity, mul(normal, g_n))
    g_tan_len = length(g_tan)
    if g_tan_len > F(0.5):
        g_n_mag = abs(g_n)
        g_dir = div(g_tan, g_tan_len)
        g_dir_d, normal_d = d3(g_dir), d3(normal)
        stick = state["material"][0] * g_n_mag
        enter, leave = g_tan_len > stick * F(1.05), g_tan_len < stick * F(0.9)
        if c.ground_sliding:
            if leave:
                c.ground_sliding = False
        elif enter:
            c.ground_sliding = True
        if not c.ground_sliding and g_tan_len <= stick:
            v_tan = dsub_scaled(c.velocity, normal_d, ddot(c.velocity, normal_d))
            downhill = ddot(v_tan, g_dir_d)
            if downhill > 0:
                c.velocity = dsub_scaled(c.velocity, g_dir_d, downhill)
        else:
            accel = smax(g_tan_len - state["material"][1] * g_n_mag, ZERO)
            if accel > ZERO:
                c.velocity = (c.velocity[0] + g_dir_d[0] * float(accel) * float(dt),
                              c.velocity[1] + g_dir_d[1] * float(accel) * float(dt),
                              c.velocity[2] + g_dir_d[2] * float(accel) * float(dt))
    return position, state


# ---------------------------------------------------------------- one fixed step of one character (:603-619, 1823-1901)
def fixed_step(rec, params, q, dt, gravity, apply_gravity=True, platforms=(), agents=None, self_index=-1):
    c = Controller(rec, params)
    dt = F(dt)
    gravity = tuple(F(x) for x in gravity)
    if apply_gravity and not (c.grounded and c.grounded_near):  # GravitySystem
        gd = d3(gravity)
        c.velocity = (c.velocity[0] + gd[0] * float(dt), c.velocity[1] + gd[1] * float(dt), c.velocity[2] + gd[2] * float(dt))
    position = f3(c.position)
    cache_decay(c)
    delta = platform_delta(position, c, platforms)
    if length_squared(delta) > F(1e-8):
        position = add(position, delta)
    was_grounded, was_grounded_near = c.grounded, c.grounded_near
    # VelocityGate.apply (:1037-1051)
    if was_grounded and was_grounded_near and c.velocity[1] < 0:
        c.velocity = (c.velocity[0], 0.0, c.velocity[2])
    rem_d = (c.velocity[0] * float(dt), c.velocity[1] * float(dt), c.velocity[2] * float(dt))
    if was_grounded and was_grounded_near and rem_d[1] < 0:
        rem_d = (rem_d[0], 0.0, rem_d[2])
    remaining = f3(rem_d)
    position, depen_n = depenetrate(c, position, q)
    if depen_n is not None:
        into = dot(remaining, depen_n)
        if into < ZERO:
            remaining = sub(remaining, mul(depen_n, into))
    position, remaining = kinematic_sweep(c, position, remaining, was_grounded, was_grounded_near, q, dt, agents, self_index)
    position, state = ground_contact(c, position, q, was_grounded_near, gravity, dt)
    # writeBack (:1802-1821)
    c.position = d3(position)
    c.grounded, c.grounded_near = state["grounded"], state["grounded_near"]
    c.ground_normal = state["normal"] if state["grounded"] else (ZERO, ONE, ZERO)
    c.ground_distance = state["distance"]
    if state["grounded"]:
        c.ground_triangle_index = state["triangle_index"]
    c.store(rec)


def collect_agent_states(states, dt, gravity, apply_gravity=True):
    """collectAgentStates (:1592-1611) for a batch in which every character carries an AgentCollisionComponent: positionF
    and linearVelocityF of everybody, taken after GravitySystem ran and before anybody moves."""
    out = []
    g = d3(tuple(F(x) for x in gravity))
    for rec in states:
        vel = tuple(float(x) for x in rec["velocity"])
        if apply_gravity and not (bool(rec["grounded"]) and bool(rec["grounded_near"])):
            vel = (vel[0] + g[0] * float(F(dt)), vel[1] + g[1] * float(F(dt)), vel[2] + g[2] * float(F(dt)))
        out.append((f3(tuple(float(x) for x in rec["position"])), f3(vel)))
    return out


# ---------------------------------------------------------------- AgentSeparationSystem (:1906-2210)
def agent_separation(states, params, q, mass_weight=None, iterations=2, separation_margin=0.2, height_margin=0.1,
                     use_query=True):
    n = len(states)
    if n <= 1:
        return
    cs = [Controller(states[i], params) for i in range(n)]
    sep_margin, h_margin = F(separation_margin), F(height_margin)
    agents = []
    max_radius = ZERO
    for i, c in enumerate(cs):
        mw = F(1.0) if mass_weight is None else F(mass_weight[i])
        inv_w = ONE / mw if mw > ZERO else ZERO
        max_radius = smax(max_radius, c.radius)
        agents.append({"position": f3(c.position), "velocity": f3(c.velocity), "radius": c.radius, "hh": c.half_height,
                       "inv_w": inv_w})
    original = [a["position"] for a in agents]
    cell_size = smax(max_radius * F(2) + sep_margin, F(0.001))

    def cell_of(pos):
        return (int(np.floor(pos[0] / cell_size)), int(np.floor(pos[2] / cell_size)))

    for _ in range(max(1, iterations)):
        cells = {}
        for i, a in enumerate(agents):
            cells.setdefault(cell_of(a["position"]), []).append(i)
        for i in range(n):
            a = dict(agents[i])  # `let a = agents[i]`
            ccell = cell_of(a["position"])
            c = cs[i]
            for dz in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    lst = cells.get((ccell[0] + dx, ccell[1] + dz))
                    if lst is None:
                        continue
                    for j in lst:
                        if not j > i:
                            continue
                        b = dict(agents[j])
                        a_min, a_max = a["position"][1] - a["hh"], a["position"][1] + a["hh"]
                        b_min, b_max = b["position"][1] - b["hh"], b["position"][1] + b["hh"]
                        ddx, ddz = a["position"][0] - b["position"][0], a["position"][2] - b["position"][2]
                        dist_sq = ddx * ddx + ddz * ddz
                        margin = smin(sep_margin, smin(cs[i].skin_width, cs[j].skin_width))
                        min_dist = a["radius"] + b["radius"] + margin
                        if a_max < b_min - h_margin or a_min > b_max + h_margin:
                            continue
                        if dist_sq >= min_dist * min_dist:
                            continue
                        dist = np.sqrt(smax(dist_sq, F(1e-8)))
                        nx, nz = ddx / dist, ddz / dist
                        pen = min_dist - dist
                        w_sum = a["inv_w"] + b["inv_w"]
                        if w_sum <= ZERO:
                            continue
                        corr = pen / w_sum
                        move_a = (nx * corr * a["inv_w"], ZERO, nz * corr * a["inv_w"])
                        move_b = (-nx * corr * b["inv_w"], ZERO, -nz * corr * b["inv_w"])
                        rel_v = sub(a["velocity"], b["velocity"])
                        vn = rel_v[0] * nx + rel_v[2] * nz
                        if vn < ZERO:
                            impulse = -vn
                            scale_a, scale_b = a["inv_w"] / w_sum, b["inv_w"] / w_sum
                            va, vb = agents[i]["velocity"], agents[j]["velocity"]
                            agents[i]["velocity"] = (va[0] + nx * impulse * scale_a, va[1], va[2] + nz * impulse * scale_a)
                            agents[j]["velocity"] = (vb[0] - nx * impulse * scale_b, vb[1], vb[2] - nz * impulse * scale_b)
                        if use_query:
                            blocked_a = blocked_b = False
                            if length(move_a) > F(1e-6):
                                h = q.blocking(c, agents[i]["position"], move_a)
                                blocked_a = h is not None and h["toi"] <= c.skin_width and h["normal"][1] < c.min_ground_dot
                            if length(move_b) > F(1e-6):
                                h = q.blocking(cs[j], agents[j]["position"], move_b)
                                blocked_b = (h is not None and h["toi"] <= cs[j].skin_width and
                                             h["normal"][1] < cs[j].min_ground_dot)
                            if blocked_a and not blocked_b:
                                move_a, move_b = (ZERO, ZERO, ZERO), (-nx * pen, ZERO, -nz * pen)
                            elif blocked_b and not blocked_a:
                                move_b, move_a = (ZERO, ZERO, ZERO), (nx * pen, ZERO, nz * pen)
                            elif blocked_a and blocked_b:
                                continue
                        agents[i]["position"] = add(agents[i]["position"], move_a)
                        agents[j]["position"] = add(agents[j]["position"], move_b)
    down = (ZERO, F(-1), ZERO)
    for i in range(n):  # AgentSeparationPostProcessor.apply + write-back (:2043-2117, 2196-2208)
        c, agent = cs[i], agents[i]
        position = agent["position"]
        if use_query:
            delta = sub(position, original[i])
            moved = False
            if length(delta) > F(1e-6):
                moved = True
                remaining, position = delta, original[i]
                for _ in range(2):
                    seg_len = length(remaining)
                    if seg_len < F(1e-6):
                        break
                    hit = q.blocking(c, position, remaining)
                    if hit is None:
                        position = add(position, remaining)
                        remaining = (ZERO, ZERO, ZERO)
                        break
                    position, remaining, done = resolve_hit(c, position, remaining, seg_len, hit, False, False, None,
                                                            AGENT_SEPARATION)
                    if done:
                        break
            if moved and c.velocity[1] <= 0 and c.snap_distance > ZERO:
                hit = q.ground(c, position, mul(down, c.snap_distance))
                if hit is not None and hit["toi"] <= c.snap_distance:
                    move = smin(smax(hit["toi"] - c.ground_snap_skin, ZERO), c.ground_snap_max_step)
                    position = add(position, mul(down, move))
                    c.grounded = True
                    c.grounded_near = bool(hit["toi"] <= smax(c.ground_snap_skin, c.skin_width))
                    c.ground_normal = (ZERO, ONE, ZERO) if hit["material"][2] else hit["triangle_normal"]
                    c.ground_triangle_index = hit["triangle_index"]
        c.position = d3(position)
        c.velocity = d3(agent["velocity"])
        c.store(states[i])
