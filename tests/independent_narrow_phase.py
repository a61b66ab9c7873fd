"""A SECOND, independent restatement of the reference's narrow phase and conservative-advancement sweep
(Game/CollisionQuery.swift:1285-1573), written straight from the Swift in scalar Python with numpy.float32
arithmetic (every operation rounded to IEEE single, left-to-right like the Swift expressions).

TEST INFRASTRUCTURE ONLY.  The reference cannot be compiled here (no swiftc), so the C++ oracle cannot be pinned by
reference outputs; this file reduces the remaining risk — a transcription slip in the oracle — by checking the
oracle bit-for-bit against a separately written transliteration (tests/test_oracle_known_answers.py).  It shares
only the documented `simd` shim with the oracle: dot = (x*x' + y*y') + z*z', normalize = v * (1/sqrt),
Swift's generic min/max.
"""
import numpy as np

F = np.float32
ZERO, ONE, HALF = F(0), F(1), F(0.5)
FLT_MAX = np.finfo(np.float32).max


def smax(x, y):  # Swift max(x, y): y >= x ? y : x
    return y if y >= x else x


def smin(x, y):  # Swift min(x, y): y < x ? y : x
    return y if y < x else x


def v(x, y, z):
    return (F(x), F(y), F(z))


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def mul(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def neg(a):
    return (-a[0], -a[1], -a[2])


def dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def length_squared(a):
    return dot(a, a)


def normalize(a):
    with np.errstate(all="ignore"):
        return mul(a, ONE / np.sqrt(dot(a, a)))


def clamp(x, lo, hi):  # :1571-1573  min(max(v, minV), maxV)
    return smin(smax(x, lo), hi)


def closest_point_on_triangle(p, a, b, c):  # :1464-1517
    ab, ac, ap = sub(b, a), sub(c, a), sub(p, a)
    d1, d2 = dot(ab, ap), dot(ac, ap)
    if d1 <= ZERO and d2 <= ZERO:
        return length_squared(sub(p, a)), a
    bp = sub(p, b)
    d3, d4 = dot(ab, bp), dot(ac, bp)
    if d3 >= ZERO and d4 <= d3:
        return length_squared(sub(p, b)), b
    vc = d1 * d4 - d3 * d2
    if vc <= ZERO and d1 >= ZERO and d3 <= ZERO:
        vv = d1 / (d1 - d3)
        point = add(a, mul(ab, vv))
        return length_squared(sub(p, point)), point
    cp = sub(p, c)
    d5, d6 = dot(ab, cp), dot(ac, cp)
    if d6 >= ZERO and d5 <= d6:
        return length_squared(sub(p, c)), c
    vb = d5 * d2 - d1 * d6
    if vb <= ZERO and d2 >= ZERO and d6 <= ZERO:
        w = d2 / (d2 - d6)
        point = add(a, mul(ac, w))
        return length_squared(sub(p, point)), point
    va = d3 * d6 - d5 * d4
    if va <= ZERO and (d4 - d3) >= ZERO and (d5 - d6) >= ZERO:
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        point = add(b, mul(sub(c, b), w))
        return length_squared(sub(p, point)), point
    denom = ONE / (va + vb + vc)
    vv, w = vb * denom, vc * denom
    point = add(add(a, mul(ab, vv)), mul(ac, w))
    return length_squared(sub(p, point)), point


def segment_segment_distance_sq(p1, q1, p2, q2):  # :1519-1569
    d1, d2, r = sub(q1, p1), sub(q2, p2), sub(p1, p2)
    a, e, f = dot(d1, d1), dot(d2, d2), dot(d2, r)
    eps = F(1e-6)
    if a <= eps and e <= eps:
        return length_squared(sub(p1, p2)), p1, p2
    if a <= eps:
        t = clamp(f / e, ZERO, ONE)
        c2 = add(p2, mul(d2, t))
        return length_squared(sub(p1, c2)), p1, c2
    c = dot(d1, r)
    if e <= eps:
        s = clamp(-c / a, ZERO, ONE)
        c1 = add(p1, mul(d1, s))
        return length_squared(sub(c1, p2)), c1, p2
    b = dot(d1, d2)
    denom = a * e - b * b
    s = clamp((b * f - c * e) / denom, ZERO, ONE) if denom != ZERO else ZERO
    t_nom = b * s + f
    if t_nom < ZERO:
        t = ZERO
        s = clamp(-c / a, ZERO, ONE)
    elif t_nom > e:
        t = ONE
        s = clamp((b - c) / a, ZERO, ONE)
    else:
        t = t_nom / e
    c1, c2 = add(p1, mul(d1, s)), add(p2, mul(d2, t))
    return length_squared(sub(c1, c2)), c1, c2


def segment_triangle_intersect(a, b, v0, v1, v2):  # :1440-1462
    d = sub(b, a)
    e1, e2 = sub(v1, v0), sub(v2, v0)
    pvec = cross(d, e2)
    det = dot(e1, pvec)
    if abs(det) < F(1e-6):
        return None
    inv_det = ONE / det
    tvec = sub(a, v0)
    u = dot(tvec, pvec) * inv_det
    if u < ZERO or u > ONE:
        return None
    qvec = cross(tvec, e1)
    vv = dot(d, qvec) * inv_det
    if vv < ZERO or (u + vv) > ONE:
        return None
    t = dot(e2, qvec) * inv_det
    if t < ZERO or t > ONE:
        return None
    return add(a, mul(d, t))


def segment_triangle_distance(center, half_height, v0, v1, v2):  # :1396-1438
    up = v(0, 1, 0)
    a, b = add(center, mul(up, half_height)), sub(center, mul(up, half_height))
    hit = segment_triangle_intersect(a, b, v0, v1, v2)
    if hit is not None:
        return ZERO, hit, hit
    best, best_seg, best_tri = FLT_MAX, a, v0
    d0, p0 = closest_point_on_triangle(a, v0, v1, v2)
    if d0 < best:
        best, best_seg, best_tri = d0, a, p0
    d1, p1 = closest_point_on_triangle(b, v0, v1, v2)
    if d1 < best:
        best, best_seg, best_tri = d1, b, p1
    for e0, e1 in ((v0, v1), (v1, v2), (v2, v0)):
        d, s, t = segment_segment_distance_sq(a, b, e0, e1)
        if d < best:
            best, best_seg, best_tri = d, s, t
    return np.sqrt(smax(best, ZERO)), best_seg, best_tri


def refine_toi(frm, direction, radius, half_height, v0, v1, v2, t0, t1, max_distance):  # :1361-1394
    c0, c1 = smax(ZERO, smin(t0, max_distance)), smax(ZERO, smin(t1, max_distance))
    lo, hi = smin(c0, c1), smax(c0, c1)
    if hi - lo < F(1e-5):
        return hi
    for _ in range(10):
        mid = HALF * (lo + hi)
        dist, _, _ = segment_triangle_distance(add(frm, mul(direction, mid)), half_height, v0, v1, v2)
        if dist <= radius:
            hi = mid
        else:
            lo = mid
    return hi


def sweep_capsule_triangle(frm, direction, max_distance, radius, half_height, v0, v1, v2):  # :1285-1359
    """Returns None or (toi, position, normal, triangleNormal, iterations)."""
    min_advance = smax(radius * F(0.02), F(1e-4))
    max_iter = min(256, int(np.ceil(max_distance / min_advance)) + 1)
    contact_eps = F(1e-5)
    tri_normal = normalize(cross(sub(v1, v0), sub(v2, v0)))
    t, last_safe = ZERO, ZERO
    iterations = 0
    for _ in range(max_iter):
        iterations += 1
        if t > max_distance:
            return None
        dist, _, _ = segment_triangle_distance(add(frm, mul(direction, t)), half_height, v0, v1, v2)
        if dist <= radius + contact_eps:
            t_hit = refine_toi(frm, direction, radius, half_height, v0, v1, v2, last_safe, t, max_distance)
            hit_dist, hit_seg, hit_tri = segment_triangle_distance(add(frm, mul(direction, t_hit)), half_height, v0, v1, v2)
            if hit_dist < F(1e-6):
                n = neg(tri_normal) if dot(tri_normal, direction) > ZERO else tri_normal
            else:
                n = normalize(sub(hit_seg, hit_tri))
            tri_n = tri_normal
            if dot(tri_n, n) < ZERO:
                tri_n = neg(tri_n)
            return t_hit, hit_tri, n, tri_n, iterations
        last_safe = t
        advance = smax(dist - radius, min_advance)
        t = t + (min_advance if advance <= ZERO else advance)
    return None


# ---------------------------------------------------------------- Game/Systems.swift:1417-1590 (agent CCD)
def clamp_interval(start, end):  # :1417-1424
    s, e = smax(start, ZERO), smin(end, ONE)
    return None if e < s else (s, e)


def interval_greater_equal(y0, vy, threshold):  # :1426-1436
    if abs(vy) < F(1e-6):
        return (ZERO, ONE) if y0 >= threshold else None
    t = (threshold - y0) / vy
    return clamp_interval(t, ONE) if vy > ZERO else clamp_interval(ZERO, t)


def interval_less_equal(y0, vy, threshold):  # :1438-1448
    if abs(vy) < F(1e-6):
        return (ZERO, ONE) if y0 <= threshold else None
    t = (threshold - y0) / vy
    return clamp_interval(ZERO, t) if vy > ZERO else clamp_interval(t, ONE)


def earliest_root(a, b, c, t_min, t_max):  # :1450-1472
    eps = F(1e-6)
    if abs(a) < eps:
        if abs(b) < eps:
            return t_min if c <= ZERO else None
        t = -c / b
        return t if (t >= t_min and t <= t_max) else None
    disc = b * b - F(4) * a * c
    if disc < ZERO:
        return None
    sqrt_d = np.sqrt(disc)
    inv2a = ONE / (F(2) * a)
    t0, t1 = (-b - sqrt_d) * inv2a, (-b + sqrt_d) * inv2a
    s, e = smax(smin(t0, t1), t_min), smin(smax(t0, t1), t_max)
    return s if e >= s else None


def capsule_pair_separation_y(y_rel, h_sum):  # :1474-1482
    if y_rel > h_sum:
        return y_rel - h_sum
    if y_rel < -h_sum:
        return y_rel + h_sum
    return ZERO


def capsule_pair_hit_normal(rel, h_sum):  # :1484-1497
    sep = (rel[0], capsule_pair_separation_y(rel[1], h_sum), rel[2])
    len_sq = length_squared(sep)
    if len_sq > F(1e-8):
        d = np.sqrt(len_sq)
        return (sep[0] / d, sep[1] / d, sep[2] / d)
    lateral = (rel[0], ZERO, rel[2])
    lat_sq = length_squared(lateral)
    if lat_sq > F(1e-8):
        d = np.sqrt(lat_sq)
        return (lateral[0] / d, lateral[1] / d, lateral[2] / d)
    return v(1, 0, 0)


def capsule_pair_overlap(rel, r_sum, h_sum):  # :1499-1503
    sep_y = capsule_pair_separation_y(rel[1], h_sum)
    return rel[0] * rel[0] + rel[2] * rel[2] + sep_y * sep_y <= r_sum * r_sum


def capsule_capsule_sweep(frm, delta, radius, half_height, other_pos, other_delta, other_radius, other_half_height):
    """:1505-1590.  Returns None or (toi, normal)."""
    rel_start, rel_delta = sub(frm, other_pos), sub(delta, other_delta)
    r_sum, h_sum = radius + other_radius, half_height + other_half_height
    rel_len, move_len = np.sqrt(dot(rel_delta, rel_delta)), np.sqrt(dot(delta, delta))
    if rel_len < F(1e-6):
        if capsule_pair_overlap(rel_start, r_sum, h_sum):
            return ZERO, capsule_pair_hit_normal(rel_start, h_sum)
        return None
    y0, vy, vx, vz, r0x, r0z = rel_start[1], rel_delta[1], rel_delta[0], rel_delta[2], rel_start[0], rel_start[2]
    best = None
    upper = interval_greater_equal(y0, vy, h_sum)
    if upper is not None:
        a = vx * vx + vz * vz + vy * vy
        b = F(2) * (r0x * vx + r0z * vz + (y0 - h_sum) * vy)
        c = r0x * r0x + r0z * r0z + (y0 - h_sum) * (y0 - h_sum) - r_sum * r_sum
        t = earliest_root(a, b, c, upper[0], upper[1])
        if t is not None:
            best = t
    lower = interval_less_equal(y0, vy, -h_sum)
    if lower is not None:
        a = vx * vx + vz * vz + vy * vy
        b = F(2) * (r0x * vx + r0z * vz + (y0 + h_sum) * vy)
        c = r0x * r0x + r0z * r0z + (y0 + h_sum) * (y0 + h_sum) - r_sum * r_sum
        t = earliest_root(a, b, c, lower[0], lower[1])
        if t is not None and (best is None or t < best):
            best = t
    if abs(vy) < F(1e-6):
        if abs(y0) <= h_sum:
            a, b, c = vx * vx + vz * vz, F(2) * (r0x * vx + r0z * vz), r0x * r0x + r0z * r0z - r_sum * r_sum
            t = earliest_root(a, b, c, ZERO, ONE)
            if t is not None and (best is None or t < best):
                best = t
    else:
        t1, t2 = (h_sum - y0) / vy, (-h_sum - y0) / vy
        overlap = clamp_interval(smin(t1, t2), smax(t1, t2))
        if overlap is not None:
            a, b, c = vx * vx + vz * vz, F(2) * (r0x * vx + r0z * vz), r0x * r0x + r0z * r0z - r_sum * r_sum
            t = earliest_root(a, b, c, overlap[0], overlap[1])
            if t is not None and (best is None or t < best):
                best = t
    if best is None:
        return None
    rel_at_hit = add(rel_start, mul(rel_delta, best))
    return best * move_len, capsule_pair_hit_normal(rel_at_hit, h_sum)


# ---------------------------------------------------------------- Game/CollisionQuery.swift:980-1283 (query layer)
# Brute force over every triangle in index order: the candidate SET of the reference's BVH walk is tree independent
# (node boxes only cull), and a strict `<` over ascending indices keeps the smallest index of equal tois — the
# canonical tie rule.  `tris` = list of (v0, v1, v2, layer) in triangle-index order, static set first.
def _vmin(a, b):
    return (min(a[0], b[0]), min(a[1], b[1]), min(a[2], b[2]))


def _vmax(a, b):
    return (max(a[0], b[0]), max(a[1], b[1]), max(a[2], b[2]))


def _disjoint(tlo, thi, lo, hi):
    return (thi[0] < lo[0] or tlo[0] > hi[0] or thi[1] < lo[1] or tlo[1] > hi[1] or thi[2] < lo[2] or tlo[2] > hi[2])


def capsule_cast(tris, n_static, frm, delta, radius, half_height, mask, blocking_only, min_normal_y):
    """capsuleCastCombined (:980-1117).  Returns None or (toi, position, normal, triangleNormal, triangleIndex)."""
    seg_len = np.sqrt(dot(delta, delta))
    if seg_len < F(1e-6):
        return None
    direction = (delta[0] / seg_len, delta[1] / seg_len, delta[2] / seg_len)
    up = v(0, 1, 0)
    a0, b0 = add(frm, mul(up, half_height)), sub(frm, mul(up, half_height))
    a1, b1 = add(a0, delta), add(b0, delta)
    ext = (radius, radius, radius)
    lo = sub(_vmin(_vmin(a0, b0), _vmin(a1, b1)), ext)
    hi = add(_vmax(_vmax(a0, b0), _vmax(a1, b1)), ext)
    best = [None, None]  # per set, then chooseNearest (:909-914)
    best_t = [seg_len, seg_len]
    for index, (v0, v1, v2, layer) in enumerate(tris):
        which = 0 if index < n_static else 1
        if (int(layer) & int(mask)) == 0:
            continue
        if _disjoint(_vmin(v0, _vmin(v1, v2)), _vmax(v0, _vmax(v1, v2)), lo, hi):
            continue
        hit = sweep_capsule_triangle(frm, direction, seg_len, radius, half_height, v0, v1, v2)
        if hit is None or not hit[0] < best_t[which]:
            continue
        toi, pos, nrm, tri_n, _ = hit
        if blocking_only and (dot(delta, nrm) >= ZERO or dot(delta, tri_n) >= ZERO):
            continue
        if min_normal_y is not None and tri_n[1] < min_normal_y:
            continue
        best_t[which] = toi
        best[which] = (toi, pos, nrm, tri_n, index)
    if best[0] is not None and best[1] is not None:
        return best[0] if best[0][0] <= best[1][0] else best[1]
    return best[0] if best[0] is not None else best[1]


def capsule_overlap_all(tris, frm, radius, half_height, mask, max_hits):
    """capsuleOverlapAll (:1119-1283) under the canonical rule: the max_hits deepest, deepest first, ties by index."""
    up = v(0, 1, 0)
    a0, b0 = add(frm, mul(up, half_height)), sub(frm, mul(up, half_height))
    ext = (radius, radius, radius)
    lo, hi = sub(_vmin(a0, b0), ext), add(_vmax(a0, b0), ext)
    hits = []
    for index, (v0, v1, v2, layer) in enumerate(tris):
        if (int(layer) & int(mask)) == 0:
            continue
        if _disjoint(_vmin(v0, _vmin(v1, v2)), _vmax(v0, _vmax(v1, v2)), lo, hi):
            continue
        dist, seg_pt, tri_pt = segment_triangle_distance(frm, half_height, v0, v1, v2)
        if dist >= radius:
            continue
        tri_normal = normalize(cross(sub(v1, v0), sub(v2, v0)))
        n = tri_normal if dist < F(1e-6) else normalize(sub(seg_pt, tri_pt))
        tri_n = neg(tri_normal) if dot(tri_normal, n) < ZERO else tri_normal
        hits.append((radius - dist, tri_pt, n, tri_n, index))
    hits.sort(key=lambda h: (-float(h[0]), h[4]))
    return hits[:max_hits]


# ---------------------------------------------------------------- Game/CollisionQuery.swift:331-417, 768-785, 916-978
def build_soup(parts):
    """TriangleMeshSet.rebuild for one set: world-space vertices = simd_mul(modelMatrix, (p, 1)) with the documented
    column order ((c0*x + c1*y) + c2*z) + c3*1, triangles with |e1 x e2|^2 <= 1e-10 dropped, numbering = what is left.
    parts: dicts with positions (n,3), indices, model (16 floats column-major), layer.  Returns (positions, triangles)
    with triangles = [(i0, i1, i2, layer)]."""
    positions, triangles = [], []
    for prt in parts:
        m = [F(x) for x in np.asarray(prt["model"], np.float32).reshape(-1)]
        base = len(positions)
        for p in np.asarray(prt["positions"], np.float32).reshape(-1, 3):
            x, y, z = F(p[0]), F(p[1]), F(p[2])
            positions.append(tuple(((m[r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r] * ONE for r in range(3)))
        idx = np.asarray(prt["indices"], np.uint32).reshape(-1)
        t = 0
        while t + 2 < len(idx):
            i0, i1, i2 = base + int(idx[t]), base + int(idx[t + 1]), base + int(idx[t + 2])
            e1, e2 = sub(positions[i1], positions[i0]), sub(positions[i2], positions[i0])
            if not length_squared(cross(e1, e2)) <= F(1e-10):
                triangles.append((i0, i1, i2, int(prt["layer"])))
            t += 3
    return positions, triangles


def raycast(tris, n_static, origin, direction, max_distance, mask):
    """raycast (:768-785, 916-978) as the minimum over all triangles (canonical rule): rayTriangle two-sided with
    eps 1e-6, accepted when t < closest (strict), per set, then chooseNearest.  Returns None or (t, index)."""
    best = [None, None]
    closest = [max_distance, max_distance]
    for index, (v0, v1, v2, layer) in enumerate(tris):
        which = 0 if index < n_static else 1
        if (int(layer) & int(mask)) == 0:
            continue
        e1, e2 = sub(v1, v0), sub(v2, v0)
        pvec = cross(direction, e2)
        det = dot(e1, pvec)
        if abs(det) < F(1e-6):
            continue
        inv_det = ONE / det
        tvec = sub(origin, v0)
        u = dot(tvec, pvec) * inv_det
        if u < ZERO or u > ONE:
            continue
        qvec = cross(tvec, e1)
        vv = dot(direction, qvec) * inv_det
        if vv < ZERO or (u + vv) > ONE:
            continue
        t = dot(e2, qvec) * inv_det
        if t >= ZERO and t < closest[which]:
            closest[which] = t
            best[which] = (t, index)
    if best[0] is not None and best[1] is not None:
        return best[0] if best[0][0] <= best[1][0] else best[1]
    return best[0] if best[0] is not None else best[1]
