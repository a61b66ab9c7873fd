"""A SECOND, independent restatement of the reference's BVH (Game/CollisionQuery.swift:487-707: top-down median split
of the centroid bounds, Hoare-style partition, sort fallback, preorder node numbering, refit) and of its three
stack-driven walks (raycastBVH :916-978 with rayAABB :1603-1631, capsuleCastBVH :1011-1117, capsuleOverlapBVHAll
:1201-1283: push left then right, so the RIGHT child is visited first; first visited wins exact ties; overlap-all stops
at maxHits in visiting order), written straight from the Swift in Python with numpy.float32 arithmetic.

TEST INFRASTRUCTURE ONLY (see tests/independent_narrow_phase.py for why): it pins the oracle's ORDER_REFERENCE mode —
the tree- and order-dependent part of the reference's answers — against a separately written transliteration.
"""
import numpy as np

import independent_narrow_phase as ind

F = np.float32
ZERO, ONE, HALF = F(0), F(1), F(0.5)
FLT_MAX = np.finfo(np.float32).max


def _vmin(a, b):
    return (ind.smin(a[0], b[0]), ind.smin(a[1], b[1]), ind.smin(a[2], b[2]))


def _vmax(a, b):
    return (ind.smax(a[0], b[0]), ind.smax(a[1], b[1]), ind.smax(a[2], b[2]))


def tri_aabb(p0, p1, p2):
    """simd_min(p0, simd_min(p1, p2)) / simd_max (:391-392, :452-453)."""
    return _vmin(p0, _vmin(p1, p2)), _vmax(p0, _vmax(p1, p2))


def _centroid(b):
    return ((b[0][0] + b[1][0]) * HALF, (b[0][1] + b[1][1]) * HALF, (b[0][2] + b[1][2]) * HALF)


class ReferenceBVH:
    """struct BVH (:496-707).  nodes: dicts bounds(lo, hi) / left / right / start / count / parent."""

    def __init__(self, aabbs):
        self.nodes = []
        self.tri_order = list(range(len(aabbs)))
        self.tri_leaf = [-1] * len(aabbs)
        self.root = -1
        if aabbs:
            self.root = self._build(aabbs, 0, len(aabbs), -1)

    def _bounds_for_range(self, aabbs, start, count):
        lo, hi = aabbs[self.tri_order[start]]
        for i in range(1, count):
            b = aabbs[self.tri_order[start + i]]
            lo, hi = _vmin(lo, b[0]), _vmax(hi, b[1])
        return lo, hi

    def _centroid_bounds_for_range(self, aabbs, start, count):
        lo = hi = _centroid(aabbs[self.tri_order[start]])
        for i in range(1, count):
            c = _centroid(aabbs[self.tri_order[start + i]])
            lo, hi = _vmin(lo, c), _vmax(hi, c)
        return lo, hi

    def _build(self, aabbs, start, count, parent):
        index = len(self.nodes)
        self.nodes.append({"bounds": self._bounds_for_range(aabbs, start, count), "left": -1, "right": -1, "start": start,
                           "count": count, "parent": parent})
        if count <= 4:  # leafTriangleLimit (:472)
            for i in range(count):
                self.tri_leaf[self.tri_order[start + i]] = index
            return index
        clo, chi = self._centroid_bounds_for_range(aabbs, start, count)
        ext = (chi[0] - clo[0], chi[1] - clo[1], chi[2] - clo[2])
        if ext[0] >= ext[1] and ext[0] >= ext[2]:
            axis = 0
        elif ext[1] >= ext[2]:
            axis = 1
        else:
            axis = 2
        pivot = (clo[axis] + chi[axis]) * HALF
        i, j = start, start + count - 1
        order = self.tri_order
        while i <= j:
            if _centroid(aabbs[order[i]])[axis] < pivot:
                i += 1
            else:
                order[i], order[j] = order[j], order[i]
                j -= 1
        end = start + count
        if i == start or i == end:  # everything fell on one side: sort by the centroid and cut in the middle (:637-653)
            order[start:end] = sorted(order[start:end], key=lambda t: _centroid(aabbs[t])[axis])
            i = start + count // 2
        left = self._build(aabbs, start, i - start, index)
        right = self._build(aabbs, i, end - i, index)
        node = self.nodes[index]
        node["left"], node["right"], node["start"], node["count"] = left, right, 0, 0
        lb, rb = self.nodes[left]["bounds"], self.nodes[right]["bounds"]
        node["bounds"] = (_vmin(lb[0], rb[0]), _vmax(lb[1], rb[1]))
        return index

    def refit(self, updated, aabbs):
        """refit (:528-575): leaf bounds from their ranges, then every ancestor, deepest first."""
        if not self.nodes:
            return
        leaves = {self.tri_leaf[t] for t in updated if self.tri_leaf[t] >= 0}
        for leaf in leaves:
            n = self.nodes[leaf]
            n["bounds"] = self._bounds_for_range(aabbs, n["start"], n["count"])
        dirty, seen = [], set()
        for leaf in leaves:
            p = self.nodes[leaf]["parent"]
            while p >= 0:
                if p not in seen:
                    seen.add(p)
                    dirty.append(p)
                p = self.nodes[p]["parent"]

        def depth(n):
            d = 0
            while n >= 0:
                d += 1
                n = self.nodes[n]["parent"]
            return d
        for p in sorted(dirty, key=depth, reverse=True):
            lb, rb = self.nodes[self.nodes[p]["left"]]["bounds"], self.nodes[self.nodes[p]["right"]]["bounds"]
            self.nodes[p]["bounds"] = (_vmin(lb[0], rb[0]), _vmax(lb[1], rb[1]))


class TriangleSet:
    """One TriangleMeshSet (:320-470) in the reference's numbering: built by independent_narrow_phase.build_soup."""

    def __init__(self, parts):
        self.parts = parts
        self.positions, self.triangles = ind.build_soup(parts)
        self.aabbs = [tri_aabb(self.positions[a], self.positions[b], self.positions[c]) for a, b, c, _ in self.triangles]
        self.bvh = ReferenceBVH(self.aabbs)
        # slices (:404-409): vertex range and surviving triangle range per part
        self.slices, v, t = [], 0, 0
        for prt in parts:
            nv = len(np.asarray(prt["positions"]).reshape(-1, 3))
            _, kept = ind.build_soup([prt])
            self.slices.append((v, v + nv, t, t + len(kept)))
            v, t = v + nv, t + len(kept)

    def update_transform(self, part_index, model):
        """updateTransforms (:419-462) for one entity: vertices, triangle AABBs (no new degenerate filter), refit."""
        v0, v1, t0, t1 = self.slices[part_index]
        prt = dict(self.parts[part_index], model=model)
        moved, _ = ind.build_soup([prt])
        self.positions[v0:v1] = moved
        for t in range(t0, t1):
            a, b, c, _ = self.triangles[t]
            self.aabbs[t] = tri_aabb(self.positions[a], self.positions[b], self.positions[c])
        if t1 > t0:
            self.bvh.refit(list(range(t0, t1)), self.aabbs)

    def tri(self, t):
        a, b, c, layer = self.triangles[t]
        return self.positions[a], self.positions[b], self.positions[c], layer


def _box_disjoint(b, lo, hi):
    return (b[1][0] < lo[0] or b[0][0] > hi[0] or b[1][1] < lo[1] or b[0][1] > hi[1] or b[1][2] < lo[2] or b[0][2] > hi[2])


def _leaf_triangles(s, node):
    return [s.bvh.tri_order[i] for i in range(node["start"], node["start"] + node["count"])]


def capsule_cast_set(s, offset, frm, delta, radius, half_height, mask, blocking_only, min_normal_y):
    """capsuleCastBVH (:1011-1117) on one set.  Returns None or (toi, position, normal, triangleNormal, index)."""
    if s.bvh.root < 0:
        return None
    seg_len = np.sqrt(ind.dot(delta, delta))
    if seg_len < F(1e-6):
        return None
    direction = (delta[0] / seg_len, delta[1] / seg_len, delta[2] / seg_len)
    up = ind.v(0, 1, 0)
    a0, b0 = ind.add(frm, ind.mul(up, half_height)), ind.sub(frm, ind.mul(up, half_height))
    a1, b1 = ind.add(a0, delta), ind.add(b0, delta)
    ext = (radius, radius, radius)
    lo = ind.sub(_vmin(_vmin(a0, b0), _vmin(a1, b1)), ext)
    hi = ind.add(_vmax(_vmax(a0, b0), _vmax(a1, b1)), ext)
    best, best_t = None, seg_len
    stack = [s.bvh.root]
    while stack:
        node = s.bvh.nodes[stack.pop()]
        if _box_disjoint(node["bounds"], lo, hi):
            continue
        if node["left"] >= 0:
            stack.append(node["left"])
            stack.append(node["right"])
            continue
        for t in _leaf_triangles(s, node):
            v0, v1, v2, layer = s.tri(t)
            if (int(layer) & int(mask)) == 0 or _box_disjoint(s.aabbs[t], lo, hi):
                continue
            hit = ind.sweep_capsule_triangle(frm, direction, seg_len, radius, half_height, v0, v1, v2)
            if hit is None or not hit[0] < best_t:
                continue
            toi, pos, nrm, tri_n, _ = hit
            if blocking_only and (ind.dot(delta, nrm) >= ZERO or ind.dot(delta, tri_n) >= ZERO):
                continue
            if min_normal_y is not None and tri_n[1] < min_normal_y:
                continue
            best_t, best = toi, (toi, pos, nrm, tri_n, t + offset)
    return best


def capsule_cast(static, dynamic, frm, delta, radius, half_height, mask, blocking_only=False, min_normal_y=None):
    """capsuleCastCombined (:980-1009): static then dynamic, `a.toi <= b.toi` keeps the static hit."""
    a = capsule_cast_set(static, 0, frm, delta, radius, half_height, mask, blocking_only, min_normal_y)
    b = capsule_cast_set(dynamic, len(static.triangles), frm, delta, radius, half_height, mask, blocking_only, min_normal_y)
    if a is not None and b is not None:
        return a if a[0] <= b[0] else b
    return a if a is not None else b


def capsule_overlap_all_set(s, offset, frm, radius, half_height, mask, max_hits):
    """capsuleOverlapBVHAll (:1201-1283): the first max_hits overlaps in visiting order."""
    if s.bvh.root < 0:
        return []
    up = ind.v(0, 1, 0)
    a0, b0 = ind.add(frm, ind.mul(up, half_height)), ind.sub(frm, ind.mul(up, half_height))
    ext = (radius, radius, radius)
    lo, hi = ind.sub(_vmin(a0, b0), ext), ind.add(_vmax(a0, b0), ext)
    hits, stack = [], [s.bvh.root]
    while stack:
        node = s.bvh.nodes[stack.pop()]
        if _box_disjoint(node["bounds"], lo, hi):
            continue
        if node["left"] >= 0:
            stack.append(node["left"])
            stack.append(node["right"])
            continue
        for t in _leaf_triangles(s, node):
            v0, v1, v2, layer = s.tri(t)
            if (int(layer) & int(mask)) == 0 or _box_disjoint(s.aabbs[t], lo, hi):
                continue
            dist, seg_pt, tri_pt = ind.segment_triangle_distance(frm, half_height, v0, v1, v2)
            if dist >= radius:
                continue
            tri_normal = ind.normalize(ind.cross(ind.sub(v1, v0), ind.sub(v2, v0)))
            n = tri_normal if dist < F(1e-6) else ind.normalize(ind.sub(seg_pt, tri_pt))
            tri_n = ind.neg(tri_normal) if ind.dot(tri_normal, n) < ZERO else tri_normal
            hits.append((radius - dist, tri_pt, n, tri_n, t + offset))
            if len(hits) >= max_hits:
                return hits
    return hits


def capsule_overlap_all(static, dynamic, frm, radius, half_height, mask, max_hits):
    """capsuleOverlapAll (:852-882): static hits, then the dynamic set with what is left of max_hits; the final
    sort-by-depth only runs when hits.count > maxHits, which cannot happen."""
    max_hits = max(1, max_hits)
    hits = capsule_overlap_all_set(static, 0, frm, radius, half_height, mask, max_hits)
    if len(hits) < max_hits:
        hits += capsule_overlap_all_set(dynamic, len(static.triangles), frm, radius, half_height, mask, max_hits - len(hits))
    return hits


def capsule_overlap_set(s, offset, frm, radius, half_height, mask):
    """capsuleOverlapBVH (:1119-1199): the deepest overlap; `depth <= bestDepth` is skipped, so the first visited of equally
    deep triangles stays.  bestDepth starts at 0."""
    if s.bvh.root < 0:
        return None
    up = ind.v(0, 1, 0)
    a0, b0 = ind.add(frm, ind.mul(up, half_height)), ind.sub(frm, ind.mul(up, half_height))
    ext = (radius, radius, radius)
    lo, hi = ind.sub(_vmin(a0, b0), ext), ind.add(_vmax(a0, b0), ext)
    best, best_depth, stack = None, ZERO, [s.bvh.root]
    while stack:
        node = s.bvh.nodes[stack.pop()]
        if _box_disjoint(node["bounds"], lo, hi):
            continue
        if node["left"] >= 0:
            stack.append(node["left"])
            stack.append(node["right"])
            continue
        for t in _leaf_triangles(s, node):
            v0, v1, v2, layer = s.tri(t)
            if (int(layer) & int(mask)) == 0 or _box_disjoint(s.aabbs[t], lo, hi):
                continue
            dist, seg_pt, tri_pt = ind.segment_triangle_distance(frm, half_height, v0, v1, v2)
            if dist >= radius:
                continue
            depth = radius - dist
            if depth <= best_depth:
                continue
            tri_normal = ind.normalize(ind.cross(ind.sub(v1, v0), ind.sub(v2, v0)))
            n = tri_normal if dist < F(1e-6) else ind.normalize(ind.sub(seg_pt, tri_pt))
            tri_n = ind.neg(tri_normal) if ind.dot(tri_normal, n) < ZERO else tri_normal
            best_depth, best = depth, (depth, tri_pt, n, tri_n, t + offset)
    return best


def capsule_overlap(static, dynamic, frm, radius, half_height, mask):
    """capsuleOverlap (:830-850): `a.depth >= b.depth` keeps the static hit."""
    a = capsule_overlap_set(static, 0, frm, radius, half_height, mask)
    b = capsule_overlap_set(dynamic, len(static.triangles), frm, radius, half_height, mask)
    if a is not None and b is not None:
        return a if a[0] >= b[0] else b
    return a if a is not None else b


def ray_aabb(origin, direction, bounds):
    """rayAABB (:1603-1631): slab test with 1/d, FLT_MAX for a zero component, no tmax < 0 rejection."""
    inv = [ONE / direction[k] if direction[k] != ZERO else FLT_MAX for k in range(3)]
    with np.errstate(over="ignore", invalid="ignore"):
        tmin, tmax = (bounds[0][0] - origin[0]) * inv[0], (bounds[1][0] - origin[0]) * inv[0]
        if tmin > tmax:
            tmin, tmax = tmax, tmin
        for k in (1, 2):
            kmin, kmax = (bounds[0][k] - origin[k]) * inv[k], (bounds[1][k] - origin[k]) * inv[k]
            if kmin > kmax:
                kmin, kmax = kmax, kmin
            if tmin > kmax or kmin > tmax:
                return None
            tmin, tmax = ind.smax(tmin, kmin), ind.smin(tmax, kmax)
    return tmin, tmax


def raycast_set(s, offset, origin, direction, max_distance, mask):
    """raycastBVH (:916-978).  Returns None or (t, index)."""
    if s.bvh.root < 0:
        return None
    closest, hit = max_distance, None
    stack = [s.bvh.root]
    while stack:
        node = s.bvh.nodes[stack.pop()]
        rng = ray_aabb(origin, direction, node["bounds"])
        if rng is None or rng[0] > closest:
            continue
        if node["left"] >= 0:
            stack.append(node["left"])
            stack.append(node["right"])
            continue
        for t in _leaf_triangles(s, node):
            v0, v1, v2, layer = s.tri(t)
            if (int(layer) & int(mask)) == 0:
                continue
            e1, e2 = ind.sub(v1, v0), ind.sub(v2, v0)
            pvec = ind.cross(direction, e2)
            det = ind.dot(e1, pvec)
            if abs(det) < F(1e-6):
                continue
            inv_det = ONE / det
            tvec = ind.sub(origin, v0)
            u = ind.dot(tvec, pvec) * inv_det
            if u < ZERO or u > ONE:
                continue
            qvec = ind.cross(tvec, e1)
            vv = ind.dot(direction, qvec) * inv_det
            if vv < ZERO or (u + vv) > ONE:
                continue
            tt = ind.dot(e2, qvec) * inv_det
            if tt >= ZERO and tt < closest:
                closest, hit = tt, (tt, t + offset)
    return hit


def raycast(static, dynamic, origin, direction, max_distance, mask):
    """raycast (:768-785) + chooseNearest (:902-907: `a.distance <= b.distance` keeps the static hit)."""
    a = raycast_set(static, 0, origin, direction, max_distance, mask)
    b = raycast_set(dynamic, len(static.triangles), origin, direction, max_distance, mask)
    if a is not None and b is not None:
        return a if a[0] <= b[0] else b
    return a if a is not None else b
