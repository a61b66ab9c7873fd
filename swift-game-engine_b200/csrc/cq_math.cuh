// cq_math.cuh — float3 helpers and the capsule/ray narrow phase, device side.
//
// Arithmetic contract: this file is compiled with -fmad=false
// -prec-div=true -prec-sqrt=true, so every a*b+c below is a separate IEEE multiply and
// add, exactly as the reference's Swift evaluates it on scalars.  Expression ORDER
// follows the reference (Game/CollisionQuery.swift, cited per function); scheduling,
// data layout, pruning and tie rules are this library's own.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

// host+device so that tests can run the very same source on the CPU against the oracle (tests/hostmath)
#define CQ_HD __host__ __device__ __forceinline__

namespace cq {

struct f3 {
    float x, y, z;
};
CQ_HD f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
CQ_HD f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
CQ_HD f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
CQ_HD f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
CQ_HD f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
CQ_HD f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// simd_dot: left-to-right, no contraction
CQ_HD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
CQ_HD f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
CQ_HD float len2(f3 a) { return dot(a, a); }
CQ_HD float len(f3 a) { return sqrtf(dot(a, a)); }
CQ_HD f3 normalize(f3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
CQ_HD f3 vmin(f3 a, f3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
CQ_HD f3 vmax(f3 a, f3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
// Swift's generic max/min on Comparable (max(x,y) = y >= x ? y : x ; min(x,y) = y < x ? y : x)
CQ_HD float smax(float x, float y) { return y >= x ? y : x; }
CQ_HD float smin(float x, float y) { return y < x ? y : x; }
CQ_HD float clamp01(float v) { return smin(smax(v, 0.0f), 1.0f); }
CQ_HD f3 xyz(float4 v) { return {v.x, v.y, v.z}; }

struct d3 {
    double x, y, z;
};
CQ_HD d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
CQ_HD d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
CQ_HD d3 operator*(d3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
CQ_HD double dot(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
CQ_HD d3 to_d3(f3 v) { return {(double)v.x, (double)v.y, (double)v.z}; }
CQ_HD f3 to_f3(d3 v) { return {(float)v.x, (float)v.y, (float)v.z}; }

// ---- triangle
struct Tri {
    f3 v0, v1, v2;
};

struct SegTriResult {
    float dist;
    f3 seg, tri;
};

// All three primitives below are written BRANCH-FREE (selects instead of early returns): 32 lanes of a
// warp evaluate 32 different (capsule, triangle) pairs, each landing in a different Voronoi region, so
// branchy code serialises into up to 7 passes (measured: 13-17 of 32 lanes active inside the distance
// function).  Every value that is finally selected is computed by exactly the reference's expression, in
// the reference's order; values of regions that are not selected are computed speculatively and dropped
// (they may be inf/NaN: nothing traps on the device).  NaN behaviour of the comparisons is preserved by
// keeping the reference's comparison direction (e.g. !(u < 0 || u > 1), not (u >= 0 && u <= 1)).

// closestPointOnTriangle — CollisionQuery.swift:1464-1517 (Ericson's Voronoi-region walk).
// Region priority as in the reference: A, B, AB, C, AC, BC, face.
CQ_HD float closest_point_on_triangle(f3 p, f3 a, f3 b, f3 c, f3 &out) {
    f3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    f3 bp = p - b;
    float d3_ = dot(ab, bp), d4 = dot(ac, bp);
    f3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    float vc = d1 * d4 - d3_ * d2;
    float vb = d5 * d2 - d1 * d6;
    float va = d3_ * d6 - d5 * d4;
    float d43 = d4 - d3_, d56 = d5 - d6;
    bool rA = d1 <= 0.0f && d2 <= 0.0f;
    bool rB = d3_ >= 0.0f && d4 <= d3_;
    bool rAB = vc <= 0.0f && d1 >= 0.0f && d3_ <= 0.0f;
    bool rC = d6 >= 0.0f && d5 <= d6;
    bool rAC = vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f;
    bool rBC = va <= 0.0f && d43 >= 0.0f && d56 >= 0.0f;
    // first true region wins
    bool sA = rA, sB = !rA && rB, sAB = !rA && !rB && rAB;
    bool n3 = !rA && !rB && !rAB;
    bool sC = n3 && rC, sAC = n3 && !rC && rAC, sBC = n3 && !rC && !rAC && rBC;
    bool sEdge = sAB || sAC || sBC;
    bool sVert = sA || sB || sC;
    // one division serves all regions: v = d1/(d1-d3) | w = d2/(d2-d6) | w = (d4-d3)/((d4-d3)+(d5-d6)) |
    // denom = 1/(va+vb+vc)
    float num = sAB ? d1 : (sAC ? d2 : (sBC ? d43 : 1.0f));
    float den = sAB ? (d1 - d3_) : (sAC ? (d2 - d6) : (sBC ? (d43 + d56) : (va + vb + vc)));
    float q = num / den;
    f3 base = sBC ? b : a;
    f3 dv = sAB ? ab : (sAC ? ac : (c - b));
    f3 pe = base + dv * q;                          // a + ab*v | a + ac*w | b + (c-b)*w
    f3 pf = a + ab * (vb * q) + ac * (vc * q);      // a + ab*v + ac*w with v = vb*denom, w = vc*denom
    f3 pv = sA ? a : (sB ? b : c);
    f3 pt = sVert ? pv : (sEdge ? pe : pf);
    out = pt;
    return len2(p - pt);
}

// segmentSegmentDistanceSq — CollisionQuery.swift:1519-1569
CQ_HD float segment_segment_dist2(f3 p1, f3 q1, f3 p2, f3 q2, f3 &c1, f3 &c2) {
    f3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    const float eps = 1e-6f;
    if (a <= eps || e <= eps) { // degenerate segment(s): rare, kept as the reference's branches (:1534-1547)
        if (a <= eps && e <= eps) {
            c1 = p1;
            c2 = p2;
            return len2(p1 - p2);
        }
        if (a <= eps) {
            float t = clamp01(f / e);
            c1 = p1;
            c2 = p2 + d2 * t;
            return len2(p1 - c2);
        }
        float c = dot(d1, r);
        float s = clamp01(-c / a);
        c1 = p1 + d1 * s;
        c2 = p2;
        return len2(c1 - p2);
    }
    float c = dot(d1, r);
    float b = dot(d1, d2);
    float denom = a * e - b * b;
    float s0 = (denom != 0.0f) ? clamp01((b * f - c * e) / denom) : 0.0f;
    float tNom = b * s0 + f;
    bool lo = tNom < 0.0f, hi = !lo && tNom > e;
    // one division: -c/a | (b-c)/a | tNom/e
    float num = lo ? -c : (hi ? (b - c) : tNom);
    float den = (lo || hi) ? a : e;
    float q = num / den;
    float s = (lo || hi) ? clamp01(q) : s0;
    float t = lo ? 0.0f : (hi ? 1.0f : q);
    c1 = p1 + d1 * s;
    c2 = p2 + d2 * t;
    return len2(c1 - c2);
}

// ---- vertical-axis specialisations ------------------------------------------------------------------
// The capsule axis is world +Y, so the segment handed to the two functions below is a = c + (0,hh,0),
// b = c - (0,hh,0) and d1 = b - a = (0, dy, 0) with dy = b.y - a.y (b.x - a.x and b.z - a.z are exactly
// zero: both are c.x -+ 0*hh).  Multiplying by an exact zero and adding the resulting (signed) zero never
// changes a non-zero value, so the reference's general expressions reduce EXACTLY (same roundings) to the
// shorter ones used here; only the sign of a zero result can differ, which no comparison and no non-zero
// output depends on.  tests/test_device_math_on_host.py checks these against the oracle's literal,
// unspecialised restatement on millions of inputs.

// segmentSegmentDistanceSq(p1: a, q1: b, p2, q2) — CollisionQuery.swift:1519-1569, d1 = (0, dy, 0), aa = dy*dy
CQ_HD float vseg_segment_dist2(f3 p1, float dy, float aa, f3 p2, f3 q2, f3 &c1, f3 &c2) {
    f3 d2 = q2 - p2, r = p1 - p2;
    float e = dot(d2, d2), f = dot(d2, r);
    const float eps = 1e-6f;
    float c = dy * r.y; // dot(d1, r)
    if (aa <= eps || e <= eps) { // degenerate segment(s): rare, kept as the reference's branches (:1534-1547)
        if (aa <= eps && e <= eps) {
            c1 = p1;
            c2 = p2;
            return len2(p1 - p2);
        }
        if (aa <= eps) {
            float t = clamp01(f / e);
            c1 = p1;
            c2 = p2 + d2 * t;
            return len2(p1 - c2);
        }
        float s = clamp01(-c / aa);
        c1 = mk3(p1.x, p1.y + dy * s, p1.z);
        c2 = p2;
        return len2(c1 - p2);
    }
    float b = dy * d2.y; // dot(d1, d2)
    float denom = aa * e - b * b;
    float s0 = (denom != 0.0f) ? clamp01((b * f - c * e) / denom) : 0.0f;
    float tNom = b * s0 + f;
    bool lo = tNom < 0.0f, hi = !lo && tNom > e;
    float num = lo ? -c : (hi ? (b - c) : tNom); // one division: -c/a | (b-c)/a | tNom/e
    float den = (lo || hi) ? aa : e;
    float q = num / den;
    float s = (lo || hi) ? clamp01(q) : s0;
    float t = lo ? 0.0f : (hi ? 1.0f : q);
    c1 = mk3(p1.x, p1.y + dy * s, p1.z); // p1 + d1*s
    c2 = p2 + d2 * t;
    return len2(c1 - c2);
}

// segmentTriangleIntersect(a, b, …) — CollisionQuery.swift:1440-1462, dir = (0, dy, 0)
CQ_HD bool vseg_triangle_intersect(f3 a, float dy, const Tri &T, f3 &out) {
    f3 e1 = T.v1 - T.v0, e2 = T.v2 - T.v0;
    float pvx = dy * e2.z, pvz = -(dy * e2.x); // cross(dir, e2) = (dy*e2.z, 0, -dy*e2.x)
    float det = e1.x * pvx + e1.z * pvz;       // dot(e1, pvec)
    float invDet = 1.0f / det;
    f3 tvec = a - T.v0;
    float u = (tvec.x * pvx + tvec.z * pvz) * invDet;
    f3 qvec = cross(tvec, e1);
    float v = (dy * qvec.y) * invDet; // dot(dir, qvec)
    float t = dot(e2, qvec) * invDet;
    out = mk3(a.x, a.y + dy * t, a.z); // a + dir*t
    return !(fabsf(det) < 1e-6f) && !(u < 0.0f || u > 1.0f) && !(v < 0.0f || (u + v) > 1.0f) && !(t < 0.0f || t > 1.0f);
}

// general-direction versions (kept for reference / tests of the specialisations)
CQ_HD bool segment_triangle_intersect(f3 a, f3 b, const Tri &T, f3 &out) {
    f3 dir = b - a;
    f3 e1 = T.v1 - T.v0, e2 = T.v2 - T.v0;
    f3 pvec = cross(dir, e2);
    float det = dot(e1, pvec);
    float invDet = 1.0f / det;
    f3 tvec = a - T.v0;
    float u = dot(tvec, pvec) * invDet;
    f3 qvec = cross(tvec, e1);
    float v = dot(dir, qvec) * invDet;
    float t = dot(e2, qvec) * invDet;
    out = a + dir * t;
    return !(fabsf(det) < 1e-6f) && !(u < 0.0f || u > 1.0f) && !(v < 0.0f || (u + v) > 1.0f) && !(t < 0.0f || t > 1.0f);
}

// segmentTriangleDistance — CollisionQuery.swift:1396-1438.  The capsule axis is world +Y.
// The two point-triangle and three segment-edge tests are ROLLED loops (#pragma unroll 1, the triangle is
// rotated in registers between trips): the fully inlined form is ~15 KB of straight-line SASS per
// evaluation, and ncu showed 40% of the stall samples as stall_no_inst (instruction-fetch bound) with four
// warps per scheduler streaming it through the ~6 KB L0 I-cache.  Candidate order is the reference's:
// point a, point b, edge v0v1, edge v1v2, edge v2v0, strict '<'.
template <bool WANT_POINTS>
CQ_HD float segment_triangle_distance(f3 center, float hh, const Tri &T, f3 &segPt, f3 &triPt) {
    const f3 up = {0.0f, 1.0f, 0.0f};
    const f3 a = center + up * hh;
    const f3 b = center - up * hh;
    const float dy = b.y - a.y; // (b - a) = (0, dy, 0)
    const float aa = dy * dy;   // dot(b - a, b - a)
    f3 hit;
    bool pierced = vseg_triangle_intersect(a, dy, T, hit);
    float best = FLT_MAX;
    f3 bs = a, bt = T.v0;
    f3 p = a;
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        f3 q;
        float d = closest_point_on_triangle(p, T.v0, T.v1, T.v2, q);
        if (d < best) {
            best = d;
            bs = p;
            bt = q;
        }
        p = b;
    }
    f3 e0 = T.v0, e1 = T.v1, e2 = T.v2;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        f3 s, t;
        float d = vseg_segment_dist2(a, dy, aa, e0, e1, s, t);
        if (d < best) {
            best = d;
            bs = s;
            bt = t;
        }
        f3 tmp = e0; // rotate: (v0,v1) -> (v1,v2) -> (v2,v0)
        e0 = e1;
        e1 = e2;
        e2 = tmp;
    }
    if (WANT_POINTS) {
        segPt = pierced ? hit : bs;
        triPt = pierced ? hit : bt;
    }
    return pierced ? 0.0f : sqrtf(smax(best, 0.0f));
}

struct CastHit {
    float toi;
    f3 position, normal, triNormal;
};

// rayTriangle — CollisionQuery.swift:1575-1601 (two-sided Moller-Trumbore)
CQ_HD bool ray_triangle(f3 origin, f3 direction, const Tri &T, float &tOut) {
    f3 e1 = T.v1 - T.v0, e2 = T.v2 - T.v0;
    f3 pvec = cross(direction, e2);
    float det = dot(e1, pvec);
    if (fabsf(det) < 1e-6f) return false;
    float invDet = 1.0f / det;
    f3 tvec = origin - T.v0;
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return false;
    f3 qvec = cross(tvec, e1);
    float v = dot(direction, qvec) * invDet;
    if (v < 0.0f || (u + v) > 1.0f) return false;
    float t = dot(e2, qvec) * invDet;
    if (t >= 0.0f) {
        tOut = t;
        return true;
    }
    return false;
}

// ---- capsule-capsule CCD between agents (vertical capsules; Systems.swift:1417-1590)
struct AgentHit {
    float toi;
    f3 normal;
    int other;
};

CQ_HD bool clamp_interval(float start, float end, float &s, float &e) { // SYS:1417-1424
    s = smax(start, 0.0f);
    e = smin(end, 1.0f);
    return !(e < s);
}
CQ_HD bool interval_ge(float y0, float vy, float threshold, float &s, float &e) { // SYS:1426-1436
    if (fabsf(vy) < 1e-6f) {
        s = 0.0f, e = 1.0f;
        return y0 >= threshold;
    }
    float t = (threshold - y0) / vy;
    return vy > 0.0f ? clamp_interval(t, 1.0f, s, e) : clamp_interval(0.0f, t, s, e);
}
CQ_HD bool interval_le(float y0, float vy, float threshold, float &s, float &e) { // SYS:1438-1448
    if (fabsf(vy) < 1e-6f) {
        s = 0.0f, e = 1.0f;
        return y0 <= threshold;
    }
    float t = (threshold - y0) / vy;
    return vy > 0.0f ? clamp_interval(0.0f, t, s, e) : clamp_interval(t, 1.0f, s, e);
}
CQ_HD bool earliest_root(float A, float B, float C, float tMin, float tMax, float &out) { // SYS:1450-1472
    const float eps = 1e-6f;
    if (fabsf(A) < eps) {
        if (fabsf(B) < eps) {
            out = tMin;
            return C <= 0.0f;
        }
        float t = -C / B;
        out = t;
        return t >= tMin && t <= tMax;
    }
    float disc = B * B - 4.0f * A * C;
    if (disc < 0.0f) return false;
    float sqrtD = sqrtf(disc);
    float inv2A = 1.0f / (2.0f * A);
    float t0 = (-B - sqrtD) * inv2A, t1 = (-B + sqrtD) * inv2A;
    float enter = smin(t0, t1), exit_ = smax(t0, t1);
    float s = smax(enter, tMin), e = smin(exit_, tMax);
    out = s;
    return e >= s;
}
CQ_HD float capsule_pair_separation_y(float yRel, float hSum) { // SYS:1474-1482
    if (yRel > hSum) return yRel - hSum;
    if (yRel < -hSum) return yRel + hSum;
    return 0.0f;
}
CQ_HD f3 capsule_pair_normal(f3 rel, float hSum) { // SYS:1484-1497
    float sepY = capsule_pair_separation_y(rel.y, hSum);
    f3 sep = {rel.x, sepY, rel.z};
    float l2 = len2(sep);
    if (l2 > 1e-8f) return sep / sqrtf(l2);
    f3 lateral = {rel.x, 0.0f, rel.z};
    float ll2 = len2(lateral);
    if (ll2 > 1e-8f) return lateral / sqrtf(ll2);
    return {1.0f, 0.0f, 0.0f};
}
CQ_HD bool capsule_pair_overlap(f3 rel, float rSum, float hSum) { // SYS:1499-1503
    float sepY = capsule_pair_separation_y(rel.y, hSum);
    float d2 = rel.x * rel.x + rel.z * rel.z + sepY * sepY;
    return d2 <= rSum * rSum;
}
// capsuleCapsuleSweep (SYS:1505-1590): relative motion against the two cap spheres and the cylinder band
CQ_HD bool capsule_pair_sweep(f3 from, f3 delta, float radius, float halfHeight, int other, f3 otherPos, f3 otherDelta,
                              float otherRadius, float otherHalfHeight, AgentHit &out) {
    f3 relStart = from - otherPos, relDelta = delta - otherDelta;
    float rSum = radius + otherRadius, hSum = halfHeight + otherHalfHeight;
    float relLen = len(relDelta), moveLen = len(delta);
    if (relLen < 1e-6f) {
        if (capsule_pair_overlap(relStart, rSum, hSum)) {
            out.toi = 0.0f, out.normal = capsule_pair_normal(relStart, hSum), out.other = other;
            return true;
        }
        return false;
    }
    float y0 = relStart.y, vy = relDelta.y, vx = relDelta.x, vz = relDelta.z, r0x = relStart.x, r0z = relStart.z;
    bool have = false;
    float bestT = 0.0f, s, e, t;
    if (interval_ge(y0, vy, hSum, s, e)) {
        float A = vx * vx + vz * vz + vy * vy;
        float B = 2.0f * (r0x * vx + r0z * vz + (y0 - hSum) * vy);
        float C = r0x * r0x + r0z * r0z + (y0 - hSum) * (y0 - hSum) - rSum * rSum;
        if (earliest_root(A, B, C, s, e, t)) bestT = t, have = true;
    }
    if (interval_le(y0, vy, -hSum, s, e)) {
        float A = vx * vx + vz * vz + vy * vy;
        float B = 2.0f * (r0x * vx + r0z * vz + (y0 + hSum) * vy);
        float C = r0x * r0x + r0z * r0z + (y0 + hSum) * (y0 + hSum) - rSum * rSum;
        if (earliest_root(A, B, C, s, e, t) && (!have || t < bestT)) bestT = t, have = true;
    }
    if (fabsf(vy) < 1e-6f) {
        if (fabsf(y0) <= hSum) {
            float A = vx * vx + vz * vz, B = 2.0f * (r0x * vx + r0z * vz), C = r0x * r0x + r0z * r0z - rSum * rSum;
            if (earliest_root(A, B, C, 0.0f, 1.0f, t) && (!have || t < bestT)) bestT = t, have = true;
        }
    } else {
        float t1 = (hSum - y0) / vy, t2 = (-hSum - y0) / vy;
        if (clamp_interval(smin(t1, t2), smax(t1, t2), s, e)) {
            float A = vx * vx + vz * vz, B = 2.0f * (r0x * vx + r0z * vz), C = r0x * r0x + r0z * r0z - rSum * rSum;
            if (earliest_root(A, B, C, s, e, t) && (!have || t < bestT)) bestT = t, have = true;
        }
    }
    if (!have) return false;
    f3 relAtHit = relStart + relDelta * bestT;
    out.toi = bestT * moveLen, out.normal = capsule_pair_normal(relAtHit, hSum), out.other = other;
    return true;
}

} // namespace cq
