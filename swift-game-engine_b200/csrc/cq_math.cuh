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

namespace cq {

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// simd_dot: left-to-right, no contraction
__device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float len2(f3 a) { return dot(a, a); }
__device__ __forceinline__ float len(f3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ f3 normalize(f3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
__device__ __forceinline__ f3 vmin(f3 a, f3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
__device__ __forceinline__ f3 vmax(f3 a, f3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
// Swift's generic max/min on Comparable (max(x,y) = y >= x ? y : x ; min(x,y) = y < x ? y : x)
__device__ __forceinline__ float smax(float x, float y) { return y >= x ? y : x; }
__device__ __forceinline__ float smin(float x, float y) { return y < x ? y : x; }
__device__ __forceinline__ float clamp01(float v) { return smin(smax(v, 0.0f), 1.0f); }
__device__ __forceinline__ f3 xyz(float4 v) { return {v.x, v.y, v.z}; }

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ d3 to_d3(f3 v) { return {(double)v.x, (double)v.y, (double)v.z}; }
__device__ __forceinline__ f3 to_f3(d3 v) { return {(float)v.x, (float)v.y, (float)v.z}; }

// ---- triangle with the per-triangle constants the distance function reuses -----------
struct Tri {
    f3 v0, v1, v2;
};

struct SegTriResult {
    float dist;
    f3 seg, tri;
};

// closestPointOnTriangle — CollisionQuery.swift:1464-1517 (Ericson's Voronoi-region walk)
__device__ __forceinline__ float closest_point_on_triangle(f3 p, f3 a, f3 b, f3 c, f3 &out) {
    f3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) {
        out = a;
        return len2(p - a);
    }
    f3 bp = p - b;
    float d3_ = dot(ab, bp), d4 = dot(ac, bp);
    if (d3_ >= 0.0f && d4 <= d3_) {
        out = b;
        return len2(p - b);
    }
    float vc = d1 * d4 - d3_ * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3_ <= 0.0f) {
        float v = d1 / (d1 - d3_);
        out = a + ab * v;
        return len2(p - out);
    }
    f3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) {
        out = c;
        return len2(p - c);
    }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        out = a + ac * w;
        return len2(p - out);
    }
    float va = d3_ * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3_) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
        out = b + (c - b) * w;
        return len2(p - out);
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    out = a + ab * v + ac * w;
    return len2(p - out);
}

// segmentSegmentDistanceSq — CollisionQuery.swift:1519-1569
__device__ __forceinline__ float segment_segment_dist2(f3 p1, f3 q1, f3 p2, f3 q2, f3 &c1, f3 &c2) {
    f3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    const float eps = 1e-6f;
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
    if (e <= eps) {
        float s = clamp01(-c / a);
        c1 = p1 + d1 * s;
        c2 = p2;
        return len2(c1 - p2);
    }
    float b = dot(d1, d2);
    float denom = a * e - b * b;
    float s = (denom != 0.0f) ? clamp01((b * f - c * e) / denom) : 0.0f;
    float t;
    float tNom = b * s + f;
    if (tNom < 0.0f) {
        t = 0.0f;
        s = clamp01(-c / a);
    } else if (tNom > e) {
        t = 1.0f;
        s = clamp01((b - c) / a);
    } else {
        t = tNom / e;
    }
    c1 = p1 + d1 * s;
    c2 = p2 + d2 * t;
    return len2(c1 - c2);
}

// segmentTriangleIntersect — CollisionQuery.swift:1440-1462
__device__ __forceinline__ bool segment_triangle_intersect(f3 a, f3 b, const Tri &T, f3 &out) {
    f3 dir = b - a;
    f3 e1 = T.v1 - T.v0, e2 = T.v2 - T.v0;
    f3 pvec = cross(dir, e2);
    float det = dot(e1, pvec);
    if (fabsf(det) < 1e-6f) return false;
    float invDet = 1.0f / det;
    f3 tvec = a - T.v0;
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return false;
    f3 qvec = cross(tvec, e1);
    float v = dot(dir, qvec) * invDet;
    if (v < 0.0f || (u + v) > 1.0f) return false;
    float t = dot(e2, qvec) * invDet;
    if (t < 0.0f || t > 1.0f) return false;
    out = a + dir * t;
    return true;
}

// segmentTriangleDistance — CollisionQuery.swift:1396-1438.  The capsule axis is world +Y.
template <bool WANT_POINTS>
__device__ __forceinline__ float segment_triangle_distance(f3 center, float hh, const Tri &T, f3 &segPt, f3 &triPt) {
    const f3 up = {0.0f, 1.0f, 0.0f};
    f3 a = center + up * hh;
    f3 b = center - up * hh;
    f3 hit;
    if (segment_triangle_intersect(a, b, T, hit)) {
        if (WANT_POINTS) {
            segPt = hit;
            triPt = hit;
        }
        return 0.0f;
    }
    float best = FLT_MAX;
    f3 bs = a, bt = T.v0;
    f3 p;
    float d = closest_point_on_triangle(a, T.v0, T.v1, T.v2, p);
    if (d < best) {
        best = d;
        if (WANT_POINTS) {
            bs = a;
            bt = p;
        }
    }
    d = closest_point_on_triangle(b, T.v0, T.v1, T.v2, p);
    if (d < best) {
        best = d;
        if (WANT_POINTS) {
            bs = b;
            bt = p;
        }
    }
    f3 s, t;
    d = segment_segment_dist2(a, b, T.v0, T.v1, s, t);
    if (d < best) {
        best = d;
        if (WANT_POINTS) {
            bs = s;
            bt = t;
        }
    }
    d = segment_segment_dist2(a, b, T.v1, T.v2, s, t);
    if (d < best) {
        best = d;
        if (WANT_POINTS) {
            bs = s;
            bt = t;
        }
    }
    d = segment_segment_dist2(a, b, T.v2, T.v0, s, t);
    if (d < best) {
        best = d;
        if (WANT_POINTS) {
            bs = s;
            bt = t;
        }
    }
    if (WANT_POINTS) {
        segPt = bs;
        triPt = bt;
    }
    return sqrtf(smax(best, 0.0f));
}

struct CastHit {
    float toi;
    f3 position, normal, triNormal;
};

// sweepCapsuleTriangle + refineTOI — CollisionQuery.swift:1285-1394, with the exact-safe prune of
// a candidate is abandoned as soon as its time of impact is provably > pruneT
// (the best accepted toi so far): lastSafeT > pruneT in the advancement loop, lo > pruneT in the
// bisection (the returned hi >= lo).  Such a candidate could never be accepted (needs toi <= bestT),
// so results are unchanged; only the work counters differ from the reference's.
// Returns 0 = no hit (or pruned), 1 = hit.  `evals` counts distance evaluations when COUNT.
template <bool COUNT>
__device__ __forceinline__ int sweep_capsule_triangle(f3 from, f3 dir, float maxDistance, float radius, float hh,
                                                      const Tri &T, float pruneT, CastHit &out, uint32_t &evals) {
    float minAdvance = smax(radius * 0.02f, 1e-4f);
    int maxIter = min(256, (int)ceilf(maxDistance / minAdvance) + 1);
    const float contactEps = 1e-5f;
    float t = 0.0f, lastSafeT = 0.0f;
    f3 dummyA, dummyB;
    for (int it = 0; it < maxIter; it++) {
        if (t > maxDistance) return 0;
        f3 center = from + dir * t;
        if (COUNT) evals++;
        float dist = segment_triangle_distance<false>(center, hh, T, dummyA, dummyB);
        if (dist <= radius + contactEps) {
            // refineTOI (CollisionQuery.swift:1361-1394)
            float c0 = smax(0.0f, smin(lastSafeT, maxDistance));
            float c1 = smax(0.0f, smin(t, maxDistance));
            float lo = smin(c0, c1), hi = smax(c0, c1);
            if (!(hi - lo < 1e-5f)) {
                for (int k = 0; k < 10; k++) {
                    if (lo > pruneT) return 0;
                    float mid = 0.5f * (lo + hi);
                    f3 cm = from + dir * mid;
                    if (COUNT) evals++;
                    float dm = segment_triangle_distance<false>(cm, hh, T, dummyA, dummyB);
                    if (dm <= radius) hi = mid;
                    else lo = mid;
                }
            }
            float tHit = hi;
            if (tHit > pruneT) return 0; // cannot be accepted (needs toi <= best); skip the contact evaluation
            f3 hc = from + dir * tHit;
            f3 hs, ht;
            if (COUNT) evals++;
            float hd = segment_triangle_distance<true>(hc, hh, T, hs, ht);
            f3 triNormal = normalize(cross(T.v1 - T.v0, T.v2 - T.v0));
            f3 n;
            if (hd < 1e-6f) n = dot(triNormal, dir) > 0.0f ? -triNormal : triNormal;
            else n = normalize(hs - ht);
            f3 triN = triNormal;
            if (dot(triN, n) < 0.0f) triN = -triN;
            out.toi = tHit;
            out.position = ht;
            out.normal = n;
            out.triNormal = triN;
            return 1;
        }
        lastSafeT = t;
        if (lastSafeT > pruneT) return 0;
        float advance = smax(dist - radius, minAdvance);
        if (advance <= 0.0f) t += minAdvance;
        else t += advance;
    }
    return 0;
}

// rayTriangle — CollisionQuery.swift:1575-1601 (two-sided Moller-Trumbore)
__device__ __forceinline__ bool ray_triangle(f3 origin, f3 direction, const Tri &T, float &tOut) {
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

} // namespace cq
