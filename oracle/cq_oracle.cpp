// cq_oracle.cpp — CPU oracle.  TEST INFRASTRUCTURE ONLY (see cq_oracle.h).
//
// A C++17 restatement of the reference's collision-query path, block by block,
// each block tagged with the Swift lines it follows:
//   CQ  = /root/reference/Game/CollisionQuery.swift
//   SYS = /root/reference/Game/Systems.swift
//   CMP = /root/reference/Game/Components.swift
// PARITY UNPINNED by the reference's own tests (it has none); pinned by
// tests/test_oracle_known_answers.py and tests/golden/*.
//
// Third-party arithmetic absent from /root/reference: Apple's `simd` module
// (system framework, no version pin; macOS 26 SDK per Game.xcodeproj).  The shim
// below restates its published semantics: left-to-right dot,
// IEEE sqrt/div, normalize = x * (1/sqrt(len²)), component-wise min/max.
// Build: g++ -O2 -ffp-contract=off -fno-fast-math  (no FMA contraction).

#include "cq_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------- simd shim
struct F3 {
    float x, y, z;
};
inline F3 f3(float x, float y, float z) { return F3{x, y, z}; }
inline F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 operator-(F3 a) { return {-a.x, -a.y, -a.z}; }
inline F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline F3 operator/(F3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline F3 cross(F3 a, F3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length_squared(F3 a) { return dot(a, a); }
inline float length(F3 a) { return sqrtf(dot(a, a)); }
inline F3 normalize(F3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
inline F3 vmin(F3 a, F3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline F3 vmax(F3 a, F3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
// Swift stdlib generic min/max on Comparable: max(x,y) = y >= x ? y : x ; min(x,y) = y < x ? y : x
template <class T> inline T smax(T x, T y) { return y >= x ? y : x; }
template <class T> inline T smin(T x, T y) { return y < x ? y : x; }

struct D3 {
    double x, y, z;
};
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline D3 d3(F3 v) { return {(double)v.x, (double)v.y, (double)v.z}; }       // SYS:428
inline F3 f3(D3 v) { return {(float)v.x, (float)v.y, (float)v.z}; }          // SYS:432

const F3 kUp = {0, 1, 0};

struct AABB {
    F3 mn, mx;
};
struct Material {
    float muS, muK;
    bool flatten;
};
const Material kDefaultMaterial = {0.8f, 0.6f, false}; // CMP:709

// ---------------------------------------------------------------- BVH (CQ:487-707)
struct BVHNode {
    AABB bounds;
    int left, right, start, count, parent;
};

struct BVH {
    std::vector<BVHNode> nodes;
    std::vector<int> triOrder, triLeaf;
    int root = -1;

    static F3 centroid(const AABB &b) { return (b.mn + b.mx) * 0.5f; } // CQ:700
    static AABB merge(const AABB &a, const AABB &b) { return {vmin(a.mn, b.mn), vmax(a.mx, b.mx)}; }

    void init(const std::vector<AABB> &aabbs) { // CQ:502-513
        nodes.clear();
        triOrder.resize(aabbs.size());
        for (size_t i = 0; i < aabbs.size(); i++) triOrder[i] = (int)i;
        triLeaf.assign(aabbs.size(), -1);
        root = -1;
        if (!aabbs.empty()) root = build(aabbs, 0, (int)aabbs.size(), -1);
    }

    AABB boundsForRange(const std::vector<AABB> &aabbs, int start, int count) const { // CQ:672-684
        const AABB &first = aabbs[triOrder[start]];
        F3 bmin = first.mn, bmax = first.mx;
        for (int i = 1; i < count; i++) {
            const AABB &b = aabbs[triOrder[start + i]];
            bmin = vmin(bmin, b.mn);
            bmax = vmax(bmax, b.mx);
        }
        return {bmin, bmax};
    }
    AABB centroidBoundsForRange(const std::vector<AABB> &aabbs, int start, int count) const { // CQ:686-698
        F3 first = centroid(aabbs[triOrder[start]]);
        F3 bmin = first, bmax = first;
        for (int i = 1; i < count; i++) {
            F3 c = centroid(aabbs[triOrder[start + i]]);
            bmin = vmin(bmin, c);
            bmax = vmax(bmax, c);
        }
        return {bmin, bmax};
    }

    int build(const std::vector<AABB> &aabbs, int start, int count, int parent) { // CQ:577-670
        int nodeIndex = (int)nodes.size();
        AABB bounds = boundsForRange(aabbs, start, count);
        nodes.push_back({bounds, -1, -1, start, count, parent});
        if (count <= 4) { // leafTriangleLimit CQ:473
            for (int i = 0; i < count; i++) triLeaf[triOrder[start + i]] = nodeIndex;
            return nodeIndex;
        }
        AABB cb = centroidBoundsForRange(aabbs, start, count);
        F3 extent = cb.mx - cb.mn;
        int axis;
        if (extent.x >= extent.y && extent.x >= extent.z) axis = 0;
        else if (extent.y >= extent.z) axis = 1;
        else axis = 2;
        float pivot;
        switch (axis) {
        case 0: pivot = (cb.mn.x + cb.mx.x) * 0.5f; break;
        case 1: pivot = (cb.mn.y + cb.mx.y) * 0.5f; break;
        default: pivot = (cb.mn.z + cb.mx.z) * 0.5f; break;
        }
        auto axisValue = [&](int tri) {
            F3 c = centroid(aabbs[tri]);
            return axis == 0 ? c.x : (axis == 1 ? c.y : c.z);
        };
        int i = start, j = start + count - 1; // CQ:617-634
        while (i <= j) {
            if (axisValue(triOrder[i]) < pivot) {
                i += 1;
            } else {
                std::swap(triOrder[i], triOrder[j]);
                j -= 1;
            }
        }
        int end = start + count;
        if (i == start || i == end) { // CQ:637-653 (Swift `sorted` -> std::stable_sort)
            std::stable_sort(triOrder.begin() + start, triOrder.begin() + end,
                             [&](int a, int b) { return axisValue(a) < axisValue(b); });
            i = start + count / 2;
        }
        int mid = i;
        int left = build(aabbs, start, mid - start, nodeIndex);
        int right = build(aabbs, mid, start + count - mid, nodeIndex);
        nodes[nodeIndex].left = left;
        nodes[nodeIndex].right = right;
        nodes[nodeIndex].start = 0;
        nodes[nodeIndex].count = 0;
        nodes[nodeIndex].bounds = merge(nodes[left].bounds, nodes[right].bounds);
        return nodeIndex;
    }

    void refit(const std::vector<int> &updated, const std::vector<AABB> &aabbs) { // CQ:528-575
        if (nodes.empty()) return;
        std::unordered_set<int> updatedLeaves;
        for (int tri : updated) {
            int leaf = triLeaf[tri];
            if (leaf >= 0) updatedLeaves.insert(leaf);
        }
        for (int leaf : updatedLeaves)
            nodes[leaf].bounds = boundsForRange(aabbs, nodes[leaf].start, nodes[leaf].count);
        std::vector<int> dirtyParents;
        std::unordered_set<int> dirtySet;
        for (int leaf : updatedLeaves) {
            int parent = nodes[leaf].parent;
            while (parent >= 0) {
                if (dirtySet.insert(parent).second) dirtyParents.push_back(parent);
                parent = nodes[parent].parent;
            }
        }
        if (!dirtyParents.empty()) {
            std::vector<int> depths(dirtyParents.size(), 0);
            for (size_t k = 0; k < dirtyParents.size(); k++) {
                int depth = 0, node = dirtyParents[k];
                while (node >= 0) {
                    depth += 1;
                    node = nodes[node].parent;
                }
                depths[k] = depth;
            }
            std::vector<int> order(dirtyParents.size());
            for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return depths[a] > depths[b]; });
            for (int k : order) {
                int p = dirtyParents[k];
                nodes[p].bounds = merge(nodes[nodes[p].left].bounds, nodes[nodes[p].right].bounds);
            }
        }
    }
};

// ---------------------------------------------------------------- TriangleMeshSet (CQ:320-470)
struct MeshSlice {
    int vertexLo, vertexHi, indexLo, indexHi, triLo, triHi;
};

struct PartCopy { // what the oracle keeps of a StaticMeshComponent + TransformComponent
    std::vector<float> positions;
    std::vector<uint32_t> indices;
    float model[16];
    uint32_t layer;
    Material material;
    std::vector<Material> triangleMaterials; // StaticMeshComponent.triangleMaterials (may be empty = nil)
    bool isDynamic;
    uint32_t entityId;
    int partIndex;
};

inline F3 transformPoint(const float *m, F3 p) { // simd_mul(modelMatrix, (p,1)) CQ:349-351
    // ((c0*x + c1*y) + c2*z) + c3*1
    F3 c0 = {m[0], m[1], m[2]}, c1 = {m[4], m[5], m[6]}, c2 = {m[8], m[9], m[10]}, c3 = {m[12], m[13], m[14]};
    return ((c0 * p.x + c1 * p.y) + c2 * p.z) + c3 * 1.0f;
}

struct TriangleMeshSet {
    std::vector<F3> positions;
    std::vector<uint32_t> indices;
    std::vector<AABB> triangleAABBs;
    std::vector<Material> triangleMaterials;
    std::vector<uint32_t> triangleLayers;
    std::vector<int> triangleParts;
    std::unordered_map<uint32_t, MeshSlice> slices;
    BVH bvh;
    bool hasBVH = false;

    void rebuild(const std::vector<const PartCopy *> &entities) { // CQ:331-417
        positions.clear();
        indices.clear();
        triangleAABBs.clear();
        triangleMaterials.clear();
        triangleLayers.clear();
        triangleParts.clear();
        slices.clear();
        const float areaEps = 1e-10f;
        for (const PartCopy *e : entities) {
            int baseVertex = (int)positions.size();
            int nv = (int)e->positions.size() / 3;
            for (int i = 0; i < nv; i++)
                positions.push_back(transformPoint(e->model, f3(e->positions[3 * i], e->positions[3 * i + 1],
                                                                e->positions[3 * i + 2])));
            const std::vector<uint32_t> &local = e->indices;
            const size_t triCount = local.size() / 3;
            const bool perTri = !e->triangleMaterials.empty() && e->triangleMaterials.size() == triCount; // CQ:365-369
            int indexStart = (int)indices.size();
            int triStart = (int)triangleAABBs.size();
            size_t tri = 0, triLocal = 0;
            while (tri + 2 < local.size()) {
                int i0 = (int)((uint32_t)baseVertex + local[tri]);
                int i1 = (int)((uint32_t)baseVertex + local[tri + 1]);
                int i2 = (int)((uint32_t)baseVertex + local[tri + 2]);
                F3 p0 = positions[i0], p1 = positions[i1], p2 = positions[i2];
                F3 e1 = p1 - p0, e2 = p2 - p0;
                if (length_squared(cross(e1, e2)) <= areaEps) { // CQ:385
                    tri += 3;
                    triLocal += 1;
                    continue;
                }
                indices.push_back((uint32_t)i0);
                indices.push_back((uint32_t)i1);
                indices.push_back((uint32_t)i2);
                triangleAABBs.push_back({vmin(p0, vmin(p1, p2)), vmax(p0, vmax(p1, p2))});
                triangleMaterials.push_back(perTri ? e->triangleMaterials[triLocal] : e->material); // triSource[triLocal], CQ:396
                triangleLayers.push_back(e->layer);
                triangleParts.push_back(e->partIndex);
                tri += 3;
                triLocal += 1;
            }
            int indexEnd = (int)indices.size(), triEnd = (int)triangleAABBs.size();
            if (indexEnd > indexStart && triEnd > triStart)
                slices[e->entityId] = {baseVertex, (int)positions.size(), indexStart, indexEnd, triStart, triEnd};
        }
        hasBVH = !triangleAABBs.empty();
        if (hasBVH) bvh.init(triangleAABBs);
    }

    void updateTransforms(const std::vector<const PartCopy *> &entities) { // CQ:419-462
        if (entities.empty() || triangleAABBs.empty()) return;
        std::vector<int> updated;
        for (const PartCopy *e : entities) {
            auto it = slices.find(e->entityId);
            if (it == slices.end()) continue;
            const MeshSlice &s = it->second;
            int nv = (int)e->positions.size() / 3;
            if (nv != s.vertexHi - s.vertexLo) continue;
            for (int i = 0; i < nv; i++)
                positions[s.vertexLo + i] = transformPoint(
                    e->model, f3(e->positions[3 * i], e->positions[3 * i + 1], e->positions[3 * i + 2]));
            int triIndex = s.triLo;
            int i = s.indexLo;
            while (i + 2 < s.indexHi) {
                F3 p0 = positions[indices[i]], p1 = positions[indices[i + 1]], p2 = positions[indices[i + 2]];
                triangleAABBs[triIndex] = {vmin(p0, vmin(p1, p2)), vmax(p0, vmax(p1, p2))};
                updated.push_back(triIndex);
                i += 3;
                triIndex += 1;
            }
        }
        if (!updated.empty() && hasBVH) bvh.refit(updated, triangleAABBs);
    }

    Material materialForTriangle(int t) const { // CQ:464-469
        if (t >= 0 && t < (int)triangleMaterials.size()) return triangleMaterials[t];
        return kDefaultMaterial;
    }
};

// ---------------------------------------------------------------- hit records
struct RaycastHit {
    float distance;
    F3 position, normal;
    int triangleIndex;
    Material material;
};
struct CapsuleCastHit {
    float toi;
    F3 position, normal, triangleNormal;
    int triangleIndex;
    Material material;
};
struct CapsuleOverlapHit {
    float depth;
    F3 position, normal, triangleNormal;
    int triangleIndex;
    Material material;
};

struct Stats { // CQ:292-318 + extras the roofline formula needs
    int64_t candidates = 0, sweepTests = 0, sweepIterations = 0, maxIterations = 0, distanceEvals = 0,
            nodesVisited = 0, ties = 0, overflows = 0;
    void add(const Stats &o) {
        candidates += o.candidates;
        sweepTests += o.sweepTests;
        sweepIterations += o.sweepIterations;
        maxIterations = std::max(maxIterations, o.maxIterations);
        distanceEvals += o.distanceEvals;
        nodesVisited += o.nodesVisited;
        ties += o.ties;
        overflows += o.overflows;
    }
};

// ---------------------------------------------------------------- narrow phase (CQ:1396-1631)
inline float clampf(float v, float lo, float hi) { return smin(smax(v, lo), hi); } // CQ:1571

struct DistSqPt {
    float d;
    F3 p;
};
struct SegSeg {
    float d;
    F3 s, t;
};
struct SegTri {
    float dist;
    F3 seg, tri;
};

bool segmentTriangleIntersect(F3 a, F3 b, F3 v0, F3 v1, F3 v2, F3 &out) { // CQ:1440-1462
    F3 dir = b - a;
    const float eps = 1e-6f;
    F3 e1 = v1 - v0, e2 = v2 - v0;
    F3 pvec = cross(dir, e2);
    float det = dot(e1, pvec);
    if (fabsf(det) < eps) return false;
    float invDet = 1.0f / det;
    F3 tvec = a - v0;
    float u = dot(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return false;
    F3 qvec = cross(tvec, e1);
    float v = dot(dir, qvec) * invDet;
    if (v < 0 || (u + v) > 1) return false;
    float t = dot(e2, qvec) * invDet;
    if (t < 0 || t > 1) return false;
    out = a + dir * t;
    return true;
}

DistSqPt closestPointOnTriangle(F3 p, F3 a, F3 b, F3 c) { // CQ:1464-1517
    F3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return {length_squared(p - a), a};
    F3 bp = p - b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return {length_squared(p - b), b};
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        float v = d1 / (d1 - d3);
        F3 point = a + ab * v;
        return {length_squared(p - point), point};
    }
    F3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return {length_squared(p - c), c};
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        float w = d2 / (d2 - d6);
        F3 point = a + ac * w;
        return {length_squared(p - point), point};
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        F3 point = b + (c - b) * w;
        return {length_squared(p - point), point};
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    F3 point = a + ab * v + ac * w;
    return {length_squared(p - point), point};
}

SegSeg segmentSegmentDistanceSq(F3 p1, F3 q1, F3 p2, F3 q2) { // CQ:1519-1569
    F3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    float s = 0, t = 0;
    const float eps = 1e-6f;
    if (a <= eps && e <= eps) return {length_squared(p1 - p2), p1, p2};
    if (a <= eps) {
        t = clampf(f / e, 0, 1);
        F3 c2 = p2 + d2 * t;
        return {length_squared(p1 - c2), p1, c2};
    }
    float c = dot(d1, r);
    if (e <= eps) {
        s = clampf(-c / a, 0, 1);
        F3 c1 = p1 + d1 * s;
        return {length_squared(c1 - p2), c1, p2};
    }
    float b = dot(d1, d2);
    float denom = a * e - b * b;
    if (denom != 0) s = clampf((b * f - c * e) / denom, 0, 1);
    else s = 0;
    float tNom = b * s + f;
    if (tNom < 0) {
        t = 0;
        s = clampf(-c / a, 0, 1);
    } else if (tNom > e) {
        t = 1;
        s = clampf((b - c) / a, 0, 1);
    } else {
        t = tNom / e;
    }
    F3 c1 = p1 + d1 * s, c2 = p2 + d2 * t;
    return {length_squared(c1 - c2), c1, c2};
}

SegTri segmentTriangleDistance(F3 center, float halfHeight, F3 v0, F3 v1, F3 v2, Stats *st) { // CQ:1396-1438
    if (st) st->distanceEvals++;
    F3 a = center + kUp * halfHeight;
    F3 b = center - kUp * halfHeight;
    F3 hit;
    if (segmentTriangleIntersect(a, b, v0, v1, v2, hit)) return {0, hit, hit};
    float bestDistSq = FLT_MAX;
    F3 bestSeg = a, bestTri = v0;
    DistSqPt c0 = closestPointOnTriangle(a, v0, v1, v2);
    if (c0.d < bestDistSq) {
        bestDistSq = c0.d;
        bestSeg = a;
        bestTri = c0.p;
    }
    DistSqPt c1 = closestPointOnTriangle(b, v0, v1, v2);
    if (c1.d < bestDistSq) {
        bestDistSq = c1.d;
        bestSeg = b;
        bestTri = c1.p;
    }
    const F3 edges[3][2] = {{v0, v1}, {v1, v2}, {v2, v0}};
    for (int k = 0; k < 3; k++) {
        SegSeg ss = segmentSegmentDistanceSq(a, b, edges[k][0], edges[k][1]);
        if (ss.d < bestDistSq) {
            bestDistSq = ss.d;
            bestSeg = ss.s;
            bestTri = ss.t;
        }
    }
    return {sqrtf(smax(bestDistSq, 0.0f)), bestSeg, bestTri};
}

float refineTOI(F3 from, F3 dir, float radius, float halfHeight, F3 v0, F3 v1, F3 v2, float t0, float t1,
                float maxDistance, Stats *st) { // CQ:1361-1394
    float clampT0 = smax(0.0f, smin(t0, maxDistance));
    float clampT1 = smax(0.0f, smin(t1, maxDistance));
    float lo = smin(clampT0, clampT1);
    float hi = smax(clampT0, clampT1);
    if (hi - lo < 1e-5f) return hi;
    for (int k = 0; k < 10; k++) {
        float mid = 0.5f * (lo + hi);
        F3 center = from + dir * mid;
        SegTri d = segmentTriangleDistance(center, halfHeight, v0, v1, v2, st);
        if (d.dist <= radius) hi = mid;
        else lo = mid;
    }
    return hi;
}

bool sweepCapsuleTriangle(F3 from, F3 dir, float maxDistance, float radius, float halfHeight, F3 v0, F3 v1,
                          F3 v2, int triangleIndex, int &iterations, CapsuleCastHit &out, Stats *st) { // CQ:1285-1359
    float minAdvance = smax(radius * 0.02f, 1e-4f);
    int maxIter = smin(256, (int)ceilf(maxDistance / minAdvance) + 1);
    const float contactEps = 1e-5f;
    F3 triNormal = normalize(cross(v1 - v0, v2 - v0));
    float t = 0, lastSafeT = 0;
    for (int it = 0; it < maxIter; it++) {
        iterations += 1;
        if (t > maxDistance) return false;
        F3 center = from + dir * t;
        SegTri d = segmentTriangleDistance(center, halfHeight, v0, v1, v2, st);
        if (d.dist <= radius + contactEps) {
            float tHit = refineTOI(from, dir, radius, halfHeight, v0, v1, v2, lastSafeT, t, maxDistance, st);
            F3 hitCenter = from + dir * tHit;
            SegTri h = segmentTriangleDistance(hitCenter, halfHeight, v0, v1, v2, st);
            F3 n;
            if (h.dist < 1e-6f) n = dot(triNormal, dir) > 0 ? -triNormal : triNormal;
            else n = normalize(h.seg - h.tri);
            F3 triN = triNormal;
            if (dot(triN, n) < 0) triN = -triN;
            out = {tHit, h.tri, n, triN, triangleIndex, kDefaultMaterial};
            return true;
        }
        lastSafeT = t;
        float advance = smax(d.dist - radius, minAdvance);
        if (advance <= 0) t += minAdvance;
        else t += advance;
    }
    return false;
}

bool rayTriangle(F3 origin, F3 direction, F3 v0, F3 v1, F3 v2, float eps, float &tOut) { // CQ:1575-1601
    F3 e1 = v1 - v0, e2 = v2 - v0;
    F3 pvec = cross(direction, e2);
    float det = dot(e1, pvec);
    if (fabsf(det) < eps) return false;
    float invDet = 1.0f / det;
    F3 tvec = origin - v0;
    float u = dot(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return false;
    F3 qvec = cross(tvec, e1);
    float v = dot(direction, qvec) * invDet;
    if (v < 0 || (u + v) > 1) return false;
    float t = dot(e2, qvec) * invDet;
    if (t >= 0) {
        tOut = t;
        return true;
    }
    return false;
}

bool rayAABB(F3 origin, F3 direction, const AABB &b, float &tminOut, float &tmaxOut) { // CQ:1603-1631
    float invX = direction.x != 0 ? 1.0f / direction.x : FLT_MAX;
    float invY = direction.y != 0 ? 1.0f / direction.y : FLT_MAX;
    float invZ = direction.z != 0 ? 1.0f / direction.z : FLT_MAX;
    float tmin = (b.mn.x - origin.x) * invX, tmax = (b.mx.x - origin.x) * invX;
    if (tmin > tmax) std::swap(tmin, tmax);
    float tymin = (b.mn.y - origin.y) * invY, tymax = (b.mx.y - origin.y) * invY;
    if (tymin > tymax) std::swap(tymin, tymax);
    if (tmin > tymax || tymin > tmax) return false;
    tmin = smax(tmin, tymin);
    tmax = smin(tmax, tymax);
    float tzmin = (b.mn.z - origin.z) * invZ, tzmax = (b.mx.z - origin.z) * invZ;
    if (tzmin > tzmax) std::swap(tzmin, tzmax);
    if (tmin > tzmax || tzmin > tmax) return false;
    tmin = smax(tmin, tzmin);
    tmax = smin(tmax, tzmax);
    tminOut = tmin;
    tmaxOut = tmax;
    return true;
}

inline bool aabbDisjoint(const AABB &b, F3 minP, F3 maxP) { // CQ:1047-1049
    return b.mx.x < minP.x || b.mn.x > maxP.x || b.mx.y < minP.y || b.mn.y > maxP.y || b.mx.z < minP.z ||
           b.mn.z > maxP.z;
}

// ---------------------------------------------------------------- StaticTriMesh (CQ:472-1283)
enum { ORDER_REFERENCE = 0, ORDER_CANONICAL = 1 };

struct QueryCtx {
    int order = ORDER_REFERENCE;
    Stats *stats = nullptr;
    bool tie = false;      // set by a query when its result depended on visiting order
    bool overflow = false; // overlapAll saw more than maxHits overlaps
};

struct StaticTriMesh {
    TriangleMeshSet staticSet, dynamicSet;

    // Candidate triangles of a box query in the order the chosen mode visits them.
    // REFERENCE: the reference's DFS (push left, push right => right first; leaf ranges in triOrder).
    // CANONICAL: the same set, ascending triangle index.  Layer mask and triangle-AABB tests applied here.
    static void collectBoxCandidates(const TriangleMeshSet &set, F3 minP, F3 maxP, uint32_t mask, QueryCtx &ctx,
                                     std::vector<int> &out) {
        out.clear();
        if (!set.hasBVH || set.bvh.root < 0) return;
        const BVH &bvh = set.bvh;
        std::vector<int> stack;
        stack.push_back(bvh.root);
        while (!stack.empty()) {
            int ni = stack.back();
            stack.pop_back();
            const BVHNode &node = bvh.nodes[ni];
            if (ctx.stats) ctx.stats->nodesVisited++;
            if (aabbDisjoint(node.bounds, minP, maxP)) continue;
            if (node.left < 0) {
                for (int i = node.start; i < node.start + node.count; i++) {
                    int tri = bvh.triOrder[i];
                    if ((set.triangleLayers[tri] & mask) == 0) continue;
                    if (aabbDisjoint(set.triangleAABBs[tri], minP, maxP)) continue;
                    out.push_back(tri);
                }
            } else {
                stack.push_back(node.left);
                stack.push_back(node.right);
            }
        }
        if (ctx.order == ORDER_CANONICAL) std::sort(out.begin(), out.end());
    }

    static void triVerts(const TriangleMeshSet &set, int tri, F3 &v0, F3 &v1, F3 &v2) {
        v0 = set.positions[set.indices[tri * 3]];
        v1 = set.positions[set.indices[tri * 3 + 1]];
        v2 = set.positions[set.indices[tri * 3 + 2]];
    }

    // ---- raycast (CQ:916-978)
    bool raycastBVH(F3 origin, F3 direction, float maxDistance, const TriangleMeshSet &set, int offset,
                    uint32_t mask, QueryCtx &ctx, RaycastHit &hit) const {
        if (!set.hasBVH || set.bvh.root < 0) return false;
        const BVH &bvh = set.bvh;
        const float eps = 1e-6f;
        float closestT = maxDistance;
        bool have = false;
        float tieT = -1;
        auto testTri = [&](int tri) {
            if ((set.triangleLayers[tri] & mask) == 0) return;
            F3 v0, v1, v2;
            triVerts(set, tri, v0, v1, v2);
            float t;
            if (!rayTriangle(origin, direction, v0, v1, v2, eps, t)) return;
            if (t < closestT) {
                F3 n = normalize(cross(v1 - v0, v2 - v0));
                F3 normal = dot(n, direction) > 0 ? -n : n;
                closestT = t;
                hit = {t, origin + direction * t, normal, tri + offset, set.materialForTriangle(tri)};
                have = true;
            } else if (have && t == closestT) {
                tieT = t;
            }
        };
        if (ctx.order == ORDER_CANONICAL) {
            // tree-independent definition: min t over ALL triangles, ascending index, strict <
            for (int tri = 0; tri < (int)set.triangleAABBs.size(); tri++) testTri(tri);
        } else {
            std::vector<int> stack;
            stack.push_back(bvh.root);
            while (!stack.empty()) {
                int ni = stack.back();
                stack.pop_back();
                const BVHNode &node = bvh.nodes[ni];
                if (ctx.stats) ctx.stats->nodesVisited++;
                float tmin, tmax;
                if (!rayAABB(origin, direction, node.bounds, tmin, tmax)) continue;
                if (tmin > closestT) continue;
                if (node.left < 0) {
                    for (int i = node.start; i < node.start + node.count; i++) {
                        if (ctx.stats) ctx.stats->candidates++;
                        testTri(bvh.triOrder[i]);
                    }
                } else {
                    stack.push_back(node.left);
                    stack.push_back(node.right);
                }
            }
        }
        if (have && tieT == closestT) ctx.tie = true;
        return have;
    }

    bool raycast(F3 origin, F3 direction, float maxDistance, uint32_t mask, QueryCtx &ctx, RaycastHit &out) const { // CQ:768-785
        RaycastHit a, b;
        bool ha = raycastBVH(origin, direction, maxDistance, staticSet, 0, mask, ctx, a);
        bool hb = raycastBVH(origin, direction, maxDistance, dynamicSet, (int)staticSet.triangleAABBs.size(), mask,
                             ctx, b);
        if (ha && hb) { // chooseNearest CQ:902-907
            out = a.distance <= b.distance ? a : b;
            return true;
        }
        if (ha) out = a;
        else if (hb) out = b;
        return ha || hb;
    }

    // ---- capsule cast (CQ:980-1117)
    bool capsuleCastBVH(F3 from, F3 delta, float radius, float halfHeight, const TriangleMeshSet &set, int offset,
                        bool blockingOnly, const float *minNormalY, uint32_t mask, QueryCtx &ctx,
                        CapsuleCastHit &bestHit) const {
        if (!set.hasBVH || set.bvh.root < 0) return false;
        float len = length(delta);
        if (len < 1e-6f) return false;
        F3 dir = delta / len;
        F3 a0 = from + kUp * halfHeight, b0 = from - kUp * halfHeight;
        F3 a1 = a0 + delta, b1 = b0 + delta;
        F3 minP = vmin(vmin(a0, b0), vmin(a1, b1));
        F3 maxP = vmax(vmax(a0, b0), vmax(a1, b1));
        F3 ext = {radius, radius, radius};
        minP = minP - ext;
        maxP = maxP + ext;

        bool have = false;
        float bestT = len;
        float tieT = -1;
        thread_local std::vector<int> cands;
        collectBoxCandidates(set, minP, maxP, mask, ctx, cands);
        int sweepIterations = 0, sweepMaxIterations = 0;
        for (int tri : cands) {
            F3 v0, v1, v2;
            triVerts(set, tri, v0, v1, v2);
            int iterCount = 0;
            CapsuleCastHit hit;
            bool got = sweepCapsuleTriangle(from, dir, len, radius, halfHeight, v0, v1, v2, tri, iterCount, hit,
                                            ctx.stats);
            sweepIterations += iterCount;
            sweepMaxIterations = std::max(sweepMaxIterations, iterCount);
            if (!got) continue;
            bool better = hit.toi < bestT;
            bool equal = have && hit.toi == bestT;
            if (!better && !equal) continue;
            hit.material = set.materialForTriangle(tri);
            hit.triangleIndex = tri + offset;
            if (blockingOnly) { // CQ:1087-1094
                if (dot(delta, hit.normal) >= 0) continue;
                if (dot(delta, hit.triangleNormal) >= 0) continue;
            }
            if (minNormalY && hit.triangleNormal.y < *minNormalY) continue; // CQ:1095
            if (better) {
                bestT = hit.toi;
                bestHit = hit;
                have = true;
            } else {
                tieT = hit.toi; // an accepted candidate with exactly the same toi: visiting order decides
            }
        }
        if (ctx.stats) {
            ctx.stats->candidates += (int64_t)cands.size();
            ctx.stats->sweepTests += (int64_t)cands.size();
            ctx.stats->sweepIterations += sweepIterations;
            ctx.stats->maxIterations = std::max<int64_t>(ctx.stats->maxIterations, sweepMaxIterations);
        }
        if (have && tieT == bestT) ctx.tie = true;
        return have;
    }

    bool capsuleCastCombined(F3 from, F3 delta, float radius, float halfHeight, bool blockingOnly,
                             const float *minNormalY, uint32_t mask, QueryCtx &ctx, CapsuleCastHit &out) const {
        float len = length(delta);
        if (len < 1e-6f) return false; // CQ:988
        CapsuleCastHit a, b;
        bool ha = capsuleCastBVH(from, delta, radius, halfHeight, staticSet, 0, blockingOnly, minNormalY, mask, ctx, a);
        bool hb = capsuleCastBVH(from, delta, radius, halfHeight, dynamicSet, (int)staticSet.triangleAABBs.size(),
                                 blockingOnly, minNormalY, mask, ctx, b);
        if (ha && hb) { // chooseNearest CQ:909-914
            out = a.toi <= b.toi ? a : b;
            return true;
        }
        if (ha) out = a;
        else if (hb) out = b;
        return ha || hb;
    }
    bool capsuleCast(F3 from, F3 delta, float r, float hh, uint32_t mask, QueryCtx &ctx, CapsuleCastHit &out) const {
        return capsuleCastCombined(from, delta, r, hh, false, nullptr, mask, ctx, out); // CQ:787
    }
    bool capsuleCastBlocking(F3 from, F3 delta, float r, float hh, uint32_t mask, QueryCtx &ctx,
                             CapsuleCastHit &out) const {
        return capsuleCastCombined(from, delta, r, hh, true, nullptr, mask, ctx, out); // CQ:801
    }
    bool capsuleCastGround(F3 from, F3 delta, float r, float hh, float minNormalY, uint32_t mask, QueryCtx &ctx,
                           CapsuleCastHit &out) const {
        return capsuleCastCombined(from, delta, r, hh, false, &minNormalY, mask, ctx, out); // CQ:815
    }

    // ---- overlaps (CQ:1119-1283)
    static void overlapBox(F3 from, float radius, float halfHeight, F3 &minP, F3 &maxP) { // CQ:1126-1133
        F3 a0 = from + kUp * halfHeight, b0 = from - kUp * halfHeight;
        minP = vmin(a0, b0);
        maxP = vmax(a0, b0);
        F3 ext = {radius, radius, radius};
        minP = minP - ext;
        maxP = maxP + ext;
    }
    static bool overlapTriangle(const TriangleMeshSet &set, int tri, int offset, F3 from, float radius,
                                float halfHeight, QueryCtx &ctx, CapsuleOverlapHit &hit) { // CQ:1160-1190
        F3 v0, v1, v2;
        triVerts(set, tri, v0, v1, v2);
        SegTri d = segmentTriangleDistance(from, halfHeight, v0, v1, v2, ctx.stats);
        if (d.dist >= radius) return false;
        float depth = radius - d.dist;
        F3 triNormal = normalize(cross(v1 - v0, v2 - v0));
        F3 n = d.dist < 1e-6f ? triNormal : normalize(d.seg - d.tri);
        F3 triN = triNormal;
        if (dot(triN, n) < 0) triN = -triN;
        hit = {depth, d.tri, n, triN, tri + offset, set.materialForTriangle(tri)};
        return true;
    }

    bool capsuleOverlapBVH(F3 from, float radius, float halfHeight, const TriangleMeshSet &set, int offset,
                           uint32_t mask, QueryCtx &ctx, CapsuleOverlapHit &best) const { // CQ:1119-1199
        if (!set.hasBVH || set.bvh.root < 0) return false;
        F3 minP, maxP;
        overlapBox(from, radius, halfHeight, minP, maxP);
        thread_local std::vector<int> cands;
        collectBoxCandidates(set, minP, maxP, mask, ctx, cands);
        if (ctx.stats) ctx.stats->candidates += (int64_t)cands.size();
        bool have = false;
        float bestDepth = 0, tieDepth = -1;
        for (int tri : cands) {
            CapsuleOverlapHit hit;
            if (!overlapTriangle(set, tri, offset, from, radius, halfHeight, ctx, hit)) continue;
            if (hit.depth <= bestDepth) { // CQ:1172
                if (have && hit.depth == bestDepth) tieDepth = hit.depth;
                continue;
            }
            bestDepth = hit.depth;
            best = hit;
            have = true;
        }
        if (have && tieDepth == bestDepth) ctx.tie = true;
        return have;
    }
    bool capsuleOverlap(F3 from, float radius, float halfHeight, uint32_t mask, QueryCtx &ctx,
                        CapsuleOverlapHit &out) const { // CQ:830-850
        CapsuleOverlapHit a, b;
        bool ha = capsuleOverlapBVH(from, radius, halfHeight, staticSet, 0, mask, ctx, a);
        bool hb = capsuleOverlapBVH(from, radius, halfHeight, dynamicSet, (int)staticSet.triangleAABBs.size(), mask,
                                    ctx, b);
        if (ha && hb) {
            out = a.depth >= b.depth ? a : b;
            return true;
        }
        if (ha) out = a;
        else if (hb) out = b;
        return ha || hb;
    }

    // all overlapping triangles of one set in visiting order (no cap)
    void overlapAllOfSet(F3 from, float radius, float halfHeight, const TriangleMeshSet &set, int offset,
                         uint32_t mask, QueryCtx &ctx, std::vector<CapsuleOverlapHit> &hits) const {
        if (!set.hasBVH || set.bvh.root < 0) return;
        F3 minP, maxP;
        overlapBox(from, radius, halfHeight, minP, maxP);
        thread_local std::vector<int> cands;
        collectBoxCandidates(set, minP, maxP, mask, ctx, cands);
        if (ctx.stats) ctx.stats->candidates += (int64_t)cands.size();
        for (int tri : cands) {
            CapsuleOverlapHit hit;
            if (overlapTriangle(set, tri, offset, from, radius, halfHeight, ctx, hit)) hits.push_back(hit);
        }
    }

    // CQ:852-882 + :1201-1283.  REFERENCE: first maxHits overlaps in DFS order (static set, then
    // dynamic with the remainder).  CANONICAL: the maxHits deepest, ties by ascending index.
    // (The oracle evaluates every overlapping candidate in both modes so that `overflow` is known; the
    //  reference stops at maxHits — identical result, fewer distance evaluations when it overflows.)
    void capsuleOverlapAll(F3 from, float radius, float halfHeight, int maxHits, uint32_t mask, QueryCtx &ctx,
                           std::vector<CapsuleOverlapHit> &hits) const {
        hits.clear();
        thread_local std::vector<CapsuleOverlapHit> s, d;
        s.clear();
        d.clear();
        overlapAllOfSet(from, radius, halfHeight, staticSet, 0, mask, ctx, s);
        overlapAllOfSet(from, radius, halfHeight, dynamicSet, (int)staticSet.triangleAABBs.size(), mask, ctx, d);
        size_t total = s.size() + d.size();
        if ((int)total > maxHits) ctx.overflow = true;
        if (ctx.order == ORDER_CANONICAL) {
            hits = s;
            hits.insert(hits.end(), d.begin(), d.end());
            std::sort(hits.begin(), hits.end(), [](const CapsuleOverlapHit &a, const CapsuleOverlapHit &b) {
                if (a.depth != b.depth) return a.depth > b.depth;
                return a.triangleIndex < b.triangleIndex;
            });
            if ((int)hits.size() > maxHits) hits.resize(maxHits);
        } else {
            for (size_t i = 0; i < s.size() && (int)hits.size() < maxHits; i++) hits.push_back(s[i]);
            for (size_t i = 0; i < d.size() && (int)hits.size() < maxHits; i++) hits.push_back(d[i]);
        }
        for (size_t i = 0; i < hits.size(); i++)
            for (size_t j = i + 1; j < hits.size(); j++)
                if (hits[i].depth == hits[j].depth) ctx.tie = true; // caller's sort-by-depth order is ambiguous
    }
};

// ---------------------------------------------------------------- move-and-slide (SYS)
struct Character { // PhysicsBodyComponent + CharacterControllerComponent working copy
    D3 velocity;
    orc_state *s;
    const orc_params *p;
};

inline F3 ld3(const float *v) { return {v[0], v[1], v[2]}; }
inline void st3(float *o, F3 v) {
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}

// ContactManifoldCache (SYS:1159-1205)
void manifoldReset(orc_state &c) {
    c.manifold_count = 0;
    c.manifold_frames = 0;
}
bool manifoldNormalFor(const orc_state &c, int tri, F3 &out) { // SYS:1169-1175
    for (int i = 0; i < c.manifold_count; i++)
        if (c.manifold_triangles[i] == tri) {
            out = ld3(c.manifold_normals[i]);
            return true;
        }
    return false;
}
void manifoldUpdate(orc_state &c, int tri, F3 normal) { // SYS:1177-1204
    F3 n = normal;
    if (length_squared(n) < 1e-8f) return;
    c.manifold_frames = 8;
    for (int i = 0; i < c.manifold_count; i++) {
        if (c.manifold_triangles[i] == tri) {
            F3 cached = ld3(c.manifold_normals[i]);
            if (dot(cached, n) < 0) n = -n;
            const float blend = 0.25f;
            F3 combined = normalize(cached * (1 - blend) + n * blend);
            st3(c.manifold_normals[i], combined);
            st3(c.side_contact_normal, combined);
            return;
        }
    }
    if (c.manifold_count >= 4) c.manifold_count -= 1; // removeLast
    for (int i = c.manifold_count; i > 0; i--) {       // insert at 0
        c.manifold_triangles[i] = c.manifold_triangles[i - 1];
        memcpy(c.manifold_normals[i], c.manifold_normals[i - 1], sizeof(float) * 3);
    }
    c.manifold_triangles[0] = tri;
    st3(c.manifold_normals[0], normalize(n));
    c.manifold_count += 1;
    memcpy(c.side_contact_normal, c.manifold_normals[0], sizeof(float) * 3);
}
// DefaultContactCachePolicy (SYS:1102-1134)
void cacheDecay(orc_state &c) {
    if (c.side_contact_frames > 0) c.side_contact_frames -= 1;
    if (c.manifold_frames > 0) {
        c.manifold_frames -= 1;
        if (c.manifold_frames == 0) {
            manifoldReset(c);
            st3(c.side_contact_normal, f3(0, 0, 0));
        }
    }
}
void cacheRecord(orc_state &c, int tri, F3 normal, bool isSideContact) {
    manifoldUpdate(c, tri, normal);
    if (isSideContact) {
        st3(c.side_contact_normal, normalize(normal));
        c.side_contact_frames = 3;
    }
}

// DepenetrationResolver.resolve (SYS:734-808)
bool depenetrate(const StaticTriMesh &q, QueryCtx &ctx, F3 &position, Character &ch, float radius,
                 float halfHeight, float skinWidth, F3 &outNormal) {
    orc_state &c = *ch.s;
    const orc_params &p = *ch.p;
    float slop = smax(skinWidth * 0.5f, 0.001f);
    bool didResolve = false;
    F3 normalSum = {0, 0, 0};
    float normalWeight = 0;
    std::vector<CapsuleOverlapHit> hits;
    for (int it = 0; it < 4; it++) {
        q.capsuleOverlapAll(position, radius, halfHeight, 8, p.collision_mask, ctx, hits);
        if (hits.empty()) break;
        std::vector<CapsuleOverlapHit> sorted = hits;
        std::stable_sort(sorted.begin(), sorted.end(),
                         [](const CapsuleOverlapHit &a, const CapsuleOverlapHit &b) { return a.depth > b.depth; });
        const CapsuleOverlapHit &deepest = sorted[0];
        bool sideContact = deepest.normal.y < p.min_ground_dot;
        int useCount = sideContact ? 1 : smin(2, (int)sorted.size());
        float maxDepth = deepest.depth;
        F3 frameNormal = {0, 0, 0};
        for (int k = 0; k < useCount; k++) {
            const CapsuleOverlapHit &hit = sorted[k];
            maxDepth = smax(maxDepth, hit.depth);
            F3 n = hit.normal;
            F3 cached;
            if (manifoldNormalFor(c, hit.triangleIndex, cached)) {
                if (dot(cached, n) < 0) n = -n;
                n = cached;
            }
            frameNormal = frameNormal + n * hit.depth;
            cacheRecord(c, hit.triangleIndex, n, hit.normal.y < p.min_ground_dot);
        }
        float frameNormalLen = length(frameNormal);
        F3 depenNormal = frameNormalLen > 1e-6f ? frameNormal / frameNormalLen : frameNormal;
        float push = sideContact ? smax(maxDepth, 0.0f) : smax(maxDepth + slop, 0.0f);
        if (sideContact) push = smin(push, skinWidth);
        if (push <= 1e-6f) break;
        position = position + depenNormal * push;
        D3 depenNormalD = d3(depenNormal);
        double vInto = dot(ch.velocity, depenNormalD);
        if (vInto < 0) ch.velocity = ch.velocity - depenNormalD * vInto;
        didResolve = true;
        normalSum = normalSum + depenNormal * maxDepth;
        normalWeight += maxDepth;
    }
    if (!didResolve) return false;
    if (normalWeight > 1e-6f) outNormal = normalize(normalSum / normalWeight);
    else outNormal = normalize(normalSum);
    return true;
}

// ---- capsule-capsule CCD between agents (SYS:1417-1590, 1023-1091)
struct AgentState { // AgentSweepState (SYS:1023-1029): snapshot taken before any character of the step moves
    F3 position, velocity;
    float radius, halfHeight;
};
struct AgentSnapshot {
    const AgentState *agents;
    int count;
};
struct CapsuleCapsuleHit {
    float toi;
    F3 normal;
    int other;
};

bool clampInterval(float start, float end, float &s, float &e) { // SYS:1417-1424
    s = smax(start, 0.0f);
    e = smin(end, 1.0f);
    return !(e < s);
}
bool intervalGreaterEqual(float y0, float vy, float threshold, float &s, float &e) { // SYS:1426-1436
    if (fabsf(vy) < 1e-6f) {
        s = 0, e = 1;
        return y0 >= threshold;
    }
    float t = (threshold - y0) / vy;
    if (vy > 0) return clampInterval(t, 1, s, e);
    return clampInterval(0, t, s, e);
}
bool intervalLessEqual(float y0, float vy, float threshold, float &s, float &e) { // SYS:1438-1448
    if (fabsf(vy) < 1e-6f) {
        s = 0, e = 1;
        return y0 <= threshold;
    }
    float t = (threshold - y0) / vy;
    if (vy > 0) return clampInterval(0, t, s, e);
    return clampInterval(t, 1, s, e);
}
bool earliestRoot(float A, float B, float C, float tMin, float tMax, float &out) { // SYS:1450-1472
    const float eps = 1e-6f;
    if (fabsf(A) < eps) {
        if (fabsf(B) < eps) {
            out = tMin;
            return C <= 0;
        }
        float t = -C / B;
        out = t;
        return t >= tMin && t <= tMax;
    }
    float disc = B * B - 4 * A * C;
    if (disc < 0) return false;
    float sqrtD = sqrtf(disc);
    float inv2A = 1 / (2 * A);
    float t0 = (-B - sqrtD) * inv2A, t1 = (-B + sqrtD) * inv2A;
    float enter = smin(t0, t1), exit_ = smax(t0, t1);
    float s = smax(enter, tMin), e = smin(exit_, tMax);
    out = s;
    return e >= s;
}
float capsuleCapsuleSeparationY(float yRel, float halfHeightSum) { // SYS:1474-1482
    if (yRel > halfHeightSum) return yRel - halfHeightSum;
    if (yRel < -halfHeightSum) return yRel + halfHeightSum;
    return 0;
}
F3 capsuleCapsuleHitNormal(F3 rel, float halfHeightSum) { // SYS:1484-1497
    float sepY = capsuleCapsuleSeparationY(rel.y, halfHeightSum);
    F3 sep = {rel.x, sepY, rel.z};
    float lenSq = length_squared(sep);
    if (lenSq > 1e-8f) return sep / sqrtf(lenSq);
    F3 lateral = {rel.x, 0, rel.z};
    float lateralLenSq = length_squared(lateral);
    if (lateralLenSq > 1e-8f) return lateral / sqrtf(lateralLenSq);
    return {1, 0, 0};
}
bool capsuleCapsuleOverlap(F3 rel, float radiusSum, float halfHeightSum) { // SYS:1499-1503
    float sepY = capsuleCapsuleSeparationY(rel.y, halfHeightSum);
    float distSq = rel.x * rel.x + rel.z * rel.z + sepY * sepY;
    return distSq <= radiusSum * radiusSum;
}
bool capsuleCapsuleSweep(F3 from, F3 delta, float radius, float halfHeight, int other, F3 otherPos, F3 otherDelta,
                         float otherRadius, float otherHalfHeight, CapsuleCapsuleHit &out) { // SYS:1505-1590
    F3 relStart = from - otherPos, relDelta = delta - otherDelta;
    float rSum = radius + otherRadius, hSum = halfHeight + otherHalfHeight;
    float relLen = length(relDelta), moveLen = length(delta);
    if (relLen < 1e-6f) {
        if (capsuleCapsuleOverlap(relStart, rSum, hSum)) {
            out = {0, capsuleCapsuleHitNormal(relStart, hSum), other};
            return true;
        }
        return false;
    }
    float y0 = relStart.y, vy = relDelta.y, vx = relDelta.x, vz = relDelta.z, r0x = relStart.x, r0z = relStart.z;
    bool have = false;
    float bestT = 0, s, e, t;
    if (intervalGreaterEqual(y0, vy, hSum, s, e)) {
        float A = vx * vx + vz * vz + vy * vy;
        float B = 2 * (r0x * vx + r0z * vz + (y0 - hSum) * vy);
        float C = r0x * r0x + r0z * r0z + (y0 - hSum) * (y0 - hSum) - rSum * rSum;
        if (earliestRoot(A, B, C, s, e, t)) {
            bestT = t;
            have = true;
        }
    }
    if (intervalLessEqual(y0, vy, -hSum, s, e)) {
        float A = vx * vx + vz * vz + vy * vy;
        float B = 2 * (r0x * vx + r0z * vz + (y0 + hSum) * vy);
        float C = r0x * r0x + r0z * r0z + (y0 + hSum) * (y0 + hSum) - rSum * rSum;
        if (earliestRoot(A, B, C, s, e, t)) {
            if (!have || t < bestT) {
                bestT = t;
                have = true;
            }
        }
    }
    if (fabsf(vy) < 1e-6f) {
        if (fabsf(y0) <= hSum) {
            float A = vx * vx + vz * vz, B = 2 * (r0x * vx + r0z * vz), C = r0x * r0x + r0z * r0z - rSum * rSum;
            if (earliestRoot(A, B, C, 0, 1, t)) {
                if (!have || t < bestT) {
                    bestT = t;
                    have = true;
                }
            }
        }
    } else {
        float t1 = (hSum - y0) / vy, t2 = (-hSum - y0) / vy;
        if (clampInterval(smin(t1, t2), smax(t1, t2), s, e)) {
            float A = vx * vx + vz * vz, B = 2 * (r0x * vx + r0z * vz), C = r0x * r0x + r0z * r0z - rSum * rSum;
            if (earliestRoot(A, B, C, s, e, t)) {
                if (!have || t < bestT) {
                    bestT = t;
                    have = true;
                }
            }
        }
    }
    if (!have) return false;
    F3 relAtHit = relStart + relDelta * bestT;
    out = {bestT * moveLen, capsuleCapsuleHitNormal(relAtHit, hSum), other};
    return true;
}
// AgentSweepSolver.bestHit (SYS:1053-1091): every other agent, strict '<' keeps the first of equal tois
bool agentBestHit(F3 position, F3 remaining, float remainingLen, float baseMoveLen, float dt, int selfIndex,
                  float selfRadius, float halfHeight, const AgentSnapshot &snap, CapsuleCapsuleHit &best) {
    bool have = false;
    float timeScale = baseMoveLen > 1e-6f ? smin(remainingLen / baseMoveLen, 1.0f) : 1.0f;
    float segmentDt = dt * timeScale;
    for (int k = 0; k < snap.count; k++) {
        if (k == selfIndex) continue;
        const AgentState &o = snap.agents[k];
        F3 otherDelta = o.velocity * segmentDt;
        CapsuleCapsuleHit hit;
        if (capsuleCapsuleSweep(position, remaining, selfRadius, halfHeight, k, o.position, otherDelta, o.radius,
                                o.halfHeight, hit)) {
            if (!have || hit.toi < best.toi) {
                best = hit;
                have = true;
            }
        }
    }
    return have;
}

struct SlideOptions { // SYS:1208-1222
    bool allowHorizontalGroundPass, adjustVelocity, useGroundSnapSkinForStatic, allowTriangleNormalGroundLike;
};
const SlideOptions kSlideKinematicMove = {false, true, true, true};
const SlideOptions kSlideAgentSeparation = {true, false, false, false};

// SlideResolver.resolveHit (SYS:1229-1375); `agent` selects the .agentHit case
bool slideResolveHit(F3 &remaining, float len, const CapsuleCastHit &sHit, Character &ch, bool wasGrounded,
                     bool wasGroundedNear, F3 &position, bool haveCachedSide, F3 cachedSideNormal,
                     const CapsuleCapsuleHit *agent = nullptr, const SlideOptions &opt = kSlideKinematicMove) {
    const orc_params &p = *ch.p;
    const orc_state &c = *ch.s;
    const bool hitIsStatic = agent == nullptr;
    if (opt.allowHorizontalGroundPass && hitIsStatic && fabsf(remaining.y) < 1e-5f &&
        sHit.normal.y >= p.min_ground_dot) { // SYS:1240-1247
        position = position + remaining;
        remaining = {0, 0, 0};
        return true;
    }
    float hitToi = hitIsStatic ? sHit.toi : agent->toi;
    F3 slideNormal = hitIsStatic ? sHit.normal : agent->normal;
    bool hitIsGroundLike = hitIsStatic && sHit.triangleNormal.y >= p.min_ground_dot;
    float contactSkin = !hitIsStatic ? 0.0f
                                     : ((opt.useGroundSnapSkinForStatic && hitIsGroundLike) ? p.ground_snap_skin
                                                                                            : p.skin_width); // SYS:1255-1271
    F3 hitTriNormal = hitIsStatic ? sHit.triangleNormal : f3(0, 0, 0);

    if (hitIsStatic && slideNormal.y < p.min_ground_dot && c.side_contact_frames > 0) { // SYS:1273-1292
        if (haveCachedSide) {
            F3 cachedN = cachedSideNormal;
            if (dot(cachedN, slideNormal) < 0) cachedN = -cachedN;
            slideNormal = cachedN;
        } else {
            F3 cached = ld3(c.side_contact_normal);
            float cachedLen = length_squared(cached);
            if (cachedLen > 1e-6f) {
                F3 cachedN = cached / sqrtf(cachedLen);
                float dotC = dot(cachedN, slideNormal);
                if (fabsf(dotC) > 0.5f) slideNormal = dotC >= 0 ? cachedN : -cachedN;
            }
        }
    }
    if (slideNormal.y < p.min_ground_dot) { // SYS:1294-1309
        if (hitIsStatic && hitIsGroundLike && opt.allowTriangleNormalGroundLike) slideNormal = hitTriNormal;
        if (slideNormal.y < p.min_ground_dot) {
            slideNormal.y = 0;
            float nLen = length(slideNormal);
            if (nLen > 1e-5f) {
                slideNormal = slideNormal / nLen;
            } else {
                position = position + remaining;
                remaining = {0, 0, 0};
                return true;
            }
        }
    }
    float into = dot(remaining, slideNormal);
    float intoEps = 1e-4f * len;
    float effectiveSkin;
    if (hitToi <= contactSkin && into < -intoEps) effectiveSkin = smin(contactSkin, hitToi * 0.5f);
    else effectiveSkin = contactSkin;
    float stickyThreshold = contactSkin * 0.1f;
    if (hitToi <= stickyThreshold && into < -intoEps) { // SYS:1320
        remaining = remaining - slideNormal * into;
        return false;
    }
    if (into >= -intoEps) { // SYS:1324
        if (wasGroundedNear && hitIsStatic && !hitIsGroundLike && remaining.y < 0) remaining.y = 0;
        position = position + remaining;
        remaining = {0, 0, 0};
        return true;
    }
    if (hitToi <= effectiveSkin && fabsf(into) <= intoEps) {
        position = position + remaining;
        remaining = {0, 0, 0};
        return true;
    }
    if (into >= 0) {
        position = position + remaining;
        remaining = {0, 0, 0};
        return true;
    }
    float rawMoveDist = smax(hitToi - effectiveSkin, 0.0f); // SYS:1343
    float moveDist = rawMoveDist;
    if (slideNormal.y >= p.min_ground_dot && remaining.y < 0 && moveDist > p.ground_sweep_max_step)
        moveDist = p.ground_sweep_max_step;
    F3 dir = remaining / len;
    position = position + dir * moveDist;
    F3 leftover = remaining - dir * moveDist;
    leftover = leftover - slideNormal * dot(leftover, slideNormal);
    if (wasGrounded && wasGroundedNear && leftover.y < 0) leftover.y = 0;
    float residual = dot(leftover, slideNormal);
    if (fabsf(residual) < 1e-5f) leftover = leftover - slideNormal * residual;
    if (length_squared(leftover) < 1e-8f) {
        remaining = {0, 0, 0};
        return true;
    }
    remaining = leftover;
    if (opt.adjustVelocity) { // SYS:1366-1372
        double vInto = dot(ch.velocity, d3(slideNormal));
        if (vInto < 0) ch.velocity = ch.velocity - d3(slideNormal) * vInto;
    }
    return false;
}

// KinematicMoveStopSystem.resolveKinematicSweep without agents (SYS:1658-1765)
void resolveKinematicSweep(const StaticTriMesh &q, QueryCtx &ctx, F3 &position, F3 &remaining, Character &ch,
                           bool wasGrounded, bool wasGroundedNear, const AgentSnapshot *snap = nullptr,
                           int selfIndex = -1, float dt = 0) {
    orc_state &c = *ch.s;
    const orc_params &p = *ch.p;
    F3 baseMove = f3(ch.velocity) * dt; // body.linearVelocityF * dt (SYS:1671-1672)
    float baseMoveLen = length(baseMove);
    bool haveLast = false;
    F3 lastSlideNormal = {0, 0, 0};
    for (int it = 0; it < p.max_slide_iterations; it++) {
        float len = length(remaining);
        if (len < 1e-6f) break;
        CapsuleCastHit sHit;
        bool haveHit = q.capsuleCastBlocking(position, remaining, p.radius, p.half_height, p.collision_mask, ctx, sHit);
        if (haveHit && sHit.normal.y < p.min_ground_dot && c.side_contact_frames > 0) { // SYS:1683-1694
            F3 cached;
            if (manifoldNormalFor(c, sHit.triangleIndex, cached)) {
                F3 cachedN = cached;
                if (dot(cachedN, sHit.normal) < 0) cachedN = -cachedN;
                sHit.normal = cachedN;
            }
        }
        CapsuleCapsuleHit aHit;
        bool haveAgent = snap && agentBestHit(position, remaining, len, baseMoveLen, dt, selfIndex, p.radius, p.half_height,
                                              *snap, aHit); // SYS:1695-1705
        bool useAgent = false;
        if (haveHit && haveAgent) { // HitSelector.selectBestHit (SYS:1378-1399)
            float staticSkin = sHit.normal.y >= p.min_ground_dot ? p.ground_snap_skin : p.skin_width;
            float staticStop = smax(sHit.toi - staticSkin, 0.0f), agentStop = smax(aHit.toi, 0.0f);
            useAgent = !(staticStop <= agentStop);
        } else if (haveAgent) {
            useAgent = true;
        }
        if (useAgent) {
            F3 hitNormal = aHit.normal;
            bool shouldBreak = slideResolveHit(remaining, len, sHit, ch, wasGrounded, wasGroundedNear, position, false,
                                               f3(0, 0, 0), &aHit);
            if (haveLast) { // SYS:1744-1754
                float dotN = dot(lastSlideNormal, hitNormal);
                if (fabsf(dotN) < 0.98f) {
                    F3 axis = cross(lastSlideNormal, hitNormal);
                    float axisLen = length(axis);
                    if (axisLen > 1e-5f) {
                        F3 axisN = axis / axisLen;
                        remaining = axisN * dot(remaining, axisN);
                    }
                }
            }
            lastSlideNormal = hitNormal;
            haveLast = true;
            if (shouldBreak) break;
        } else if (haveHit) {
            F3 hitNormal = sHit.normal;
            bool haveCachedSide = false;
            F3 cachedSide = {0, 0, 0};
            if (sHit.normal.y < p.min_ground_dot && c.side_contact_frames > 0) // SYS:1719-1726
                haveCachedSide = manifoldNormalFor(c, sHit.triangleIndex, cachedSide);
            bool shouldBreak = slideResolveHit(remaining, len, sHit, ch, wasGrounded, wasGroundedNear, position,
                                               haveCachedSide, cachedSide);
            if (sHit.normal.y < p.min_ground_dot) cacheRecord(c, sHit.triangleIndex, sHit.normal, true); // SYS:1738
            if (haveLast) { // SYS:1744-1754
                float dotN = dot(lastSlideNormal, hitNormal);
                if (fabsf(dotN) < 0.98f) {
                    F3 axis = cross(lastSlideNormal, hitNormal);
                    float axisLen = length(axis);
                    if (axisLen > 1e-5f) {
                        F3 axisN = axis / axisLen;
                        remaining = axisN * dot(remaining, axisN);
                    }
                }
            }
            lastSlideNormal = hitNormal;
            haveLast = true;
            if (shouldBreak) break;
        } else {
            position = position + remaining;
            remaining = {0, 0, 0};
            break;
        }
    }
}

// ---- AgentSeparationSystem (SYS:1906-2210): sequential pair resolution in entity order over a uniform XZ grid,
// then a per-agent slide from the pre-separation position and a snap to the ground
struct SepAgent { // AgentSeparationSystem.Agent (SYS:1907-1915); the controller copy is the character's orc_state
    F3 position, velocity;
    float radius, halfHeight, invWeight;
};
struct SepCell {
    int64_t x, z;
    bool operator==(const SepCell &o) const { return x == o.x && z == o.z; }
};
struct SepCellHash {
    size_t operator()(const SepCell &c) const {
        return std::hash<int64_t>()(c.x * 0x9E3779B97F4A7C15ll ^ (c.z + 0x7F4A7C15ll) * 0xC2B2AE3D27D4EB4Fll);
    }
};
struct SepGrid { // AgentSeparationGrid (SYS:1917-1944)
    float cellSize;
    std::unordered_map<SepCell, std::vector<int>, SepCellHash> cells;
    SepCell cellCoord(F3 pos) const { return {(int64_t)floorf(pos.x / cellSize), (int64_t)floorf(pos.z / cellSize)}; }
    void rebuild(const std::vector<SepAgent> &agents) {
        cells.clear();
        for (int i = 0; i < (int)agents.size(); i++) cells[cellCoord(agents[i].position)].push_back(i);
    }
};

// AgentSeparationResolver.resolve (SYS:1946-2041)
void agentSeparationResolve(std::vector<SepAgent> &agents, const SepGrid &grid, const orc_params &p, float separationMargin,
                            float heightMargin, const StaticTriMesh *query, QueryCtx &ctx, int64_t *pairCount) {
    for (int i = 0; i < (int)agents.size(); i++) {
        const SepAgent a = agents[i]; // `let a = agents[i]`: a copy taken at the start of i's turn
        SepCell cell = grid.cellCoord(a.position);
        for (int dz = -1; dz <= 1; dz++)
            for (int dx = -1; dx <= 1; dx++) {
                auto it = grid.cells.find(SepCell{cell.x + dx, cell.z + dz});
                if (it == grid.cells.end()) continue;
                for (int j : it->second) {
                    if (!(j > i)) continue;
                    const SepAgent b = agents[j];
                    float aMin = a.position.y - a.halfHeight, aMax = a.position.y + a.halfHeight;
                    float bMin = b.position.y - b.halfHeight, bMax = b.position.y + b.halfHeight;
                    float ddx = a.position.x - b.position.x, ddz = a.position.z - b.position.z;
                    float distSq = ddx * ddx + ddz * ddz;
                    float skinAllowance = smin(p.skin_width, p.skin_width);
                    float margin = smin(separationMargin, skinAllowance);
                    float minDist = a.radius + b.radius + margin;
                    bool heightSeparated = aMax < bMin - heightMargin || aMin > bMax + heightMargin;
                    if (heightSeparated) continue;
                    if (distSq >= minDist * minDist) continue;
                    float dist = sqrtf(smax(distSq, 1e-8f));
                    float nx = ddx / dist, nz = ddz / dist;
                    float penetration = minDist - dist;
                    float wSum = a.invWeight + b.invWeight;
                    if (wSum <= 0) continue;
                    if (pairCount) (*pairCount)++;
                    float corr = penetration / wSum;
                    F3 moveA = {nx * corr * a.invWeight, 0, nz * corr * a.invWeight};
                    F3 moveB = {-nx * corr * b.invWeight, 0, -nz * corr * b.invWeight};
                    F3 relV = a.velocity - b.velocity;
                    float vn = relV.x * nx + relV.z * nz;
                    if (vn < 0) { // SYS:1985-1993
                        float impulse = -vn;
                        float scaleA = a.invWeight / wSum, scaleB = b.invWeight / wSum;
                        agents[i].velocity.x += nx * impulse * scaleA;
                        agents[i].velocity.z += nz * impulse * scaleA;
                        agents[j].velocity.x -= nx * impulse * scaleB;
                        agents[j].velocity.z -= nz * impulse * scaleB;
                    }
                    if (query) { // SYS:1994-2033
                        const float eps = 1e-6f;
                        bool blockedA = false, blockedB = false;
                        CapsuleCastHit hit;
                        if (length(moveA) > eps &&
                            query->capsuleCastBlocking(agents[i].position, moveA, a.radius, a.halfHeight, p.collision_mask, ctx,
                                                       hit) &&
                            hit.toi <= p.skin_width && hit.normal.y < p.min_ground_dot)
                            blockedA = true;
                        if (length(moveB) > eps &&
                            query->capsuleCastBlocking(agents[j].position, moveB, b.radius, b.halfHeight, p.collision_mask, ctx,
                                                       hit) &&
                            hit.toi <= p.skin_width && hit.normal.y < p.min_ground_dot)
                            blockedB = true;
                        if (blockedA && !blockedB) {
                            moveA = {0, 0, 0};
                            moveB = {-nx * penetration, 0, -nz * penetration};
                        } else if (blockedB && !blockedA) {
                            moveB = {0, 0, 0};
                            moveA = {nx * penetration, 0, nz * penetration};
                        } else if (blockedA && blockedB) {
                            continue;
                        }
                    }
                    agents[i].position = agents[i].position + moveA;
                    agents[j].position = agents[j].position + moveB;
                }
            }
    }
}

// AgentSeparationPostProcessor.apply (SYS:2043-2117) + the write-back of fixedUpdate (SYS:2196-2208)
void agentSeparationPost(const StaticTriMesh *query, QueryCtx &ctx, const SepAgent &agent, F3 start, orc_state &s,
                         const orc_params &p) {
    F3 position = agent.position;
    if (query) {
        F3 delta = position - start;
        float len = length(delta);
        bool moved = false;
        if (len > 1e-6f) {
            moved = true;
            F3 remaining = delta;
            position = start;
            Character ch;
            ch.s = &s;
            ch.p = &p;
            ch.velocity = {s.velocity[0], s.velocity[1], s.velocity[2]};
            for (int it = 0; it < 2; it++) {
                float segLen = length(remaining);
                if (segLen < 1e-6f) break;
                CapsuleCastHit hit;
                if (query->capsuleCastBlocking(position, remaining, agent.radius, agent.halfHeight, p.collision_mask, ctx, hit)) {
                    bool done = slideResolveHit(remaining, segLen, hit, ch, false, false, position, false, f3(0, 0, 0), nullptr,
                                                kSlideAgentSeparation);
                    if (done) break;
                } else {
                    position = position + remaining;
                    remaining = {0, 0, 0};
                    break;
                }
            }
        }
        if (moved && s.velocity[1] <= 0) { // body.linearVelocity.y (Double, still the pre-separation value)
            if (p.snap_distance > 0) {
                F3 down = {0, -1, 0};
                F3 snapDelta = down * p.snap_distance;
                CapsuleCastHit hit;
                if (query->capsuleCastGround(position, snapDelta, agent.radius, agent.halfHeight, p.min_ground_dot,
                                             p.collision_mask, ctx, hit) &&
                    hit.toi <= p.snap_distance) {
                    float rawMove = smax(hit.toi - p.ground_snap_skin, 0.0f);
                    float moveDist = smin(rawMove, p.ground_snap_max_step);
                    position = position + down * moveDist;
                    s.grounded = 1;
                    s.grounded_near = hit.toi <= smax(p.ground_snap_skin, p.skin_width) ? 1 : 0;
                    st3(s.ground_normal, hit.material.flatten ? f3(0, 1, 0) : hit.triangleNormal);
                    s.ground_triangle_index = hit.triangleIndex;
                }
            }
        }
    }
    s.position[0] = (double)position.x, s.position[1] = (double)position.y, s.position[2] = (double)position.z;
    s.velocity[0] = (double)agent.velocity.x, s.velocity[1] = (double)agent.velocity.y, s.velocity[2] = (double)agent.velocity.z;
}

struct GroundContactState { // SYS:810-817
    bool grounded, groundedNear;
    F3 normal;
    Material material;
    int triangleIndex;
    float distance;
};
struct GroundProbeResult { // SYS:819-824
    GroundContactState state;
    bool canSnap, nearGround, haveHit;
    CapsuleCastHit hit;
};

// GroundProbe.resolve (SYS:826-943)
GroundProbeResult groundProbe(const StaticTriMesh &q, QueryCtx &ctx, F3 position, const Character &ch,
                              bool wasGroundedNear, F3 prevNormal) {
    const orc_params &p = *ch.p;
    GroundProbeResult r;
    r.state = {false, false, f3(0, 1, 0), kDefaultMaterial, -1, FLT_MAX};
    r.canSnap = false;
    r.nearGround = false;
    r.haveHit = false;
    F3 down = {0, -1, 0};
    F3 snapDelta = down * p.snap_distance;
    CapsuleCastHit centerHit;
    bool haveCenter = false;
    if (p.snap_distance > 0)
        haveCenter = q.capsuleCastGround(position, snapDelta, p.radius, p.half_height, p.min_ground_dot,
                                         p.collision_mask, ctx, centerHit);
    if (p.fall_probe_distance > 0) { // SYS:855-866
        F3 fallDelta = down * p.fall_probe_distance;
        CapsuleCastHit fallHit;
        if (q.capsuleCastGround(position, fallDelta, p.radius, p.half_height, p.min_ground_dot, p.collision_mask, ctx,
                                fallHit))
            r.state.distance = fallHit.toi;
    }
    if (!haveCenter || !(centerHit.toi <= p.snap_distance)) return r; // SYS:868-871

    float baseCenterY = position.y - p.half_height;
    float bottomY = baseCenterY - p.radius;
    float groundTol = smax(p.skin_width, p.ground_snap_skin);
    bool validGroundPoint = centerHit.position.y <= bottomY + groundTol;
    float groundNearThreshold = smax(p.ground_snap_skin, p.skin_width);
    bool nearGround = centerHit.toi <= groundNearThreshold;
    r.state.groundedNear = nearGround;
    r.state.distance = centerHit.toi;
    bool groundGateVel = ch.velocity.y <= 0;
    double vInto = dot(ch.velocity, d3(centerHit.normal));
    bool groundGateSpeed = vInto >= -(double)p.ground_snap_max_speed;
    bool groundGateToi = centerHit.toi <= p.ground_snap_max_toi;
    bool canSnap = validGroundPoint && groundGateVel && (nearGround || groundGateSpeed || groundGateToi);
    if (wasGroundedNear && centerHit.toi <= p.snap_distance) canSnap = validGroundPoint;

    if (validGroundPoint && (nearGround || canSnap)) { // SYS:890-925
        r.state.grounded = true;
        r.state.material = centerHit.material;
        r.state.triangleIndex = centerHit.triangleIndex;
        F3 normalSum = centerHit.triangleNormal;
        const float flatDot = 0.98f;
        if (centerHit.triangleNormal.y < flatDot && (wasGroundedNear || nearGround)) {
            float offset = p.radius * 0.6f;
            const float offs[4][2] = {{offset, 0}, {-offset, 0}, {0, offset}, {0, -offset}};
            float combineTol = smax(smax(p.ground_snap_skin, p.skin_width), 0.05f);
            for (int k = 0; k < 4; k++) {
                F3 samplePos = position + f3(offs[k][0], 0, offs[k][1]);
                CapsuleCastHit hit;
                if (q.capsuleCastGround(samplePos, snapDelta, p.radius, p.half_height, p.min_ground_dot,
                                        p.collision_mask, ctx, hit) &&
                    hit.toi <= centerHit.toi + combineTol) {
                    if (dot(hit.triangleNormal, centerHit.triangleNormal) > 0.98f)
                        normalSum = normalSum + hit.triangleNormal;
                }
            }
        }
        float nLen = length(normalSum);
        r.state.normal = nLen > 1e-6f ? normalSum / nLen : centerHit.triangleNormal;
    }
    if (r.state.grounded && wasGroundedNear) { // SYS:927-934
        float dotN = dot(prevNormal, r.state.normal);
        if (dotN > 0.9f) {
            const float blend = 0.2f;
            r.state.normal = normalize(prevNormal * (1 - blend) + r.state.normal * blend);
        }
    }
    if (r.state.grounded && r.state.material.flatten) r.state.normal = {0, 1, 0};
    r.canSnap = canSnap;
    r.nearGround = nearGround;
    r.haveHit = true;
    r.hit = centerHit;
    return r;
}

// PlatformCarry.computeDelta (SYS:644-732); a platform = world AABB of its mesh + its motion this step
F3 platformCarryDelta(F3 position, const orc_params &c, const orc_platform *platforms, int nPlatforms) {
    if (nPlatforms <= 0) return {0, 0, 0};
    float capsuleHalf = c.half_height + c.radius;
    float baseY = position.y - capsuleHalf;
    F3 capMin = {position.x - c.radius, position.y - capsuleHalf, position.z - c.radius};
    F3 capMax = {position.x + c.radius, position.y + capsuleHalf, position.z + c.radius};
    float sideTol = smax(c.skin_width, c.ground_snap_skin);
    F3 bestCarry = {0, 0, 0}, pushDelta = {0, 0, 0};
    for (int k = 0; k < nPlatforms; k++) {
        F3 pDelta = ld3(platforms[k].delta);
        if (length_squared(pDelta) < 1e-8f) continue;
        F3 amin = ld3(platforms[k].aabb_min), amax = ld3(platforms[k].aabb_max);
        F3 emin = amin - f3(sideTol, sideTol, sideTol), emax = amax + f3(sideTol, sideTol, sideTol);
        bool overlap = capMin.x <= emax.x && capMax.x >= emin.x && capMin.y <= emax.y && capMax.y >= emin.y &&
                       capMin.z <= emax.z && capMax.z >= emin.z;
        if (!overlap) continue;
        bool withinXZ = position.x >= amin.x - c.radius && position.x <= amax.x + c.radius &&
                        position.z >= amin.z - c.radius && position.z <= amax.z + c.radius;
        float topY = amax.y;
        float topTol = c.snap_distance + smax(c.skin_width, c.ground_snap_skin) + 0.05f;
        bool onTop = withinXZ && baseY >= topY - topTol && baseY <= topY + topTol;
        if (onTop) {
            if (length_squared(pDelta) > length_squared(bestCarry)) bestCarry = pDelta;
        } else {
            float yMin = amin.y - capsuleHalf, yMax = amax.y + capsuleHalf;
            if (position.y >= yMin && position.y <= yMax) {
                bool outsideX = position.x < amin.x - c.radius || position.x > amax.x + c.radius;
                bool outsideZ = position.z < amin.z - c.radius || position.z > amax.z + c.radius;
                if (!outsideX && !outsideZ) continue;
                float cx = smax(amin.x, smin(position.x, amax.x));
                float cz = smax(amin.z, smin(position.z, amax.z));
                float dx = position.x - cx, dz = position.z - cz;
                float sideDistSq = dx * dx + dz * dz;
                float sidePushTol = c.radius + sideTol;
                if (sideDistSq <= sidePushTol * sidePushTol) {
                    float dirLen = sqrtf(smax(sideDistSq, 0.0f));
                    if (dirLen > 1e-5f) {
                        F3 dir = {dx / dirLen, 0, dz / dirLen};
                        float moveToward = dot(f3(pDelta.x, 0, pDelta.z), dir);
                        if (moveToward > 0) pushDelta = pushDelta + f3(pDelta.x, 0, pDelta.z);
                    }
                }
            }
        }
    }
    if (length_squared(bestCarry) > 1e-8f) return bestCarry;
    if (length_squared(pushDelta) > 1e-8f) return pushDelta;
    return {0, 0, 0};
}

// one character, one fixed step: KinematicMoveStopSystem.fixedUpdate loop body (SYS:1842-1901)
void moveAndSlideOne(const StaticTriMesh &q, QueryCtx &ctx, orc_state &s, const orc_params &p, float dt,
                     F3 gravity, uint32_t flags, const orc_platform *platforms = nullptr, int nPlatforms = 0,
                     const AgentSnapshot *snap = nullptr, int selfIndex = -1) {
    Character ch;
    ch.s = &s;
    ch.p = &p;
    ch.velocity = {s.velocity[0], s.velocity[1], s.velocity[2]};
    if (flags & 1u) { // GravitySystem.fixedUpdate (SYS:603-619), body assumed .dynamic
        if (!(s.grounded && s.grounded_near)) ch.velocity = ch.velocity + d3(gravity) * (double)dt;
    }
    F3 position = {(float)s.position[0], (float)s.position[1], (float)s.position[2]}; // positionF
    cacheDecay(s);                                                                   // SYS:1848
    { // applyPlatformDelta (SYS:1618-1633)
        F3 platformDelta = platformCarryDelta(position, p, platforms, nPlatforms);
        if (length_squared(platformDelta) > 1e-8f) position = position + platformDelta;
    }
    bool wasGrounded = s.grounded != 0, wasGroundedNear = s.grounded_near != 0;
    // VelocityGate.apply (SYS:1037-1051)
    if (wasGrounded && wasGroundedNear && ch.velocity.y < 0) ch.velocity.y = 0;
    D3 remD = ch.velocity * (double)dt;
    if (wasGrounded && wasGroundedNear && remD.y < 0) remD.y = 0;
    F3 remaining = f3(remD);
    // applyPreSweepDepenetration (SYS:1635-1656)
    F3 depenNormal;
    if (depenetrate(q, ctx, position, ch, p.radius, p.half_height, p.skin_width, depenNormal)) {
        float into = dot(remaining, depenNormal);
        if (into < 0) remaining = remaining - depenNormal * into;
    }
    resolveKinematicSweep(q, ctx, position, remaining, ch, wasGrounded, wasGroundedNear, snap, selfIndex, dt);
    // resolveGroundContact (SYS:1767-1800)
    GroundProbeResult probe = groundProbe(q, ctx, position, ch, wasGroundedNear, ld3(s.ground_normal));
    GroundContactState gs = probe.state;
    if (probe.canSnap && probe.haveHit) { // GroundSnap.apply (SYS:945-963)
        float rawMove = smax(probe.hit.toi - p.ground_snap_skin, 0.0f);
        float moveDist = rawMove;
        if (probe.nearGround && moveDist > p.ground_snap_max_step) moveDist = p.ground_snap_max_step;
        position = position + f3(0, -1, 0) * moveDist;
        double vIntoSnap = dot(ch.velocity, d3(probe.hit.normal));
        if (vIntoSnap < 0) ch.velocity = ch.velocity - d3(probe.hit.normal) * vIntoSnap;
    }
    if (gs.grounded) {
        float normalUpDelta = gs.normal.y - s.ground_normal[1];
        if (gs.triangleIndex != s.ground_triangle_index && normalUpDelta > 0.02f) s.ground_transition_frames = 3;
    }
    // SlopeFriction.apply (SYS:965-1021)
    if (!gs.grounded) {
        s.ground_sliding = 0;
    } else {
        F3 normal = normalize(gs.normal);
        if (normal.y > 0.98f) {
            s.ground_transition_frames = 0;
            s.ground_sliding = 0;
        } else if (s.ground_transition_frames > 0) {
            s.ground_transition_frames -= 1;
            s.ground_sliding = 0;
        } else {
            float gN = dot(gravity, normal);
            F3 gTan = gravity - normal * gN;
            float gTanLen = length(gTan);
            const float slopeAccelEps = 0.5f;
            if (gTanLen > slopeAccelEps) {
                float gNMag = fabsf(gN);
                F3 gTanDir = gTan / gTanLen;
                D3 gTanDirD = d3(gTanDir), normalD = d3(normal);
                float stickLimit = gs.material.muS * gNMag;
                bool enterSlide = gTanLen > stickLimit * 1.05f;
                bool exitSlide = gTanLen < stickLimit * 0.9f;
                if (s.ground_sliding) {
                    if (exitSlide) s.ground_sliding = 0;
                } else if (enterSlide) {
                    s.ground_sliding = 1;
                }
                if (!s.ground_sliding && gTanLen <= stickLimit) {
                    D3 v = ch.velocity;
                    D3 vTan = v - normalD * dot(v, normalD);
                    double downhillSpeed = dot(vTan, gTanDirD);
                    if (downhillSpeed > 0) ch.velocity = ch.velocity - gTanDirD * downhillSpeed;
                } else {
                    float slideAccelMag = smax(gTanLen - gs.material.muK * gNMag, 0.0f);
                    if (slideAccelMag > 0) ch.velocity = ch.velocity + gTanDirD * (double)slideAccelMag * (double)dt;
                }
            }
        }
    }
    // writeBack (SYS:1802-1821)
    s.position[0] = (double)position.x;
    s.position[1] = (double)position.y;
    s.position[2] = (double)position.z;
    s.velocity[0] = ch.velocity.x;
    s.velocity[1] = ch.velocity.y;
    s.velocity[2] = ch.velocity.z;
    s.grounded = gs.grounded ? 1 : 0;
    s.grounded_near = gs.groundedNear ? 1 : 0;
    st3(s.ground_normal, gs.grounded ? gs.normal : f3(0, 1, 0));
    s.ground_distance = gs.distance;
    if (gs.grounded) s.ground_triangle_index = gs.triangleIndex;
}

template <class Fn> void parallelFor(int n, int nThreads, Fn fn) {
    if (nThreads <= 1 || n < 2) {
        fn(0, 0, n);
        return;
    }
    nThreads = std::min(nThreads, n);
    std::vector<std::thread> threads;
    for (int t = 0; t < nThreads; t++) {
        int lo = (int)((int64_t)n * t / nThreads), hi = (int)((int64_t)n * (t + 1) / nThreads);
        threads.emplace_back([=] { fn(t, lo, hi); });
    }
    for (auto &th : threads) th.join();
}

void exportStats(const std::vector<Stats> &per, orc_stats *out) {
    if (!out) return;
    Stats s;
    for (const Stats &p : per) s.add(p);
    out->candidates = s.candidates;
    out->sweep_tests = s.sweepTests;
    out->sweep_iterations = s.sweepIterations;
    out->max_iterations = s.maxIterations;
    out->distance_evals = s.distanceEvals;
    out->nodes_visited = s.nodesVisited;
    out->ties = s.ties;
    out->overflows = s.overflows;
}

} // namespace

// ---------------------------------------------------------------- C interface
struct orc_world {
    std::vector<PartCopy> parts;
    StaticTriMesh mesh;
    std::vector<uint8_t> flagScratch;
};

extern "C" {

orc_world *orc_world_create(const orc_part *parts, int32_t n_parts) { return orc_world_create_ex(parts, n_parts, nullptr, 0); }

void orc_world_triangle_material(const orc_world *w, int32_t triangle_index, float out[3]) { // CQ:464-469 over both sets (:782,1004)
    const int nStatic = (int)w->mesh.staticSet.triangleAABBs.size();
    const Material m = triangle_index >= nStatic ? w->mesh.dynamicSet.materialForTriangle(triangle_index - nStatic)
                                                  : w->mesh.staticSet.materialForTriangle(triangle_index);
    out[0] = m.muS, out[1] = m.muK, out[2] = m.flatten ? 1.0f : 0.0f;
}

orc_world *orc_world_create_ex(const orc_part *parts, int32_t n_parts, const orc_triangle_materials *tri_materials,
                               int32_t n_tri_materials) { // StaticTriMesh.init CQ:717-726
    orc_world *w = new orc_world();
    w->parts.resize(n_parts);
    for (int i = 0; i < n_parts; i++) {
        PartCopy &pc = w->parts[i];
        pc.positions.assign(parts[i].positions_xyz, parts[i].positions_xyz + (size_t)parts[i].n_verts * 3);
        pc.indices.assign(parts[i].indices, parts[i].indices + parts[i].n_indices);
        memcpy(pc.model, parts[i].model, sizeof(pc.model));
        pc.layer = parts[i].layer;
        pc.material = {parts[i].mu_s, parts[i].mu_k, parts[i].flatten_ground != 0};
        pc.isDynamic = parts[i].is_dynamic != 0;
        pc.entityId = parts[i].entity_id;
        pc.partIndex = i;
        for (int k = 0; k < n_tri_materials; k++)
            if (tri_materials[k].entity_id == pc.entityId && tri_materials[k].n > 0 && tri_materials[k].materials) {
                pc.triangleMaterials.clear();
                for (int t = 0; t < tri_materials[k].n; t++) {
                    const orc_surface_material &m = tri_materials[k].materials[t];
                    pc.triangleMaterials.push_back({m.mu_s, m.mu_k, m.flatten_ground != 0});
                }
            }
    }
    std::vector<const PartCopy *> statics, dynamics; // partitionEntities CQ:886-900
    for (const PartCopy &pc : w->parts) (pc.isDynamic ? dynamics : statics).push_back(&pc);
    w->mesh.staticSet.rebuild(statics);
    w->mesh.dynamicSet.rebuild(dynamics);
    return w;
}

void orc_world_destroy(orc_world *w) { delete w; }

void orc_world_counts(const orc_world *w, int32_t which, int32_t out[3]) {
    const TriangleMeshSet &s = which ? w->mesh.dynamicSet : w->mesh.staticSet;
    out[0] = (int32_t)s.positions.size();
    out[1] = (int32_t)s.triangleAABBs.size();
    out[2] = s.hasBVH ? (int32_t)s.bvh.nodes.size() : 0;
}

void orc_world_read_soup(const orc_world *w, int32_t which, float *positions, uint32_t *indices, float *aabbs,
                         uint32_t *layers, int32_t *parts) {
    const TriangleMeshSet &s = which ? w->mesh.dynamicSet : w->mesh.staticSet;
    if (positions) memcpy(positions, s.positions.data(), s.positions.size() * sizeof(F3));
    if (indices) memcpy(indices, s.indices.data(), s.indices.size() * sizeof(uint32_t));
    if (aabbs) memcpy(aabbs, s.triangleAABBs.data(), s.triangleAABBs.size() * sizeof(AABB));
    if (layers) memcpy(layers, s.triangleLayers.data(), s.triangleLayers.size() * sizeof(uint32_t));
    if (parts) memcpy(parts, s.triangleParts.data(), s.triangleParts.size() * sizeof(int32_t));
}

void orc_world_update_transforms(orc_world *w, const uint32_t *entity_ids, const float *models, int32_t n) {
    std::vector<const PartCopy *> statics, dynamics; // CQ:750-766
    for (int i = 0; i < n; i++)
        for (PartCopy &pc : w->parts)
            if (pc.entityId == entity_ids[i]) {
                memcpy(pc.model, models + 16 * i, sizeof(pc.model));
                (pc.isDynamic ? dynamics : statics).push_back(&pc);
            }
    w->mesh.staticSet.updateTransforms(statics);
    w->mesh.dynamicSet.updateTransforms(dynamics);
}

int32_t orc_world_check_bvh(const orc_world *w, int32_t which) {
    const TriangleMeshSet &s = which ? w->mesh.dynamicSet : w->mesh.staticSet;
    if (!s.hasBVH) return 1;
    const BVH &b = s.bvh;
    auto same = [](const AABB &x, const AABB &y) { return memcmp(&x, &y, sizeof(AABB)) == 0; };
    for (const BVHNode &n : b.nodes) {
        if (n.left < 0) {
            if (!same(n.bounds, b.boundsForRange(s.triangleAABBs, n.start, n.count))) return 0;
        } else if (!same(n.bounds, BVH::merge(b.nodes[n.left].bounds, b.nodes[n.right].bounds))) {
            return 0;
        }
    }
    return 1;
}

// AgentSeparationSystem.fixedUpdate (SYS:2136-2210): every character of the batch is a solid agent;
// mass_weight per agent (null = 1.0 each, AgentCollisionComponent default); use_query = setQuery(...) was given
void orc_agent_separation(orc_world *w, orc_state *inout, int32_t n, const orc_params *params, const float *mass_weight,
                          int32_t iterations, float separation_margin, float height_margin, int32_t use_query, int32_t order,
                          int32_t n_threads, int64_t *pair_count) {
    if (pair_count) *pair_count = 0;
    if (n <= 1) return; // `guard agents.count > 1`
    const orc_params &p = *params;
    std::vector<SepAgent> agents(n);
    std::vector<F3> originalPositions(n);
    float maxRadius = 0;
    for (int i = 0; i < n; i++) {
        float mw = mass_weight ? mass_weight[i] : 1.0f;
        float invWeight = mw > 0 ? 1.0f / mw : 0.0f;
        maxRadius = smax(maxRadius, p.radius);
        F3 pos = {(float)inout[i].position[0], (float)inout[i].position[1], (float)inout[i].position[2]};
        F3 vel = {(float)inout[i].velocity[0], (float)inout[i].velocity[1], (float)inout[i].velocity[2]};
        agents[i] = {pos, vel, p.radius, p.half_height, invWeight};
        originalPositions[i] = pos;
    }
    SepGrid grid;
    grid.cellSize = smax(maxRadius * 2 + separation_margin, 0.001f);
    QueryCtx ctx;
    ctx.order = order;
    const StaticTriMesh *query = use_query ? &w->mesh : nullptr;
    iterations = std::max(1, iterations);
    for (int it = 0; it < iterations; it++) {
        grid.rebuild(agents);
        agentSeparationResolve(agents, grid, p, separation_margin, height_margin, query, ctx, pair_count);
    }
    parallelFor(n, n_threads, [&](int, int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            QueryCtx c2;
            c2.order = order;
            agentSeparationPost(query, c2, agents[i], originalPositions[i], inout[i], p);
        }
    });
}

void orc_raycast(orc_world *w, const orc_ray *rays, int32_t n, orc_ray_hit *out, int32_t order, int32_t n_threads,
                 orc_stats *stats) {
    std::vector<Stats> per(std::max(1, n_threads));
    parallelFor(n, n_threads, [&](int t, int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            QueryCtx ctx;
            ctx.order = order;
            ctx.stats = stats ? &per[t] : nullptr;
            RaycastHit h;
            orc_ray_hit &o = out[i];
            if (w->mesh.raycast(ld3(rays[i].origin), ld3(rays[i].direction), rays[i].max_distance, rays[i].mask, ctx, h)) {
                o.distance = h.distance;
                st3(o.position, h.position);
                st3(o.normal, h.normal);
                o.triangle_index = h.triangleIndex;
            } else {
                memset(&o, 0, sizeof(o));
                o.triangle_index = -1;
            }
            if (ctx.tie) per[t].ties++;
        }
    });
    exportStats(per, stats);
}

static void writeCastHit(orc_cast_hit &o, bool have, const CapsuleCastHit &h) {
    if (have) {
        o.toi = h.toi;
        st3(o.position, h.position);
        st3(o.normal, h.normal);
        st3(o.triangle_normal, h.triangleNormal);
        o.triangle_index = h.triangleIndex;
    } else {
        memset(&o, 0, sizeof(o));
        o.triangle_index = -1;
    }
}

void orc_capsule_cast(orc_world *w, const orc_cast *q, int32_t n, int32_t mode, orc_cast_hit *out, int32_t order,
                      int32_t n_threads, orc_stats *stats) {
    std::vector<Stats> per(std::max(1, n_threads));
    parallelFor(n, n_threads, [&](int t, int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            QueryCtx ctx;
            ctx.order = order;
            ctx.stats = stats ? &per[t] : nullptr;
            CapsuleCastHit h;
            bool have;
            F3 from = ld3(q[i].from), delta = ld3(q[i].delta);
            if (mode == 1) have = w->mesh.capsuleCastBlocking(from, delta, q[i].radius, q[i].half_height, q[i].mask, ctx, h);
            else if (mode == 2)
                have = w->mesh.capsuleCastGround(from, delta, q[i].radius, q[i].half_height, q[i].min_normal_y,
                                                 q[i].mask, ctx, h);
            else have = w->mesh.capsuleCast(from, delta, q[i].radius, q[i].half_height, q[i].mask, ctx, h);
            writeCastHit(out[i], have, h);
            if (ctx.tie) per[t].ties++;
        }
    });
    exportStats(per, stats);
}

static void writeOverlapHit(orc_overlap_hit &o, bool have, const CapsuleOverlapHit &h) {
    if (have) {
        o.depth = h.depth;
        st3(o.position, h.position);
        st3(o.normal, h.normal);
        st3(o.triangle_normal, h.triangleNormal);
        o.triangle_index = h.triangleIndex;
    } else {
        memset(&o, 0, sizeof(o));
        o.triangle_index = -1;
    }
}

void orc_capsule_overlap(orc_world *w, const orc_capsule *q, int32_t n, orc_overlap_hit *out, int32_t order,
                         int32_t n_threads, orc_stats *stats) {
    std::vector<Stats> per(std::max(1, n_threads));
    parallelFor(n, n_threads, [&](int t, int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            QueryCtx ctx;
            ctx.order = order;
            ctx.stats = stats ? &per[t] : nullptr;
            CapsuleOverlapHit h;
            bool have = w->mesh.capsuleOverlap(ld3(q[i].from), q[i].radius, q[i].half_height, q[i].mask, ctx, h);
            writeOverlapHit(out[i], have, h);
            if (ctx.tie) per[t].ties++;
        }
    });
    exportStats(per, stats);
}

void orc_capsule_overlap_all(orc_world *w, const orc_capsule *q, int32_t n, int32_t max_hits, orc_overlap_hit *out,
                             int32_t *counts, uint8_t *overflow, int32_t order, int32_t n_threads,
                             orc_stats *stats) {
    max_hits = std::max(1, max_hits); // CQ:157
    std::vector<Stats> per(std::max(1, n_threads));
    parallelFor(n, n_threads, [&](int t, int lo, int hi) {
        std::vector<CapsuleOverlapHit> hits;
        for (int i = lo; i < hi; i++) {
            QueryCtx ctx;
            ctx.order = order;
            ctx.stats = stats ? &per[t] : nullptr;
            w->mesh.capsuleOverlapAll(ld3(q[i].from), q[i].radius, q[i].half_height, max_hits, q[i].mask, ctx, hits);
            // the reference's callers sort by depth (SYS:759); emit in that order, stable
            std::stable_sort(hits.begin(), hits.end(),
                             [](const CapsuleOverlapHit &a, const CapsuleOverlapHit &b) { return a.depth > b.depth; });
            counts[i] = (int32_t)hits.size();
            for (int k = 0; k < max_hits; k++)
                writeOverlapHit(out[(size_t)i * max_hits + k], k < (int)hits.size(), k < (int)hits.size() ? hits[k] : CapsuleOverlapHit{});
            if (overflow) overflow[i] = ctx.overflow ? 1 : 0;
            if (ctx.tie) per[t].ties++;
            if (ctx.overflow) per[t].overflows++;
        }
    });
    exportStats(per, stats);
}

void orc_move_and_slide(orc_world *w, orc_state *inout, int32_t n, const orc_params *params, float dt,
                        const float gravity[3], uint32_t flags, int32_t order, int32_t n_threads,
                        orc_stats *stats) {
    orc_move_and_slide_ex(w, inout, n, params, dt, gravity, flags, order, n_threads, stats, nullptr, 0);
}

void orc_move_and_slide_ex(orc_world *w, orc_state *inout, int32_t n, const orc_params *params, float dt,
                           const float gravity[3], uint32_t flags, int32_t order, int32_t n_threads,
                           orc_stats *stats, const orc_platform *platforms, int32_t n_platforms) {
    std::vector<Stats> per(std::max(1, n_threads));
    F3 g = ld3(gravity);
    // flags & 2: every character is a solid agent (AgentCollisionComponent defaults); collectAgentStates (SYS:1592-1611)
    // snapshots position / velocity of all of them BEFORE the loop — after GravitySystem, which ran as system #3
    std::vector<AgentState> agents;
    AgentSnapshot snap = {nullptr, 0};
    if (flags & 2u) {
        agents.resize(n);
        for (int i = 0; i < n; i++) {
            D3 v = {inout[i].velocity[0], inout[i].velocity[1], inout[i].velocity[2]};
            if ((flags & 1u) && !(inout[i].grounded && inout[i].grounded_near)) v = v + d3(g) * (double)dt;
            agents[i] = {f3((float)inout[i].position[0], (float)inout[i].position[1], (float)inout[i].position[2]), f3(v),
                         params->radius, params->half_height};
        }
        snap = {agents.data(), n};
    }
    parallelFor(n, n_threads, [&](int t, int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            QueryCtx ctx;
            ctx.order = order;
            ctx.stats = stats ? &per[t] : nullptr;
            moveAndSlideOne(w->mesh, ctx, inout[i], *params, dt, g, flags, platforms, n_platforms,
                            (flags & 2u) ? &snap : nullptr, i);
            if (ctx.tie) per[t].ties++;
            if (ctx.overflow) per[t].overflows++;
        }
    });
    exportStats(per, stats);
}

float orc_segment_triangle_distance(const float center[3], float half_height, const float v0[3], const float v1[3],
                                    const float v2[3], float seg_pt[3], float tri_pt[3]) {
    SegTri r = segmentTriangleDistance(ld3(center), half_height, ld3(v0), ld3(v1), ld3(v2), nullptr);
    st3(seg_pt, r.seg);
    st3(tri_pt, r.tri);
    return r.dist;
}
void orc_segment_triangle_distance_batch(int32_t n, const float *centers, const float *hh, const float *tris,
                                         float *dist, float *seg, float *tri) {
    for (int i = 0; i < n; i++) {
        SegTri r = segmentTriangleDistance(ld3(centers + 3 * i), hh[i], ld3(tris + 9 * i), ld3(tris + 9 * i + 3),
                                           ld3(tris + 9 * i + 6), nullptr);
        dist[i] = r.dist;
        st3(seg + 3 * i, r.seg);
        st3(tri + 3 * i, r.tri);
    }
}
// capsuleCapsuleSweep, packed: from/delta/otherPos/otherDelta (n,3), radius/halfHeight pairs (n,4) = r, hh, oR, oHh
// -> hit (n), toi (n), normal (n,3)
void orc_capsule_capsule_sweep_batch(int32_t n, const float *from, const float *delta, const float *otherPos,
                                     const float *otherDelta, const float *dims, int32_t *hit, float *toi, float *normal) {
    for (int i = 0; i < n; i++) {
        CapsuleCapsuleHit h = {0, f3(0, 0, 0), -1};
        hit[i] = capsuleCapsuleSweep(ld3(from + 3 * i), ld3(delta + 3 * i), dims[4 * i], dims[4 * i + 1], i,
                                     ld3(otherPos + 3 * i), ld3(otherDelta + 3 * i), dims[4 * i + 2], dims[4 * i + 3], h)
                     ? 1
                     : 0;
        toi[i] = h.toi;
        st3(normal + 3 * i, h.normal);
    }
}
void orc_ray_triangle_batch(int32_t n, const float *origins, const float *dirs, const float *tris, float *tout,
                            int32_t *hit) {
    for (int i = 0; i < n; i++) {
        float t = 0;
        hit[i] = rayTriangle(ld3(origins + 3 * i), ld3(dirs + 3 * i), ld3(tris + 9 * i), ld3(tris + 9 * i + 3),
                             ld3(tris + 9 * i + 6), 1e-6f, t)
                     ? 1
                     : 0;
        tout[i] = t;
    }
}
float orc_closest_point_on_triangle(const float p[3], const float a[3], const float b[3], const float c[3],
                                    float out_pt[3]) {
    DistSqPt r = closestPointOnTriangle(ld3(p), ld3(a), ld3(b), ld3(c));
    st3(out_pt, r.p);
    return r.d;
}
float orc_segment_segment_distance_sq(const float p1[3], const float q1[3], const float p2[3], const float q2[3],
                                      float c1[3], float c2[3]) {
    SegSeg r = segmentSegmentDistanceSq(ld3(p1), ld3(q1), ld3(p2), ld3(q2));
    st3(c1, r.s);
    st3(c2, r.t);
    return r.d;
}

} // extern "C"
