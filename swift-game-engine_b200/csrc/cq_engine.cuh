// cq_engine.cuh — the flat per-lane query engine.
//
// Why: a straightforward "thread per query" kernel nests three data-dependent loops (BVH walk -> leaf
// triangles -> conservative advancement with 1..256 trips + 10 bisections).  Measured on B200 (ncu,
// profiles/r1_mas_v1_summary.txt) that shape runs with 2.0 of 32 lanes active: the warp serialises.
// Here every lane is a small resumable state machine and the kernel's main loop has exactly three
// stages that all 32 lanes pass through together:
//     L  (logic)     lanes whose query just finished consume the result / post the next query
//     T  (traverse)  lanes without an active candidate walk the LBVH until they hold one (or finish)
//     E  (evaluate)  every lane holding a candidate does ONE segment-triangle distance evaluation
// E is ~80% of the instructions and is executed convergently; the loop exit is warp-uniform
// (__all_sync), so the compiler's reconvergence points sit exactly at the three `if`s.
//
// Arithmetic is the reference's (CollisionQuery.swift:1285-1438), expression for expression; see
// cq_math.cuh.  The exact-safe prunes of SURVEY.md §A.4-3 are applied (a candidate whose toi is
// provably > the best accepted toi is dropped); ties on toi go to the smallest triangle index.
#pragma once
#include "cq_world.cuh"

namespace cq {

enum { PH_NONE = 0, PH_ADV = 1, PH_BIS = 2, PH_FIN = 3, PH_OVL = 4 };
#define CQ_KIND_OVERLAP 3 /* LaneQ.mode value for the two-deepest overlap query used by move-and-slide */

struct LaneQ {
    // ---- query
    f3 from, dir, delta;
    float L, radius, hh, minNormalY, minAdvance;
    uint32_t mask;
    int mode;   // CQ_MODE_ALL / BLOCKING / GROUND, or CQ_KIND_OVERLAP
    int maxIter;
    f3 qlo, qhi;
    // ---- traversal
    int sp, set, leafPos, leafEnd;
    bool travDone;
    // ---- active candidate
    int phase;
    Tri T;
    int gid, part;
    float t, lastSafeT, lo, hi;
    int it, k;
    // ---- result.  cast: bestT/bestTri/bestPart + contact.  overlap: two deepest, aliased as
    //      d0=bestT t0=bestTri n0=bestN | d1=bestPos.x t1=bestPart n1=bestTriN
    float bestT;
    int bestTri, bestPart;
    f3 bestPos, bestN, bestTriN;
};

__device__ __forceinline__ void q_idle(LaneQ &q) {
    q.phase = PH_NONE;
    q.travDone = true;
    q.sp = 0;
    q.leafPos = q.leafEnd = 0;
    q.set = 1;
}

__device__ __forceinline__ void q_push_root(const WorldView &W, LaneQ &q, int *stack, Counters &ctr, bool count) {
    const SetHeader *hp = q.set ? W.set[1].hdr : W.set[0].hdr;
    SetHeader h = *hp;
    q.sp = 0;
    if (h.rootRef == CQ_REF_EMPTY) return;
    if (count) {
        ctr.queries++;
        ctr.nodes++;
    }
    if (box_disjoint(mk3(h.lo[0], h.lo[1], h.lo[2]), mk3(h.hi[0], h.hi[1], h.hi[2]), q.qlo, q.qhi)) return;
    stack[q.sp++] = h.rootRef;
}

// capsuleCastCombined prologue — CollisionQuery.swift:980-1043.  Returns false when the query is nil
// without any traversal (|delta| < 1e-6).
template <bool COUNT>
__device__ __forceinline__ bool q_begin_cast(const WorldView &W, LaneQ &q, int *stack, f3 from, f3 delta, float radius,
                                             float hh, uint32_t mask, int mode, float minNormalY, Counters &ctr) {
    q.bestTri = -1;
    q.bestPart = -1;
    q.phase = PH_NONE;
    q.leafPos = q.leafEnd = 0;
    q.sp = 0;
    float L = len(delta);
    if (L < 1e-6f) {
        q.travDone = true;
        return false;
    }
    q.from = from;
    q.delta = delta;
    q.L = L;
    q.dir = delta / L;
    q.radius = radius;
    q.hh = hh;
    q.mask = mask;
    q.mode = mode;
    q.minNormalY = minNormalY;
    q.minAdvance = smax(radius * 0.02f, 1e-4f);                      // :1295
    q.maxIter = min(256, (int)ceilf(L / q.minAdvance) + 1);           // :1296
    const f3 up = {0.0f, 1.0f, 0.0f};
    f3 a0 = from + up * hh, b0 = from - up * hh;
    f3 a1 = a0 + delta, b1 = b0 + delta;
    f3 ext = {radius, radius, radius};
    q.qlo = vmin(vmin(a0, b0), vmin(a1, b1)) - ext;
    q.qhi = vmax(vmax(a0, b0), vmax(a1, b1)) + ext;
    q.bestT = L;
    q.set = 0;
    q.travDone = false;
    q_push_root(W, q, stack, ctr, COUNT);
    return true;
}

// capsuleOverlapAll prologue — CollisionQuery.swift:1209-1216 (two deepest kept, Systems.swift:764-767)
template <bool COUNT>
__device__ __forceinline__ void q_begin_overlap(const WorldView &W, LaneQ &q, int *stack, f3 from, float radius, float hh,
                                                uint32_t mask, Counters &ctr) {
    q.phase = PH_NONE;
    q.leafPos = q.leafEnd = 0;
    q.from = from;
    q.radius = radius;
    q.hh = hh;
    q.mask = mask;
    q.mode = CQ_KIND_OVERLAP;
    q.dir = mk3(0, 0, 0);
    overlap_box(from, radius, hh, q.qlo, q.qhi);
    q.bestT = 0.0f, q.bestTri = -1, q.bestN = mk3(0, 0, 0);         // deepest
    q.bestPos.x = 0.0f, q.bestPart = -1, q.bestTriN = mk3(0, 0, 0); // second deepest
    q.set = 0;
    q.travDone = false;
    q_push_root(W, q, stack, ctr, COUNT);
}

// ---- stage T: advance the traversal until the lane holds a candidate triangle or the query is finished
template <bool COUNT>
__device__ __forceinline__ void q_next_candidate(const WorldView &W, LaneQ &q, int *stack, Counters &ctr) {
    while (q.phase == PH_NONE && !q.travDone) {
        if (q.leafPos < q.leafEnd) {
            int slot = q.leafPos++;
            const float4 *p0 = q.set ? W.set[1].tv0 : W.set[0].tv0;
            const float4 *p1 = q.set ? W.set[1].tv1 : W.set[0].tv1;
            const float4 *p2 = q.set ? W.set[1].tv2 : W.set[0].tv2;
            float4 a = __ldg(p0 + slot), b = __ldg(p1 + slot), c = __ldg(p2 + slot);
            if ((__float_as_uint(a.w) & q.mask) == 0u) continue; // layer mask, CollisionQuery.swift:1057
            f3 v0 = xyz(a), v1 = xyz(b), v2 = xyz(c);
            f3 tlo = vmin(v0, vmin(v1, v2)), thi = vmax(v0, vmax(v1, v2));
            if (box_disjoint(tlo, thi, q.qlo, q.qhi)) continue; // :1060-1065
            if (COUNT) ctr.cands++;
            q.T.v0 = v0, q.T.v1 = v1, q.T.v2 = v2;
            q.gid = __float_as_int(b.w) + (q.set ? W.set[1].triOffset : 0);
            q.part = __float_as_int(c.w);
            if (q.mode == CQ_KIND_OVERLAP) {
                q.phase = PH_OVL;
            } else {
                q.phase = PH_ADV;
                q.t = 0.0f;
                q.lastSafeT = 0.0f;
                q.it = 0;
            }
        } else if (q.sp > 0) {
            int ref = stack[--q.sp];
            if (ref < 0) {
                int enc = ~ref;
                q.leafPos = enc >> 2;
                q.leafEnd = q.leafPos + (enc & 3) + 1;
            } else {
                const Node *n = (q.set ? W.set[1].nodes : W.set[0].nodes) + ref;
                float4 n0 = __ldg(&n->n0), n1 = __ldg(&n->n1), n2 = __ldg(&n->n2), n3 = __ldg(&n->n3);
                if (COUNT) ctr.nodes += 2;
                if (!box_disjoint(xyz(n2), xyz(n3), q.qlo, q.qhi)) stack[q.sp++] = __float_as_int(n1.w);
                if (!box_disjoint(xyz(n0), xyz(n1), q.qlo, q.qhi)) stack[q.sp++] = __float_as_int(n0.w);
            }
        } else if (q.set == 0) {
            q.set = 1; // static set done -> dynamic set (CollisionQuery.swift:990-1008)
            q_push_root(W, q, stack, ctr, COUNT);
        } else {
            q.travDone = true;
        }
    }
}

// ---- stage E: one distance evaluation + the candidate's state transition
template <bool COUNT> __device__ __forceinline__ void q_eval_step(LaneQ &q, Counters &ctr) {
    const int ph = q.phase;
    float tc = ph == PH_ADV ? q.t : (ph == PH_BIS ? 0.5f * (q.lo + q.hi) : q.hi);
    f3 center = ph == PH_OVL ? q.from : q.from + q.dir * tc;
    f3 sp, tp;
    if (COUNT) ctr.evals++;
    float dist = segment_triangle_distance<true>(center, q.hh, q.T, sp, tp);
    if (ph == PH_ADV) { // sweepCapsuleTriangle loop body, CollisionQuery.swift:1303-1356
        if (dist <= q.radius + 1e-5f) {
            // refineTOI prologue, :1371-1377
            float c0 = smax(0.0f, smin(q.lastSafeT, q.L));
            float c1 = smax(0.0f, smin(q.t, q.L));
            q.lo = smin(c0, c1);
            q.hi = smax(c0, c1);
            if (q.hi - q.lo < 1e-5f) {
                q.phase = q.hi > q.bestT ? PH_NONE : PH_FIN;
            } else {
                q.k = 0;
                q.phase = q.lo > q.bestT ? PH_NONE : PH_BIS;
            }
        } else {
            q.lastSafeT = q.t;
            float advance = smax(dist - q.radius, q.minAdvance);
            q.t += advance <= 0.0f ? q.minAdvance : advance;
            q.it++;
            // next trip: `for _ in 0..<maxIter { if t > maxDistance return nil ...`; prune: toi >= lastSafeT
            if (q.it >= q.maxIter || q.t > q.L || q.lastSafeT > q.bestT) q.phase = PH_NONE;
        }
    } else if (ph == PH_BIS) { // refineTOI bisection, :1379-1392 (threshold is radius, not radius+eps)
        if (dist <= q.radius) q.hi = tc;
        else q.lo = tc;
        q.k++;
        if (q.k == 10) q.phase = q.hi > q.bestT ? PH_NONE : PH_FIN;
        else if (q.lo > q.bestT) q.phase = PH_NONE;
    } else if (ph == PH_FIN) { // contact at tHit = hi, :1325-1346, then the acceptance test of :1084-1099
        q.phase = PH_NONE;
        f3 triNormal = normalize(cross(q.T.v1 - q.T.v0, q.T.v2 - q.T.v0));
        f3 n;
        if (dist < 1e-6f) n = dot(triNormal, q.dir) > 0.0f ? -triNormal : triNormal;
        else n = normalize(sp - tp);
        f3 triN = triNormal;
        if (dot(triN, n) < 0.0f) triN = -triN;
        bool better = tc < q.bestT;
        bool tieWin = q.bestTri >= 0 && tc == q.bestT && q.gid < q.bestTri;
        bool ok = better || tieWin;
        if (q.mode == CQ_MODE_BLOCKING) ok = ok && !(dot(q.delta, n) >= 0.0f) && !(dot(q.delta, triN) >= 0.0f);
        else if (q.mode == CQ_MODE_GROUND) ok = ok && !(triN.y < q.minNormalY);
        if (ok) {
            q.bestT = tc;
            q.bestTri = q.gid;
            q.bestPart = q.part;
            q.bestPos = tp;
            q.bestN = n;
            q.bestTriN = triN;
        }
    } else { // PH_OVL: capsuleOverlapBVHAll leaf body, :1248-1271; keep the two deepest (depth desc, index asc)
        q.phase = PH_NONE;
        if (dist < q.radius) {
            float depth = q.radius - dist;
            float d0 = q.bestT, d1 = q.bestPos.x;
            int t0 = q.bestTri, t1 = q.bestPart;
            bool before0 = t0 < 0 || depth > d0 || (depth == d0 && q.gid < t0);
            bool before1 = t1 < 0 || depth > d1 || (depth == d1 && q.gid < t1);
            if (before0 || before1) {
                f3 triNormal = normalize(cross(q.T.v1 - q.T.v0, q.T.v2 - q.T.v0));
                f3 n = dist < 1e-6f ? triNormal : normalize(sp - tp);
                if (before0) {
                    q.bestPos.x = d0, q.bestPart = t0, q.bestTriN = q.bestN;
                    q.bestT = depth, q.bestTri = q.gid, q.bestN = n;
                } else {
                    q.bestPos.x = depth, q.bestPart = q.gid, q.bestTriN = n;
                }
            }
        }
    }
}

} // namespace cq
