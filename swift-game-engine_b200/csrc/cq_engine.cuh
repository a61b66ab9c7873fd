// cq_engine.cuh — the flat per-lane query engine.
//
// Why: a straightforward "thread per query" kernel nests three data-dependent loops (BVH walk -> leaf
// triangles -> conservative advancement with 1..256 trips + 10 bisections).  Measured on B200 (ncu,
// profiles/r1_mas_v1_summary.txt) that shape runs with 2.0 of 32 lanes active: the warp serialises.
// Here every lane is a small resumable state machine and the kernel's main loop is
//
//     while (lanes alive) {
//         front end (divergent, cheap):  lanes without a candidate either run their owner's logic
//                                        (consume a finished query, post the next) or ACQUIRE the next
//                                        candidate triangle (BVH walk -> per-lane candidate list -> select)
//         back end  (convergent, ~80% of the instructions): every lane does ONE segment-triangle
//                                        distance evaluation and the candidate's state transition
//     }
//
// with a warp-uniform exit (__all_sync), so the reconvergence points sit exactly at those two blocks.
//
// Time-ordered conservative advancement.  The reference sweeps candidate triangles one after another,
// each to completion (CollisionQuery.swift:1055-1105).  Per-character work measured that way is very
// heavy-tailed (median 37 distance evaluations, p99 1135): candidates that "creep" beside the capsule
// advance by minAdvance for up to 256 trips before anything prunes them.  The result of a cast does not
// depend on the visiting order (it is the accepted candidate with the smallest (toi, index); SURVEY.md
// §A.4), so this engine keeps up to CQ_LIST candidates of the query in a per-lane shared-memory list and
// always advances the one whose next evaluation time t is smallest.  The first contact found is then
// (nearly) the earliest one, and every other candidate is dropped by the exact-safe prune
// `lastSafeT > bestT` after at most one more evaluation.  Same arithmetic per evaluation, same result,
// far fewer evaluations, and no heavy tail.
//
// Arithmetic is the reference's (CollisionQuery.swift:1285-1438), expression for expression; see
// cq_math.cuh.  Ties on toi go to the smallest triangle index.
#pragma once
#include "cq_world.cuh"

namespace cq {

enum { PH_NONE = 0, PH_ADV = 1, PH_BIS = 2, PH_FIN = 3, PH_OVL = 4 };
#define CQ_KIND_OVERLAP 3 /* LaneQ.mode value for the two-deepest overlap query used by move-and-slide */
#define CQ_LIST 8         /* candidates kept per lane between BVH walks */
#define CQ_DEAD 3.0e38f

// Per-lane candidate list in shared memory, interleaved so that lane l's word (e, f) lives at
// base[(e * 4 + f) * nThreads + l]: conflict-free when all lanes touch the same (e, f).
// fields: 0 = slot | set << 31, 1 = next evaluation time t (CQ_DEAD = finished), 2 = lastSafeT, 3 = trips done
struct CandList {
    float *base;
    int stride;
    __device__ __forceinline__ float &f(int e, int fld) const { return base[(e * 4 + fld) * stride]; }
};

struct LaneQ {
    // ---- query
    f3 from, dir, delta;
    float L, radius, hh, minNormalY, minAdvance;
    uint32_t mask;
    int mode; // CQ_MODE_ALL / BLOCKING / GROUND, or CQ_KIND_OVERLAP
    int maxIter;
    f3 qlo, qhi;
    // ---- traversal
    int sp, set, leafPos, leafEnd;
    bool travDone;
    bool done; // query finished: result fields are final
    // ---- candidate list bookkeeping
    int nList, cur;
    float minOther; // smallest t among the list's other live candidates (switch threshold)
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
    q.done = true;
    q.sp = 0;
    q.leafPos = q.leafEnd = 0;
    q.set = 1;
    q.nList = 0;
    q.cur = -1;
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

// capsuleCastCombined prologue — CollisionQuery.swift:980-1043.
template <bool COUNT>
__device__ __forceinline__ void q_begin_cast(const WorldView &W, LaneQ &q, int *stack, f3 from, f3 delta, float radius,
                                             float hh, uint32_t mask, int mode, float minNormalY, Counters &ctr) {
    q.bestTri = -1;
    q.bestPart = -1;
    q.phase = PH_NONE;
    q.leafPos = q.leafEnd = 0;
    q.sp = 0;
    q.nList = 0;
    q.cur = -1;
    float L = len(delta);
    if (L < 1e-6f) { // nil without any traversal (:988)
        q.travDone = true;
        q.done = true;
        return;
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
    q.minAdvance = smax(radius * 0.02f, 1e-4f);            // :1295
    q.maxIter = min(256, (int)ceilf(L / q.minAdvance) + 1); // :1296
    const f3 up = {0.0f, 1.0f, 0.0f};
    f3 a0 = from + up * hh, b0 = from - up * hh;
    f3 a1 = a0 + delta, b1 = b0 + delta;
    f3 ext = {radius, radius, radius};
    q.qlo = vmin(vmin(a0, b0), vmin(a1, b1)) - ext;
    q.qhi = vmax(vmax(a0, b0), vmax(a1, b1)) + ext;
    q.bestT = L;
    q.set = 0;
    q.travDone = false;
    q.done = false;
    q_push_root(W, q, stack, ctr, COUNT);
}

// capsuleOverlapAll prologue — CollisionQuery.swift:1209-1216 (two deepest kept, Systems.swift:764-767)
template <bool COUNT>
__device__ __forceinline__ void q_begin_overlap(const WorldView &W, LaneQ &q, int *stack, f3 from, float radius, float hh,
                                                uint32_t mask, Counters &ctr) {
    q.phase = PH_NONE;
    q.leafPos = q.leafEnd = 0;
    q.nList = 0;
    q.cur = -1;
    q.from = from;
    q.radius = radius;
    q.hh = hh;
    q.mask = mask;
    q.mode = CQ_KIND_OVERLAP;
    q.dir = mk3(0, 0, 0);
    q.L = 0.0f;
    q.maxIter = 1;
    overlap_box(from, radius, hh, q.qlo, q.qhi);
    q.bestT = 0.0f, q.bestTri = -1, q.bestN = mk3(0, 0, 0);         // deepest
    q.bestPos.x = 0.0f, q.bestPart = -1, q.bestTriN = mk3(0, 0, 0); // second deepest
    q.set = 0;
    q.travDone = false;
    q.done = false;
    q_push_root(W, q, stack, ctr, COUNT);
}

// ---- front end: make the lane hold a candidate (q.phase != PH_NONE) or finish the query (q.done)
template <bool COUNT>
__device__ __forceinline__ void q_acquire(const WorldView &W, LaneQ &q, int *stack, const CandList &cl, Counters &ctr) {
    const bool overlap = q.mode == CQ_KIND_OVERLAP;
    while (true) {
        // 1. select: live candidate with the smallest next evaluation time (prune first)
        int best = -1;
        float bestKey = CQ_DEAD, second = CQ_DEAD;
        for (int e = 0; e < q.nList; e++) {
            float t = cl.f(e, 1);
            if (t >= CQ_DEAD) continue;
            if (!overlap && cl.f(e, 2) > q.bestT) { // exact-safe prune: toi >= lastSafeT > bestT
                cl.f(e, 1) = CQ_DEAD;
                continue;
            }
            if (t < bestKey) {
                second = bestKey;
                bestKey = t;
                best = e;
            } else if (t < second) {
                second = t;
            }
        }
        if (best >= 0) {
            uint32_t enc = __float_as_uint(cl.f(best, 0));
            int set = enc >> 31, slot = enc & 0x7fffffffu;
            const float4 *p0 = set ? W.set[1].tv0 : W.set[0].tv0;
            const float4 *p1 = set ? W.set[1].tv1 : W.set[0].tv1;
            const float4 *p2 = set ? W.set[1].tv2 : W.set[0].tv2;
            float4 a = __ldg(p0 + slot), b = __ldg(p1 + slot), c = __ldg(p2 + slot);
            q.T.v0 = xyz(a), q.T.v1 = xyz(b), q.T.v2 = xyz(c);
            q.gid = __float_as_int(b.w) + (set ? W.set[1].triOffset : 0);
            q.part = __float_as_int(c.w);
            q.cur = best;
            q.minOther = second;
            q.t = bestKey;
            q.lastSafeT = cl.f(best, 2);
            q.it = __float_as_int(cl.f(best, 3));
            q.phase = overlap ? PH_OVL : PH_ADV;
            return;
        }
        // 2. list exhausted: refill from the BVH walk, or finish
        q.nList = 0;
        q.cur = -1;
        if (q.travDone) {
            q.done = true;
            return;
        }
        while (!q.travDone && q.nList < CQ_LIST) {
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
                int e = q.nList++;
                cl.f(e, 0) = __uint_as_float((uint32_t)slot | ((uint32_t)q.set << 31));
                cl.f(e, 1) = 0.0f;
                cl.f(e, 2) = 0.0f;
                cl.f(e, 3) = __int_as_float(0);
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
        if (q.nList == 0) { // nothing found: travDone must be true now
            q.done = true;
            return;
        }
    }
}

__device__ __forceinline__ void q_retire(LaneQ &q, const CandList &cl) { // current candidate is finished
    cl.f(q.cur, 1) = CQ_DEAD;
    q.phase = PH_NONE;
}

// ---- back end: one distance evaluation + the candidate's state transition
template <bool COUNT> __device__ __forceinline__ void q_eval_step(LaneQ &q, const CandList &cl, Counters &ctr) {
    const int ph = q.phase;
    float tc = ph == PH_ADV ? q.t : (ph == PH_BIS ? 0.5f * (q.lo + q.hi) : q.hi);
    f3 center = ph == PH_OVL ? q.from : q.from + q.dir * tc;
    f3 sp, tp;
    if (COUNT) ctr.evals++;
    float dist = segment_triangle_distance<true>(center, q.hh, q.T, sp, tp);
    if (ph == PH_ADV) { // sweepCapsuleTriangle loop body, CollisionQuery.swift:1303-1356
        if (dist <= q.radius + 1e-5f) {
            // refineTOI prologue, :1371-1377.  The refine + contact evaluation run uninterrupted.
            float c0 = smax(0.0f, smin(q.lastSafeT, q.L));
            float c1 = smax(0.0f, smin(q.t, q.L));
            q.lo = smin(c0, c1);
            q.hi = smax(c0, c1);
            if (q.hi - q.lo < 1e-5f) {
                if (q.hi > q.bestT) q_retire(q, cl);
                else q.phase = PH_FIN;
            } else {
                q.k = 0;
                if (q.lo > q.bestT) q_retire(q, cl);
                else q.phase = PH_BIS;
            }
        } else {
            q.lastSafeT = q.t;
            float advance = smax(dist - q.radius, q.minAdvance);
            q.t += advance <= 0.0f ? q.minAdvance : advance;
            q.it++;
            // next trip: `for _ in 0..<maxIter { if t > maxDistance return nil ...`; prune: toi >= lastSafeT
            if (q.it >= q.maxIter || q.t > q.L || q.lastSafeT > q.bestT) {
                q_retire(q, cl);
            } else if (q.t > q.minOther) { // another candidate is now earlier in time: park this one
                cl.f(q.cur, 1) = q.t;
                cl.f(q.cur, 2) = q.lastSafeT;
                cl.f(q.cur, 3) = __int_as_float(q.it);
                q.phase = PH_NONE;
            }
        }
    } else if (ph == PH_BIS) { // refineTOI bisection, :1379-1392 (threshold is radius, not radius+eps)
        if (dist <= q.radius) q.hi = tc;
        else q.lo = tc;
        q.k++;
        if (q.k == 10) {
            if (q.hi > q.bestT) q_retire(q, cl);
            else q.phase = PH_FIN;
        } else if (q.lo > q.bestT) {
            q_retire(q, cl);
        }
    } else if (ph == PH_FIN) { // contact at tHit = hi, :1325-1346, then the acceptance test of :1084-1099
        q_retire(q, cl);
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
        q_retire(q, cl);
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
