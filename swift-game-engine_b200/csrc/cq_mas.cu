// cq_mas.cu — in-kernel move-and-slide: one fixed step of the reference's
// KinematicMoveStopSystem.fixedUpdate per-entity body (Game/Systems.swift:1842-1901) for a batch of
// independent characters, one thread per character:
//   cache decay (SYS:1105) -> VelocityGate (SYS:1037) -> pre-sweep depenetration (SYS:734,1635)
//   -> <= maxSlideIterations blocking casts + SlideResolver plane clipping (SYS:1229,1658)
//   -> GroundProbe snap/fall/offset casts (SYS:826) -> GroundSnap (SYS:945) -> SlopeFriction (SYS:965)
//   -> writeBack (SYS:1802).
// All casts of one character go through ONE call site driven by a small phase machine, so lanes that
// are in different slide iterations / probe stages still execute the conservative-advancement code
// together.  Contact-plane clipping state (remaining, slide normal, position) stays in registers; the
// rarely touched contact-manifold cache is updated in place in the character record.
// Double precision exactly where the reference uses Double (velocity; SYS:792,882,958,1045,1368).
#include "cq_internal.h"

namespace cq {

#define MAS_THREADS 128

struct MasArgs {
    cq_controller_params p;
    float dt;
    float gx, gy, gz;
    uint32_t flags;
};

__device__ __forceinline__ f3 ld3(const float *p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ void st3(float *o, f3 v) {
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}

// ---- ContactManifoldCache / DefaultContactCachePolicy on the record in global memory (SYS:1102-1205)
__device__ __forceinline__ bool manifold_normal_for(const cq_character_state *c, int tri, f3 &out) {
    int cnt = c->manifold_count;
    for (int i = 0; i < cnt; i++)
        if (c->manifold_triangles[i] == tri) {
            out = ld3(c->manifold_normals[i]);
            return true;
        }
    return false;
}

__device__ __noinline__ void manifold_update(cq_character_state *c, int tri, f3 normal) {
    f3 n = normal;
    if (len2(n) < 1e-8f) return;
    c->manifold_frames = 8;
    int cnt = c->manifold_count;
    for (int i = 0; i < cnt; i++) {
        if (c->manifold_triangles[i] == tri) {
            f3 cached = ld3(c->manifold_normals[i]);
            if (dot(cached, n) < 0.0f) n = -n;
            const float blend = 0.25f;
            f3 combined = normalize(cached * (1.0f - blend) + n * blend);
            st3(c->manifold_normals[i], combined);
            st3(c->side_contact_normal, combined);
            return;
        }
    }
    if (cnt >= CQ_MANIFOLD_MAX) cnt -= 1; // removeLast
    for (int i = cnt; i > 0; i--) {      // insert at 0
        c->manifold_triangles[i] = c->manifold_triangles[i - 1];
        st3(c->manifold_normals[i], ld3(c->manifold_normals[i - 1]));
    }
    f3 nn = normalize(n);
    c->manifold_triangles[0] = tri;
    st3(c->manifold_normals[0], nn);
    c->manifold_count = cnt + 1;
    st3(c->side_contact_normal, nn);
}

__device__ __forceinline__ void cache_record(cq_character_state *c, int tri, f3 normal, bool isSide) {
    manifold_update(c, tri, normal);
    if (isSide) {
        st3(c->side_contact_normal, normalize(normal));
        c->side_contact_frames = 3;
    }
}

enum { PH_SLIDE = 0, PH_SNAP, PH_FALL, PH_GATE, PH_OFFSET, PH_DONE };

template <bool COUNT>
__global__ void __launch_bounds__(MAS_THREADS) k_move_and_slide(WorldView W, cq_character_state *__restrict__ states, int n,
                                                                MasArgs A, unsigned long long *gctr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Counters ctr = {0, 0, 0, 0};
    if (i < n) {
        cq_character_state *S = states + i;
        const cq_controller_params &P = A.p;
        const f3 gravity = {A.gx, A.gy, A.gz};
        d3 vel = {S->velocity[0], S->velocity[1], S->velocity[2]};
        const bool wasGrounded = S->grounded != 0, wasGroundedNear = S->grounded_near != 0;
        if (A.flags & CQ_MAS_APPLY_GRAVITY) { // GravitySystem (SYS:603-619)
            if (!(wasGrounded && wasGroundedNear)) vel = vel + to_d3(gravity) * (double)A.dt;
        }
        f3 position = {(float)S->position[0], (float)S->position[1], (float)S->position[2]}; // positionF
        // decay (SYS:1105-1116)
        {
            int scf = S->side_contact_frames;
            if (scf > 0) S->side_contact_frames = scf - 1;
            int mf = S->manifold_frames;
            if (mf > 0) {
                mf -= 1;
                S->manifold_frames = mf;
                if (mf == 0) {
                    S->manifold_count = 0;
                    st3(S->side_contact_normal, mk3(0, 0, 0));
                }
            }
        }
        // VelocityGate (SYS:1037-1051)
        if (wasGrounded && wasGroundedNear && vel.y < 0.0) vel.y = 0.0;
        d3 remD = vel * (double)A.dt;
        if (wasGrounded && wasGroundedNear && remD.y < 0.0) remD.y = 0.0;
        f3 remaining = to_f3(remD);

        // ---- pre-sweep depenetration (SYS:734-808, 1635-1656)
        {
            const float slop = smax(P.skin_width * 0.5f, 0.001f);
            bool didResolve = false;
            f3 normalSum = {0, 0, 0};
            float normalWeight = 0.0f;
#pragma unroll 1
            for (int it = 0; it < 4; it++) {
                // the two deepest overlaps in (depth desc, index asc) order; only they are used (SYS:764-767)
                float d0 = 0.0f, d1 = 0.0f;
                int t0 = -1, t1 = -1;
                f3 n0 = {0, 0, 0}, n1 = {0, 0, 0};
                capsule_overlap_visit<COUNT>(W, position, P.radius, P.half_height, P.collision_mask, ctr,
                                             [&](float depth, int gid, int part, const Tri &T, float dist, f3 sp, f3 tp) {
                                                 bool before0 = t0 < 0 || depth > d0 || (depth == d0 && gid < t0);
                                                 bool before1 = t1 < 0 || depth > d1 || (depth == d1 && gid < t1);
                                                 if (!before0 && !before1) return;
                                                 OverlapRec r;
                                                 overlap_contact(T, dist, sp, tp, P.radius, r);
                                                 if (before0) {
                                                     d1 = d0, t1 = t0, n1 = n0;
                                                     d0 = depth, t0 = gid, n0 = r.normal;
                                                 } else {
                                                     d1 = depth, t1 = gid, n1 = r.normal;
                                                 }
                                             });
                if (t0 < 0) break; // hits.isEmpty
                bool sideContact = n0.y < P.min_ground_dot;
                int useCount = sideContact ? 1 : (t1 >= 0 ? 2 : 1);
                float maxDepth = d0;
                f3 frameNormal = {0, 0, 0};
                for (int k = 0; k < useCount; k++) {
                    float hd = k == 0 ? d0 : d1;
                    int ht = k == 0 ? t0 : t1;
                    f3 hn = k == 0 ? n0 : n1;
                    maxDepth = smax(maxDepth, hd);
                    f3 nn = hn;
                    f3 cached;
                    if (manifold_normal_for(S, ht, cached)) nn = cached; // SYS:770-776
                    frameNormal = frameNormal + nn * hd;
                    cache_record(S, ht, nn, hn.y < P.min_ground_dot);
                }
                float fl = len(frameNormal);
                f3 depenNormal = fl > 1e-6f ? frameNormal / fl : frameNormal;
                float push = sideContact ? smax(maxDepth, 0.0f) : smax(maxDepth + slop, 0.0f);
                if (sideContact) push = smin(push, P.skin_width);
                if (push <= 1e-6f) break;
                position = position + depenNormal * push;
                d3 dn = to_d3(depenNormal);
                double vInto = dot(vel, dn);
                if (vInto < 0.0) vel = vel - dn * vInto;
                didResolve = true;
                normalSum = normalSum + depenNormal * maxDepth;
                normalWeight += maxDepth;
            }
            if (didResolve) {
                f3 depenN = normalWeight > 1e-6f ? normalize(normalSum / normalWeight) : normalize(normalSum);
                float into = dot(remaining, depenN);
                if (into < 0.0f) remaining = remaining - depenN * into;
            }
        }

        // ---- slide iterations + ground probe: every cast goes through the single call site below
        int phase = PH_SLIDE;
        int slideIt = 0, offsetIt = 0;
        bool haveLast = false;
        f3 lastSlideNormal = {0, 0, 0};
        float slideLen = 0.0f;
        // ground-probe temporaries (SYS:826-943)
        bool haveCenter = false, grounded = false, groundedNear = false, canSnap = false, nearGround = false;
        bool probeHit = false; // GroundProbeResult.hit != nil
        float centerToi = 0.0f, centerPosY = 0.0f;
        f3 centerNormal = {0, 0, 0}, centerTriNormal = {0, 0, 0};
        int centerTri = -1, centerPart = -1;
        float gDistance = FLT_MAX;
        f3 gNormal = {0, 1, 0};
        f3 normalSum = {0, 0, 0};
        float combineTol = 0.0f;
        const f3 down = {0.0f, -1.0f, 0.0f};
        const f3 snapDelta = down * P.snap_distance;

#pragma unroll 1
        while (phase != PH_DONE) {
            f3 qfrom = position, qdelta = snapDelta;
            int mode = CQ_MODE_GROUND;
            if (phase == PH_SLIDE) {
                if (slideIt >= P.max_slide_iterations) {
                    phase = PH_SNAP;
                    continue;
                }
                slideLen = len(remaining);
                if (slideLen < 1e-6f) { // SYS:1676
                    phase = PH_SNAP;
                    continue;
                }
                qdelta = remaining;
                mode = CQ_MODE_BLOCKING;
            } else if (phase == PH_SNAP) {
                if (!(P.snap_distance > 0.0f)) { // SYS:845
                    phase = PH_FALL;
                    continue;
                }
            } else if (phase == PH_FALL) {
                if (!(P.fall_probe_distance > 0.0f)) { // SYS:855
                    phase = PH_GATE;
                    continue;
                }
                qdelta = down * P.fall_probe_distance;
            } else if (phase == PH_GATE) {
                // SYS:868-925: validity gates after the centre + fall casts
                phase = PH_DONE;
                if (!haveCenter || !(centerToi <= P.snap_distance)) continue;
                probeHit = true;
                float baseCenterY = position.y - P.half_height;
                float bottomY = baseCenterY - P.radius;
                float groundTol = smax(P.skin_width, P.ground_snap_skin);
                bool validGroundPoint = centerPosY <= bottomY + groundTol;
                float groundNearThreshold = smax(P.ground_snap_skin, P.skin_width);
                nearGround = centerToi <= groundNearThreshold;
                groundedNear = nearGround;
                gDistance = centerToi;
                bool groundGateVel = vel.y <= 0.0;
                double vInto = dot(vel, to_d3(centerNormal));
                bool groundGateSpeed = vInto >= -(double)P.ground_snap_max_speed;
                bool groundGateToi = centerToi <= P.ground_snap_max_toi;
                canSnap = validGroundPoint && groundGateVel && (nearGround || groundGateSpeed || groundGateToi);
                if (wasGroundedNear && centerToi <= P.snap_distance) canSnap = validGroundPoint;
                if (validGroundPoint && (nearGround || canSnap)) {
                    grounded = true;
                    normalSum = centerTriNormal;
                    if (centerTriNormal.y < 0.98f && (wasGroundedNear || nearGround)) { // SYS:897
                        combineTol = smax(smax(P.ground_snap_skin, P.skin_width), 0.05f);
                        offsetIt = 0;
                        phase = PH_OFFSET;
                    }
                }
                continue;
            } else { // PH_OFFSET (SYS:898-921)
                float offset = P.radius * 0.6f;
                float ox = offsetIt == 0 ? offset : (offsetIt == 1 ? -offset : 0.0f);
                float oz = offsetIt == 2 ? offset : (offsetIt == 3 ? -offset : 0.0f);
                qfrom = position + mk3(ox, 0.0f, oz);
            }

            CastResult res;
            capsule_cast<COUNT>(W, qfrom, qdelta, P.radius, P.half_height, P.collision_mask, mode, P.min_ground_dot, res, ctr);

            if (phase == PH_SLIDE) {
                slideIt++;
                if (res.tri < 0) { // SYS:1759-1763
                    position = position + remaining;
                    remaining = mk3(0, 0, 0);
                    phase = PH_SNAP;
                    continue;
                }
                f3 hitN = res.hit.normal;
                const f3 hitTriN = res.hit.triNormal;
                const float hitToi = res.hit.toi;
                bool haveCachedSide = false;
                f3 cachedSide = {0, 0, 0};
                const int sideFrames = S->side_contact_frames;
                if (hitN.y < P.min_ground_dot && sideFrames > 0) { // SYS:1683-1694
                    f3 cached;
                    if (manifold_normal_for(S, res.tri, cached)) {
                        if (dot(cached, hitN) < 0.0f) cached = -cached;
                        hitN = cached;
                    }
                }
                if (hitN.y < P.min_ground_dot && sideFrames > 0) // SYS:1719-1726
                    haveCachedSide = manifold_normal_for(S, res.tri, cachedSide);

                // ---- SlideResolver.resolveHit, kinematicMove options, static hit (SYS:1229-1375)
                bool shouldBreak = false;
                {
                    f3 slideNormal = hitN;
                    bool groundLike = hitTriN.y >= P.min_ground_dot;
                    float contactSkin = groundLike ? P.ground_snap_skin : P.skin_width;
                    if (slideNormal.y < P.min_ground_dot && sideFrames > 0) { // SYS:1273-1292
                        if (haveCachedSide) {
                            f3 cn = cachedSide;
                            if (dot(cn, slideNormal) < 0.0f) cn = -cn;
                            slideNormal = cn;
                        } else {
                            f3 cached = ld3(S->side_contact_normal);
                            float cl = len2(cached);
                            if (cl > 1e-6f) {
                                f3 cn = cached / sqrtf(cl);
                                float dc = dot(cn, slideNormal);
                                if (fabsf(dc) > 0.5f) slideNormal = dc >= 0.0f ? cn : -cn;
                            }
                        }
                    }
                    bool resolved = false;
                    if (slideNormal.y < P.min_ground_dot) { // SYS:1294-1309
                        if (groundLike) slideNormal = hitTriN;
                        if (slideNormal.y < P.min_ground_dot) {
                            slideNormal.y = 0.0f;
                            float nl = len(slideNormal);
                            if (nl > 1e-5f) {
                                slideNormal = slideNormal / nl;
                            } else {
                                position = position + remaining;
                                remaining = mk3(0, 0, 0);
                                shouldBreak = true;
                                resolved = true;
                            }
                        }
                    }
                    if (!resolved) {
                        float into = dot(remaining, slideNormal);
                        float intoEps = 1e-4f * slideLen;
                        float effectiveSkin =
                            (hitToi <= contactSkin && into < -intoEps) ? smin(contactSkin, hitToi * 0.5f) : contactSkin;
                        float sticky = contactSkin * 0.1f;
                        if (hitToi <= sticky && into < -intoEps) { // SYS:1320
                            remaining = remaining - slideNormal * into;
                            shouldBreak = false;
                        } else if (into >= -intoEps) { // SYS:1324
                            if (wasGroundedNear && !groundLike && remaining.y < 0.0f) remaining.y = 0.0f;
                            position = position + remaining;
                            remaining = mk3(0, 0, 0);
                            shouldBreak = true;
                        } else if ((hitToi <= effectiveSkin && fabsf(into) <= intoEps) || into >= 0.0f) {
                            position = position + remaining;
                            remaining = mk3(0, 0, 0);
                            shouldBreak = true;
                        } else {
                            float moveDist = smax(hitToi - effectiveSkin, 0.0f); // SYS:1343
                            if (slideNormal.y >= P.min_ground_dot && remaining.y < 0.0f && moveDist > P.ground_sweep_max_step)
                                moveDist = P.ground_sweep_max_step;
                            f3 dir = remaining / slideLen;
                            position = position + dir * moveDist;
                            f3 leftover = remaining - dir * moveDist;
                            leftover = leftover - slideNormal * dot(leftover, slideNormal);
                            if (wasGrounded && wasGroundedNear && leftover.y < 0.0f) leftover.y = 0.0f;
                            float residual = dot(leftover, slideNormal);
                            if (fabsf(residual) < 1e-5f) leftover = leftover - slideNormal * residual;
                            if (len2(leftover) < 1e-8f) {
                                remaining = mk3(0, 0, 0);
                                shouldBreak = true;
                            } else {
                                remaining = leftover;
                                d3 sn = to_d3(slideNormal);
                                double vInto = dot(vel, sn); // SYS:1367-1372
                                if (vInto < 0.0) vel = vel - sn * vInto;
                                shouldBreak = false;
                            }
                        }
                    }
                }
                if (hitN.y < P.min_ground_dot) cache_record(S, res.tri, hitN, true); // SYS:1738-1743
                if (haveLast) {                                                     // SYS:1744-1754
                    float dn = dot(lastSlideNormal, hitN);
                    if (fabsf(dn) < 0.98f) {
                        f3 axis = cross(lastSlideNormal, hitN);
                        float al = len(axis);
                        if (al > 1e-5f) {
                            f3 an = axis / al;
                            remaining = an * dot(remaining, an);
                        }
                    }
                }
                lastSlideNormal = hitN;
                haveLast = true;
                if (shouldBreak) phase = PH_SNAP;
            } else if (phase == PH_SNAP) {
                haveCenter = res.tri >= 0;
                if (haveCenter) {
                    centerToi = res.hit.toi;
                    centerPosY = res.hit.position.y;
                    centerNormal = res.hit.normal;
                    centerTriNormal = res.hit.triNormal;
                    centerTri = res.tri;
                    centerPart = res.part;
                }
                phase = PH_FALL;
            } else if (phase == PH_FALL) {
                if (res.tri >= 0) gDistance = res.hit.toi; // SYS:864
                phase = PH_GATE;
            } else { // PH_OFFSET
                if (res.tri >= 0 && res.hit.toi <= centerToi + combineTol) {
                    if (dot(res.hit.triNormal, centerTriNormal) > 0.98f) normalSum = normalSum + res.hit.triNormal;
                }
                if (++offsetIt == 4) phase = PH_DONE;
            }
        }

        // ---- finish GroundProbe.resolve (SYS:923-937)
        float matMuS = 0.8f, matMuK = 0.6f;
        bool matFlatten = false;
        if (grounded) {
            float nl = len(normalSum);
            gNormal = nl > 1e-6f ? normalSum / nl : centerTriNormal;
            if (wasGroundedNear) {
                f3 prevNormal = ld3(S->ground_normal);
                if (dot(prevNormal, gNormal) > 0.9f) {
                    const float blend = 0.2f;
                    gNormal = normalize(prevNormal * (1.0f - blend) + gNormal * blend);
                }
            }
            if (centerPart >= 0 && centerPart < W.nParts) {
                float4 m = __ldg(W.materials + centerPart);
                matMuS = m.x, matMuK = m.y, matFlatten = m.z != 0.0f;
            }
            if (matFlatten) gNormal = mk3(0, 1, 0);
        }
        // ---- GroundSnap.apply (SYS:945-963)
        if (canSnap && probeHit) {
            float moveDist = smax(centerToi - P.ground_snap_skin, 0.0f);
            if (nearGround && moveDist > P.ground_snap_max_step) moveDist = P.ground_snap_max_step;
            position = position + down * moveDist;
            d3 cn = to_d3(centerNormal);
            double vIntoSnap = dot(vel, cn);
            if (vIntoSnap < 0.0) vel = vel - cn * vIntoSnap;
        }
        // ---- resolveGroundContact tail + SlopeFriction.apply (SYS:1787-1798, 965-1021)
        int transitionFrames = S->ground_transition_frames;
        bool sliding = S->ground_sliding != 0;
        if (grounded) {
            float normalUpDelta = gNormal.y - S->ground_normal[1];
            if (centerTri != S->ground_triangle_index && normalUpDelta > 0.02f) transitionFrames = 3;
        }
        if (!grounded) {
            sliding = false;
        } else {
            f3 normal = normalize(gNormal);
            if (normal.y > 0.98f) {
                transitionFrames = 0;
                sliding = false;
            } else if (transitionFrames > 0) {
                transitionFrames -= 1;
                sliding = false;
            } else {
                float gN = dot(gravity, normal);
                f3 gTan = gravity - normal * gN;
                float gTanLen = len(gTan);
                if (gTanLen > 0.5f) {
                    float gNMag = fabsf(gN);
                    f3 gTanDir = gTan / gTanLen;
                    d3 gTanDirD = to_d3(gTanDir), normalD = to_d3(normal);
                    float stickLimit = matMuS * gNMag;
                    bool enterSlide = gTanLen > stickLimit * 1.05f;
                    bool exitSlide = gTanLen < stickLimit * 0.9f;
                    if (sliding) {
                        if (exitSlide) sliding = false;
                    } else if (enterSlide) {
                        sliding = true;
                    }
                    if (!sliding && gTanLen <= stickLimit) {
                        d3 vTan = vel - normalD * dot(vel, normalD);
                        double downhill = dot(vTan, gTanDirD);
                        if (downhill > 0.0) vel = vel - gTanDirD * downhill;
                    } else {
                        float slideAccelMag = smax(gTanLen - matMuK * gNMag, 0.0f);
                        if (slideAccelMag > 0.0f) vel = vel + gTanDirD * (double)slideAccelMag * (double)A.dt;
                    }
                }
            }
        }
        // ---- writeBack (SYS:1802-1821)
        S->position[0] = (double)position.x;
        S->position[1] = (double)position.y;
        S->position[2] = (double)position.z;
        S->velocity[0] = vel.x;
        S->velocity[1] = vel.y;
        S->velocity[2] = vel.z;
        S->grounded = grounded ? 1 : 0;
        S->grounded_near = groundedNear ? 1 : 0;
        S->ground_sliding = sliding ? 1 : 0;
        S->ground_transition_frames = transitionFrames;
        st3(S->ground_normal, grounded ? gNormal : mk3(0, 1, 0));
        S->ground_distance = gDistance;
        if (grounded) S->ground_triangle_index = centerTri;
    }
    // counters
    if (COUNT) {
        uint32_t v[4] = {ctr.nodes, ctr.cands, ctr.evals, ctr.queries};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned long long s = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(gctr + k, s);
        }
    }
}

int launch_move_and_slide(cq_world *w, cq_character_state *d_inout, int n, const cq_controller_params &p, float dt,
                          const float g[3], uint32_t flags, cudaStream_t st) {
    if (n <= 0) return CQ_OK;
    MasArgs A;
    A.p = p;
    A.dt = dt;
    A.gx = g[0], A.gy = g[1], A.gz = g[2];
    A.flags = flags;
    int blocks = (n + MAS_THREADS - 1) / MAS_THREADS;
    if (w->counting) k_move_and_slide<true><<<blocks, MAS_THREADS, 0, st>>>(w->view, d_inout, n, A, w->dCounters);
    else k_move_and_slide<false><<<blocks, MAS_THREADS, 0, st>>>(w->view, d_inout, n, A, w->dCounters);
    w->launches++;
    return check_cuda(cudaGetLastError(), "k_move_and_slide");
}

} // namespace cq
