// Host build of the DEVICE narrow-phase source (swift-game-engine_b200/csrc/cq_math.cuh) so that the CPU
// test-suite can compare it bit-exactly with the oracle on millions of random inputs — before any GPU time
// is spent.  Compiled by nvcc as host code with -Xcompiler -ffp-contract=off (tests/test_device_math_on_host.py).
#include "../../swift-game-engine_b200/csrc/cq_math.cuh"

extern "C" {

float hm_segment_triangle_distance(const float *c, float hh, const float *v0, const float *v1, const float *v2,
                                   float *seg, float *tri) {
    cq::Tri T = {{v0[0], v0[1], v0[2]}, {v1[0], v1[1], v1[2]}, {v2[0], v2[1], v2[2]}};
    cq::f3 s, t;
    float d = cq::segment_triangle_distance<true>(cq::f3{c[0], c[1], c[2]}, hh, T, s, t);
    seg[0] = s.x, seg[1] = s.y, seg[2] = s.z;
    tri[0] = t.x, tri[1] = t.y, tri[2] = t.z;
    return d;
}

// n cases, packed: centers (n,3), hh (n), tris (n,9) -> dist (n), seg (n,3), tri (n,3)
void hm_segment_triangle_distance_batch(int n, const float *centers, const float *hh, const float *tris, float *dist,
                                        float *seg, float *tri) {
    for (int i = 0; i < n; i++)
        dist[i] = hm_segment_triangle_distance(centers + 3 * i, hh[i], tris + 9 * i, tris + 9 * i + 3, tris + 9 * i + 6,
                                               seg + 3 * i, tri + 3 * i);
}

void hm_ray_triangle_batch(int n, const float *origins, const float *dirs, const float *tris, float *tout, int *hit) {
    for (int i = 0; i < n; i++) {
        cq::Tri T = {{tris[9 * i], tris[9 * i + 1], tris[9 * i + 2]},
                     {tris[9 * i + 3], tris[9 * i + 4], tris[9 * i + 5]},
                     {tris[9 * i + 6], tris[9 * i + 7], tris[9 * i + 8]}};
        float t = 0;
        hit[i] = cq::ray_triangle(cq::f3{origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]},
                                  cq::f3{dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]}, T, t)
                     ? 1
                     : 0;
        tout[i] = t;
    }
}

void hm_capsule_capsule_sweep_batch(int n, const float *from, const float *delta, const float *otherPos,
                                    const float *otherDelta, const float *dims, int *hit, float *toi, float *normal) {
    for (int i = 0; i < n; i++) {
        cq::AgentHit h = {0.0f, {0.0f, 0.0f, 0.0f}, -1};
        hit[i] = cq::capsule_pair_sweep(cq::f3{from[3 * i], from[3 * i + 1], from[3 * i + 2]},
                                        cq::f3{delta[3 * i], delta[3 * i + 1], delta[3 * i + 2]}, dims[4 * i], dims[4 * i + 1], i,
                                        cq::f3{otherPos[3 * i], otherPos[3 * i + 1], otherPos[3 * i + 2]},
                                        cq::f3{otherDelta[3 * i], otherDelta[3 * i + 1], otherDelta[3 * i + 2]}, dims[4 * i + 2],
                                        dims[4 * i + 3], h)
                     ? 1
                     : 0;
        toi[i] = h.toi;
        normal[3 * i] = h.normal.x, normal[3 * i + 1] = h.normal.y, normal[3 * i + 2] = h.normal.z;
    }
}
}
