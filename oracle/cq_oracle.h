/*
 * cq_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  A plain C++ restatement of the reference's
 * algorithm (Game/CollisionQuery.swift:320-1632, the move-and-slide driver
 * Game/Systems.swift:603-1821 and AgentSeparationSystem :1906-2210).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (libcq.so) never links, loads or calls anything in this directory.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference has no tests, golden vectors
 * or fixtures for this path (GameTests/GameTests.swift:13-15 is an empty
 * template) and cannot be compiled here (no swiftc, Apple-only `simd`).  The
 * oracle is pinned instead by analytic known-answer tests, a brute-force
 * cross-check and committed golden files it produced itself (tests/golden/).
 *
 * The record layouts below are byte-identical to include/cq.h on purpose (the
 * same numpy dtypes feed both sides) but are declared independently.
 */
#ifndef CQ_ORACLE_H
#define CQ_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

typedef struct orc_part {
    const float *positions_xyz;
    const uint32_t *indices;
    int32_t n_verts;
    int32_t n_indices;
    float model[16]; /* column-major */
    uint32_t layer;
    float mu_s, mu_k;
    uint8_t flatten_ground;
    uint8_t is_dynamic;
    uint16_t _pad;
    uint32_t entity_id;
} orc_part;

typedef struct orc_ray { float origin[3], direction[3], max_distance; uint32_t mask; } orc_ray;
typedef struct orc_ray_hit { float distance, position[3], normal[3]; int32_t triangle_index; } orc_ray_hit;
typedef struct orc_cast { float from[3], delta[3], radius, half_height; uint32_t mask; float min_normal_y; } orc_cast;
typedef struct orc_cast_hit { float toi, position[3], normal[3], triangle_normal[3]; int32_t triangle_index; } orc_cast_hit;
typedef struct orc_capsule { float from[3], radius, half_height; uint32_t mask; } orc_capsule;
typedef struct orc_overlap_hit { float depth, position[3], normal[3], triangle_normal[3]; int32_t triangle_index; } orc_overlap_hit;

typedef struct orc_params {
    float radius, half_height, skin_width, ground_snap_skin, snap_distance, fall_probe_distance,
        ground_snap_max_speed, ground_snap_max_toi, ground_snap_max_step, ground_sweep_max_step;
    int32_t max_slide_iterations;
    float min_ground_dot;
    uint32_t collision_mask;
} orc_params;

typedef struct orc_state {
    double position[3], velocity[3];
    float ground_normal[3];
    float ground_distance;
    float side_contact_normal[3];
    int32_t ground_triangle_index, ground_transition_frames, side_contact_frames, manifold_frames,
        manifold_count;
    int32_t manifold_triangles[4];
    float manifold_normals[4][3];
    uint8_t grounded, grounded_near, ground_sliding, _pad[5];
} orc_state;

/* per-query work counters = the reference's CollisionQueryStats (CollisionQuery.swift:280-290) */
typedef struct orc_stats {
    int64_t candidates;      /* capsuleCandidateCount */
    int64_t sweep_tests;     /* capsuleSweepCount */
    int64_t sweep_iterations;/* capsuleSweepIterations (CA loop trips) */
    int64_t max_iterations;  /* capsuleSweepMaxIterations */
    int64_t distance_evals;  /* every segmentTriangleDistance call (CA + refine + final) */
    int64_t nodes_visited;   /* reference-BVH nodes popped */
    int64_t ties;            /* queries where >= 2 accepted candidates share the best key */
    int64_t overflows;       /* overlapAll queries with more than maxHits overlapping triangles */
} orc_stats;

/* order: 0 = REFERENCE (the reference's own BVH + DFS visiting order, first-visited wins ties)
 *        1 = CANONICAL (tree-independent rule the GPU implements: ties -> smallest triangle
 *            index; overlapAll -> the max_hits deepest; raycast -> brute force, no slab culling) */
#define ORC_ORDER_REFERENCE 0
#define ORC_ORDER_CANONICAL 1

orc_world *orc_world_create(const orc_part *parts, int32_t n_parts);
/* StaticMeshComponent.triangleMaterials of some parts: used when n == the part's triangle count, ignored otherwise
 * (CollisionQuery.swift:363-369). */
typedef struct orc_surface_material { float mu_s, mu_k; uint8_t flatten_ground; uint8_t _pad[3]; } orc_surface_material;
typedef struct orc_triangle_materials { uint32_t entity_id; int32_t n; const orc_surface_material *materials; } orc_triangle_materials;
orc_world *orc_world_create_ex(const orc_part *parts, int32_t n_parts, const orc_triangle_materials *tri_materials,
                               int32_t n_tri_materials);
/* TriangleMeshSet.materialForTriangle (CollisionQuery.swift:464-469) of a global triangle index: out = mu_s, mu_k, flatten */
void orc_world_triangle_material(const orc_world *w, int32_t triangle_index, float out[3]);
void orc_world_destroy(orc_world *w);
/* which: 0 static, 1 dynamic. out[0]=n_vertices out[1]=n_triangles out[2]=n_bvh_nodes */
void orc_world_counts(const orc_world *w, int32_t which, int32_t out[3]);
void orc_world_read_soup(const orc_world *w, int32_t which, float *positions, uint32_t *indices,
                         float *aabbs, uint32_t *layers, int32_t *parts);
void orc_world_update_transforms(orc_world *w, const uint32_t *entity_ids, const float *models, int32_t n);
/* reference BVH node bounds after build/refit, n_nodes*6 floats, and a from-scratch recomputation
 * flag: returns 1 if every internal node equals merge(children) and every leaf equals its range. */
int32_t orc_world_check_bvh(const orc_world *w, int32_t which);

/* AgentSeparationSystem.fixedUpdate (Systems.swift:2136-2210) over the batch (every character a solid agent):
 * `iterations` sequential grid + pair-resolution sweeps in index order, then per agent a <= 2-cast slide from its
 * pre-separation position and a ground snap.  mass_weight: per agent or NULL (1.0); use_query 0 = no world casts. */
void orc_agent_separation(orc_world *w, orc_state *inout, int32_t n, const orc_params *params, const float *mass_weight,
                          int32_t iterations, float separation_margin, float height_margin, int32_t use_query, int32_t order,
                          int32_t n_threads, int64_t *pair_count);
void orc_raycast(orc_world *w, const orc_ray *rays, int32_t n, orc_ray_hit *out, int32_t order,
                 int32_t n_threads, orc_stats *stats);
void orc_capsule_cast(orc_world *w, const orc_cast *q, int32_t n, int32_t mode, orc_cast_hit *out,
                      int32_t order, int32_t n_threads, orc_stats *stats);
void orc_capsule_overlap(orc_world *w, const orc_capsule *q, int32_t n, orc_overlap_hit *out,
                         int32_t order, int32_t n_threads, orc_stats *stats);
void orc_capsule_overlap_all(orc_world *w, const orc_capsule *q, int32_t n, int32_t max_hits,
                             orc_overlap_hit *out, int32_t *counts, uint8_t *overflow, int32_t order,
                             int32_t n_threads, orc_stats *stats);
void orc_move_and_slide(orc_world *w, orc_state *inout, int32_t n, const orc_params *params, float dt,
                        const float gravity[3], uint32_t flags, int32_t order, int32_t n_threads,
                        orc_stats *stats);

typedef struct orc_platform { float aabb_min[3], aabb_max[3], delta[3]; } orc_platform;
void orc_move_and_slide_ex(orc_world *w, orc_state *inout, int32_t n, const orc_params *params, float dt,
                           const float gravity[3], uint32_t flags, int32_t order, int32_t n_threads,
                           orc_stats *stats, const orc_platform *platforms, int32_t n_platforms);

/* narrow-phase primitives exposed for known-answer tests */
float orc_segment_triangle_distance(const float center[3], float half_height, const float v0[3],
                                    const float v1[3], const float v2[3], float seg_pt[3], float tri_pt[3]);
void orc_segment_triangle_distance_batch(int32_t n, const float *centers, const float *hh, const float *tris,
                                         float *dist, float *seg, float *tri);
void orc_ray_triangle_batch(int32_t n, const float *origins, const float *dirs, const float *tris, float *tout,
                            int32_t *hit);
/* capsuleCapsuleSweep (Systems.swift:1505-1590), packed; dims (n,4) = radius, halfHeight, otherRadius, otherHalfHeight */
void orc_capsule_capsule_sweep_batch(int32_t n, const float *from, const float *delta, const float *otherPos,
                                     const float *otherDelta, const float *dims, int32_t *hit, float *toi, float *normal);
float orc_closest_point_on_triangle(const float p[3], const float a[3], const float b[3],
                                    const float c[3], float out_pt[3]);
float orc_segment_segment_distance_sq(const float p1[3], const float q1[3], const float p2[3],
                                      const float q2[3], float c1[3], float c2[3]);

#ifdef __cplusplus
}
#endif
#endif
