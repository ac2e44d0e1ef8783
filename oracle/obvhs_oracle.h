/*
 * obvhs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (C++17, g++) of the obvhs 0.3.1 hot path:
 *   PLOC BVH2 build -> parallel reinsertion -> BVH2->CWBVH collapse -> CWBVH ray traversal.
 * It is the parity checker for the CUDA path in obvhs_b200/csrc and the "port" CPU baseline in bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (obvhs_b200) never links, imports or executes anything in this directory.
 *
 * Parity status: the Rust reference cannot be compiled in this image (no rustc/cargo). The oracle is pinned
 * against the reference's own known answers (tests/test_oracle_golden.py):
 *   - kitchen.obj 32x18 normal hash == 1343358762 for the fastest/fast/medium presets
 *     (examples/obj_cwbvh.rs:142-181),
 *   - icosphere(1)+PLANE closest hit primitive == 62 (src/cwbvh/traverse_macro.rs:33-56),
 *   - 4x4 flat plane, 256x256 rays all hit with normal +Y (tests/mod.rs:105-124),
 *   - degenerate builds traverse without hit (tests/mod.rs:35-102).
 * Those pin triangle intersection, traversal order and closest-hit results. Morton tie order, BVH2 topology and
 * CWBVH node bytes are NOT pinned by any reference test or fixture ("parity unpinned" for those: the reference's
 * own sort is unstable); there the oracle is a line-by-line restatement with the documented
 * deterministic tie rule (ties by ascending original index).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layouts mirror the reference's repr(C) Pod types byte for byte. */
typedef struct { float min[3]; float _p0; float max[3]; float _p1; } OrcAabb;           /* src/aabb.rs:11-16, 32 B  */
typedef struct { float v0[3]; float _p0; float v1[3]; float _p1; float v2[3]; float _p2; } OrcTriangle; /* src/triangle.rs:8-13, 48 B */
typedef struct { OrcAabb aabb; uint32_t prim_count, first_index, meta1, meta2; } OrcBvh2Node; /* src/bvh2/node.rs:40-66, 48 B */
typedef struct {
    float p[3]; uint8_t e[3]; uint8_t imask; uint32_t child_base_idx; uint32_t primitive_base_idx;
    uint8_t child_meta[8];
    uint8_t child_min_x[8], child_max_x[8], child_min_y[8], child_max_y[8], child_min_z[8], child_max_z[8];
} OrcCwBvhNode;                                                                          /* src/cwbvh/node.rs:12-54, 80 B */
typedef struct { float origin[3]; float _p0; float direction[3]; float _p1; float inv_direction[3]; float _p2;
                 float tmin, tmax; float _p3[2]; } OrcRay;                                /* src/ray.rs:15-30, 64 B */
typedef struct { uint32_t primitive_id, geometry_id, instance_id; float t; } OrcRayHit;   /* src/ray.rs:63-70, 16 B */

typedef struct OrcBvh2 OrcBvh2;
typedef struct OrcCwBvh OrcCwBvh;

/* -- geometry helpers -------------------------------------------------------------------------------- */
void orc_tri_aabbs(const OrcTriangle* tris, size_t n, OrcAabb* out);                    /* triangle.rs:28-30 */
void orc_make_rays(const float* origin_dir6, size_t n, float tmin, float tmax, OrcRay* out); /* ray.rs:34-52 */

/* -- PLOC -------------------------------------------------------------------------------------------- */
/* Morton stage only (ploc/mod.rs:187-243,287-288,771-827): codes (lo,hi 64-bit words per prim, in ORIGINAL order)
 * and the stable sorted order. precision = 64 or 128. total_aabb_out = scene AABB. */
void orc_morton_sort(const OrcAabb* aabbs, size_t n, int precision, uint64_t* codes_lo, uint64_t* codes_hi,
                     uint32_t* order_out, OrcAabb* total_aabb_out);
/* Full PLOC build (ploc/mod.rs:169-503). threads>1 uses OpenMP for the loops the reference runs under rayon. */
OrcBvh2* orc_ploc_build(const OrcAabb* aabbs, const uint32_t* indices, size_t n, uint32_t search_distance,
                        int precision, size_t search_depth_threshold, int threads);
void     orc_bvh2_free(OrcBvh2*);
size_t   orc_bvh2_node_count(const OrcBvh2*);
size_t   orc_bvh2_prim_count(const OrcBvh2*);
size_t   orc_bvh2_max_depth(const OrcBvh2*);
size_t   orc_bvh2_ploc_iterations(const OrcBvh2*);
void     orc_bvh2_get(const OrcBvh2*, OrcBvh2Node* nodes, uint32_t* primitive_indices, uint32_t* parents /*may be NULL*/);
OrcBvh2* orc_bvh2_from(const OrcBvh2Node* nodes, size_t n_nodes, const uint32_t* primitive_indices, size_t n_prims,
                       size_t max_depth);
/* returns 0 when valid, else a negative code; msg (>=256 B) receives a description (bvh2/mod.rs:786-981) */
int      orc_bvh2_validate(const OrcBvh2*, const OrcAabb* prim_aabbs, size_t n, int tight_fit, char* msg);
void     orc_bvh2_compute_parents(OrcBvh2*);                                            /* bvh2/mod.rs:586-619 */
void  orc_bvh2_reorder_in_stack_traversal_order(OrcBvh2*);                              /* bvh2/mod.rs:462-500 */
void     orc_bvh2_collapse(OrcBvh2*, uint32_t max_prims, float traversal_cost);             /* bvh2/leaf_collapser.rs:21-192 */
int      orc_bvh2_has_parents(const OrcBvh2*);
void     orc_bvh2_refit_all(OrcBvh2*);                                                  /* bvh2/mod.rs:527-569 */
/* PlocBuilder::full_rebuild / partial_rebuild, compute_rebuild_path_flags (ploc/rebuild.rs:12-183). should_remove and
 * flags_out hold one byte per node. Tie rule of partial rebuilds: equal Morton codes keep ascending node index. */
void     orc_ploc_full_rebuild(OrcBvh2*, uint32_t search_distance, int precision, size_t search_depth_threshold, int threads);
void     orc_ploc_partial_rebuild(OrcBvh2*, const uint8_t* should_remove, uint32_t search_distance, int precision,
                                  size_t search_depth_threshold, int threads);
void     orc_compute_rebuild_path_flags(const OrcBvh2*, const uint32_t* leaves, size_t n, uint8_t* flags_out);
void     orc_bvh2_set_node_aabbs(OrcBvh2*, const uint32_t* node_ids, const OrcAabb* aabbs, size_t n); /* Bvh2Node::set_aabb */
void     orc_bvh2_set_leaf_aabbs(OrcBvh2*, const OrcAabb* prim_aabbs);                  /* config 5 helper */

/* -- reinsertion (bvh2/reinsertion.rs:40-382) --------------------------------------------------------- */
void     orc_reinsertion_run(OrcBvh2*, float batch_size_ratio, const float* ratio_seq, size_t n_seq, int threads);
/* one batch, exposing the intermediate products for stage-by-stage parity */
void     orc_reinsertion_run_with_candidates(OrcBvh2*, const uint32_t* node_ids, size_t n, uint32_t iterations, int threads); /* reinsertion.rs:66-90 */
size_t   orc_reinsertion_last_applied(const OrcBvh2*);

/* -- BVH2 -> CWBVH (cwbvh/bvh2_to_cwbvh.rs:490-510) --------------------------------------------------- */
OrcCwBvh* orc_bvh2_to_cwbvh(const OrcBvh2*, uint32_t max_prims_per_leaf, int order_children, int include_exact_node_aabbs);
OrcCwBvh* orc_cwbvh_from(const OrcCwBvhNode* nodes, size_t n_nodes, const uint32_t* primitive_indices, size_t n_prims,
                         const OrcAabb* total_aabb);
void      orc_cwbvh_free(OrcCwBvh*);
size_t    orc_cwbvh_node_count(const OrcCwBvh*);
size_t    orc_cwbvh_prim_count(const OrcCwBvh*);
void      orc_cwbvh_get(const OrcCwBvh*, OrcCwBvhNode* nodes, uint32_t* primitive_indices, OrcAabb* total_aabb);
int       orc_cwbvh_validate(const OrcCwBvh*, const OrcAabb* prim_aabbs, size_t n, char* msg); /* cwbvh/mod.rs:747-908 */
size_t    orc_cwbvh_exact_node_aabbs(const OrcCwBvh*, OrcAabb* out, size_t cap);            /* cwbvh/mod.rs:47 (0 when absent) */

/* one-call builder (cwbvh/builder.rs:20-85). core_seconds mirrors core_build_time. */
OrcCwBvh* orc_build_cwbvh_from_tris(const OrcTriangle* tris, size_t n, uint32_t search_distance,
                                    size_t search_depth_threshold, float reinsertion_batch_ratio, int precision,
                                    uint32_t max_prims_per_leaf, int pre_split, int threads, double* core_seconds);

/* -- spatial pre-splits (splits.rs:16-158) ------------------------------------------------------------- */
/* split_aabbs_precise on caller arrays holding n entries with room for cap; returns the new count (> cap: nothing past
 * cap was written). */
size_t orc_split_aabbs_precise(OrcAabb* aabbs, uint32_t* indices, size_t n, size_t cap, const OrcTriangle* tris,
                               float area_thresh_low, float area_thresh_high, float split_factor_low, float split_factor_high,
                               uint32_t max_iterations, uint32_t split_tests);
/* the builders' prologue (cwbvh/builder.rs:27-54): triangle AABBs, sequential-f32 average and max half area, then
 * split_aabbs_preset; returns the split count */
size_t orc_presplit_tris(const OrcTriangle* tris, size_t n, OrcAabb* aabbs_out, uint32_t* indices_out, size_t cap,
                         float* avg_half_area, float* largest_half_area);
void   orc_bvh2_set_uses_spatial_splits(OrcBvh2*, int);   /* bvh2/mod.rs:84: relaxes validate() as the reference does */
void   orc_cwbvh_set_uses_spatial_splits(OrcCwBvh*, int); /* cwbvh/mod.rs:54 */

/* -- traversal (cwbvh/mod.rs:169-245, traverse_macro.rs:59-126, node.rs:86-231, simd.rs:17-100) -------- */
/* bvh_tris are the triangles pre-permuted by primitive_indices (examples/obj_cwbvh.rs:63-67).
 * counters (optional, 2 x u64): [0] += nodes visited, [1] += triangles tested, summed over all rays.
 * use_simd: 1 = SSE2 4-wide node test (simd.rs), 0 = scalar node test (node.rs intersect_ray_basic). */
void orc_cwbvh_ray_traverse(const OrcCwBvh*, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n,
                            OrcRayHit* hits, int threads, int use_simd, uint64_t* counters);
void orc_cwbvh_ray_traverse_miss(const OrcCwBvh*, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n,
                                 uint8_t* miss, int threads, int use_simd, uint64_t* counters);
/* all-hit variant (cwbvh/mod.rs:233-245): counts intersections with t < +inf per ray */
void orc_cwbvh_ray_traverse_anyhit_count(const OrcCwBvh*, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n,
                                         uint32_t* counts, int threads);
/* Bvh2::ray_traverse / ray_traverse_miss / counting ray_traverse_anyhit (bvh2/mod.rs:148-334) over triangles permuted by
 * primitive_indices; counters[0] += node AABB tests, counters[1] += triangle tests */
void orc_bvh2_ray_traverse(const OrcBvh2*, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, OrcRayHit* hits, int threads,
                           uint64_t* counters);
void orc_bvh2_ray_traverse_miss(const OrcBvh2*, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, uint8_t* miss, int threads,
                                uint64_t* counters);
void orc_bvh2_ray_traverse_anyhit_count(const OrcBvh2*, const OrcTriangle* bvh_tris, const OrcRay* rays, size_t n, uint32_t* counts,
                                        int threads);
/* build_bvh2_from_tris (bvh2/builder.rs:17-91) */
OrcBvh2* orc_build_bvh2_from_tris(const OrcTriangle* tris, size_t n, uint32_t search_distance, size_t search_depth_threshold,
                                  float reinsertion_batch_ratio, float post_collapse_multiplier, int precision,
                                  uint32_t max_prims_per_leaf, float collapse_traversal_cost, int pre_split, int threads,
                                  double* core_seconds);
/* Broad-phase queries with an `eval` that always continues: Bvh2::aabb_traverse / point_traverse (bvh2/mod.rs:365-456) report
 * leaf NODE ids; the traverse! macro over CwBvhNode::intersect_aabb / contains_point (cwbvh/node.rs:157-200,
 * traverse_macro.rs:59-126) reports primitive slots (state.primitive_id), children ordered by dir3 (new_traversal).
 * counts[i] = reports of query i; ids receive them query after query in call order, up to cap. Returns the total.
 * OrcAabb queries use the min/max lanes; points are Vec3A (4 floats each). */
size_t orc_bvh2_aabb_traverse(const OrcBvh2*, const OrcAabb* queries, size_t n, uint32_t* counts, uint32_t* leaf_ids, size_t cap);
size_t orc_bvh2_point_traverse(const OrcBvh2*, const float* points4, size_t n, uint32_t* counts, uint32_t* leaf_ids, size_t cap);
size_t orc_cwbvh_aabb_traverse(const OrcCwBvh*, const OrcAabb* queries, size_t n, const float* dir3, uint32_t* counts,
                               uint32_t* prim_ids, size_t cap);
size_t orc_cwbvh_point_traverse(const OrcCwBvh*, const float* points4, size_t n, const float* dir3, uint32_t* counts,
                                uint32_t* prim_ids, size_t cap);
float orc_triangle_intersect(const OrcTriangle* tri, const OrcRay* ray);               /* triangle.rs:35-76 */
void  orc_triangle_normal(const OrcTriangle* tri, float* out3);                         /* triangle.rs:20-24 */
void  orc_cwbvh_order_node_children(OrcCwBvh*, const OrcAabb* prim_aabbs, size_t node_index, int direct_layout); /* cwbvh/mod.rs:538-735 */
void  orc_cwbvh_order_children(OrcCwBvh*, const OrcAabb* prim_aabbs, int direct_layout);                          /* cwbvh/mod.rs:520-524 */
float orc_aabb_half_area(const OrcAabb* a);                                              /* aabb.rs:151-154 */
void  orc_aabb_union(const OrcAabb* a, const OrcAabb* b, OrcAabb* out);                 /* aabb.rs:84-89 */
float orc_aabb_intersect_ray(const OrcAabb* a, const OrcRay* ray);                      /* aabb.rs:186-206 */
int   orc_aabb_intersect_aabb(const OrcAabb* a, const OrcAabb* b);                      /* aabb.rs:181-183 */
int   orc_aabb_contains_point(const OrcAabb* a, const float* p3);                       /* aabb.rs:70-72 */
int   orc_max_threads(void);

#ifdef __cplusplus
}
#endif
