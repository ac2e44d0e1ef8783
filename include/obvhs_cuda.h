/*
 * obvhs_cuda.h -- C ABI of libobvhs_cuda (sm_100a), the B200-native drop-in for the data-parallel hot path of
 * obvhs 0.3.1: PLOC BVH2 build -> parallel reinsertion -> BVH2->CWBVH collapse -> CWBVH ray traversal.
 *
 * The reference (pure Rust) has no FFI layer; each entry point below names the Rust interface it replaces
 * (paths relative to the reference checkout). INTEGRATION.md shows the `extern "C"` block + safe wrappers a
 * maintainer adds to obvhs to route these calls here.
 *
 * Conventions
 *   - every call returns int: 0 = OBVHS_OK, <0 = ObvhsStatus error. Nothing throws across the boundary;
 *     obvhs_cuda_last_error() returns a description of the last failure on that context.
 *   - POD structs are byte-identical to the reference's #[repr(C)] Pod types (sizes static-asserted in the library).
 *   - data pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes). Host buffers are staged
 *     through the context's stream; device buffers are used in place.
 *   - a context owns one CUDA stream (its own, or the caller's) and the reusable scratch of the builder
 *     (the reference's PlocBuilder keeps current_nodes/next_nodes/mortons for reuse, ploc/mod.rs:35-45).
 *     One context is thread-compatible; distinct contexts are independent.
 *   - there is NO CPU fallback: without a CUDA device obvhs_cuda_create fails with OBVHS_ERR_CUDA.
 */
#ifndef OBVHS_CUDA_H
#define OBVHS_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    OBVHS_OK = 0,
    OBVHS_ERR_INVALID_ARG = -1,
    OBVHS_ERR_CUDA = -2,
    OBVHS_ERR_UNSUPPORTED = -3, /* depth beyond the fixed stacks, >= 2^30 primitives */
    OBVHS_ERR_NAN_INPUT = -4,   /* the reference panics / goes out of bounds on NaN AABBs (ploc/mod.rs:451) */
    OBVHS_ERR_STACK_OVERFLOW = -5,
    OBVHS_ERR_CAPACITY = -6,    /* a growing output (Vec::push in the reference) does not fit the caller's arrays */
    OBVHS_ERR_NCCL = -7         /* NCCL missing at run time, or a collective failed */
} ObvhsStatus;

/* src/aabb.rs:11-16 -- two Vec3A lanes; the 4th float of each lane is padding (never read, never compared). */
typedef struct { float min[3]; float _pad0; float max[3]; float _pad1; } ObvhsAabb;
/* src/triangle.rs:8-13 */
typedef struct { float v0[3]; float _pad0; float v1[3]; float _pad1; float v2[3]; float _pad2; } ObvhsTriangle;
/* src/bvh2/node.rs:40-66 (default 48-byte layout). prim_count == 0 => inner node, children at first_index, +1. */
typedef struct { ObvhsAabb aabb; uint32_t prim_count; uint32_t first_index; uint32_t meta1; uint32_t meta2; } ObvhsBvh2Node;
/* src/cwbvh/node.rs:12-54 -- 80 bytes, no padding */
typedef struct {
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base_idx;
    uint32_t primitive_base_idx;
    uint8_t child_meta[8];
    uint8_t child_min_x[8], child_max_x[8], child_min_y[8], child_max_y[8], child_min_z[8], child_max_z[8];
} ObvhsCwBvhNode;
/* src/ray.rs:15-30 */
typedef struct {
    float origin[3]; float _pad0; float direction[3]; float _pad1; float inv_direction[3]; float _pad2;
    float tmin, tmax; float _pad3[2];
} ObvhsRay;
/* The arguments of Ray::new(origin, direction, min, max) (src/ray.rs:34-52), 32 bytes: what a caller holds BEFORE the
 * constructor fills inv_direction. The *_ray_new_* entry points run the constructor on the device (safe_inverse,
 * ray.rs:6-12, bit for bit), so a host batch crosses PCIe at half the size of the Ray array. */
typedef struct { float origin[3]; float tmin; float direction[3]; float tmax; } ObvhsRayNew;
/* Origin and direction alone, 24 bytes: the per-ray part of Ray::new_inf(origin, direction) (src/ray.rs:55-57), which is how the
 * camera, bounce and shadow loops of the reference's examples construct every ray (examples/demoscene.rs:152,178,189,213), and of
 * Ray::new(eye, direction, 0.0, f32::MAX) (examples/cornell_box_cwbvh.rs:116, obj_cwbvh.rs:104). The *_ray_od_* entry points take
 * ONE (tmin, tmax) for the batch and run the constructor on the device: 24 instead of 64 bytes per ray cross PCIe. */
typedef struct { float origin[3]; float direction[3]; } ObvhsRayOd;
/* src/ray.rs:63-70; RayHit::none() = ids 0xffffffff, t = +inf */
typedef struct { uint32_t primitive_id, geometry_id, instance_id; float t; } ObvhsRayHit;

/* What CwBvh::ray_traverse with the triangle closure writes into a RayHit (cwbvh/mod.rs:184-189): primitive_id and t. geometry_id and
 * instance_id stay RayHit::none()'s INVALID on this path, so a host batch can bring back 8 instead of 16 bytes per ray. */
typedef struct { uint32_t primitive_id; float t; } ObvhsRayHit8;

/* src/lib.rs:208-231 BvhBuildParams, field for field. ploc_search_distance is the u32 form of PlocSearchDistance
 * (ploc/mod.rs:534-562: 1,2,6,14,24,32); sort_precision is 64 or 128 (ploc/mod.rs:658-661). */
typedef struct {
    uint32_t pre_split; /* spatial pre-splits of large triangles (src/splits.rs) before PLOC; triangle builders only */
    uint32_t ploc_search_distance;
    uint64_t search_depth_threshold;
    float reinsertion_batch_ratio;
    float post_collapse_reinsertion_batch_ratio_multiplier; /* Bvh2-only in the reference; ignored for CwBvh */
    uint32_t sort_precision;
    uint32_t max_prims_per_leaf; /* CwBvh builders clamp to 1..3 (cwbvh/builder.rs:74) */
    float collapse_traversal_cost; /* Bvh2-only; ignored */
} ObvhsBuildParams;

typedef struct ObvhsContext ObvhsContext;
typedef struct ObvhsBvh2 ObvhsBvh2;   /* device-resident Bvh2 (src/bvh2/mod.rs:31-85) */
typedef struct ObvhsCwBvh ObvhsCwBvh; /* device-resident CwBvh (src/cwbvh/mod.rs:43-55) + optional permuted triangles */

/* ---- context ------------------------------------------------------------------------------------------ */
/* stream: a cudaStream_t the caller owns, or NULL to let the context create one. */
int obvhs_cuda_create(int device, void* stream, ObvhsContext** out);
void obvhs_cuda_destroy(ObvhsContext* ctx);
const char* obvhs_cuda_last_error(const ObvhsContext* ctx);
int obvhs_cuda_synchronize(ObvhsContext* ctx);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
uint64_t obvhs_cuda_launch_count(const ObvhsContext* ctx);
/* Tuning knobs that do not change any result. key "traverse": "auto" (default: every 128-ray block of the batch is
 * judged on the device and goes to the kernel that suits it), "static" (one ray per thread) or "persistent[:refill[:chunk]]" (persistent warps refilled from a ray cursor when
 * `refill` of 32 lanes have finished, `chunk` consecutive rays per fetch). key "host_slice": rays per pipelined slice of a host-resident ray batch ("0" = sized for the kernel
 * in use). key "trace": "1"/"0" stage timing on stderr
 * (the reference's scope!/timeit! macros, lib.rs:158-205). key "traverse_variant": "<id>" picks the persistent kernel's
 * scheduling / stack / register-cap variant (traverse.cu; 0 = default). Environment: OBVHS_TRAVERSE, OBVHS_TRAVERSE_VARIANT, OBVHS_TRACE set the defaults. */
int obvhs_cuda_set_option(ObvhsContext* ctx, const char* key, const char* value);
/* 6 built-in presets of src/lib.rs:233-305 by name: fastest_build, very_fast_build, fast_build, medium_build,
 * slow_build, very_slow_build. */
int obvhs_cuda_build_params_preset(const char* name, ObvhsBuildParams* out);

/* ---- spatial pre-splits (src/splits.rs) ------------------------------------------------------------------ */
/* split_aabbs_precise(&mut aabbs, &mut indices, triangles, area_thresh_low, area_thresh_high, split_factor_low,
 * split_factor_high, max_iterations, split_tests)  (splits.rs:49-125). aabbs / indices hold n entries and have room
 * for `capacity`; entries may shrink in place and the right halves are appended in the reference's order.
 * *count_out receives the new length; when it exceeds capacity nothing is written and OBVHS_ERR_CAPACITY is returned
 * (the Vec would have grown: call again with at least *count_out entries of room). */
int obvhs_cuda_split_aabbs_precise(ObvhsContext* ctx, ObvhsAabb* aabbs, uint32_t* indices, size_t n, size_t capacity,
                                   const ObvhsTriangle* tris, size_t n_tris, float area_thresh_low, float area_thresh_high,
                                   float split_factor_low, float split_factor_high, uint32_t max_iterations,
                                   uint32_t split_tests, size_t* count_out);
/* split_aabbs_preset(&mut aabbs, &mut indices, triangles, avg_half_area, largest_half_area)  (splits.rs:16-34) */
int obvhs_cuda_split_aabbs_preset(ObvhsContext* ctx, ObvhsAabb* aabbs, uint32_t* indices, size_t n, size_t capacity,
                                  const ObvhsTriangle* tris, size_t n_tris, float avg_half_area, float largest_half_area,
                                  size_t* count_out);
/* The pre-split prologue of build_cwbvh_from_tris / build_bvh2_from_tris (cwbvh/builder.rs:27-54): triangle AABBs,
 * their average (sequential f32 sum, as the reference) and largest half area, then split_aabbs_preset. aabbs_out /
 * indices_out may be NULL to query the count. */
int obvhs_cuda_presplit_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, ObvhsAabb* aabbs_out,
                             uint32_t* indices_out, size_t capacity, size_t* count_out, float* avg_half_area,
                             float* largest_half_area);
/* Bvh2::uses_spatial_splits / CwBvh::uses_spatial_splits (bvh2/mod.rs:84, cwbvh/mod.rs:54): primitive_indices may name a
 * triangle several times, so set_triangles accepts fewer triangles than primitives. Set by the builders. */
int obvhs_cuda_bvh2_uses_spatial_splits(const ObvhsBvh2* bvh);
void obvhs_cuda_bvh2_set_uses_spatial_splits(ObvhsBvh2* bvh, int value);
int obvhs_cuda_cwbvh_uses_spatial_splits(const ObvhsCwBvh* bvh);
void obvhs_cuda_cwbvh_set_uses_spatial_splits(ObvhsCwBvh* bvh, int value);

/* ---- PLOC (src/ploc/mod.rs) ----------------------------------------------------------------------------- */
/* Stage probe for parity tests: leaf init + scene AABB (ploc/mod.rs:187-243), Morton codes (:287-288,782-785;
 * morton.rs:35-89) in ORIGINAL order and the sorted order (:811-827, ties by ascending index). Any output may be
 * NULL. precision 64: codes_hi is zero filled. */
int obvhs_cuda_morton_sort(ObvhsContext* ctx, const ObvhsAabb* aabbs, size_t n, uint32_t sort_precision,
                           uint64_t* codes_lo, uint64_t* codes_hi, uint32_t* order, ObvhsAabb* total_aabb);
/* PlocBuilder::build(search_distance, aabbs, indices, sort_precision, search_depth_threshold) -> Bvh2
 * (ploc/mod.rs:95-102). indices == NULL means 0..n. */
int obvhs_cuda_ploc_build(ObvhsContext* ctx, const ObvhsAabb* aabbs, const uint32_t* indices, size_t n,
                          uint32_t search_distance, uint32_t sort_precision, size_t search_depth_threshold,
                          ObvhsBvh2** out);
/* Same over &[Triangle] (Boundable for Triangle, triangle.rs:28-30): the AABB is computed inside leaf init. */
int obvhs_cuda_ploc_build_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, uint32_t search_distance,
                               uint32_t sort_precision, size_t search_depth_threshold, ObvhsBvh2** out);

/* PlocBuilder::full_rebuild(&mut bvh, search_distance, sort_precision, search_depth_threshold)  (ploc/rebuild.rs:56-80):
 * rebuilds the tree in place from its current leaves (inner nodes are ignored; multi-primitive leaves stay as they are). */
int obvhs_cuda_ploc_full_rebuild(ObvhsContext* ctx, ObvhsBvh2* bvh, uint32_t search_distance, uint32_t sort_precision,
                                 size_t search_depth_threshold);
/* PlocBuilder::partial_rebuild(&mut bvh, should_remove, ...)  (ploc/rebuild.rs:101-135). should_remove holds one byte per
 * node (the reference takes a closure Fn(usize) -> bool; compute_rebuild_path_flags produces exactly this array). Flagged
 * chains are dissolved, the untouched subtrees and the flagged leaves are merged again by PLOC, and the new inner nodes
 * reuse the freed slot pairs from the end of the node array downwards (ploc/mod.rs:449-462). Tie rule: collected nodes
 * with equal Morton codes keep ascending node index. */
int obvhs_cuda_ploc_partial_rebuild(ObvhsContext* ctx, ObvhsBvh2* bvh, const uint8_t* should_remove, uint32_t search_distance,
                                    uint32_t sort_precision, size_t search_depth_threshold);
/* compute_rebuild_path_flags(&bvh, leaves, &mut flags)  (ploc/rebuild.rs:12-43): flags[i] = 1 for the given leaf nodes and
 * every ancestor. Needs parents (obvhs_cuda_bvh2_compute_parents) like the reference, else OBVHS_ERR_INVALID_ARG.
 * flags: node_count bytes, host or device. */
int obvhs_cuda_compute_rebuild_path_flags(ObvhsContext* ctx, const ObvhsBvh2* bvh, const uint32_t* leaves, size_t n_leaves,
                                          uint8_t* flags);
/* Bvh2Node::set_aabb (bvh2/node.rs) for n nodes: bvh.nodes[node_ids[k]].set_aabb(aabbs[k]) (examples/physics.rs:446) */
int obvhs_cuda_bvh2_set_node_aabbs(ObvhsContext* ctx, ObvhsBvh2* bvh, const uint32_t* node_ids, const ObvhsAabb* aabbs, size_t n);

/* ---- Bvh2 (src/bvh2/mod.rs) ------------------------------------------------------------------------------ */
void obvhs_cuda_bvh2_free(ObvhsBvh2* bvh);
size_t obvhs_cuda_bvh2_node_count(const ObvhsBvh2* bvh);
size_t obvhs_cuda_bvh2_prim_count(const ObvhsBvh2* bvh);
size_t obvhs_cuda_bvh2_max_depth(const ObvhsBvh2* bvh);       /* Bvh2::max_depth, ploc/mod.rs:501 */
size_t obvhs_cuda_bvh2_ploc_iterations(const ObvhsBvh2* bvh); /* `depth` at loop exit, ploc/mod.rs:327-493 */
int obvhs_cuda_bvh2_children_ordered_after_parents(const ObvhsBvh2* bvh);
/* nodes: node_count x 48 B; primitive_indices: prim_count x u32; parents: node_count x u32 (needs parents computed).
 * Any may be NULL. */
int obvhs_cuda_bvh2_download(ObvhsContext* ctx, const ObvhsBvh2* bvh, ObvhsBvh2Node* nodes, uint32_t* primitive_indices,
                             uint32_t* parents);
int obvhs_cuda_bvh2_upload(ObvhsContext* ctx, const ObvhsBvh2Node* nodes, size_t node_count,
                           const uint32_t* primitive_indices, size_t prim_count, size_t max_depth,
                           int children_ordered_after_parents, ObvhsBvh2** out);
int obvhs_cuda_bvh2_compute_parents(ObvhsContext* ctx, ObvhsBvh2* bvh);                 /* bvh2/mod.rs:586-619 */
int obvhs_cuda_bvh2_refit_all(ObvhsContext* ctx, ObvhsBvh2* bvh);
/* Bvh2::reorder_in_stack_traversal_order (src/bvh2/mod.rs:462-500): nodes re-indexed in the pop order of the reference's stack
 * (parents before children, each sibling pair followed by the subtree of its second node, then of its first); parents are
 * recomputed when present; children_are_ordered_after_parents becomes true. */
int obvhs_cuda_bvh2_reorder_in_stack_traversal_order(ObvhsContext* ctx, ObvhsBvh2* bvh);                       /* bvh2/mod.rs:462-500 */
/* rewrite every leaf's AABB from per-primitive AABBs (dynamic scenes, examples/physics.rs:500-539), then refit_all */
int obvhs_cuda_bvh2_set_leaf_aabbs(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsAabb* prim_aabbs, size_t n);

/* ReinsertionOptimizer::run(&mut bvh, batch_size_ratio, ratio_sequence) (bvh2/reinsertion.rs:40-57,92-111).
 * ratio_sequence == NULL selects the default 1/1, 1/3, ... 1/31. applied_out (optional) receives the number of
 * reinsertions applied. */
int obvhs_cuda_reinsertion_run(ObvhsContext* ctx, ObvhsBvh2* bvh, float batch_size_ratio, const float* ratio_sequence,
                               size_t n_sequence, uint64_t* applied_out);

/* ReinsertionOptimizer::run_with_candidates(&mut bvh, candidates, iterations) (bvh2/reinsertion.rs:66-90,113-118): the
 * given node ids (each in [1, node_count), host or device memory), in the given order, are the candidates of every one
 * of `iterations` rounds. The root or an out-of-range id is OBVHS_ERR_INVALID_ARG (the reference panics). */
int obvhs_cuda_reinsertion_run_with_candidates(ObvhsContext* ctx, ObvhsBvh2* bvh, const uint32_t* node_ids, size_t n,
                                               uint32_t iterations, uint64_t* applied_out);

/* collapse(&mut bvh, max_prims, traversal_cost) (bvh2/leaf_collapser.rs:21-192): SAH leaf collapse; nodes and
 * primitive_indices are rewritten in the reference's order, parents are recomputed only if they existed. */
int obvhs_cuda_bvh2_collapse(ObvhsContext* ctx, ObvhsBvh2* bvh, uint32_t max_prims, float traversal_cost);
/* build_bvh2_from_tris(triangles, config, core_build_time) (bvh2/builder.rs:17-91): PLOC -> reinsertion -> collapse ->
 * reinsertion(ratio * post_collapse multiplier). The permuted triangles are attached to the result. */
int obvhs_cuda_build_bvh2_from_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, const ObvhsBuildParams* params,
                                    double* core_build_seconds, ObvhsBvh2** out);
/* build_bvh2<T: Boundable>(primitives, config, core_build_time) -> Bvh2  (src/bvh2/builder.rs:103-140) over AABBs:
 * PLOC -> reinsertion -> SAH leaf collapse -> reinsertion. config.pre_split is ignored, as in the reference. */
int obvhs_cuda_build_bvh2(ObvhsContext* ctx, const ObvhsAabb* aabbs, size_t n, const ObvhsBuildParams* params,
                          double* core_build_seconds, ObvhsBvh2** out);
/* bvh_tris[i] = tris[primitive_indices[i]] kept on the device inside the handle (examples/demoscene.rs:66-70) */
int obvhs_cuda_bvh2_set_triangles(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsTriangle* tris, size_t n);
/* Batched Bvh2::ray_traverse / ray_traverse_miss / counting ray_traverse_anyhit over triangles (bvh2/mod.rs:148-334,
 * aabb.rs:186-206, triangle.rs:35-76). hit.primitive_id indexes primitive_indices order; as in the reference a ray that
 * exhausts its stack without a hit leaves hit.t = ray.tmax (bvh2/mod.rs:326), a ray that misses the root leaves +inf.
 * counters (host or device, 2 x u64): [0] += node AABB tests, [1] += triangle tests. */
int obvhs_cuda_bvh2_ray_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n, ObvhsRayHit* hits);
int obvhs_cuda_bvh2_ray_traverse_miss_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n, uint8_t* miss);
int obvhs_cuda_bvh2_ray_traverse_anyhit_count_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n,
                                                    uint32_t* counts);
int obvhs_cuda_bvh2_ray_traverse_batch_counted(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRay* rays, size_t n,
                                               ObvhsRayHit* hits, uint64_t* counters);

/* ---- CwBvh (src/cwbvh) ------------------------------------------------------------------------------------ */
/* bvh2_to_cwbvh(&bvh2, max_prims_per_leaf, order_children, include_exact_node_aabbs) (bvh2_to_cwbvh.rs:490-510).
 * include_exact_node_aabbs != 0 also fills CwBvh::exact_node_aabbs (see obvhs_cuda_cwbvh_exact_node_aabbs). */
int obvhs_cuda_bvh2_to_cwbvh(ObvhsContext* ctx, const ObvhsBvh2* bvh, uint32_t max_prims_per_leaf, int order_children,
                             int include_exact_node_aabbs, ObvhsCwBvh** out);
/* CwBvh::exact_node_aabbs (src/cwbvh/mod.rs:47; filled when bvh2_to_cwbvh was called with include_exact_node_aabbs,
 * bvh2_to_cwbvh.rs:60-80): one Aabb per Bvh2 node slot, entry i < node_count = the unquantised box of wide node i, the rest
 * Aabb::empty(). *count receives the number of entries (0 when absent); out may be NULL to query it. */
int obvhs_cuda_cwbvh_exact_node_aabbs(ObvhsContext* ctx, const ObvhsCwBvh* bvh, ObvhsAabb* out, size_t capacity, size_t* count);
/* CwBvh::compute_parents (src/cwbvh/mod.rs:494-509): parents[node_index] = the node whose inner child slot references it,
 * parents[0] = 0. `parents` holds node_count entries (host or device). */
int obvhs_cuda_cwbvh_compute_parents(ObvhsContext* ctx, const ObvhsCwBvh* bvh, uint32_t* parents);
/* CwBvh::order_children(&mut self, primitives, direct_layout) (src/cwbvh/mod.rs:520-735): every node's children re-assigned to
 * the octant slots by the greedy cost table over child centres (leaf children: union of their primitives' boxes; inner children:
 * the child's own box, exact when the tree carries exact_node_aabbs), inner children's records moved inside their range.
 * prim_aabbs: the primitives as Boundable::aabb(), indexed by primitive id (direct_layout = 0) or already in
 * primitive_indices order (direct_layout != 0). Note: permuted triangles attached to the handle stay valid (primitive order
 * does not change). */
int obvhs_cuda_cwbvh_order_children(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsAabb* prim_aabbs, size_t n, int direct_layout);
/* build_cwbvh_from_tris(triangles, config, core_build_time) (cwbvh/builder.rs:20-85). core_build_seconds (optional)
 * is INCREMENTED by the device time of PLOC -> reinsertion -> collapse, as the reference's `+=` does. The permuted
 * triangle array (examples/obj_cwbvh.rs:63-67) is attached to the result so it can be traversed directly. */
int obvhs_cuda_build_cwbvh_from_tris(ObvhsContext* ctx, const ObvhsTriangle* tris, size_t n, const ObvhsBuildParams* params,
                                     double* core_build_seconds, ObvhsCwBvh** out);
/* build_cwbvh<T: Boundable>(primitives, config, core_build_time) -> CwBvh  (src/cwbvh/builder.rs:98-123) over AABBs:
 * PLOC -> reinsertion -> collapse. config.pre_split is ignored, as in the reference. The handle carries no triangles. */
int obvhs_cuda_build_cwbvh(ObvhsContext* ctx, const ObvhsAabb* aabbs, size_t n, const ObvhsBuildParams* params,
                           double* core_build_seconds, ObvhsCwBvh** out);
void obvhs_cuda_cwbvh_free(ObvhsCwBvh* bvh);
size_t obvhs_cuda_cwbvh_node_count(const ObvhsCwBvh* bvh);
size_t obvhs_cuda_cwbvh_prim_count(const ObvhsCwBvh* bvh);
int obvhs_cuda_cwbvh_download(ObvhsContext* ctx, const ObvhsCwBvh* bvh, ObvhsCwBvhNode* nodes, uint32_t* primitive_indices,
                              ObvhsAabb* total_aabb);
int obvhs_cuda_cwbvh_upload(ObvhsContext* ctx, const ObvhsCwBvhNode* nodes, size_t node_count,
                            const uint32_t* primitive_indices, size_t prim_count, const ObvhsAabb* total_aabb,
                            ObvhsCwBvh** out);
/* bvh_tris[i] = tris[primitive_indices[i]] (examples/obj_cwbvh.rs:63-67), kept on the device inside the handle. */
int obvhs_cuda_cwbvh_set_triangles(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsTriangle* tris, size_t n);
/* Bytes per primitive of the handle's internal triangle buffer (the third pointer of obvhs_cuda_cwbvh_device_ptrs): the
 * triangles permuted by primitive_indices are kept as the reference's RtTriangle {v0, e1, e2, ng} (src/rt_triangle.rs:160-183,
 * 64 bytes), whose intersect() returns exactly what Triangle::intersect() does. */
size_t obvhs_cuda_cwbvh_triangle_bytes(void);
/* device addresses of the buffers (for NCCL broadcast of a finished tree); bvh_tris is NULL until set. */
int obvhs_cuda_cwbvh_device_ptrs(const ObvhsCwBvh* bvh, void** nodes, void** primitive_indices, void** bvh_tris);
/* empty device-resident CwBvh with the given sizes, to receive a broadcast */
int obvhs_cuda_cwbvh_alloc(ObvhsContext* ctx, size_t node_count, size_t prim_count, int with_triangles,
                           const ObvhsAabb* total_aabb, ObvhsCwBvh** out);

/* Batched CwBvh::ray_traverse(ray, &mut hit, |ray, id| bvh_tris[id].intersect(ray)) (cwbvh/mod.rs:169-198,
 * traverse_macro.rs:59-126, triangle.rs:35-76): hits[i] starts as RayHit::none(); hit.primitive_id indexes
 * primitive_indices order (NOT the original triangle id), exactly like the reference. */
int obvhs_cuda_cwbvh_ray_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n,
                                        ObvhsRayHit* hits);
/* Batched CwBvh::ray_traverse_miss (cwbvh/mod.rs:201-225): miss[i] = 1 when nothing is hit before ray.tmax. */
int obvhs_cuda_cwbvh_ray_traverse_miss_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n,
                                             uint8_t* miss);
/* Batched CwBvh::ray_traverse_anyhit (cwbvh/mod.rs:233-245) with a counting closure: counts[i] = number of
 * triangles the ray intersects (t < +inf) among those the node tests let through. */
int obvhs_cuda_cwbvh_ray_traverse_anyhit_count_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays,
                                                     size_t n, uint32_t* counts);
/* Same closest-hit traversal, also accumulating the per-launch totals the roofline model needs:
 * counters[0] += nodes visited, counters[1] += triangles tested (host or device pointer to 2 x u64). */
int obvhs_cuda_cwbvh_ray_traverse_batch_counted(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRay* rays, size_t n,
                                                ObvhsRayHit* hits, uint64_t* counters);
int obvhs_cuda_cwbvh_ray_new_traverse_batch_counted(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n,
                                                    ObvhsRayHit* hits, uint64_t* counters);
/* Ray::new (src/ray.rs:34-52) for n argument records -> n Ray structs (host or device pointers). */
int obvhs_cuda_ray_new_batch(ObvhsContext* ctx, const ObvhsRayNew* args, size_t n, ObvhsRay* rays);
/* ray_traverse / ray_traverse_miss / ray_traverse_anyhit of rays[i] = Ray::new(args[i]): identical results to the *_batch
 * calls above on obvhs_cuda_ray_new_batch's output; the constructor runs on the device chunk by chunk inside the pipelined
 * H2D -> traversal -> D2H of a host batch. */
int obvhs_cuda_cwbvh_ray_new_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n,
                                            ObvhsRayHit* hits);
int obvhs_cuda_cwbvh_ray_new_traverse_miss_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args, size_t n,
                                                 uint8_t* miss);
int obvhs_cuda_cwbvh_ray_new_traverse_anyhit_count_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayNew* args,
                                                         size_t n, uint32_t* counts);
int obvhs_cuda_bvh2_ray_new_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRayNew* args, size_t n,
                                           ObvhsRayHit* hits);
int obvhs_cuda_bvh2_ray_new_traverse_miss_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRayNew* args, size_t n,
                                                uint8_t* miss);
/* The same for rays[i] = Ray::new(od[i].origin, od[i].direction, tmin, tmax) -- tmin = 0, tmax = INFINITY is Ray::new_inf
 * (src/ray.rs:55-57). Bit-identical results to the calls above on the expanded rays. */
int obvhs_cuda_cwbvh_ray_od_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayOd* od, size_t n, float tmin, float tmax,
                                           ObvhsRayHit* hits);
/* the same closest hits as {primitive_id, t} records: hits[i] == {h.primitive_id, h.t} of obvhs_cuda_cwbvh_ray_od_traverse_batch's h */
int obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayOd* od, size_t n, float tmin, float tmax,
                                                ObvhsRayHit8* hits);
int obvhs_cuda_cwbvh_ray_od_traverse_miss_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsRayOd* od, size_t n, float tmin,
                                                float tmax, uint8_t* miss);
int obvhs_cuda_bvh2_ray_od_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsRayOd* od, size_t n, float tmin, float tmax,
                                          ObvhsRayHit* hits);
/* ---- multi-GPU: replicate a finished tree -------------------------------------------------------------------
 * The reference's CwBvh is a plain Clone of three Vecs and an Aabb (src/cwbvh/mod.rs:43-55) and ray_traverse only reads &self
 * (:169): the build runs on ONE GPU, the finished tree is broadcast over NVLink / NVSwitch with NCCL, and every rank traverses
 * its own range of rays against its replica (no collective on the traversal path). One process (or thread) per GPU, one
 * context per rank. NCCL is bound at run time (dlopen libnccl.so.2, override with OBVHS_NCCL_LIB); OBVHS_ERR_NCCL without it.
 *   obvhs_cuda_nccl_unique_id : on ONE rank; the caller ships the 128 bytes to the others (file, socket, MPI, torch.distributed)
 *   obvhs_cuda_comm_init      : collective over `world` ranks; binds an ncclComm_t to the context (freed with it)
 *   obvhs_cuda_cwbvh_broadcast: collective; *bvh is the finished tree on `root`; elsewhere NULL or a handle from an earlier
 *                               broadcast (its buffers are reused when the sizes match) and receives the replica: nodes,
 *                               primitive_indices, total_aabb, flags and the permuted triangles when the root has them.
 *                               A 64-byte header broadcast, then ONE grouped NCCL launch, all on the context's stream; the
 *                               receivers make one host round trip (the header), the root none. */
#define OBVHS_NCCL_UNIQUE_ID_BYTES 128
int obvhs_cuda_nccl_unique_id(uint8_t id[OBVHS_NCCL_UNIQUE_ID_BYTES]);
int obvhs_cuda_comm_init(ObvhsContext* ctx, const uint8_t id[OBVHS_NCCL_UNIQUE_ID_BYTES], int rank, int world);
int obvhs_cuda_cwbvh_broadcast(ObvhsContext* ctx, ObvhsCwBvh** bvh, int root);

/* ---- broad-phase queries (batched; the per-report closure of the reference becomes a list of reports) ---------------
 * Bvh2::aabb_traverse(aabb, eval) / Bvh2::point_traverse(point, eval)  (src/bvh2/mod.rs:365-456) for n queries with an eval
 * that always continues: counts[i] = number of leaf nodes reported for query i; leaf_ids receives the reported leaf NODE
 * ids query after query (query i starts at the sum of counts[0..i)), each query's ids in exactly the order the reference
 * calls eval. counts / leaf_ids may be NULL; *total = sum of counts. When leaf_ids != NULL and *total > capacity nothing
 * is written to leaf_ids and OBVHS_ERR_CAPACITY is returned (counts and *total are valid). points: Vec3A, 4 floats each. */
int obvhs_cuda_bvh2_aabb_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const ObvhsAabb* queries, size_t n,
                                        uint32_t* counts, uint32_t* leaf_ids, size_t capacity, size_t* total);
int obvhs_cuda_bvh2_point_traverse_batch(ObvhsContext* ctx, const ObvhsBvh2* bvh, const float* points, size_t n,
                                         uint32_t* counts, uint32_t* leaf_ids, size_t capacity, size_t* total);
/* traverse!(bvh, node, state, node.intersect_aabb(&aabb, state.oct_inv4), { state.primitive_id }) and the contains_point
 * form (src/cwbvh/traverse_macro.rs:59-126, src/cwbvh/node.rs:157-200): reports primitive slots (indices into
 * primitive_indices). traversal_direction: 3 host floats ordering the children (CwBvh::new_traversal, cwbvh/mod.rs:146),
 * NULL = Vec3A::ZERO as in the reference's tests. Same output convention as above. */
int obvhs_cuda_cwbvh_aabb_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const ObvhsAabb* queries, size_t n,
                                         const float* traversal_direction, uint32_t* counts, uint32_t* primitive_ids,
                                         size_t capacity, size_t* total);
int obvhs_cuda_cwbvh_point_traverse_batch(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const float* points, size_t n,
                                          const float* traversal_direction, uint32_t* counts, uint32_t* primitive_ids,
                                          size_t capacity, size_t* total);

/* Ray::new for n rays (src/ray.rs:34-52): origin_dir is n x 6 floats (ox,oy,oz,dx,dy,dz). */
int obvhs_cuda_make_rays(ObvhsContext* ctx, const float* origin_dir, size_t n, float tmin, float tmax, ObvhsRay* rays);

#ifdef __cplusplus
}
#endif
#endif /* OBVHS_CUDA_H */
