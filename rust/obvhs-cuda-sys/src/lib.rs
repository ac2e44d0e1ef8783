//! Raw bindings to `libobvhs_cuda` (`include/obvhs_cuda.h`): the B200 (sm_100a) implementation of obvhs' data-parallel hot path
//! -- PLOC BVH2 build, parallel reinsertion, BVH2 -> CWBVH collapse, CWBVH / BVH2 ray traversal, broad-phase queries -- plus the
//! NCCL broadcast of a finished tree. **Written but NOT compiled in the build image (no rustc/cargo)**; the `extern "C"` block is generated from the header by
//! `scripts/gen_rust_sys.py`, and `tests/test_rust_shim.py` checks that every C symbol is declared here with the same arity.
//!
//! POD layouts are byte-identical to the `#[repr(C)]` types of obvhs 0.3.1 they mirror (sizes asserted below and, on the C side, in
//! `obvhs_b200/csrc/common.cuh`): inside the obvhs crate these structs are replaced by `crate::{aabb::Aabb, triangle::Triangle,
//! bvh2::node::Bvh2Node, cwbvh::node::CwBvhNode, ray::{Ray, RayHit}}` (INTEGRATION.md section 3 lists the safe wrappers).
//! Every call returns 0 or a negative `ObvhsStatus`; nothing unwinds across the boundary.
#![allow(non_camel_case_types, clippy::too_many_arguments, clippy::missing_safety_doc)]
use core::ffi::{c_char, c_int, c_void};

pub const OBVHS_OK: c_int = 0;
pub const OBVHS_ERR_INVALID_ARG: c_int = -1;
pub const OBVHS_ERR_CUDA: c_int = -2;
pub const OBVHS_ERR_UNSUPPORTED: c_int = -3;
pub const OBVHS_ERR_NAN_INPUT: c_int = -4;
pub const OBVHS_ERR_STACK_OVERFLOW: c_int = -5;
pub const OBVHS_ERR_CAPACITY: c_int = -6;
pub const OBVHS_ERR_NCCL: c_int = -7;
pub const OBVHS_NCCL_UNIQUE_ID_BYTES: usize = 128;

/// src/aabb.rs:11-16 (two Vec3A lanes; the 4th float of each lane is padding)
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default)]
pub struct Aabb { pub min: [f32; 3], pub _pad0: f32, pub max: [f32; 3], pub _pad1: f32 }
/// src/triangle.rs:8-13
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default)]
pub struct Triangle { pub v0: [f32; 3], pub _pad0: f32, pub v1: [f32; 3], pub _pad1: f32, pub v2: [f32; 3], pub _pad2: f32 }
/// src/bvh2/node.rs:40-66 (default 48-byte layout)
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default)]
pub struct Bvh2Node { pub aabb: Aabb, pub prim_count: u32, pub first_index: u32, pub meta1: u32, pub meta2: u32 }
/// src/cwbvh/node.rs:12-54 (80 bytes, no padding)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct CwBvhNode {
    pub p: [f32; 3], pub e: [u8; 3], pub imask: u8, pub child_base_idx: u32, pub primitive_base_idx: u32, pub child_meta: [u8; 8],
    pub child_min_x: [u8; 8], pub child_max_x: [u8; 8], pub child_min_y: [u8; 8], pub child_max_y: [u8; 8], pub child_min_z: [u8; 8], pub child_max_z: [u8; 8],
}
/// src/ray.rs:15-30
#[repr(C, align(16))]
#[derive(Clone, Copy, Debug, Default)]
pub struct Ray {
    pub origin: [f32; 3], pub _pad0: f32, pub direction: [f32; 3], pub _pad1: f32, pub inv_direction: [f32; 3], pub _pad2: f32,
    pub tmin: f32, pub tmax: f32, pub _pad3: [f32; 2],
}
/// the arguments of `Ray::new(origin, direction, min, max)` (src/ray.rs:34-52): the constructor runs on the device
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RayNew { pub origin: [f32; 3], pub tmin: f32, pub direction: [f32; 3], pub tmax: f32 }
/// origin and direction of `Ray::new_inf(origin, direction)` (src/ray.rs:55-57); one (tmin, tmax) per batch
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RayOd { pub origin: [f32; 3], pub direction: [f32; 3] }
/// what the triangle closure writes into a `RayHit` (cwbvh/mod.rs:184-189)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RayHit8 { pub primitive_id: u32, pub t: f32 }
/// src/ray.rs:63-70
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RayHit { pub primitive_id: u32, pub geometry_id: u32, pub instance_id: u32, pub t: f32 }
/// src/lib.rs:208-231 BvhBuildParams, field for field
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct BuildParams {
    pub pre_split: u32,
    pub ploc_search_distance: u32,
    pub search_depth_threshold: u64,
    pub reinsertion_batch_ratio: f32,
    pub post_collapse_reinsertion_batch_ratio_multiplier: f32,
    pub sort_precision: u32,
    pub max_prims_per_leaf: u32,
    pub collapse_traversal_cost: f32,
}
const _: () = {
    assert!(core::mem::size_of::<Aabb>() == 32);
    assert!(core::mem::size_of::<Triangle>() == 48);
    assert!(core::mem::size_of::<Bvh2Node>() == 48);
    assert!(core::mem::size_of::<CwBvhNode>() == 80);
    assert!(core::mem::size_of::<Ray>() == 64);
    assert!(core::mem::size_of::<RayNew>() == 32);
    assert!(core::mem::size_of::<RayOd>() == 24);
    assert!(core::mem::size_of::<RayHit8>() == 8);
    assert!(core::mem::size_of::<RayHit>() == 16);
};

#[repr(C)] pub struct Context { _p: [u8; 0] }
#[repr(C)] pub struct Bvh2 { _p: [u8; 0] }
#[repr(C)] pub struct CwBvh { _p: [u8; 0] }

extern "C" {
    pub fn obvhs_cuda_create(device: c_int, stream: *mut c_void, out: *mut *mut Context) -> c_int;
    pub fn obvhs_cuda_destroy(ctx: *mut Context);
    pub fn obvhs_cuda_last_error(ctx: *const Context) -> *const c_char;
    pub fn obvhs_cuda_synchronize(ctx: *mut Context) -> c_int;
    pub fn obvhs_cuda_launch_count(ctx: *const Context) -> u64;
    pub fn obvhs_cuda_set_option(ctx: *mut Context, key: *const c_char, value: *const c_char) -> c_int;
    pub fn obvhs_cuda_build_params_preset(name: *const c_char, out: *mut BuildParams) -> c_int;
    pub fn obvhs_cuda_split_aabbs_precise(
        ctx: *mut Context,
        aabbs: *mut Aabb,
        indices: *mut u32,
        n: usize,
        capacity: usize,
        tris: *const Triangle,
        n_tris: usize,
        area_thresh_low: f32,
        area_thresh_high: f32,
        split_factor_low: f32,
        split_factor_high: f32,
        max_iterations: u32,
        split_tests: u32,
        count_out: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_split_aabbs_preset(
        ctx: *mut Context,
        aabbs: *mut Aabb,
        indices: *mut u32,
        n: usize,
        capacity: usize,
        tris: *const Triangle,
        n_tris: usize,
        avg_half_area: f32,
        largest_half_area: f32,
        count_out: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_presplit_tris(
        ctx: *mut Context,
        tris: *const Triangle,
        n: usize,
        aabbs_out: *mut Aabb,
        indices_out: *mut u32,
        capacity: usize,
        count_out: *mut usize,
        avg_half_area: *mut f32,
        largest_half_area: *mut f32,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_uses_spatial_splits(bvh: *const Bvh2) -> c_int;
    pub fn obvhs_cuda_bvh2_set_uses_spatial_splits(bvh: *mut Bvh2, value: c_int);
    pub fn obvhs_cuda_cwbvh_uses_spatial_splits(bvh: *const CwBvh) -> c_int;
    pub fn obvhs_cuda_cwbvh_set_uses_spatial_splits(bvh: *mut CwBvh, value: c_int);
    pub fn obvhs_cuda_morton_sort(
        ctx: *mut Context,
        aabbs: *const Aabb,
        n: usize,
        sort_precision: u32,
        codes_lo: *mut u64,
        codes_hi: *mut u64,
        order: *mut u32,
        total_aabb: *mut Aabb,
    ) -> c_int;
    pub fn obvhs_cuda_ploc_build(
        ctx: *mut Context,
        aabbs: *const Aabb,
        indices: *const u32,
        n: usize,
        search_distance: u32,
        sort_precision: u32,
        search_depth_threshold: usize,
        out: *mut *mut Bvh2,
    ) -> c_int;
    pub fn obvhs_cuda_ploc_build_tris(
        ctx: *mut Context,
        tris: *const Triangle,
        n: usize,
        search_distance: u32,
        sort_precision: u32,
        search_depth_threshold: usize,
        out: *mut *mut Bvh2,
    ) -> c_int;
    pub fn obvhs_cuda_ploc_full_rebuild(
        ctx: *mut Context,
        bvh: *mut Bvh2,
        search_distance: u32,
        sort_precision: u32,
        search_depth_threshold: usize,
    ) -> c_int;
    pub fn obvhs_cuda_ploc_partial_rebuild(
        ctx: *mut Context,
        bvh: *mut Bvh2,
        should_remove: *const u8,
        search_distance: u32,
        sort_precision: u32,
        search_depth_threshold: usize,
    ) -> c_int;
    pub fn obvhs_cuda_compute_rebuild_path_flags(
        ctx: *mut Context,
        bvh: *const Bvh2,
        leaves: *const u32,
        n_leaves: usize,
        flags: *mut u8,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_set_node_aabbs(
        ctx: *mut Context,
        bvh: *mut Bvh2,
        node_ids: *const u32,
        aabbs: *const Aabb,
        n: usize,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_free(bvh: *mut Bvh2);
    pub fn obvhs_cuda_bvh2_node_count(bvh: *const Bvh2) -> usize;
    pub fn obvhs_cuda_bvh2_prim_count(bvh: *const Bvh2) -> usize;
    pub fn obvhs_cuda_bvh2_max_depth(bvh: *const Bvh2) -> usize;
    pub fn obvhs_cuda_bvh2_ploc_iterations(bvh: *const Bvh2) -> usize;
    pub fn obvhs_cuda_bvh2_children_ordered_after_parents(bvh: *const Bvh2) -> c_int;
    pub fn obvhs_cuda_bvh2_download(
        ctx: *mut Context,
        bvh: *const Bvh2,
        nodes: *mut Bvh2Node,
        primitive_indices: *mut u32,
        parents: *mut u32,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_upload(
        ctx: *mut Context,
        nodes: *const Bvh2Node,
        node_count: usize,
        primitive_indices: *const u32,
        prim_count: usize,
        max_depth: usize,
        children_ordered_after_parents: c_int,
        out: *mut *mut Bvh2,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_compute_parents(ctx: *mut Context, bvh: *mut Bvh2) -> c_int;
    pub fn obvhs_cuda_bvh2_refit_all(ctx: *mut Context, bvh: *mut Bvh2) -> c_int;
    pub fn obvhs_cuda_bvh2_reorder_in_stack_traversal_order(ctx: *mut Context, bvh: *mut Bvh2) -> c_int;
    pub fn obvhs_cuda_bvh2_set_leaf_aabbs(ctx: *mut Context, bvh: *mut Bvh2, prim_aabbs: *const Aabb, n: usize) -> c_int;
    pub fn obvhs_cuda_reinsertion_run(
        ctx: *mut Context,
        bvh: *mut Bvh2,
        batch_size_ratio: f32,
        ratio_sequence: *const f32,
        n_sequence: usize,
        applied_out: *mut u64,
    ) -> c_int;
    pub fn obvhs_cuda_reinsertion_run_with_candidates(
        ctx: *mut Context,
        bvh: *mut Bvh2,
        node_ids: *const u32,
        n: usize,
        iterations: u32,
        applied_out: *mut u64,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_collapse(ctx: *mut Context, bvh: *mut Bvh2, max_prims: u32, traversal_cost: f32) -> c_int;
    pub fn obvhs_cuda_build_bvh2_from_tris(
        ctx: *mut Context,
        tris: *const Triangle,
        n: usize,
        params: *const BuildParams,
        core_build_seconds: *mut f64,
        out: *mut *mut Bvh2,
    ) -> c_int;
    pub fn obvhs_cuda_build_bvh2(
        ctx: *mut Context,
        aabbs: *const Aabb,
        n: usize,
        params: *const BuildParams,
        core_build_seconds: *mut f64,
        out: *mut *mut Bvh2,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_set_triangles(ctx: *mut Context, bvh: *mut Bvh2, tris: *const Triangle, n: usize) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_traverse_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        rays: *const Ray,
        n: usize,
        hits: *mut RayHit,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_traverse_miss_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        rays: *const Ray,
        n: usize,
        miss: *mut u8,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_traverse_anyhit_count_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        rays: *const Ray,
        n: usize,
        counts: *mut u32,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_traverse_batch_counted(
        ctx: *mut Context,
        bvh: *const Bvh2,
        rays: *const Ray,
        n: usize,
        hits: *mut RayHit,
        counters: *mut u64,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_to_cwbvh(
        ctx: *mut Context,
        bvh: *const Bvh2,
        max_prims_per_leaf: u32,
        order_children: c_int,
        include_exact_node_aabbs: c_int,
        out: *mut *mut CwBvh,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_exact_node_aabbs(
        ctx: *mut Context,
        bvh: *const CwBvh,
        out: *mut Aabb,
        capacity: usize,
        count: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_compute_parents(ctx: *mut Context, bvh: *const CwBvh, parents: *mut u32) -> c_int;
    pub fn obvhs_cuda_cwbvh_order_children(
        ctx: *mut Context,
        bvh: *mut CwBvh,
        prim_aabbs: *const Aabb,
        n: usize,
        direct_layout: c_int,
    ) -> c_int;
    pub fn obvhs_cuda_build_cwbvh_from_tris(
        ctx: *mut Context,
        tris: *const Triangle,
        n: usize,
        params: *const BuildParams,
        core_build_seconds: *mut f64,
        out: *mut *mut CwBvh,
    ) -> c_int;
    pub fn obvhs_cuda_build_cwbvh(
        ctx: *mut Context,
        aabbs: *const Aabb,
        n: usize,
        params: *const BuildParams,
        core_build_seconds: *mut f64,
        out: *mut *mut CwBvh,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_free(bvh: *mut CwBvh);
    pub fn obvhs_cuda_cwbvh_node_count(bvh: *const CwBvh) -> usize;
    pub fn obvhs_cuda_cwbvh_prim_count(bvh: *const CwBvh) -> usize;
    pub fn obvhs_cuda_cwbvh_download(
        ctx: *mut Context,
        bvh: *const CwBvh,
        nodes: *mut CwBvhNode,
        primitive_indices: *mut u32,
        total_aabb: *mut Aabb,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_upload(
        ctx: *mut Context,
        nodes: *const CwBvhNode,
        node_count: usize,
        primitive_indices: *const u32,
        prim_count: usize,
        total_aabb: *const Aabb,
        out: *mut *mut CwBvh,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_set_triangles(ctx: *mut Context, bvh: *mut CwBvh, tris: *const Triangle, n: usize) -> c_int;
    pub fn obvhs_cuda_cwbvh_triangle_bytes() -> usize;
    pub fn obvhs_cuda_cwbvh_device_ptrs(
        bvh: *const CwBvh,
        nodes: *mut *mut c_void,
        primitive_indices: *mut *mut c_void,
        bvh_tris: *mut *mut c_void,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_alloc(
        ctx: *mut Context,
        node_count: usize,
        prim_count: usize,
        with_triangles: c_int,
        total_aabb: *const Aabb,
        out: *mut *mut CwBvh,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_traverse_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        rays: *const Ray,
        n: usize,
        hits: *mut RayHit,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_traverse_miss_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        rays: *const Ray,
        n: usize,
        miss: *mut u8,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_traverse_anyhit_count_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        rays: *const Ray,
        n: usize,
        counts: *mut u32,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_traverse_batch_counted(
        ctx: *mut Context,
        bvh: *const CwBvh,
        rays: *const Ray,
        n: usize,
        hits: *mut RayHit,
        counters: *mut u64,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_new_traverse_batch_counted(
        ctx: *mut Context,
        bvh: *const CwBvh,
        args: *const RayNew,
        n: usize,
        hits: *mut RayHit,
        counters: *mut u64,
    ) -> c_int;
    pub fn obvhs_cuda_ray_new_batch(ctx: *mut Context, args: *const RayNew, n: usize, rays: *mut Ray) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_new_traverse_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        args: *const RayNew,
        n: usize,
        hits: *mut RayHit,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_new_traverse_miss_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        args: *const RayNew,
        n: usize,
        miss: *mut u8,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_new_traverse_anyhit_count_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        args: *const RayNew,
        n: usize,
        counts: *mut u32,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_new_traverse_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        args: *const RayNew,
        n: usize,
        hits: *mut RayHit,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_new_traverse_miss_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        args: *const RayNew,
        n: usize,
        miss: *mut u8,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_od_traverse_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        od: *const RayOd,
        n: usize,
        tmin: f32,
        tmax: f32,
        hits: *mut RayHit,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        od: *const RayOd,
        n: usize,
        tmin: f32,
        tmax: f32,
        hits: *mut RayHit8,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_ray_od_traverse_miss_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        od: *const RayOd,
        n: usize,
        tmin: f32,
        tmax: f32,
        miss: *mut u8,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_ray_od_traverse_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        od: *const RayOd,
        n: usize,
        tmin: f32,
        tmax: f32,
        hits: *mut RayHit,
    ) -> c_int;
    pub fn obvhs_cuda_nccl_unique_id(id: *mut u8) -> c_int;
    pub fn obvhs_cuda_comm_init(ctx: *mut Context, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn obvhs_cuda_cwbvh_broadcast(ctx: *mut Context, bvh: *mut *mut CwBvh, root: c_int) -> c_int;
    pub fn obvhs_cuda_bvh2_aabb_traverse_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        queries: *const Aabb,
        n: usize,
        counts: *mut u32,
        leaf_ids: *mut u32,
        capacity: usize,
        total: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_bvh2_point_traverse_batch(
        ctx: *mut Context,
        bvh: *const Bvh2,
        points: *const f32,
        n: usize,
        counts: *mut u32,
        leaf_ids: *mut u32,
        capacity: usize,
        total: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_aabb_traverse_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        queries: *const Aabb,
        n: usize,
        traversal_direction: *const f32,
        counts: *mut u32,
        primitive_ids: *mut u32,
        capacity: usize,
        total: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_cwbvh_point_traverse_batch(
        ctx: *mut Context,
        bvh: *const CwBvh,
        points: *const f32,
        n: usize,
        traversal_direction: *const f32,
        counts: *mut u32,
        primitive_ids: *mut u32,
        capacity: usize,
        total: *mut usize,
    ) -> c_int;
    pub fn obvhs_cuda_make_rays(
        ctx: *mut Context,
        origin_dir: *const f32,
        n: usize,
        tmin: f32,
        tmax: f32,
        rays: *mut Ray,
    ) -> c_int;
}
