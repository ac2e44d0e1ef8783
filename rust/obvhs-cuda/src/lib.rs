//! Safe wrappers over `obvhs-cuda-sys` that keep obvhs 0.3.1's names, argument order and error behaviour for the hot path
//! (INTEGRATION.md section 3 maps every item to the reference file:line it replaces). **Written but NOT compiled in the build image**
//! (no rustc/cargo); the C++ mirror `include/obvhs.hpp` has the same shape and IS compiled and run on the GPU
//! (`tests/cpp/host_api.cpp`), and the ctypes binding `obvhs_b200/api.py` drives the parity tests.
//!
//! Inside the obvhs crate the POD types below are the crate's own (`Aabb`, `Triangle`, `Ray`, `RayHit`, `CwBvhNode`, ...): they are
//! `#[repr(C)]` + `Pod` there already, with the byte layouts `obvhs_cuda_sys` asserts.
use obvhs_cuda_sys as sys;
pub use sys::{Aabb, BuildParams as BvhBuildParams, Bvh2Node, CwBvhNode, Ray, RayHit, RayNew, RayOd, Triangle};
use std::{ffi::CStr, ptr, time::Duration};

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }
pub type Result<T> = std::result::Result<T, Error>;

/// One device + stream + reusable builder scratch (the reference's `PlocBuilder` keeps its Vecs for reuse, ploc/mod.rs:35-54).
pub struct Context { h: *mut sys::Context }
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut h = ptr::null_mut();
        let rc = unsafe { sys::obvhs_cuda_create(device, ptr::null_mut(), &mut h) };
        if rc != 0 { return Err(Error { code: rc, message: "obvhs_cuda_create failed: no CUDA device (there is no CPU fallback)".into() }); }
        Ok(Self { h })
    }
    fn check(&self, rc: i32) -> Result<()> {
        if rc == 0 { return Ok(()); }
        let message = unsafe { CStr::from_ptr(sys::obvhs_cuda_last_error(self.h)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }
    /// `obvhs_cuda_comm_init`: collective over `world` ranks; `id` comes from `nccl_unique_id()` on one rank.
    pub fn comm_init(&mut self, id: &[u8; sys::OBVHS_NCCL_UNIQUE_ID_BYTES], rank: i32, world: i32) -> Result<()> {
        self.check(unsafe { sys::obvhs_cuda_comm_init(self.h, id.as_ptr(), rank, world) })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::obvhs_cuda_destroy(self.h) } } }

pub fn nccl_unique_id() -> Result<[u8; sys::OBVHS_NCCL_UNIQUE_ID_BYTES]> {
    let mut id = [0u8; sys::OBVHS_NCCL_UNIQUE_ID_BYTES];
    let rc = unsafe { sys::obvhs_cuda_nccl_unique_id(id.as_mut_ptr()) };
    if rc != 0 { return Err(Error { code: rc, message: "NCCL is not available".into() }); }
    Ok(id)
}

/// `PlocSearchDistance` (ploc/mod.rs:534-562) and `SortPrecision` (:658-661) as the u32 forms the C ABI takes.
#[derive(Clone, Copy)] #[repr(u32)] pub enum PlocSearchDistance { Minimum = 1, VeryLow = 2, Low = 6, Medium = 14, High = 24, VeryHigh = 32 }
#[derive(Clone, Copy)] #[repr(u32)] pub enum SortPrecision { U64 = 64, U128 = 128 }

/// Device-resident `Bvh2` (bvh2/mod.rs:31-85).
pub struct Bvh2 { h: *mut sys::Bvh2 }
impl Drop for Bvh2 { fn drop(&mut self) { unsafe { sys::obvhs_cuda_bvh2_free(self.h) } } }
impl Bvh2 {
    /// `Bvh2::reorder_in_stack_traversal_order(&mut self)` (bvh2/mod.rs:462-500)
    pub fn reorder_in_stack_traversal_order(&mut self, ctx: &Context) -> Result<()> {
        ctx.check(unsafe { sys::obvhs_cuda_bvh2_reorder_in_stack_traversal_order(ctx.h, self.h) })
    }
}

/// `PlocBuilder::build(&mut self, search_distance, aabbs, indices, sort_precision, search_depth_threshold) -> Bvh2` (ploc/mod.rs:95-102)
pub struct PlocBuilder<'c> { pub ctx: &'c Context }
impl PlocBuilder<'_> {
    pub fn build(&mut self, search_distance: PlocSearchDistance, aabbs: &[Aabb], indices: &[u32], sort_precision: SortPrecision,
                 search_depth_threshold: usize) -> Result<Bvh2> {
        assert!(indices.is_empty() || indices.len() == aabbs.len());
        let idx = if indices.is_empty() { ptr::null() } else { indices.as_ptr() };
        let mut h = ptr::null_mut();
        self.ctx.check(unsafe { sys::obvhs_cuda_ploc_build(self.ctx.h, aabbs.as_ptr(), idx, aabbs.len(), search_distance as u32,
                                                          sort_precision as u32, search_depth_threshold, &mut h) })?;
        Ok(Bvh2 { h })
    }
}

/// `ReinsertionOptimizer::run(&mut self, bvh, batch_size_ratio, ratio_sequence)` (bvh2/reinsertion.rs:40-57)
#[derive(Default)] pub struct ReinsertionOptimizer { pub applied: u64 }
impl ReinsertionOptimizer {
    pub fn run(&mut self, ctx: &Context, bvh: &mut Bvh2, batch_size_ratio: f32, ratio_sequence: Option<Vec<f32>>) -> Result<()> {
        let (p, n) = ratio_sequence.as_ref().map_or((ptr::null(), 0), |s| (s.as_ptr(), s.len()));
        ctx.check(unsafe { sys::obvhs_cuda_reinsertion_run(ctx.h, bvh.h, batch_size_ratio, p, n, &mut self.applied) })
    }
}

/// Device-resident `CwBvh` (cwbvh/mod.rs:43-55) with the triangles permuted by `primitive_indices` (examples/obj_cwbvh.rs:63-67).
pub struct CwBvh { h: *mut sys::CwBvh }
impl Drop for CwBvh { fn drop(&mut self) { unsafe { sys::obvhs_cuda_cwbvh_free(self.h) } } }

/// `bvh2_to_cwbvh(&bvh2, max_prims_per_leaf, order_children, include_exact_node_aabbs) -> CwBvh` (cwbvh/bvh2_to_cwbvh.rs:490-510)
pub fn bvh2_to_cwbvh(ctx: &Context, bvh2: &Bvh2, max_prims_per_leaf: u32, order_children: bool, include_exact_node_aabbs: bool) -> Result<CwBvh> {
    let mut h = ptr::null_mut();
    ctx.check(unsafe { sys::obvhs_cuda_bvh2_to_cwbvh(ctx.h, bvh2.h, max_prims_per_leaf, order_children as i32, include_exact_node_aabbs as i32, &mut h) })?;
    Ok(CwBvh { h })
}

/// `build_cwbvh_from_tris(triangles, config, core_build_time) -> CwBvh` (cwbvh/builder.rs:20-85)
pub fn build_cwbvh_from_tris(ctx: &Context, triangles: &[Triangle], config: BvhBuildParams, core_build_time: &mut Duration) -> Result<CwBvh> {
    let (mut h, mut secs) = (ptr::null_mut(), 0.0f64);
    ctx.check(unsafe { sys::obvhs_cuda_build_cwbvh_from_tris(ctx.h, triangles.as_ptr(), triangles.len(), &config, &mut secs, &mut h) })?;
    *core_build_time += Duration::from_secs_f64(secs);
    Ok(CwBvh { h })
}

impl CwBvh {
    /// `(nodes, primitive_indices, total_aabb)` exactly as the reference's `CwBvh` holds them.
    pub fn download(&self, ctx: &Context) -> Result<(Vec<CwBvhNode>, Vec<u32>, Aabb)> {
        let (m, n) = unsafe { (sys::obvhs_cuda_cwbvh_node_count(self.h), sys::obvhs_cuda_cwbvh_prim_count(self.h)) };
        let (mut nodes, mut prims, mut total) = (vec![CwBvhNode::default(); m], vec![0u32; n], Aabb::default());
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_download(ctx.h, self.h, nodes.as_mut_ptr(), prims.as_mut_ptr(), &mut total) })?;
        Ok((nodes, prims, total))
    }
    /// Batched `CwBvh::ray_traverse(ray, &mut hit, |ray, id| bvh_tris[id].intersect(ray))` (cwbvh/mod.rs:169-198). `hit.primitive_id`
    /// indexes `primitive_indices` order, as in the reference. The per-ray closure cannot cross the boundary: it is fixed to the
    /// triangle test (BASELINE north_star).
    pub fn ray_traverse_batch(&self, ctx: &Context, rays: &[Ray], hits: &mut [RayHit]) -> Result<()> {
        assert_eq!(rays.len(), hits.len());
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_ray_traverse_batch(ctx.h, self.h, rays.as_ptr(), rays.len(), hits.as_mut_ptr()) })
    }
    /// The same over the arguments of `Ray::new` (ray.rs:34-52): the constructor runs on the device, half the PCIe bytes.
    pub fn ray_new_traverse_batch(&self, ctx: &Context, args: &[RayNew], hits: &mut [RayHit]) -> Result<()> {
        assert_eq!(args.len(), hits.len());
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_ray_new_traverse_batch(ctx.h, self.h, args.as_ptr(), args.len(), hits.as_mut_ptr()) })
    }
    /// The same for `Ray::new(od.origin, od.direction, tmin, tmax)` with one pair of bounds per batch (`Ray::new_inf` = 0, INFINITY):
    /// 24 bytes per ray cross PCIe.
    pub fn ray_od_traverse_batch(&self, ctx: &Context, od: &[RayOd], tmin: f32, tmax: f32, hits: &mut [RayHit]) -> Result<()> {
        assert_eq!(od.len(), hits.len());
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_ray_od_traverse_batch(ctx.h, self.h, od.as_ptr(), od.len(), tmin, tmax, hits.as_mut_ptr()) })
    }
    /// Batched `CwBvh::ray_traverse_miss` (cwbvh/mod.rs:201-225)
    pub fn ray_traverse_miss_batch(&self, ctx: &Context, rays: &[Ray], miss: &mut [u8]) -> Result<()> {
        assert_eq!(rays.len(), miss.len());
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_ray_traverse_miss_batch(ctx.h, self.h, rays.as_ptr(), rays.len(), miss.as_mut_ptr()) })
    }
    /// `CwBvh::order_children(&mut self, primitives: &[T], direct_layout)` (cwbvh/mod.rs:520-524) with `T::aabb()` collected by the
    /// caller (`primitives.iter().map(Boundable::aabb)`).
    pub fn order_children(&mut self, ctx: &Context, prim_aabbs: &[Aabb], direct_layout: bool) -> Result<()> {
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_order_children(ctx.h, self.h, prim_aabbs.as_ptr(), prim_aabbs.len(), direct_layout as i32) })
    }
    /// Replicate the tree built on `root` to every rank of the context's communicator (`Clone` across GPUs): pass the tree on
    /// `root`, `None` (or an earlier replica to refill) elsewhere.
    pub fn broadcast(tree: Option<CwBvh>, ctx: &Context, root: i32) -> Result<CwBvh> {
        let mut h = tree.as_ref().map_or(ptr::null_mut(), |t| t.h);
        std::mem::forget(tree);  // the library keeps, refills or replaces the handle
        ctx.check(unsafe { sys::obvhs_cuda_cwbvh_broadcast(ctx.h, &mut h, root) })?;
        Ok(CwBvh { h })
    }
}
