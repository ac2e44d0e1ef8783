// common.cuh -- shared device helpers and host-side containers of libobvhs_cuda (sm_100a only).
//
// Float rules that make the GPU path bit-exact with the reference's CPU arithmetic:
//   * the library is compiled with -fmad=false: rustc never contracts a*b+c;
//   * glam Vec3A::min/max are _mm_min_ps/_mm_max_ps: min(a,b) = a<b ? a : b (second operand on ties/NaN).
//     fminf/fmaxf order -0 < +0 and drop NaNs, so they are NOT used for tree AABBs.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/obvhs_cuda.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

static_assert(sizeof(ObvhsAabb) == 32, "Aabb");
static_assert(sizeof(ObvhsTriangle) == 48, "Triangle");
static_assert(sizeof(ObvhsBvh2Node) == 48, "Bvh2Node");
static_assert(sizeof(ObvhsCwBvhNode) == 80, "CwBvhNode");
static_assert(sizeof(ObvhsRay) == 64, "Ray");
static_assert(sizeof(ObvhsRayNew) == 32, "RayNew");
static_assert(sizeof(ObvhsRayHit) == 16, "RayHit");

#define OBVHS_RT_TRIANGLE_BYTES 64  // rt_triangle.rs:160-168 RtTriangle {v0, e1, e2, ng}: the handles' internal triangle layout
#define OBVHS_SM_COUNT 148  // B200: grids are sized in multiples of this

// ----------------------------------------------------------------------------------------------------------
// Device-side BVH2 node: the reference's own 32-byte `small_bvh2_node` layout (bvh2/node.rs:12-36):
// {min.xyz, prim_count, max.xyz, first_index} = two 16-byte vectors. Expanded to the 48-byte default layout only
// on download.
// ----------------------------------------------------------------------------------------------------------
struct __align__(16) Node32 {
    float minx, miny, minz;
    u32 prim_count;
    float maxx, maxy, maxz;
    u32 first_index;
};
static_assert(sizeof(Node32) == 32, "Node32");

struct Box {
    float minx, miny, minz, maxx, maxy, maxz;
};

__device__ __forceinline__ float smin(float a, float b) { return a < b ? a : b; }  // _mm_min_ps(a,b)
__device__ __forceinline__ float smax(float a, float b) { return a > b ? a : b; }  // _mm_max_ps(a,b)

// aabb.rs:84-89  self.union(other): min(self,other), max(self,other) per lane
__device__ __forceinline__ Box box_union(const Box& a, const Box& b) {
    Box r;
    r.minx = smin(a.minx, b.minx);
    r.miny = smin(a.miny, b.miny);
    r.minz = smin(a.minz, b.minz);
    r.maxx = smax(a.maxx, b.maxx);
    r.maxy = smax(a.maxy, b.maxy);
    r.maxz = smax(a.maxz, b.maxz);
    return r;
}
// aabb.rs:151-154  (d.x + d.y) * d.z + d.x * d.y, no contraction
__device__ __forceinline__ float box_half_area(const Box& a) {
    float dx = __fsub_rn(a.maxx, a.minx), dy = __fsub_rn(a.maxy, a.miny), dz = __fsub_rn(a.maxz, a.minz);
    return __fadd_rn(__fmul_rn(__fadd_rn(dx, dy), dz), __fmul_rn(dx, dy));
}
__device__ __forceinline__ Box node_box(const Node32& n) { return Box{n.minx, n.miny, n.minz, n.maxx, n.maxy, n.maxz}; }
__device__ __forceinline__ Node32 make_node32(const Box& b, u32 prim_count, u32 first_index) {
    Node32 n;
    n.minx = b.minx; n.miny = b.miny; n.minz = b.minz; n.prim_count = prim_count;
    n.maxx = b.maxx; n.maxy = b.maxy; n.maxz = b.maxz; n.first_index = first_index;
    return n;
}
__device__ __forceinline__ Node32 load_node(const Node32* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = q[0], b = q[1];
    Node32 n;
    n.minx = a.x; n.miny = a.y; n.minz = a.z; n.prim_count = __float_as_uint(a.w);
    n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.first_index = __float_as_uint(b.w);
    return n;
}
// L2-coherent load (bypasses the non-coherent L1) for data written by other CTAs of the same launch
__device__ __forceinline__ Node32 load_node_cg(const Node32* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = __ldcg(q), b = __ldcg(q + 1);
    Node32 n;
    n.minx = a.x; n.miny = a.y; n.minz = a.z; n.prim_count = __float_as_uint(a.w);
    n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.first_index = __float_as_uint(b.w);
    return n;
}
__device__ __forceinline__ void store_node(Node32* p, const Node32& n) {
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4(n.minx, n.miny, n.minz, __uint_as_float(n.prim_count));
    q[1] = make_float4(n.maxx, n.maxy, n.maxz, __uint_as_float(n.first_index));
}
// bvh2/node.rs:154-180: siblings are adjacent, the left one has an odd index
__device__ __host__ __forceinline__ u32 sibling_id(u32 id) { return (id & 1u) ? id + 1 : id - 1; }
__device__ __host__ __forceinline__ u32 left_sibling_id(u32 id) { return (id & 1u) ? id : id - 1; }

// ----------------------------------------------------------------------------------------------------------
// Host-side objects
// ----------------------------------------------------------------------------------------------------------
struct ObvhsContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    std::string last_error;
    u64 launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* pinned = nullptr;  // 4 KB of pinned host memory for small read-backs
    int sm_count = OBVHS_SM_COUNT;
    // Scratch arena: grow-only device blocks reused by every call (the reference's builders keep their Vecs for reuse,
    // ploc/mod.rs:35-54). Allocation is a pointer bump; API calls release back to their entry mark.
    struct ArenaBlock {
        char* p;
        size_t cap;
    };
    std::vector<ArenaBlock> arena_blocks;
    size_t arena_block = 0, arena_off = 0;  // bump position: block index + offset inside it
    size_t arena_peak = 0, arena_used = 0;
    int api_depth = 0;
    // Result buffers (trees outlive the call that built them): a size-matched cache of cudaMalloc blocks owned by the context.
    // Freed trees return their blocks here, so a rebuild of a same-sized scene allocates nothing (the driver's stream-ordered
    // pool was measured to hand back fresh memory every few rebuilds of the 10M-triangle scene: 5-50 ms per allocation).
    // Handles keep the context alive (refs); obvhs_cuda_destroy only drops the creator's reference.
    std::atomic<int> refs{1};
    std::mutex result_mu;
    struct ResultBlock {
        void* p;
        size_t cap;
    };
    std::vector<ResultBlock> result_cache;
    std::unordered_map<void*, size_t> result_live;
    // host-batch pipeline (traverse_common): copy-in / copy-out streams beside `stream`, and a pool of timing-free events
    cudaStream_t copy_in = nullptr, copy_out = nullptr, compute_alt = nullptr;
    size_t host_slice = 0;  // obvhs_cuda_set_option("host_slice", "<rays>"), 0 = automatic
    std::vector<cudaEvent_t> event_pool;
    int traverse_mode = 2, traverse_refill = 4, traverse_chunk = 32;  // obvhs_cuda_set_option("traverse", "auto|static|persistent[:refill[:chunk]]")
    int traverse_variant = 0;                                           // obvhs_cuda_set_option("traverse_variant", "<id>"), traverse.cu
    size_t traverse_resident_lanes = 0;                                 // lanes the persistent kernel keeps resident (set by its launcher)
    // multi-GPU (comm.cu): an NCCL communicator bound to this context's device, and the 64-byte device header of a broadcast
    void* comm = nullptr;  // ncclComm_t
    int comm_rank = 0, comm_world = 1;
    void* comm_header = nullptr;
    bool staged_host = false;  // a stage_in() of this API call copied from HOST memory: the call synchronises before it returns
    bool trace = false;  // OBVHS_TRACE=1: per-stage wall times on stderr (the reference's scope!/timeit! macros, lib.rs:158-205)
};

// Stage scope. Always an NVTX range named like the reference's profiling scope of the same stage (`crate::scope!`: build_ploc,
// build_ploc_from_leaves, preallocate_builder, sort_nodes -- ploc/mod.rs:56,70,177,265,779; reinsertion_optimize,
// reinsertion_optimize_candidates -- reinsertion.rs:41,66; calculate_cost, convert_to_cwbvh -- bvh2_to_cwbvh.rs:70,215; collapse;
// split_aabbs_precise), visible in Nsight Systems / ncu --nvtx; NVTX3 is header-only and a no-op without a tool attached. With
// OBVHS_TRACE=1 it also synchronises the stream at entry / exit and prints the elapsed wall time of the stage.
struct TraceScope {
    ObvhsContext* ctx;
    const char* name;
    double t0 = 0;
    u64 l0 = 0;
    static double now();
    TraceScope(ObvhsContext* c, const char* n) : ctx(c), name(n) {
        const char* label = n;
        while (*label == ' ') label++;  // (the indentation only structures the OBVHS_TRACE print-out)
        nvtxRangePushA(label);
        if (ctx->trace) {
            cudaStreamSynchronize(ctx->stream);
            t0 = now();
            l0 = ctx->launches;
        }
    }
    TraceScope(const TraceScope&) = delete;
    TraceScope& operator=(const TraceScope&) = delete;
    ~TraceScope() {
        if (ctx->trace) {
            cudaStreamSynchronize(ctx->stream);
            fprintf(stderr, "[obvhs trace] %-28s %9.3f ms  %5llu launches\n", name, (now() - t0) * 1e3, (unsigned long long)(ctx->launches - l0));
        }
        nvtxRangePop();
    }
};

struct ObvhsBvh2 {
    ObvhsContext* owner = nullptr;  // holds a reference
    int device = 0;
    Node32* nodes = nullptr;          // node_count
    u32* primitive_indices = nullptr;  // prim_count
    u32* parents = nullptr;            // node_count, or null when not computed (Bvh2::parents: Option)
    void* bvh_tris = nullptr;  // triangles permuted by primitive_indices as 64-byte RtTriangles (see traverse.cu), or null
    size_t node_count = 0, prim_count = 0;
    size_t max_depth = 96;  // bvh2/mod.rs:87 DEFAULT_MAX_STACK_DEPTH
    size_t ploc_iterations = 0;
    bool children_are_ordered_after_parents = false;
    bool uses_spatial_splits = false;  // bvh2/mod.rs:84: primitive_indices may name a primitive more than once
};

struct ObvhsCwBvh {
    ObvhsContext* owner = nullptr;  // holds a reference
    int device = 0;
    ObvhsCwBvhNode* nodes = nullptr;
    u32* primitive_indices = nullptr;
    void* bvh_tris = nullptr;  // triangles permuted by primitive_indices as 64-byte RtTriangles (see traverse.cu), or null
    size_t node_count = 0, prim_count = 0;
    ObvhsAabb total_aabb = {};
    ObvhsAabb* exact_node_aabbs = nullptr;  // cwbvh/mod.rs:47, exact_count entries (the Bvh2's node count), or null
    size_t exact_count = 0;
    bool uses_spatial_splits = false;  // cwbvh/mod.rs:54
};

#define OBVHS_SET_ERR(ctx, ...)                       \
    do {                                              \
        char _b[512];                                 \
        snprintf(_b, sizeof(_b), __VA_ARGS__);        \
        (ctx)->last_error = _b;                       \
    } while (0)

#define CU_TRY(ctx, expr)                                                                           \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            OBVHS_SET_ERR(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return OBVHS_ERR_CUDA;                                                                  \
        }                                                                                           \
    } while (0)

#define ST_TRY(expr)              \
    do {                          \
        int _s = (expr);          \
        if (_s != OBVHS_OK) return _s; \
    } while (0)

// checks the launch and counts it (obvhs_cuda_launch_count)
#define KERNEL_CHECK(ctx)                      \
    do {                                       \
        (ctx)->launches++;                     \
        CU_TRY(ctx, cudaGetLastError());       \
    } while (0)

// One-time per-kernel set-up (cudaFuncSetAttribute, occupancy queries) is PER DEVICE: the flags are indexed by device ordinal so a
// process that drives several GPUs through several contexts configures every one of them.
template <class T>
struct PerDevice {
    T v[64] = {};
    T& operator[](int device) { return v[device & 63]; }
};

static inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// Scratch from the context's arena (see ObvhsContext). Released in bulk when the API call that allocated it returns.
void* obvhs_arena_alloc(size_t bytes);  // uses the context of the API call in flight on this thread; nullptr on failure
template <class T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t count, cudaStream_t) {
        if (count == 0) count = 1;
        p = static_cast<T*>(obvhs_arena_alloc(count * sizeof(T)));
        return p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
};

// Result buffers (they outlive the call, so they do not come from the arena): stream-ordered pool allocations.
cudaError_t obvhs_result_alloc(ObvhsContext* ctx, void** p, size_t bytes);
void obvhs_result_free(ObvhsContext* ctx, void* p);
void obvhs_context_retain(ObvhsContext* ctx);
void obvhs_context_release(ObvhsContext* ctx);  // destroys the context when the last reference goes

// true when ptr is device (or managed) memory
bool obvhs_is_device_ptr(const void* ptr);
// Returns a device view of `src` (count elements of T). Host memory is copied into `stage` on the stream.
template <class T>
static inline int stage_in(ObvhsContext* ctx, const T* src, size_t count, DevBuf<T>& stage, const T** out) {
    if (count == 0 || src == nullptr) {
        *out = nullptr;
        return OBVHS_OK;
    }
    if (obvhs_is_device_ptr(src)) {
        *out = src;
        return OBVHS_OK;
    }
    CU_TRY(ctx, stage.alloc(count, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(stage.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->staged_host = true;  // the caller may reuse `src` as soon as the API call returns (DeviceScope synchronises)
    *out = stage.p;
    return OBVHS_OK;
}
// copies device data to a caller pointer that may be host or device; host copies are synchronised
template <class T>
static inline int copy_out(ObvhsContext* ctx, T* dst, const T* src_dev, size_t count) {
    if (!dst || count == 0) return OBVHS_OK;
    if (obvhs_is_device_ptr(dst)) {
        CU_TRY(ctx, cudaMemcpyAsync(dst, src_dev, count * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        CU_TRY(ctx, cudaMemcpyAsync(dst, src_dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return OBVHS_OK;
}

// comm.cu
int comm_unique_id(uint8_t* id);
int comm_init(ObvhsContext* ctx, const uint8_t* id, int rank, int world);
void comm_destroy(ObvhsContext* ctx);
int comm_broadcast_cwbvh(ObvhsContext* ctx, ObvhsCwBvh** bvh, int root);

// ---- stage entry points implemented across the .cu files ---------------------------------------------------
// ploc.cu
struct PlocMortonOut {
    u64* codes_lo = nullptr;  // device, original order (optional)
    u64* codes_hi = nullptr;
    u32* order = nullptr;
    ObvhsAabb* total = nullptr;  // device, 1 element
};
int ploc_build_device(ObvhsContext* ctx, const ObvhsAabb* d_aabbs, const ObvhsTriangle* d_tris, const u32* d_indices, size_t n,
                      u32 search_distance, u32 sort_precision, size_t search_depth_threshold, ObvhsBvh2** out,
                      const PlocMortonOut* probe);
// ploc.cu : PlocBuilder::full_rebuild / partial_rebuild / compute_rebuild_path_flags (src/ploc/rebuild.rs); flags are one byte per node
int ploc_full_rebuild_device(ObvhsContext* ctx, ObvhsBvh2* bvh, u32 search_distance, u32 sort_precision, size_t search_depth_threshold);
int ploc_partial_rebuild_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const u8* d_should_remove, u32 search_distance, u32 sort_precision,
                                size_t search_depth_threshold);
int ploc_compute_rebuild_path_flags_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, const u32* d_leaves, size_t n_leaves, u8* d_flags);
// query.cu : batched Bvh2::aabb_traverse / point_traverse and the CwBvh traverse! macro over intersect_aabb / contains_point.
// query_kind 0 = boxes (2 float4 per query), 1 = points (1 float4). counts / ids may be host or device.
int bvh2_query_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, int query_kind, const float4* d_queries, size_t n, u32* counts, u32* ids,
                      size_t capacity, size_t* total_out);
int cwbvh_compute_parents_device(ObvhsContext* ctx, const ObvhsCwBvh* bvh, u32* d_parents);
int cwbvh_query_device(ObvhsContext* ctx, const ObvhsCwBvh* bvh, int query_kind, const float4* d_queries, size_t n, const float* host_dir3,
                       u32* counts, u32* ids, size_t capacity, size_t* total_out);
// splits.cu : spatial pre-splits (src/splits.rs). Arrays live in the arena of the API call in flight and grow like the Vecs.
struct SplitArrays {
    ObvhsAabb* aabbs = nullptr;
    u32* indices = nullptr;
    size_t len = 0, cap = 0;
};
int split_aabbs_precise_device(ObvhsContext* ctx, SplitArrays& a, const ObvhsTriangle* d_tris, float area_thresh_low, float area_thresh_high,
                               float split_factor_low, float split_factor_high, u32 max_iterations, u32 split_tests);
int presplit_tris_device(ObvhsContext* ctx, const ObvhsTriangle* d_tris, size_t n, SplitArrays& a, float* avg_largest_host, cudaEvent_t ev_start);
// sort.cu : stable LSD radix sort (onesweep) of (key, value) pairs; keys in `keys` / `vals`, scratch in *_alt.
// On return *sorted_keys / *sorted_vals point at whichever buffer holds the result.
int radix_sort_pairs_u64(ObvhsContext* ctx, u64* keys, u64* keys_alt, u32* vals, u32* vals_alt, size_t n, int key_bytes,
                         u64** sorted_keys, u32** sorted_vals);
int radix_sort_pairs_u32(ObvhsContext* ctx, u32* keys, u32* keys_alt, u32* vals, u32* vals_alt, size_t n, int key_bytes,
                         u32** sorted_keys, u32** sorted_vals);
// bvh2.cu
int bvh2_compute_parents_device(ObvhsContext* ctx, ObvhsBvh2* bvh);
int bvh2_compute_parents_into(ObvhsContext* ctx, const ObvhsBvh2* bvh, u32* d_parents);
int bvh2_refit_all_device(ObvhsContext* ctx, ObvhsBvh2* bvh);
int bvh2_reorder_in_stack_traversal_order_device(ObvhsContext* ctx, ObvhsBvh2* bvh);
int bvh2_set_node_aabbs_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const u32* d_node_ids, const ObvhsAabb* d_aabbs, size_t n);
// collapse.cu
int bvh2_collapse_device(ObvhsContext* ctx, ObvhsBvh2* bvh, u32 max_prims, float traversal_cost);
// reinsertion.cu
int reinsertion_run_device(ObvhsContext* ctx, ObvhsBvh2* bvh, float ratio, const float* seq, size_t n_seq, u64* applied_out);
int reinsertion_run_candidates_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const u32* d_node_ids, size_t n, u32 iterations, u64* applied_out);
// cwbvh_build.cu
int cwbvh_order_children_device(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsAabb* d_prim_aabbs, size_t n_prims, bool direct_layout);
int bvh2_to_cwbvh_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, u32 max_prims_per_leaf, bool order_children, ObvhsCwBvh** out,
                         bool include_exact_node_aabbs = false);
// traverse.cu
// How the rays of a batch lie in memory. kind 0: ObvhsRay[n] (64 B). kind 1: ObvhsRayNew[n] (32 B, Ray::new runs inside the kernel).
// kind 2: ObvhsRayOd[n] (24 B: origin, direction) with ONE (tmin, tmax) for the batch.
struct RayFormat {
    u32 kind;
    float tmin, tmax;
    u32 hit8;  // closest-hit results as ObvhsRayHit8 {primitive_id, t} (8 B) instead of ObvhsRayHit (16 B)
    __host__ __device__ size_t bytes() const { return kind == 0 ? 64 : kind == 1 ? 32 : 24; }
};
int cwbvh_traverse_device(ObvhsContext* ctx, const ObvhsCwBvh* bvh, const void* d_rays, const RayFormat& fmt, size_t n, int mode, void* d_out,
                          u64* d_counters);
int cwbvh_permute_tris_device(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsTriangle* d_tris, size_t n);
int bvh2_traverse_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, const void* d_rays, const RayFormat& fmt, size_t n, int mode, void* d_out,
                         u64* d_counters);
int bvh2_permute_tris_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsTriangle* d_tris, size_t n_tris);
size_t traverse_host_chunk_min(const ObvhsContext* ctx, size_t prim_count, bool* persistent);
int make_rays_device(ObvhsContext* ctx, const float* d_od, size_t n, float tmin, float tmax, ObvhsRay* d_rays);
