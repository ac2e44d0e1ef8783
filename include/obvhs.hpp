// obvhs.hpp -- C++17 host side above the C ABI (include/obvhs_cuda.h), header-only.
//
// The reference is a compiled (Rust) library with no FFI layer; its toolchain is absent here, so this is the compiled-language
// mirror of its public interface for the hot path: the same names, argument order and meaning as the Rust items cited on each
// declaration (paths relative to the reference checkout), RAII instead of Drop, exceptions where the reference panics.
// Every array argument may be host or device memory (`const T*` + count). There is no CPU fallback: Context's constructor
// throws when no CUDA device / library is present. The Python mirror (obvhs_b200/api.py) binds the same C ABI.
#ifndef OBVHS_HPP
#define OBVHS_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "obvhs_cuda.h"

namespace obvhs {

using Aabb = ObvhsAabb;              // src/aabb.rs:13-16
using Triangle = ObvhsTriangle;      // src/triangle.rs:9-13
using Bvh2Node = ObvhsBvh2Node;      // src/bvh2/node.rs:40-66
using CwBvhNode = ObvhsCwBvhNode;    // src/cwbvh/node.rs:14-54
using Ray = ObvhsRay;                // src/ray.rs:15-30
using RayNew = ObvhsRayNew;          // the arguments of Ray::new, src/ray.rs:34
using RayOd = ObvhsRayOd;            // origin + direction of Ray::new_inf, src/ray.rs:55-57 (one tmin/tmax per batch)
using RayHit = ObvhsRayHit;          // src/ray.rs:63-70
using RayHit8 = ObvhsRayHit8;        // {primitive_id, t}: what the triangle closure writes into a RayHit
constexpr uint32_t INVALID_ID = 0xffffffffu;  // src/ray.rs:72

// the reference panics; the C ABI returns a status; this layer throws
struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// src/ray.rs:6-12
inline float safe_inverse(float x) {
    constexpr float EPS = 1.1920929e-07f;
    return std::fabs(x) <= EPS ? std::copysign(1.0f, x) / EPS : 1.0f / x;
}
// Ray::new(origin, direction, min, max), src/ray.rs:34-52
inline Ray ray_new(const float origin[3], const float direction[3], float tmin, float tmax) {
    Ray r;
    std::memset(&r, 0, sizeof(r));
    for (int k = 0; k < 3; k++) {
        r.origin[k] = origin[k];
        r.direction[k] = direction[k];
        r.inv_direction[k] = safe_inverse(direction[k]);
    }
    r.tmin = tmin;
    r.tmax = tmax;
    return r;
}
// Ray::new_inf, src/ray.rs:55-57
inline Ray ray_new_inf(const float origin[3], const float direction[3]) { return ray_new(origin, direction, 0.0f, INFINITY); }

// src/ploc/mod.rs:534-562
enum class PlocSearchDistance : uint32_t { Minimum = 1, VeryLow = 2, Low = 6, Medium = 14, High = 24, VeryHigh = 32 };
// src/ploc/mod.rs:658-661
enum class SortPrecision : uint32_t { U64 = 64, U128 = 128 };

// src/lib.rs:208-305
struct BvhBuildParams : ObvhsBuildParams {
    static BvhBuildParams preset(const char* name) {
        BvhBuildParams p;
        if (obvhs_cuda_build_params_preset(name, &p) != OBVHS_OK) throw Error(OBVHS_ERR_INVALID_ARG, std::string("unknown preset ") + name);
        return p;
    }
    static BvhBuildParams fastest_build() { return preset("fastest_build"); }
    static BvhBuildParams very_fast_build() { return preset("very_fast_build"); }
    static BvhBuildParams fast_build() { return preset("fast_build"); }
    static BvhBuildParams medium_build() { return preset("medium_build"); }
    static BvhBuildParams slow_build() { return preset("slow_build"); }
    static BvhBuildParams very_slow_build() { return preset("very_slow_build"); }
};

// One per GPU (and per CUDA stream): owns the scratch arena, the result cache and the copy streams of the library.
class Context {
public:
    explicit Context(int device = 0, void* cuda_stream = nullptr) {
        ObvhsContext* h = nullptr;
        int rc = obvhs_cuda_create(device, cuda_stream, &h);
        if (rc != OBVHS_OK || !h) throw Error(rc, "obvhs_cuda_create failed: no CUDA device or driver (there is no CPU fallback)");
        h_.reset(h, obvhs_cuda_destroy);
    }
    ObvhsContext* get() const { return h_.get(); }
    void check(int rc) const {
        if (rc != OBVHS_OK) throw Error(rc, obvhs_cuda_last_error(h_.get()));
    }
    void synchronize() const { check(obvhs_cuda_synchronize(h_.get())); }
    // Multi-GPU (one process per GPU; no equivalent in the reference, whose CwBvh is cloned freely): `nccl_unique_id` on one rank,
    // the 128 bytes shipped to the others by the caller, then `comm_init` on every rank (collective). See CwBvh::broadcast.
    static std::array<uint8_t, OBVHS_NCCL_UNIQUE_ID_BYTES> nccl_unique_id() {
        std::array<uint8_t, OBVHS_NCCL_UNIQUE_ID_BYTES> id{};
        int rc = obvhs_cuda_nccl_unique_id(id.data());
        if (rc != OBVHS_OK) throw Error(rc, "obvhs_cuda_nccl_unique_id failed (libnccl.so.2 not loadable?)");
        return id;
    }
    void comm_init(const std::array<uint8_t, OBVHS_NCCL_UNIQUE_ID_BYTES>& id, int rank, int world) const {
        check(obvhs_cuda_comm_init(h_.get(), id.data(), rank, world));
    }
    uint64_t launch_count() const { return obvhs_cuda_launch_count(h_.get()); }
    // "traverse": auto | static | persistent[:refill[:chunk]]; "host_slice": rays per pipelined slice; "trace": 0 | 1
    void set_option(const char* key, const char* value) const {
        int rc = obvhs_cuda_set_option(h_.get(), key, value);
        if (rc != OBVHS_OK) throw Error(rc, obvhs_cuda_last_error(h_.get()));
    }

private:
    std::shared_ptr<ObvhsContext> h_;
};

// result of the batched broad-phase queries: counts[i] reports for query i, ids of query i start at sum(counts[0..i))
struct QueryResult {
    std::vector<uint32_t> counts, ids;
};

namespace detail {
template <class Call>
QueryResult query(const Context& ctx, size_t n, Call call) {
    QueryResult r;
    r.counts.resize(n);
    size_t total = 0;
    ctx.check(call(r.counts.data(), (uint32_t*)nullptr, (size_t)0, &total));  // count pass
    r.ids.resize(total);
    if (total) ctx.check(call(r.counts.data(), r.ids.data(), total, &total));
    return r;
}
}  // namespace detail

// Device-resident Bvh2, src/bvh2/mod.rs:31-85
class Bvh2 {
public:
    Bvh2(Context ctx, ObvhsBvh2* h) : ctx_(std::move(ctx)), h_(h, obvhs_cuda_bvh2_free) {}
    // from host arrays (e.g. a tree built by the reference itself)
    static Bvh2 upload(const Context& ctx, const Bvh2Node* nodes, size_t node_count, const uint32_t* primitive_indices, size_t prim_count,
                       size_t max_depth, bool children_are_ordered_after_parents) {
        ObvhsBvh2* h = nullptr;
        ctx.check(obvhs_cuda_bvh2_upload(ctx.get(), nodes, node_count, primitive_indices, prim_count, max_depth,
                                         children_are_ordered_after_parents, &h));
        return Bvh2(ctx, h);
    }
    ObvhsBvh2* get() const { return h_.get(); }
    const Context& context() const { return ctx_; }
    size_t node_count() const { return obvhs_cuda_bvh2_node_count(h_.get()); }
    size_t prim_count() const { return obvhs_cuda_bvh2_prim_count(h_.get()); }
    size_t max_depth() const { return obvhs_cuda_bvh2_max_depth(h_.get()); }
    bool children_are_ordered_after_parents() const { return obvhs_cuda_bvh2_children_ordered_after_parents(h_.get()) != 0; }
    bool uses_spatial_splits() const { return obvhs_cuda_bvh2_uses_spatial_splits(h_.get()) != 0; }
    // nodes / primitive_indices / parents (src/bvh2/mod.rs:33-46) as host vectors; parents only when computed and asked for
    void download(std::vector<Bvh2Node>* nodes, std::vector<uint32_t>* primitive_indices, std::vector<uint32_t>* parents = nullptr) const {
        if (nodes) nodes->resize(node_count());
        if (primitive_indices) primitive_indices->resize(prim_count());
        if (parents) parents->resize(node_count());
        ctx_.check(obvhs_cuda_bvh2_download(ctx_.get(), h_.get(), nodes ? nodes->data() : nullptr,
                                            primitive_indices ? primitive_indices->data() : nullptr, parents ? parents->data() : nullptr));
    }
    void compute_parents() { ctx_.check(obvhs_cuda_bvh2_compute_parents(ctx_.get(), h_.get())); }  // src/bvh2/mod.rs:586-619
    // Bvh2::reorder_in_stack_traversal_order, src/bvh2/mod.rs:462-500
    void reorder_in_stack_traversal_order() { ctx_.check(obvhs_cuda_bvh2_reorder_in_stack_traversal_order(ctx_.get(), h_.get())); }
    void refit_all() { ctx_.check(obvhs_cuda_bvh2_refit_all(ctx_.get(), h_.get())); }              // src/bvh2/mod.rs:527-569
    // rewrites every leaf box from per-primitive boxes and refits (the update loop of examples/physics.rs)
    void set_leaf_aabbs(const Aabb* prim_aabbs, size_t n) { ctx_.check(obvhs_cuda_bvh2_set_leaf_aabbs(ctx_.get(), h_.get(), prim_aabbs, n)); }
    // resize_node for a batch of nodes, src/bvh2/mod.rs:755-761
    void set_node_aabbs(const uint32_t* node_ids, const Aabb* aabbs, size_t n) {
        ctx_.check(obvhs_cuda_bvh2_set_node_aabbs(ctx_.get(), h_.get(), node_ids, aabbs, n));
    }
    // collapse(&mut bvh, max_prims, traversal_cost), src/bvh2/leaf_collapser.rs:21
    void collapse(uint32_t max_prims, float traversal_cost) { ctx_.check(obvhs_cuda_bvh2_collapse(ctx_.get(), h_.get(), max_prims, traversal_cost)); }
    // the triangles the fixed intersection closure tests (permuted by primitive_indices inside the handle)
    void set_triangles(const Triangle* tris, size_t n) { ctx_.check(obvhs_cuda_bvh2_set_triangles(ctx_.get(), h_.get(), tris, n)); }
    // Bvh2::ray_traverse / ray_traverse_miss / ray_traverse_anyhit over a batch, src/bvh2/mod.rs:148-238
    void ray_traverse(const Ray* rays, size_t n, RayHit* hits) const { ctx_.check(obvhs_cuda_bvh2_ray_traverse_batch(ctx_.get(), h_.get(), rays, n, hits)); }
    void ray_traverse(const RayNew* args, size_t n, RayHit* hits) const {
        ctx_.check(obvhs_cuda_bvh2_ray_new_traverse_batch(ctx_.get(), h_.get(), args, n, hits));
    }
    void ray_traverse(const RayOd* od, size_t n, RayHit* hits, float tmin = 0.0f, float tmax = INFINITY) const {
        ctx_.check(obvhs_cuda_bvh2_ray_od_traverse_batch(ctx_.get(), h_.get(), od, n, tmin, tmax, hits));
    }
    void ray_traverse_miss(const Ray* rays, size_t n, uint8_t* miss) const {
        ctx_.check(obvhs_cuda_bvh2_ray_traverse_miss_batch(ctx_.get(), h_.get(), rays, n, miss));
    }
    void ray_traverse_miss(const RayNew* args, size_t n, uint8_t* miss) const {
        ctx_.check(obvhs_cuda_bvh2_ray_new_traverse_miss_batch(ctx_.get(), h_.get(), args, n, miss));
    }
    void ray_traverse_anyhit_count(const Ray* rays, size_t n, uint32_t* counts) const {
        ctx_.check(obvhs_cuda_bvh2_ray_traverse_anyhit_count_batch(ctx_.get(), h_.get(), rays, n, counts));
    }
    // Bvh2::aabb_traverse / point_traverse with an eval that always continues, src/bvh2/mod.rs:365-456: reported LEAF NODE ids
    QueryResult aabb_traverse(const Aabb* queries, size_t n) const {
        return detail::query(ctx_, n, [&](uint32_t* c, uint32_t* ids, size_t cap, size_t* total) {
            return obvhs_cuda_bvh2_aabb_traverse_batch(ctx_.get(), h_.get(), queries, n, c, ids, cap, total);
        });
    }
    QueryResult point_traverse(const float* points_xyzw, size_t n) const {
        return detail::query(ctx_, n, [&](uint32_t* c, uint32_t* ids, size_t cap, size_t* total) {
            return obvhs_cuda_bvh2_point_traverse_batch(ctx_.get(), h_.get(), points_xyzw, n, c, ids, cap, total);
        });
    }

private:
    Context ctx_;
    std::shared_ptr<ObvhsBvh2> h_;
};

// Device-resident CwBvh, src/cwbvh/mod.rs:43-55
class CwBvh {
public:
    CwBvh(Context ctx, ObvhsCwBvh* h) : ctx_(std::move(ctx)), h_(h, obvhs_cuda_cwbvh_free) {}
    static CwBvh upload(const Context& ctx, const CwBvhNode* nodes, size_t node_count, const uint32_t* primitive_indices, size_t prim_count,
                        const Aabb& total_aabb) {
        ObvhsCwBvh* h = nullptr;
        ctx.check(obvhs_cuda_cwbvh_upload(ctx.get(), nodes, node_count, primitive_indices, prim_count, &total_aabb, &h));
        return CwBvh(ctx, h);
    }
    ObvhsCwBvh* get() const { return h_.get(); }
    const Context& context() const { return ctx_; }
    size_t node_count() const { return obvhs_cuda_cwbvh_node_count(h_.get()); }
    size_t prim_count() const { return obvhs_cuda_cwbvh_prim_count(h_.get()); }
    bool uses_spatial_splits() const { return obvhs_cuda_cwbvh_uses_spatial_splits(h_.get()) != 0; }
    void download(std::vector<CwBvhNode>* nodes, std::vector<uint32_t>* primitive_indices, Aabb* total_aabb = nullptr) const {
        if (nodes) nodes->resize(node_count());
        if (primitive_indices) primitive_indices->resize(prim_count());
        ctx_.check(obvhs_cuda_cwbvh_download(ctx_.get(), h_.get(), nodes ? nodes->data() : nullptr,
                                             primitive_indices ? primitive_indices->data() : nullptr, total_aabb));
    }
    // CwBvh::compute_parents, src/cwbvh/mod.rs:494-509
    std::vector<uint32_t> compute_parents() const {
        std::vector<uint32_t> parents(node_count());
        ctx_.check(obvhs_cuda_cwbvh_compute_parents(ctx_.get(), h_.get(), parents.data()));
        return parents;
    }
    // CwBvh::exact_node_aabbs, src/cwbvh/mod.rs:47 (empty when the tree was converted without them)
    std::vector<Aabb> exact_node_aabbs() const {
        size_t count = 0;
        ctx_.check(obvhs_cuda_cwbvh_exact_node_aabbs(ctx_.get(), h_.get(), nullptr, 0, &count));
        std::vector<Aabb> out(count);
        if (count) ctx_.check(obvhs_cuda_cwbvh_exact_node_aabbs(ctx_.get(), h_.get(), out.data(), count, &count));
        return out;
    }
    void set_triangles(const Triangle* tris, size_t n) { ctx_.check(obvhs_cuda_cwbvh_set_triangles(ctx_.get(), h_.get(), tris, n)); }
    // CwBvh::order_children(&mut self, primitives, direct_layout), src/cwbvh/mod.rs:520-524, the primitives given as their AABBs
    void order_children(const Aabb* prim_aabbs, size_t n, bool direct_layout) {
        ctx_.check(obvhs_cuda_cwbvh_order_children(ctx_.get(), h_.get(), prim_aabbs, n, direct_layout));
    }
    // Collective over the ranks of ctx.comm_init: `root_tree` is the finished tree on rank `root` (returned as is) and null
    // elsewhere, where a replica is returned (nodes, primitive_indices, total_aabb, flags, permuted triangles). One 64-byte header
    // broadcast + one grouped NCCL launch on the context's stream. (Re-filling an earlier replica in place is offered by the C
    // entry point only: it may free the handle it is given, which a shared handle cannot express.)
    static CwBvh broadcast(const Context& ctx, const CwBvh* root_tree, int root) {
        ObvhsCwBvh* h = root_tree ? root_tree->get() : nullptr;
        ctx.check(obvhs_cuda_cwbvh_broadcast(ctx.get(), &h, root));
        if (root_tree && h == root_tree->get()) return *root_tree;
        return CwBvh(ctx, h);
    }
    // CwBvh::ray_traverse / ray_traverse_miss / ray_traverse_anyhit over a batch with the triangle closure,
    // src/cwbvh/mod.rs:169-245; hit.primitive_id indexes primitive_indices order, as in the reference
    void ray_traverse(const Ray* rays, size_t n, RayHit* hits) const { ctx_.check(obvhs_cuda_cwbvh_ray_traverse_batch(ctx_.get(), h_.get(), rays, n, hits)); }
    void ray_traverse(const RayNew* args, size_t n, RayHit* hits) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_new_traverse_batch(ctx_.get(), h_.get(), args, n, hits));
    }
    // rays[i] = Ray::new(od[i].origin, od[i].direction, tmin, tmax); defaults = Ray::new_inf (src/ray.rs:55-57)
    void ray_traverse(const RayOd* od, size_t n, RayHit* hits, float tmin = 0.0f, float tmax = INFINITY) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_od_traverse_batch(ctx_.get(), h_.get(), od, n, tmin, tmax, hits));
    }
    void ray_traverse(const RayOd* od, size_t n, RayHit8* hits, float tmin = 0.0f, float tmax = INFINITY) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch(ctx_.get(), h_.get(), od, n, tmin, tmax, hits));
    }
    void ray_traverse_miss(const RayOd* od, size_t n, uint8_t* miss, float tmin = 0.0f, float tmax = INFINITY) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_od_traverse_miss_batch(ctx_.get(), h_.get(), od, n, tmin, tmax, miss));
    }
    void ray_traverse_miss(const Ray* rays, size_t n, uint8_t* miss) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_traverse_miss_batch(ctx_.get(), h_.get(), rays, n, miss));
    }
    void ray_traverse_miss(const RayNew* args, size_t n, uint8_t* miss) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_new_traverse_miss_batch(ctx_.get(), h_.get(), args, n, miss));
    }
    void ray_traverse_anyhit_count(const Ray* rays, size_t n, uint32_t* counts) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_traverse_anyhit_count_batch(ctx_.get(), h_.get(), rays, n, counts));
    }
    void ray_traverse_anyhit_count(const RayNew* args, size_t n, uint32_t* counts) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_new_traverse_anyhit_count_batch(ctx_.get(), h_.get(), args, n, counts));
    }
    // closest hit plus the per-launch totals the roofline model needs: counters[0] += nodes visited, [1] += triangles tested
    void ray_traverse_counted(const Ray* rays, size_t n, RayHit* hits, uint64_t counters[2]) const {
        ctx_.check(obvhs_cuda_cwbvh_ray_traverse_batch_counted(ctx_.get(), h_.get(), rays, n, hits, counters));
    }
    // traverse!(.., node.intersect_aabb / contains_point(.., state.oct_inv4), ..), src/cwbvh/node.rs:157-200: primitive SLOTS
    QueryResult aabb_traverse(const Aabb* queries, size_t n, const float traversal_direction[3]) const {
        return detail::query(ctx_, n, [&](uint32_t* c, uint32_t* ids, size_t cap, size_t* total) {
            return obvhs_cuda_cwbvh_aabb_traverse_batch(ctx_.get(), h_.get(), queries, n, traversal_direction, c, ids, cap, total);
        });
    }
    QueryResult point_traverse(const float* points_xyzw, size_t n, const float traversal_direction[3]) const {
        return detail::query(ctx_, n, [&](uint32_t* c, uint32_t* ids, size_t cap, size_t* total) {
            return obvhs_cuda_cwbvh_point_traverse_batch(ctx_.get(), h_.get(), points_xyzw, n, traversal_direction, c, ids, cap, total);
        });
    }

private:
    Context ctx_;
    std::shared_ptr<ObvhsCwBvh> h_;
};

// src/ploc/mod.rs:35-137 and src/ploc/rebuild.rs
class PlocBuilder {
public:
    explicit PlocBuilder(Context ctx) : ctx_(std::move(ctx)) {}
    // build(search_distance, aabbs, indices, sort_precision, search_depth_threshold), src/ploc/mod.rs:95-102;
    // indices may be null (= 0..n)
    Bvh2 build(PlocSearchDistance search_distance, const Aabb* aabbs, size_t n, const uint32_t* indices, SortPrecision sort_precision,
               size_t search_depth_threshold) const {
        ObvhsBvh2* h = nullptr;
        ctx_.check(obvhs_cuda_ploc_build(ctx_.get(), aabbs, indices, n, (uint32_t)search_distance, (uint32_t)sort_precision, search_depth_threshold, &h));
        return Bvh2(ctx_, h);
    }
    // the same with T = Triangle (Boundable::aabb computed on the device)
    Bvh2 build(PlocSearchDistance search_distance, const Triangle* tris, size_t n, SortPrecision sort_precision, size_t search_depth_threshold) const {
        ObvhsBvh2* h = nullptr;
        ctx_.check(obvhs_cuda_ploc_build_tris(ctx_.get(), tris, n, (uint32_t)search_distance, (uint32_t)sort_precision, search_depth_threshold, &h));
        return Bvh2(ctx_, h);
    }
    // src/ploc/rebuild.rs:56-80
    void full_rebuild(Bvh2& bvh, PlocSearchDistance search_distance, SortPrecision sort_precision, size_t search_depth_threshold) const {
        ctx_.check(obvhs_cuda_ploc_full_rebuild(ctx_.get(), bvh.get(), (uint32_t)search_distance, (uint32_t)sort_precision, search_depth_threshold));
    }
    // src/ploc/rebuild.rs:101-135; should_remove holds one flag per node (the reference's closure, evaluated up front)
    void partial_rebuild(Bvh2& bvh, const uint8_t* should_remove, PlocSearchDistance search_distance, SortPrecision sort_precision,
                         size_t search_depth_threshold) const {
        ctx_.check(obvhs_cuda_ploc_partial_rebuild(ctx_.get(), bvh.get(), should_remove, (uint32_t)search_distance, (uint32_t)sort_precision,
                                                   search_depth_threshold));
    }

private:
    Context ctx_;
};

// Bvh2::compute_primitives_to_nodes(nodes, primitive_indices, &mut primitives_to_nodes), src/bvh2/mod.rs:647-665, over downloaded
// arrays (host-side bookkeeping): the leaf node holding each primitive id, INVALID_ID when none
inline std::vector<uint32_t> compute_primitives_to_nodes(const std::vector<Bvh2Node>& nodes, const std::vector<uint32_t>& primitive_indices) {
    std::vector<uint32_t> out(primitive_indices.size(), INVALID_ID);
    for (size_t node_id = 0; node_id < nodes.size(); node_id++)
        for (uint32_t k = nodes[node_id].first_index, end = k + nodes[node_id].prim_count; nodes[node_id].prim_count != 0 && k < end; k++)
            out[primitive_indices[k]] = (uint32_t)node_id;
    return out;
}

// compute_rebuild_path_flags(bvh, leaves, flags), src/ploc/rebuild.rs:12-43
inline std::vector<uint8_t> compute_rebuild_path_flags(const Bvh2& bvh, const uint32_t* leaves, size_t n_leaves) {
    std::vector<uint8_t> flags(bvh.node_count());
    bvh.context().check(obvhs_cuda_compute_rebuild_path_flags(bvh.context().get(), bvh.get(), leaves, n_leaves, flags.data()));
    return flags;
}

// src/bvh2/reinsertion.rs:22-57
class ReinsertionOptimizer {
public:
    explicit ReinsertionOptimizer(Context ctx) : ctx_(std::move(ctx)) {}
    // run(&mut bvh, batch_size_ratio, ratio_sequence); returns the number of reinsertions applied
    uint64_t run(Bvh2& bvh, float batch_size_ratio, const std::vector<float>* ratio_sequence = nullptr) const {
        uint64_t applied = 0;
        ctx_.check(obvhs_cuda_reinsertion_run(ctx_.get(), bvh.get(), batch_size_ratio, ratio_sequence ? ratio_sequence->data() : nullptr,
                                              ratio_sequence ? ratio_sequence->size() : 0, &applied));
        return applied;
    }
    // run_with_candidates(&mut bvh, candidates, iterations), src/bvh2/reinsertion.rs:62-90
    uint64_t run_with_candidates(Bvh2& bvh, const uint32_t* node_ids, size_t n, uint32_t iterations) const {
        uint64_t applied = 0;
        ctx_.check(obvhs_cuda_reinsertion_run_with_candidates(ctx_.get(), bvh.get(), node_ids, n, iterations, &applied));
        return applied;
    }

private:
    Context ctx_;
};

// bvh2_to_cwbvh(&bvh2, max_prims_per_leaf, order_children, include_exact_node_aabbs), src/cwbvh/bvh2_to_cwbvh.rs:490-510
inline CwBvh bvh2_to_cwbvh(const Bvh2& bvh2, uint32_t max_prims_per_leaf, bool order_children, bool include_exact_node_aabbs) {
    ObvhsCwBvh* h = nullptr;
    bvh2.context().check(obvhs_cuda_bvh2_to_cwbvh(bvh2.context().get(), bvh2.get(), max_prims_per_leaf, order_children, include_exact_node_aabbs, &h));
    return CwBvh(bvh2.context(), h);
}

// build_cwbvh_from_tris(triangles, config, core_build_time), src/cwbvh/builder.rs:20-85. *core_build_seconds is incremented
// like the reference's &mut Duration. The result carries the permuted triangles and is ready to traverse.
inline CwBvh build_cwbvh_from_tris(const Context& ctx, const Triangle* tris, size_t n, const BvhBuildParams& config, double* core_build_seconds = nullptr) {
    ObvhsCwBvh* h = nullptr;
    double secs = 0.0;
    ctx.check(obvhs_cuda_build_cwbvh_from_tris(ctx.get(), tris, n, &config, &secs, &h));
    if (core_build_seconds) *core_build_seconds += secs;
    return CwBvh(ctx, h);
}
// build_cwbvh<T: Boundable>(primitives, config, core_build_time) over the primitives' boxes, src/cwbvh/builder.rs:98-123
inline CwBvh build_cwbvh(const Context& ctx, const Aabb* aabbs, size_t n, const BvhBuildParams& config, double* core_build_seconds = nullptr) {
    ObvhsCwBvh* h = nullptr;
    double secs = 0.0;
    ctx.check(obvhs_cuda_build_cwbvh(ctx.get(), aabbs, n, &config, &secs, &h));
    if (core_build_seconds) *core_build_seconds += secs;
    return CwBvh(ctx, h);
}
// build_bvh2_from_tris, src/bvh2/builder.rs:17-91
inline Bvh2 build_bvh2_from_tris(const Context& ctx, const Triangle* tris, size_t n, const BvhBuildParams& config, double* core_build_seconds = nullptr) {
    ObvhsBvh2* h = nullptr;
    double secs = 0.0;
    ctx.check(obvhs_cuda_build_bvh2_from_tris(ctx.get(), tris, n, &config, &secs, &h));
    if (core_build_seconds) *core_build_seconds += secs;
    return Bvh2(ctx, h);
}
// build_bvh2<T: Boundable>, src/bvh2/builder.rs:103-140
inline Bvh2 build_bvh2(const Context& ctx, const Aabb* aabbs, size_t n, const BvhBuildParams& config, double* core_build_seconds = nullptr) {
    ObvhsBvh2* h = nullptr;
    double secs = 0.0;
    ctx.check(obvhs_cuda_build_bvh2(ctx.get(), aabbs, n, &config, &secs, &h));
    if (core_build_seconds) *core_build_seconds += secs;
    return Bvh2(ctx, h);
}

// split_aabbs_preset(&mut aabbs, &mut indices, triangles, avg_half_area, largest_half_area), src/splits.rs:16-47: the two
// vectors grow exactly as the reference's do
inline void split_aabbs_preset(const Context& ctx, std::vector<Aabb>& aabbs, std::vector<uint32_t>& indices, const Triangle* tris, size_t n_tris,
                               float avg_half_area, float largest_half_area) {
    const size_t n = aabbs.size();
    size_t capacity = n + n / 2 + 16, count = 0;
    for (;;) {
        aabbs.resize(capacity);
        indices.resize(capacity);
        int rc = obvhs_cuda_split_aabbs_preset(ctx.get(), aabbs.data(), indices.data(), n, capacity, tris, n_tris, avg_half_area, largest_half_area, &count);
        if (rc == OBVHS_ERR_CAPACITY) {  // nothing was written: call again with the room it asked for
            aabbs.resize(n);
            indices.resize(n);
            capacity = count;
            continue;
        }
        ctx.check(rc);
        break;
    }
    aabbs.resize(count);
    indices.resize(count);
}

// Ray::new over a batch on the device, src/ray.rs:34-52
inline void ray_new_batch(const Context& ctx, const RayNew* args, size_t n, Ray* rays) { ctx.check(obvhs_cuda_ray_new_batch(ctx.get(), args, n, rays)); }

}  // namespace obvhs

#endif  // OBVHS_HPP
