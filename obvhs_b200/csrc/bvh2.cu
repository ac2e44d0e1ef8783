// bvh2.cu -- Bvh2 container operations on the device: compute_parents, refit_all, layout conversion.
//
//   compute_parents   src/bvh2/mod.rs:586-619   parents[first] = parents[first+1] = i for every inner node i
//   refit_all         src/bvh2/mod.rs:527-569   every inner node = first.union(second), children before parents.
//                     The reference sweeps indices in reverse (or a stack order); any bottom-up order gives the same
//                     bits because min/max are exact and the operand order (first, second) is fixed. Here: one thread
//                     per leaf climbs with an arrival counter per inner node; the second arriver refits the node.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) compute_parents_kernel(const Node32* __restrict__ nodes, u32 n, u32* __restrict__ parents) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) parents[0] = 0;
    float4 a = __ldg(reinterpret_cast<const float4*>(nodes + i));
    if (__float_as_uint(a.w) == 0) {
        u32 f = __float_as_uint(__ldg(reinterpret_cast<const float4*>(nodes + i) + 1).w);
        parents[f] = i;
        parents[f + 1] = i;
    }
}

__global__ void __launch_bounds__(256) refit_bottom_up_kernel(Node32* nodes, u32 n, const u32* __restrict__ parents, u32* arrivals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    if (__ldcg(&nodes[i].prim_count) == 0) return;  // start at leaves only
    u32 p = parents[i];
    for (;;) {
        __threadfence();
        if (atomicAdd(&arrivals[p], 1u) == 0) return;  // first arriver: the sibling subtree is not finished yet
        Node32 me = load_node_cg(nodes + p);
        Node32 c0 = load_node_cg(nodes + me.first_index), c1 = load_node_cg(nodes + me.first_index + 1);
        Box u = box_union(node_box(c0), node_box(c1));  // first.union(second), bvh2/mod.rs:537-539
        store_node(nodes + p, make_node32(u, me.prim_count, me.first_index));
        if (p == 0) return;
        p = parents[p];
    }
}

__global__ void __launch_bounds__(256) set_leaf_aabbs_kernel(Node32* nodes, u32 n, const u32* __restrict__ prim_idx,
                                                             const float4* __restrict__ aabbs) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Node32 nd = load_node(nodes + i);
    if (nd.prim_count == 0) return;
    u32 p = prim_idx[nd.first_index];
    float4 lo = __ldg(aabbs + (size_t)p * 2), hi = __ldg(aabbs + (size_t)p * 2 + 1);
    Box b{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
    for (u32 k = 1; k < nd.prim_count; k++) {
        u32 q = prim_idx[nd.first_index + k];
        float4 l2 = __ldg(aabbs + (size_t)q * 2), h2 = __ldg(aabbs + (size_t)q * 2 + 1);
        b = box_union(b, Box{l2.x, l2.y, l2.z, h2.x, h2.y, h2.z});
    }
    store_node(nodes + i, make_node32(b, nd.prim_count, nd.first_index));
}

// 32-byte device node <-> 48-byte Bvh2Node (bvh2/node.rs:40-66)
__global__ void __launch_bounds__(256) expand_nodes_kernel(const Node32* __restrict__ in, u32 n, float4* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Node32 nd = load_node(in + i);
    out[(size_t)i * 3 + 0] = make_float4(nd.minx, nd.miny, nd.minz, 0.f);
    out[(size_t)i * 3 + 1] = make_float4(nd.maxx, nd.maxy, nd.maxz, 0.f);
    out[(size_t)i * 3 + 2] = make_float4(__uint_as_float(nd.prim_count), __uint_as_float(nd.first_index), 0.f, 0.f);
}
__global__ void __launch_bounds__(256) pack_nodes_kernel(const float4* __restrict__ in, u32 n, Node32* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = in[(size_t)i * 3], b = in[(size_t)i * 3 + 1], c = in[(size_t)i * 3 + 2];
    store_node(out + i, make_node32(Box{a.x, a.y, a.z, b.x, b.y, b.z}, __float_as_uint(c.x), __float_as_uint(c.y)));
}

// Bvh2Node::set_aabb (bvh2/node.rs:103-107) for a list of nodes; prim_count / first_index are kept
__global__ void __launch_bounds__(256) set_node_aabbs_kernel(Node32* nodes, u32 n_nodes, const u32* __restrict__ ids, const float4* __restrict__ aabbs,
                                                             u32 n) {
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const u32 id = ids[k];
    if (id >= n_nodes) return;
    Node32 nd = load_node(nodes + id);
    const float4 lo = aabbs[2 * (size_t)k], hi = aabbs[2 * (size_t)k + 1];
    nd.minx = lo.x; nd.miny = lo.y; nd.minz = lo.z;
    nd.maxx = hi.x; nd.maxy = hi.y; nd.maxz = hi.z;
    store_node(nodes + id, nd);
}

// ---- Bvh2::reorder_in_stack_traversal_order (src/bvh2/mod.rs:462-500) --------------------------------------------------------
// The reference pops sibling pairs off a stack that receives the FIRST child's pair before the SECOND's, so a pair is followed
// by the whole subtree under its second node, then by the subtree under its first node; the k-th popped pair lands at new
// indices 2k+1, 2k+2 (the root keeps index 0). With pairs(x) = number of sibling pairs below node x (= inner nodes of its
// subtree), the pair below the second node of pair k is pair k + 1 and the pair below its first node is pair
// k + 1 + pairs(second). Two passes: pairs() bottom-up with arrival counters (as refit_all), then a top-down breadth-first
// sweep that carries every pair's rank and writes both nodes -- with their first_index already remapped -- to their final
// slots. One cooperative launch, a grid barrier per tree level.
__global__ void __launch_bounds__(256) subtree_pairs_kernel(const Node32* __restrict__ nodes, u32 n, const u32* __restrict__ parents, u32* arrivals,
                                                            u32* pairs) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    if (__float_as_uint(__ldg(reinterpret_cast<const float4*>(nodes + i)).w) == 0) return;  // start at leaves (pairs = 0, pre-zeroed)
    u32 p = parents[i];
    for (;;) {
        __threadfence();
        if (atomicAdd(&arrivals[p], 1u) == 0) return;  // first arriver: the sibling subtree is not finished yet
        const u32 first = __float_as_uint(__ldg(reinterpret_cast<const float4*>(nodes + p) + 1).w);
        pairs[p] = 1u + __ldcg(&pairs[first]) + __ldcg(&pairs[first + 1]);
        if (p == 0) return;
        p = parents[p];
    }
}

struct ReorderArgs {
    const Node32* in;
    Node32* out;
    const u32* pairs;
    uint2* queue[2];  // (old index of the pair's first node, rank of the pair)
    u32* qcount;      // [3], slot = level % 3
};
__global__ void __launch_bounds__(256) reorder_bfs_kernel(ReorderArgs a) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const u32 lane = threadIdx.x & 31u;
    if (tid == 0) {
        Node32 root = load_node(a.in);
        u32 n0 = 0;
        if (root.prim_count == 0) {
            a.queue[0][0] = make_uint2(root.first_index, 0u);
            root.first_index = 1;  // mapping[first_index] of the first popped pair
            n0 = 1;
        }
        store_node(a.out, root);
        a.qcount[0] = n0;
        a.qcount[1] = 0;
        a.qcount[2] = 0;
    }
    grid.sync();
    for (u32 level = 0;; level++) {
        const u32 n = __ldcg(&a.qcount[level % 3]);
        if (n == 0) break;
        const uint2* q = a.queue[level & 1];
        uint2* qn = a.queue[(level + 1) & 1];
        u32* qn_count = &a.qcount[(level + 1) % 3];
        for (u32 base = blockIdx.x * blockDim.x; base < n; base += nthreads) {  // CTA-uniform trip count (warp-aggregated pushes)
            const u32 idx = base + threadIdx.x;
            uint2 push_a = make_uint2(0, 0), push_b = make_uint2(0, 0);
            bool has_a = false, has_b = false;
            if (idx < n) {
                const uint2 e = __ldcg(q + idx);
                const u32 cur = e.x, r = e.y;
                Node32 na = load_node(a.in + cur), nb = load_node(a.in + cur + 1);
                const u32 pb = nb.prim_count == 0 ? __ldg(a.pairs + cur + 1) : 0u;
                if (na.prim_count == 0) {
                    has_a = true;
                    push_a = make_uint2(na.first_index, r + 1u + pb);
                    na.first_index = 2u * push_a.y + 1u;
                }
                if (nb.prim_count == 0) {
                    has_b = true;
                    push_b = make_uint2(nb.first_index, r + 1u);
                    nb.first_index = 2u * push_b.y + 1u;
                }
                store_node(a.out + 2u * r + 1u, na);
                store_node(a.out + 2u * r + 2u, nb);
            }
            const u32 ba = __ballot_sync(0xffffffffu, has_a), bb = __ballot_sync(0xffffffffu, has_b);
            const u32 total = __popc(ba) + __popc(bb);
            u32 wbase = 0;
            if (lane == 0 && total) wbase = atomicAdd(qn_count, total);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            const u32 lt = (1u << lane) - 1u;
            if (has_a) qn[wbase + __popc(ba & lt)] = push_a;
            if (has_b) qn[wbase + __popc(ba) + __popc(bb & lt)] = push_b;
        }
        if (tid == 0) a.qcount[(level + 2) % 3] = 0;  // the slot of level + 2: nobody reads or writes it now
        grid.sync();
    }
}

}  // namespace

int bvh2_reorder_in_stack_traversal_order_device(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    if (bvh->node_count < 2) return OBVHS_OK;  // bvh2/mod.rs:463-465
    cudaStream_t s = ctx->stream;
    const u32 n = (u32)bvh->node_count;
    DevBuf<u32> parents_tmp, arrivals, pairs, qcount;
    DevBuf<uint2> q0, q1;
    const u32* parents = bvh->parents;
    if (!parents) {
        CU_TRY(ctx, parents_tmp.alloc(n, s));
        ST_TRY(bvh2_compute_parents_into(ctx, bvh, parents_tmp.p));
        parents = parents_tmp.p;
    }
    CU_TRY(ctx, arrivals.alloc(n, s));
    CU_TRY(ctx, pairs.alloc(n, s));
    CU_TRY(ctx, q0.alloc((size_t)n / 2 + 1, s));
    CU_TRY(ctx, q1.alloc((size_t)n / 2 + 1, s));
    CU_TRY(ctx, qcount.alloc(4, s));
    CU_TRY(ctx, cudaMemsetAsync(arrivals.p, 0, (size_t)n * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(pairs.p, 0, (size_t)n * 4, s));
    subtree_pairs_kernel<<<div_up(n, 256), 256, 0, s>>>(bvh->nodes, n, parents, arrivals.p, pairs.p);
    KERNEL_CHECK(ctx);
    Node32* out = nullptr;
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&out, (size_t)n * sizeof(Node32)));
    ReorderArgs ra{bvh->nodes, out, pairs.p, {q0.p, q1.p}, qcount.p};
    static PerDevice<int> per_sm_dev;
    int& per_sm = per_sm_dev[ctx->device];
    if (per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reorder_bfs_kernel, 256, 0);
        if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const int blocks = std::min(per_sm * ctx->sm_count, std::max(1, div_up(n / 2, 256)));
    void* args[] = {&ra};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)reorder_bfs_kernel, dim3(blocks), dim3(256), args, 0, s);
    if (e != cudaSuccess) {
        obvhs_result_free(ctx, out);
        CU_TRY(ctx, e);
    }
    KERNEL_CHECK(ctx);
    obvhs_result_free(bvh->owner, bvh->nodes);  // (stream order: the kernel above is the last reader)
    bvh->nodes = out;
    if (bvh->parents) ST_TRY(bvh2_compute_parents_device(ctx, bvh));  // update_parents (bvh2/mod.rs:492-494)
    bvh->children_are_ordered_after_parents = true;                   // :498
    return OBVHS_OK;
}

int bvh2_compute_parents_device(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    if (bvh->node_count == 0) return OBVHS_OK;
    if (!bvh->parents) CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->parents, bvh->node_count * 4));
    compute_parents_kernel<<<div_up(bvh->node_count, 256), 256, 0, ctx->stream>>>(bvh->nodes, (u32)bvh->node_count, bvh->parents);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int bvh2_compute_parents_into(ObvhsContext* ctx, const ObvhsBvh2* bvh, u32* d_parents) {
    if (bvh->node_count == 0) return OBVHS_OK;
    compute_parents_kernel<<<div_up(bvh->node_count, 256), 256, 0, ctx->stream>>>(bvh->nodes, (u32)bvh->node_count, d_parents);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int bvh2_set_node_aabbs_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const u32* d_node_ids, const ObvhsAabb* d_aabbs, size_t n) {
    if (n == 0) return OBVHS_OK;
    set_node_aabbs_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(bvh->nodes, (u32)bvh->node_count, d_node_ids,
                                                                  reinterpret_cast<const float4*>(d_aabbs), (u32)n);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int bvh2_refit_all_device(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    if (bvh->node_count < 2) return OBVHS_OK;
    if (!bvh->parents) ST_TRY(bvh2_compute_parents_device(ctx, bvh));
    DevBuf<u32> arrivals;
    CU_TRY(ctx, arrivals.alloc(bvh->node_count, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(arrivals.p, 0, bvh->node_count * 4, ctx->stream));
    refit_bottom_up_kernel<<<div_up(bvh->node_count, 256), 256, 0, ctx->stream>>>(bvh->nodes, (u32)bvh->node_count, bvh->parents, arrivals.p);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int bvh2_set_leaf_aabbs_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const ObvhsAabb* d_aabbs) {
    if (bvh->node_count == 0) return OBVHS_OK;
    set_leaf_aabbs_kernel<<<div_up(bvh->node_count, 256), 256, 0, ctx->stream>>>(bvh->nodes, (u32)bvh->node_count, bvh->primitive_indices,
                                                                                reinterpret_cast<const float4*>(d_aabbs));
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int bvh2_expand_nodes_device(ObvhsContext* ctx, const Node32* in, size_t n, ObvhsBvh2Node* d_out) {
    if (n == 0) return OBVHS_OK;
    expand_nodes_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(in, (u32)n, reinterpret_cast<float4*>(d_out));
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}
int bvh2_pack_nodes_device(ObvhsContext* ctx, const ObvhsBvh2Node* d_in, size_t n, Node32* out) {
    if (n == 0) return OBVHS_OK;
    pack_nodes_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(d_in), (u32)n, out);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}
