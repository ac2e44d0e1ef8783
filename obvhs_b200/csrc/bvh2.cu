// bvh2.cu -- Bvh2 container operations on the device: compute_parents, refit_all, layout conversion.
//
//   compute_parents   src/bvh2/mod.rs:586-619   parents[first] = parents[first+1] = i for every inner node i
//   refit_all         src/bvh2/mod.rs:527-569   every inner node = first.union(second), children before parents.
//                     The reference sweeps indices in reverse (or a stack order); any bottom-up order gives the same
//                     bits because min/max are exact and the operand order (first, second) is fixed. Here: one thread
//                     per leaf climbs with an arrival counter per inner node; the second arriver refits the node.
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

}  // namespace

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
