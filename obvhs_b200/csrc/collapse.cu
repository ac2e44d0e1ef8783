// collapse.cu -- SAH leaf collapse of a Bvh2 on the device.
//
// Replaces src/bvh2/leaf_collapser.rs:21-192 `collapse(bvh, max_prims, traversal_cost)`:
//   1. bottom-up pass (leaf_collapser.rs:48-87, bottom_up_traverse :238-282): one thread per leaf climbs with an arrival
//      flag per inner node; the second arriver owns the node, reads the primitive counts its children ended up with and
//      decides whether the two child leaves collapse into it (`collapse_cost <= base_cost`, or both hold the same
//      primitive). The decision only depends on finished children, so the result is independent of thread timing.
//   2. inclusive prefix sums of the surviving-node flags and of the per-node primitive counts (:89-98), here ONE scan
//      over packed (node, prim) pairs;
//   3. emit (:100-175): surviving nodes move to their scanned slot; a node that owns primitives becomes a leaf whose
//      primitives are gathered by the reference's stackless top-down walk (left to right, parent pointers on the way up),
//      so `primitive_indices` comes out in exactly the reference's order.
// Parents are recomputed afterwards only if the tree had them before (:183-190).
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ Node32 ldg_node32(const Node32* __restrict__ nodes, u32 id) {
    const float4* q = reinterpret_cast<const float4*>(nodes + id);
    float4 a = __ldg(q), b = __ldg(q + 1);
    Node32 n;
    n.minx = a.x; n.miny = a.y; n.minz = a.z; n.prim_count = __float_as_uint(a.w);
    n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.first_index = __float_as_uint(b.w);
    return n;
}

// counts[i] = (node_count, prim_count) of leaf_collapser.rs:41-42
__global__ void __launch_bounds__(256) collapse_bottom_up_kernel(const Node32* __restrict__ nodes, u32 n, const u32* __restrict__ parents,
                                                                 uint2* counts, u32* flags, u32 max_prims, float traversal_cost) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    const Node32 self = ldg_node32(nodes, i);
    if (self.prim_count == 0) return;                // paths start at leaves (:250-252)
    counts[i] = make_uint2(1u, self.prim_count);     // :51-52
    u32 j = i;
    while (j != 0) {
        j = parents[j];
        __threadfence();
        if (atomicAdd(&flags[j], 1u) != 1u) break;  // the first arriver stops; the second processes the node (:262-266)
        const Node32 node = ldg_node32(nodes, j);
        const u32 first_child = node.first_index;
        const uint2 lc = __ldcg(&counts[first_child]), rc = __ldcg(&counts[first_child + 1]);
        const u32 left_count = lc.y, right_count = rc.y, total_count = left_count + right_count;
        uint2 mine = make_uint2(1u, 0u);
        if (left_count > 0 && right_count > 0 && total_count <= max_prims) {  // :62
            const Node32 left = ldg_node32(nodes, first_child), right = ldg_node32(nodes, first_child + 1);
            const float collapse_cost = __fmul_rn(box_half_area(node_box(node)), __fsub_rn((float)total_count, traversal_cost));
            const float base_cost = __fadd_rn(__fmul_rn(box_half_area(node_box(left)), (float)left_count),
                                              __fmul_rn(box_half_area(node_box(right)), (float)right_count));
            const bool both_have_same_prim = (left.first_index == right.first_index) && total_count == 2;
            if (collapse_cost <= base_cost || both_have_same_prim) {  // :74-82
                mine.y = total_count;
                __stcg(&counts[first_child], make_uint2(0u, 0u));
                __stcg(&counts[first_child + 1], make_uint2(0u, 0u));
            }
        }
        __stcg(&counts[j], mine);
    }
}

// ---- inclusive scan of uint2 (component-wise), three small kernels ---------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint2 add2(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ uint2 block_exclusive_scan2(uint2 v, uint2* total) {  // exclusive scan of one value per thread
    __shared__ uint2 warp_sums[SCAN_THREADS / 32];
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint2 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 x = __shfl_up_sync(0xffffffffu, incl.x, o), y = __shfl_up_sync(0xffffffffu, incl.y, o);
        if (lane >= (u32)o) incl = add2(incl, make_uint2(x, y));
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    uint2 base = make_uint2(0u, 0u), tot = make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; k++) {
        if ((u32)k < w) base = add2(base, warp_sums[k]);
        tot = add2(tot, warp_sums[k]);
    }
    __syncthreads();
    *total = tot;
    return make_uint2(base.x + incl.x - v.x, base.y + incl.y - v.y);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint2* __restrict__ in, u32 n, uint2* __restrict__ tile_sums) {
    const u32 base = blockIdx.x * SCAN_TILE;
    uint2 s = make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        u32 i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) s = add2(s, in[i]);
    }
    uint2 tot;
    block_exclusive_scan2(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(uint2* tile_sums, u32 tiles) {  // one block: exclusive scan in place
    uint2 carry = make_uint2(0u, 0u);
    for (u32 base = 0; base < tiles; base += SCAN_THREADS) {
        u32 i = base + threadIdx.x;
        uint2 v = i < tiles ? tile_sums[i] : make_uint2(0u, 0u), tot;
        uint2 ex = block_exclusive_scan2(v, &tot);
        if (i < tiles) tile_sums[i] = add2(carry, ex);
        carry = add2(carry, tot);
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint2* __restrict__ in, u32 n, const uint2* __restrict__ tile_offsets,
                                                                  uint2* __restrict__ out) {
    const u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;  // blocked arrangement: thread t owns 8 consecutive items
    uint2 v[SCAN_ITEMS];
    uint2 s = make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : make_uint2(0u, 0u);
        s = add2(s, v[k]);
    }
    uint2 tot;
    uint2 run = add2(block_exclusive_scan2(s, &tot), tile_offsets[blockIdx.x]);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        run = add2(run, v[k]);
        if (base + k < n) out[base + k] = run;
    }
}

// leaf_collapser.rs:100-175; incl[i] = inclusive sums (node_counts, prim_counts)
__global__ void __launch_bounds__(256) collapse_emit_kernel(const Node32* __restrict__ nodes, u32 n, const u32* __restrict__ parents,
                                                            const u32* __restrict__ prim_idx, const uint2* __restrict__ incl,
                                                            Node32* __restrict__ out_nodes, u32* __restrict__ out_idx) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Node32 nd = ldg_node32(nodes, i);
    if (i == 0) {  // :111-115 (the root never becomes a leaf here: see bvh2_collapse_device)
        nd.first_index = incl[nd.first_index - 1].x;
        store_node(out_nodes, nd);
        return;
    }
    const uint2 prev = incl[i - 1], cur = incl[i];
    if (prev.x == cur.x) return;  // removed: it sits inside a collapsed subtree (:155-157)
    const u32 node_id = prev.x;
    if (prev.y == cur.y) {  // an inner node that stays inner (:160-162)
        nd.first_index = incl[nd.first_index - 1].x;
    } else {                // a leaf, original or collapsed (:163-170)
        u32 first_prim = prev.y;
        nd.prim_count = cur.y - prev.y;
        nd.first_index = first_prim;
        u32 j = i;  // top_down_traverse :122-151
        for (;;) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(nodes + j)), b = __ldg(reinterpret_cast<const float4*>(nodes + j) + 1);
            const u32 pc = __float_as_uint(a.w), fi = __float_as_uint(b.w);
            if (pc != 0) {
                for (u32 k = 0; k < pc; k++) out_idx[first_prim + k] = prim_idx[fi + k];
                first_prim += pc;
                while (!(j & 1u) && j != i) j = parents[j];  // climb while j is a right sibling
                if (j == i) break;
                j = sibling_id(j);
            } else {
                j = fi;
            }
        }
    }
    store_node(out_nodes + node_id, nd);
}

}  // namespace

int bvh2_collapse_device(ObvhsContext* ctx, ObvhsBvh2* bvh, u32 max_prims, float traversal_cost) {
    cudaStream_t s = ctx->stream;
    const size_t n = bvh->node_count;
    if (max_prims <= 1 || (u32)n <= max_prims * 2 + 1) return OBVHS_OK;                    // leaf_collapser.rs:25-27
    if (bvh->prim_count != 0 && (u32)bvh->prim_count <= max_prims) return OBVHS_OK;       // :29-31
    if (n == 0) return OBVHS_OK;                                                           // :33 (a leaf root has n == 1, caught above)
    TraceScope ts(ctx, "collapse");
    const bool previously_had_parents = bvh->parents != nullptr;
    DevBuf<u32> tmp_parents, flags;
    const u32* parents = bvh->parents;
    if (!parents) {
        CU_TRY(ctx, tmp_parents.alloc(n, s));
        ST_TRY(bvh2_compute_parents_into(ctx, bvh, tmp_parents.p));
        parents = tmp_parents.p;
    }
    DevBuf<uint2> counts, incl, tile_sums;
    const u32 tiles = (u32)div_up(n, SCAN_TILE);
    CU_TRY(ctx, counts.alloc(n, s));
    CU_TRY(ctx, incl.alloc(n, s));
    CU_TRY(ctx, tile_sums.alloc(tiles, s));
    CU_TRY(ctx, flags.alloc(n, s));
    CU_TRY(ctx, cudaMemsetAsync(flags.p, 0, n * 4, s));
    collapse_bottom_up_kernel<<<div_up(n, 256), 256, 0, s>>>(bvh->nodes, (u32)n, parents, counts.p, flags.p, max_prims, traversal_cost);
    KERNEL_CHECK(ctx);
    scan_tile_sums_kernel<<<tiles, SCAN_THREADS, 0, s>>>(counts.p, (u32)n, tile_sums.p);
    KERNEL_CHECK(ctx);
    scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, s>>>(tile_sums.p, tiles);
    KERNEL_CHECK(ctx);
    scan_apply_kernel<<<tiles, SCAN_THREADS, 0, s>>>(counts.p, (u32)n, tile_sums.p, incl.p);
    KERNEL_CHECK(ctx);
    // sizes of the collapsed tree: the only host round trip
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, incl.p + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(h + 2, incl.p, 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    const u32 new_nodes = h[0], new_prims = h[1], root_prims = h[3];
    if (root_prims != 0) {  // leaf_collapser.rs:104-110; unreachable when every leaf holds at least one primitive
        OBVHS_SET_ERR(ctx, "collapse: the root became a leaf (%u primitives) -- not supported", root_prims);
        return OBVHS_ERR_UNSUPPORTED;
    }
    // (even when nothing collapsed the reference rewrites primitive_indices in node order, so the emit pass always runs)
    Node32* out_nodes = nullptr;
    u32* out_idx = nullptr;
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&out_nodes, (size_t)new_nodes * sizeof(Node32)));
    if (cudaError_t e = obvhs_result_alloc(ctx, (void**)&out_idx, std::max<size_t>(1, new_prims) * 4); e != cudaSuccess) {
        obvhs_result_free(ctx, out_nodes);
        CU_TRY(ctx, e);
    }
    collapse_emit_kernel<<<div_up(n, 256), 256, 0, s>>>(bvh->nodes, (u32)n, parents, bvh->primitive_indices, incl.p, out_nodes, out_idx);
    ctx->launches++;
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
        obvhs_result_free(ctx, out_nodes);
        obvhs_result_free(ctx, out_idx);
        CU_TRY(ctx, e);
    }
    obvhs_result_free(ctx, bvh->nodes);
    obvhs_result_free(ctx, bvh->primitive_indices);
    bvh->nodes = out_nodes;
    bvh->primitive_indices = out_idx;
    bvh->node_count = new_nodes;
    bvh->prim_count = new_prims;
    if (previously_had_parents) {
        obvhs_result_free(ctx, bvh->parents);  // sized for the old node count; the cache hands back a fitting block
        bvh->parents = nullptr;
        ST_TRY(bvh2_compute_parents_device(ctx, bvh));  // update_parents, :183-186
    }
    return OBVHS_OK;
}
