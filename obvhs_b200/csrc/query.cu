// query.cu -- batched broad-phase queries: one query box / point per thread.
//
// Replaces, for a batch of queries with an `eval` that always continues,
//   Bvh2::aabb_traverse / point_traverse                          src/bvh2/mod.rs:365-456   (reports leaf NODE ids)
//   traverse!(bvh, node, state, node.intersect_aabb(..) | node.contains_point(..), { state.primitive_id })
//                                                                 src/cwbvh/traverse_macro.rs:59-126, src/cwbvh/node.rs:157-200
//                                                                 (reports primitive slots)
// as used by the physics example's collision broad phase (examples/physics.rs:566-588) and the reference's traverse_aabb /
// traverse_point tests (tests/mod.rs:178-325).
//
// The reference hands every report to a closure; a batch has no closure, so the result is the list of reports itself:
// pass 1 counts per query, an exclusive prefix sum places the queries, pass 2 repeats the walk and writes each query's ids
// in exactly the reference's call order (same stack discipline, left before right / highest bit first).
#include "common.cuh"
#include "compact.cuh"

namespace {

struct QueryBox {  // QUERY 0: lo/hi = the Aabb; QUERY 1: lo = the point
    float lox, loy, loz, hix, hiy, hiz;
};
template <int QUERY>
__device__ __forceinline__ QueryBox load_query(const float4* __restrict__ q, size_t i) {
    if (QUERY == 0) {
        const float4 a = __ldg(q + 2 * i), b = __ldg(q + 2 * i + 1);
        return QueryBox{a.x, a.y, a.z, b.x, b.y, b.z};
    }
    const float4 a = __ldg(q + i);
    return QueryBox{a.x, a.y, a.z, 0.f, 0.f, 0.f};
}
// aabb.rs:181-183 (self = node box, other = query) and aabb.rs:70-72; NaN compares false in both
template <int QUERY>
__device__ __forceinline__ bool box_test(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, const QueryBox& q) {
    if (QUERY == 0) return !(mnx > q.hix || mny > q.hiy || mnz > q.hiz || mxx < q.lox || mxy < q.loy || mxz < q.loz);
    return q.lox >= mnx && q.loy >= mny && q.loz >= mnz && q.lox <= mxx && q.loy <= mxy && q.loz <= mxz;
}

struct Emitter {  // FILL = false: count only
    u32* out;
    u32 count;
    template <bool FILL>
    __device__ __forceinline__ void emit(u32 id) {
        if (FILL) out[count] = id;
        count++;
    }
};

// The u32 stack of the Bvh2 queries: the reference's fixed StackStack<u32, 96 | 192> (saturating push), or -- beyond max_depth
// 192, where fast_stack! switches to HeapStack::new_with_capacity(max_depth) (faststack.rs:44-47) -- a slice of a global arena,
// interleaved over the threads of the grid.
template <int CAP>
struct QueryStack {
    u32 data[CAP];
    __device__ __forceinline__ u32 cap() const { return CAP; }
    __device__ __forceinline__ u32& at(u32 i) { return data[i]; }
};
struct QueryHeapStack {
    u32* base;
    u32 stride, capacity;
    __device__ __forceinline__ u32 cap() const { return capacity; }
    __device__ __forceinline__ u32& at(u32 i) { return base[(size_t)i * stride]; }
};

// bvh2/mod.rs:365-456
template <int QUERY, bool FILL, class Stack>
__device__ __forceinline__ void bvh2_query(const float4* __restrict__ nodes, u32 node_count, const QueryBox& q, Emitter& e, Stack& stack) {
    if (node_count == 0) return;
    {
        const float4 lo = __ldg(nodes), hi = __ldg(nodes + 1);
        if (__float_as_uint(lo.w) != 0) {  // the root is a leaf (:370-376)
            if (box_test<QUERY>(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, q)) e.emit<FILL>(0u);
            return;
        }
        const u32 cap1 = stack.cap() - 1u;
        u32 sp = 0;
        stack.at(sp++) = __float_as_uint(hi.w);
        while (sp > 0) {
            const u32 node_index = stack.at(--sp);
            const float4* np = nodes + (size_t)node_index * 2;
            const float4 llo = __ldg(np), lhi = __ldg(np + 1), rlo = __ldg(np + 2), rhi = __ldg(np + 3);
            if (box_test<QUERY>(llo.x, llo.y, llo.z, lhi.x, lhi.y, lhi.z, q)) {
                if (__float_as_uint(llo.w) != 0) e.emit<FILL>(node_index);
                else {
                    stack.at(sp) = __float_as_uint(lhi.w);
                    sp = min(sp + 1u, cap1);
                }
            }
            if (box_test<QUERY>(rlo.x, rlo.y, rlo.z, rhi.x, rhi.y, rhi.z, q)) {
                if (__float_as_uint(rlo.w) != 0) e.emit<FILL>(node_index + 1);
                else {
                    stack.at(sp) = __float_as_uint(rhi.w);
                    sp = min(sp + 1u, cap1);
                }
            }
        }
    }
}

// cwbvh/node.rs:157-200: the query is moved into the node's quantisation grid ((q - p) * (1 / extent)) and compared with the
// 8-bit child boxes; hit children contribute child_bits << bit_index exactly as in intersect_ray (node.rs:207-231).
template <int QUERY>
__device__ __forceinline__ u32 cw_node_query(const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3, const uint4 q4, const QueryBox& q,
                                             u32 oct_inv4) {
    const float px = __uint_as_float(q0.x), py = __uint_as_float(q0.y), pz = __uint_as_float(q0.z);
    const float rx = __fdiv_rn(1.0f, __uint_as_float((q0.w & 0xffu) << 23)), ry = __fdiv_rn(1.0f, __uint_as_float(((q0.w >> 8) & 0xffu) << 23)),
                rz = __fdiv_rn(1.0f, __uint_as_float(((q0.w >> 16) & 0xffu) << 23));
    QueryBox a;
    a.lox = __fmul_rn(__fsub_rn(q.lox, px), rx); a.loy = __fmul_rn(__fsub_rn(q.loy, py), ry); a.loz = __fmul_rn(__fsub_rn(q.loz, pz), rz);
    if (QUERY == 0) {
        a.hix = __fmul_rn(__fsub_rn(q.hix, px), rx); a.hiy = __fmul_rn(__fsub_rn(q.hiy, py), ry); a.hiz = __fmul_rn(__fsub_rn(q.hiz, pz), rz);
    } else {
        a.hix = a.hiy = a.hiz = 0.f;
    }
    // q2 = {min_x[0..3], min_x[4..7], max_x[0..3], max_x[4..7]}, q3 = y, q4 = z; q1.z / q1.w = child_meta[0..3] / [4..7]
    const u32 mnx[2] = {q2.x, q2.y}, mxx[2] = {q2.z, q2.w}, mny[2] = {q3.x, q3.y}, mxy[2] = {q3.z, q3.w}, mnz[2] = {q4.x, q4.y}, mxz[2] = {q4.z, q4.w};
    const u32 meta[2] = {q1.z, q1.w};
    u32 hit_mask = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const u32 m = meta[h];
        const u32 is_inner = (m & (m << 1)) & 0x10101010u;
        const u32 inner_mask = (is_inner >> 4) * 0xffu;
        const u32 bit_index = (m ^ (oct_inv4 & inner_mask)) & 0x1f1f1f1fu;
        const u32 child_bits = (m >> 5) & 0x07070707u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int sh = 8 * j;
            const bool hit = box_test<QUERY>((float)((mnx[h] >> sh) & 0xffu), (float)((mny[h] >> sh) & 0xffu), (float)((mnz[h] >> sh) & 0xffu),
                                             (float)((mxx[h] >> sh) & 0xffu), (float)((mxy[h] >> sh) & 0xffu), (float)((mxz[h] >> sh) & 0xffu), a);
            if (hit) hit_mask |= ((child_bits >> sh) & 0xffu) << ((bit_index >> sh) & 0xffu);
        }
    }
    return hit_mask;
}

template <int QUERY, bool FILL>
__device__ __forceinline__ void cwbvh_query(const uint4* __restrict__ nodes, u32 root_group, u32 oct_inv4, const QueryBox& q, Emitter& e) {
    uint2 stack[32];  // StackStack<UVec2, 32>, cwbvh/mod.rs:60
    u32 sp = 0;
    uint2 cur = make_uint2(0u, root_group), prim = make_uint2(0u, 0u);
    for (;;) {
        while (prim.y != 0) {  // traverse_macro.rs:64-72
            const u32 local = 31u - __clz(prim.y);
            prim.y &= ~(1u << local);
            e.emit<FILL>(prim.x + local);
        }
        prim = make_uint2(0u, 0u);
        if (cur.y & 0xff000000u) {
            const u32 hits_imask = cur.y;
            const u32 child_index_offset = 31u - __clz(hits_imask);
            const u32 child_index_base = cur.x;
            cur.y &= ~(1u << child_index_offset);
            if (cur.y & 0xff000000u) {
                stack[sp] = cur;
                sp = min(sp + 1u, 31u);
            }
            const u32 slot_index = (child_index_offset - 24u) ^ (oct_inv4 & 0xffu);
            const u32 relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index));
            const uint4* np = nodes + (size_t)(child_index_base + relative_index) * 5;
            const uint4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3), q4 = __ldg(np + 4);
            const u32 hitmask = cw_node_query<QUERY>(q0, q1, q2, q3, q4, q, oct_inv4);
            cur.x = q1.x;
            prim.x = q1.y;
            cur.y = (hitmask & 0xff000000u) | (q0.w >> 24);
            prim.y = hitmask & 0x00ffffffu;
        } else {
            cur = make_uint2(0u, 0u);
        }
        if (prim.y == 0 && (cur.y & 0xff000000u) == 0) {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
}

// TREE 0 = Bvh2 (CAP 96), 1 = Bvh2 (CAP 192), 2 = CwBvh, 3 = Bvh2 with a heap stack of heap_cap entries per thread (grid-stride:
// the launch has a bounded number of threads so that the arena stays small)
template <int TREE, int QUERY, bool FILL>
__global__ void __launch_bounds__(128) query_kernel(const void* __restrict__ nodes, u32 node_count, u32 root_group, u32 oct_inv4,
                                                    const float4* __restrict__ queries, u32 n, u32* __restrict__ counts,
                                                    const u32* __restrict__ offsets, u32* __restrict__ ids, u32* heap, u32 heap_cap) {
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    for (u32 i = tid; i < n; i += nthreads) {
        const QueryBox q = load_query<QUERY>(queries, i);
        Emitter e{FILL ? ids + offsets[i] : nullptr, 0u};
        if (TREE == 0) {
            QueryStack<96> st;
            bvh2_query<QUERY, FILL>(reinterpret_cast<const float4*>(nodes), node_count, q, e, st);
        } else if (TREE == 1) {
            QueryStack<192> st;
            bvh2_query<QUERY, FILL>(reinterpret_cast<const float4*>(nodes), node_count, q, e, st);
        } else if (TREE == 3) {
            QueryHeapStack st{heap + tid, nthreads, heap_cap};
            bvh2_query<QUERY, FILL>(reinterpret_cast<const float4*>(nodes), node_count, q, e, st);
        } else {
            cwbvh_query<QUERY, FILL>(reinterpret_cast<const uint4*>(nodes), root_group, oct_inv4, q, e);
        }
        if (!FILL) counts[i] = e.count;
    }
}

struct CountOf {
    const u32* counts;
    __device__ u32 operator()(u32 i) const { return counts[i]; }
};
struct StoreOffset {
    u32* offsets;
    __device__ void operator()(u32 i, u32 exclusive, u32) const { offsets[i] = exclusive; }
};

// 64-bit total of the per-query counts: the offsets are a u32 prefix scan, which silently wraps from 2^32 reports on
__global__ void __launch_bounds__(256) sum_counts_u64_kernel(const u32* __restrict__ counts, u32 n, unsigned long long* total) {
    unsigned long long v = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += counts[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, v);
}

template <int TREE, int QUERY>
int run_query(ObvhsContext* ctx, const void* nodes, u32 node_count, u32 root_group, u32 oct_inv4, const float4* d_queries, size_t n, u32* counts,
              u32* ids, size_t capacity, size_t* total_out, u32 heap_cap = 0) {
    cudaStream_t s = ctx->stream;
    if (total_out) *total_out = 0;
    if (n == 0) return OBVHS_OK;
    if (n >= (1u << 31)) {
        OBVHS_SET_ERR(ctx, "query batch too large: %zu", n);
        return OBVHS_ERR_UNSUPPORTED;
    }
    const u32 un = (u32)n;
    DevBuf<u32> d_counts, d_offsets, tiles, d_total, d_ids;
    DevBuf<unsigned long long> d_total64;
    CU_TRY(ctx, d_total64.alloc(1, s));
    CU_TRY(ctx, cudaMemsetAsync(d_total64.p, 0, 8, s));
    CU_TRY(ctx, d_counts.alloc(n, s));
    CU_TRY(ctx, d_offsets.alloc(n, s));
    CU_TRY(ctx, tiles.alloc((size_t)div_up(n, CP_TILE) + 1, s));
    CU_TRY(ctx, d_total.alloc(1, s));
    // heap stacks (TREE 3): at most 256 Ki threads, each with heap_cap entries of the arena
    DevBuf<u32> heap;
    int blocks = div_up(n, 128);
    if (TREE == 3) {
        blocks = std::min(blocks, 2048);
        CU_TRY(ctx, heap.alloc((size_t)blocks * 128 * heap_cap, s));
    }
    query_kernel<TREE, QUERY, false><<<blocks, 128, 0, s>>>(nodes, node_count, root_group, oct_inv4, d_queries, un, d_counts.p, nullptr, nullptr,
                                                             heap.p, heap_cap);
    KERNEL_CHECK(ctx);
    ST_TRY(scan_values(ctx, CountOf{d_counts.p}, StoreOffset{d_offsets.p}, un, tiles.p, d_total.p));
    sum_counts_u64_kernel<<<std::min(div_up(n, 256), ctx->sm_count * 8), 256, 0, s>>>(d_counts.p, un, d_total64.p);
    KERNEL_CHECK(ctx);
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, d_total.p, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(h + 2, d_total64.p, 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    unsigned long long total64 = 0;
    memcpy(&total64, h + 2, 8);
    if (total64 >= (1ull << 32)) {  // the u32 offsets would have wrapped: refuse instead of writing ids at wrapped positions
        if (total_out) *total_out = (size_t)total64;
        OBVHS_SET_ERR(ctx, "query batch reports %llu ids: 2^32 or more per batch is not supported (split the batch)", total64);
        return OBVHS_ERR_UNSUPPORTED;
    }
    const size_t total = h[0];
    if (total_out) *total_out = total;
    ST_TRY(copy_out(ctx, counts, (const u32*)d_counts.p, n));
    if (!ids) return OBVHS_OK;
    if (total > capacity) {
        OBVHS_SET_ERR(ctx, "query batch reports %zu ids, capacity %zu (call again with more room)", total, capacity);
        return OBVHS_ERR_CAPACITY;
    }
    if (total == 0) return OBVHS_OK;
    u32* d_out = ids;
    const bool out_dev = obvhs_is_device_ptr(ids);
    if (!out_dev) {
        CU_TRY(ctx, d_ids.alloc(total, s));
        d_out = d_ids.p;
    }
    query_kernel<TREE, QUERY, true><<<blocks, 128, 0, s>>>(nodes, node_count, root_group, oct_inv4, d_queries, un, nullptr, d_offsets.p, d_out, heap.p,
                                                            heap_cap);
    KERNEL_CHECK(ctx);
    if (!out_dev) ST_TRY(copy_out(ctx, ids, (const u32*)d_out, total));
    return OBVHS_OK;
}

u32 octant_inv4(const float* dir) {  // cwbvh/mod.rs:1001-1010; NULL = Vec3A::ZERO (all axes >= 0)
    if (!dir) return 0x07070707u;
    return (dir[0] < 0.0f ? 0u : 0x04040404u) | (dir[1] < 0.0f ? 0u : 0x02020202u) | (dir[2] < 0.0f ? 0u : 0x01010101u);
}

}  // namespace

int bvh2_query_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, int query_kind, const float4* d_queries, size_t n, u32* counts, u32* ids,
                      size_t capacity, size_t* total_out) {
    const u32 nc0 = (u32)bvh->node_count;
    if (bvh->max_depth > 192) {  // the reference switches to HeapStack::new_with_capacity(max_depth) there (fast_stack!)
        const u32 cap = (u32)bvh->max_depth;
        if (query_kind == 0) return run_query<3, 0>(ctx, bvh->nodes, nc0, 0, 0, d_queries, n, counts, ids, capacity, total_out, cap);
        return run_query<3, 1>(ctx, bvh->nodes, nc0, 0, 0, d_queries, n, counts, ids, capacity, total_out, cap);
    }
    const bool deep = bvh->max_depth > 96;
    const u32 nc = (u32)bvh->node_count;
    if (query_kind == 0)
        return deep ? run_query<1, 0>(ctx, bvh->nodes, nc, 0, 0, d_queries, n, counts, ids, capacity, total_out)
                    : run_query<0, 0>(ctx, bvh->nodes, nc, 0, 0, d_queries, n, counts, ids, capacity, total_out);
    return deep ? run_query<1, 1>(ctx, bvh->nodes, nc, 0, 0, d_queries, n, counts, ids, capacity, total_out)
                : run_query<0, 1>(ctx, bvh->nodes, nc, 0, 0, d_queries, n, counts, ids, capacity, total_out);
}

namespace {
// CwBvh::compute_parents (cwbvh/mod.rs:494-509): one thread per (node, child slot); an occupied inner slot writes the node's index
// into its child's entry. Every node but the root is referenced by exactly one slot, so the writes never collide; parents[0] = 0.
__global__ void __launch_bounds__(256) cwbvh_parents_kernel(const uint4* __restrict__ nodes, u32 node_count, u32* __restrict__ parents) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 node = t >> 3, ch = t & 7u;
    if (node >= node_count) return;
    if (t == 0) parents[0] = 0;
    const uint4 q0 = __ldg(nodes + (size_t)node * 5), q1 = __ldg(nodes + (size_t)node * 5 + 1);
    const u32 imask = q0.w >> 24;
    const u32 meta = ((ch < 4 ? q1.z : q1.w) >> ((ch & 3u) * 8u)) & 0xffu;  // child_meta[ch]
    if (meta == 0 || !(imask & (1u << ch))) return;                         // is_child_empty / is_leaf (node.rs:279-286)
    const u32 slot = (meta & 31u) - 24u;                                    // child_node_index (node.rs:299-304)
    parents[q1.x + __popc(imask & ~(0xffffffffu << slot))] = node;
}
}  // namespace

int cwbvh_compute_parents_device(ObvhsContext* ctx, const ObvhsCwBvh* bvh, u32* d_parents) {
    if (bvh->node_count == 0) return OBVHS_OK;
    CU_TRY(ctx, cudaMemsetAsync(d_parents, 0, bvh->node_count * sizeof(u32), ctx->stream));  // vec![0; nodes.len()]
    cwbvh_parents_kernel<<<div_up(bvh->node_count * 8, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(bvh->nodes), (u32)bvh->node_count,
                                                                                    d_parents);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

int cwbvh_query_device(ObvhsContext* ctx, const ObvhsCwBvh* bvh, int query_kind, const float4* d_queries, size_t n, const float* host_dir3,
                       u32* counts, u32* ids, size_t capacity, size_t* total_out) {
    const u32 root_group = bvh->node_count ? 0x80000000u : 0u;  // cwbvh/mod.rs:149-153
    const u32 oct = octant_inv4(host_dir3);
    if (query_kind == 0) return run_query<2, 0>(ctx, bvh->nodes, (u32)bvh->node_count, root_group, oct, d_queries, n, counts, ids, capacity, total_out);
    return run_query<2, 1>(ctx, bvh->nodes, (u32)bvh->node_count, root_group, oct, d_queries, n, counts, ids, capacity, total_out);
}
