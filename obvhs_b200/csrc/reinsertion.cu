// reinsertion.cu -- parallel reinsertion optimisation of a Bvh2 (Meister & Bittner PRBVH as restated by obvhs).
//
// Replaces ReinsertionOptimizer::run / optimize_impl / find_candidates / optimize_candidates, find_reinsertion and
// reinsert_node (src/bvh2/reinsertion.rs:40-57, 92-208, 233-382) and the refits of Bvh2::refit_from_fast
// (src/bvh2/mod.rs:722-751).
//
//   K8   candidate select   half areas of nodes [1, 2*node_count) -> stable radix sort on the f32 key of -cost
//   K9   find_reinsertion   one thread per candidate, branch-and-bound with a 192-entry (f32,u32) stack
//   K10  apply              the reference applies the gain-sorted list SEQUENTIALLY, skipping entries that touch a node
//                           already touched. The five conflict cells {to, from, sibling(from), parent(to), parent(from)}
//                           of an ACCEPTED entry never change before it is applied (any entry whose parent pointer was
//                           rewritten by an earlier accepted entry also has a touched cell), so the accepted set is the
//                           greedy maximal independent set in rank order over static cells. It is computed with
//                           deterministic reservations: undecided entries atomicMin their rank into their cells; an
//                           entry that holds all five is accepted and marks them; entries seeing a mark are rejected;
//                           repeat. Accepted entries write disjoint nodes/parents and are applied in one kernel; the
//                           dirty ancestor paths are then refit bottom-up (pending-child counters), which produces the
//                           same tight boxes as the reference's per-entry refit_from_fast walks.
#include "common.cuh"

namespace {

// rdst RadixKey for f32: total order preserving map (negative -> flip all, else flip sign)
__device__ __forceinline__ u32 f32_radix_key(float f) {
    u32 u = __float_as_uint(f);
    u32 mask = (u32)((int)u >> 31) | 0x80000000u;
    return u ^ mask;
}

__device__ __forceinline__ float node_half_area(const Node32* __restrict__ nodes, u32 id) {
    const float4* q = reinterpret_cast<const float4*>(nodes + id);
    float4 a = __ldg(q), b = __ldg(q + 1);
    return box_half_area(Box{a.x, a.y, a.z, b.x, b.y, b.z});
}
__device__ __forceinline__ Node32 ldg_node(const Node32* __restrict__ nodes, u32 id) {
    const float4* q = reinterpret_cast<const float4*>(nodes + id);
    float4 a = __ldg(q), b = __ldg(q + 1);
    Node32 n;
    n.minx = a.x; n.miny = a.y; n.minz = a.z; n.prim_count = __float_as_uint(a.w);
    n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.first_index = __float_as_uint(b.w);
    return n;
}

// K8: reinsertion.rs:121-139
__global__ void __launch_bounds__(256) cand_init_kernel(const Node32* __restrict__ nodes, u32 m, u32* __restrict__ keys, u32* __restrict__ vals) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    u32 id = j + 1;
    keys[j] = f32_radix_key(-node_half_area(nodes, id));
    vals[j] = id;
}

constexpr int RSTACK = 192;  // fast_stack!((f32,u32), (96,192), max_depth*2, ...) reinsertion.rs:147 with max_depth <= 96

// K9: reinsertion.rs:233-334. Stack pushes saturate at the last slot and pop_fast saturates at 0 (faststack.rs:299-310).
__global__ void __launch_bounds__(128) find_reinsertion_kernel(const Node32* __restrict__ nodes, const u32* __restrict__ parents,
                                                               const u32* __restrict__ cand_ids, u32 count, u32* __restrict__ r_from,
                                                               u32* __restrict__ r_to, float* __restrict__ r_diff, u32* __restrict__ gain_keys,
                                                               u32* __restrict__ gain_vals) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const u32 node_id = cand_ids[j];
    float s_area[RSTACK];
    u32 s_id[RSTACK];
    u32 sp = 0;
    u32 best_to = 0;
    float best_diff = 0.0f;
    const Node32 self = ldg_node(nodes, node_id);
    const Box aabb = node_box(self);
    const float node_area = box_half_area(aabb);
    const u32 parent_id = parents[node_id];
    const float parent_area = node_half_area(nodes, parent_id);
    float area_diff = parent_area;
    u32 sib = sibling_id(node_id);
    Box pivot_bbox = node_box(ldg_node(nodes, sib));
    u32 pivot_id = parent_id;
    for (;;) {
        s_area[sp] = area_diff;
        s_id[sp] = sib;
        sp = min(sp + 1u, (u32)RSTACK - 1u);
        while (sp != 0) {
            sp = sp - 1;
            float top_area_diff = s_area[sp];
            u32 top_sibling_id = s_id[sp];
            if (top_area_diff - node_area <= best_diff) continue;
            const Node32 dst = ldg_node(nodes, top_sibling_id);
            const Box dbox = node_box(dst);
            float merged_area = box_half_area(box_union(dbox, aabb));
            float reinsert_area = top_area_diff - merged_area;
            if (reinsert_area > best_diff) {
                best_to = top_sibling_id;
                best_diff = reinsert_area;
            }
            if (dst.prim_count == 0) {
                float child_area = reinsert_area + box_half_area(dbox);
                s_area[sp] = child_area;
                s_id[sp] = dst.first_index;
                sp = min(sp + 1u, (u32)RSTACK - 1u);
                s_area[sp] = child_area;
                s_id[sp] = dst.first_index + 1;
                sp = min(sp + 1u, (u32)RSTACK - 1u);
            }
        }
        if (pivot_id != parent_id) {
            pivot_bbox = box_union(pivot_bbox, node_box(ldg_node(nodes, sib)));
            area_diff += node_half_area(nodes, pivot_id) - box_half_area(pivot_bbox);
        }
        if (pivot_id == 0) break;
        sib = sibling_id(pivot_id);
        pivot_id = parents[pivot_id];
    }
    u32 from = node_id;
    if (best_to == sibling_id(from) || best_to == parent_id) {  // reinsertion.rs:328-333 -> Reinsertion::default()
        from = 0;
        best_to = 0;
        best_diff = 0.0f;
    }
    r_from[j] = from;
    r_to[j] = best_to;
    r_diff[j] = best_diff;
    gain_keys[j] = f32_radix_key(-best_diff);  // descending area_diff, ties by candidate rank (stable)
    gain_vals[j] = j;
}

struct ReinsertState {
    u32 active;     // entries with area_diff > 0 (they form a prefix of the gain-sorted list)
    u32 undecided;  // entries still undecided after the current resolution iteration
    u32 applied;    // running total of applied reinsertions
    u32 pad;
};

// cells of entry r (gain order): {to, from, sibling(from), parent(to), parent(from)} (reinsertion.rs:198-208)
__global__ void __launch_bounds__(256) conflicts_prep_kernel(const u32* __restrict__ order, const u32* __restrict__ r_from,
                                                             const u32* __restrict__ r_to, const float* __restrict__ r_diff,
                                                             const u32* __restrict__ parents, u32 count, u32* __restrict__ cells,
                                                             u32* __restrict__ status, ReinsertState* st) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    u32 j = order[r];
    float d = r_diff[j];
    if (!(d > 0.0f)) {  // reinsertion.rs:178-180
        status[r] = 2;
        return;
    }
    u32 from = r_from[j], to = r_to[j];
    u32* c = cells + (size_t)r * 5;
    c[0] = to;
    c[1] = from;
    c[2] = sibling_id(from);
    c[3] = parents[to];
    c[4] = parents[from];
    status[r] = 0;
    atomicMax(&st->active, r + 1);
}

__global__ void __launch_bounds__(256) resolve_reserve_kernel(const u32* __restrict__ cells, u32* __restrict__ status, const ReinsertState* st,
                                                              const u32* __restrict__ touched, u32 round_stamp,
                                                              unsigned long long* __restrict__ reserve, u32 iter_stamp) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st->active || status[r] != 0) return;
    const u32* c = cells + (size_t)r * 5;
    u32 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4];
    if (touched[c0] == round_stamp || touched[c1] == round_stamp || touched[c2] == round_stamp || touched[c3] == round_stamp ||
        touched[c4] == round_stamp) {
        status[r] = 2;
        return;
    }
    // newer iterations carry a smaller high word, so stale reservations of earlier iterations always lose
    unsigned long long v = ((unsigned long long)(~iter_stamp) << 32) | r;
    atomicMin(reserve + c0, v);
    atomicMin(reserve + c1, v);
    atomicMin(reserve + c2, v);
    atomicMin(reserve + c3, v);
    atomicMin(reserve + c4, v);
}

__global__ void __launch_bounds__(256) resolve_commit_kernel(const u32* __restrict__ cells, u32* __restrict__ status, ReinsertState* st,
                                                             u32* __restrict__ touched, u32 round_stamp,
                                                             const unsigned long long* __restrict__ reserve, u32 iter_stamp) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    bool undecided = false;
    if (r < st->active && status[r] == 0) {
        const u32* c = cells + (size_t)r * 5;
        unsigned long long v = ((unsigned long long)(~iter_stamp) << 32) | r;
        u32 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4];
        if (reserve[c0] == v && reserve[c1] == v && reserve[c2] == v && reserve[c3] == v && reserve[c4] == v) {
            status[r] = 1;
            touched[c0] = round_stamp;
            touched[c1] = round_stamp;
            touched[c2] = round_stamp;
            touched[c3] = round_stamp;
            touched[c4] = round_stamp;
        } else {
            undecided = true;
        }
    }
    u32 b = __ballot_sync(0xffffffffu, undecided);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&st->undecided, (u32)__popc(b));
}

// reinsert_node without its two refits (reinsertion.rs:336-382). Accepted entries touch disjoint nodes/parents.
__global__ void __launch_bounds__(256) apply_kernel(const u32* __restrict__ cells, const u32* __restrict__ status, ReinsertState* st,
                                                    Node32* nodes, u32* parents) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = r < st->active && status[r] == 1;
    if (acc) {
        const u32* c = cells + (size_t)r * 5;
        u32 to = c[0], from = c[1], sib = c[2], parent_id = c[4];
        Node32 sibling_node = load_node(nodes + sib);
        Node32 dst_node = load_node(nodes + to);
        Node32 new_to = dst_node;
        new_to.prim_count = 0;  // make_inner(left_sibling(from)); its box is refit below
        new_to.first_index = left_sibling_id(from);
        store_node(nodes + to, new_to);
        store_node(nodes + sib, dst_node);
        store_node(nodes + parent_id, sibling_node);
        if (dst_node.prim_count == 0) {
            parents[dst_node.first_index] = sib;
            parents[dst_node.first_index + 1] = sib;
        }
        if (sibling_node.prim_count == 0) {
            parents[sibling_node.first_index] = parent_id;
            parents[sibling_node.first_index + 1] = parent_id;
        }
        parents[sib] = to;
        parents[from] = to;
    }
    u32 b = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&st->applied, (u32)__popc(b));
}

// Dirty paths: every accepted entry dirties `to` and the old parent of `from`, and all their ancestors. pending[x] counts
// the arrivals node x waits for before it can be refit: one per dirty child, plus one "self" token when x is a start
// node. Whoever brings pending[x] to zero refits x and carries on to its parent, so every dirty node is refit exactly
// once, after all dirty nodes below it.
__global__ void __launch_bounds__(256) refit_mark_kernel(const u32* __restrict__ cells, u32* __restrict__ status, const ReinsertState* st,
                                                         const u32* __restrict__ parents, u32* mark, u32* pending, u32 round_stamp) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st->active || (status[r] & 3u) != 1) return;
    const u32* c = cells + (size_t)r * 5;
    u32 starts[2] = {c[0], c[4]};
    for (int k = 0; k < 2; k++) {
        u32 node = starts[k];
        atomicAdd(&pending[node], 1u);  // self token, released by this entry's thread in refit_dirty_kernel
        if (atomicExch(&mark[node], round_stamp) == round_stamp) continue;
        while (node != 0) {
            u32 p = parents[node];
            atomicAdd(&pending[p], 1u);
            if (atomicExch(&mark[p], round_stamp) == round_stamp) break;
            node = p;
        }
    }
}

__global__ void __launch_bounds__(256) refit_dirty_kernel(const u32* __restrict__ cells, const u32* __restrict__ status, const ReinsertState* st,
                                                          const u32* __restrict__ parents, Node32* nodes, u32* pending) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st->active || (status[r] & 3u) != 1) return;
    const u32* c = cells + (size_t)r * 5;
    u32 starts[2] = {c[0], c[4]};
    for (int k = 0; k < 2; k++) {
        u32 node = starts[k];
        for (;;) {
            if (atomicSub(&pending[node], 1u) != 1u) break;  // somebody below is still due
            Node32 me = load_node_cg(nodes + node);
            if (me.prim_count == 0) {
                Node32 c0 = load_node_cg(nodes + me.first_index), c1 = load_node_cg(nodes + me.first_index + 1);
                store_node(nodes + node, make_node32(box_union(node_box(c0), node_box(c1)), 0u, me.first_index));
            }
            if (node == 0) break;
            __threadfence();
            node = parents[node];
        }
    }
}

__global__ void reinsert_state_reset_kernel(ReinsertState* st, int what) {
    if (what == 0) {
        st->active = 0;
        st->undecided = 0;
    } else {
        st->undecided = 0;
    }
}

}  // namespace

int reinsertion_run_device(ObvhsContext* ctx, ObvhsBvh2* bvh, float ratio, const float* seq, size_t n_seq, u64* applied_out) {
    if (applied_out) *applied_out = 0;
    cudaStream_t s = ctx->stream;
    const size_t len = bvh->node_count;
    if (len == 0 || !(ratio > 0.0f)) return OBVHS_OK;  // reinsertion.rs:43-45 (NaN ratio: `<=` is false in Rust; treated as no-op here)
    if (len == 1) return OBVHS_OK;                      // root is a leaf
    if (bvh->max_depth > 96) {
        OBVHS_SET_ERR(ctx, "reinsertion: max_depth %zu > 96 needs a heap stack (faststack.rs) -- not supported", bvh->max_depth);
        return OBVHS_ERR_UNSUPPORTED;
    }
    if (!bvh->parents) ST_TRY(bvh2_compute_parents_device(ctx, bvh));  // init_parents_if_uninit
    bvh->children_are_ordered_after_parents = false;                    // reinsertion.rs:93
    std::vector<float> default_seq;
    if (!seq) {
        for (int k = 1; k < 32; k += 2) default_seq.push_back(1.0f / (float)k);
        seq = default_seq.data();
        n_seq = default_seq.size();
    }
    // reinsertion.rs:104-107 per round sizes
    std::vector<size_t> node_counts(n_seq);
    size_t max_nc = 0;
    for (size_t k = 0; k < n_seq; k++) {
        volatile float f0 = (float)len * ratio;
        volatile float f = f0 * seq[k];
        double fd = (double)f;
        size_t batch = !(fd > 0.0) ? 0 : (fd >= 18446744073709551616.0 ? ~(size_t)0 : (size_t)fd);  // `as usize` saturates
        if (batch < 1) batch = 1;
        size_t nc = batch >= len ? len : std::min(len, batch + 1);
        node_counts[k] = nc;
        max_nc = std::max(max_nc, nc);
    }
    const size_t max_take = std::min(len, max_nc * 2), max_count = max_nc - 1;
    if (max_count == 0) return OBVHS_OK;
    DevBuf<u32> ckeys, ckeys_alt, cvals, cvals_alt, r_from, r_to, gkeys, gkeys_alt, gvals, gvals_alt, cells, status, touched, mark, pending;
    DevBuf<float> r_diff;
    DevBuf<unsigned long long> reserve;
    DevBuf<ReinsertState> st;
    CU_TRY(ctx, ckeys.alloc(max_take, s));
    CU_TRY(ctx, ckeys_alt.alloc(max_take, s));
    CU_TRY(ctx, cvals.alloc(max_take, s));
    CU_TRY(ctx, cvals_alt.alloc(max_take, s));
    CU_TRY(ctx, r_from.alloc(max_count, s));
    CU_TRY(ctx, r_to.alloc(max_count, s));
    CU_TRY(ctx, r_diff.alloc(max_count, s));
    CU_TRY(ctx, gkeys.alloc(max_count, s));
    CU_TRY(ctx, gkeys_alt.alloc(max_count, s));
    CU_TRY(ctx, gvals.alloc(max_count, s));
    CU_TRY(ctx, gvals_alt.alloc(max_count, s));
    CU_TRY(ctx, cells.alloc(max_count * 5, s));
    CU_TRY(ctx, status.alloc(max_count, s));
    CU_TRY(ctx, touched.alloc(len, s));
    CU_TRY(ctx, mark.alloc(len, s));
    CU_TRY(ctx, pending.alloc(len, s));
    CU_TRY(ctx, reserve.alloc(len, s));
    CU_TRY(ctx, st.alloc(1, s));
    CU_TRY(ctx, cudaMemsetAsync(touched.p, 0, len * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(mark.p, 0, len * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(pending.p, 0, len * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(reserve.p, 0xff, len * 8, s));
    CU_TRY(ctx, cudaMemsetAsync(st.p, 0, sizeof(ReinsertState), s));
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    u32 iter_stamp = 0;
    for (size_t k = 0; k < n_seq; k++) {
        const size_t nc = node_counts[k];
        const u32 take = (u32)std::min(len, nc * 2), count = (u32)(nc - 1);
        if (count == 0 || take < 2) continue;
        const u32 round_stamp = (u32)k + 1;
        const u32 m = take - 1;
        TraceScope* tsp = new TraceScope(ctx, "  reins_candidates_sort");
        cand_init_kernel<<<div_up(m, 256), 256, 0, s>>>(bvh->nodes, m, ckeys.p, cvals.p);
        KERNEL_CHECK(ctx);
        u32 *sk, *cand_ids;
        ST_TRY(radix_sort_pairs_u32(ctx, ckeys.p, ckeys_alt.p, cvals.p, cvals_alt.p, m, 4, &sk, &cand_ids));
        delete tsp;
        tsp = new TraceScope(ctx, "  reins_find");
        find_reinsertion_kernel<<<div_up(count, 128), 128, 0, s>>>(bvh->nodes, bvh->parents, cand_ids, count, r_from.p, r_to.p, r_diff.p,
                                                                  gkeys.p, gvals.p);
        KERNEL_CHECK(ctx);
        delete tsp;
        tsp = new TraceScope(ctx, "  reins_gain_sort_resolve");
        u32 *gk, *order;
        ST_TRY(radix_sort_pairs_u32(ctx, gkeys.p, gkeys_alt.p, gvals.p, gvals_alt.p, count, 4, &gk, &order));
        reinsert_state_reset_kernel<<<1, 1, 0, s>>>(st.p, 0);
        KERNEL_CHECK(ctx);
        conflicts_prep_kernel<<<div_up(count, 256), 256, 0, s>>>(order, r_from.p, r_to.p, r_diff.p, bvh->parents, count, cells.p, status.p, st.p);
        KERNEL_CHECK(ctx);
        CU_TRY(ctx, cudaMemcpyAsync(h, st.p, sizeof(ReinsertState), cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        const u32 active = h[0];
        if (active == 0) {
            delete tsp;
            continue;
        }
        const int blocks = div_up(active, 256);
        for (;;) {
            iter_stamp++;
            resolve_reserve_kernel<<<blocks, 256, 0, s>>>(cells.p, status.p, st.p, touched.p, round_stamp, reserve.p, iter_stamp);
            KERNEL_CHECK(ctx);
            resolve_commit_kernel<<<blocks, 256, 0, s>>>(cells.p, status.p, st.p, touched.p, round_stamp, reserve.p, iter_stamp);
            KERNEL_CHECK(ctx);
            CU_TRY(ctx, cudaMemcpyAsync(h, st.p, sizeof(ReinsertState), cudaMemcpyDeviceToHost, s));
            reinsert_state_reset_kernel<<<1, 1, 0, s>>>(st.p, 1);
            KERNEL_CHECK(ctx);
            CU_TRY(ctx, cudaStreamSynchronize(s));
            if (h[1] == 0) break;
        }
        delete tsp;
        TraceScope ts_apply(ctx, "  reins_apply_refit");
        apply_kernel<<<blocks, 256, 0, s>>>(cells.p, status.p, st.p, bvh->nodes, bvh->parents);
        KERNEL_CHECK(ctx);
        refit_mark_kernel<<<blocks, 256, 0, s>>>(cells.p, status.p, st.p, bvh->parents, mark.p, pending.p, round_stamp);
        KERNEL_CHECK(ctx);
        refit_dirty_kernel<<<blocks, 256, 0, s>>>(cells.p, status.p, st.p, bvh->parents, bvh->nodes, pending.p);
        KERNEL_CHECK(ctx);
    }
    CU_TRY(ctx, cudaMemcpyAsync(h, st.p, sizeof(ReinsertState), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    if (applied_out) *applied_out = h[2];
    return OBVHS_OK;
}
