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
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

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
    u32 active;        // entries with area_diff > 0 (they form a prefix of the gain-sorted list)
    u32 undecided[3];  // entries still undecided after resolution iteration i, in slot i % 3
    u32 applied;       // running total of applied reinsertions
    u32 error;         // 1: resolution iteration counter overflow
    u32 pad[2];
};

struct ResolveArgs {
    const u32* order;      // gain-sorted rank -> candidate rank
    const u32* r_from;
    const u32* r_to;
    const float* r_diff;
    u32 count;
    u32* cells;            // 5 per entry
    u32* status;           // 0 undecided, 1 accepted, 2 rejected / inactive
    ReinsertState* st;
    u32* touched;          // per node, == round_stamp when touched this round
    unsigned long long* reserve;  // per node, min over ((~stamp) << 32 | rank)
    u32* mark;             // per node, == round_stamp when on a dirty path
    u32* pending;          // per node, arrivals still due before the node can be refit
    Node32* nodes;
    u32* parents;
    u32 round_stamp;       // 1-based round index
};

// K10, one cooperative launch per round: conflict cells -> greedy-MIS resolution by deterministic reservations ->
// apply -> dirty-path marking -> bottom-up refit. Phases are separated by grid-wide barriers; no host round trips.
constexpr int RESOLVE_THREADS = 1024;  // few, fat blocks: a grid-wide barrier costs more the more blocks take part
// SINGLE: the whole round fits one CTA (a few thousand entries): the phases are separated by __syncthreads instead of
// grid-wide barriers (about 2 us each; a round has ~20 of them, which was the entire 40 us of a kitchen-sized round).
struct PhaseBarrier {
    cg::grid_group grid;
    bool single;
    __device__ __forceinline__ void sync() {
        if (single) __syncthreads();
        else grid.sync();
    }
};
template <bool SINGLE>
__global__ void __launch_bounds__(RESOLVE_THREADS) reinsert_resolve_apply_kernel(ResolveArgs a) {
    PhaseBarrier grid{cg::this_grid(), SINGLE};
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    ReinsertState* st = a.st;
    if (tid == 0) st->undecided[0] = st->undecided[1] = st->undecided[2] = 0;
    // ---- cells of entry r (gain order): {to, from, sibling(from), parent(to), parent(from)} (reinsertion.rs:198-208)
    for (u32 r = tid; r < a.count; r += nthreads) {
        u32 j = a.order[r];
        float d = a.r_diff[j];
        if (!(d > 0.0f)) {  // reinsertion.rs:178-180
            a.status[r] = 2;
            continue;
        }
        u32 from = a.r_from[j], to = a.r_to[j];
        u32* c = a.cells + (size_t)r * 5;
        c[0] = to;
        c[1] = from;
        c[2] = sibling_id(from);
        c[3] = a.parents[to];
        c[4] = a.parents[from];
        a.status[r] = 0;
        atomicMax(&st->active, r + 1);
    }
    grid.sync();
    const u32 active = __ldcg(&st->active);
    // ---- resolution: the accepted set of the reference's sequential sweep (see the file header)
    for (u32 iter = 1;; iter++) {
        if (iter > 0xffffu) {
            if (tid == 0) st->error = 1;
            break;
        }
        const u32 stamp = (a.round_stamp << 16) | iter;  // strictly increasing over the whole run
        if (tid == 0) st->undecided[(iter + 1) % 3] = 0;  // slot of the NEXT iteration; the previous one may still be read
        for (u32 r = tid; r < active; r += nthreads) {
            if (__ldcg(&a.status[r]) != 0) continue;
            const u32* c = a.cells + (size_t)r * 5;
            u32 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4];
            if (__ldcg(&a.touched[c0]) == a.round_stamp || __ldcg(&a.touched[c1]) == a.round_stamp || __ldcg(&a.touched[c2]) == a.round_stamp ||
                __ldcg(&a.touched[c3]) == a.round_stamp || __ldcg(&a.touched[c4]) == a.round_stamp) {
                a.status[r] = 2;
                continue;
            }
            // newer stamps carry a smaller high word, so stale reservations always lose against current ones
            unsigned long long v = ((unsigned long long)(~stamp) << 32) | r;
            atomicMin(a.reserve + c0, v);
            atomicMin(a.reserve + c1, v);
            atomicMin(a.reserve + c2, v);
            atomicMin(a.reserve + c3, v);
            atomicMin(a.reserve + c4, v);
        }
        grid.sync();
        u32 und = 0;
        for (u32 r = tid; r < active; r += nthreads) {
            if (__ldcg(&a.status[r]) != 0) continue;
            const u32* c = a.cells + (size_t)r * 5;
            unsigned long long v = ((unsigned long long)(~stamp) << 32) | r;
            u32 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4];
            if (__ldcg(a.reserve + c0) == v && __ldcg(a.reserve + c1) == v && __ldcg(a.reserve + c2) == v && __ldcg(a.reserve + c3) == v &&
                __ldcg(a.reserve + c4) == v) {
                a.status[r] = 1;
                a.touched[c0] = a.round_stamp;
                a.touched[c1] = a.round_stamp;
                a.touched[c2] = a.round_stamp;
                a.touched[c3] = a.round_stamp;
                a.touched[c4] = a.round_stamp;
            } else {
                und++;
            }
        }
        if (und) atomicAdd(&st->undecided[iter % 3], und);
        grid.sync();
        if (__ldcg(&st->undecided[iter % 3]) == 0) break;
    }
    // ---- apply: reinsert_node without its two refits (reinsertion.rs:336-382); accepted entries are disjoint
    u32 applied = 0;
    for (u32 r = tid; r < active; r += nthreads) {
        if (__ldcg(&a.status[r]) != 1) continue;
        const u32* c = a.cells + (size_t)r * 5;
        u32 to = c[0], from = c[1], sib = c[2], parent_id = c[4];
        Node32 sibling_node = load_node_cg(a.nodes + sib);
        Node32 dst_node = load_node_cg(a.nodes + to);
        Node32 new_to = dst_node;
        new_to.prim_count = 0;  // make_inner(left_sibling(from)); its box is refit below
        new_to.first_index = left_sibling_id(from);
        store_node(a.nodes + to, new_to);
        store_node(a.nodes + sib, dst_node);
        store_node(a.nodes + parent_id, sibling_node);
        if (dst_node.prim_count == 0) {
            a.parents[dst_node.first_index] = sib;
            a.parents[dst_node.first_index + 1] = sib;
        }
        if (sibling_node.prim_count == 0) {
            a.parents[sibling_node.first_index] = parent_id;
            a.parents[sibling_node.first_index + 1] = parent_id;
        }
        a.parents[sib] = to;
        a.parents[from] = to;
        applied++;
    }
    if (applied) atomicAdd(&st->applied, applied);
    grid.sync();
    // ---- dirty paths: every accepted entry dirties `to` and the old parent of `from`, and all their ancestors.
    // pending[x] = arrivals node x waits for: one per dirty child plus one self token when x is a start node.
    for (u32 r = tid; r < active; r += nthreads) {
        if (__ldcg(&a.status[r]) != 1) continue;
        const u32* c = a.cells + (size_t)r * 5;
        u32 starts[2] = {c[0], c[4]};
        for (int k = 0; k < 2; k++) {
            u32 node = starts[k];
            atomicAdd(&a.pending[node], 1u);  // self token, released by this entry in the refit phase
            if (atomicExch(&a.mark[node], a.round_stamp) == a.round_stamp) continue;
            while (node != 0) {
                u32 p = __ldcg(&a.parents[node]);
                atomicAdd(&a.pending[p], 1u);
                if (atomicExch(&a.mark[p], a.round_stamp) == a.round_stamp) break;
                node = p;
            }
        }
    }
    grid.sync();
    // ---- refit: whoever brings pending[x] to zero refits x (first.union(second)) and carries on to its parent
    for (u32 r = tid; r < active; r += nthreads) {
        if (__ldcg(&a.status[r]) != 1) continue;
        const u32* c = a.cells + (size_t)r * 5;
        u32 starts[2] = {c[0], c[4]};
        for (int k = 0; k < 2; k++) {
            u32 node = starts[k];
            for (;;) {
                if (atomicSub(&a.pending[node], 1u) != 1u) break;  // somebody below is still due
                Node32 me = load_node_cg(a.nodes + node);
                if (me.prim_count == 0) {
                    Node32 c0 = load_node_cg(a.nodes + me.first_index), c1 = load_node_cg(a.nodes + me.first_index + 1);
                    store_node(a.nodes + node, make_node32(box_union(node_box(c0), node_box(c1)), 0u, me.first_index));
                }
                if (node == 0) break;
                __threadfence();
                node = __ldcg(&a.parents[node]);
            }
        }
    }
    if (tid == 0) st->active = 0;  // next round starts from zero (nobody reads `active` after the last barrier)
}

}  // namespace

// Scratch of one optimisation run + one round of optimize_candidates (reinsertion.rs:141-208): find_reinsertion for the
// first `count` candidates, stable sort by descending gain, conflict-checked apply (greedy in gain order) + refit.
struct ReinsertRun {
    ObvhsContext* ctx;
    ObvhsBvh2* bvh;
    DevBuf<u32> r_from, r_to, gkeys, gkeys_alt, gvals, gvals_alt, cells, status, touched, mark, pending;
    DevBuf<float> r_diff;
    DevBuf<unsigned long long> reserve;
    DevBuf<ReinsertState> st;
    int coop_blocks = 0;

    int init(ObvhsContext* c, ObvhsBvh2* b, size_t max_count) {
        ctx = c;
        bvh = b;
        cudaStream_t s = ctx->stream;
        const size_t len = bvh->node_count;
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
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&coop_blocks, reinsert_resolve_apply_kernel<false>, RESOLVE_THREADS, 0));
        coop_blocks = std::max(1, coop_blocks) * ctx->sm_count;
        return OBVHS_OK;
    }

    int round(const u32* cand_ids, u32 count, u32 round_stamp) {
        cudaStream_t s = ctx->stream;
        TraceScope* tsp = new TraceScope(ctx, "  reins_find");
        find_reinsertion_kernel<<<div_up(count, 128), 128, 0, s>>>(bvh->nodes, bvh->parents, cand_ids, count, r_from.p, r_to.p, r_diff.p,
                                                                  gkeys.p, gvals.p);
        KERNEL_CHECK(ctx);
        delete tsp;
        TraceScope ts(ctx, "  reins_gain_sort_resolve");
        u32 *gk, *order;
        ST_TRY(radix_sort_pairs_u32(ctx, gkeys.p, gkeys_alt.p, gvals.p, gvals_alt.p, count, 4, &gk, &order));
        ResolveArgs ra;
        ra.order = order; ra.r_from = r_from.p; ra.r_to = r_to.p; ra.r_diff = r_diff.p; ra.count = count;
        ra.cells = cells.p; ra.status = status.p; ra.st = st.p; ra.touched = touched.p; ra.reserve = reserve.p;
        ra.mark = mark.p; ra.pending = pending.p; ra.nodes = bvh->nodes; ra.parents = bvh->parents; ra.round_stamp = round_stamp;
        void* args[] = {&ra};
        if (count <= 4 * RESOLVE_THREADS) {  // one CTA, block-level barriers, a plain launch
            reinsert_resolve_apply_kernel<true><<<1, RESOLVE_THREADS, 0, s>>>(ra);
        } else {
            int blocks = std::min(coop_blocks, std::max(1, div_up(count, RESOLVE_THREADS)));
            CU_TRY(ctx, cudaLaunchCooperativeKernel((void*)reinsert_resolve_apply_kernel<false>, dim3(blocks), dim3(RESOLVE_THREADS), args, 0, s));
        }
        KERNEL_CHECK(ctx);
        return OBVHS_OK;
    }

    int finish(u64* applied_out) {
        u32* h = reinterpret_cast<u32*>(ctx->pinned);
        CU_TRY(ctx, cudaMemcpyAsync(h, st.p, sizeof(ReinsertState), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (h[5]) {
            OBVHS_SET_ERR(ctx, "reinsertion: conflict resolution did not converge");
            return OBVHS_ERR_CUDA;
        }
        if (applied_out) *applied_out = h[4];
        return OBVHS_OK;
    }
};

static int reinsertion_prologue(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    if (bvh->max_depth > 96) {
        OBVHS_SET_ERR(ctx, "reinsertion: max_depth %zu > 96 needs a heap stack (faststack.rs) -- not supported", bvh->max_depth);
        return OBVHS_ERR_UNSUPPORTED;
    }
    if (!bvh->parents) ST_TRY(bvh2_compute_parents_device(ctx, bvh));  // init_parents_if_uninit
    bvh->children_are_ordered_after_parents = false;                    // reinsertion.rs:93,113
    return OBVHS_OK;
}

// ReinsertionOptimizer::run_with_candidates (reinsertion.rs:66-90, 113-118): the given node ids, in the given order, are
// the candidates of every one of `iterations` rounds.
int reinsertion_run_candidates_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const u32* d_node_ids, size_t n, u32 iterations, u64* applied_out) {
    if (applied_out) *applied_out = 0;
    if (bvh->node_count <= 1 || n == 0 || iterations == 0) return OBVHS_OK;  // reinsertion.rs:69-71
    if (iterations > 60000) {
        OBVHS_SET_ERR(ctx, "reinsertion: too many iterations (%u)", iterations);
        return OBVHS_ERR_UNSUPPORTED;
    }
    ST_TRY(reinsertion_prologue(ctx, bvh));
    ReinsertRun run;
    ST_TRY(run.init(ctx, bvh, n));
    for (u32 k = 0; k < iterations; k++) ST_TRY(run.round(d_node_ids, (u32)n, k + 1));
    return run.finish(applied_out);
}

int reinsertion_run_device(ObvhsContext* ctx, ObvhsBvh2* bvh, float ratio, const float* seq, size_t n_seq, u64* applied_out) {
    if (applied_out) *applied_out = 0;
    cudaStream_t s = ctx->stream;
    const size_t len = bvh->node_count;
    if (len == 0 || !(ratio > 0.0f)) return OBVHS_OK;  // reinsertion.rs:43-45 (NaN ratio: `<=` is false in Rust; treated as no-op here)
    if (len == 1) return OBVHS_OK;                      // root is a leaf
    ST_TRY(reinsertion_prologue(ctx, bvh));
    std::vector<float> default_seq;
    if (!seq) {
        for (int k = 1; k < 32; k += 2) default_seq.push_back(1.0f / (float)k);
        seq = default_seq.data();
        n_seq = default_seq.size();
    }
    if (n_seq > 60000) {
        OBVHS_SET_ERR(ctx, "reinsertion: ratio sequence too long (%zu)", n_seq);
        return OBVHS_ERR_UNSUPPORTED;
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
    DevBuf<u32> ckeys, ckeys_alt, cvals, cvals_alt;
    CU_TRY(ctx, ckeys.alloc(max_take, s));
    CU_TRY(ctx, ckeys_alt.alloc(max_take, s));
    CU_TRY(ctx, cvals.alloc(max_take, s));
    CU_TRY(ctx, cvals_alt.alloc(max_take, s));
    ReinsertRun run;
    ST_TRY(run.init(ctx, bvh, max_count));
    for (size_t k = 0; k < n_seq; k++) {
        const size_t nc = node_counts[k];
        const u32 take = (u32)std::min(len, nc * 2), count = (u32)(nc - 1);
        if (count == 0 || take < 2) continue;
        const u32 m = take - 1;
        u32 *sk, *cand_ids;
        {
            TraceScope ts(ctx, "  reins_candidates_sort");
            cand_init_kernel<<<div_up(m, 256), 256, 0, s>>>(bvh->nodes, m, ckeys.p, cvals.p);
            KERNEL_CHECK(ctx);
            ST_TRY(radix_sort_pairs_u32(ctx, ckeys.p, ckeys_alt.p, cvals.p, cvals_alt.p, m, 4, &sk, &cand_ids));
        }
        ST_TRY(run.round(cand_ids, count, (u32)k + 1));
    }
    return run.finish(applied_out);
}
