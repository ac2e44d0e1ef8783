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
#include <cstdlib>

#include "common.cuh"
#include "sort_tile.cuh"

// A whole ReinsertionOptimizer::run is ONE cooperative launch (reinsertion_run_kernel): every round's candidate keys, candidate
// sort, find_reinsertion, gain sort, conflict resolution, apply and refit are phases of the same kernel, separated by a
// barrier over the CTAs that take part in the round. The rounds of a run are a chain of small dependent steps (16 rounds x
// ~15 barriers): as 80 separate launches they cost ~190 us per round whatever their size (3.7 ms of a 10.9 ms build at 10 M
// triangles, 1.45 ms of 2.3 ms on the kitchen); as one launch 3.0 ms and 1.2 ms (per-phase times: OBVHS_TRACE=1).
//   * per-round participation: round r runs on K_r CTAs, sized for its candidate count; a barrier over 50 CTAs is much cheaper
//     than one over 592, and the later rounds are small (batch ~ 1/(2k-1)). CTAs beyond K_r wait for the round's done flag.
//   * the conflict-resolution loop finishes in ONE CTA (block barriers) once few entries are undecided.
namespace cg = cooperative_groups;

namespace {

// rdst RadixKey for f32: total order preserving map (negative -> flip all, else flip sign)
__device__ __forceinline__ u32 f32_radix_key(float f) {
    u32 u = __float_as_uint(f);
    u32 mask = (u32)((int)u >> 31) | 0x80000000u;
    return u ^ mask;
}

// Node reads of the read-only phases (candidate keys, find_reinsertion): ordinary L1-cached loads. The tree is rewritten by the
// apply / refit phases of the previous round of the SAME launch, but every phase boundary is a RoundBarrier whose
// __threadfence() is MEMBAR.SC.GPU + CCTL.IVALL on sm_100a: the SM's L1 is invalidated before the phase starts, exactly what
// cooperative_groups' grid.sync() relies on. (Reading through L2 only, ld.global.cg, made the branch-and-bound search 4-6x
// slower: it re-reads the upper levels of the tree constantly. ld.global.nc is not an option: not covered by the fence.)
template <bool NC>
__device__ __forceinline__ Node32 find_load_node(const Node32* nodes, u32 id) {
    if (NC) {
        const float4* q = reinterpret_cast<const float4*>(nodes + id);
        float4 a = __ldg(q), b = __ldg(q + 1);
        Node32 n;
        n.minx = a.x; n.miny = a.y; n.minz = a.z; n.prim_count = __float_as_uint(a.w);
        n.maxx = b.x; n.maxy = b.y; n.maxz = b.z; n.first_index = __float_as_uint(b.w);
        return n;
    }
    return load_node(nodes + id);
}
__device__ __forceinline__ float node_half_area_cg(const Node32* nodes, u32 id) { return box_half_area(node_box(load_node(nodes + id))); }

constexpr int RSTACK = 192;  // fast_stack!((f32,u32), (96,192), max_depth*2, ...) reinsertion.rs:287 with max_depth <= 96

// The (f32, u32) stack of find_reinsertion: 192 entries of local memory (StackStack), or -- for max_depth > 96, where the
// reference switches to HeapStack::new_with_capacity(max_depth * 2) (faststack.rs:44-47) -- a slice of a global arena,
// interleaved over the threads of the grid so that a warp's accesses at equal depth coalesce. Two types, so that the common
// case carries no pointers and no run-time choice.
struct LocalFindStack {
    float s_area[RSTACK];
    u32 s_id[RSTACK];
    __device__ __forceinline__ u32 cap() const { return RSTACK; }
    __device__ __forceinline__ void put(u32 i, float a, u32 id) {
        s_area[i] = a;
        s_id[i] = id;
    }
    __device__ __forceinline__ void get(u32 i, float& a, u32& id) const {
        a = s_area[i];
        id = s_id[i];
    }
};
struct HeapFindStack {
    float* h_area;
    u32* h_id;
    u32 stride, capacity;
    __device__ __forceinline__ u32 cap() const { return capacity; }
    __device__ __forceinline__ void put(u32 i, float a, u32 id) {
        h_area[(size_t)i * stride] = a;
        h_id[(size_t)i * stride] = id;
    }
    __device__ __forceinline__ void get(u32 i, float& a, u32& id) const {
        a = h_area[(size_t)i * stride];
        id = h_id[(size_t)i * stride];
    }
};

// K9: reinsertion.rs:233-334. Stack pushes saturate at the last slot and pop_fast saturates at 0 (faststack.rs:299-310).
// The reads of pivot level k+1 (sibling node, pivot node, the pivot's parent index) are issued at the start of level k: their
// addresses are known one level ahead, the tree is read-only in this phase, and the values are consumed only where the reference
// reads them. The first stack entry of a level (always the sibling) skips the push/pop round trip through the stack.
// -DOBVHS_FIND_DEBUG (build.build_variant) adds per-round search statistics to the OBVHS_TRACE print-out.
#ifndef OBVHS_RANK_SORT_MAX
#define OBVHS_RANK_SORT_MAX 6144  // (the keys must fit the sort tile's shared memory: 26 KB)
#endif
#ifndef OBVHS_RANK_LANES
#define OBVHS_RANK_LANES 8
#endif
#ifndef OBVHS_FIND_PER_CTA
#define OBVHS_FIND_PER_CTA 128  // searches per 256-thread CTA of a round (64: dynamic frames +1 %, kitchen build -2 %; 256: -3 % / 0; 16 rank lanes: +3 % / -2 %)
#endif
#ifdef OBVHS_FIND_DEBUG
#define FIND_DBG(...) __VA_ARGS__
#else
#define FIND_DBG(...)
#endif
template <bool NC, class Stack>
__device__ __forceinline__ void find_reinsertion(const Node32* nodes, const u32* parents, u32 node_id, Stack& stk, u32& out_from, u32& out_to,
                                                 float& out_diff, const bool keep_next FIND_DBG(, u32& dbg_visits, u32& dbg_pops, u32& dbg_load_cycles, u32& dbg_get_cycles)) {
    const u32 cap1 = stk.cap() - 1u;
    u32 sp = 0;
    u32 best_to = 0;
    float best_diff = 0.0f;
    const u32 parent_id = NC ? __ldg(parents + node_id) : parents[node_id];
    const Node32 self = find_load_node<NC>(nodes, node_id);
    u32 sib = sibling_id(node_id);
    Node32 sib_node = find_load_node<NC>(nodes, sib);
    Node32 piv_node = find_load_node<NC>(nodes, parent_id);
    u32 next_pivot = NC ? __ldg(parents + parent_id) : parents[parent_id];
    const Box aabb = node_box(self);
    const float node_area = box_half_area(aabb);
    float area_diff = box_half_area(node_box(piv_node));  // parent_area
    Box pivot_bbox = node_box(sib_node);
    u32 pivot_id = parent_id;
    // ONE flat loop: an iteration either pops an entry (skipping pruned ones) or moves to the next pivot and takes that level's
    // first entry. The reference's two nested loops, run by 32 lanes with 32 different searches, cost the warp the SUM over the
    // levels of the longest subtree search at that level (lanes wait at the end of every inner loop); flat, it costs the longest
    // lane's own total: find of round 0 383 -> 282 us at 10 M triangles, 111 -> 82 us on the kitchen. The late rounds (a few dozen
    // searches of <= 17 visits / 32 pops / 9 levels) stay at ~15 us = ~45 iterations x ~650 cycles of dependent instructions of a
    // lone warp; node loads are ~105 cycles of that (OBVHS_FIND_DEBUG). Requesting the child popped next (first_index + 1) at push
    // time changed nothing there and cost round 0 at 10 M a third (282 -> 377 us: half of those reads are pruned at the pop).
    Node32 n_sib = sib_node, n_piv = piv_node;
    u32 n_next = 0;
    bool started = false;
    // keep_next (rounds with at most one search per lane, where the longest search is the cost: -25 % on the kitchen's first
    // round; with several searches per lane the extra path costs the 10 M scene's first rounds 3-10 %): the entry popped right
    // after an inner node's two pushes is always its second child, with the best gain unchanged in between, so it stays in
    // registers (no store / load round trip between learning first_index and requesting that child), and a pair that the pop test
    // would reject anyway (the best gain only grows) is not pushed at all.
    bool have_next = false;
    float next_area = 0.0f;
    u32 next_id = 0;
    for (;;) {
        float top_area_diff;
        u32 top_sibling_id;
        bool first = false, have;
        if (have_next) {
            top_area_diff = next_area;
            top_sibling_id = next_id;
            have_next = false;
            have = true;
            FIND_DBG(dbg_pops++;)
        } else if (sp == 0) {
            if (started) {  // the level's stack ran empty: reinsertion.rs:315-326
                if (pivot_id != parent_id) {
                    pivot_bbox = box_union(pivot_bbox, node_box(sib_node));
                    area_diff += box_half_area(node_box(piv_node)) - box_half_area(pivot_bbox);
                }
                if (pivot_id == 0) break;
                sib = sibling_id(pivot_id);
                pivot_id = next_pivot;
                sib_node = n_sib;
                piv_node = n_piv;
                next_pivot = n_next;
            }
            started = true;
            if (pivot_id != 0) {  // the reads of the level after this one
                n_sib = find_load_node<NC>(nodes, sibling_id(pivot_id));
                n_piv = find_load_node<NC>(nodes, next_pivot);
                n_next = NC ? __ldg(parents + next_pivot) : parents[next_pivot];
            }
            top_area_diff = area_diff;  // stack.push((area_diff, sibling_id)) + the pop that follows it
            top_sibling_id = sib;
            first = true;
            have = !(top_area_diff - node_area <= best_diff);
        } else {
            do {
                sp = sp - 1;
                stk.get(sp, top_area_diff, top_sibling_id);
                FIND_DBG(dbg_pops++;)
                have = !(top_area_diff - node_area <= best_diff);
            } while (!have && sp != 0);
        }
        if (!have) continue;
        FIND_DBG(dbg_visits++;)
        FIND_DBG(long long l0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(l0));)
        const Node32 dst = first ? sib_node : find_load_node<NC>(nodes, top_sibling_id);
        FIND_DBG(long long l1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(l1) : "f"(dst.minx), "f"(dst.maxx)); dbg_load_cycles += (u32)(l1 - l0);)
        const Box dbox = node_box(dst);
        float merged_area = box_half_area(box_union(dbox, aabb));
        float reinsert_area = top_area_diff - merged_area;
        if (reinsert_area > best_diff) {
            best_to = top_sibling_id;
            best_diff = reinsert_area;
        }
        if (dst.prim_count == 0) {
            float child_area = reinsert_area + box_half_area(dbox);
            if (!keep_next) {
                stk.put(sp, child_area, dst.first_index);
                sp = min(sp + 1u, cap1);
                stk.put(sp, child_area, dst.first_index + 1);
                sp = min(sp + 1u, cap1);
            } else if (!(child_area - node_area <= best_diff)) {
                stk.put(sp, child_area, dst.first_index);
                sp = min(sp + 1u, cap1);
                have_next = true;
                next_area = child_area;
                next_id = dst.first_index + 1;
            }
        }
    }
    u32 from = node_id;
    if (best_to == sibling_id(from) || best_to == parent_id) {  // reinsertion.rs:328-333 -> Reinsertion::default()
        from = 0;
        best_to = 0;
        best_diff = 0.0f;
    }
    out_from = from;
    out_to = best_to;
    out_diff = best_diff;
}

struct ReinsertState {
    u32 active;        // entries with area_diff > 0 (they form a prefix of the gain-sorted list)
    u32 undecided[3];  // entries still undecided after resolution iteration i, in slot i % 3
    u32 applied;       // running total of applied reinsertions
    u32 error;         // 1: resolution iteration counter overflow
    u32 iterations;    // resolution iterations of the whole run (tracing)
    u32 barriers;      // barriers block 0 went through (tracing)
};

// One round of the run. m = number of candidate-sort entries (nodes 1..m), 0 when the candidates are given; count = candidates
// searched; blocks = CTAs taking part; *_off = this round's zeroed sort scratch (ghist[4*256] then status[4*tiles*256]).
struct RoundPlan {
    u32 m, count, blocks, cand_off, gain_off;
};

struct RunArgs {
    Node32* nodes;
    u32* parents;
    const RoundPlan* plan;
    u32 n_rounds;
    const u32* fixed_candidates;  // run_with_candidates: the same ids every round (no candidate sort)
    u32 *ckeys, *ckeys_alt, *cvals, *cvals_alt;
    u32 *gkeys, *gkeys_alt, *gvals, *gvals_alt;
    u32* sort_scratch;
    u32* round_barrier;  // one counter per round (zeroed), + done flags behind them
    u32* round_done;
    u32 *r_from, *r_to;
    float* r_diff;
    u32* cells;            // 5 per entry
    u32* status;           // 0 undecided, 1 accepted, 2 rejected / inactive
    u32* und_list[2];      // ranks still undecided, ping-pong between resolution iterations
    ReinsertState* st;
    u32* touched;          // per node, == round_stamp when touched this round
    unsigned long long* reserve;  // per node, min over ((~stamp) << 32 | rank)
    u32* mark;             // per node, == round_stamp when on a dirty path
    u32* pending;          // per node, arrivals still due before the node can be refit
    unsigned long long* trace_ns;  // OBVHS_TRACE: globaltimer of block 0 at the end of each phase, 10 per round (or null)
    unsigned long long* find_dbg;  // -DOBVHS_FIND_DEBUG builds: per round {max cycles, sum cycles, max visits, max pops} of one search
    float* heap_area;      // find_reinsertion stacks for max_depth > 96 (null: local memory)
    u32* heap_id;
    u32 heap_cap;
};

__device__ __forceinline__ void phase_stamp(const RunArgs& a, u32 round, u32 phase) {
    if (a.trace_ns && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace_ns[round * 10 + phase] = t;
    }
}

// atomicSub(p, 1) with release semantics: the stores before it (a refitted box) are visible to whoever observes the new count.
// MEMBAR.ALL.GPU + ATOMG; unlike __threadfence() (MEMBAR.SC.GPU + CCTL.IVALL) it leaves the SM's L1 alone.
__device__ __forceinline__ u32 atomic_dec_release(u32* p) {
    u32 old;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(0xffffffffu) : "memory");
    return old;
}

// Barrier over the first `k` CTAs of the grid on a monotonic counter (one counter per round, zeroed by the host). All CTAs of
// a cooperative launch are co-resident, so spinning is safe.
struct RoundBarrier {
    u32* counter;
    u32 k, gen;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            gen++;
            __threadfence();
            if (k > 1) {  // (a round of ONE CTA only needs the fence: it still invalidates the SM's L1 for the words atomics changed)
                atomicAdd(counter, 1u);
                const u32 target = gen * k;
                while (*reinterpret_cast<volatile u32*>(counter) < target) {
                }
                __threadfence();
            }
        }
        __syncthreads();
    }
};

// Stable LSD radix sort of n (key, value) pairs by the k CTAs of the round; the four digit histograms are already in ghist.
// Returns (through the references) the buffers that hold the result.
__device__ __forceinline__ void round_sort(u32*& keys, u32*& keys_alt, u32*& vals, u32*& vals_alt, u32 n, u32* ghist, u32* status, RoundBarrier& bar,
                                           unsigned char* smem_raw, u32* s_goffs, u32* s_ws) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 tiles = (n + SortCfg<u32>::TILE - 1) / SortCfg<u32>::TILE;
    for (int p = 0; p < 4; p++) {
        {  // exclusive scan of this pass's 256 bins (every CTA its own copy)
            const u32 c = __ldcg(&ghist[p * 256 + tid]);
            u32 x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_ws[warp] = x;
            __syncthreads();
            u32 wbase = 0;
            for (int k = 0; k < warp; k++) wbase += s_ws[k];
            s_goffs[tid] = wbase + x - c;
            __syncthreads();
        }
        for (u32 tile = blockIdx.x; tile < tiles; tile += bar.k)
            onesweep_tile<u32, true, true>(keys, keys_alt, vals, vals_alt, n, 8 * p, s_goffs, status + (size_t)p * tiles * 256, tile, smem_raw);
        bar.sync();
        u32* t = keys; keys = keys_alt; keys_alt = t;
        t = vals; vals = vals_alt; vals_alt = t;
    }
}

// Small inputs (a few thousand pairs: every round of a kitchen-sized scene, the late rounds of a large one): stable sort by
// COUNTING. Every CTA of the round stages all n keys in shared memory and ranks its own slice of them -- rank(i) = #{j : k_j <
// k_i} + #{j < i : k_j == k_i} -- then scatters straight to the sorted position: one phase, no passes, no histograms. The O(n^2)
// compares are spread over the round's CTAs (n = 4556 on 18 CTAs: 4.5 k compares per thread).
constexpr u32 RANK_SORT_MAX = OBVHS_RANK_SORT_MAX;  // beyond this the four radix passes are cheaper (one lane per key: 4555 keys 46 us by counting, 20 us by radix)
__device__ __forceinline__ void rank_sort(const u32* keys, const u32* vals, u32* vals_out, u32 n, RoundBarrier& bar, unsigned char* smem_raw) {
    u32* sk = reinterpret_cast<u32*>(smem_raw);
    const u32 n4 = (n + 3u) & ~3u;
    for (u32 j = threadIdx.x; j < n4; j += blockDim.x) sk[j] = j < n ? __ldcg(keys + j) : 0xffffffffu;  // padding ranks behind every key
    __syncthreads();
    const u32 nthreads = bar.k * blockDim.x;
    const uint4* sk4 = reinterpret_cast<const uint4*>(sk);
    // S = 1, 2, 4 or 8 adjacent lanes share one key when the round has the threads for it: each scans a slice of the keys.
    // Measured on the 1 M-triangle dynamic frames (16 rounds of 640-20 k candidates): 1.42 -> 1.27 ms per frame against one lane per
    // key up to 2048 keys and four radix passes beyond (2048 keys x 8 lanes 1.31 ms, 4096 x 4 1.34, 4096 x 8 1.29, 6144 x 8 1.27)
    u32 S = 1;
    while (S < OBVHS_RANK_LANES && 2 * S * n <= nthreads) S *= 2;
    const u32 part = threadIdx.x & (S - 1u);
    const u32 q4 = n4 / 4, per = (q4 + S - 1) / S, j4_lo = part * per, j4_hi = min(q4, j4_lo + per);
    for (u32 i0 = (blockIdx.x * blockDim.x + threadIdx.x) / S; i0 < ((n + 31u) & ~31u); i0 += nthreads / S) {  // warp-uniform trip count
        const bool live = i0 < n;
        const u32 i = live ? i0 : 0u;
        const u32 ki = sk[i];
        u32 rank = 0;
        // keys before i count when <=, keys after i when < (a stable rank); four keys per shared-memory load
        for (u32 j4 = j4_lo; j4 < j4_hi; j4++) {
            const uint4 k = sk4[j4];
            const u32 j = j4 * 4;
            rank += (k.x < ki || (k.x == ki && j < i)) ? 1u : 0u;
            rank += (k.y < ki || (k.y == ki && j + 1 < i)) ? 1u : 0u;
            rank += (k.z < ki || (k.z == ki && j + 2 < i)) ? 1u : 0u;
            rank += (k.w < ki || (k.w == ki && j + 3 < i)) ? 1u : 0u;
        }
        for (u32 o = 1; o < S; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (live && part == 0) vals_out[rank] = __ldcg(vals + i);
    }
    bar.sync();
}

// the four 8-bit digit histograms of the keys a CTA produces, accumulated in shared memory and flushed to ghist
__device__ __forceinline__ void hist_clear(u32* sh) {
    for (int t = threadIdx.x; t < 4 * 256; t += blockDim.x) sh[t] = 0;
    __syncthreads();
}
__device__ __forceinline__ void hist_add(u32* sh, u32 key) {
#pragma unroll
    for (int p = 0; p < 4; p++) atomicAdd(&sh[p * 256 + ((key >> (8 * p)) & 0xffu)], 1u);
}
__device__ __forceinline__ void hist_flush(u32* sh, u32* ghist) {
    __syncthreads();
    for (int t = threadIdx.x; t < 4 * 256; t += blockDim.x) {
        const u32 c = sh[t];
        if (c) atomicAdd(&ghist[t], c);
    }
    __syncthreads();
}

constexpr int RUN_THREADS = SORT_THREADS;   // 256: the sort tile routine is written for it
constexpr int RUN_MIN_CTAS = 2;             // two CTAs per SM: barriers over 296 CTAs cost less than the extra search lanes of 592 bring
constexpr u32 RESOLVE_SINGLE_BELOW = 256;   // undecided entries from which ONE CTA finishes the resolution loop (one entry per thread)

// K10 conflict resolution + apply + refit of one round: the reference applies the gain-sorted list SEQUENTIALLY, skipping
// entries that touch a node already touched (see the file header of the previous revision / DESIGN.md K10): the accepted set is
// the greedy maximal independent set in rank order over static cells, computed with deterministic reservations.
__device__ __forceinline__ void resolve_round(const RunArgs& a, const u32* order, u32 count, u32 round_stamp, RoundBarrier& bar) {
    const u32 round = round_stamp - 1;
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = bar.k * blockDim.x;
    ReinsertState* st = a.st;
    if (tid == 0) st->undecided[0] = st->undecided[1] = st->undecided[2] = 0;
    // ---- cells of entry r (gain order): {to, from, sibling(from), parent(to), parent(from)} (reinsertion.rs:198-208)
    for (u32 r = tid; r < count; r += nthreads) {
        u32 j = __ldcg(order + r);
        float d = __ldcg(a.r_diff + j);
        if (!(d > 0.0f)) {  // reinsertion.rs:178-180
            a.status[r] = 2;
            continue;
        }
        u32 from = __ldcg(a.r_from + j), to = __ldcg(a.r_to + j);
        u32* c = a.cells + (size_t)r * 5;
        c[0] = to;
        c[1] = from;
        c[2] = sibling_id(from);
        c[3] = __ldcg(a.parents + to);
        c[4] = __ldcg(a.parents + from);
        a.status[r] = 0;
        atomicMax(&st->active, r + 1);
    }
    bar.sync();
    phase_stamp(a, round, 5);
    const u32 active = __ldcg(&st->active);
    // ---- resolution. Every iteration works on the list of ranks still undecided (iteration 1: all active ranks) and appends
    // the ones it leaves undecided to the other list. Grid-wide while the list is long, then block 0 alone (block barriers only).
    bool single = active <= RESOLVE_SINGLE_BELOW;
    u32 iter = 1, n_in = active;
    for (;; iter++) {
        if (iter > 0xffffu) {
            if (tid == 0) st->error = 1;
            break;
        }
        if (single && blockIdx.x != 0) break;  // block 0 finishes; everybody meets at the barrier below the loop
        const u32 stride = single ? blockDim.x : nthreads;
        const u32 first = single ? threadIdx.x : tid;
        const u32 stamp = (round_stamp << 16) | iter;  // strictly increasing over the whole run
        const u32* lin = a.und_list[iter & 1];
        u32* lout = a.und_list[(iter + 1) & 1];
        if (first == 0) st->undecided[(iter + 1) % 3] = 0;  // slot of the NEXT iteration; the previous one may still be read
        for (u32 q = first; q < n_in; q += stride) {
            const u32 r = iter == 1 ? q : __ldcg(lin + q);
            if (__ldcg(&a.status[r]) != 0) continue;
            const u32* c = a.cells + (size_t)r * 5;
            u32 c0 = __ldcg(c), c1 = __ldcg(c + 1), c2 = __ldcg(c + 2), c3 = __ldcg(c + 3), c4 = __ldcg(c + 4);
            if (__ldcg(&a.touched[c0]) == round_stamp || __ldcg(&a.touched[c1]) == round_stamp || __ldcg(&a.touched[c2]) == round_stamp ||
                __ldcg(&a.touched[c3]) == round_stamp || __ldcg(&a.touched[c4]) == round_stamp) {
                __stcg(&a.status[r], 2u);
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
        if (single) {
            __threadfence();
            __syncthreads();
        } else bar.sync();
        for (u32 q0 = first - (first & 31u); q0 < n_in; q0 += stride) {  // warp-uniform trip count: one list reservation per warp
            const u32 q = q0 + (threadIdx.x & 31u);
            bool und = false;
            u32 r = 0;
            if (q < n_in) {
                r = iter == 1 ? q : __ldcg(lin + q);
                if (__ldcg(&a.status[r]) == 0) {
                    const u32* c = a.cells + (size_t)r * 5;
                    unsigned long long v = ((unsigned long long)(~stamp) << 32) | r;
                    u32 c0 = __ldcg(c), c1 = __ldcg(c + 1), c2 = __ldcg(c + 2), c3 = __ldcg(c + 3), c4 = __ldcg(c + 4);
                    if (__ldcg(a.reserve + c0) == v && __ldcg(a.reserve + c1) == v && __ldcg(a.reserve + c2) == v && __ldcg(a.reserve + c3) == v &&
                        __ldcg(a.reserve + c4) == v) {
                        __stcg(&a.status[r], 1u);
                        __stcg(&a.touched[c0], round_stamp);
                        __stcg(&a.touched[c1], round_stamp);
                        __stcg(&a.touched[c2], round_stamp);
                        __stcg(&a.touched[c3], round_stamp);
                        __stcg(&a.touched[c4], round_stamp);
                    } else {
                        und = true;
                    }
                }
            }
            const u32 bal = __ballot_sync(0xffffffffu, und);
            if (bal) {
                u32 base = 0;
                if ((threadIdx.x & 31u) == 0) base = atomicAdd(&st->undecided[iter % 3], (u32)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (und) lout[base + __popc(bal & ((1u << (threadIdx.x & 31u)) - 1u))] = r;
            }
        }
        if (single) {
            __threadfence();
            __syncthreads();
        } else bar.sync();
        const u32 left = __ldcg(&st->undecided[iter % 3]);
        if (left == 0) break;
        n_in = left;
        if (!single && left <= RESOLVE_SINGLE_BELOW) single = true;  // (every CTA reads the same count: a uniform decision)
    }
    if (tid == 0) atomicAdd(&st->iterations, iter);
    bar.sync();
    phase_stamp(a, round, 6);
    // ---- apply: reinsert_node without its two refits (reinsertion.rs:336-382); accepted entries are disjoint
    u32 applied = 0;
    for (u32 r = tid; r < active; r += nthreads) {
        if (__ldcg(&a.status[r]) != 1) continue;
        const u32* c = a.cells + (size_t)r * 5;
        u32 to = __ldcg(c), from = __ldcg(c + 1), sib = __ldcg(c + 2), parent_id = __ldcg(c + 4);
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
    bar.sync();
    phase_stamp(a, round, 7);
    // ---- dirty paths: every accepted entry dirties `to` and the old parent of `from`, and all their ancestors.
    // pending[x] = arrivals node x waits for: one per dirty child plus one self token when x is a start node.
    // (Both walks below are chains of dependent L2 round trips, one tree level after the other: the next level's parent index is
    // requested together with the current level's atomic, and the refit carries the box it just computed upwards.)
    for (u32 r = tid; r < active; r += nthreads) {
        if (__ldcg(&a.status[r]) != 1) continue;
        const u32* c = a.cells + (size_t)r * 5;
        u32 starts[2] = {__ldcg(c), __ldcg(c + 4)};
        for (int k = 0; k < 2; k++) {
            u32 node = starts[k];
            atomicAdd(&a.pending[node], 1u);  // self token, released by this entry in the refit phase
            u32 p = node != 0 ? __ldcg(&a.parents[node]) : 0u;
            if (atomicExch(&a.mark[node], round_stamp) == round_stamp) continue;
            while (node != 0) {
                atomicAdd(&a.pending[p], 1u);
                const u32 pp = p != 0 ? __ldcg(&a.parents[p]) : 0u;
                if (atomicExch(&a.mark[p], round_stamp) == round_stamp) break;
                node = p;
                p = pp;
            }
        }
    }
    bar.sync();
    phase_stamp(a, round, 8);
    // ---- refit: whoever brings pending[x] to zero refits x (first.union(second)) and carries on to its parent. Above the
    // start node the children of x are the node we come from (its box is in registers) and its sibling, and x's first_index is
    // the left one of the two (bvh2/node.rs:154-180), so a level costs: atomic, sibling load, store + fence.
    for (u32 r = tid; r < active; r += nthreads) {
        if (__ldcg(&a.status[r]) != 1) continue;
        const u32* c = a.cells + (size_t)r * 5;
        u32 starts[2] = {__ldcg(c), __ldcg(c + 4)};
        for (int k = 0; k < 2; k++) {
            u32 node = starts[k];
            if (atomicSub(&a.pending[node], 1u) != 1u) continue;  // somebody below is still due
            const Node32 me = load_node_cg(a.nodes + node);
            u32 par = node != 0 ? __ldcg(&a.parents[node]) : 0u;
            Box cur = node_box(me);
            if (me.prim_count == 0) {
                const Node32 c0 = load_node_cg(a.nodes + me.first_index), c1 = load_node_cg(a.nodes + me.first_index + 1);
                cur = box_union(node_box(c0), node_box(c1));
                store_node(a.nodes + node, make_node32(cur, 0u, me.first_index));
            }
            while (node != 0) {
                const u32 p = par;
                if (atomic_dec_release(&a.pending[p]) != 1u) break;
                const Node32 sn = load_node_cg(a.nodes + sibling_id(node));
                par = p != 0 ? __ldcg(&a.parents[p]) : 0u;
                const bool node_is_left = (node & 1u) != 0;
                cur = node_is_left ? box_union(cur, node_box(sn)) : box_union(node_box(sn), cur);
                store_node(a.nodes + p, make_node32(cur, 0u, left_sibling_id(node)));
                node = p;
            }
        }
    }
    if (tid == 0) st->active = 0;  // the next round starts from zero (nobody reads `active` after the last barrier)
}

template <int MIN_CTAS, bool NC>
__global__ void __launch_bounds__(RUN_THREADS, MIN_CTAS) reinsertion_run_kernel(RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_goffs[256];
    __shared__ u32 s_ws[SORT_WARPS];
    u32* s_hist = reinterpret_cast<u32*>(smem_raw);  // 4 x 256 counters; the sort tiles reuse the same bytes afterwards
    for (u32 round = 0; round < a.n_rounds; round++) {
        const RoundPlan pl = a.plan[round];
        if (round > 0) {  // the tree of the previous round must be complete (its participants may have been other CTAs)
            if (threadIdx.x == 0) {
                while (*reinterpret_cast<volatile u32*>(a.round_done + round - 1) == 0) {
                }
                __threadfence();
            }
            __syncthreads();
        }
        if (pl.count == 0 || pl.blocks == 0) {  // an empty round (the host does not plan any; kept for safety)
            if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile u32*>(a.round_done + round) = 1u;
            continue;
        }
        if (blockIdx.x >= pl.blocks) continue;
        RoundBarrier bar{a.round_barrier + round, pl.blocks, 0};
        const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = pl.blocks * blockDim.x;
        const u32 round_stamp = round + 1;
        const u32* cand_ids = a.fixed_candidates;
        phase_stamp(a, round, 0);
        if (!cand_ids) {
            // ---- K8 candidates (reinsertion.rs:121-139): half areas of nodes [1, m], stable radix sort on the f32 key of -cost
            u32* ghist = a.sort_scratch + pl.cand_off;
            hist_clear(s_hist);
            for (u32 j = tid; j < pl.m; j += nthreads) {
                const u32 key = f32_radix_key(-node_half_area_cg(a.nodes, j + 1));
                a.ckeys[j] = key;
                a.cvals[j] = j + 1;
                hist_add(s_hist, key);
            }
            hist_flush(s_hist, ghist);
            bar.sync();
            phase_stamp(a, round, 1);
            u32 *k0 = a.ckeys, *k1 = a.ckeys_alt, *v0 = a.cvals, *v1 = a.cvals_alt;
            if (pl.m <= RANK_SORT_MAX) {
                rank_sort(k0, v0, v1, pl.m, bar, smem_raw);
                v0 = v1;
            } else round_sort(k0, k1, v0, v1, pl.m, ghist, ghist + 4 * 256, bar, smem_raw, s_goffs, s_ws);
            cand_ids = v0;
        }
        phase_stamp(a, round, 2);
        // ---- K9 find_reinsertion for the first `count` candidates; gain keys: descending area_diff, ties by candidate rank
        u32* ghist = a.sort_scratch + pl.gain_off;
        hist_clear(s_hist);
        const bool keep_next = pl.count <= nthreads;
        for (u32 j = tid; j < pl.count; j += nthreads) {
            u32 from, to;
            float diff;
            if (a.heap_area) {
                HeapFindStack stk{a.heap_area + (size_t)blockIdx.x * blockDim.x + threadIdx.x, a.heap_id + (size_t)blockIdx.x * blockDim.x + threadIdx.x,
                                  gridDim.x * blockDim.x, a.heap_cap};
#ifdef OBVHS_FIND_DEBUG
                u32 dv = 0, dp = 0, dl = 0, dg = 0;
                find_reinsertion<NC>(a.nodes, a.parents, __ldcg(cand_ids + j), stk, from, to, diff, keep_next, dv, dp, dl, dg);
#else
                find_reinsertion<NC>(a.nodes, a.parents, __ldcg(cand_ids + j), stk, from, to, diff, keep_next);
#endif
            } else {
                LocalFindStack stk;
#ifdef OBVHS_FIND_DEBUG
                u32 dv = 0, dp = 0, dl = 0, dg = 0;
                const long long c0 = clock64();
                find_reinsertion<NC>(a.nodes, a.parents, __ldcg(cand_ids + j), stk, from, to, diff, keep_next, dv, dp, dl, dg);
                const unsigned long long dc = (unsigned long long)(clock64() - c0);
                if (a.find_dbg) {
                    atomicMax(a.find_dbg + round * 8, dc);
                    atomicAdd(a.find_dbg + round * 8 + 1, dc);
                    atomicMax(a.find_dbg + round * 8 + 2, (unsigned long long)dv);
                    atomicMax(a.find_dbg + round * 8 + 3, (unsigned long long)dp);
                    atomicAdd(a.find_dbg + round * 8 + 4, (unsigned long long)dv);
                    atomicAdd(a.find_dbg + round * 8 + 5, (unsigned long long)dp);
                    atomicAdd(a.find_dbg + round * 8 + 6, (unsigned long long)dl);
                    atomicAdd(a.find_dbg + round * 8 + 7, (unsigned long long)dg);
                }
#else
                find_reinsertion<NC>(a.nodes, a.parents, __ldcg(cand_ids + j), stk, from, to, diff, keep_next);
#endif
            }
            a.r_from[j] = from;
            a.r_to[j] = to;
            a.r_diff[j] = diff;
            const u32 key = f32_radix_key(-diff);
            a.gkeys[j] = key;
            a.gvals[j] = j;
            hist_add(s_hist, key);
        }
        hist_flush(s_hist, ghist);
        bar.sync();
        phase_stamp(a, round, 3);
        u32 *k0 = a.gkeys, *k1 = a.gkeys_alt, *v0 = a.gvals, *v1 = a.gvals_alt;
        if (pl.count <= RANK_SORT_MAX) {
            rank_sort(k0, v0, v1, pl.count, bar, smem_raw);
            v0 = v1;
        } else round_sort(k0, k1, v0, v1, pl.count, ghist, ghist + 4 * 256, bar, smem_raw, s_goffs, s_ws);
        phase_stamp(a, round, 4);
        resolve_round(a, v0, pl.count, round_stamp, bar);
        bar.sync();
        phase_stamp(a, round, 9);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.st->barriers += bar.gen;
            __threadfence();
            *reinterpret_cast<volatile u32*>(a.round_done + round) = 1u;
        }
    }
}

}  // namespace

static int reinsertion_prologue(ObvhsContext* ctx, ObvhsBvh2* bvh) {
    if (!bvh->parents) ST_TRY(bvh2_compute_parents_device(ctx, bvh));  // init_parents_if_uninit
    bvh->children_are_ordered_after_parents = false;                    // reinsertion.rs:93,113
    return OBVHS_OK;
}

// rounds: (m, count) per round; candidates given (fixed != nullptr) or selected by area every round
static int reinsertion_launch(ObvhsContext* ctx, ObvhsBvh2* bvh, std::vector<RoundPlan>& plan, const u32* fixed, u64* applied_out) {
    cudaStream_t s = ctx->stream;
    const size_t len = bvh->node_count;
    size_t max_m = 0, max_count = 0;
    for (const RoundPlan& p : plan) {
        max_m = std::max<size_t>(max_m, p.m);
        max_count = std::max<size_t>(max_count, p.count);
    }
    if (max_count == 0) return OBVHS_OK;
    const size_t smem = onesweep_smem<u32>();
    void* kernel = (void*)reinsertion_run_kernel<RUN_MIN_CTAS, false>;
    static PerDevice<int> per_sm_dev;
    int& per_sm = per_sm_dev[ctx->device];
    if (per_sm == 0) {
        CU_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RUN_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    // max_depth > 96: the reference's HeapStack of max_depth * 2 entries per search (faststack.rs:44-47); here a global arena,
    // with fewer resident CTAs so that it stays small
    const bool heap = bvh->max_depth > 96;
    const size_t heap_cap = heap ? bvh->max_depth * 2 : 0;
    int resident = per_sm * ctx->sm_count;
    if (heap) resident = std::min(resident, std::max(ctx->sm_count / 2, (int)(((size_t)1 << 30) / (heap_cap * 8 * RUN_THREADS))));
    if (resident < 1) resident = 1;
    // CTAs per round: one search per thread where possible, one sort tile per CTA, never more than are resident
    size_t scratch_words = 0;
    int grid = 1;
    for (RoundPlan& p : plan) {
        if (p.count == 0) continue;
        size_t want = std::max<size_t>((p.count + OBVHS_FIND_PER_CTA - 1) / OBVHS_FIND_PER_CTA, (std::max<size_t>(p.m, p.count) + SortCfg<u32>::TILE - 1) / SortCfg<u32>::TILE);
        // the counting sort (up to RANK_SORT_MAX keys) costs n / S comparisons per thread with S lanes per key: give it eight lanes per key
        // from 512 keys on, where it is the longest phase of the round
        if (p.m >= 512 && p.m <= RANK_SORT_MAX) want = std::max<size_t>(want, (p.m * OBVHS_RANK_LANES + RUN_THREADS - 1) / RUN_THREADS);
        p.blocks = (u32)std::max<size_t>(1, std::min<size_t>(want, (size_t)resident));
        grid = std::max(grid, (int)p.blocks);
        const size_t tiles_c = (p.m + SortCfg<u32>::TILE - 1) / SortCfg<u32>::TILE, tiles_g = (p.count + SortCfg<u32>::TILE - 1) / SortCfg<u32>::TILE;
        p.cand_off = (u32)scratch_words;
        scratch_words += p.m ? 4 * 256 + 4 * tiles_c * 256 : 0;
        p.gain_off = (u32)scratch_words;
        scratch_words += 4 * 256 + 4 * tiles_g * 256;
        if (scratch_words >= ((size_t)1 << 31)) {
            OBVHS_SET_ERR(ctx, "reinsertion: ratio sequence too long for one run (%zu rounds)", plan.size());
            return OBVHS_ERR_UNSUPPORTED;
        }
    }
    RunArgs a = {};
    DevBuf<RoundPlan> d_plan;
    DevBuf<u32> ckeys, ckeys_alt, cvals, cvals_alt, gkeys, gkeys_alt, gvals, gvals_alt, scratch, rbar, r_from, r_to, cells, status, und0, und1, touched, mark, pending, heap_id;
    DevBuf<float> r_diff, heap_area;
    DevBuf<unsigned long long> reserve, trace_ns, find_dbg;
    DevBuf<ReinsertState> st;
    const size_t n_rounds = plan.size();
    if (ctx->trace) {
        CU_TRY(ctx, trace_ns.alloc(n_rounds * 10, s));
        CU_TRY(ctx, cudaMemsetAsync(trace_ns.p, 0, n_rounds * 80, s));
#ifdef OBVHS_FIND_DEBUG
        CU_TRY(ctx, find_dbg.alloc(n_rounds * 8, s));
        CU_TRY(ctx, cudaMemsetAsync(find_dbg.p, 0, n_rounds * 64, s));
#endif
    }
    CU_TRY(ctx, d_plan.alloc(n_rounds, s));
    if (!fixed) {
        CU_TRY(ctx, ckeys.alloc(max_m, s));
        CU_TRY(ctx, ckeys_alt.alloc(max_m, s));
        CU_TRY(ctx, cvals.alloc(max_m, s));
        CU_TRY(ctx, cvals_alt.alloc(max_m, s));
    }
    CU_TRY(ctx, gkeys.alloc(max_count, s));
    CU_TRY(ctx, gkeys_alt.alloc(max_count, s));
    CU_TRY(ctx, gvals.alloc(max_count, s));
    CU_TRY(ctx, gvals_alt.alloc(max_count, s));
    CU_TRY(ctx, scratch.alloc(scratch_words, s));
    CU_TRY(ctx, rbar.alloc(2 * n_rounds, s));
    CU_TRY(ctx, r_from.alloc(max_count, s));
    CU_TRY(ctx, r_to.alloc(max_count, s));
    CU_TRY(ctx, r_diff.alloc(max_count, s));
    CU_TRY(ctx, cells.alloc(max_count * 5, s));
    CU_TRY(ctx, status.alloc(max_count, s));
    CU_TRY(ctx, und0.alloc(max_count, s));
    CU_TRY(ctx, und1.alloc(max_count, s));
    CU_TRY(ctx, touched.alloc(len, s));
    CU_TRY(ctx, mark.alloc(len, s));
    CU_TRY(ctx, pending.alloc(len, s));
    CU_TRY(ctx, reserve.alloc(len, s));
    CU_TRY(ctx, st.alloc(1, s));
    if (heap) {
        CU_TRY(ctx, heap_area.alloc(heap_cap * (size_t)grid * RUN_THREADS, s));
        CU_TRY(ctx, heap_id.alloc(heap_cap * (size_t)grid * RUN_THREADS, s));
    }
    CU_TRY(ctx, cudaMemcpyAsync(d_plan.p, plan.data(), n_rounds * sizeof(RoundPlan), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemsetAsync(scratch.p, 0, scratch_words * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(rbar.p, 0, 2 * n_rounds * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(touched.p, 0, len * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(mark.p, 0, len * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(pending.p, 0, len * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(reserve.p, 0xff, len * 8, s));
    CU_TRY(ctx, cudaMemsetAsync(st.p, 0, sizeof(ReinsertState), s));
    a.nodes = bvh->nodes; a.parents = bvh->parents; a.plan = d_plan.p; a.n_rounds = (u32)n_rounds; a.fixed_candidates = fixed;
    a.ckeys = ckeys.p; a.ckeys_alt = ckeys_alt.p; a.cvals = cvals.p; a.cvals_alt = cvals_alt.p;
    a.gkeys = gkeys.p; a.gkeys_alt = gkeys_alt.p; a.gvals = gvals.p; a.gvals_alt = gvals_alt.p;
    a.sort_scratch = scratch.p; a.round_barrier = rbar.p; a.round_done = rbar.p + n_rounds;
    a.r_from = r_from.p; a.r_to = r_to.p; a.r_diff = r_diff.p; a.cells = cells.p; a.status = status.p; a.und_list[0] = und0.p; a.und_list[1] = und1.p; a.st = st.p;
    a.touched = touched.p; a.reserve = reserve.p; a.mark = mark.p; a.pending = pending.p;
    a.trace_ns = ctx->trace ? trace_ns.p : nullptr;
    a.find_dbg = find_dbg.p;
    a.heap_area = heap ? heap_area.p : nullptr; a.heap_id = heap ? heap_id.p : nullptr; a.heap_cap = (u32)heap_cap;
    void* args[] = {&a};
    CU_TRY(ctx, cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(RUN_THREADS), args, smem, s));
    KERNEL_CHECK(ctx);
    ReinsertState* h = reinterpret_cast<ReinsertState*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, st.p, sizeof(ReinsertState), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    if (ctx->trace) {
        fprintf(stderr, "[obvhs trace]   reinsertion run: %zu rounds, grid %d x %d, %u resolution iterations, %u barriers, %u applied\n", n_rounds, grid,
                RUN_THREADS, h->iterations, h->barriers, h->applied);
        std::vector<unsigned long long> t(n_rounds * 10);
        CU_TRY(ctx, cudaMemcpy(t.data(), trace_ns.p, n_rounds * 80, cudaMemcpyDeviceToHost));
        static const char* names[9] = {"cand keys", "cand sort", "find", "gain sort", "cells", "resolve", "apply", "dirty", "refit"};
        for (size_t r = 0; r < n_rounds; r++) {
            fprintf(stderr, "[obvhs trace]     round %2zu (m %7u, count %7u, %3u CTAs):", r, plan[r].m, plan[r].count, plan[r].blocks);
            for (int k = 0; k < 9; k++) {
                const unsigned long long t0 = t[r * 10 + k], t1 = t[r * 10 + k + 1];
                fprintf(stderr, " %s %.1f", names[k], t1 >= t0 && t0 ? (t1 - t0) * 1e-3 : 0.0);
            }
            fprintf(stderr, " us\n");
        }
#ifdef OBVHS_FIND_DEBUG
        std::vector<unsigned long long> fd(n_rounds * 8);
        CU_TRY(ctx, cudaMemcpy(fd.data(), find_dbg.p, n_rounds * 64, cudaMemcpyDeviceToHost));
        for (size_t r = 0; r < n_rounds; r++) {
            const unsigned long long* f = &fd[r * 8];
            const double c = plan[r].count;
            fprintf(stderr, "[obvhs trace]     find %2zu: longest search %llu cycles, mean %.0f; visits max %llu mean %.1f; pops max %llu mean %.1f; per visit: node load %.0f cycles; per pop: stack get %.0f cycles\n",
                    r, f[0], f[1] / c, f[2], f[4] / c, f[3], f[5] / c, (double)f[6] / (double)(f[4] ? f[4] : 1), (double)f[7] / (double)(f[5] ? f[5] : 1));
        }
#endif
    }
    if (h->error) {
        OBVHS_SET_ERR(ctx, "reinsertion: conflict resolution did not converge");
        return OBVHS_ERR_CUDA;
    }
    if (applied_out) *applied_out = h->applied;
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
    TraceScope ts(ctx, "reinsertion_optimize_candidates");
    ST_TRY(reinsertion_prologue(ctx, bvh));
    std::vector<RoundPlan> plan(iterations, RoundPlan{0u, (u32)n, 0u, 0u, 0u});
    return reinsertion_launch(ctx, bvh, plan, d_node_ids, applied_out);
}

int reinsertion_run_device(ObvhsContext* ctx, ObvhsBvh2* bvh, float ratio, const float* seq, size_t n_seq, u64* applied_out) {
    if (applied_out) *applied_out = 0;
    const size_t len = bvh->node_count;
    if (len == 0 || !(ratio > 0.0f)) return OBVHS_OK;  // reinsertion.rs:43-45 (NaN ratio: `<=` is false in Rust; treated as no-op here)
    if (len == 1) return OBVHS_OK;                      // root is a leaf
    TraceScope ts(ctx, "reinsertion_optimize");
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
    std::vector<RoundPlan> plan;
    for (size_t k = 0; k < n_seq; k++) {
        volatile float f0 = (float)len * ratio;
        volatile float f = f0 * seq[k];
        double fd = (double)f;
        size_t batch = !(fd > 0.0) ? 0 : (fd >= 18446744073709551616.0 ? ~(size_t)0 : (size_t)fd);  // `as usize` saturates
        if (batch < 1) batch = 1;
        const size_t nc = batch >= len ? len : std::min(len, batch + 1);
        const u32 take = (u32)std::min(len, nc * 2), count = (u32)(nc - 1);
        if (count == 0 || take < 2) continue;
        plan.push_back(RoundPlan{take - 1, count, 0u, 0u, 0u});
    }
    if (plan.empty()) return OBVHS_OK;
    return reinsertion_launch(ctx, bvh, plan, nullptr, applied_out);
}
