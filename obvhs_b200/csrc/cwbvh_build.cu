// cwbvh_build.cu -- BVH2 -> CWBVH (compressed wide BVH8, 80-byte quantised nodes).
//
// Replaces bvh2_to_cwbvh / Bvh2Converter::{calculate_cost_impl, get_children, order_children, convert_to_cwbvh_impl,
// count_primitives} (src/cwbvh/bvh2_to_cwbvh.rs:34-510). The reference is two sequential recursions; here:
//
//   K11 cwbvh_cost_kernel  bottom-up DP (one thread per BVH2 leaf climbs, the second arriver at an inner node owns it):
//                          the 7 Decisions of the node (cost f32 + packed kind/left/right byte), the primitive count P
//                          of its subtree and K = number of CWBVH nodes its subtree produces when it is emitted as a
//                          wide node (1 + sum of K over the INTERNAL children get_children() collects).
//   K12 cwbvh_emit_kernel  top-down, one thread per wide node, one launch per CWBVH level. The reference's pre-order
//                          recursion allocates with two running counters (nodes.len(), primitive_indices.len());
//                          in closed form, for the j-th INTERNAL child c_j of wide node N in slot order:
//                              node(c_j)       = child_base(N) + j
//                              child_base(c_j) = child_base(N) + k(N) + sum_{i<j} (K(c_i) - 1)
//                              prim_base(c_j)  = prim_base(N) + direct_prims(N) + sum_{i<j} P(c_i)
//                          so every node lands at the index and with the bytes the sequential recursion gives.
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"
#include "cwbvh_exponent.h"

namespace cg = cooperative_groups;

namespace {

constexpr u32 KIND_LEAF = 0, KIND_INTERNAL = 1, KIND_DISTRIBUTE = 2;
constexpr float PRIM_COST = 0.3f;           // bvh2_to_cwbvh.rs:30
constexpr float DENOM = 1.0f / 255.0f;      // cwbvh/mod.rs:36-38
constexpr u32 INVALID32 = 0xffffffffu;

// Decision (bvh2_to_cwbvh.rs:470-484), 7 per BVH2 node: cost[i] and meta[i] = kind | left << 2 | right << 5
// (left/right = 7 encodes the reference's INVALID 0xff).
// S[i] = number of CWBVH nodes created BELOW this node when it is expanded with decision i, i.e. the sum over the
// children get_children(node, i) collects of (1 + S[0] of the child) for the INTERNAL ones. It turns the reference's
// running node allocator into a bottom-up quantity that only needs the two children's records. 64 bytes = 4 x 16-B vectors.
struct __align__(16) Dec {
    float cost[7];
    u32 meta_lo;
    u32 meta_hi;
    u32 S[7];
};
static_assert(sizeof(Dec) == 64, "Dec");

__device__ __forceinline__ Dec load_dec_cg(const Dec* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3);
    Dec r;
    r.cost[0] = __uint_as_float(a.x); r.cost[1] = __uint_as_float(a.y); r.cost[2] = __uint_as_float(a.z); r.cost[3] = __uint_as_float(a.w);
    r.cost[4] = __uint_as_float(b.x); r.cost[5] = __uint_as_float(b.y); r.cost[6] = __uint_as_float(b.z); r.meta_lo = b.w;
    r.meta_hi = c.x; r.S[0] = c.y; r.S[1] = c.z; r.S[2] = c.w;
    r.S[3] = d.x; r.S[4] = d.y; r.S[5] = d.z; r.S[6] = d.w;
    return r;
}
__device__ __forceinline__ void store_dec(Dec* p, const Dec& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(__float_as_uint(r.cost[0]), __float_as_uint(r.cost[1]), __float_as_uint(r.cost[2]), __float_as_uint(r.cost[3]));
    q[1] = make_uint4(__float_as_uint(r.cost[4]), __float_as_uint(r.cost[5]), __float_as_uint(r.cost[6]), r.meta_lo);
    q[2] = make_uint4(r.meta_hi, r.S[0], r.S[1], r.S[2]);
    q[3] = make_uint4(r.S[3], r.S[4], r.S[5], r.S[6]);
}
// What the emit pass needs of a record, as ONE 16-byte vector per node: {meta_lo, meta_hi, S[0], P} (P = primitives in the subtree).
// The cost pass writes it next to the 64-byte record and never reads it (the frontier form carries the primitive counts through
// the arrival words and the work items; the climbing form for small trees keeps a 4-byte P array that stays in L2); the emit pass
// then touches one sector per BVH2 node instead of two of the record plus one of P (emit 1.03 -> 0.92 ms at 10 M triangles,
// 1.32 -> 1.12 on the soup).
__device__ __forceinline__ void store_hot(uint4* p, const Dec& d, u32 prims) { *p = make_uint4(d.meta_lo, d.meta_hi, d.S[0], prims); }
__device__ __forceinline__ u32 dec_meta_of(const Dec& d, u32 i) { return ((i < 4 ? d.meta_lo >> (8 * i) : d.meta_hi >> (8 * (i - 4)))) & 0xffu; }

struct CwGlobals {
    u32 error;           // 1: DISTRIBUTE/INVALID decision on the emit path (non-finite costs), 2: child left unassigned, 3: count mismatch
    u32 queue_count[3];  // wide nodes queued by level L for level L+1, in slot L % 3
    u32 emitted;         // wide nodes written
    u32 levels;
    unsigned long long level_ns[48];  // OBVHS_TRACE: globaltimer at the end of each level
    u32 level_len[48];
};

// bvh2_to_cwbvh.rs:246-257: every decision of a leaf is LEAF with cost area * prims * 0.3
__device__ __forceinline__ Dec leaf_dec(const Node32& nd) {
    const float cost_leaf = box_half_area(node_box(nd)) * (float)nd.prim_count * PRIM_COST;
    Dec d;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        d.cost[k] = cost_leaf;
        d.S[k] = 0;
    }
    d.meta_lo = 0;  // kind LEAF, left/right 0 (Decision::default indices)
    d.meta_hi = 0;
    return d;
}

// One inner node of calculate_cost_impl (bvh2_to_cwbvh.rs:259-343) from the records of its two children.
__device__ __forceinline__ Dec inner_dec(const Dec& L, const Dec& R, float ha, u32 num_primitives, u32 max_prims_per_leaf) {
    Dec d;
    u32 meta[7];
    // What child k of a side contributes to S when the parent's decision picks index k for it (get_children,
    // bvh2_to_cwbvh.rs:346-397): a DISTRIBUTE child is expanded (its own S[k]), anything else is collected as a child
    // (a wide node counts itself plus everything below it). Computed once with static indices and carried through the
    // arg-min below, instead of indexing the records with the winning (run-time) indices afterwards.
    const u32 wideL = (dec_meta_of(L, 0) & 3u) == KIND_INTERNAL ? 1u + L.S[0] : 0u;
    const u32 wideR = (dec_meta_of(R, 0) & 3u) == KIND_INTERNAL ? 1u + R.S[0] : 0u;
    u32 effL[7], effR[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        effL[k] = (dec_meta_of(L, k) & 3u) == KIND_DISTRIBUTE ? L.S[k] : wideL;
        effR[k] = (dec_meta_of(R, k) & 3u) == KIND_DISTRIBUTE ? R.S[k] : wideR;
    }
    {  // i = 0
        float cost_leaf = num_primitives <= max_prims_per_leaf ? ((float)num_primitives * ha) * PRIM_COST : __int_as_float(0x7f800000);
        float cost_distribute = __int_as_float(0x7f800000);
        u32 dl = 7, dr = 7, sv = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            float c = L.cost[k] + R.cost[6 - k];
            if (c < cost_distribute) {
                cost_distribute = c;
                dl = k;
                dr = 6 - k;
                sv = effL[k] + effR[6 - k];
            }
        }
        float cost_internal = cost_distribute + ha;
        if (cost_leaf < cost_internal) {
            d.cost[0] = cost_leaf;
            meta[0] = KIND_LEAF | dl << 2 | dr << 5;
        } else {
            d.cost[0] = cost_internal;
            meta[0] = KIND_INTERNAL | dl << 2 | dr << 5;
        }
        d.S[0] = sv;  // 0 while no pair was chosen (dl == dr == 7)
    }
#pragma unroll
    for (int ii = 1; ii < 7; ii++) {
        float cost_distribute = d.cost[ii - 1];
        u32 dl = 7, dr = 7, sv = d.S[ii - 1];
#pragma unroll
        for (int k = 0; k < ii; k++) {
            float c = L.cost[k] + R.cost[ii - k - 1];
            if (c < cost_distribute) {
                cost_distribute = c;
                dl = k;
                dr = ii - k - 1;
                sv = effL[k] + effR[ii - k - 1];
            }
        }
        d.cost[ii] = cost_distribute;
        if (dl != 7) meta[ii] = KIND_DISTRIBUTE | dl << 2 | dr << 5;
        else meta[ii] = meta[ii - 1];  // decisions[node_i] = decisions[node_i - 1] (and with it the same S)
        d.S[ii] = sv;
    }
    d.meta_lo = meta[0] | meta[1] << 8 | meta[2] << 16 | meta[3] << 24;
    d.meta_hi = meta[4] | meta[5] << 8 | meta[6] << 16;
    return d;
}

// K11 (climbing form, used for small trees): calculate_cost_impl (bvh2_to_cwbvh.rs:220-344), bottom-up: one thread per BVH2
// leaf climbs, the second arriver at an inner node computes its record. A climbing thread CARRIES the record of the child it comes from in registers, so only the
// sibling's record is read back; records of leaves are never stored (they are a function of the leaf node itself and are
// recomputed by whoever needs them, here and in the emit pass), which halves the pass's DRAM writes.
constexpr int COST_THREADS = 64;  // small CTAs: a CTA lives as long as its longest climber, and most threads stop after one or two levels
__global__ void __launch_bounds__(COST_THREADS) cwbvh_cost_kernel(const Node32* __restrict__ nodes, const u32* __restrict__ parents, u32 n_nodes,
                                                         u32 max_prims_per_leaf, Dec* dec, uint4* hot, u32* P, u32* arrivals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    Node32 nd = load_node(nodes + i);
    if (nd.prim_count == 0) return;  // start at leaves
    Dec mine = leaf_dec(nd);
    u32 my_prims = nd.prim_count, me = i;
    bool mine_is_leaf = true;
    if (i == 0) {  // a single-leaf tree: the host reads S[0] of the root record
        store_dec(dec, mine);
        store_hot(hot, mine, my_prims);
        P[0] = my_prims;
        return;
    }
    for (;;) {
        const u32 node = parents[me];
        if (!mine_is_leaf) {  // publish before signalling: the sibling's thread (or the emit pass) reads it
            store_dec(dec + me, mine);
            store_hot(hot + me, mine, my_prims);
            P[me] = my_prims;
            __threadfence();
        }
        if (atomicAdd(&arrivals[node], 1u) == 0) return;
        const u32 sib = sibling_id(me);
        const Node32 pn = load_node(nodes + node), sn = load_node(nodes + sib);
        Dec sd = load_dec_cg(dec + sib);  // (unwritten memory when the sibling is a leaf: replaced below)
        u32 sib_prims = __ldcg(&P[sib]);
        if (sn.prim_count != 0) {
            sd = leaf_dec(sn);
            sib_prims = sn.prim_count;
        }
        const bool me_left = (me & 1u) != 0;  // bvh2/node.rs:154-180: the left sibling has the odd index
        const Dec L = me_left ? mine : sd, R = me_left ? sd : mine;
        const float ha = box_half_area(node_box(pn));
        const u32 num_primitives = my_prims + sib_prims;
        const Dec d = inner_dec(L, R, ha, num_primitives, max_prims_per_leaf);
        if (node == 0) {
            store_dec(dec, d);
            store_hot(hot, d, num_primitives);
            P[0] = num_primitives;
            return;
        }
        mine = d;
        my_prims = num_primitives;
        me = node;
        mine_is_leaf = false;
    }
}

// K11 (frontier form). The climbing kernel above runs at 5.6 of 32 lanes per issued instruction: every level halves the live
// lanes of a warp and the survivors of different warps are never packed together. Here a ROUND processes the dense list of
// nodes whose two children are finished (the frontier); the second arriver at a parent appends it to the next round's list
// (one global atomic per CTA chunk). Rounds are separated by grid-wide barriers, which also publish the records, so no
// per-node fences are needed. Round count = height of the tree.
struct CostArgs {
    const Node32* nodes;
    const u32* parents;
    u32 n_nodes, max_prims_per_leaf;
    Dec* dec;
    uint4* hot;
    u32* arrivals;   // per inner node: arrivals in bits 0-1, "left / right child is a leaf" in bits 2 / 3, primitives of the children that arrived in bits 4-31
    uint4* queue[2];  // frontier lists, (n_nodes + 1) / 2 entries each: {node, first_child | left_is_leaf << 30 | right_is_leaf << 31, primitives in the subtree, -}
    u32* qcount;     // [3], slot = round % 3
};
constexpr int FRONT_THREADS = 256;
#ifndef OBVHS_FRONT_MIN_CTAS
#define OBVHS_FRONT_MIN_CTAS 4  // 64 registers (73 unconstrained = 3 CTAs): the pass waits on memory, 1.42 -> 1.35 ms at 10 M triangles
#endif
constexpr int FRONT_MIN_CTAS = OBVHS_FRONT_MIN_CTAS;

// appends `item` of every thread with `flag` to q (count in *qn): block scan + one atomic per call. Whole CTA must call.
__device__ __forceinline__ void frontier_push(bool flag, uint4 item, uint4* q, u32* qn) {
    __shared__ u32 s_w[FRONT_THREADS / 32];
    __shared__ u32 s_base;
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32 bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_w[w] = __popc(bal);
    __syncthreads();
    u32 before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < FRONT_THREADS / 32; k++) {
        const u32 c = s_w[k];
        if ((u32)k < w) before += c;
        total += c;
    }
    if (threadIdx.x == 0 && total) s_base = atomicAdd(qn, total);
    __syncthreads();
    if (flag) q[s_base + before + __popc(bal & ((1u << lane) - 1u))] = item;
}

// Child `child` (with `prims` primitives below it) is finished: count it at its parent and leave a note whether it is a leaf and
// how many primitives it brings. The second arriver knows both children (siblings are adjacent, the left one odd:
// bvh2/node.rs:154-180), what kind each is and the parent's primitive count, so whoever computes the parent can request
// everything it needs -- its own box, and per child either the leaf node or the decision record -- in ONE round trip instead of
// three dependent ones (node -> child nodes -> child records), and no per-node primitive count is read back at all.
// (28 bits of primitives per word: the host refuses trees beyond 2^28 primitives.)
__device__ __forceinline__ bool frontier_arrive(const CostArgs& a, u32 child, bool child_is_leaf, u32 prims, uint4& entry) {
    const u32 parent = a.parents[child];
    const bool left = (child & 1u) != 0;
    const u32 mine = 1u | (child_is_leaf ? (left ? 4u : 8u) : 0u) | prims << 4;
    const u32 old = atomicAdd(&a.arrivals[parent], mine);
    if ((old & 3u) != 1u) return false;
    entry = make_uint4(parent, left_sibling_id(child) | (((old | mine) >> 2) & 3u) << 30, (old >> 4) + prims, 0u);
    return true;
}

// record of inner node e.x from its two finished children (leaf children are recomputed from their nodes)
__device__ __forceinline__ void frontier_node(const CostArgs& a, const uint4 e) {
    const u32 node = e.x, first = e.y & 0x3fffffffu;
    const bool lleaf = (e.y >> 30) & 1u, rleaf = (e.y >> 31) & 1u;
    // all reads first: 2 vectors for a leaf (its node), 4 + the primitive count for an inner child (its record)
    const uint4* lsrc = lleaf ? reinterpret_cast<const uint4*>(a.nodes + first) : reinterpret_cast<const uint4*>(a.dec + first);
    const uint4* rsrc = rleaf ? reinterpret_cast<const uint4*>(a.nodes + first + 1) : reinterpret_cast<const uint4*>(a.dec + first + 1);
    const uint4* msrc = reinterpret_cast<const uint4*>(a.nodes + node);
    const uint4 m0 = __ldcg(msrc), m1 = __ldcg(msrc + 1);
    const uint4 l0 = __ldcg(lsrc), l1 = __ldcg(lsrc + 1), r0 = __ldcg(rsrc), r1 = __ldcg(rsrc + 1);
    uint4 l2 = make_uint4(0, 0, 0, 0), l3 = l2, r2 = l2, r3 = l2;
    if (!lleaf) {
        l2 = __ldcg(lsrc + 2);
        l3 = __ldcg(lsrc + 3);
    }
    if (!rleaf) {
        r2 = __ldcg(rsrc + 2);
        r3 = __ldcg(rsrc + 3);
    }
    auto as_node = [](const uint4 q0, const uint4 q1) {
        Node32 n;
        n.minx = __uint_as_float(q0.x); n.miny = __uint_as_float(q0.y); n.minz = __uint_as_float(q0.z); n.prim_count = q0.w;
        n.maxx = __uint_as_float(q1.x); n.maxy = __uint_as_float(q1.y); n.maxz = __uint_as_float(q1.z); n.first_index = q1.w;
        return n;
    };
    auto as_dec = [](const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3) {
        Dec r;
        r.cost[0] = __uint_as_float(q0.x); r.cost[1] = __uint_as_float(q0.y); r.cost[2] = __uint_as_float(q0.z); r.cost[3] = __uint_as_float(q0.w);
        r.cost[4] = __uint_as_float(q1.x); r.cost[5] = __uint_as_float(q1.y); r.cost[6] = __uint_as_float(q1.z); r.meta_lo = q1.w;
        r.meta_hi = q2.x; r.S[0] = q2.y; r.S[1] = q2.z; r.S[2] = q2.w;
        r.S[3] = q3.x; r.S[4] = q3.y; r.S[5] = q3.z; r.S[6] = q3.w;
        return r;
    };
    Dec L, R;
    if (lleaf) L = leaf_dec(as_node(l0, l1));
    else L = as_dec(l0, l1, l2, l3);
    if (rleaf) R = leaf_dec(as_node(r0, r1));
    else R = as_dec(r0, r1, r2, r3);
    const u32 num_primitives = e.z;
    const Dec d = inner_dec(L, R, box_half_area(node_box(as_node(m0, m1))), num_primitives, a.max_prims_per_leaf);
    store_dec(a.dec + node, d);
    store_hot(a.hot + node, d, num_primitives);
}

// Below this many ready nodes a round is cheaper as plain climbing (the rest of the tree is a few thousand nodes, and every
// further round would cost a grid-wide barrier): each thread takes one ready node and keeps going up while it is the
// second arriver, publishing with a fence as in cwbvh_cost_kernel.
constexpr u32 FRONT_CLIMB_BELOW = 32768;

__global__ void __launch_bounds__(FRONT_THREADS, FRONT_MIN_CTAS) cwbvh_cost_frontier_kernel(CostArgs a) {
    cg::grid_group grid = cg::this_grid();
    const u32 nthreads = gridDim.x * blockDim.x;
    // round 0: every leaf arrives at its parent
    for (u32 base = blockIdx.x * blockDim.x; base < a.n_nodes; base += nthreads) {
        const u32 i = base + threadIdx.x;
        bool push = false;
        uint4 e = make_uint4(0, 0, 0, 0);
        if (i < a.n_nodes) {
            const u32 prim_count = __float_as_uint(__ldg(reinterpret_cast<const float4*>(a.nodes + i)).w);
            if (prim_count != 0) {
                if (i == 0) {  // a single-leaf tree: the host reads S[0] of the root record
                    const Dec d = leaf_dec(load_node(a.nodes));
                    store_dec(a.dec, d);
                    store_hot(a.hot, d, prim_count);
                } else {
                    push = frontier_arrive(a, i, true, prim_count, e);
                }
            }
        }
        frontier_push(push, e, a.queue[0], &a.qcount[1]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.qcount[2] = 0;
    grid.sync();
    for (u32 round = 1;; round++) {
        const uint4* q = a.queue[(round - 1) & 1];
        uint4* qnext = a.queue[round & 1];
        const u32 n = __ldcg(&a.qcount[round % 3]);
        if (n == 0) break;
        if (n < FRONT_CLIMB_BELOW) {
            for (u32 idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += nthreads) {
                uint4 e = __ldcg(q + idx);
                for (;;) {
                    frontier_node(a, e);
                    if (e.x == 0) break;
                    __threadfence();
                    if (!frontier_arrive(a, e.x, false, e.z, e)) break;  // the other child's thread will do the parent
                }
            }
            break;
        }
        u32* qn_next = &a.qcount[(round + 1) % 3];
        for (u32 base = blockIdx.x * blockDim.x; base < n; base += nthreads) {
            const u32 idx = base + threadIdx.x;
            bool push = false;
            uint4 e = make_uint4(0, 0, 0, 0);
            if (idx < n) {
                e = __ldcg(q + idx);
                frontier_node(a, e);
                if (e.x != 0) push = frontier_arrive(a, e.x, false, e.z, e);
            }
            frontier_push(push, e, qnext, qn_next);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.qcount[(round + 2) % 3] = 0;  // the slot of round + 2; nobody reads or writes it now
        grid.sync();
    }
}

// What a group lane knows about one BVH2 node while get_children runs: everything comes from ONE round trip
// (node, both meta words and S[0] of its decision record, subtree primitive count).
struct Entry {
    u32 node;     // BVH2 node index, INVALID32 = empty
    u32 idx;      // decision index to expand with (meaningful when expand != 0)
    u32 expand;   // 1: still a DISTRIBUTE entry, to be replaced by its two children
    u32 meta_lo, meta_hi, S0, P;
    Node32 n;
};
__device__ __forceinline__ u32 entry_meta(const Entry& e, u32 i) { return ((i < 4 ? e.meta_lo >> (8 * i) : e.meta_hi >> (8 * (i - 4)))) & 0xffu; }
__device__ __forceinline__ void entry_load(Entry& e, const Node32* __restrict__ nodes, const uint4* __restrict__ hot) {
    e.n = load_node(nodes + e.node);
    const uint4 h = __ldcg(hot + e.node);
    e.meta_lo = h.x;
    e.meta_hi = h.y;
    e.S0 = h.z;
    e.P = h.w;
    if (e.n.prim_count != 0) {  // leaves have no stored record (see cwbvh_cost_kernel): all LEAF decisions, nothing below
        e.meta_lo = 0;
        e.meta_hi = 0;
        e.S0 = 0;
        e.P = e.n.prim_count;
    }
}

// K12: convert_to_cwbvh_impl (bvh2_to_cwbvh.rs:75-193) for one wide node, executed by a GROUP OF 8 LANES (four nodes per
// warp). Lane c ends up owning child c: its box, kind, quantised bytes, primitives and work item. The whole routine is a
// chain of dependent memory round trips, so it is organised to make that chain short and identical for the four groups
// of a warp (loops are warp-uniform, every round is one trip):
//   get_children (bvh2_to_cwbvh.rs:346-397) is run breadth-wise: in every round all DISTRIBUTE entries of the group are
//   replaced by their two children at once, positions by an 8-lane prefix sum (keeps the recursion's left-to-right order);
//   order_children (:402-467) is an 8-lane arg-min per round; slot-ordered prefix sums give the allocator offsets.
// `gmask` = lanes of the group, `gl` = lane in group, `sbuf` = 80 bytes of shared memory of the group.
struct GroupOut {
    u32 x;  // CWBVH node index being written
};
__device__ void emit_wide_node(const Node32* __restrict__ nodes, const u32* __restrict__ bvh2_prims, const uint4* __restrict__ hot,
                               bool have, uint4 item, u32* __restrict__ next_queue_idx, uint4* __restrict__ next_queue,
                               u32* queue_counter, uint4* __restrict__ out_nodes, u32* __restrict__ out_prims, int order_children, CwGlobals* g,
                               u32 gmask, int gl, u32* sbuf) {
    const int lane = threadIdx.x & 31, gbase = lane & ~7;
    const u32 x = item.x, child_base = item.z, prim_base = item.w;
    // ---- get_children, breadth-wise ------------------------------------------------------------------------------
    Entry e;
    e.node = INVALID32;
    e.idx = 0;
    e.expand = 0;
    e.meta_lo = e.meta_hi = e.S0 = e.P = 0;
    e.n = Node32{0, 0, 0, 0, 0, 0, 0, 0};
    if (have && gl == 0) {
        e.node = item.y;
        entry_load(e, nodes, hot);
        e.expand = e.n.prim_count == 0 ? 1u : 0u;  // a leaf root is its own single child
    }
    // the wide node's own box (lane 0 holds the BVH2 node) and exponents: lanes 0..2 take one axis each
    Box aabb;
    aabb.minx = __shfl_sync(0xffffffffu, e.n.minx, gbase); aabb.miny = __shfl_sync(0xffffffffu, e.n.miny, gbase);
    aabb.minz = __shfl_sync(0xffffffffu, e.n.minz, gbase); aabb.maxx = __shfl_sync(0xffffffffu, e.n.maxx, gbase);
    aabb.maxy = __shfl_sync(0xffffffffu, e.n.maxy, gbase); aabb.maxz = __shfl_sync(0xffffffffu, e.n.maxz, gbase);
    u32 err = 0;
    // staging for the in-group re-arrangement: 8 entries x 14 words {node, idx | flags, meta_lo, meta_hi, S0, P, node[8]}
    constexpr u32 FRESH = 0x100u;
    for (int round = 0; round < 8; round++) {
        if (!__ballot_sync(0xffffffffu, e.expand)) break;  // warp-uniform: the four groups of the warp stay in step
        u32 dl = 0, dr = 0;
        if (e.expand) {
            u32 m = entry_meta(e, e.idx);
            dl = (m >> 2) & 7u;
            dr = (m >> 5) & 7u;
            if (dl == 7u || dr == 7u) err = 1;  // INVALID decision (non-finite costs): the reference indexes out of bounds
        }
        const u32 sz = e.node == INVALID32 ? 0u : (e.expand ? 2u : 1u);
        u32 off = sz;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, off, o, 8);
            if (gl >= o) off += y;
        }
        const u32 total = __shfl_sync(0xffffffffu, off, gbase + 7);
        off -= sz;
        if (total > 8) err = 1;
        __syncwarp();
        if (total <= 8) {
            if (sz == 2) {  // replaced by its two children, left first (the recursion's order)
                sbuf[off * 14] = e.n.first_index;
                sbuf[off * 14 + 1] = dl | FRESH;
                sbuf[off * 14 + 14] = e.n.first_index + 1;
                sbuf[off * 14 + 15] = dr | FRESH;
            } else if (sz == 1) {  // a settled child only moves
                u32* d = sbuf + off * 14;
                d[0] = e.node; d[1] = e.idx; d[2] = e.meta_lo; d[3] = e.meta_hi; d[4] = e.S0; d[5] = e.P;
                d[6] = __float_as_uint(e.n.minx); d[7] = __float_as_uint(e.n.miny); d[8] = __float_as_uint(e.n.minz); d[9] = e.n.prim_count;
                d[10] = __float_as_uint(e.n.maxx); d[11] = __float_as_uint(e.n.maxy); d[12] = __float_as_uint(e.n.maxz); d[13] = e.n.first_index;
            }
        }
        __syncwarp();
        e.node = INVALID32;
        e.expand = 0;
        if ((u32)gl < total && total <= 8) {
            const u32* d = sbuf + gl * 14;
            e.node = d[0];
            e.idx = d[1] & 0xffu;
            if (d[1] & FRESH) {
                entry_load(e, nodes, hot);  // ONE trip: node, decision record, primitive count
                e.expand = (entry_meta(e, e.idx) & 3u) == KIND_DISTRIBUTE ? 1u : 0u;
            } else {
                e.meta_lo = d[2]; e.meta_hi = d[3]; e.S0 = d[4]; e.P = d[5];
                e.n.minx = __uint_as_float(d[6]); e.n.miny = __uint_as_float(d[7]); e.n.minz = __uint_as_float(d[8]); e.n.prim_count = d[9];
                e.n.maxx = __uint_as_float(d[10]); e.n.maxy = __uint_as_float(d[11]); e.n.maxz = __uint_as_float(d[12]); e.n.first_index = d[13];
            }
        }
        __syncwarp();
    }
    if (__ballot_sync(0xffffffffu, e.expand) & gmask) err = 1;  // did not settle in 8 rounds
    if (__ballot_sync(0xffffffffu, err != 0) & gmask) {
        if (gl == 0 && have) g->error = 1;
        have = false;
    }
    const bool active = have && e.node != INVALID32;
    const u32 child = e.node;
    const Box cb = node_box(e.n);
    const u32 kind = active ? (e.meta_lo & 3u) : 3u;  // decisions[child * 7].kind
    const u32 child_count = __popc(__ballot_sync(0xffffffffu, active) & gmask);
    // node.e (bvh2_to_cwbvh.rs:82-99)
    const float pmin = gl == 0 ? aabb.minx : gl == 1 ? aabb.miny : aabb.minz;
    const float pmax = gl == 0 ? aabb.maxx : gl == 1 ? aabb.maxy : aabb.maxz;
    float rcp_mine = 0.f;
    u32 e_mine = gl < 3 ? obvhs_cwbvh_exponent(smax(pmax - pmin, 1e-20f) * DENOM, &rcp_mine) : 0u;
    const u32 ex = __shfl_sync(0xffffffffu, e_mine, gbase + 0), ey = __shfl_sync(0xffffffffu, e_mine, gbase + 1),
              ez = __shfl_sync(0xffffffffu, e_mine, gbase + 2);
    const float rcpx = __shfl_sync(0xffffffffu, rcp_mine, gbase + 0), rcpy = __shfl_sync(0xffffffffu, rcp_mine, gbase + 1),
                rcpz = __shfl_sync(0xffffffffu, rcp_mine, gbase + 2);
    // ---- order_children (bvh2_to_cwbvh.rs:402-467): greedy, globally cheapest (child, slot) first, ties to the first in
    // (child, slot) scan order, costs that are not < f32::MAX never assigned
    int my_slot = -1;
    if (order_children) {
        float cost[8];
        {
            float cx = (aabb.maxx + aabb.minx) * 0.5f, cy = (aabb.maxy + aabb.miny) * 0.5f, cz = (aabb.maxz + aabb.minz) * 0.5f;
            float vx = (cb.maxx + cb.minx) * 0.5f - cx, vy = (cb.maxy + cb.miny) * 0.5f - cy, vz = (cb.maxz + cb.minz) * 0.5f - cz;
#pragma unroll
            for (int sl = 0; sl < 8; sl++) {
                // direction_lut (bvh2_to_cwbvh.rs:40-50): bit 2 -> -x, bit 1 -> -y, bit 0 -> -z; dot = (x + y) + z
                float dx = (sl & 4) ? -1.0f : 1.0f, dy = (sl & 2) ? -1.0f : 1.0f, dz = (sl & 1) ? -1.0f : 1.0f;
                cost[sl] = (dx * vx + dy * vy) + dz * vz;
            }
        }
        u32 filled = 0;
#pragma unroll 1
        for (int round = 0; round < 8; round++) {
            float best_cost = 3.40282347e+38f;
            int best_s = -1;
            if (active && my_slot < 0) {
#pragma unroll
                for (int sl = 0; sl < 8; sl++) {
                    if (!(filled & (1u << sl)) && cost[sl] < best_cost) {
                        best_cost = cost[sl];
                        best_s = sl;
                    }
                }
            }
            // group arg-min over (cost, lane); lanes without a candidate never win
            float wc = best_cost;
            int wl = best_s >= 0 ? gl : 99, ws = best_s;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                float oc = __shfl_xor_sync(0xffffffffu, wc, o);
                int ol = __shfl_xor_sync(0xffffffffu, wl, o), os = __shfl_xor_sync(0xffffffffu, ws, o);
                bool take = (ol != 99) && (wl == 99 || oc < wc || (oc == wc && ol < wl));
                if (take) {
                    wc = oc;
                    wl = ol;
                    ws = os;
                }
            }
            if (!__ballot_sync(0xffffffffu, wl != 99)) break;  // warp-uniform exit
            if (wl != 99) {
                filled |= 1u << ws;
                if (wl == gl) my_slot = ws;
            }
        }
    } else if (active) {
        my_slot = gl;
    }
    if (__ballot_sync(0xffffffffu, active && my_slot < 0) & gmask) {  // the reference indexes out of bounds (panics) here
        if (gl == 0) g->error = 2;
        have = false;
    }
    // ---- LEAF children: count_primitives (bvh2_to_cwbvh.rs:197-211): DFS, left first; at most 3 primitives, 5 nodes
    u32 pc = 0, prim_ids[3] = {0, 0, 0};
    bool bad = false;
    const bool is_leaf_kind = have && active && kind == KIND_LEAF;
    if (is_leaf_kind) {
        pc = e.P;
        if (pc < 1 || pc > 3) {
            bad = true;
        } else if (e.n.prim_count != 0) {
            prim_ids[0] = bvh2_prims[e.n.first_index];
        } else {
            Node32 l = load_node(nodes + e.n.first_index), r = load_node(nodes + e.n.first_index + 1);
            u32 k = 0;
            if (l.prim_count != 0) {
                prim_ids[k++] = bvh2_prims[l.first_index];
            } else {
                Node32 ll = load_node(nodes + l.first_index), lr = load_node(nodes + l.first_index + 1);
                bad |= ll.prim_count == 0 || lr.prim_count == 0;
                prim_ids[k++] = bvh2_prims[ll.first_index];
                prim_ids[k++] = bvh2_prims[lr.first_index];
            }
            if (!bad && r.prim_count != 0) {
                if (k < 3) prim_ids[k++] = bvh2_prims[r.first_index];
                else bad = true;
            } else if (!bad) {
                Node32 rl = load_node(nodes + r.first_index), rr = load_node(nodes + r.first_index + 1);
                bad |= rl.prim_count == 0 || rr.prim_count == 0 || k != 1;
                if (!bad) {
                    prim_ids[k++] = bvh2_prims[rl.first_index];
                    prim_ids[k++] = bvh2_prims[rr.first_index];
                }
            }
            bad |= k != pc;
        }
    } else if (have && active && kind != KIND_INTERNAL) {
        bad = true;  // DISTRIBUTE is unreachable here in the reference
    }
    if (__ballot_sync(0xffffffffu, bad) & gmask) {
        if (gl == 0) g->error = 1;
        have = false;
    }
    const bool internal = have && active && kind == KIND_INTERNAL;
    // ---- prefix sums in SLOT order over the group (the recursion visits slots 0..7)
    u32 prims_before = 0, int_before = 0, k_before = 0, p_before = 0, num_internal = 0, num_primitives = 0, imask = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int os = __shfl_sync(0xffffffffu, my_slot, gbase + c);
        u32 opc = __shfl_sync(0xffffffffu, pc, gbase + c);
        u32 oint = __shfl_sync(0xffffffffu, internal ? 1u : 0u, gbase + c);
        u32 oS = __shfl_sync(0xffffffffu, e.S0, gbase + c), oP = __shfl_sync(0xffffffffu, e.P, gbase + c);
        if (os < 0) continue;
        num_primitives += opc;
        num_internal += oint;
        if (oint) imask |= 1u << os;
        if (os < my_slot) {
            prims_before += opc;
            int_before += oint;
            if (oint) {
                k_before += oS;
                p_before += oP;
            }
        }
    }
    // ---- the group's 80-byte image of the node in shared memory
    __syncwarp();
    for (int k = gl; k < 20; k += 8) sbuf[k] = 0;
    __syncwarp();
    u8* sb = reinterpret_cast<u8*>(sbuf);
    if (have && gl == 0) {
        sbuf[0] = __float_as_uint(aabb.minx);
        sbuf[1] = __float_as_uint(aabb.miny);
        sbuf[2] = __float_as_uint(aabb.minz);
        sbuf[3] = ex | ey << 8 | ez << 16 | imask << 24;
        sbuf[4] = child_base;
        sbuf[5] = prim_base;
    }
    if (have && active) {
        // bvh2_to_cwbvh.rs:128-141: floor/ceil, clamp 0..255 (glam clamp = max then min, SSE operand rule), `as u8`
        float lo0 = floorf((cb.minx - aabb.minx) * rcpx), lo1 = floorf((cb.miny - aabb.miny) * rcpy), lo2 = floorf((cb.minz - aabb.minz) * rcpz);
        float hi0 = ceilf((cb.maxx - aabb.minx) * rcpx), hi1 = ceilf((cb.maxy - aabb.miny) * rcpy), hi2 = ceilf((cb.maxz - aabb.minz) * rcpz);
        sb[32 + my_slot] = (u8)(u32)smin(smax(lo0, 0.0f), 255.0f);
        sb[40 + my_slot] = (u8)(u32)smin(smax(hi0, 0.0f), 255.0f);
        sb[48 + my_slot] = (u8)(u32)smin(smax(lo1, 0.0f), 255.0f);
        sb[56 + my_slot] = (u8)(u32)smin(smax(hi1, 0.0f), 255.0f);
        sb[64 + my_slot] = (u8)(u32)smin(smax(lo2, 0.0f), 255.0f);
        sb[72 + my_slot] = (u8)(u32)smin(smax(hi2, 0.0f), 255.0f);
        if (internal) {
            sb[24 + my_slot] = (u8)((24u + (u32)my_slot) | 0x20u);
        } else {
            u32 unary = pc == 1 ? 0x20u : pc == 2 ? 0x60u : 0xe0u;
            sb[24 + my_slot] = (u8)((prims_before & 0xffu) | unary);
            for (u32 k = 0; k < pc; k++) out_prims[prim_base + prims_before + k] = prim_ids[k];
        }
    }
    // ---- work items of the INTERNAL children; one queue reservation per group
    u32 qbase = 0;
    if (have && gl == 0 && num_internal) qbase = atomicAdd(queue_counter, num_internal);
    qbase = __shfl_sync(0xffffffffu, qbase, gbase);
    if (internal) {
        // child_base(c_j) = child_base(N) + k(N) + sum over earlier INTERNAL slots of (K - 1) = S[0]; prim_base likewise with P
        next_queue[qbase + int_before] =
            make_uint4(child_base + int_before, child, child_base + num_internal + k_before, prim_base + num_primitives + p_before);
    }
    __syncwarp();
    if (have && gl < 5) {
        const uint4* src = reinterpret_cast<const uint4*>(sbuf);
        out_nodes[(size_t)x * 5 + gl] = src[gl];
    }
    __syncwarp();
    (void)next_queue_idx;
    (void)child_count;
    (void)lane;
}

struct EmitArgs {
    const Node32* nodes;
    const u32* bvh2_prims;
    const uint4* hot;  // {meta_lo, meta_hi, S[0], P} per BVH2 node (store_hot)
    uint4* queue_a;  // work items {cwbvh node index, bvh2 node, child_base, prim_base}
    uint4* queue_b;
    uint4* out_nodes;
    u32* out_prims;
    int order_children;
    u32 expected;  // M = 1 + S(root, 0)
    CwGlobals* g;
    float4* exact;  // CwBvh::exact_node_aabbs (bvh2_to_cwbvh.rs:78-80), or null
};

constexpr int EMIT_THREADS = 512;

// All CWBVH levels in ONE cooperative launch: level L's wide nodes are emitted by 8-lane groups in a grid-stride loop,
// their INTERNAL children are queued for level L+1, a grid-wide barrier separates the levels. No host round trips.
__global__ void __launch_bounds__(EMIT_THREADS) cwbvh_emit_all_kernel(EmitArgs a) {
    cg::grid_group grid = cg::this_grid();
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const u32 warp_id = tid >> 5, nwarps = nthreads >> 5;
    const int gl = threadIdx.x & 7, giw = (threadIdx.x & 31) >> 3;  // lane in group, group in warp
    const u32 gmask = 0xffu << (giw * 8);
    __shared__ __align__(16) u32 s_buf[(EMIT_THREADS / 8) * 112];
    u32* sbuf = s_buf + (threadIdx.x >> 3) * 112;
    CwGlobals* g = a.g;
    if (tid == 0) {
        a.queue_a[0] = make_uint4(0u, 0u, 1u, 0u);  // convert_to_cwbvh_impl(0, 0): nodes = [default] -> child_base = 1
        g->queue_count[0] = g->queue_count[1] = g->queue_count[2] = 0;
        g->error = 0;
        g->emitted = 0;
        g->levels = 0;
    }
    grid.sync();
    if (tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g->level_ns[0] = t;
        g->level_len[0] = 1;
    }
    u32 qlen = 1, emitted = 0;
    uint4 *cur = a.queue_a, *nxt = a.queue_b;
    for (u32 level = 0; qlen > 0; level++) {
        if (tid == 0) g->queue_count[(level + 1) % 3] = 0;  // slot of the next level
        for (u32 t0 = warp_id * 4; t0 < qlen; t0 += nwarps * 4) {  // warp-uniform trip count, four nodes per warp
            const u32 t = t0 + giw;
            const bool have = t < qlen;
            uint4 item = make_uint4(0, 0, 0, 0);
            if (have) item = __ldcg(&cur[t]);
if (a.exact && have && gl == 0) {  // exact_node_aabbs[node_index_bvh8] = *aabb (bvh2_to_cwbvh.rs:78-80)
                const float4* src = reinterpret_cast<const float4*>(a.nodes + item.y);
                const float4 lo = __ldg(src), hi = __ldg(src + 1);
                a.exact[2 * (size_t)item.x] = make_float4(lo.x, lo.y, lo.z, 0.f);
                a.exact[2 * (size_t)item.x + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
            }
                        emit_wide_node(a.nodes, a.bvh2_prims, a.hot, have, item, nullptr, nxt, &g->queue_count[level % 3], a.out_nodes, a.out_prims,
                           a.order_children, g, gmask, gl, sbuf);
        }
        grid.sync();
        emitted += qlen;
        qlen = __ldcg(&g->queue_count[level % 3]);
        if (__ldcg(&g->error) != 0) break;
        if (emitted + qlen > a.expected) {
            if (tid == 0) g->error = 3;
            break;
        }
        uint4* t2 = cur;
        cur = nxt;
        nxt = t2;
        if (tid == 0) {
            g->levels = level + 1;
            if (level < 47) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                g->level_ns[level + 1] = t;
                g->level_len[level + 1] = qlen;
            }
        }
    }
    if (tid == 0) g->emitted = emitted;
}

// ---- CwBvh::order_children as a separate pass (src/cwbvh/mod.rs:520-735) -----------------------------------------------------
// The reference loops `for i in 0..nodes.len() { order_node_children(i) }`. Node i only permutes its own eight slots and moves the
// 80-byte records (and exact boxes) of its inner children inside its own range [child_base_idx, child_base_idx + inner_count); what
// it reads of a child (p, e -- or the exact box -- for the centre) does not change when that child is reordered later, and a
// parent always has a lower index than its children. So the sequential loop equals a top-down sweep level by level: one
// cooperative launch, an 8-lane group per node (lane = old slot), a grid barrier per CWBVH level.
struct OrderArgs {
    uint4* nodes;
    float4* exact;                // CwBvh::exact_node_aabbs or null
    const float4* prim_aabbs;     // the primitives' boxes (Boundable::aabb), two float4 each
    const u32* primitive_indices;
    int direct_layout;
    u32 n_prims;                  // number of boxes in prim_aabbs
    u32* queue[2];
    u32* qcount;                  // [3]
    u32* error;                   // 1: a child could not be assigned (non-finite centres): the reference asserts; 2: a primitive index beyond prim_aabbs
};
constexpr int ORDER_THREADS = 256;
__global__ void __launch_bounds__(ORDER_THREADS) cwbvh_order_children_kernel(OrderArgs a) {
    cg::grid_group grid = cg::this_grid();
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const u32 ngroups = nthreads >> 3, group = tid >> 3;
    const int gl = threadIdx.x & 7, lane = threadIdx.x & 31, gbase = lane & ~7;
    const u32 gmask = 0xffu << gbase;
    __shared__ __align__(16) u32 s_img[(ORDER_THREADS / 8) * 20];
    u32* img = s_img + (threadIdx.x >> 3) * 20;
    if (tid == 0) {
        a.queue[0][0] = 0;
        a.qcount[0] = 1;
        a.qcount[1] = a.qcount[2] = 0;
    }
    grid.sync();
    for (u32 level = 0;; level++) {
        const u32 n = __ldcg(&a.qcount[level % 3]);
        if (n == 0) break;
        const u32* q = a.queue[level & 1];
        u32* qn = a.queue[(level + 1) & 1];
        u32* qn_count = &a.qcount[(level + 1) % 3];
        for (u32 t0 = (group & ~3u); t0 < n; t0 += ngroups) {  // warp-uniform trip count: the four groups of a warp stay in step
            const u32 t = t0 + (group & 3u);
            const bool have = t < n;
            const u32 x = have ? __ldcg(q + t) : 0u;
            const uint4 q0 = __ldcg(a.nodes + (size_t)x * 5), q1 = __ldcg(a.nodes + (size_t)x * 5 + 1);
            const uint4 q2 = __ldcg(a.nodes + (size_t)x * 5 + 2), q3 = __ldcg(a.nodes + (size_t)x * 5 + 3), q4 = __ldcg(a.nodes + (size_t)x * 5 + 4);
            const u32 imask = q0.w >> 24, child_base = q1.x, prim_base = q1.y;
            const u32 meta = ((gl < 4 ? q1.z : q1.w) >> ((gl & 3) * 8)) & 0xffu;
            const bool empty = !have || meta == 0;
            const bool leaf = (imask & (1u << gl)) == 0;  // node.rs:279-281 (true for empty slots too)
            const bool inner = !empty && !leaf;
            // the node's own centre always comes from the compressed box (cwbvh/mod.rs:549), max = p + e * 255 (node.rs:261-265)
            const float px = __uint_as_float(q0.x), py = __uint_as_float(q0.y), pz = __uint_as_float(q0.z);
            const float ex = __uint_as_float((q0.w & 0xffu) << 23), ey = __uint_as_float(((q0.w >> 8) & 0xffu) << 23),
                        ez = __uint_as_float(((q0.w >> 16) & 0xffu) << 23);
            const float cx = ((px + ex * 255.0f) + px) * 0.5f, cy = ((py + ey * 255.0f) + py) * 0.5f, cz = ((pz + ez * 255.0f) + pz) * 0.5f;
            // centre of this lane's child
            u32 old_child_idx = 0;
            float ccx = 0.f, ccy = 0.f, ccz = 0.f;
            if (inner) {
                const u32 slot_index = (meta & 31u) - 24u;  // node.rs:299-304
                old_child_idx = child_base + __popc(imask & ~(0xffffffffu << slot_index));
                float mnx, mny, mnz, mxx, mxy, mxz;
                if (a.exact) {  // node_aabb(): the exact box when the tree carries them (cwbvh/mod.rs:737-745)
                    const float4 lo = __ldcg(a.exact + 2 * (size_t)old_child_idx), hi = __ldcg(a.exact + 2 * (size_t)old_child_idx + 1);
                    mnx = lo.x; mny = lo.y; mnz = lo.z; mxx = hi.x; mxy = hi.y; mxz = hi.z;
                } else {
                    const uint4 c0 = __ldcg(a.nodes + (size_t)old_child_idx * 5);
                    mnx = __uint_as_float(c0.x); mny = __uint_as_float(c0.y); mnz = __uint_as_float(c0.z);
                    mxx = mnx + __uint_as_float((c0.w & 0xffu) << 23) * 255.0f;
                    mxy = mny + __uint_as_float(((c0.w >> 8) & 0xffu) << 23) * 255.0f;
                    mxz = mnz + __uint_as_float(((c0.w >> 16) & 0xffu) << 23) * 255.0f;
                }
                ccx = (mxx + mnx) * 0.5f; ccy = (mxy + mny) * 0.5f; ccz = (mxz + mnz) * 0.5f;
            } else if (!empty) {  // child_primitives (node.rs:290-295): union of the primitives' boxes, starting from Aabb::empty()
                const u32 start = prim_base + (meta & 31u), count = __popc(meta & 0xe0u);
                const float FMAX = 3.40282347e+38f;
                Box b{FMAX, FMAX, FMAX, -FMAX, -FMAX, -FMAX};
                for (u32 i = 0; i < count; i++) {
                    u32 pi = start + i;
                    if (!a.direct_layout) pi = __ldg(a.primitive_indices + pi);
                    if (pi >= a.n_prims) {  // the reference would panic on the slice index
                        *a.error = 2;
                        break;
                    }
                    const float4 lo = __ldg(a.prim_aabbs + 2 * (size_t)pi), hi = __ldg(a.prim_aabbs + 2 * (size_t)pi + 1);
                    b = box_union(b, Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z});
                }
                ccx = (b.maxx + b.minx) * 0.5f; ccy = (b.maxy + b.miny) * 0.5f; ccz = (b.maxz + b.minz) * 0.5f;
            }
            // cost table of this child (cwbvh/mod.rs:584-600) and the greedy assignment (:605-636): globally cheapest (child, slot)
            // first, ties to the first in (child, slot) scan order, costs that are not < f32::MAX never assigned
            float cost[8];
            {
                const float vx = ccx - cx, vy = ccy - cy, vz = ccz - cz;
#pragma unroll
                for (int sl = 0; sl < 8; sl++) {
                    const float dx = (sl & 4) ? -1.0f : 1.0f, dy = (sl & 2) ? -1.0f : 1.0f, dz = (sl & 1) ? -1.0f : 1.0f;
                    cost[sl] = (dx * vx + dy * vy) + dz * vz;
                }
            }
            int my_slot = -1;
            u32 filled = 0;
#pragma unroll 1
            for (int round = 0; round < 8; round++) {
                float best_cost = 3.40282347e+38f;
                int best_s = -1;
                if (!empty && my_slot < 0) {
#pragma unroll
                    for (int sl = 0; sl < 8; sl++) {
                        if (!(filled & (1u << sl)) && cost[sl] < best_cost) {
                            best_cost = cost[sl];
                            best_s = sl;
                        }
                    }
                }
                float wc = best_cost;
                int wl = best_s >= 0 ? gl : 99, ws = best_s;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    const float oc = __shfl_xor_sync(0xffffffffu, wc, o);
                    const int ol = __shfl_xor_sync(0xffffffffu, wl, o), os = __shfl_xor_sync(0xffffffffu, ws, o);
                    const bool take = (ol != 99) && (wl == 99 || oc < wc || (oc == wc && ol < wl));
                    if (take) {
                        wc = oc;
                        wl = ol;
                        ws = os;
                    }
                }
                if (!__ballot_sync(0xffffffffu, wl != 99)) break;  // warp-uniform exit
                if (wl != 99) {
                    filled |= 1u << ws;
                    if (wl == gl) my_slot = ws;
                }
            }
            if (__ballot_sync(0xffffffffu, !empty && my_slot < 0) & gmask) {
                if (gl == 0) *a.error = 1;
            }
            const bool ok = !empty && my_slot >= 0;
            const u32 new_imask = (__ballot_sync(0xffffffffu, ok && inner ? true : false) >> gbase) & 0xffu;  // by OLD lane; rebuilt by slot below
            // imask by NEW slot: OR of (1 << my_slot) over the inner lanes of the group
            u32 im = (ok && inner) ? (1u << my_slot) : 0u;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) im |= __shfl_xor_sync(0xffffffffu, im, o);
            (void)new_imask;
            // ---- the node's new 80-byte image (cwbvh/mod.rs:638-663)
            // `let mut new_node = old_node` (:643): only imask and child_meta are cleared, so a slot left empty keeps the quantised
            // bytes the OLD node had in that slot -- dead data to a traversal, but part of the bit-exact image
            __syncwarp();
            u8* ib = reinterpret_cast<u8*>(img);
            if (have && gl == 0) {
                uint4* iq = reinterpret_cast<uint4*>(img);
                iq[0] = make_uint4(q0.x, q0.y, q0.z, (q0.w & 0x00ffffffu) | (im << 24));
                iq[1] = make_uint4(child_base, prim_base, 0u, 0u);
                iq[2] = q2;
                iq[3] = q3;
                iq[4] = q4;
            }
            __syncwarp();
            if (ok) {
                const int sh = (gl & 3) * 8;
                ib[24 + my_slot] = inner ? (u8)((24u + (u32)my_slot) | 0x20u) : (u8)meta;
                ib[32 + my_slot] = (u8)(((gl < 4 ? q2.x : q2.y) >> sh) & 0xffu);
                ib[40 + my_slot] = (u8)(((gl < 4 ? q2.z : q2.w) >> sh) & 0xffu);
                ib[48 + my_slot] = (u8)(((gl < 4 ? q3.x : q3.y) >> sh) & 0xffu);
                ib[56 + my_slot] = (u8)(((gl < 4 ? q3.z : q3.w) >> sh) & 0xffu);
                ib[64 + my_slot] = (u8)(((gl < 4 ? q4.x : q4.y) >> sh) & 0xffu);
                ib[72 + my_slot] = (u8)(((gl < 4 ? q4.z : q4.w) >> sh) & 0xffu);
            }
            // ---- move the inner children's records (and exact boxes) to the index their new slot implies (:665-733)
            uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0, c2 = c0, c3 = c0, c4 = c0;
            float4 e0 = make_float4(0, 0, 0, 0), e1 = e0;
            const bool mover = ok && inner;
            if (mover) {
                const uint4* src = a.nodes + (size_t)old_child_idx * 5;
                c0 = __ldcg(src); c1 = __ldcg(src + 1); c2 = __ldcg(src + 2); c3 = __ldcg(src + 3); c4 = __ldcg(src + 4);
                if (a.exact) {
                    e0 = __ldcg(a.exact + 2 * (size_t)old_child_idx);
                    e1 = __ldcg(a.exact + 2 * (size_t)old_child_idx + 1);
                }
            }
            __syncwarp();  // every record of the range is in registers before any is overwritten
            u32 new_child_idx = 0;
            if (mover) {
                new_child_idx = child_base + __popc(im & ~(0xffffffffu << my_slot));
                uint4* dst = a.nodes + (size_t)new_child_idx * 5;
                dst[0] = c0; dst[1] = c1; dst[2] = c2; dst[3] = c3; dst[4] = c4;
                if (a.exact) {
                    a.exact[2 * (size_t)new_child_idx] = e0;
                    a.exact[2 * (size_t)new_child_idx + 1] = e1;
                }
            }
            if (have && gl < 5) a.nodes[(size_t)x * 5 + gl] = reinterpret_cast<const uint4*>(img)[gl];
            // ---- next level: this node's inner children (one queue reservation per warp)
            const u32 movers = __ballot_sync(0xffffffffu, mover);
            u32 wbase = 0;
            if (lane == 0 && movers) wbase = atomicAdd(qn_count, (u32)__popc(movers));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (mover) qn[wbase + __popc(movers & ((1u << lane) - 1u))] = new_child_idx;
            __syncwarp();
        }
        if (tid == 0) a.qcount[(level + 2) % 3] = 0;
        grid.sync();
    }
}

// vec![Aabb::empty(); bvh2.nodes.len()] (bvh2_to_cwbvh.rs:60-62; aabb.rs:166-171)
__global__ void fill_empty_aabbs_kernel(float4* out, u32 n) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float FMAX = 3.40282347e+38f;
    out[2 * (size_t)i] = make_float4(FMAX, FMAX, FMAX, 0.f);
    out[2 * (size_t)i + 1] = make_float4(-FMAX, -FMAX, -FMAX, 0.f);
}

__global__ void root_aabb_kernel(const Node32* nodes, float* out8) {
    Node32 r = load_node(nodes);
    out8[0] = r.minx; out8[1] = r.miny; out8[2] = r.minz; out8[3] = 0.f;
    out8[4] = r.maxx; out8[5] = r.maxy; out8[6] = r.maxz; out8[7] = 0.f;
}

}  // namespace

int bvh2_to_cwbvh_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, u32 max_prims_per_leaf, bool order_children, ObvhsCwBvh** out,
                         bool include_exact_node_aabbs) {
    cudaStream_t s = ctx->stream;
    ObvhsCwBvh* cw = new ObvhsCwBvh();
    cw->device = ctx->device;
    cw->owner = ctx;
    cw->uses_spatial_splits = bvh->uses_spatial_splits;  // bvh2_to_cwbvh.rs:508
    obvhs_context_retain(ctx);
    struct Guard {
        ObvhsCwBvh* b;
        ~Guard() { if (b) obvhs_cuda_cwbvh_free(b); }
    } guard{cw};
    if (bvh->node_count == 0) {  // bvh2_to_cwbvh.rs:496-498 CwBvh::default()
        guard.b = nullptr;
        *out = cw;
        return OBVHS_OK;
    }
    const u32 n_nodes = (u32)bvh->node_count;
    DevBuf<u32> parents_tmp, P, arrivals;
    DevBuf<uint4> hot;
    DevBuf<uint4> queue_a, queue_b;
    DevBuf<Dec> dec;
    DevBuf<CwGlobals> g;
    DevBuf<float> root_box;
    const u32* parents = bvh->parents;
    if (!parents) {  // the converter does not need Bvh2::parents; the bottom-up pass does -> scratch copy
        CU_TRY(ctx, parents_tmp.alloc(n_nodes, s));
        ST_TRY(bvh2_compute_parents_into(ctx, bvh, parents_tmp.p));
        parents = parents_tmp.p;
    }
    std::optional<TraceScope> tsp;
    tsp.emplace(ctx, "  calculate_cost");
    CU_TRY(ctx, hot.alloc(n_nodes, s));
    CU_TRY(ctx, arrivals.alloc(n_nodes, s));
    CU_TRY(ctx, dec.alloc(n_nodes, s));
    CU_TRY(ctx, g.alloc(1, s));
    CU_TRY(ctx, root_box.alloc(8, s));
    CU_TRY(ctx, cudaMemsetAsync(arrivals.p, 0, (size_t)n_nodes * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(g.p, 0, sizeof(CwGlobals), s));
    // small trees: the frontier kernel would switch to climbing after its first round anyway, and a plain launch is cheaper
    // than a cooperative one (kitchen, 114 k nodes: 0.11 vs 0.14 ms)
    if (n_nodes < 8 * FRONT_CLIMB_BELOW) {
        CU_TRY(ctx, P.alloc(n_nodes, s));
        cwbvh_cost_kernel<<<div_up(n_nodes, COST_THREADS), COST_THREADS, 0, s>>>(bvh->nodes, parents, n_nodes, max_prims_per_leaf, dec.p, hot.p, P.p, arrivals.p);
        KERNEL_CHECK(ctx);
    } else {
        if (bvh->prim_count >= ((size_t)1 << 28)) {  // frontier_arrive keeps primitive counts in 28 bits of the arrival word
            OBVHS_SET_ERR(ctx, "bvh2_to_cwbvh: more than 2^28 primitives are not supported");
            return OBVHS_ERR_UNSUPPORTED;
        }
        DevBuf<uint4> q0, q1;
        DevBuf<u32> qcount;
        CU_TRY(ctx, q0.alloc((size_t)n_nodes / 2 + 1, s));
        CU_TRY(ctx, q1.alloc((size_t)n_nodes / 2 + 1, s));
        CU_TRY(ctx, qcount.alloc(4, s));
        CU_TRY(ctx, cudaMemsetAsync(qcount.p, 0, 16, s));
        CostArgs ca;
        ca.nodes = bvh->nodes; ca.parents = parents; ca.n_nodes = n_nodes; ca.max_prims_per_leaf = max_prims_per_leaf;
        ca.dec = dec.p; ca.hot = hot.p; ca.arrivals = arrivals.p; ca.queue[0] = q0.p; ca.queue[1] = q1.p; ca.qcount = qcount.p;
        int per_sm = 0;
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cwbvh_cost_frontier_kernel, FRONT_THREADS, 0));
        const int blocks = std::min(std::max(1, per_sm) * ctx->sm_count, std::max(1, div_up(n_nodes, FRONT_THREADS)));
        void* args[] = {&ca};
        CU_TRY(ctx, cudaLaunchCooperativeKernel((void*)cwbvh_cost_frontier_kernel, dim3(blocks), dim3(FRONT_THREADS), args, 0, s));
        KERNEL_CHECK(ctx);
    }
    root_aabb_kernel<<<1, 1, 0, s>>>(bvh->nodes, root_box.p);
    KERNEL_CHECK(ctx);
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, &dec.p->S[0], 4, cudaMemcpyDeviceToHost, s));  // M = 1 + S(root, 0)
    CU_TRY(ctx, cudaMemcpyAsync(h + 8, root_box.p, 32, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    tsp.reset();
    TraceScope ts_emit(ctx, "  convert_to_cwbvh");
    const u32 M = h[0] + 1;
    memcpy(&cw->total_aabb, h + 8, 32);  // bvh2_to_cwbvh.rs:506 total_aabb = bvh2.nodes[0].aabb
    if (M == 0 || M > n_nodes) {
        OBVHS_SET_ERR(ctx, "bvh2_to_cwbvh: invalid wide node count %u", M);
        return OBVHS_ERR_CUDA;
    }
    cw->node_count = M;
    cw->prim_count = bvh->prim_count;
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&cw->nodes, (size_t)M * sizeof(ObvhsCwBvhNode)));
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&cw->primitive_indices, std::max<size_t>(1, cw->prim_count) * 4));
    CU_TRY(ctx, queue_a.alloc(M, s));
    CU_TRY(ctx, queue_b.alloc(M, s));
    if (include_exact_node_aabbs) {
        cw->exact_count = n_nodes;
        CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&cw->exact_node_aabbs, (size_t)n_nodes * sizeof(ObvhsAabb)));
        fill_empty_aabbs_kernel<<<div_up(n_nodes, 256), 256, 0, s>>>(reinterpret_cast<float4*>(cw->exact_node_aabbs), n_nodes);
        KERNEL_CHECK(ctx);
    }
    {
        EmitArgs ea;
        ea.nodes = bvh->nodes; ea.bvh2_prims = bvh->primitive_indices; ea.hot = hot.p;
        ea.queue_a = queue_a.p; ea.queue_b = queue_b.p; ea.out_nodes = reinterpret_cast<uint4*>(cw->nodes);
        ea.out_prims = cw->primitive_indices; ea.order_children = order_children ? 1 : 0; ea.expected = M; ea.g = g.p;
        ea.exact = reinterpret_cast<float4*>(cw->exact_node_aabbs);
        int per_sm = 0;
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cwbvh_emit_all_kernel, EMIT_THREADS, 0));
        // few, fat blocks: the cost of a grid-wide barrier grows with the number of participating blocks
        int blocks = std::min(std::max(1, per_sm) * ctx->sm_count, std::max(1, div_up((size_t)M * 8, EMIT_THREADS)));
        void* args[] = {&ea};
        CU_TRY(ctx, cudaLaunchCooperativeKernel((void*)cwbvh_emit_all_kernel, dim3(blocks), dim3(EMIT_THREADS), args, 0, s));
        KERNEL_CHECK(ctx);
    }
    CU_TRY(ctx, cudaMemcpyAsync(h, g.p, sizeof(CwGlobals), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    if (ctx->trace) {
        const CwGlobals* hg = reinterpret_cast<const CwGlobals*>(h);
        for (u32 l = 0; l + 1 <= hg->levels && l < 46; l++)
            fprintf(stderr, "[obvhs trace]     emit level %2u: %8u nodes %9.1f us\n", l, hg->level_len[l], (hg->level_ns[l + 1] - hg->level_ns[l]) * 1e-3);
    }
    if (h[0] != 0 || h[4] != M) {
        OBVHS_SET_ERR(ctx, "bvh2_to_cwbvh: %s (emitted %u of %u nodes; non-finite AABBs, or Bvh2 leaves holding more than one primitive -- bvh2_to_cwbvh.rs:201 wants the uncollapsed tree? the reference panics here)",
                      h[0] == 2 ? "order_children left a child unassigned" : h[0] == 3 ? "node count mismatch" : "invalid decision on the emit path",
                      h[4], M);
        return OBVHS_ERR_NAN_INPUT;
    }
    guard.b = nullptr;
    *out = cw;
    return OBVHS_OK;
}

// CwBvh::order_children(&mut self, primitives, direct_layout) (src/cwbvh/mod.rs:520-524) with the primitives given as their AABBs
int cwbvh_order_children_device(ObvhsContext* ctx, ObvhsCwBvh* bvh, const ObvhsAabb* d_prim_aabbs, size_t n_prims, bool direct_layout) {
    if (bvh->node_count == 0) return OBVHS_OK;
    cudaStream_t s = ctx->stream;
    DevBuf<u32> q0, q1, qcount;
    CU_TRY(ctx, q0.alloc(bvh->node_count, s));
    CU_TRY(ctx, q1.alloc(bvh->node_count, s));
    CU_TRY(ctx, qcount.alloc(4, s));
    CU_TRY(ctx, cudaMemsetAsync(qcount.p, 0, 16, s));
    OrderArgs a;
    a.nodes = reinterpret_cast<uint4*>(bvh->nodes);
    a.exact = reinterpret_cast<float4*>(bvh->exact_node_aabbs);
    a.prim_aabbs = reinterpret_cast<const float4*>(d_prim_aabbs);
    a.primitive_indices = bvh->primitive_indices;
    a.direct_layout = direct_layout ? 1 : 0;
    a.n_prims = (u32)std::min<size_t>(n_prims, 0xffffffffu);
    a.queue[0] = q0.p;
    a.queue[1] = q1.p;
    a.qcount = qcount.p;
    a.error = qcount.p + 3;
    static PerDevice<int> per_sm_dev;
    int& per_sm = per_sm_dev[ctx->device];
    if (per_sm == 0) {
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cwbvh_order_children_kernel, ORDER_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
    }
    const int blocks = std::min(per_sm * ctx->sm_count, std::max(1, div_up(bvh->node_count * 8, ORDER_THREADS)));
    void* args[] = {&a};
    CU_TRY(ctx, cudaLaunchCooperativeKernel((void*)cwbvh_order_children_kernel, dim3(blocks), dim3(ORDER_THREADS), args, 0, s));
    KERNEL_CHECK(ctx);
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, qcount.p + 3, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    if (h[0] == 2) {
        OBVHS_SET_ERR(ctx, "order_children: a leaf refers to a primitive beyond the %zu boxes given (wrong direct_layout?); the tree may be partly reordered", n_prims);
        return OBVHS_ERR_INVALID_ARG;
    }
    if (h[0] != 0) {
        OBVHS_SET_ERR(ctx, "order_children: a child could not be assigned to a slot (non-finite boxes? the reference asserts here)");
        return OBVHS_ERR_NAN_INPUT;
    }
    return OBVHS_OK;
}
