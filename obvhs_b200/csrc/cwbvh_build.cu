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
struct Dec {
    float cost[7];
    u32 meta_lo, meta_hi;
};
static_assert(sizeof(Dec) == 36, "Dec");

__device__ __forceinline__ u32 dec_meta(const Dec* __restrict__ dec, u32 node, u32 i) {
    const u8* m = reinterpret_cast<const u8*>(&dec[node].meta_lo);
    return __ldcg(m + i);
}

struct CwGlobals {
    u32 error;           // 1: DISTRIBUTE/INVALID decision on the emit path (non-finite costs), 2: child left unassigned, 3: count mismatch
    u32 queue_count[3];  // wide nodes queued by level L for level L+1, in slot L % 3
    u32 emitted;         // wide nodes written
    u32 levels;
};

// get_children (bvh2_to_cwbvh.rs:346-397), iterative. Returns child_count; children in the recursion's order.
__device__ __forceinline__ u32 get_children(const Node32* __restrict__ nodes, const Dec* __restrict__ dec, u32 node_index, u32 children[8],
                                            bool coherent, u32* err) {
    u32 child_count = 0;
    u32 n0_prim, n0_first;
    {
        Node32 nd = coherent ? load_node_cg(nodes + node_index) : load_node(nodes + node_index);
        n0_prim = nd.prim_count;
        n0_first = nd.first_index;
    }
    if (n0_prim != 0) {
        children[0] = node_index;
        return 1;
    }
    // stack entries: node (bits 0..27 are not enough for 2^30 nodes -> two arrays), decision index i, or a direct child
    u32 st_node[8];
    u8 st_i[8];  // 0..6 = expand with decision i; 0xff = emit as child
    int sp = 0;
    st_node[0] = node_index;
    st_i[0] = 0;
    sp = 1;
    u32 first_of_root = n0_first;
    while (sp > 0) {
        sp--;
        u32 node = st_node[sp];
        u32 i = st_i[sp];
        if (i == 0xff) {
            if (child_count < 8) children[child_count] = node;
            child_count++;
            continue;
        }
        u32 first = (node == node_index) ? first_of_root : __ldcg(&nodes[node].first_index);
        u32 m = dec_meta(dec, node, i);
        u32 dl = (m >> 2) & 7u, dr = (m >> 5) & 7u;
        if (dl == 7u || dr == 7u) {
            *err = 1;
            return 0;
        }
        bool left_dist = (dec_meta(dec, first, dl) & 3u) == KIND_DISTRIBUTE;
        bool right_dist = (dec_meta(dec, first + 1, dr) & 3u) == KIND_DISTRIBUTE;
        if (sp + 2 > 8) {
            *err = 1;
            return 0;
        }
        // right is handled after everything the left expands to
        st_node[sp] = first + 1;
        st_i[sp] = right_dist ? (u8)dr : (u8)0xff;
        sp++;
        st_node[sp] = first;
        st_i[sp] = left_dist ? (u8)dl : (u8)0xff;
        sp++;
    }
    if (child_count > 8) {
        *err = 1;
        return 0;
    }
    return child_count;
}

// K11: calculate_cost_impl (bvh2_to_cwbvh.rs:220-344), bottom-up.
__global__ void __launch_bounds__(256) cwbvh_cost_kernel(const Node32* __restrict__ nodes, const u32* __restrict__ parents, u32 n_nodes,
                                                         u32 max_prims_per_leaf, Dec* dec, u32* P, u32* K, u32* arrivals, CwGlobals* g) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    Node32 nd = load_node(nodes + i);
    if (nd.prim_count == 0) return;  // start at leaves
    {
        float ha = box_half_area(node_box(nd));
        float cost_leaf = ha * (float)nd.prim_count * PRIM_COST;
        Dec d;
#pragma unroll
        for (int k = 0; k < 7; k++) d.cost[k] = cost_leaf;
        d.meta_lo = 0;  // kind LEAF, left/right 0 (Decision::default indices)
        d.meta_hi = 0;
        dec[i] = d;
        P[i] = nd.prim_count;
        K[i] = 1;
    }
    if (i == 0) return;
    u32 node = parents[i];
    for (;;) {
        __threadfence();
        if (atomicAdd(&arrivals[node], 1u) == 0) return;
        Node32 me = load_node_cg(nodes + node);
        const u32 first = me.first_index;
        float ha = box_half_area(node_box(me));
        float lc[7], rc[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            lc[k] = __ldcg(&dec[first].cost[k]);
            rc[k] = __ldcg(&dec[first + 1].cost[k]);
        }
        u32 num_primitives = __ldcg(&P[first]) + __ldcg(&P[first + 1]);
        Dec d;
        u8 meta[8];
        {  // i = 0
            float cost_leaf = num_primitives <= max_prims_per_leaf ? ((float)num_primitives * ha) * PRIM_COST : __int_as_float(0x7f800000);
            float cost_distribute = __int_as_float(0x7f800000);
            u32 dl = 7, dr = 7;
#pragma unroll
            for (int k = 0; k < 7; k++) {
                float c = lc[k] + rc[6 - k];
                if (c < cost_distribute) {
                    cost_distribute = c;
                    dl = k;
                    dr = 6 - k;
                }
            }
            float cost_internal = cost_distribute + ha;
            if (cost_leaf < cost_internal) {
                d.cost[0] = cost_leaf;
                meta[0] = (u8)(KIND_LEAF | dl << 2 | dr << 5);
            } else {
                d.cost[0] = cost_internal;
                meta[0] = (u8)(KIND_INTERNAL | dl << 2 | dr << 5);
            }
        }
#pragma unroll
        for (int ii = 1; ii < 7; ii++) {
            float cost_distribute = d.cost[ii - 1];
            u32 dl = 7, dr = 7;
#pragma unroll
            for (int k = 0; k < ii; k++) {
                float c = lc[k] + rc[ii - k - 1];
                if (c < cost_distribute) {
                    cost_distribute = c;
                    dl = k;
                    dr = ii - k - 1;
                }
            }
            d.cost[ii] = cost_distribute;
            if (dl != 7) meta[ii] = (u8)(KIND_DISTRIBUTE | dl << 2 | dr << 5);
            else meta[ii] = meta[ii - 1];  // decisions[node_i] = decisions[node_i - 1]
        }
        meta[7] = 0;
        d.meta_lo = meta[0] | meta[1] << 8 | meta[2] << 16 | (u32)meta[3] << 24;
        d.meta_hi = meta[4] | meta[5] << 8 | meta[6] << 16;
        dec[node] = d;
        P[node] = num_primitives;
        __threadfence();  // get_children below reads this node's own decisions through L2
        // K(node): CWBVH nodes produced by this subtree when `node` is emitted as a wide node
        u32 children[8];
        u32 err = 0;
        u32 cc = get_children(nodes, dec, node, children, true, &err);
        u32 k_total = 1;
        for (u32 c = 0; c < cc && c < 8; c++) {
            u32 ch = children[c];
            if ((dec_meta(dec, ch, 0) & 3u) == KIND_INTERNAL) k_total += __ldcg(&K[ch]);
        }
        // an INVALID decision only matters if this node is really emitted; the emit kernel reports it
        K[node] = k_total;
        if (node == 0) return;
        node = parents[node];
    }
}

struct WorkItem {
    u32 bvh2_node, child_base, prim_base;
};

// K12: convert_to_cwbvh_impl (bvh2_to_cwbvh.rs:75-193) for one wide node (CWBVH index x).
__device__ void emit_wide_node(const Node32* __restrict__ nodes, const u32* __restrict__ bvh2_prims, const Dec* __restrict__ dec,
                               const u32* __restrict__ P, const u32* __restrict__ K, u32 x, u32* __restrict__ next_queue, u32* queue_counter,
                               WorkItem* work, uint4* __restrict__ out_nodes, u32* __restrict__ out_prims, int order_children, CwGlobals* g) {
    WorkItem w;
    w.bvh2_node = __ldcg(&work[x].bvh2_node);
    w.child_base = __ldcg(&work[x].child_base);
    w.prim_base = __ldcg(&work[x].prim_base);
    const Node32 me = load_node(nodes + w.bvh2_node);
    const Box aabb = node_box(me);
    // node.p, node.e (bvh2_to_cwbvh.rs:82-99)
    float px = aabb.minx, py = aabb.miny, pz = aabb.minz;
    float rcpx, rcpy, rcpz;
    u32 ex = obvhs_cwbvh_exponent(smax(aabb.maxx - aabb.minx, 1e-20f) * DENOM, &rcpx);
    u32 ey = obvhs_cwbvh_exponent(smax(aabb.maxy - aabb.miny, 1e-20f) * DENOM, &rcpy);
    u32 ez = obvhs_cwbvh_exponent(smax(aabb.maxz - aabb.minz, 1e-20f) * DENOM, &rcpz);
    u32 children[8];
    u32 err = 0;
    u32 child_count = get_children(nodes, dec, w.bvh2_node, children, false, &err);
    if (err) {
        g->error = 1;
        return;
    }
    for (u32 c = child_count; c < 8; c++) children[c] = INVALID32;
    Box cbox[8];  // boxes of the children, indexed like children[] (get_children order)
    for (u32 c = 0; c < child_count; c++) cbox[c] = node_box(load_node(nodes + children[c]));
    // slot_of[c]: slot of child c. order_children (bvh2_to_cwbvh.rs:402-467), greedy assignment
    int slot_child[8];  // slot -> child position in children[] or -1
#pragma unroll
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    if (order_children) {
        float cx = (aabb.maxx + aabb.minx) * 0.5f, cy = (aabb.maxy + aabb.miny) * 0.5f, cz = (aabb.maxz + aabb.minz) * 0.5f;
        float vx[8], vy[8], vz[8];
        for (u32 c = 0; c < child_count; c++) {
            vx[c] = (cbox[c].maxx + cbox[c].minx) * 0.5f - cx;
            vy[c] = (cbox[c].maxy + cbox[c].miny) * 0.5f - cy;
            vz[c] = (cbox[c].maxz + cbox[c].minz) * 0.5f - cz;
        }
        u32 assigned = 0, filled = 0;  // bit masks over children / slots
        for (;;) {
            float min_cost = 3.40282347e+38f;
            int min_slot = -1, min_index = -1;
            for (u32 c = 0; c < child_count; c++) {
                if (assigned & (1u << c)) continue;
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    // direction_lut (bvh2_to_cwbvh.rs:40-50): bit 2 -> -x, bit 1 -> -y, bit 0 -> -z; dot = (x + y) + z
                    float dx = (s & 4) ? -1.0f : 1.0f, dy = (s & 2) ? -1.0f : 1.0f, dz = (s & 1) ? -1.0f : 1.0f;
                    float cost = (dx * vx[c] + dy * vy[c]) + dz * vz[c];
                    if (!(filled & (1u << s)) && cost < min_cost) {
                        min_cost = cost;
                        min_slot = s;
                        min_index = (int)c;
                    }
                }
            }
            if (min_slot < 0) break;
            filled |= 1u << min_slot;
            assigned |= 1u << min_index;
            slot_child[min_slot] = min_index;
        }
        if (assigned != ((1u << child_count) - 1u)) {  // the reference indexes out of bounds (panics) here
            g->error = 2;
            return;
        }
    } else {
        for (u32 c = 0; c < child_count; c++) slot_child[c] = (int)c;
    }
    // pass 1 over the slots: kinds, direct primitives, internal child bookkeeping
    u32 imask = 0, num_internal = 0, num_primitives = 0;
    u32 meta_b[8], qlo[3][8], qhi[3][8];
    u32 kind_of[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        meta_b[s] = 0;
        qlo[0][s] = qlo[1][s] = qlo[2][s] = 0;
        qhi[0][s] = qhi[1][s] = qhi[2][s] = 0;
        kind_of[s] = 3;
        int c = slot_child[s];
        if (c < 0) continue;
        const Box cb = cbox[c];
        // bvh2_to_cwbvh.rs:128-141: floor/ceil, clamp 0..255 (glam clamp = max then min, SSE operand rule), `as u8`
        float lo[3] = {floorf((cb.minx - px) * rcpx), floorf((cb.miny - py) * rcpy), floorf((cb.minz - pz) * rcpz)};
        float hi[3] = {ceilf((cb.maxx - px) * rcpx), ceilf((cb.maxy - py) * rcpy), ceilf((cb.maxz - pz) * rcpz)};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float l = smin(smax(lo[a], 0.0f), 255.0f), h = smin(smax(hi[a], 0.0f), 255.0f);
            qlo[a][s] = (u32)l;  // values are integers in [0,255] (NaN -> 0 by smax)
            qhi[a][s] = (u32)h;
        }
        u32 child = children[c];
        u32 kind = dec_meta(dec, child, 0) & 3u;
        kind_of[s] = kind;
        if (kind == KIND_LEAF) {
            // count_primitives (bvh2_to_cwbvh.rs:197-211): DFS, left first, pushes primitive ids
            u32 pc = 0;
            u32 stk[8];
            int sp = 0;
            stk[sp++] = child;
            while (sp > 0) {
                u32 nidx = stk[--sp];
                Node32 nn = load_node(nodes + nidx);
                if (nn.prim_count != 0) {
                    if (pc < 3) out_prims[w.prim_base + num_primitives + pc] = bvh2_prims[nn.first_index];
                    pc += nn.prim_count;
                } else if (sp + 2 <= 8) {
                    stk[sp++] = nn.first_index + 1;
                    stk[sp++] = nn.first_index;
                } else {
                    pc = 99;
                    break;
                }
            }
            u32 unary = pc == 1 ? 0x20u : pc == 2 ? 0x60u : pc == 3 ? 0xe0u : 0u;
            if (!unary) {
                g->error = 1;
                return;
            }
            meta_b[s] = (num_primitives & 0xffu) | unary;
            num_primitives += pc;
        } else if (kind == KIND_INTERNAL) {
            imask |= 1u << s;
            meta_b[s] = (24u + (u32)s) | 0x20u;
            num_internal++;
        } else {
            g->error = 1;
            return;
        }
    }
    // pass 2: work items of the internal children
    if (num_internal) {
        u32 qbase = atomicAdd(queue_counter, num_internal);
        u32 j = 0, k_before = 0, p_before = 0;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if (kind_of[s] != KIND_INTERNAL) continue;
            u32 child = children[slot_child[s]];
            WorkItem wi;
            wi.bvh2_node = child;
            wi.child_base = w.child_base + num_internal + k_before;
            wi.prim_base = w.prim_base + num_primitives + p_before;
            u32 ci = w.child_base + j;
            work[ci] = wi;
            next_queue[qbase + j] = ci;
            k_before += K[child] - 1;
            p_before += P[child];
            j++;
        }
    }
    // the 80 bytes (cwbvh/node.rs:14-54)
    auto pack4 = [](const u32* b) { return b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24; };
    uint4 q0 = make_uint4(__float_as_uint(px), __float_as_uint(py), __float_as_uint(pz), ex | ey << 8 | ez << 16 | imask << 24);
    uint4 q1 = make_uint4(w.child_base, w.prim_base, pack4(meta_b), pack4(meta_b + 4));
    uint4 q2 = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    uint4 q3 = make_uint4(pack4(qlo[1]), pack4(qlo[1] + 4), pack4(qhi[1]), pack4(qhi[1] + 4));
    uint4 q4 = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    uint4* o = out_nodes + (size_t)x * 5;
    o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; o[4] = q4;
}

struct EmitArgs {
    const Node32* nodes;
    const u32* bvh2_prims;
    const Dec* dec;
    const u32* P;
    const u32* K;
    u32* queue_a;
    u32* queue_b;
    WorkItem* work;
    uint4* out_nodes;
    u32* out_prims;
    int order_children;
    u32 expected;  // M = K[root]
    CwGlobals* g;
};

// All CWBVH levels in ONE cooperative launch: level L's wide nodes are emitted by a grid-stride loop, their INTERNAL
// children are queued for level L+1, a grid-wide barrier separates the levels. No host round trips.
__global__ void __launch_bounds__(128) cwbvh_emit_all_kernel(EmitArgs a) {
    cg::grid_group grid = cg::this_grid();
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    CwGlobals* g = a.g;
    if (tid == 0) {
        a.work[0] = WorkItem{0u, 1u, 0u};  // convert_to_cwbvh_impl(0, 0): nodes = [default] -> child_base = 1
        a.queue_a[0] = 0;
        g->queue_count[0] = g->queue_count[1] = g->queue_count[2] = 0;
        g->error = 0;
        g->emitted = 0;
        g->levels = 0;
    }
    grid.sync();
    u32 qlen = 1, emitted = 0;
    u32 *cur = a.queue_a, *nxt = a.queue_b;
    for (u32 level = 0; qlen > 0; level++) {
        if (tid == 0) g->queue_count[(level + 1) % 3] = 0;  // slot of the next level
        for (u32 t = tid; t < qlen; t += nthreads)
            emit_wide_node(a.nodes, a.bvh2_prims, a.dec, a.P, a.K, __ldcg(&cur[t]), nxt, &g->queue_count[level % 3], a.work, a.out_nodes, a.out_prims,
                           a.order_children, g);
        grid.sync();
        emitted += qlen;
        qlen = __ldcg(&g->queue_count[level % 3]);
        if (__ldcg(&g->error) != 0) break;
        if (emitted + qlen > a.expected) {
            if (tid == 0) g->error = 3;
            break;
        }
        u32* t2 = cur;
        cur = nxt;
        nxt = t2;
        if (tid == 0) g->levels = level + 1;
    }
    if (tid == 0) g->emitted = emitted;
}

__global__ void root_aabb_kernel(const Node32* nodes, float* out8) {
    Node32 r = load_node(nodes);
    out8[0] = r.minx; out8[1] = r.miny; out8[2] = r.minz; out8[3] = 0.f;
    out8[4] = r.maxx; out8[5] = r.maxy; out8[6] = r.maxz; out8[7] = 0.f;
}

}  // namespace

int bvh2_to_cwbvh_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, u32 max_prims_per_leaf, bool order_children, ObvhsCwBvh** out) {
    cudaStream_t s = ctx->stream;
    ObvhsCwBvh* cw = new ObvhsCwBvh();
    cw->device = ctx->device;
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
    DevBuf<u32> parents_tmp, P, K, arrivals, queue_a, queue_b;
    DevBuf<Dec> dec;
    DevBuf<CwGlobals> g;
    DevBuf<WorkItem> work;
    DevBuf<float> root_box;
    const u32* parents = bvh->parents;
    if (!parents) {  // the converter does not need Bvh2::parents; the bottom-up pass does -> scratch copy
        CU_TRY(ctx, parents_tmp.alloc(n_nodes, s));
        ST_TRY(bvh2_compute_parents_into(ctx, bvh, parents_tmp.p));
        parents = parents_tmp.p;
    }
    TraceScope* tsp = new TraceScope(ctx, "  cwbvh_cost");
    CU_TRY(ctx, P.alloc(n_nodes, s));
    CU_TRY(ctx, K.alloc(n_nodes, s));
    CU_TRY(ctx, arrivals.alloc(n_nodes, s));
    CU_TRY(ctx, dec.alloc(n_nodes, s));
    CU_TRY(ctx, g.alloc(1, s));
    CU_TRY(ctx, root_box.alloc(8, s));
    CU_TRY(ctx, cudaMemsetAsync(arrivals.p, 0, (size_t)n_nodes * 4, s));
    CU_TRY(ctx, cudaMemsetAsync(g.p, 0, sizeof(CwGlobals), s));
    cwbvh_cost_kernel<<<div_up(n_nodes, 256), 256, 0, s>>>(bvh->nodes, parents, n_nodes, max_prims_per_leaf, dec.p, P.p, K.p, arrivals.p, g.p);
    KERNEL_CHECK(ctx);
    root_aabb_kernel<<<1, 1, 0, s>>>(bvh->nodes, root_box.p);
    KERNEL_CHECK(ctx);
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, K.p, 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaMemcpyAsync(h + 8, root_box.p, 32, cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    delete tsp;
    TraceScope ts_emit(ctx, "  cwbvh_emit");
    const u32 M = h[0];
    memcpy(&cw->total_aabb, h + 8, 32);  // bvh2_to_cwbvh.rs:506 total_aabb = bvh2.nodes[0].aabb
    if (M == 0 || M > n_nodes) {
        OBVHS_SET_ERR(ctx, "bvh2_to_cwbvh: invalid wide node count %u", M);
        return OBVHS_ERR_CUDA;
    }
    cw->node_count = M;
    cw->prim_count = bvh->prim_count;
    CU_TRY(ctx, cudaMallocAsync((void**)&cw->nodes, (size_t)M * sizeof(ObvhsCwBvhNode), s));
    CU_TRY(ctx, cudaMallocAsync((void**)&cw->primitive_indices, std::max<size_t>(1, cw->prim_count) * 4, s));
    CU_TRY(ctx, work.alloc(M, s));
    CU_TRY(ctx, queue_a.alloc(M, s));
    CU_TRY(ctx, queue_b.alloc(M, s));
    {
        EmitArgs ea;
        ea.nodes = bvh->nodes; ea.bvh2_prims = bvh->primitive_indices; ea.dec = dec.p; ea.P = P.p; ea.K = K.p;
        ea.queue_a = queue_a.p; ea.queue_b = queue_b.p; ea.work = work.p; ea.out_nodes = reinterpret_cast<uint4*>(cw->nodes);
        ea.out_prims = cw->primitive_indices; ea.order_children = order_children ? 1 : 0; ea.expected = M; ea.g = g.p;
        int per_sm = 0;
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cwbvh_emit_all_kernel, 128, 0));
        int blocks = std::min(std::max(1, per_sm) * ctx->sm_count, std::max(1, div_up(M, 128)));
        void* args[] = {&ea};
        CU_TRY(ctx, cudaLaunchCooperativeKernel((void*)cwbvh_emit_all_kernel, dim3(blocks), dim3(128), args, 0, s));
        KERNEL_CHECK(ctx);
    }
    CU_TRY(ctx, cudaMemcpyAsync(h, g.p, sizeof(CwGlobals), cudaMemcpyDeviceToHost, s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    if (h[0] != 0 || h[4] != M) {
        OBVHS_SET_ERR(ctx, "bvh2_to_cwbvh: %s (emitted %u of %u nodes; non-finite AABBs? the reference panics here)",
                      h[0] == 2 ? "order_children left a child unassigned" : h[0] == 3 ? "node count mismatch" : "invalid decision on the emit path",
                      h[4], M);
        return OBVHS_ERR_NAN_INPUT;
    }
    guard.b = nullptr;
    *out = cw;
    return OBVHS_OK;
}
