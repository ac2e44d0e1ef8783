// ploc.cu -- PLOC BVH2 build on one B200 (sm_100a).
//
// Replaces PlocBuilder::build / build_with_bvh / build_ploc / build_ploc_from_leaves  (src/ploc/mod.rs:95-503),
// sort_nodes_by_morton (:771-847) and morton_encode_u64_unorm (src/ploc/morton.rs:35-58):
//   K1  leaf_init_kernel      leaf nodes + scene AABB (block reduce -> 6 ordered-int atomics)      ploc/mod.rs:187-243
//   K2  morton_kernel         f64 normalise, 21 bits/axis, bit interleave                         ploc/mod.rs:287-288,782-785
//   K3  onesweep radix sort   (sort.cu), stable: ties by ascending original index                  ploc/mod.rs:811-827
//   K4  gather_nodes_kernel   nodes into sorted order                                             ploc/mod.rs:829-846
//   K5+K6 ploc_fused_kernel   one pass per iteration: the tile's window of sorted cluster AABBs staged in smem by a TMA bulk
//                             copy, radius-r nearest neighbour (ploc/mod.rs:329-419,575-651), then the reference's SEQUENTIAL
//                             merge sweep restated as flags + a decoupled-look-back scan (kept/parent outputs, merges) +
//                             scatter (ploc/mod.rs:420-487); ploc_mid_kernel / ploc_tail_kernel for the small iterations
// Node order, child slots (allocated from the END of bvh.nodes in sweep order) and AABB bits equal the sequential
// reference sweep exactly.
#include <cooperative_groups.h>

#include "common.cuh"

namespace {

// ---- ordered-int float atomics -----------------------------------------------------------------------------
__device__ __forceinline__ u32 f2ord(float f) {
    u32 u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(u32 o) {
    u32 u = o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu);
    return __uint_as_float(u);
}

struct PlocState {       // device-side loop state, double buffered by iteration parity
    u32 count;           // nodes alive in this iteration
    u32 insert_index;    // ploc/mod.rs:464-469 running allocator, counts DOWN from 2n-1
};
struct PlocGlobals {
    u32 total_ord[6];    // scene AABB as ordered ints: min xyz (atomicMin), max xyz (atomicMax)
    u32 nan_flag;
    u32 pad;
    double scale[3], offset[3];  // ploc/mod.rs:287-288
    float total[8];      // decoded scene AABB (Aabb layout)
    PlocState state[2];
    u32 ticket;
    u32 pad2;
    u32 mid_depth, mid_parity;  // written by ploc_mid_kernel: iterations done so far / parity of the live state
    // large iterations (ploc_fused_kernel), launched several at a time without a host round trip in between:
    u32 fused_depth;     // iterations completed so far
    u32 fused_stop;      // 1: the cluster count reached the hand-over size, the remaining launches do nothing; 2: no progress
    u32 fused_ticket[2]; // tile tickets, by iteration parity
};

__global__ void ploc_globals_init_kernel(PlocGlobals* g, u32 n) {
    if (threadIdx.x < 3) {
        g->total_ord[threadIdx.x] = 0xffffffffu;
        g->total_ord[3 + threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) {
        g->nan_flag = 0;
        g->state[0].count = n;
        g->state[0].insert_index = 2 * n - 1;
        g->state[1].count = 0;
        g->state[1].insert_index = 0;
        g->ticket = 0;
        g->fused_depth = 0;
        g->fused_stop = 0;
        g->fused_ticket[0] = g->fused_ticket[1] = 0;
    }
}

// K1. One leaf per primitive: Bvh2Node::new(aabb, 1, prim_index) (ploc/mod.rs:188-194). TRIS: the AABB comes from
// Triangle::aabb (triangle.rs:28-30 -> Aabb::from_points: first vertex, then extend with v1, v2).
template <bool TRIS>
__global__ void __launch_bounds__(256) leaf_init_kernel(const float4* __restrict__ src, const u32* __restrict__ indices, u32 n,
                                                        Node32* __restrict__ leaves, PlocGlobals* g) {
    float mnx = __int_as_float(0x7f800000), mny = mnx, mnz = mnx, mxx = -mnx, mxy = -mnx, mxz = -mnx;
    bool nan = false;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Box b;
        if (TRIS) {
            float4 v0 = __ldg(src + (size_t)i * 3), v1 = __ldg(src + (size_t)i * 3 + 1), v2 = __ldg(src + (size_t)i * 3 + 2);
            b.minx = smin(smin(v0.x, v1.x), v2.x); b.miny = smin(smin(v0.y, v1.y), v2.y); b.minz = smin(smin(v0.z, v1.z), v2.z);
            b.maxx = smax(smax(v0.x, v1.x), v2.x); b.maxy = smax(smax(v0.y, v1.y), v2.y); b.maxz = smax(smax(v0.z, v1.z), v2.z);
        } else {
            float4 lo = __ldg(src + (size_t)i * 2), hi = __ldg(src + (size_t)i * 2 + 1);
            b = Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
        }
        nan |= (b.minx != b.minx) | (b.miny != b.miny) | (b.minz != b.minz) | (b.maxx != b.maxx) | (b.maxy != b.maxy) | (b.maxz != b.maxz);
        store_node(leaves + i, make_node32(b, 1u, indices ? indices[i] : i));
        // total_aabb.extend(min).extend(max) (ploc/mod.rs:189-190): min and max of both corners
        mnx = fminf(mnx, fminf(b.minx, b.maxx)); mny = fminf(mny, fminf(b.miny, b.maxy)); mnz = fminf(mnz, fminf(b.minz, b.maxz));
        mxx = fmaxf(mxx, fmaxf(b.minx, b.maxx)); mxy = fmaxf(mxy, fmaxf(b.miny, b.maxy)); mxz = fmaxf(mxz, fmaxf(b.minz, b.maxz));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    __shared__ float red[8][6];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        red[w][0] = mnx; red[w][1] = mny; red[w][2] = mnz; red[w][3] = mxx; red[w][4] = mxy; red[w][5] = mxz;
    }
    if (__syncthreads_or(nan) && threadIdx.x == 0) g->nan_flag = 1;
    if (threadIdx.x < 6) {
        float v = red[0][threadIdx.x];
        for (int k = 1; k < 8; k++) v = threadIdx.x < 3 ? fminf(v, red[k][threadIdx.x]) : fmaxf(v, red[k][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&g->total_ord[threadIdx.x], f2ord(v));
        else atomicMax(&g->total_ord[threadIdx.x], f2ord(v));
    }
}

// ploc/mod.rs:287-288: scale = 1/diag (f64), offset = -min*scale. diag is computed in f32 (aabb.rs:106-108) then widened.
__global__ void morton_params_kernel(PlocGlobals* g) {
    int k = threadIdx.x;
    if (k < 3) {
        float mn = ord2f(g->total_ord[k]), mx = ord2f(g->total_ord[3 + k]);
        float diag = mx - mn;
        double scale = 1.0 / (double)diag;
        g->scale[k] = scale;
        g->offset[k] = -(double)mn * scale;
        g->total[k] = mn;
        g->total[4 + k] = mx;
    }
    if (k == 3) g->total[3] = g->total[7] = 0.f;
}

// morton.rs:35-44
__device__ __forceinline__ u64 split_by_3_u64(u32 a) {
    u64 x = (u64)a & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// K2. ploc/mod.rs:782-785: center (f32) -> f64 * scale + offset (mul then add, no FMA) -> morton_encode_u64_unorm.
// `as u32` saturates and maps NaN to 0 == cvt.rzi.u32.f64 (__double2uint_rz).
__global__ void __launch_bounds__(256) morton_kernel(const Node32* __restrict__ leaves, u32 n, const PlocGlobals* __restrict__ g,
                                                     u64* __restrict__ keys, u32* __restrict__ vals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Node32 nd = load_node(leaves + i);
    float cx = (nd.maxx + nd.minx) * 0.5f, cy = (nd.maxy + nd.miny) * 0.5f, cz = (nd.maxz + nd.minz) * 0.5f;  // aabb.rs:113-115
    double px = __dadd_rn(__dmul_rn((double)cx, g->scale[0]), g->offset[0]);
    double py = __dadd_rn(__dmul_rn((double)cy, g->scale[1]), g->offset[1]);
    double pz = __dadd_rn(__dmul_rn((double)cz, g->scale[2]), g->offset[2]);
    u32 x = __double2uint_rz(__dmul_rn(px, 2097152.0)), y = __double2uint_rz(__dmul_rn(py, 2097152.0)),
        z = __double2uint_rz(__dmul_rn(pz, 2097152.0));
    keys[i] = split_by_3_u64(x) | split_by_3_u64(y) << 1 | split_by_3_u64(z) << 2;
    vals[i] = i;
}

// K2 for SortPrecision::U128 (ploc/mod.rs:686-701,726-747; morton.rs:65-89): 42 bits per axis, `as u64` saturating cast
// (cvt.rzi.u64.f64), masked to 42 bits, interleaved into a 126-bit code kept as two 64-bit words. Code bit 3i+a is bit i
// of axis a, so the low word holds x[0..21], y[0..20], z[0..20] and the high word (code >> 64) holds y[21..41] at 3j,
// z[21..41] at 3j+1 and x[22..41] at 3j+2.
__global__ void __launch_bounds__(256) morton128_kernel(const Node32* __restrict__ leaves, u32 n, const PlocGlobals* __restrict__ g,
                                                        u64* __restrict__ keys_lo, u64* __restrict__ keys_hi, u32* __restrict__ vals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Node32 nd = load_node(leaves + i);
    float cx = (nd.maxx + nd.minx) * 0.5f, cy = (nd.maxy + nd.miny) * 0.5f, cz = (nd.maxz + nd.minz) * 0.5f;  // aabb.rs:113-115
    double px = __dadd_rn(__dmul_rn((double)cx, g->scale[0]), g->offset[0]);
    double py = __dadd_rn(__dmul_rn((double)cy, g->scale[1]), g->offset[1]);
    double pz = __dadd_rn(__dmul_rn((double)cz, g->scale[2]), g->offset[2]);
    const double k42 = 4398046511104.0;  // (1u64 << 42) as f64
    const u64 m42 = 0x3ffffffffffull;
    u64 x = __double2ull_rz(__dmul_rn(px, k42)) & m42, y = __double2ull_rz(__dmul_rn(py, k42)) & m42,
        z = __double2ull_rz(__dmul_rn(pz, k42)) & m42;
    keys_lo[i] = split_by_3_u64((u32)x) | split_by_3_u64((u32)y) << 1 | split_by_3_u64((u32)z) << 2 | ((x >> 21) & 1ull) << 63;
    keys_hi[i] = split_by_3_u64((u32)(y >> 21)) | split_by_3_u64((u32)(z >> 21)) << 1 | split_by_3_u64((u32)(x >> 22)) << 2;
    vals[i] = i;
}

// second half of the 16-level LSD sort: high words brought into the order the low-word sort produced
__global__ void __launch_bounds__(256) gather_keys_kernel(const u64* __restrict__ src, const u32* __restrict__ order, u32 n,
                                                          u64* __restrict__ dst) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __ldg(src + order[i]);
}

// K4. ploc/mod.rs:829: sorted[i] = leaves[order[i]]
__global__ void __launch_bounds__(256) gather_nodes_kernel(const Node32* __restrict__ leaves, const u32* __restrict__ order, u32 n,
                                                           Node32* __restrict__ sorted) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;  // two threads per node, one float4 each
    if (t >= 2 * n) return;
    u32 i = t >> 1, h = t & 1;
    reinterpret_cast<float4*>(sorted)[t] = __ldg(reinterpret_cast<const float4*>(leaves + order[i]) + h);
}

// ---- TMA bulk copy (global -> shared, 1-D) completing on an mbarrier -----------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(phase)
        : "memory");
}

constexpr int SEARCH_TILE = 128;  // (granularity of the scan-status allocation: the smallest tile any kernel variant uses)

__device__ __forceinline__ Box smem_box(const Node32* w, int j) {
    const float4* q = reinterpret_cast<const float4*>(w + j);
    float4 a = q[0], b = q[1];
    return Box{a.x, a.y, a.z, b.x, b.y, b.z};
}

// The neighbour choice of cluster i (window entry li). r1: the reference's r=1 fast path (ploc/mod.rs:329-382): -1 iff
// cost(i-1,i) < cost(i,i+1), first element +1, last element -1. Otherwise find_best_node (ploc/mod.rs:624-650): scan
// i-R..i-1 then i+1..i+R with `cost <= best` so the LAST minimum wins; cost(lo,hi) = half_area(union(nodes[lo], nodes[hi]))
// with the lower index first.
template <int R>
__device__ __forceinline__ int search_offset(const Node32* win, int li, u32 i, u32 count, int r1) {
    const Box me = smem_box(win, li);
    if (r1) {
        if (i == count - 1) return -1;
        float last = i > 0 ? box_half_area(box_union(smem_box(win, li - 1), me)) : __int_as_float(0x7f800000);
        float cost = box_half_area(box_union(me, smem_box(win, li + 1)));
        return last < cost ? -1 : 1;
    }
    int best = 0;
    float best_cost = __int_as_float(0x7f800000);
    const int nb = (int)min((u32)R, i), ne = (int)min((u32)R, count - 1 - i);
    for (int o = -nb; o < 0; o++) {
        float c = box_half_area(box_union(smem_box(win, li + o), me));
        if (c <= best_cost) {
            best = o;
            best_cost = c;
        }
    }
    for (int o = 1; o <= ne; o++) {
        float c = box_half_area(box_union(me, smem_box(win, li + o)));
        if (c <= best_cost) {
            best = o;
            best_cost = c;
        }
    }
    return best;
}

constexpr u64 SCAN_AGG = 1ull << 62, SCAN_INCL = 2ull << 62, SCAN_MASK = (1ull << 62) - 1;

// K6. The sequential sweep of ploc/mod.rs:423-487 visits index = 0..count: a node whose choice is not mutual is
// carried to `next`; of a mutual pair the LOWER index is skipped and the HIGHER index emits the parent at its own
// position, with left = nodes[higher], right = nodes[lower] stored at insert_index-2, -1 (slots taken downwards in
// sweep order). Positions are therefore exclusive prefix sums of (carried | emits) and of (emits): one chained scan
// over a packed (outputs, merges) pair.
// free_slots != null is the reference's REBUILD mode (ploc/mod.rs:449-462): the m-th merge of the whole build (in sweep
// order) takes the m-th freed slot pair counted from the END of bvh.nodes, i.e. free_slots[m] with the table sorted
// descending; insert_index then only counts merges (insert_start - 2 * merges).
__device__ __forceinline__ u32 child_slot(const u32* __restrict__ free_slots, u32 insert_start, u32 insert_base, u32 mi) {
    return free_slots ? free_slots[(insert_start - insert_base) / 2 + mi] : insert_base - 2 * (mi + 1);
}

// K5 + K6 of one iteration in ONE pass over the clusters (iterations above PLOC_MID_MAX clusters). A tile of 1024 clusters
// and a halo of 2R on each side is brought into shared memory by one TMA bulk copy; the CTA computes the neighbour choice of its
// own clusters AND of the R clusters on each side (what the mutual-pair test of its own clusters needs), then runs the merge
// sweep -- flags, packed decoupled-look-back scan, scatter -- reading the cluster boxes from shared memory. Against separate
// search and merge launches this reads every cluster once instead of twice, needs no merge[] array, and halves the launches.
// All loop state lives on the device (g->state[parity], g->fused_*): the host enqueues several iterations back to back, sized
// for the last count it knows (CTAs beyond the live tiles exit), and reads the state back once per batch instead of once per
// iteration. `depth` is the iteration this launch would be; a launch does nothing once g->fused_stop is set.
#ifndef OBVHS_FUSED_THREADS
#define OBVHS_FUSED_THREADS 256
#endif
#ifndef OBVHS_FUSED_ITEMS
#define OBVHS_FUSED_ITEMS 4
#endif
constexpr int FUSED_THREADS = OBVHS_FUSED_THREADS, FUSED_ITEMS = OBVHS_FUSED_ITEMS, FUSED_TILE = FUSED_THREADS * FUSED_ITEMS;

// One tile of one fused iteration. TMA: the window arrives by a bulk copy (separate launches: `cur` was written by the previous
// kernel); otherwise by ld.global.cg (ploc_mid_kernel: `cur` was written by other CTAs of the same launch). `bar_phase`: parity
// of the mbarrier phase to wait for (a CTA of the mid kernel reuses its barrier tile after tile).
template <int R, bool TMA, int ITEMS = FUSED_ITEMS>
__device__ __forceinline__ void ploc_fused_tile(const Node32* cur, Node32* next, Node32* __restrict__ bvh_nodes, PlocGlobals* g, int parity, u32 depth,
                                                u32 count, u32 insert_base, int r1, u32 tile, u64* st, const u32* __restrict__ free_slots,
                                                u32 insert_start, Node32* win, signed char* sm, u64* bar, u32 bar_phase, u64* s_wsum, u64* s_excl) {
    constexpr int TILE = FUSED_THREADS * ITEMS;  // (the mid kernel shrinks its tiles with the cluster count; the arrays are sized for FUSED_TILE)
    const u32 tiles = (count + TILE - 1) / TILE;
    const u32 tile0 = tile * TILE;
    const u32 lo = tile0 >= 2u * R ? tile0 - 2u * R : 0u;
    const u32 hi = min(count, tile0 + TILE + 2u * R);
    if (TMA) {
        if (threadIdx.x == 0) {
            const u32 bytes = (hi - lo) * (u32)sizeof(Node32);
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s(win, cur + lo, bytes, bar);
        }
        mbar_wait(bar, bar_phase);
    } else {
        for (u32 j = threadIdx.x; j < (hi - lo) * 2; j += FUSED_THREADS)
            reinterpret_cast<float4*>(win)[j] = __ldcg(reinterpret_cast<const float4*>(cur + lo) + j);
        __syncthreads();
    }
    // ---- search: own clusters, then the halo of R on each side. sm[j] <-> cluster tile0 - R + j (0 outside the array)
    // (striped over the threads: consecutive lanes read consecutive 32-byte boxes; the blocked mapping the scan below uses would
    // put a warp's 128-bit shared-memory loads on the same banks -- measured 2.5x slower at R = 6)
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const u32 j = k * FUSED_THREADS + threadIdx.x, i = tile0 + j;
        if (i < count) sm[j + R] = (signed char)search_offset<R>(win, (int)(i - lo), i, count, r1);
    }
    if (threadIdx.x < 2 * R) {
        const int j = threadIdx.x < R ? (int)threadIdx.x : TILE + (int)threadIdx.x;  // left halo 0..R-1, right halo TILE+R..TILE+2R-1
        const long long i = (long long)tile0 - R + j;
        sm[j] = (i >= 0 && i < (long long)count) ? (signed char)search_offset<R>(win, (int)(i - lo), (u32)i, count, r1) : (signed char)0;
    }
    __syncthreads();
    // ---- merge sweep (K6 above): flags, packed (outputs, merges) scan, scatter
    u32 flags = 0;
    u64 local = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const u32 i = tile0 + threadIdx.x * ITEMS + k;
        if (i < count) {
            const int li = threadIdx.x * ITEMS + k + R;
            const int m = sm[li];
            const int mb = sm[li + m];
            const bool mutual = (m + mb) == 0;
            const bool emits = mutual && m < 0;
            const bool outp = !mutual || emits;
            flags |= (outp ? 1u : 0u) << (2 * k) | (emits ? 2u : 0u) << (2 * k);
            local += (outp ? 1ull : 0ull) + (emits ? (1ull << 31) : 0ull);
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 x = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_wsum[w] = x;
    __syncthreads();
    u64 wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < FUSED_THREADS / 32; k++) {
        const u64 sv = s_wsum[k];
        if (k < w) wbase += sv;
        total += sv;
    }
    const u64 thread_excl = wbase + x - local;
    if (threadIdx.x < 32) {
        // Decoupled look-back by warp 0. Every iteration is a cold start (all tiles begin together, only tile 0 holds an
        // inclusive value), so the chain of inclusive values advances one look-back window per L2 round trip: the window is
        // 4 x 32 predecessors, all requested before any is consumed (32 per trip cost a 1024-tile iteration ~22 us).
        volatile u64* vst = st;
        u64 excl = 0;
        if (tile == 0) {
            if (lane == 0) vst[0] = SCAN_INCL | total;
        } else {
            if (lane == 0) vst[tile] = SCAN_AGG | total;
            long long t = (long long)tile - 1;  // lane k of sub-window j looks at tile t - 32 j - k
            bool done = false;
            while (!done) {
                u64 sv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const long long mine = t - 32 * j - lane;
                    sv[j] = mine >= 0 ? vst[mine] : (2ull << 62);  // SCAN_INCL | 0: tiles before the first one
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (done) break;
                    const u32 incl = __ballot_sync(0xffffffffu, (sv[j] & SCAN_INCL) != 0);
                    const u32 ready = __ballot_sync(0xffffffffu, (sv[j] & (SCAN_INCL | SCAN_AGG)) != 0);
                    // consume lanes 0..first-1 while they are ready; stop at the first inclusive value
                    const u32 not_ready = ~ready;
                    const int first_gap = not_ready ? __ffs(not_ready) - 1 : 32;
                    const int first_incl = incl ? __ffs(incl) - 1 : 32;
                    const int upto = min(first_gap, first_incl + 1);  // number of lanes whose value is consumed
                    u64 v = lane < upto ? (sv[j] & SCAN_MASK) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    excl += v;
                    t -= upto;
                    if (first_incl < first_gap) done = true;
                    else if (upto < 32) break;  // a gap: re-read from there
                }
            }
            if (lane == 0) vst[tile] = SCAN_INCL | (excl + total);
        }
        if (lane == 0) *s_excl = excl;
        if (lane == 0 && tile == tiles - 1) {  // loop state of the next iteration
            const u64 all = excl + total;
            const u32 outs = (u32)(all & 0x7fffffffull), merges = (u32)(all >> 31);
            g->state[parity ^ 1].count = outs;
            g->state[parity ^ 1].insert_index = insert_base - 2 * merges;
            g->fused_depth = depth + 1;
            if (outs >= count || outs == 0) g->fused_stop = 2;  // no progress (non-finite boxes): the host reports it
        }
    }
    __syncthreads();
    u64 run = *s_excl + thread_excl;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const u32 f = (flags >> (2 * k)) & 3u;
        if (f & 1u) {
            const u32 i = tile0 + threadIdx.x * ITEMS + k;
            const u32 pos = (u32)(run & 0x7fffffffull);
            const Node32 left = load_node(win + (i - lo));
            if (f & 2u) {
                const u32 mi = (u32)(run >> 31);
                const int m = sm[threadIdx.x * ITEMS + k + R];
                const Node32 right = load_node(win + (i - lo) + m);
                const u32 slot = child_slot(free_slots, insert_start, insert_base, mi);
                store_node(bvh_nodes + slot, left);
                store_node(bvh_nodes + slot + 1, right);
                store_node(next + pos, make_node32(box_union(node_box(left), node_box(right)), 0u, slot));
                run += 1ull + (1ull << 31);
            } else {
                store_node(next + pos, left);
                run += 1ull;
            }
        }
    }
    __syncthreads();  // the shared arrays are reused by the caller's next tile
}

template <int R>
__global__ void __launch_bounds__(FUSED_THREADS) ploc_fused_kernel(Node32* buf0, Node32* buf1, Node32* __restrict__ bvh_nodes, PlocGlobals* g,
                                                                   u32 depth, u32 search_depth_threshold, u64* scan_status, u32 status_stride,
                                                                   const u32* __restrict__ free_slots, u32 insert_start, u32 stop_at) {
    __shared__ __align__(128) Node32 win[FUSED_TILE + 4 * R];
    __shared__ signed char sm[FUSED_TILE + 2 * R];
    __shared__ __align__(8) u64 bar;
    __shared__ u64 s_wsum[FUSED_THREADS / 32];
    __shared__ u64 s_excl;
    __shared__ u32 s_tile;
    const int parity = (int)(depth & 1u);
    // reset the chained-scan state and the ticket of the NEXT iteration (its tiles are a subset of this grid)
    if (threadIdx.x == 0) {
        scan_status[(size_t)(parity ^ 1) * status_stride + blockIdx.x] = 0;
        if (blockIdx.x == 0) g->fused_ticket[parity ^ 1] = 0;
        mbar_init(&bar, 1);
    }
    // one thread reads the loop state for the whole CTA (fused_stop may be set by another CTA of this very launch: every
    // thread of a CTA must take the same exit)
    __shared__ u32 s_state[2];
    if (threadIdx.x == 0) {
        const u32 stop = *reinterpret_cast<volatile u32*>(&g->fused_stop);
        const u32 c = g->state[parity].count;
        u32 t = 0xffffffffu;
        if (stop == 0) {
            if (c <= stop_at) g->fused_stop = 1;  // small enough for the cooperative mid kernel: this and every later launch do nothing
            else t = atomicAdd(&g->fused_ticket[parity], 1u);  // tiles in ticket order: the look-back only waits for running CTAs
        }
        s_state[0] = c;
        s_state[1] = g->state[parity].insert_index;
        s_tile = t;
    }
    __syncthreads();
    const u32 count = s_state[0], insert_base = s_state[1];
    const u32 tile = s_tile;
    if (tile >= (count + FUSED_TILE - 1) / FUSED_TILE) return;
    const int r1 = (R == 1 || depth < search_depth_threshold) ? 1 : 0;
    ploc_fused_tile<R, true>(parity ? buf1 : buf0, parity ? buf0 : buf1, bvh_nodes, g, parity, depth, count, insert_base, r1, tile,
                             scan_status + (size_t)parity * status_stride, free_slots, insert_start, win, sm, &bar, 0u, s_wsum, &s_excl);
}

// Iterations between "too small to be worth a launch + a host round trip each" and the single-CTA tail: ONE cooperative
// launch loops { fused search + merge tiles -> grid barrier } until at most PLOC_TAIL clusters are left. Tiles are dealt
// round-robin (tile t to CTA t % gridDim.x), so the merge scan's look-back only ever waits for CTAs that are running.
// The window is loaded with plain L2 loads (the clusters were written by other CTAs of this launch).
// g->state[2] receives {count, insert_index} bookkeeping as usual; g->mid_depth / mid_parity report where the loop stopped.
#ifndef OBVHS_MID_SHRINK
#define OBVHS_MID_SHRINK 1
#endif
constexpr int PLOC_TAIL = 2048;  // clusters the single-CTA tail kernel takes over at
constexpr u32 PLOC_MID_MAX = 1u << 20;
template <int R>
__global__ void __launch_bounds__(FUSED_THREADS) ploc_mid_kernel(Node32* bufA, Node32* bufB, Node32* bvh_nodes, PlocGlobals* g, int parity, u32 depth,
                                                                u32 search_depth_threshold, u64* scan_status, u32 status_stride,
                                                                const u32* __restrict__ free_slots, u32 insert_start) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ __align__(128) Node32 win[FUSED_TILE + 4 * R];
    __shared__ signed char sm[FUSED_TILE + 2 * R];
    __shared__ u64 s_wsum[FUSED_THREADS / 32];
    __shared__ u64 s_excl;
    Node32 *cur = bufA, *next = bufB;  // the caller passes them in the order of `parity`
    for (;;) {
        const u32 count = __ldcg(&g->state[parity].count);
        const u32 insert_base = __ldcg(&g->state[parity].insert_index);
        if (count <= PLOC_TAIL) break;
        const int r1 = (R == 1 || depth < search_depth_threshold) ? 1 : 0;
        u64* st = scan_status + (size_t)(depth & 1u) * status_stride;
        u64* st_next = scan_status + (size_t)((depth + 1) & 1u) * status_stride;
        for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < (count + FUSED_THREADS - 1) / FUSED_THREADS; t += gridDim.x * blockDim.x)
            st_next[t] = 0;  // next iteration's scan state (for the smallest tile it may choose)
        // fused search + merge, tiles dealt round-robin (tile t to CTA t % gridDim.x): the look-back only waits for running CTAs.
        // An iteration costs a grid barrier plus the time of the tiles one CTA gets, so the tiles shrink with the cluster count:
        // once every CTA has at most one tile, 256 clusters per tile instead of 1024 cut its serial part by four (PLOC iterations of
        // the 10 M-triangle build 1.87 -> 1.66 ms; moving the two thresholds by 2-8x changes nothing measurable).
        const u32 one_item = gridDim.x * FUSED_THREADS;  // clusters the grid covers with one cluster per thread
        if (OBVHS_MID_SHRINK && count <= one_item) {
            for (u32 tile = blockIdx.x; tile < (count + FUSED_THREADS - 1) / FUSED_THREADS; tile += gridDim.x)
                ploc_fused_tile<R, false, 1>(cur, next, bvh_nodes, g, parity, depth, count, insert_base, r1, tile, st, free_slots, insert_start, win, sm,
                                             nullptr, 0u, s_wsum, &s_excl);
        } else if (OBVHS_MID_SHRINK && count <= 2 * one_item) {
            for (u32 tile = blockIdx.x; tile < (count + 2 * FUSED_THREADS - 1) / (2 * FUSED_THREADS); tile += gridDim.x)
                ploc_fused_tile<R, false, 2>(cur, next, bvh_nodes, g, parity, depth, count, insert_base, r1, tile, st, free_slots, insert_start, win, sm,
                                             nullptr, 0u, s_wsum, &s_excl);
        } else {
            for (u32 tile = blockIdx.x; tile < (count + FUSED_TILE - 1) / FUSED_TILE; tile += gridDim.x)
                ploc_fused_tile<R, false>(cur, next, bvh_nodes, g, parity, depth, count, insert_base, r1, tile, st, free_slots, insert_start, win, sm,
                                          nullptr, 0u, s_wsum, &s_excl);
        }
        grid.sync();
        Node32* tmp = cur;
        cur = next;
        next = tmp;
        parity ^= 1;
        depth++;
        if (__ldcg(&g->state[parity].count) >= count) break;  // no progress (non-finite boxes): the host reports it
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        g->mid_depth = depth;
        g->mid_parity = (u32)parity;
    }
}

// The last iterations (count <= PLOC_TAIL) in ONE CTA: clusters live in shared memory, each iteration is search ->
// flags -> block scan -> scatter, separated by __syncthreads instead of kernel launches and host round trips. Same
// arithmetic, same slots as K5/K6. Ends with bvh.nodes[0] = the last cluster (ploc/mod.rs:499).
constexpr int TAIL_THREADS = 1024, TAIL_ITEMS = PLOC_TAIL / TAIL_THREADS;
template <int R>
__global__ void __launch_bounds__(TAIL_THREADS) ploc_tail_kernel(const Node32* __restrict__ cur_g, Node32* __restrict__ bvh_nodes, PlocGlobals* g,
                                                                 int parity, u32 depth, u32 search_depth_threshold,
                                                                 const u32* __restrict__ free_slots, u32 insert_start) {
    extern __shared__ __align__(128) unsigned char tail_smem[];
    // (both buffers are addressed off the shared array itself, not through a pointer table, so the accesses compile to LDS/STS)
    Node32* const base = reinterpret_cast<Node32*>(tail_smem);
    __shared__ signed char sm[PLOC_TAIL];
    __shared__ u32 s_wsum[TAIL_THREADS / 32];
    u32 count = g->state[parity].count;
    u32 insert = g->state[parity].insert_index;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (u32 j = tid; j < count * 2; j += TAIL_THREADS) reinterpret_cast<float4*>(base)[j] = __ldg(reinterpret_cast<const float4*>(cur_g) + j);
    int src = 0;
    __syncthreads();
    while (count > 32) {  // (the last iterations, a few dozen of them with a handful of clusters each, run in one warp below)
        const Node32* cur = base + (src ? PLOC_TAIL : 0);
        Node32* next = base + (src ? 0 : PLOC_TAIL);
        const int r1 = (R == 1 || depth < search_depth_threshold) ? 1 : 0;
        const int ipt = count > (u32)TAIL_THREADS ? TAIL_ITEMS : 1;  // one cluster per thread as soon as they fit
#pragma unroll
        for (int k = 0; k < TAIL_ITEMS; k++) {
            u32 i = tid * ipt + k;
            if (k < ipt && i < count) sm[i] = (signed char)search_offset<R>(cur, (int)i, i, count, r1);
        }
        __syncthreads();
        u32 flags = 0, local = 0;  // packed: outputs in the low 16 bits, merges in the high 16
#pragma unroll
        for (int k = 0; k < TAIL_ITEMS; k++) {
            u32 i = tid * ipt + k;
            if (k < ipt && i < count) {
                int m = sm[i];
                int mb = sm[(int)i + m];
                bool mutual = (m + mb) == 0;
                bool emits = mutual && m < 0;
                bool outp = !mutual || emits;
                flags |= ((outp ? 1u : 0u) | (emits ? 2u : 0u)) << (2 * k);
                local += (outp ? 1u : 0u) + (emits ? 0x10000u : 0u);
            }
        }
        u32 x = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_wsum[w] = x;
        __syncthreads();
        u32 wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < TAIL_THREADS / 32; k++) {
            u32 sv = s_wsum[k];
            if (k < w) wbase += sv;
            total += sv;
        }
        u32 run = wbase + x - local;
#pragma unroll
        for (int k = 0; k < TAIL_ITEMS; k++) {
            u32 f = (flags >> (2 * k)) & 3u;
            if (f & 1u) {
                u32 i = tid * ipt + k;
                u32 pos = run & 0xffffu;
                Node32 left = cur[i];
                if (f & 2u) {
                    u32 mi = run >> 16;
                    Node32 right = cur[(int)i + sm[i]];
                    u32 slot = child_slot(free_slots, insert_start, insert, mi);
                    store_node(bvh_nodes + slot, left);
                    store_node(bvh_nodes + slot + 1, right);
                    next[pos] = make_node32(box_union(node_box(left), node_box(right)), 0u, slot);
                    run += 0x10001u;
                } else {
                    next[pos] = left;
                    run += 1u;
                }
            }
        }
        __syncthreads();
        if ((total & 0xffffu) >= count) break;  // no progress (non-finite boxes): leave count > 1, the host reports it
        count = total & 0xffffu;
        insert -= 2 * (total >> 16);
        src ^= 1;
        depth++;
    }
    // At most 32 clusters: one warp, one cluster per lane, warp barriers and ballots instead of block barriers and scans.
    if (w == 0) {
        const u32 lt = (1u << lane) - 1u;
        while (count > 1 && count <= 32) {
            const Node32* cur = base + (src ? PLOC_TAIL : 0);
            Node32* next = base + (src ? 0 : PLOC_TAIL);
            const int r1 = (R == 1 || depth < search_depth_threshold) ? 1 : 0;
            const bool have = (u32)lane < count;
            if (have) sm[lane] = (signed char)search_offset<R>(cur, lane, (u32)lane, count, r1);
            __syncwarp();
            bool outp = false, emits = false;
            int m = 0;
            if (have) {
                m = sm[lane];
                const int mb = sm[lane + m];
                const bool mutual = (m + mb) == 0;
                emits = mutual && m < 0;
                outp = !mutual || emits;
            }
            const u32 out_mask = __ballot_sync(0xffffffffu, outp), emit_mask = __ballot_sync(0xffffffffu, emits);
            if (outp) {
                const u32 pos = __popc(out_mask & lt);
                const Node32 left = cur[lane];
                if (emits) {
                    const Node32 right = cur[lane + m];
                    const u32 slot = child_slot(free_slots, insert_start, insert, (u32)__popc(emit_mask & lt));
                    store_node(bvh_nodes + slot, left);
                    store_node(bvh_nodes + slot + 1, right);
                    next[pos] = make_node32(box_union(node_box(left), node_box(right)), 0u, slot);
                } else {
                    next[pos] = left;
                }
            }
            __syncwarp();
            const u32 outs = (u32)__popc(out_mask);
            if (outs >= count) break;  // no progress
            count = outs;
            insert -= 2 * (u32)__popc(emit_mask);
            src ^= 1;
            depth++;
        }
    }
    if (tid < 2) reinterpret_cast<float4*>(bvh_nodes)[tid] = reinterpret_cast<const float4*>(base + (src ? PLOC_TAIL : 0))[tid];  // ploc/mod.rs:499
    if (tid == 0) {
        g->state[0].count = count;
        g->state[0].insert_index = insert;
        g->state[1].count = depth;  // total iterations, read back by the host
    }
}

__global__ void iota_kernel(u32* p, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

template <int R>
void launch_fused(ObvhsContext* ctx, u32 count, Node32* buf0, Node32* buf1, Node32* bvh_nodes, PlocGlobals* g, u32 depth, u32 thr, u64* scan_status,
                  u32 status_stride, const u32* free_slots, u32 insert_start, u32 stop_at) {
    ploc_fused_kernel<R><<<div_up(count, FUSED_TILE), FUSED_THREADS, 0, ctx->stream>>>(buf0, buf1, bvh_nodes, g, depth, thr, scan_status, status_stride,
                                                                                       free_slots, insert_start, stop_at);
}

template <int R>
cudaError_t launch_mid(ObvhsContext* ctx, u32 count, Node32* cur, Node32* next, Node32* bvh_nodes, PlocGlobals* g, int parity, u32 depth, u32 thr,
                       u64* scan_status, u32 status_stride, const u32* free_slots, u32 insert_start) {
    static PerDevice<int> per_sm_dev;
    int& per_sm = per_sm_dev[ctx->device];
    if (per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ploc_mid_kernel<R>, FUSED_THREADS, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    // as few CTAs as the work needs: the cost of a grid-wide barrier grows with the number of participants
    const int blocks = std::min(per_sm * ctx->sm_count, std::max(1, div_up(count, OBVHS_MID_SHRINK ? FUSED_THREADS : FUSED_TILE)));
    void* args[] = {&cur, &next, &bvh_nodes, &g, &parity, &depth, &thr, &scan_status, &status_stride, &free_slots, &insert_start};
    return cudaLaunchCooperativeKernel((void*)ploc_mid_kernel<R>, dim3(blocks), dim3(FUSED_THREADS), args, 0, ctx->stream);
}

template <int R>
cudaError_t launch_tail(ObvhsContext* ctx, const Node32* cur, Node32* bvh_nodes, PlocGlobals* g, int parity, u32 depth, u32 thr,
                        const u32* free_slots, u32 insert_start) {
    constexpr int smem = 2 * PLOC_TAIL * (int)sizeof(Node32);
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev[ctx->device];
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ploc_tail_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    ploc_tail_kernel<R><<<1, TAIL_THREADS, smem, ctx->stream>>>(cur, bvh_nodes, g, parity, depth, thr, free_slots, insert_start);
    return cudaGetLastError();
}

}  // namespace

static int ploc_from_leaves(ObvhsContext* ctx, ObvhsBvh2* bvh, PlocGlobals* gp, Node32* bufA_p, Node32* bufB_p, size_t n, u32 search_distance,
                            u32 sort_precision, size_t search_depth_threshold, const PlocMortonOut* probe, const u32* free_slots,
                            bool check_nan);

int ploc_build_device(ObvhsContext* ctx, const ObvhsAabb* d_aabbs, const ObvhsTriangle* d_tris, const u32* d_indices, size_t n,
                      u32 search_distance, u32 sort_precision, size_t search_depth_threshold, ObvhsBvh2** out,
                      const PlocMortonOut* probe) {
    if (sort_precision != 64 && sort_precision != 128) {
        OBVHS_SET_ERR(ctx, "sort precision %u is not SortPrecision::U64 (64) or ::U128 (128)", sort_precision);
        return OBVHS_ERR_INVALID_ARG;
    }
    switch (search_distance) {  // PlocSearchDistance::from(u32), ploc/mod.rs:550-562
        case 1: case 2: case 6: case 14: case 24: case 32: break;
        default:
            OBVHS_SET_ERR(ctx, "search distance %u is not one of 1,2,6,14,24,32", search_distance);
            return OBVHS_ERR_INVALID_ARG;
    }
    if (n >= (1u << 30)) {
        OBVHS_SET_ERR(ctx, "too many primitives: %zu", n);
        return OBVHS_ERR_UNSUPPORTED;
    }
    cudaStream_t s = ctx->stream;
    ObvhsBvh2* bvh = new ObvhsBvh2();
    bvh->device = ctx->device;
    bvh->owner = ctx;
    obvhs_context_retain(ctx);
    bvh->prim_count = n;
    struct Guard {
        ObvhsBvh2* b;
        ~Guard() { if (b) obvhs_cuda_bvh2_free(b); }
    } guard{bvh};
    // bvh2/mod.rs:106-121 reset_for_reuse: primitive_indices = indices
    if (n) {
        CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->primitive_indices, n * 4));
        if (d_indices) CU_TRY(ctx, cudaMemcpyAsync(bvh->primitive_indices, d_indices, n * 4, cudaMemcpyDeviceToDevice, s));
        else {
            iota_kernel<<<div_up(n, 256), 256, 0, s>>>(bvh->primitive_indices, (u32)n);
            KERNEL_CHECK(ctx);
        }
    }
    if (n == 0) {  // ploc/mod.rs:183-185
        if (probe && probe->total) CU_TRY(ctx, cudaMemsetAsync(probe->total, 0, sizeof(ObvhsAabb), s));
        guard.b = nullptr;
        *out = bvh;
        return OBVHS_OK;
    }
    const u32 un = (u32)n;
    bvh->node_count = 2 * n - 1;
    CU_TRY(ctx, obvhs_result_alloc(ctx, (void**)&bvh->nodes, bvh->node_count * sizeof(Node32)));

    DevBuf<PlocGlobals> g;
    DevBuf<Node32> bufA, bufB;
    CU_TRY(ctx, g.alloc(1, s));
    CU_TRY(ctx, bufA.alloc(n, s));
    CU_TRY(ctx, bufB.alloc(n, s));
    {
        TraceScope ts(ctx, "  ploc_leaves");
        ploc_globals_init_kernel<<<1, 32, 0, s>>>(g.p, un);
        KERNEL_CHECK(ctx);
        const int grid_stride_blocks = (int)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 8);
        if (d_tris) leaf_init_kernel<true><<<grid_stride_blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(d_tris), d_indices, un, bufA.p, g.p);
        else leaf_init_kernel<false><<<grid_stride_blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(d_aabbs), d_indices, un, bufA.p, g.p);
        KERNEL_CHECK(ctx);
    }
    ST_TRY(ploc_from_leaves(ctx, bvh, g.p, bufA.p, bufB.p, n, search_distance, sort_precision, search_depth_threshold, probe, nullptr, true));
    bvh->children_are_ordered_after_parents = true;  // ploc/mod.rs:502
    guard.b = nullptr;
    *out = bvh;
    return OBVHS_OK;
}

// build_ploc_from_leaves (ploc/mod.rs:265-503) on `n` leaves/subtree roots in bufA (any order): Morton codes against the scene
// box held in g->total_ord, sort, gather, then the search/merge iterations writing children into bvh->nodes (allocated by
// the caller) and the last cluster into bvh->nodes[0]. free_slots: see child_slot(). Sets max_depth / ploc_iterations.
static int ploc_from_leaves(ObvhsContext* ctx, ObvhsBvh2* bvh, PlocGlobals* gp, Node32* bufA_p, Node32* bufB_p, size_t n, u32 search_distance,
                            u32 sort_precision, size_t search_depth_threshold, const PlocMortonOut* probe, const u32* free_slots,
                            bool check_nan) {
    cudaStream_t s = ctx->stream;
    const u32 un = (u32)n;
    const u32 insert_start = 2 * un - 1;
    struct P {  // keeps the names of the code below
        PlocGlobals* p;
    } g{gp};
    struct B {
        Node32* p;
    } bufA{bufA_p}, bufB{bufB_p};
    TraceScope ts_all(ctx, " build_ploc_from_leaves");
    DevBuf<u64> keys, keys_alt;
    DevBuf<u32> vals, vals_alt;
    DevBuf<u64> scan_status;
    std::optional<TraceScope> ts_alloc;
    ts_alloc.emplace(ctx, "  preallocate_builder");
    CU_TRY(ctx, keys.alloc(n, s));
    CU_TRY(ctx, keys_alt.alloc(n, s));
    CU_TRY(ctx, vals.alloc(n, s));
    CU_TRY(ctx, vals_alt.alloc(n, s));
    const u32 status_stride = (u32)div_up(n, SEARCH_TILE) + 1;  // (one region per iteration parity for the fused kernel)
    CU_TRY(ctx, scan_status.alloc(2 * (size_t)status_stride, s));
    CU_TRY(ctx, cudaMemsetAsync(scan_status.p, 0, 2 * (size_t)status_stride * sizeof(u64), s));
    ts_alloc.reset();

    std::optional<TraceScope> tsp;  // (staged scopes: every early return closes the open one)
    tsp.emplace(ctx, "  ploc_morton");
    morton_params_kernel<<<1, 32, 0, s>>>(g.p);
    KERNEL_CHECK(ctx);
    const bool wide = sort_precision == 128;
    DevBuf<u64> keys_hi;
    if (wide) {
        CU_TRY(ctx, keys_hi.alloc(n, s));
        morton128_kernel<<<div_up(n, 256), 256, 0, s>>>(bufA.p, un, g.p, keys.p, keys_hi.p, vals.p);
    } else {
        morton_kernel<<<div_up(n, 256), 256, 0, s>>>(bufA.p, un, g.p, keys.p, vals.p);
    }
    KERNEL_CHECK(ctx);
    if (probe) {
        if (probe->codes_lo) CU_TRY(ctx, cudaMemcpyAsync(probe->codes_lo, keys.p, n * 8, cudaMemcpyDeviceToDevice, s));
        if (probe->codes_hi) {
            if (wide) CU_TRY(ctx, cudaMemcpyAsync(probe->codes_hi, keys_hi.p, n * 8, cudaMemcpyDeviceToDevice, s));
            else CU_TRY(ctx, cudaMemsetAsync(probe->codes_hi, 0, n * 8, s));
        }
        if (probe->total) CU_TRY(ctx, cudaMemcpyAsync(probe->total, g.p->total, sizeof(ObvhsAabb), cudaMemcpyDeviceToDevice, s));
    }
    tsp.reset();
    tsp.emplace(ctx, "  sort_nodes");
    u64* skeys;
    u32* order;
    ST_TRY(radix_sort_pairs_u64(ctx, keys.p, keys_alt.p, vals.p, vals_alt.p, n, 8, &skeys, &order));
    if (wide) {
        // Morton128::LEVELS = 16 (ploc/mod.rs:694-701): a stable LSD sort over the low word followed by a stable LSD sort over
        // the high word. The sorted low words are dead after the first half, so their buffer receives the gathered high words.
        gather_keys_kernel<<<div_up(n, 256), 256, 0, s>>>(keys_hi.p, order, un, skeys);
        KERNEL_CHECK(ctx);
        u64* k_alt = skeys == keys.p ? keys_alt.p : keys.p;
        u32* v_alt = order == vals.p ? vals_alt.p : vals.p;
        ST_TRY(radix_sort_pairs_u64(ctx, skeys, k_alt, order, v_alt, n, 8, &skeys, &order));
    }
    if (probe && probe->order) CU_TRY(ctx, cudaMemcpyAsync(probe->order, order, n * 4, cudaMemcpyDeviceToDevice, s));
    gather_nodes_kernel<<<div_up(2 * n, 256), 256, 0, s>>>(bufA.p, order, un, bufB.p);
    KERNEL_CHECK(ctx);

    tsp.reset();
    TraceScope ts_iter(ctx, "  ploc_iterations");
    Node32 *cur = bufB.p, *next = bufA.p;
    u32* h_state = reinterpret_cast<u32*>(ctx->pinned);
    u32 count = un;
    size_t depth = 0;
    bool nan_checked = !check_nan;
    while (count > PLOC_TAIL) {
        const int parity = (int)(depth & 1);
        if (count <= PLOC_MID_MAX) {
            // every remaining iteration above the tail size in one cooperative launch (no launches / host round trips per iteration)
            const u32 thr32 = (u32)std::min<size_t>(search_depth_threshold, 0xffffffffu);
            cudaError_t e;
            switch (search_distance) {
                case 1: e = launch_mid<1>(ctx, count, cur, next, bvh->nodes, g.p, parity, (u32)depth, thr32, scan_status.p, status_stride, free_slots, insert_start); break;
                case 2: e = launch_mid<2>(ctx, count, cur, next, bvh->nodes, g.p, parity, (u32)depth, thr32, scan_status.p, status_stride, free_slots, insert_start); break;
                case 6: e = launch_mid<6>(ctx, count, cur, next, bvh->nodes, g.p, parity, (u32)depth, thr32, scan_status.p, status_stride, free_slots, insert_start); break;
                case 14: e = launch_mid<14>(ctx, count, cur, next, bvh->nodes, g.p, parity, (u32)depth, thr32, scan_status.p, status_stride, free_slots, insert_start); break;
                case 24: e = launch_mid<24>(ctx, count, cur, next, bvh->nodes, g.p, parity, (u32)depth, thr32, scan_status.p, status_stride, free_slots, insert_start); break;
                default: e = launch_mid<32>(ctx, count, cur, next, bvh->nodes, g.p, parity, (u32)depth, thr32, scan_status.p, status_stride, free_slots, insert_start); break;
            }
            ctx->launches++;
            CU_TRY(ctx, e);
            CU_TRY(ctx, cudaMemcpyAsync(h_state, &g.p->state[0], 2 * sizeof(PlocState), cudaMemcpyDeviceToHost, s));
            CU_TRY(ctx, cudaMemcpyAsync(h_state + 6, &g.p->mid_depth, 8, cudaMemcpyDeviceToHost, s));
            if (!nan_checked) CU_TRY(ctx, cudaMemcpyAsync(h_state + 4, &g.p->nan_flag, 4, cudaMemcpyDeviceToHost, s));
            CU_TRY(ctx, cudaStreamSynchronize(s));
            if (!nan_checked && h_state[4]) {
                OBVHS_SET_ERR(ctx, "NaN in input AABBs (the reference goes out of bounds here, ploc/mod.rs:451)");
                return OBVHS_ERR_NAN_INPUT;
            }
            nan_checked = true;
            const u32 new_depth = h_state[6], new_parity = h_state[7], new_count = h_state[2 * new_parity];
            if (new_count > PLOC_TAIL) {  // the kernel only stops early when an iteration made no progress
                OBVHS_SET_ERR(ctx, "PLOC made no progress (%u clusters left); non-finite AABBs?", new_count);
                return OBVHS_ERR_NAN_INPUT;
            }
            if (((new_depth - (u32)depth) & 1u) != 0) std::swap(cur, next);
            depth = new_depth;
            count = new_count;
            break;
        }
        // large iterations: FUSED_BATCH fused search+merge launches back to back, then ONE read-back of the loop state
        constexpr int FUSED_BATCH = 4;
        const u32 thr32 = (u32)std::min<size_t>(search_depth_threshold, 0xffffffffu);
        Node32 *buf0 = (depth & 1) ? next : cur, *buf1 = (depth & 1) ? cur : next;  // buffer of the even / odd iterations
        for (int k = 0; k < FUSED_BATCH; k++) {
            const u32 d = (u32)depth + (u32)k;
            switch (search_distance) {
                case 1: launch_fused<1>(ctx, count, buf0, buf1, bvh->nodes, g.p, d, thr32, scan_status.p, status_stride, free_slots, insert_start, PLOC_MID_MAX); break;
                case 2: launch_fused<2>(ctx, count, buf0, buf1, bvh->nodes, g.p, d, thr32, scan_status.p, status_stride, free_slots, insert_start, PLOC_MID_MAX); break;
                case 6: launch_fused<6>(ctx, count, buf0, buf1, bvh->nodes, g.p, d, thr32, scan_status.p, status_stride, free_slots, insert_start, PLOC_MID_MAX); break;
                case 14: launch_fused<14>(ctx, count, buf0, buf1, bvh->nodes, g.p, d, thr32, scan_status.p, status_stride, free_slots, insert_start, PLOC_MID_MAX); break;
                case 24: launch_fused<24>(ctx, count, buf0, buf1, bvh->nodes, g.p, d, thr32, scan_status.p, status_stride, free_slots, insert_start, PLOC_MID_MAX); break;
                default: launch_fused<32>(ctx, count, buf0, buf1, bvh->nodes, g.p, d, thr32, scan_status.p, status_stride, free_slots, insert_start, PLOC_MID_MAX); break;
            }
            KERNEL_CHECK(ctx);
        }
        CU_TRY(ctx, cudaMemcpyAsync(h_state, &g.p->state[0], 2 * sizeof(PlocState), cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaMemcpyAsync(h_state + 6, &g.p->fused_depth, 8, cudaMemcpyDeviceToHost, s));
        const bool read_nan = !nan_checked;
        if (read_nan) CU_TRY(ctx, cudaMemcpyAsync(h_state + 4, &g.p->nan_flag, 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        nan_checked = true;
        if (read_nan && h_state[4]) {
            OBVHS_SET_ERR(ctx, "NaN in input AABBs (the reference goes out of bounds here, ploc/mod.rs:451)");
            return OBVHS_ERR_NAN_INPUT;
        }
        const u32 new_depth = h_state[6], stop = h_state[7];
        const u32 new_count = h_state[2 * (new_depth & 1u)];
        if (stop == 2 || new_depth <= depth || new_count >= count || new_count == 0) {
            OBVHS_SET_ERR(ctx, "PLOC made no progress (count %u -> %u); non-finite AABBs?", count, new_count);
            return OBVHS_ERR_NAN_INPUT;
        }
        if (((new_depth - (u32)depth) & 1u) != 0) std::swap(cur, next);
        depth = new_depth;
        count = new_count;
    }
    {
        // tail: everything that is left (count <= PLOC_TAIL, possibly the whole build) in one single-CTA launch
        const int parity = (int)(depth & 1);
        const u32 thr = (u32)std::min<size_t>(search_depth_threshold, 0xffffffffu);
        cudaError_t e;
        switch (search_distance) {
            case 1: e = launch_tail<1>(ctx, cur, bvh->nodes, g.p, parity, (u32)depth, thr, free_slots, insert_start); break;
            case 2: e = launch_tail<2>(ctx, cur, bvh->nodes, g.p, parity, (u32)depth, thr, free_slots, insert_start); break;
            case 6: e = launch_tail<6>(ctx, cur, bvh->nodes, g.p, parity, (u32)depth, thr, free_slots, insert_start); break;
            case 14: e = launch_tail<14>(ctx, cur, bvh->nodes, g.p, parity, (u32)depth, thr, free_slots, insert_start); break;
            case 24: e = launch_tail<24>(ctx, cur, bvh->nodes, g.p, parity, (u32)depth, thr, free_slots, insert_start); break;
            default: e = launch_tail<32>(ctx, cur, bvh->nodes, g.p, parity, (u32)depth, thr, free_slots, insert_start); break;
        }
        ctx->launches++;
        CU_TRY(ctx, e);
        CU_TRY(ctx, cudaMemcpyAsync(h_state, &g.p->state[0], 2 * sizeof(PlocState), cudaMemcpyDeviceToHost, s));
        if (!nan_checked) CU_TRY(ctx, cudaMemcpyAsync(h_state + 4, &g.p->nan_flag, 4, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        if (!nan_checked && h_state[4]) {
            OBVHS_SET_ERR(ctx, "NaN in input AABBs (the reference goes out of bounds here, ploc/mod.rs:451)");
            return OBVHS_ERR_NAN_INPUT;
        }
        if (h_state[0] != 1) {
            OBVHS_SET_ERR(ctx, "PLOC tail ended with %u clusters; non-finite AABBs?", h_state[0]);
            return OBVHS_ERR_NAN_INPUT;
        }
        depth = h_state[2];
    }
    bvh->max_depth = std::max<size_t>(96, depth + 1);  // ploc/mod.rs:501
    bvh->ploc_iterations = depth;
    return OBVHS_OK;
}

// ---- PlocBuilder::full_rebuild / partial_rebuild, compute_rebuild_path_flags (src/ploc/rebuild.rs) ---------------------------
#include "compact.cuh"

namespace {

// rebuild.rs:153: total_aabb = *bvh.nodes[0].aabb() -- the root box (possibly stale) only scales the Morton codes
__global__ void rebuild_globals_kernel(PlocGlobals* g, const Node32* __restrict__ nodes, u32 n_leaves) {
    if (threadIdx.x == 0) {
        const Node32 root = load_node(nodes);
        g->total_ord[0] = f2ord(root.minx); g->total_ord[1] = f2ord(root.miny); g->total_ord[2] = f2ord(root.minz);
        g->total_ord[3] = f2ord(root.maxx); g->total_ord[4] = f2ord(root.maxy); g->total_ord[5] = f2ord(root.maxz);
        g->nan_flag = 0;
        g->state[0].count = n_leaves;
        g->state[0].insert_index = 2 * n_leaves - 1;
        g->state[1].count = 0;
        g->state[1].insert_index = 0;
        g->ticket = 0;
        g->fused_depth = 0;
        g->fused_stop = 0;
        g->fused_ticket[0] = g->fused_ticket[1] = 0;
    }
}

// rebuild.rs:27-42: every given leaf flags itself and its ancestors. A thread may stop at a node another thread has
// flagged: whoever set that flag keeps climbing, so every path still reaches the root.
__global__ void __launch_bounds__(256) path_flags_kernel(const u32* __restrict__ leaves, u32 n_leaves, const u32* __restrict__ parents,
                                                         u32 n_nodes, u8* flags) {
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_leaves) return;
    u32 index = leaves[k];
    if (index >= n_nodes) return;
    volatile u8* vf = flags;
    vf[index] = 1;
    while (index > 0) {
        index = parents[index];
        if (vf[index]) break;
        vf[index] = 1;
    }
}

// rebuild.rs:112-131: the stack walk reaches a node iff every proper ancestor below the root is flagged (they are inner by
// construction). A reached node is collected when it is unflagged or a leaf, otherwise its children are reached too.
// cls: bit 0 = collected, bit 1 = reached (its slot is freed by set_invalid).
__global__ void __launch_bounds__(256) rebuild_classify_kernel(const Node32* __restrict__ nodes, u32 n, const u32* __restrict__ parents,
                                                               const u8* __restrict__ should_remove, u8* __restrict__ cls) {
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    if (v == 0) {
        cls[0] = 0;
        return;
    }
    bool reached = true;
    for (u32 a = parents[v]; a != 0; a = parents[a])
        if (!should_remove[a]) {
            reached = false;
            break;
        }
    const u32 prim_count = __float_as_uint(__ldg(reinterpret_cast<const float4*>(nodes + v)).w);
    const bool collected = reached && (!should_remove[v] || prim_count != 0);
    cls[v] = (collected ? 1 : 0) | (reached ? 2 : 0);
}

struct IsLeaf {  // rebuild.rs:70-74
    const Node32* nodes;
    __device__ bool operator()(u32 i) const { return __float_as_uint(__ldg(reinterpret_cast<const float4*>(nodes + i)).w) != 0; }
};
struct IsCollected {
    const u8* cls;
    __device__ bool operator()(u32 i) const { return (cls[i] & 1) != 0; }
};
struct IsFreedPairDesc {  // item i <-> the i-th odd node index counted from the end of the array
    const u8* cls;
    u32 last_odd;
    __device__ bool operator()(u32 i) const { return (cls[last_odd - 2 * i] & 2) != 0; }
};
struct EmitNode {
    const Node32* nodes;
    Node32* out;
    __device__ void operator()(u32 i, u32 rank) const { store_node(out + rank, load_node(nodes + i)); }
};
struct EmitFreedPair {
    u32* out;
    u32 last_odd;
    __device__ void operator()(u32 i, u32 rank) const { out[rank] = last_odd - 2 * i; }
};

int read_u32(ObvhsContext* ctx, const u32* d, u32* out) {
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = h[0];
    return OBVHS_OK;
}

int rebuild_from_leaves(ObvhsContext* ctx, ObvhsBvh2* bvh, PlocGlobals* g, Node32* leaves, u32 n_leaves, const u32* free_slots,
                        u32 search_distance, u32 sort_precision, size_t search_depth_threshold) {
    DevBuf<Node32> bufB;
    CU_TRY(ctx, bufB.alloc(n_leaves, ctx->stream));
    rebuild_globals_kernel<<<1, 32, 0, ctx->stream>>>(g, bvh->nodes, n_leaves);
    KERNEL_CHECK(ctx);
    ST_TRY(ploc_from_leaves(ctx, bvh, g, leaves, bufB.p, n_leaves, search_distance, sort_precision, search_depth_threshold, nullptr, free_slots,
                            false));
    bvh->children_are_ordered_after_parents = free_slots == nullptr;  // ploc/mod.rs:502: !REBUILD
    if (bvh->parents) ST_TRY(bvh2_compute_parents_device(ctx, bvh));   // rebuild.rs:177-179
    if (bvh->bvh_tris) {  // primitive_indices did not change: the permuted triangles stay valid
    }
    return OBVHS_OK;
}

int check_rebuild_args(ObvhsContext* ctx, u32 search_distance, u32 sort_precision) {
    if (sort_precision != 64 && sort_precision != 128) {
        OBVHS_SET_ERR(ctx, "sort precision %u is not SortPrecision::U64 (64) or ::U128 (128)", sort_precision);
        return OBVHS_ERR_INVALID_ARG;
    }
    switch (search_distance) {
        case 1: case 2: case 6: case 14: case 24: case 32: return OBVHS_OK;
        default:
            OBVHS_SET_ERR(ctx, "search distance %u is not one of 1,2,6,14,24,32", search_distance);
            return OBVHS_ERR_INVALID_ARG;
    }
}

}  // namespace

int ploc_compute_rebuild_path_flags_device(ObvhsContext* ctx, const ObvhsBvh2* bvh, const u32* d_leaves, size_t n_leaves, u8* d_flags) {
    if (bvh->node_count < 2) return OBVHS_OK;  // rebuild.rs:17-19
    if (!bvh->parents) {                       // rebuild.rs:20-24 panics
        OBVHS_SET_ERR(ctx, "parents must be computed before compute_rebuild_path_flags (call bvh2_compute_parents first)");
        return OBVHS_ERR_INVALID_ARG;
    }
    CU_TRY(ctx, cudaMemsetAsync(d_flags, 0, bvh->node_count, ctx->stream));
    if (n_leaves) {
        path_flags_kernel<<<div_up(n_leaves, 256), 256, 0, ctx->stream>>>(d_leaves, (u32)n_leaves, bvh->parents, (u32)bvh->node_count, d_flags);
        KERNEL_CHECK(ctx);
    }
    return OBVHS_OK;
}

int ploc_full_rebuild_device(ObvhsContext* ctx, ObvhsBvh2* bvh, u32 search_distance, u32 sort_precision, size_t search_depth_threshold) {
    if (bvh->node_count < 2) return OBVHS_OK;  // rebuild.rs:63-65
    ST_TRY(check_rebuild_args(ctx, search_distance, sort_precision));
    TraceScope ts(ctx, "full_rebuild");
    const u32 n = (u32)bvh->node_count;
    DevBuf<PlocGlobals> g;
    DevBuf<Node32> leaves;
    DevBuf<u32> tiles;
    CU_TRY(ctx, g.alloc(1, ctx->stream));
    CU_TRY(ctx, leaves.alloc((n + 1) / 2, ctx->stream));
    CU_TRY(ctx, tiles.alloc((size_t)div_up(n, CP_TILE) + 1, ctx->stream));
    // a tree over L leaves has 2L-1 nodes, so the leaves fit; collect them in node order (rebuild.rs:69-74)
    ST_TRY(compact(ctx, IsLeaf{bvh->nodes}, EmitNode{bvh->nodes, leaves.p}, n, tiles.p, &g.p->ticket));
    u32 n_leaves = 0;
    ST_TRY(read_u32(ctx, &g.p->ticket, &n_leaves));
    if (2 * (size_t)n_leaves - 1 != bvh->node_count) {
        OBVHS_SET_ERR(ctx, "full_rebuild: %u leaves in a tree of %zu nodes (not a full binary tree)", n_leaves, bvh->node_count);
        return OBVHS_ERR_INVALID_ARG;
    }
    return rebuild_from_leaves(ctx, bvh, g.p, leaves.p, n_leaves, nullptr, search_distance, sort_precision, search_depth_threshold);
}

int ploc_partial_rebuild_device(ObvhsContext* ctx, ObvhsBvh2* bvh, const u8* d_should_remove, u32 search_distance, u32 sort_precision,
                                size_t search_depth_threshold) {
    if (bvh->node_count < 2) return OBVHS_OK;  // rebuild.rs:108-110
    ST_TRY(check_rebuild_args(ctx, search_distance, sort_precision));
    TraceScope ts(ctx, "partial_rebuild");
    const u32 n = (u32)bvh->node_count;
    DevBuf<PlocGlobals> g;
    DevBuf<u32> tmp_parents, tiles, free_slots;
    DevBuf<u8> cls;
    DevBuf<Node32> leaves;
    const u32* parents = bvh->parents;
    if (!parents) {  // the walk needs them; Bvh2::parents stays None as in the reference
        CU_TRY(ctx, tmp_parents.alloc(n, ctx->stream));
        ST_TRY(bvh2_compute_parents_into(ctx, bvh, tmp_parents.p));
        parents = tmp_parents.p;
    }
    CU_TRY(ctx, g.alloc(1, ctx->stream));
    CU_TRY(ctx, cls.alloc(n, ctx->stream));
    CU_TRY(ctx, tiles.alloc((size_t)div_up(n, CP_TILE) + 1, ctx->stream));
    rebuild_classify_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(bvh->nodes, n, parents, d_should_remove, cls.p);
    KERNEL_CHECK(ctx);
    // collected nodes in ascending node index (the tie rule of partial rebuilds), freed slot pairs in descending order
    const u32 last_odd = n - 2, n_pairs = (n - 1) / 2;
    CU_TRY(ctx, leaves.alloc((n + 1) / 2, ctx->stream));
    CU_TRY(ctx, free_slots.alloc(n_pairs, ctx->stream));
    // more collected nodes than (n+1)/2 cannot happen: they are the leaves of the binary tree formed by the reached nodes
    ST_TRY(compact(ctx, IsCollected{cls.p}, EmitNode{bvh->nodes, leaves.p}, n, tiles.p, &g.p->ticket));
    u32 n_leaves = 0, n_free = 0;
    ST_TRY(read_u32(ctx, &g.p->ticket, &n_leaves));
    ST_TRY(compact(ctx, IsFreedPairDesc{cls.p, last_odd}, EmitFreedPair{free_slots.p, last_odd}, n_pairs, tiles.p, &g.p->ticket));
    ST_TRY(read_u32(ctx, &g.p->ticket, &n_free));
    if (n_leaves < 2 || n_free + 1 != n_leaves) {
        OBVHS_SET_ERR(ctx, "partial_rebuild: %u collected nodes but %u freed slot pairs (inconsistent tree)", n_leaves, n_free);
        return OBVHS_ERR_INVALID_ARG;
    }
    return rebuild_from_leaves(ctx, bvh, g.p, leaves.p, n_leaves, free_slots.p, search_distance, sort_precision, search_depth_threshold);
}
