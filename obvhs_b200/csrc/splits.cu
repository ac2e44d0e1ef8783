// splits.cu -- spatial pre-splits of large triangles on the device.
//
// Replaces src/splits.rs:16-158 (`split_aabbs_preset`, `split_aabbs_precise`, `split_triangle`) and the builders' pre-split
// prologue (cwbvh/builder.rs:27-54 == bvh2/builder.rs:24-51: triangle AABBs, average and largest half area).
//
// The reference's loop is sequential, but its data flow is not: inside one of the <= 12 iterations every candidate reads
// and rewrites only its OWN aabb, and the only order-dependent quantity is where the right halves are appended
// (`aabbs.len()` at the time the candidate is processed). That slot is `len + (splits among earlier candidates)`, an
// exclusive prefix sum over the candidates in list order; `candidates.retain(..)` is an order-preserving compaction.
// So one iteration = evaluate (one thread per candidate) -> scan -> scatter -> compact, and the arrays come out in
// exactly the reference's order.
//
// The average half area is a SEQUENTIAL f32 sum in the reference (`avg_area += half_area`, cwbvh/builder.rs:38); float
// addition does not reassociate, so one thread adds in index order while the rest of its CTA stages the next chunk
// in shared memory (4 cycles per dependent FADD; it sits outside the reference's core_build_time bracket too).
#include <algorithm>

#include "common.cuh"
#include "compact.cuh"

namespace {

struct SplitState {
    float avg, largest;        // cwbvh/builder.rs:29-43
    u32 largest_bits, total;   // atomicMax target (non-negative floats order like their bits); scan total
    float area_thresh_low, area_thresh_high, split_factor_low, split_factor_high;  // splits.rs:49-58
};

__device__ __forceinline__ Box ld_box(const float4* __restrict__ aabbs, u32 i) {
    const float4 lo = aabbs[2 * (size_t)i], hi = aabbs[2 * (size_t)i + 1];
    return Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
}
__device__ __forceinline__ void st_box(float4* aabbs, u32 i, const Box& b) {
    aabbs[2 * (size_t)i] = make_float4(b.minx, b.miny, b.minz, 0.f);
    aabbs[2 * (size_t)i + 1] = make_float4(b.maxx, b.maxy, b.maxz, 0.f);
}

// Triangle::aabb + half_area + f32::max of the half areas (cwbvh/builder.rs:31-41)
__global__ void __launch_bounds__(256) presplit_aabb_kernel(const float4* __restrict__ tris, u32 n, float4* __restrict__ aabbs,
                                                            u32* __restrict__ indices, float* __restrict__ half, SplitState* st) {
    float mx = 0.f;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 v0 = __ldg(tris + (size_t)i * 3), v1 = __ldg(tris + (size_t)i * 3 + 1), v2 = __ldg(tris + (size_t)i * 3 + 2);
        Box b;
        b.minx = smin(smin(v0.x, v1.x), v2.x); b.miny = smin(smin(v0.y, v1.y), v2.y); b.minz = smin(smin(v0.z, v1.z), v2.z);
        b.maxx = smax(smax(v0.x, v1.x), v2.x); b.maxy = smax(smax(v0.y, v1.y), v2.y); b.maxz = smax(smax(v0.z, v1.z), v2.z);
        st_box(aabbs, i, b);
        indices[i] = i;
        const float h = box_half_area(b);
        half[i] = h;
        mx = fmaxf(h, mx);  // f32::max: NaN is ignored
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(&st->largest_bits, __float_as_uint(mx));
}

// avg_area = (((h0 + h1) + h2) + ...) / n in f32, then split_aabbs_preset's thresholds (splits.rs:23-33)
constexpr int SUM_THREADS = 256, SUM_CHUNK = 4096;
__global__ void __launch_bounds__(SUM_THREADS) presplit_sum_kernel(const float* __restrict__ half, u32 n, SplitState* st) {
    __shared__ float buf[2][SUM_CHUNK];
    float acc = 0.f;
    const u32 chunks = (n + SUM_CHUNK - 1) / SUM_CHUNK;
    for (u32 k = threadIdx.x; k < SUM_CHUNK; k += SUM_THREADS) buf[0][k] = k < n ? half[k] : 0.f;
    __syncthreads();
    for (u32 c = 0; c < chunks; c++) {
        const float* cur = buf[c & 1];
        float* nxt = buf[(c + 1) & 1];
        if (threadIdx.x == 0) {
            const u32 m = min((u32)SUM_CHUNK, n - c * SUM_CHUNK);
            u32 k = 0;
            for (; k + 8 <= m; k += 8) {
                const float4 a = *reinterpret_cast<const float4*>(cur + k), b = *reinterpret_cast<const float4*>(cur + k + 4);
                acc = __fadd_rn(acc, a.x); acc = __fadd_rn(acc, a.y); acc = __fadd_rn(acc, a.z); acc = __fadd_rn(acc, a.w);
                acc = __fadd_rn(acc, b.x); acc = __fadd_rn(acc, b.y); acc = __fadd_rn(acc, b.z); acc = __fadd_rn(acc, b.w);
            }
            for (; k < m; k++) acc = __fadd_rn(acc, cur[k]);
        } else if (c + 1 < chunks) {
            const u32 base = (c + 1) * SUM_CHUNK;
            for (u32 k = threadIdx.x - 1; k < SUM_CHUNK; k += SUM_THREADS - 1) nxt[k] = base + k < n ? half[base + k] : 0.f;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float avg = __fdiv_rn(acc, (float)n);
        const float largest = __uint_as_float(st->largest_bits);
        st->avg = avg;
        st->largest = largest;
        st->area_thresh_low = __fmul_rn(avg, 3.0f);
        st->area_thresh_high = fmaxf(__fmul_rn(avg, 4.0f), __fadd_rn(__fmul_rn(avg, 0.9f), __fmul_rn(largest, 0.1f)));
        st->split_factor_low = 1.8f;
        st->split_factor_high = 1.6f;
    }
}

// flag sources
struct AreaOfAabb {      // splits.rs:62-66: aabb.half_area() > area_thresh_low, over all aabbs
    const float4* aabbs;
    const SplitState* st;
    __device__ bool operator()(u32 i) const { return box_half_area(ld_box(aabbs, i)) > st->area_thresh_low; }
};
struct AreaOfCandidate {  // splits.rs:121: candidates.retain(|c| aabbs[*c].half_area() > area_thresh_low)
    const float4* aabbs;
    const u32* cand;
    const SplitState* st;
    __device__ bool operator()(u32 i) const { return box_half_area(ld_box(aabbs, cand[i])) > st->area_thresh_low; }
};
struct StoredFlag {
    const u8* flags;
    __device__ bool operator()(u32 i) const { return flags[i] != 0; }
};

// sinks: called as sink(item, rank) for every flagged item, rank = number of flagged items before it
struct EmitIndex {  // candidates.push(i) in index order (splits.rs:62-66)
    u32* out;
    __device__ void operator()(u32 i, u32 rank) const { out[rank] = i; }
};
struct EmitCandidate {  // retain keeps the candidate VALUE
    const u32* cand;
    u32* out;
    __device__ void operator()(u32 i, u32 rank) const { out[rank] = cand[i]; }
};
struct EmitSplit {  // splits.rs:112-116: candidates.push(aabbs.len()); aabbs.push(right); indices.push(index)
    float4* aabbs;
    u32* indices;
    u32* cand;
    const float4* rights;
    u32 len, count;
    __device__ void operator()(u32 i, u32 rank) const {
        const u32 slot = len + rank;
        aabbs[2 * (size_t)slot] = rights[2 * (size_t)i];
        aabbs[2 * (size_t)slot + 1] = rights[2 * (size_t)i + 1];
        indices[slot] = indices[cand[i]];
        cand[count + rank] = slot;
    }
};
// ---- one candidate: splits.rs:71-117 ------------------------------------------------------------------------------
__device__ __forceinline__ float axis_of(const float4& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ void extend(Box& b, float x, float y, float z) {  // aabb.rs:76-80: union(self, from_point(p))
    b.minx = smin(b.minx, x); b.miny = smin(b.miny, y); b.minz = smin(b.minz, z);
    b.maxx = smax(b.maxx, x); b.maxy = smax(b.maxy, y); b.maxz = smax(b.maxz, z);
}
__device__ __forceinline__ Box intersection(const Box& a, const Box& b) {  // aabb.rs:97-103
    return Box{smax(a.minx, b.minx), smax(a.miny, b.miny), smax(a.minz, b.minz), smin(a.maxx, b.maxx), smin(a.maxy, b.maxy), smin(a.maxz, b.maxz)};
}
// splits.rs:129-158; Vec3A::mul_add is fused on every glam backend
__device__ __forceinline__ void split_triangle(int dim, float pos, const float4 (&v)[3], Box& left, Box& right) {
    const float FMAX = 3.40282347e+38f;
    left = Box{FMAX, FMAX, FMAX, -FMAX, -FMAX, -FMAX};  // Aabb::INVALID, aabb.rs:23-26
    right = left;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float4 v0 = v[i], v1 = v[(i + 1) % 3];
        const float v0d = axis_of(v0, dim), v1d = axis_of(v1, dim);
        if (v0d <= pos) extend(left, v0.x, v0.y, v0.z);
        if (v0d >= pos) extend(right, v0.x, v0.y, v0.z);
        if ((v0d < pos && pos < v1d) || (v1d < pos && pos < v0d)) {
            const float inv_length = __fdiv_rn(1.0f, __fsub_rn(v1d, v0d));
            const float t = __fmul_rn(__fsub_rn(pos, v0d), inv_length);
            const float cx = __fmaf_rn(t, __fsub_rn(v1.x, v0.x), v0.x), cy = __fmaf_rn(t, __fsub_rn(v1.y, v0.y), v0.y),
                        cz = __fmaf_rn(t, __fsub_rn(v1.z, v0.z), v0.z);
            extend(left, cx, cy, cz);
            extend(right, cx, cy, cz);
        }
    }
}

__global__ void __launch_bounds__(128) split_eval_kernel(float4* aabbs, const u32* __restrict__ indices, const u32* __restrict__ cand,
                                                         u32 count, const float4* __restrict__ tris, const SplitState* __restrict__ st,
                                                         u32 split_tests, float4* __restrict__ rights, u8* __restrict__ flags) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u32 c = cand[i];
    const Box aabb = ld_box(aabbs, c);
    const u32 index = indices[c];
    const float dx = __fsub_rn(aabb.maxx, aabb.minx), dy = __fsub_rn(aabb.maxy, aabb.miny), dz = __fsub_rn(aabb.maxz, aabb.minz);
    const int axis = dx < dy ? (dy < dz ? 2 : 1) : (dx < dz ? 2 : 0);  // aabb.rs:124-134 largest_axis
    const float4 tv[3] = {__ldg(tris + (size_t)index * 3), __ldg(tris + (size_t)index * 3 + 1), __ldg(tris + (size_t)index * 3 + 2)};
    const float amin = axis == 0 ? aabb.minx : (axis == 1 ? aabb.miny : aabb.minz);
    const float amax = axis == 0 ? aabb.maxx : (axis == 1 ? aabb.maxy : aabb.maxz);
    float best_cost = 3.40282347e+38f;
    Box left = aabb, right = aabb;
    for (u32 k = 1; k < split_tests; k++) {
        const float n = __fdiv_rn((float)k, (float)split_tests);
        const float pos = __fadd_rn(__fmul_rn(amin, n), __fmul_rn(amax, __fsub_rn(1.0f, n)));
        Box tmp_left = aabb, tmp_right = aabb;
        if (axis == 0) { tmp_left.maxx = pos; tmp_right.minx = pos; }
        else if (axis == 1) { tmp_left.maxy = pos; tmp_right.miny = pos; }
        else { tmp_left.maxz = pos; tmp_right.minz = pos; }
        Box t_left, t_right;
        split_triangle(axis, pos, tv, t_left, t_right);
        tmp_left = intersection(t_left, tmp_left);
        tmp_right = intersection(t_right, tmp_right);
        const float area = __fadd_rn(box_half_area(tmp_left), box_half_area(tmp_right));
        if (area < best_cost) {
            best_cost = area;
            left = tmp_left;
            right = tmp_right;
        }
    }
    const float old_cost = box_half_area(aabb);
    const bool split = (st->area_thresh_high > old_cost && __fmul_rn(best_cost, st->split_factor_high) < old_cost) ||
                       __fmul_rn(best_cost, st->split_factor_low) < old_cost;
    flags[i] = split ? 1 : 0;
    if (split) {
        st_box(aabbs, c, left);
        st_box(rights, i, right);
    }
}

__global__ void split_state_init_kernel(SplitState* st, float lo, float hi, float f_lo, float f_hi) {
    st->avg = 0.f;
    st->largest = 0.f;
    st->largest_bits = 0;
    st->total = 0;
    st->area_thresh_low = lo;
    st->area_thresh_high = hi;
    st->split_factor_low = f_lo;
    st->split_factor_high = f_hi;
}

template <class T>
int grow(ObvhsContext* ctx, T*& p, size_t& cap, size_t used, size_t need) {
    if (need <= cap) return OBVHS_OK;
    const size_t ncap = std::max(need, cap + cap / 2);
    T* q = static_cast<T*>(obvhs_arena_alloc(ncap * sizeof(T)));
    if (!q) {
        OBVHS_SET_ERR(ctx, "pre-splits: out of device memory growing to %zu entries", ncap);
        return OBVHS_ERR_CUDA;
    }
    if (used) CU_TRY(ctx, cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    p = q;
    cap = ncap;
    return OBVHS_OK;
}

int read_total(ObvhsContext* ctx, const SplitState* st, u32* out) {
    u32* h = reinterpret_cast<u32*>(ctx->pinned);
    CU_TRY(ctx, cudaMemcpyAsync(h, &st->total, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = h[0];
    return OBVHS_OK;
}

// splits.rs:60-125 on device arrays: a.aabbs / a.indices hold a.len entries with room for a.cap and are re-pointed at
// larger arena blocks as the set grows (Vec::push); the thresholds are read from *st.
int split_loop(ObvhsContext* ctx, SplitArrays& a, const ObvhsTriangle* d_tris, SplitState* st, u32 max_iterations, u32 split_tests) {
    const size_t n = a.len;
    if (n == 0 || max_iterations == 0) return OBVHS_OK;
    if (n >= (1u << 30)) {
        OBVHS_SET_ERR(ctx, "pre-splits: too many primitives: %zu", n);
        return OBVHS_ERR_UNSUPPORTED;
    }
    cudaStream_t s = ctx->stream;
    u32* tiles_p = nullptr;
    size_t tile_cap = 0;
    ST_TRY(grow(ctx, tiles_p, tile_cap, 0, (size_t)div_up(n, CP_TILE) + 1));
    // candidates = positions of the aabbs above the low threshold, ascending (:62-66)
    u32 *cand = nullptr, *cand_alt = nullptr;
    size_t cand_cap = 0, cand_alt_cap = 0;
    u32 count = 0;
    {
        const u32 tiles = (u32)div_up(n, CP_TILE);
        AreaOfAabb f{reinterpret_cast<const float4*>(a.aabbs), st};
        cp_count_kernel<AreaOfAabb><<<tiles, CP_THREADS, 0, s>>>(f, (u32)n, tiles_p);
        KERNEL_CHECK(ctx);
        cp_offsets_kernel<<<1, CP_THREADS, 0, s>>>(tiles_p, tiles, &st->total);
        KERNEL_CHECK(ctx);
        ST_TRY(read_total(ctx, st, &count));
        if (count == 0) return OBVHS_OK;
        ST_TRY(grow(ctx, cand, cand_cap, 0, (size_t)count * 2));
        cp_scatter_kernel<AreaOfAabb, EmitIndex><<<tiles, CP_THREADS, 0, s>>>(f, EmitIndex{cand}, (u32)n, tiles_p);
        KERNEL_CHECK(ctx);
    }
    ObvhsAabb* rights = nullptr;
    u8* flags = nullptr;
    size_t rights_cap = 0, flags_cap = 0;
    for (u32 it = 0; it < max_iterations && count > 0; it++) {
        if (a.len + count >= (1u << 30)) {
            OBVHS_SET_ERR(ctx, "pre-splits: too many primitives: %zu", a.len + count);
            return OBVHS_ERR_UNSUPPORTED;
        }
        if (a.len + count > a.cap) {  // every candidate may append one entry this iteration
            size_t c1 = a.cap, c2 = a.cap;
            ST_TRY(grow(ctx, a.aabbs, c1, a.len, a.len + count));
            ST_TRY(grow(ctx, a.indices, c2, a.len, c1));
            a.cap = c1;
        }
        ST_TRY(grow(ctx, rights, rights_cap, 0, count));
        ST_TRY(grow(ctx, flags, flags_cap, 0, count));
        ST_TRY(grow(ctx, cand, cand_cap, count, (size_t)count * 2));
        ST_TRY(grow(ctx, tiles_p, tile_cap, 0, (size_t)div_up((size_t)count * 2, CP_TILE) + 1));
        split_eval_kernel<<<div_up(count, 128), 128, 0, s>>>(reinterpret_cast<float4*>(a.aabbs), a.indices, cand, count,
                                                            reinterpret_cast<const float4*>(d_tris), st, split_tests,
                                                            reinterpret_cast<float4*>(rights), flags);
        KERNEL_CHECK(ctx);
        ST_TRY(compact(ctx, StoredFlag{flags},
                       EmitSplit{reinterpret_cast<float4*>(a.aabbs), a.indices, cand, reinterpret_cast<const float4*>(rights), (u32)a.len, count},
                       count, tiles_p, &st->total));
        u32 added = 0;
        ST_TRY(read_total(ctx, st, &added));
        if (added == 0) break;  // :118-119
        a.len += added;
        const u32 all = count + added;
        ST_TRY(grow(ctx, cand_alt, cand_alt_cap, 0, (size_t)all * 2));
        ST_TRY(compact(ctx, AreaOfCandidate{reinterpret_cast<const float4*>(a.aabbs), cand, st}, EmitCandidate{cand, cand_alt}, all, tiles_p, &st->total));
        ST_TRY(read_total(ctx, st, &count));  // :121-122
        std::swap(cand, cand_alt);
        std::swap(cand_cap, cand_alt_cap);
    }
    return OBVHS_OK;
}

}  // namespace

// split_aabbs_precise (splits.rs:49-125) over device arrays
int split_aabbs_precise_device(ObvhsContext* ctx, SplitArrays& a, const ObvhsTriangle* d_tris, float area_thresh_low, float area_thresh_high,
                               float split_factor_low, float split_factor_high, u32 max_iterations, u32 split_tests) {
    TraceScope ts(ctx, "split_aabbs_precise");
    DevBuf<SplitState> st;
    CU_TRY(ctx, st.alloc(1, ctx->stream));
    split_state_init_kernel<<<1, 1, 0, ctx->stream>>>(st.p, area_thresh_low, area_thresh_high, split_factor_low, split_factor_high);
    KERNEL_CHECK(ctx);
    return split_loop(ctx, a, d_tris, st.p, max_iterations, split_tests);
}

// The builders' pre-split branch (cwbvh/builder.rs:27-54, bvh2/builder.rs:24-51): AABBs of the triangles, average (sequential
// f32 sum) and largest half area, then split_aabbs_preset (splits.rs:16-34). ev_start, when given, is recorded where the
// reference starts its core_build_time clock (after the AABB pass, before the splits).
int presplit_tris_device(ObvhsContext* ctx, const ObvhsTriangle* d_tris, size_t n, SplitArrays& a, float* avg_largest_host, cudaEvent_t ev_start) {
    a = SplitArrays{};
    cudaStream_t s = ctx->stream;
    if (n == 0) {
        if (ev_start) CU_TRY(ctx, cudaEventRecord(ev_start, s));
        return OBVHS_OK;
    }
    if (n >= (1u << 30)) {
        OBVHS_SET_ERR(ctx, "pre-splits: too many primitives: %zu", n);
        return OBVHS_ERR_UNSUPPORTED;
    }
    size_t c1 = 0, c2 = 0;
    const size_t cap0 = n + n / 8 + 1024;
    ST_TRY(grow(ctx, a.aabbs, c1, 0, cap0));
    ST_TRY(grow(ctx, a.indices, c2, 0, cap0));
    a.cap = cap0;
    a.len = n;
    DevBuf<SplitState> st;
    DevBuf<float> half;
    CU_TRY(ctx, st.alloc(1, s));
    CU_TRY(ctx, half.alloc(n, s));
    {
        TraceScope ts(ctx, "presplit_aabbs_avg");
        split_state_init_kernel<<<1, 1, 0, s>>>(st.p, 0.f, 0.f, 0.f, 0.f);
        KERNEL_CHECK(ctx);
        const int blocks = (int)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 8);
        presplit_aabb_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(d_tris), (u32)n, reinterpret_cast<float4*>(a.aabbs), a.indices,
                                                   half.p, st.p);
        KERNEL_CHECK(ctx);
        presplit_sum_kernel<<<1, SUM_THREADS, 0, s>>>(half.p, (u32)n, st.p);
        KERNEL_CHECK(ctx);
    }
    if (ev_start) CU_TRY(ctx, cudaEventRecord(ev_start, s));
    if (avg_largest_host) {
        float* h = reinterpret_cast<float*>(ctx->pinned) + 16;
        CU_TRY(ctx, cudaMemcpyAsync(h, &st.p->avg, 8, cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        avg_largest_host[0] = h[0];
        avg_largest_host[1] = h[1];
    }
    TraceScope ts(ctx, "split_aabbs_preset");
    return split_loop(ctx, a, d_tris, st.p, 12, 12);
}
