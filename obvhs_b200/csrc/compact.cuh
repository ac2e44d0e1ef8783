// compact.cuh -- order-preserving stream compaction / exclusive prefix sums over one flag per item:
// tile counts -> single-CTA tile offsets -> scatter. `F` is a device functor bool(u32 item); `S` a device functor
// void(u32 item, u32 rank) called for every flagged item with rank = number of flagged items before it.
// Included by the .cu files that replace sequential push/retain loops of the reference (splits.rs, ploc/rebuild.rs).
#pragma once
#include "common.cuh"

namespace {

constexpr int CP_THREADS = 256, CP_ITEMS = 8, CP_TILE = CP_THREADS * CP_ITEMS;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* total) {
    __shared__ u32 warp_sums[CP_THREADS / 32];
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += x;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < CP_THREADS / 32; k++) {
        if ((u32)k < w) base += warp_sums[k];
        tot += warp_sums[k];
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

template <class F>
__global__ void __launch_bounds__(CP_THREADS) cp_count_kernel(F f, u32 n, u32* __restrict__ tile_sums) {
    const u32 base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++)
        if (base + k < n && f(base + k)) s++;
    u32 tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(CP_THREADS) cp_offsets_kernel(u32* tile_sums, u32 tiles, u32* total_out) {  // one block, in place
    u32 carry = 0;
    for (u32 base = 0; base < tiles; base += CP_THREADS) {
        const u32 i = base + threadIdx.x;
        u32 v = i < tiles ? tile_sums[i] : 0u, tot;
        const u32 ex = block_exclusive_scan(v, &tot);
        if (i < tiles) tile_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
template <class F, class S>
__global__ void __launch_bounds__(CP_THREADS) cp_scatter_kernel(F f, S sink, u32 n, const u32* __restrict__ tile_offsets) {
    const u32 base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    bool fl[CP_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
        fl[k] = base + k < n && f(base + k);
        s += fl[k] ? 1u : 0u;
    }
    u32 tot;
    u32 run = block_exclusive_scan(s, &tot) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++)
        if (fl[k]) sink(base + k, run++);
}

template <class F, class S>
int compact(ObvhsContext* ctx, F f, S sink, u32 n, u32* tile_sums, u32* total_out) {
    const u32 tiles = (u32)div_up(n, CP_TILE);
    cp_count_kernel<F><<<tiles, CP_THREADS, 0, ctx->stream>>>(f, n, tile_sums);
    KERNEL_CHECK(ctx);
    cp_offsets_kernel<<<1, CP_THREADS, 0, ctx->stream>>>(tile_sums, tiles, total_out);
    KERNEL_CHECK(ctx);
    cp_scatter_kernel<F, S><<<tiles, CP_THREADS, 0, ctx->stream>>>(f, sink, n, tile_sums);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

// ---- exclusive prefix sums over one u32 VALUE per item: F is u32(u32 item); S is void(u32 item, u32 exclusive, u32 value) ----
template <class F>
__global__ void __launch_bounds__(CP_THREADS) cpv_sum_kernel(F f, u32 n, u32* __restrict__ tile_sums) {
    const u32 base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++)
        if (base + k < n) s += f(base + k);
    u32 tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
template <class F, class S>
__global__ void __launch_bounds__(CP_THREADS) cpv_scatter_kernel(F f, S sink, u32 n, const u32* __restrict__ tile_offsets) {
    const u32 base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    u32 v[CP_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++) {
        v[k] = base + k < n ? f(base + k) : 0u;
        s += v[k];
    }
    u32 tot;
    u32 run = block_exclusive_scan(s, &tot) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < CP_ITEMS; k++)
        if (base + k < n) {
            sink(base + k, run, v[k]);
            run += v[k];
        }
}
template <class F, class S>
int scan_values(ObvhsContext* ctx, F f, S sink, u32 n, u32* tile_sums, u32* total_out) {
    const u32 tiles = (u32)div_up(n, CP_TILE);
    cpv_sum_kernel<F><<<tiles, CP_THREADS, 0, ctx->stream>>>(f, n, tile_sums);
    KERNEL_CHECK(ctx);
    cp_offsets_kernel<<<1, CP_THREADS, 0, ctx->stream>>>(tile_sums, tiles, total_out);
    KERNEL_CHECK(ctx);
    cpv_scatter_kernel<F, S><<<tiles, CP_THREADS, 0, ctx->stream>>>(f, sink, n, tile_sums);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

}  // namespace
