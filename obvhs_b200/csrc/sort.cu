// sort.cu -- stable LSD radix sort of (key, u32 value) pairs: "onesweep" (one read + one write of the pairs per
// 8-bit digit, chained-scan / decoupled look-back across tiles) with warp-level multisplit ranking (ballot per digit bit).
//
// Replaces the reference's sorts on the hot path:
//   Morton sort            src/ploc/mod.rs:811-827  (par_sort_unstable_by_key / sort_unstable_by_key / rdst radix)
//   candidate sort         src/bvh2/reinsertion.rs:138 (rdst radix on -cost, 4 levels :224-231)
//   gain sort              src/bvh2/reinsertion.rs:167-173
// The reference's sorts are unstable and tie order is unpinned; the contract shared with the oracle is
// "ties keep ascending original order", i.e. a STABLE sort. Stability here comes from: tiles are taken in ticket
// order and chained by look-back; inside a tile warps own consecutive chunks; inside a warp items are ranked item by
// item with lanes ordered by lane id (warp-striped layout == memory order).
#include <cooperative_groups.h>

#include "common.cuh"
#include "sort_tile.cuh"

namespace {

// all digit histograms in one read of the keys
template <typename K>
__global__ void __launch_bounds__(256) sort_hist_kernel(const K* __restrict__ keys, size_t n, int passes, u32* __restrict__ ghist) {
    __shared__ u32 sh[8 * 256];
    for (int t = threadIdx.x; t < passes * 256; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        K k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&sh[p * 256 + (u32)((k >> (8 * p)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < passes * 256; t += blockDim.x) {
        u32 c = sh[t];
        if (c) atomicAdd(&ghist[t], c);
    }
}

// exclusive scan of each pass's 256 bins (one block per pass)
__global__ void __launch_bounds__(256) sort_scan_kernel(const u32* __restrict__ ghist, u32* __restrict__ goffs) {
    __shared__ u32 wsum[8];
    int d = threadIdx.x, lane = d & 31, w = d >> 5;
    u32 c = ghist[blockIdx.x * 256 + d];
    u32 x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    u32 base = 0;
    for (int k = 0; k < w; k++) base += wsum[k];
    goffs[blockIdx.x * 256 + d] = base + x - c;
}

template <typename K, bool WRITE_KEYS>
__global__ void __launch_bounds__(SORT_THREADS, SortCfg<K>::MIN_CTAS) onesweep_kernel(const K* __restrict__ kin, K* __restrict__ kout,
                                                                const u32* __restrict__ vin, u32* __restrict__ vout, size_t n,
                                                                int shift, const u32* __restrict__ goffs, u32* status, u32* ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    onesweep_tile<K, WRITE_KEYS, false>(kin, kout, vin, vout, n, shift, goffs, status, s_tile, smem_raw);
}

// Mid-size sorts (a few thousand to a million pairs: the reinsertion rounds of a large scene, the Morton sort of a small one)
// in ONE cooperative launch: digit histograms -> grid barrier -> per pass { every CTA scans the pass's histogram into shared
// memory, tiles dealt round-robin } with a grid barrier between passes. Replaces 2 + passes launches that are each shorter than
// their own launch latency.
template <typename K>
__global__ void __launch_bounds__(SORT_THREADS, SortCfg<K>::MIN_CTAS) sort_mid_kernel(K* keys, K* keys_alt, u32* vals, u32* vals_alt, u32 n, int passes,
                                                                                    u32* ghist, u32* status, u32 tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 s_goffs[256];
    __shared__ u32 s_ws[SORT_WARPS];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {  // all digit histograms in one read of the keys
        u32* sh = reinterpret_cast<u32*>(smem_raw);
        for (int t = tid; t < passes * 256; t += SORT_THREADS) sh[t] = 0;
        __syncthreads();
        for (u32 i = blockIdx.x * SORT_THREADS + tid; i < n; i += gridDim.x * SORT_THREADS) {
            const K k = keys[i];
            for (int p = 0; p < passes; p++) atomicAdd(&sh[p * 256 + (u32)((k >> (8 * p)) & 0xff)], 1u);
        }
        __syncthreads();
        for (int t = tid; t < passes * 256; t += SORT_THREADS) {
            const u32 c = sh[t];
            if (c) atomicAdd(&ghist[t], c);
        }
    }
    grid.sync();
    K *kin = keys, *kout = keys_alt;
    u32 *vin = vals, *vout = vals_alt;
    for (int p = 0; p < passes; p++) {
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
        for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x)
            onesweep_tile<K, true, true>(kin, kout, vin, vout, n, 8 * p, s_goffs, status + (size_t)p * tiles * 256, tile, smem_raw);
        grid.sync();
        K* tk = kin; kin = kout; kout = tk;
        u32* tv = vin; vin = vout; vout = tv;
    }
}

// ---- single-block sort for small inputs: every pass of the same stable LSD scheme inside one CTA, data in shared memory.
// THREADS is chosen by the input size (256 / 512 / 1024): the per-pass fixed costs (histogram zeroing, the scan over the warps'
// counters, the barriers) grow with the number of warps, and most reinsertion rounds of a small scene sort < 2048 pairs.
template <typename K, int THREADS>
struct BlockSortCfg {
    static constexpr int ITEMS = sizeof(K) == 8 ? 4 : 8;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int MAX_N = THREADS * ITEMS;
    static constexpr size_t SMEM = (size_t)MAX_N * (sizeof(K) + 4) * 2 + (size_t)(WARPS * 256 + 256) * 4;
};

template <typename K, int THREADS>
__global__ void __launch_bounds__(THREADS) block_sort_kernel(const K* __restrict__ kin, const u32* __restrict__ vin, K* __restrict__ kout,
                                                            u32* __restrict__ vout, u32 n, int passes) {
    using Cfg = BlockSortCfg<K, THREADS>;
    constexpr int ITEMS = Cfg::ITEMS, MAX_N = Cfg::MAX_N, BS_WARPS = Cfg::WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* kbuf[2] = {reinterpret_cast<K*>(smem_raw), reinterpret_cast<K*>(smem_raw) + MAX_N};
    u32* vbuf[2] = {reinterpret_cast<u32*>(kbuf[1] + MAX_N), reinterpret_cast<u32*>(kbuf[1] + MAX_N) + MAX_N};
    u32* whist = vbuf[1] + MAX_N;            // [BS_WARPS][256]
    u32* dstart = whist + BS_WARPS * 256;    // [256]
    __shared__ u32 s_wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < MAX_N; j += THREADS) {
        bool ok = (u32)j < n;
        kbuf[0][j] = ok ? kin[j] : (K)~(K)0;  // padding sorts behind every valid key and stays there (stable)
        vbuf[0][j] = ok ? vin[j] : 0u;
    }
    int src = 0;
    const u32 lt = (1u << lane) - 1u;
    u32* mywh = whist + warp * 256;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        for (int j = tid; j < BS_WARPS * 256; j += THREADS) whist[j] = 0;
        __syncthreads();
        K key[ITEMS];
        u32 val[ITEMS], rank[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            int idx = warp * (ITEMS * 32) + i * 32 + lane;
            key[i] = kbuf[src][idx];
            val[i] = vbuf[src][idx];
            u32 d = (u32)((key[i] >> shift) & 0xff);
            u32 peers = __match_any_sync(0xffffffffu, d);  // one CTA, latency-bound: the single MATCH beats eight ballots here (measured)
            u32 pre = mywh[d];
            rank[i] = pre + __popc(peers & lt);
            __syncwarp();
            if ((peers & lt) == 0) mywh[d] = pre + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // digit d = tid (256 digits; with 256 threads every thread owns one): exclusive scan across warps, then across digits
        for (int d0 = 0; d0 < 256; d0 += THREADS) {
            const int d = d0 + tid;
            u32 run = 0, x = 0;
            if (d < 256) {
                for (int k = 0; k < BS_WARPS; k++) {
                    u32 c = whist[k * 256 + d];
                    whist[k * 256 + d] = run;
                    run += c;
                }
                x = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    u32 y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                if (lane == 31) s_wsum[d >> 5] = x;
            }
            __syncthreads();
            if (d < 256) {
                u32 wbase = 0;
                for (int k = 0; k < (d >> 5); k++) wbase += s_wsum[k];
                dstart[d] = wbase + x - run;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            u32 d = (u32)((key[i] >> shift) & 0xff);
            u32 pos = dstart[d] + mywh[d] + rank[i];
            kbuf[src ^ 1][pos] = key[i];
            vbuf[src ^ 1][pos] = val[i];
        }
        __syncthreads();
        src ^= 1;
    }
    for (u32 j = tid; j < n; j += THREADS) {
        kout[j] = kbuf[src][j];
        vout[j] = vbuf[src][j];
    }
}

template <typename K, int THREADS>
static int launch_block_sort(ObvhsContext* ctx, const K* keys, const u32* vals, K* keys_alt, u32* vals_alt, u32 n, int passes) {
    using Cfg = BlockSortCfg<K, THREADS>;
    static PerDevice<bool> attr_dev;
    bool& attr = attr_dev[ctx->device];
    if (!attr) {
        CU_TRY(ctx, cudaFuncSetAttribute(block_sort_kernel<K, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr = true;
    }
    block_sort_kernel<K, THREADS><<<1, THREADS, Cfg::SMEM, ctx->stream>>>(keys, vals, keys_alt, vals_alt, n, passes);
    KERNEL_CHECK(ctx);
    return OBVHS_OK;
}

constexpr size_t SORT_MID_MAX = (size_t)1 << 20;  // pairs up to which one cooperative launch does the whole sort

template <typename K>
int radix_sort_pairs(ObvhsContext* ctx, K* keys, K* keys_alt, u32* vals, u32* vals_alt, size_t n, int passes, K** sorted_keys,
                     u32** sorted_vals) {
    *sorted_keys = keys;
    *sorted_vals = vals;
    if (n < 2 || passes <= 0) return OBVHS_OK;
    if (passes > 8 || n >= (size_t)STATUS_MASK) {
        OBVHS_SET_ERR(ctx, "radix sort: unsupported size n=%zu passes=%d", n, passes);
        return OBVHS_ERR_UNSUPPORTED;
    }
    if (n <= (size_t)BlockSortCfg<K, 1024>::MAX_N) {  // one launch, no scratch
        if (n <= (size_t)BlockSortCfg<K, 256>::MAX_N) ST_TRY((launch_block_sort<K, 256>(ctx, keys, vals, keys_alt, vals_alt, (u32)n, passes)));
        else if (n <= (size_t)BlockSortCfg<K, 512>::MAX_N) ST_TRY((launch_block_sort<K, 512>(ctx, keys, vals, keys_alt, vals_alt, (u32)n, passes)));
        else ST_TRY((launch_block_sort<K, 1024>(ctx, keys, vals, keys_alt, vals_alt, (u32)n, passes)));
        *sorted_keys = keys_alt;
        *sorted_vals = vals_alt;
        return OBVHS_OK;
    }
    const size_t tiles = (n + SortCfg<K>::TILE - 1) / SortCfg<K>::TILE;
    // scratch: ghist[passes*256] goffs[passes*256] ticket[passes (padded to 8)] status[passes*tiles*256]
    const size_t words = (size_t)passes * 512 + 8 + (size_t)passes * tiles * 256;
    DevBuf<u32> scratch;
    CU_TRY(ctx, scratch.alloc(words, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(scratch.p, 0, words * 4, ctx->stream));
    u32* ghist = scratch.p;
    u32* goffs = ghist + (size_t)passes * 256;
    u32* ticket = goffs + (size_t)passes * 256;
    u32* status = ticket + 8;
    static PerDevice<bool> attr_dev;  // (one instance per key type: this is a template)
    bool& attr_ok = attr_dev[ctx->device];
    if (!attr_ok) {
        CU_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<K>()));
        CU_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<K>()));
        CU_TRY(ctx, cudaFuncSetAttribute(sort_mid_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<K>()));
        // three / four CTAs per SM need 140-180 KB of shared memory: ask for the largest carve-out
        CU_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<K, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CU_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<K, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CU_TRY(ctx, cudaFuncSetAttribute(sort_mid_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_ok = true;
    }
    if (n <= SORT_MID_MAX) {
        static PerDevice<int> per_sm_dev;
        int& per_sm = per_sm_dev[ctx->device];
        if (per_sm == 0) {
            CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sort_mid_kernel<K>, SORT_THREADS, onesweep_smem<K>()));
            if (per_sm < 1) per_sm = 1;
        }
        const int blocks = (int)std::min<size_t>((size_t)per_sm * ctx->sm_count, tiles);
        u32 un = (u32)n, utiles = (u32)tiles;
        void* args[] = {&keys, &keys_alt, &vals, &vals_alt, &un, &passes, &ghist, &status, &utiles};
        CU_TRY(ctx, cudaLaunchCooperativeKernel((void*)sort_mid_kernel<K>, dim3(blocks), dim3(SORT_THREADS), args, onesweep_smem<K>(), ctx->stream));
        KERNEL_CHECK(ctx);
        *sorted_keys = (passes & 1) ? keys_alt : keys;
        *sorted_vals = (passes & 1) ? vals_alt : vals;
        return OBVHS_OK;
    }
    int hist_blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
    sort_hist_kernel<K><<<hist_blocks, 256, 0, ctx->stream>>>(keys, n, passes, ghist);
    KERNEL_CHECK(ctx);
    sort_scan_kernel<<<passes, 256, 0, ctx->stream>>>(ghist, goffs);
    KERNEL_CHECK(ctx);
    K *kin = keys, *kout = keys_alt;
    u32 *vin = vals, *vout = vals_alt;
    for (int p = 0; p < passes; p++) {
        onesweep_kernel<K, true><<<(unsigned)tiles, SORT_THREADS, onesweep_smem<K>(), ctx->stream>>>(
            kin, kout, vin, vout, n, 8 * p, goffs + p * 256, status + (size_t)p * tiles * 256, ticket + p);
        KERNEL_CHECK(ctx);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    *sorted_keys = kin;
    *sorted_vals = vin;
    return OBVHS_OK;
}

}  // namespace

int radix_sort_pairs_u64(ObvhsContext* ctx, u64* keys, u64* keys_alt, u32* vals, u32* vals_alt, size_t n, int key_bytes,
                         u64** sorted_keys, u32** sorted_vals) {
    return radix_sort_pairs<u64>(ctx, keys, keys_alt, vals, vals_alt, n, key_bytes, sorted_keys, sorted_vals);
}
int radix_sort_pairs_u32(ObvhsContext* ctx, u32* keys, u32* keys_alt, u32* vals, u32* vals_alt, size_t n, int key_bytes,
                         u32** sorted_keys, u32** sorted_vals) {
    return radix_sort_pairs<u32>(ctx, keys, keys_alt, vals, vals_alt, n, key_bytes, sorted_keys, sorted_vals);
}
