// sort_tile.cuh -- the tile routine of the stable LSD radix sort ("onesweep": warp multisplit ranks, decoupled look-back across
// tiles), shared by sort.cu (Morton / stand-alone sorts) and reinsertion.cu (the candidate and gain sorts inside the single
// cooperative launch of a reinsertion run). See sort.cu for the scheme and the stability argument.
#pragma once
#include "common.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
// pairs per thread / per tile: the 10M-key Morton sort (u64 keys) streams best with 4096-pair tiles at 3 CTAs per SM; the
// reinsertion sorts (u32 keys, at most a few hundred thousand pairs) want more, smaller tiles to fill the 148 SMs
// (tuning hooks: obvhs_b200/build.py build_variant compiles the library with other values. Measured on the 10 M-key Morton sort,
// 8 passes: 16 items x 3 CTAs 900 us, 12 x 4 1150, 8 x 4 1378, 8 x 5 1642, 20 x 2 888 -- larger tiles win, the 20 x 2 gain is noise-level)
#ifndef OBVHS_SORT_ITEMS64
#define OBVHS_SORT_ITEMS64 16
#endif
#ifndef OBVHS_SORT_CTAS64
#define OBVHS_SORT_CTAS64 3
#endif
#ifndef OBVHS_SORT_ITEMS32
#define OBVHS_SORT_ITEMS32 8
#endif
template <typename K>
struct SortCfg {
    static constexpr int ITEMS = sizeof(K) == 8 ? OBVHS_SORT_ITEMS64 : OBVHS_SORT_ITEMS32;
    static constexpr int TILE = SORT_THREADS * ITEMS;
    static constexpr int MIN_CTAS = sizeof(K) == 8 ? OBVHS_SORT_CTAS64 : 4;
};
constexpr u32 FLAG_AGG = 1u << 30, FLAG_INCL = 2u << 30, STATUS_MASK = (1u << 30) - 1;

__device__ __forceinline__ u32 ld_status(const u32* p) { return *reinterpret_cast<const volatile u32*>(p); }
__device__ __forceinline__ void st_status(u32* p, u32 v) { *reinterpret_cast<volatile u32*>(p) = v; }

// Lanes of the warp holding the same 8-bit digit. Eight ballots instead of `match.any.sync`: MATCH iterates once per
// DISTINCT value in the warp (about 30 for random digits), which made the ranking loop the bottleneck of a 10M-key pass
// (164 us for 240 MB); the ballots are full-rate and independent of the data.
__device__ __forceinline__ u32 digit_peers(u32 d) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const u32 bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// One tile of one pass. goff: exclusive digit offsets of the pass (global, or the CTA's shared copy). CG: the inputs were
// written earlier in the SAME launch by other CTAs (sort_mid_kernel), so they are read through L2.
template <typename K, bool WRITE_KEYS, bool CG>
__device__ __forceinline__ void onesweep_tile(const K* kin, K* kout, const u32* vin, u32* vout, size_t n, int shift, const u32* goffs, u32* status,
                                              u32 tile, unsigned char* smem_raw) {
    constexpr int SORT_ITEMS = SortCfg<K>::ITEMS, SORT_TILE = SortCfg<K>::TILE;
    K* skeys = reinterpret_cast<K*>(smem_raw);
    u32* svals = reinterpret_cast<u32*>(skeys + SORT_TILE);
    u32* whist = svals + SORT_TILE;          // [SORT_WARPS][256] per-warp digit counts -> exclusive offsets across warps
    u32* dstart = whist + SORT_WARPS * 256;  // [256] start of each digit inside the tile
    u32* gbase = dstart + 256;               // [256] global position of local position 0 of each digit
    __shared__ u32 s_wsum[SORT_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < SORT_WARPS; k++) whist[k * 256 + tid] = 0;
    __syncthreads();
    const size_t base = (size_t)tile * SORT_TILE;
    const u32 valid = (u32)min((size_t)SORT_TILE, n - base);
    const size_t wstart = base + (size_t)warp * (SORT_ITEMS * 32);

    K key[SORT_ITEMS];
    u32 val[SORT_ITEMS];
    u32 rank2[SORT_ITEMS / 2];  // ranks are < 4096: two per register (register budget: 3-4 CTAs per SM)
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        size_t idx = wstart + i * 32 + lane;
        bool ok = idx < n;
        key[i] = ok ? (CG ? __ldcg(kin + idx) : kin[idx]) : (K)~(K)0;  // padding sorts behind every valid key of the tile
        val[i] = ok ? (CG ? __ldcg(vin + idx) : vin[idx]) : 0u;
    }
    u32* mywh = whist + warp * 256;
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        u32 d = (u32)((key[i] >> shift) & 0xff);
        u32 peers = digit_peers(d);
        u32 pre = mywh[d];
        const u32 r = pre + __popc(peers & lt);
        if (i & 1) rank2[i / 2] |= r << 16;
        else rank2[i / 2] = r;
        __syncwarp();
        if ((peers & lt) == 0) mywh[d] = pre + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // digit `tid`: exclusive scan across warps, tile count
    u32 run = 0;
#pragma unroll
    for (int k = 0; k < SORT_WARPS; k++) {
        u32 c = whist[k * 256 + tid];
        whist[k * 256 + tid] = run;
        run += c;
    }
    // exclusive scan over digits -> dstart
    u32 x = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    u32 wbase = 0;
    for (int k = 0; k < warp; k++) wbase += s_wsum[k];
    const u32 my_start = wbase + x - run;
    dstart[tid] = my_start;
    // decoupled look-back per digit; padding keys all carry digit 255 and are not counted
    u32 agg = run - ((tid == 255) ? (SORT_TILE - valid) : 0u);
    u32 excl = 0;
    u32* st = status + (size_t)tile * 256 + tid;
    if (tile == 0) {
        st_status(st, FLAG_INCL | agg);
    } else {
        st_status(st, FLAG_AGG | agg);
        // Windowed look-back: LOOKBACK predecessor words are requested at once (independent L2 round trips in flight) and
        // then consumed in order. A one-word-at-a-time walk made the first wave of CTAs (hundreds of tiles that only have
        // aggregates yet) pay one full L2 latency per predecessor: 160 of the 164 us of a 10M-key pass. (A window of 32 after the
        // local scatter, when the key registers are free, measured slower: 966 vs 900 us for the eight passes.)
        constexpr int LOOKBACK = 8;
        long long t = (long long)tile - 1;
        bool done = false;
        while (!done) {
            u32 w[LOOKBACK];
#pragma unroll
            for (int k = 0; k < LOOKBACK; k++) w[k] = (t - k >= 0) ? ld_status(status + (size_t)(t - k) * 256 + tid) : FLAG_INCL;
            int used = 0;
#pragma unroll
            for (int k = 0; k < LOOKBACK; k++) {
                if (!done && used == k) {
                    const u32 sv = w[k];
                    if (sv & FLAG_INCL) {
                        excl += sv & STATUS_MASK;
                        done = true;
                    } else if (sv & FLAG_AGG) {
                        excl += sv & STATUS_MASK;
                        used = k + 1;
                    }  // else: not published yet -- stop consuming, re-read from this tile
                }
            }
            t -= used;
        }
        st_status(st, FLAG_INCL | (excl + agg));
    }
    gbase[tid] = goffs[tid] + excl - my_start;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        u32 d = (u32)((key[i] >> shift) & 0xff);
        u32 pos = dstart[d] + mywh[d] + ((rank2[i / 2] >> ((i & 1) * 16)) & 0xffffu);
        skeys[pos] = key[i];
        svals[pos] = val[i];
    }
    __syncthreads();
    for (u32 j = tid; j < valid; j += SORT_THREADS) {
        K k = skeys[j];
        u32 d = (u32)((k >> shift) & 0xff);
        u32 o = gbase[d] + j;
        if (WRITE_KEYS) kout[o] = k;
        vout[o] = svals[j];
    }
    __syncthreads();  // the shared arrays are reused by the caller's next tile
}

template <typename K>
constexpr size_t onesweep_smem() {
    return (size_t)SortCfg<K>::TILE * (sizeof(K) + 4) + (SORT_WARPS * 256 + 512) * 4;
}

}  // namespace
