// sort.cuh -- stable LSD radix sort of (cell key, particle index) pairs, 8 bits per pass, one kernel per pass.
// Replaces `particle_indices.par_sort_unstable_by_key(cell key)` (src/sph/neighborhood_search.rs:111-118).
// Stability gives the canonical tie order (ties keep their previous relative order), see DESIGN.md "tie order".
//
// Structure ("onesweep"): the four global digit histograms are accumulated by the kernel that generates the keys
// (radix_hist_* below, called from k_keygen / k_advect_keygen / k_kickdrift_keygen), so the keys are never read just to be
// counted.  Each pass is then ONE kernel: a CTA takes the next tile of 4096 pairs (ticket order), ranks its keys per
// warp with match.any, publishes the tile's digit counts, obtains the counts of all earlier tiles by decoupled look-back
// over the published status words (count and state share one 32-bit word, so no fence pairs are needed), reorders the
// tile by digit in shared memory and writes each digit's run contiguously.  Per pass every pair is read once and written
// once.  A pass whose digit is the same for every key (high digits of a small domain) degenerates to a plain copy.
#pragma once
#include "common.cuh"

namespace yasph {

#ifndef YASPH_RS_THREADS
#define YASPH_RS_THREADS 512
#endif
#ifndef YASPH_RS_ITEMS
#define YASPH_RS_ITEMS 12
#endif
constexpr int RS_THREADS = YASPH_RS_THREADS;  // threads of a radix pass CTA (>= RS_BINS: thread d < RS_BINS owns digit d)
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = YASPH_RS_ITEMS;
constexpr int KG_THREADS = 256;              // threads of the key-generating kernels
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // pairs per tile of the default shape (the smallest tile: sizes the scratch)
// A pass is latency-bound per tile (ranking, digit prefix, look-back, reorder: phases in series), so a launch whose tiles do not all
// fit the GPU at once pays a second, nearly empty wave.  The host picks the smallest items-per-thread of this list whose tile count
// fits one wave (two CTAs per SM), and the default when none does (many waves: the tail no longer matters).
constexpr int RS_ITEMS_CHOICES[2] = {RS_ITEMS, RS_ITEMS + 2};  // (+4 spills too much under the 64-register cap: measured slower)
constexpr int RS_BINS = 256;
constexpr int RS_PASSES = 4;
#ifndef YASPH_RS_LOOKBACK
#define YASPH_RS_LOOKBACK 4  // measured at 2 M keys, all tiles resident at once: 1 / 2 / 3 / 4 / 8 / 16 -> sort 73 / 71 / 66 / 67 / 71 / 77 us
#endif
constexpr int RS_LOOKBACK = YASPH_RS_LOOKBACK;  // predecessor tiles inspected per look-back round trip
constexpr uint32_t RS_FLAG_LOCAL = 1u << 30, RS_FLAG_GLOBAL = 2u << 30, RS_VALUE_MASK = (1u << 30) - 1u;

inline uint32_t radix_num_tiles(uint32_t n) { return (n + RS_TILE - 1) / RS_TILE; }
// scratch layout (uint32): [RS_PASSES] tile tickets | [RS_PASSES][RS_BINS] global histograms | [RS_PASSES][ntiles][RS_BINS] status
inline size_t radix_scratch_words(uint32_t n) { return RS_PASSES + (size_t)RS_PASSES * RS_BINS + (size_t)RS_PASSES * radix_num_tiles(n) * RS_BINS; }

// ---- global digit histograms, fused into the key-generating kernels ------------------------------------------------------
// Call pattern inside a kernel: radix_hist_init(sh); ... radix_hist_add(sh, key, valid) by whole warps for each key ...;
// radix_hist_flush(sh, scratch).  The key-generating kernels are grid-stride loops, so a CTA flushes once for many keys.
struct RadixHistSmem {
    uint32_t h[RS_PASSES][RS_BINS];
};
__device__ __forceinline__ void radix_hist_init(RadixHistSmem& sh) {
    for (uint32_t q = threadIdx.x; q < RS_PASSES * RS_BINS; q += blockDim.x) (&sh.h[0][0])[q] = 0u;
    __syncthreads();
}
// Every thread of the warp calls this (valid == false for threads past the end).  The input is nearly sorted: most warps
// hold 32 keys that agree in everything but the lowest digit, and then one lane adds 32 to three counters.  Otherwise the
// lanes of a warp form a few runs of equal digits and each run's first lane adds the run length: one shuffle, one ballot
// and a handful of integer instructions per digit; no lane pair hits the same counter at once unless the same digit recurs
// in separate runs.
__device__ __forceinline__ void radix_hist_add_digit(RadixHistSmem& sh, int p, uint32_t key, bool valid, uint32_t lane) {
    const uint32_t d = valid ? ((key >> (8 * p)) & 0xFFu) : 0x1FFu;
    const uint32_t prev = __shfl_up_sync(0xffffffffu, d, 1);
    const bool head = lane == 0 || d != prev;
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (head && valid) {
        const unsigned later = lane == 31 ? 0u : (hm & (0xFFFFFFFEu << lane));
        const uint32_t len = (later ? (uint32_t)__ffs(later) - 1u : 32u) - lane;
        atomicAdd(&sh.h[p][d], len);
    }
}
__device__ __forceinline__ void radix_hist_add(RadixHistSmem& sh, uint32_t key, bool valid) {
    const uint32_t lane = lane_id();
    radix_hist_add_digit(sh, 0, key, valid, lane);
    const uint32_t hi = valid ? (key >> 8) : 0xFFFFFFFFu;  // a valid key has hi < 2^24
    const uint32_t hi0 = __shfl_sync(0xffffffffu, hi, 0);
    if (__all_sync(0xffffffffu, hi == hi0)) {
        if (lane < 3 && hi0 != 0xFFFFFFFFu) atomicAdd(&sh.h[lane + 1][(hi0 >> (8 * lane)) & 0xFFu], 32u);
    } else {
#pragma unroll
        for (int p = 1; p < RS_PASSES; ++p) radix_hist_add_digit(sh, p, key, valid, lane);
    }
}
__device__ __forceinline__ void radix_hist_flush(RadixHistSmem& sh, uint32_t* __restrict__ scratch) {
    __syncthreads();
    uint32_t* ghist = scratch + RS_PASSES;
    for (uint32_t q = threadIdx.x; q < RS_PASSES * RS_BINS; q += blockDim.x) {
        const uint32_t v = (&sh.h[0][0])[q];
        if (v) atomicAdd(&ghist[q], v);
    }
}

#ifdef YASPH_RADIX_TIMING
__device__ unsigned long long g_radix_dbg[8];
#define RT_MARK(i) { long long now_ = clock64(); if (threadIdx.x == 0) atomicAdd(&g_radix_dbg[i], (unsigned long long)(now_ - rz)); rz = now_; }
#else
#define RT_MARK(i)
#endif
// ---- one pass -------------------------------------------------------------------------------------------------------------
template <int ITEMS>
struct RadixPassSmem {
    uint32_t cnt[RS_WARPS][RS_BINS];  // per-warp digit counts, then exclusive prefix over the warps
    uint32_t lbin[RS_BINS];           // first position of the digit in the tile's digit-sorted order
    uint32_t gbin[RS_BINS];           // global position of the tile's first key of the digit, minus lbin
    uint32_t skey[RS_THREADS * ITEMS];
    uint32_t sval[RS_THREADS * ITEMS];
    uint32_t wsum[RS_WARPS];
    uint32_t tile;
    uint32_t trivial;
};
template <int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, 2)
    k_radix_pass(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                 uint32_t* __restrict__ vals_out, uint32_t n, int pass, uint32_t* __restrict__ scratch, uint32_t ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RS_ITEMS = ITEMS, RS_TILE = RS_THREADS * ITEMS;  // this instance's tile shape
    RadixPassSmem<ITEMS>& S = *reinterpret_cast<RadixPassSmem<ITEMS>*>(smem_raw);
    const int shift = pass * 8;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    const unsigned lt = lanemask_lt();
    const uint32_t* ghist = scratch + RS_PASSES + pass * RS_BINS;
    volatile uint32_t* status = scratch + RS_PASSES + RS_PASSES * RS_BINS + (size_t)pass * ntiles * RS_BINS;
    static_assert(RS_THREADS >= RS_BINS && RS_BINS == 256, "thread d < RS_BINS owns digit d");
#ifdef YASPH_RADIX_TIMING
    long long rz = clock64();
#endif
    for (uint32_t q = tid; q < RS_WARPS * RS_BINS; q += RS_THREADS) (&S.cnt[0][0])[q] = 0u;
    pdl_enter();
    if (tid == 0) {
        S.tile = atomicAdd(&scratch[pass], 1u);
        S.trivial = 0u;
    }
    __syncthreads();
    const uint32_t tile = S.tile;
    const bool owner = tid < RS_BINS;
    const uint32_t gh = owner ? ghist[tid] : 0u;
    if (owner && gh == n) S.trivial = 1u;  // every key has this digit: the pass is the identity
    const uint32_t base = tile * RS_TILE + warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0xFFFFFFFFu;
        val[r] = valid ? (vals_in != nullptr ? vals_in[i] : i) : 0u;  // vals_in == null: the identity (the first pass of a sort)
    }
    __syncthreads();
    RT_MARK(0)  // ticket + load
    if (S.trivial) {
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r) {
            const uint32_t i = base + r * 32 + lane;
            if (i < n) {
                keys_out[i] = key[r];
                vals_out[i] = val[r];
            }
        }
        return;
    }
    // 1. rank inside the warp (stable: rows in order, lanes in order)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((key[r] >> shift) & 0xFFu) : 0x100u;
        const unsigned mask = __match_any_sync(0xffffffffu, d);
        const uint32_t b = valid ? S.cnt[warp][d] : 0u;
        __syncwarp();
        if (valid && lane == (unsigned)(__ffs(mask) - 1)) S.cnt[warp][d] = b + (uint32_t)__popc(mask);
        __syncwarp();
        rank[r] = b + (uint32_t)__popc(mask & lt);
    }
    __syncthreads();
    RT_MARK(1)  // rank
    // 2. per digit (thread == digit): prefix over the warps, tile count, publish, look back
    uint32_t tcount = 0;
    if (owner) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t t = S.cnt[w][tid];
            S.cnt[w][tid] = tcount;
            tcount += t;
        }
        status[(size_t)tile * RS_BINS + tid] = tcount | RS_FLAG_LOCAL;
    }
    // exclusive scan of the tile counts over the digits (position of each digit in the tile's sorted order) and of the
    // global histogram (position of each digit in the output)
    uint32_t lex, gex;
    {
        uint32_t a = tcount, g = gh;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xffffffffu, a, o), ug = __shfl_up_sync(0xffffffffu, g, o);
            if (lane >= (uint32_t)o) {
                a += ua;
                g += ug;
            }
        }
        if (lane == 31 && owner) {
            S.wsum[warp] = a;
            S.lbin[warp] = g;  // borrowed until the sync below
        }
        __syncthreads();
        uint32_t wa = 0, wg = 0;
        for (uint32_t w = 0; w < warp && w < RS_BINS / 32; ++w) {
            wa += S.wsum[w];
            wg += S.lbin[w];
        }
        __syncthreads();
        lex = wa + a - tcount;
        gex = wg + g - gh;
    }
    if (owner) {
        // Decoupled look-back, RS_LOOKBACK predecessors per round trip: their status words are loaded together (independent
        // loads in flight at once) and then consumed nearest first, so the prefix chain advances a window of tiles per L2 latency.
        uint32_t excl = 0;
        for (int j = (int)tile - 1; j >= 0; j -= RS_LOOKBACK) {
            uint32_t sw[RS_LOOKBACK];
#pragma unroll
            for (int u = 0; u < RS_LOOKBACK; ++u) sw[u] = j - u >= 0 ? (uint32_t)status[(size_t)(j - u) * RS_BINS + tid] : (uint32_t)(2u << 30);
            bool done = false;
#pragma unroll
            for (int u = 0; u < RS_LOOKBACK; ++u) {
                if (!done) {
                    while ((sw[u] & ~RS_VALUE_MASK) == 0u) sw[u] = status[(size_t)(j - u) * RS_BINS + tid];  // not published yet
                    excl += sw[u] & RS_VALUE_MASK;
                    done = (sw[u] & RS_FLAG_GLOBAL) != 0u;
                }
            }
            if (done) break;
        }
        status[(size_t)tile * RS_BINS + tid] = (excl + tcount) | RS_FLAG_GLOBAL;
        S.lbin[tid] = lex;
        S.gbin[tid] = gex + excl - lex;
    }
    __syncthreads();
    RT_MARK(2)  // digit prefix, publish, look back
    // 3. reorder the tile by digit in shared memory
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & 0xFFu;
            const uint32_t lp = S.lbin[d] + S.cnt[warp][d] + rank[r];
            S.skey[lp] = key[r];
            S.sval[lp] = val[r];
        }
    }
    __syncthreads();
    RT_MARK(3)  // reorder in shared memory
    // 4. write every digit's run contiguously
    const uint32_t tile_n = min((uint32_t)RS_TILE, n - tile * RS_TILE);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t lp = r * RS_THREADS + tid;
        if (lp < tile_n) {
            const uint32_t k = S.skey[lp];
            const uint32_t gp = S.gbin[(k >> shift) & 0xFFu] + lp;
            keys_out[gp] = k;
            vals_out[gp] = S.sval[lp];
        }
    }
    RT_MARK(4)  // write out
#ifdef YASPH_RADIX_TIMING
    if (tid == 0) atomicAdd(&g_radix_dbg[5], 1ull);
#endif
}

}  // namespace yasph
