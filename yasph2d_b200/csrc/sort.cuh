// sort.cuh -- stable LSD radix sort of (cell key, particle index) pairs, 8 bits per pass.
// Replaces `particle_indices.par_sort_unstable_by_key(cell key)` (src/sph/neighborhood_search.rs:111-118).
// Stability gives the canonical tie order (ties keep their previous relative order), see DESIGN.md "tie order".
//
// Per pass: k_radix_count (per-tile digit histogram -> table[digit][tile]), exclusive scan of the table in
// digit-major order (scan.cuh), k_radix_scatter (warp match-any ranking, stable).
#pragma once
#include "scan.cuh"

namespace yasph {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per block
constexpr int RS_BINS = 256;

inline uint32_t radix_num_tiles(uint32_t n) { return (n + RS_TILE - 1) / RS_TILE; }

__global__ void __launch_bounds__(RS_THREADS) k_radix_count(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                                                           uint32_t* __restrict__ table, uint32_t ntiles) {
    __shared__ uint32_t hist[RS_BINS];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t base = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? ((keys[i] >> shift) & 0xFFu) : 0x100u;
        unsigned mask = __match_any_sync(0xffffffffu, d);
        if (valid && lane == (unsigned)(__ffs(mask) - 1)) atomicAdd(&hist[d], (uint32_t)__popc(mask));
    }
    __syncthreads();
    table[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
    k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                    uint32_t* __restrict__ vals_out, uint32_t n, int shift, const uint32_t* __restrict__ table_scanned, uint32_t ntiles) {
    __shared__ uint32_t cnt[RS_WARPS][RS_BINS];
    __shared__ uint32_t gbase[RS_BINS];
    for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t base = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    const unsigned lt = lanemask_lt();
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0xFFFFFFFFu;
        val[r] = valid ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? ((key[r] >> shift) & 0xFFu) : 0x100u;
        unsigned mask = __match_any_sync(0xffffffffu, d);
        uint32_t b = valid ? cnt[warp][d] : 0u;
        __syncwarp();
        if (valid && lane == (unsigned)(__ffs(mask) - 1)) cnt[warp][d] = b + (uint32_t)__popc(mask);
        __syncwarp();
        rank[r] = b + (uint32_t)__popc(mask & lt);
    }
    __syncthreads();
    {  // per digit: exclusive scan over the warps, plus the tile's global base
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t t = cnt[w][d];
            cnt[w][d] = run;
            run += t;
        }
        gbase[d] = table_scanned[(size_t)d * ntiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        if (i < n) {
            uint32_t d = (key[r] >> shift) & 0xFFu;
            uint32_t pos = gbase[d] + cnt[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

}  // namespace yasph
