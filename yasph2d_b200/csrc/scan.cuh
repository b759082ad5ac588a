// scan.cuh -- device-wide exclusive scan (reduce / scan-chunk-sums / apply) with functor input and output, so the
// head-flag computation and the stream compaction that follow a scan are fused into its two data passes.
// Replaces the sequential cell creation loop of the reference (src/sph/neighborhood_search.rs:146-165).
#pragma once
#include "common.cuh"

namespace yasph {

constexpr int SCAN_THREADS = 256;
#ifndef YASPH_SCAN_ITEMS
#define YASPH_SCAN_ITEMS 16
#endif
constexpr int SCAN_ITEMS = YASPH_SCAN_ITEMS;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;  // 4096
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
    const unsigned lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += u;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Every input functor of this library returns a PAIR OF FLAGS: bit 0 and bit 32 of the 64-bit value, each 0 or 1 (cell head /
// tile head, selected for list A / list B).  The two data passes exploit that: a warp's exclusive prefix and total come
// from two ballots and population counts instead of a five-step shuffle scan of 64-bit values.
__device__ __forceinline__ void warp_flag_scan(unsigned long long v, unsigned lt, unsigned long long& excl, unsigned long long& total) {
    const unsigned ba = __ballot_sync(0xffffffffu, (uint32_t)v != 0u), bb = __ballot_sync(0xffffffffu, (uint32_t)(v >> 32) != 0u);
    excl = (unsigned long long)__popc(ba & lt) | ((unsigned long long)__popc(bb & lt) << 32);
    total = (unsigned long long)__popc(ba) | ((unsigned long long)__popc(bb) << 32);
}

// chunk_sums[b] = sum of in(i) over chunk b
template <typename T, typename In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(In in, uint32_t n, T* __restrict__ chunk_sums) {
    static_assert(sizeof(T) == 8, "pair of flags");
    __shared__ T wsum[SCAN_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned lt = lanemask_lt();
    const uint32_t base = blockIdx.x * SCAN_CHUNK + warp * (32 * SCAN_ITEMS);
    T s = 0;  // warp-uniform
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        T ex, tot;
        warp_flag_scan(i < n ? in(i) : (T)0, lt, ex, tot);
        s += tot;
    }
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) t += wsum[w];
        chunk_sums[blockIdx.x] = t;
    }
}

// in-place exclusive scan of chunk_sums[0..nchunks) by ONE block; total -> *total_out (may be null)
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chunks(T* __restrict__ chunk_sums, uint32_t nchunks, T* __restrict__ total_out) {
    __shared__ T wsum[SCAN_WARPS];
    __shared__ T carry_s;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nchunks; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        T v = i < nchunks ? chunk_sums[i] : (T)0;
        T inc = warp_inclusive_scan(v);
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        T woff = 0;
        for (uint32_t w = 0; w < warp; ++w) woff += wsum[w];
        T carry = carry_s;
        if (i < nchunks) chunk_sums[i] = carry + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// out(i, exclusive_prefix(i), in(i)) for every i
template <typename T, typename In, typename Out>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(In in, uint32_t n, const T* __restrict__ chunk_offsets, Out out) {
    __shared__ T wsum[SCAN_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t base = blockIdx.x * SCAN_CHUNK + warp * (32 * SCAN_ITEMS);
    static_assert(sizeof(T) == 8, "pair of flags");
    const unsigned lt = lanemask_lt();
    T v[SCAN_ITEMS], ex[SCAN_ITEMS];
    T carry = 0;  // warp-uniform
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        v[r] = i < n ? in(i) : (T)0;
        T e, tot;
        warp_flag_scan(v[r], lt, e, tot);
        ex[r] = carry + e;
        carry += tot;
    }
    if (lane == 0) wsum[warp] = carry;
    __syncthreads();
    T off = chunk_offsets[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) off += wsum[w];
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        uint32_t i = base + r * 32 + lane;
        if (i < n) out(i, off + ex[r], v[r]);
    }
}

// ---- the same scan in ONE kernel ---------------------------------------------------------------------------------------------
// Chunks are taken in ticket order; a chunk publishes its total as soon as it has counted its flags, and obtains its exclusive prefix
// by summing the totals of ALL earlier chunks (one warp, coalesced loads, spinning on the few that are not published yet) -- no
// serial chain of inclusive prefixes, and the input is read once.  The chunk that holds the last element also hands the grand total to
// `fin` (sentinels, counts), so no extra launch follows.  status[0] = ticket counter, status[1 + b] = total of chunk b | 1 << 63
// (a total's tile half is <= 4096, so bit 63 is free), status[1 + nchunks] = chunks that are done.  The words are zero at launch and
// the chunk that finishes last zeroes them again, so back-to-back uses need no memset between them.
template <typename In, typename Out, typename Fin>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_fused(In in, uint32_t n, unsigned long long* __restrict__ status, Out out, Fin fin) {
    typedef unsigned long long T;
    __shared__ T wsum[SCAN_WARPS];
    __shared__ T off_s;
    __shared__ uint32_t chunk_s;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned lt = lanemask_lt();
    pdl_enter();
    if (threadIdx.x == 0) chunk_s = atomicAdd(reinterpret_cast<unsigned int*>(status), 1u);
    __syncthreads();
    const uint32_t b = chunk_s;
    const uint32_t base = b * SCAN_CHUNK + warp * (32 * SCAN_ITEMS);
    T v[SCAN_ITEMS], ex[SCAN_ITEMS];
    T carry = 0;  // warp-uniform
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        v[r] = i < n ? in(i) : (T)0;
        T e, tot;
        warp_flag_scan(v[r], lt, e, tot);
        ex[r] = carry + e;
        carry += tot;
    }
    if (lane == 0) wsum[warp] = carry;
    __syncthreads();
    volatile T* st = status + 1;
    if (warp == 0) {
        T total = 0;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) total += wsum[w];
        if (lane == 0) st[b] = total | (1ull << 63);
        // exclusive prefix over the chunks: every lane sums a strided share of the earlier totals
        // (eight loads in flight per lane and round trip; a total that is not published yet is polled afterwards)
        T acc = 0;
        for (uint32_t j0 = lane; j0 < b; j0 += 32 * 8) {
            T s[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] = j0 + 32u * u < b ? st[j0 + 32u * u] : (1ull << 63);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                while ((s[u] >> 63) == 0ull) s[u] = st[j0 + 32u * u];
                acc += s[u] & ~(1ull << 63);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) off_s = acc;
    }
    __syncthreads();
    T off = off_s;
    for (uint32_t w = 0; w < warp; ++w) off += wsum[w];
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) out(i, off + ex[r], v[r]);
    }
    if (threadIdx.x == SCAN_THREADS - 1 && (uint64_t)(b + 1) * SCAN_CHUNK >= n) fin(off + carry);  // the last warp's offset + its total = the grand total
    // clean-up by the chunk that finishes last: nobody reads a status word any more
    __shared__ bool last_s;
    __syncthreads();
    const uint32_t nch = gridDim.x;
    if (threadIdx.x == 0) last_s = atomicAdd(reinterpret_cast<unsigned int*>(status + 1 + nch), 1u) == nch - 1u;
    __syncthreads();
    if (last_s)
        for (uint32_t q = threadIdx.x; q < nch + 2u; q += SCAN_THREADS) status[q] = 0ull;
}

struct U32In {
    const uint32_t* p;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
struct U32Out {
    uint32_t* p;
    __device__ __forceinline__ void operator()(uint32_t i, uint32_t ex, uint32_t) const { p[i] = ex; }
};

inline uint32_t scan_num_chunks(uint32_t n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK; }

}  // namespace yasph
