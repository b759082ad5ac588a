// slab.cuh -- device side of the 1-D slab decomposition (multi-GPU): ordered selections, migrant / ghost packing and the
// per-pass halo pack / unpack.  No counterpart in the reference (single address space); the data it moves
// is the reference's Particles SoA (src/sph/fluidparticleworld.rs:11-23) plus the DFSPH solver arrays (dfsph.rs:36-40).
//
// Ordering contract.  Every list below is produced by an ORDERED selection (scan.cuh), never by atomics: the ghosts of a
// rank appear in its sorted array in exactly the order in which the owning rank lists the same particles in its own
// sorted array (both sides sort stably by cell key and the ghost message is packed in the sender's pre-sort order, so
// ties inside a cell keep the same relative order on both sides).  Per-pass halo exchanges therefore need no ids: entry k
// of the sender's send list is entry k of the receiver's ghost list.
#pragma once
#include "scan.cuh"

namespace yasph {

// ---- ordered selection of two index lists in one scan: value = flag A | flag B << 32 ------------------------------------
// by classification flag (migrants)
struct FlagSelIn {
    const uint8_t* pflag;
    uint8_t va, vb;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        const uint8_t f = pflag[i];
        return (f == va ? 1ull : 0ull) | ((f == vb ? 1ull : 0ull) << 32);
    }
};
// by cell column of the key among the particles that stay (pflag == 0): A: column == ca, B: column == cb
// (0xFFFFFFFF disables a list: rank 0 has no left neighbour, the last rank no right one)
struct ColumnSelIn {
    const uint32_t* keys;
    const uint8_t* pflag;  // may be null: every particle counts
    uint32_t ca, cb;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        if (pflag && pflag[i]) return 0ull;
        const uint32_t k = keys[i];
        if (k == YASPH_KEY_DROPPED) return 0ull;
        const uint32_t col = compact_1by1(k);
        return (col == ca ? 1ull : 0ull) | ((col == cb ? 1ull : 0ull) << 32);
    }
};
// ghosts of the sorted structure: A: column < col_lo (from the left rank), B: column >= col_hi (from the right rank)
struct GhostSelIn {
    const uint32_t* keys;
    uint32_t col_lo, col_hi;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        const uint32_t col = compact_1by1(keys[i]);
        return (col < col_lo ? 1ull : 0ull) | ((col >= col_hi ? 1ull : 0ull) << 32);
    }
};
// owned particles of the sorted structure (A only)
struct OwnSelIn {
    const uint8_t* pflag;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const { return pflag[i] ? 0ull : 1ull; }
};
struct IndexPairOut {
    uint32_t* a;
    uint32_t* b;
    uint8_t* pflag;  // if set: pflag[i] = 1 for selected (ghost) entries, 0 otherwise
    uint32_t cap;
    __device__ __forceinline__ void operator()(uint32_t i, unsigned long long ex, unsigned long long v) const {
        const uint32_t ea = (uint32_t)(ex & 0xFFFFFFFFull), eb = (uint32_t)(ex >> 32);
        if ((v & 0xFFFFFFFFull) && ea < cap) a[ea] = i;
        if ((v >> 32) && eb < cap) b[eb] = i;
        if (pflag) pflag[i] = v ? 1 : 0;
    }
};

// ---- particle records (migrants, ghosts): SoA block of `count` records, 8-byte arrays first ---------------------------
struct RecordArrays {
    float2* a2[3];
    float* a1[3];
    int n2, n1;
};
__host__ __device__ inline size_t record_bytes(int n2, int n1) { return (size_t)n2 * 8 + (size_t)n1 * 4; }
// buf <- arrays[idx[k]], k < count
__global__ void k_pack_records(RecordArrays arr, const uint32_t* __restrict__ idx, uint32_t count, unsigned char* __restrict__ buf) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t s = idx[k];
    float2* b2 = reinterpret_cast<float2*>(buf);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n2) b2[(size_t)q * count + k] = arr.a2[q][s];
    float* b1 = reinterpret_cast<float*>(buf + (size_t)arr.n2 * 8 * count);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n1) b1[(size_t)q * count + k] = arr.a1[q][s];
}
// arrays[first + k] <- buf, k < count; the new particles are not ghosts of the current structure (pflag = 0)
__global__ void k_unpack_records(RecordArrays arr, uint32_t first, uint32_t count, const unsigned char* __restrict__ buf, uint8_t* __restrict__ pflag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float2* b2 = reinterpret_cast<const float2*>(buf);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n2) arr.a2[q][first + k] = b2[(size_t)q * count + k];
    const float* b1 = reinterpret_cast<const float*>(buf + (size_t)arr.n2 * 8 * count);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n1) arr.a1[q][first + k] = b1[(size_t)q * count + k];
    pflag[first + k] = 0;
}
// a migrant that arrived must lie inside the slab (particles move less than a cell per step; a slab is many cells wide)
__global__ void k_check_arrivals(const uint8_t* __restrict__ pflag, uint32_t first, uint32_t end, Control* ctl) {
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end && pflag[i] != SLAB_STAY) atomicAdd(&ctl->err_slab, 1u);
}

// ---- per-pass halo exchange of one per-particle field (4 or 8 bytes per particle) ---------------------------------------
template <typename T>
__global__ void k_halo_pack(const T* __restrict__ field, const uint32_t* __restrict__ idx_l, uint32_t nl, const uint32_t* __restrict__ idx_r,
                            uint32_t nr, T* __restrict__ buf_l, T* __restrict__ buf_r) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nl)
        buf_l[k] = field[idx_l[k]];
    else if (k < nl + nr)
        buf_r[k - nl] = field[idx_r[k - nl]];
}
template <typename T>
__global__ void k_halo_unpack(T* __restrict__ field, const uint32_t* __restrict__ idx_l, uint32_t nl, const uint32_t* __restrict__ idx_r, uint32_t nr,
                              const T* __restrict__ buf_l, const T* __restrict__ buf_r) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nl)
        field[idx_l[k]] = buf_l[k];
    else if (k < nl + nr)
        field[idx_r[k - nl]] = buf_r[k - nl];
}

// compaction / scatter of one field between the local sorted array (owned + ghosts) and an owned-only array
template <typename T>
__global__ void k_own_compact(const T* __restrict__ field, const uint32_t* __restrict__ own_idx, uint32_t n_own, T* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_own) out[k] = field[own_idx[k]];
}
template <typename T>
__global__ void k_own_scatter(T* __restrict__ field, const uint32_t* __restrict__ own_idx, uint32_t n_own, const T* __restrict__ in) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_own) field[own_idx[k]] = in[k];
}
__global__ void k_iota(uint32_t* __restrict__ out, uint32_t n, uint32_t base) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = base + k;
}

}  // namespace yasph
