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
// by cell column of the key among the particles that stay (pflag == 0): A: column in [a_lo, a_hi), B: column in [b_lo, b_hi)
// (an empty range disables a list: rank 0 has no left neighbour, the last rank no right one)
struct ColumnSelIn {
    const uint32_t* keys;
    const uint8_t* pflag;  // may be null: every particle counts
    uint32_t a_lo, a_hi, b_lo, b_hi;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        if (pflag && pflag[i]) return 0ull;
        const uint32_t k = keys[i];
        if (k == YASPH_KEY_DROPPED) return 0ull;
        const uint32_t col = compact_1by1(k);
        return ((col >= a_lo && col < a_hi) ? 1ull : 0ull) | (((col >= b_lo && col < b_hi) ? 1ull : 0ull) << 32);
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

// ---- the four lists of a neighbourhood update in ONE ordered selection ------------------------------------------------------
// After key generation has classified every local particle (slab_classify: stays / migrates left / right / dropped ghost), the
// update needs, each in ascending index order: the migrants to the left and to the right rank, and the stayers in the first /
// last W owned columns (the ghost layers the left / right rank keeps of this slab).  One reduce / scan / apply over four flags.
struct Select4In {
    const uint32_t* keys;
    const uint8_t* pflag;
    uint32_t a_lo, a_hi, b_lo, b_hi;  // ghost-send column ranges (left | right); empty = side disabled
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const uint8_t f = pflag[i];
        if (f == SLAB_MIG_LEFT) return 1u;
        if (f == SLAB_MIG_RIGHT) return 2u;
        if (f != SLAB_STAY) return 0u;
        const uint32_t col = compact_1by1(keys[i]);
        return ((col >= a_lo && col < a_hi) ? 4u : 0u) | ((col >= b_lo && col < b_hi) ? 8u : 0u);
    }
};
__device__ __forceinline__ uint4 operator+(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// per-warp exclusive prefix and total of four flags over 32 lanes
__device__ __forceinline__ void warp_flag_scan4(uint32_t m, unsigned lt, uint4& excl, uint4& total) {
    const unsigned b0 = __ballot_sync(0xffffffffu, m & 1u), b1 = __ballot_sync(0xffffffffu, m & 2u), b2 = __ballot_sync(0xffffffffu, m & 4u),
                   b3 = __ballot_sync(0xffffffffu, m & 8u);
    excl = make_uint4(__popc(b0 & lt), __popc(b1 & lt), __popc(b2 & lt), __popc(b3 & lt));
    total = make_uint4(__popc(b0), __popc(b1), __popc(b2), __popc(b3));
}
// One kernel (the scheme of scan.cuh's k_scan_fused): a chunk publishes its four counts early -- 15 bits each in one
// 64-bit status word, bit 63 = published -- and sums the counts of all earlier chunks itself.  status[0] = ticket counter,
// status[1 + b] = counts of chunk b; the caller zeroes 1 + nchunks words.  The chunk that holds the last element writes the totals.
static_assert(SCAN_CHUNK < (1 << 15), "four 15-bit counts per status word");
__global__ void __launch_bounds__(SCAN_THREADS) k_select4_fused(Select4In in, uint32_t n, unsigned long long* __restrict__ status, uint32_t* __restrict__ out0,
                                                                uint32_t* __restrict__ out1, uint32_t* __restrict__ out2, uint32_t* __restrict__ out3, uint32_t cap,
                                                                uint32_t* __restrict__ totals_out) {
    __shared__ uint4 wsum[SCAN_WARPS];
    __shared__ uint4 off_s;
    __shared__ uint32_t chunk_s;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const unsigned lt = lanemask_lt();
    pdl_enter();
    if (threadIdx.x == 0) chunk_s = atomicAdd(reinterpret_cast<unsigned int*>(status), 1u);
    __syncthreads();
    const uint32_t b = chunk_s;
    const uint32_t base = b * SCAN_CHUNK + warp * (32 * SCAN_ITEMS);
    uint32_t m[SCAN_ITEMS];
    uint4 ex[SCAN_ITEMS];
    uint4 carry = make_uint4(0, 0, 0, 0);  // warp-uniform
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        m[r] = i < n ? in(i) : 0u;
        uint4 e, tot;
        warp_flag_scan4(m[r], lt, e, tot);
        ex[r] = carry + e;
        carry = carry + tot;
    }
    if (lane == 0) wsum[warp] = carry;
    __syncthreads();
    volatile unsigned long long* st = status + 1;
    if (warp == 0) {
        uint4 total = wsum[0];
#pragma unroll
        for (int w = 1; w < SCAN_WARPS; ++w) total = total + wsum[w];
        if (lane == 0)
            st[b] = (unsigned long long)total.x | ((unsigned long long)total.y << 15) | ((unsigned long long)total.z << 30) | ((unsigned long long)total.w << 45) | (1ull << 63);
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (uint32_t j0 = lane; j0 < b; j0 += 32 * 8) {  // eight loads in flight per lane and round trip (scan.cuh: k_scan_fused)
            unsigned long long v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = j0 + 32u * u < b ? st[j0 + 32u * u] : (1ull << 63);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                while ((v[u] >> 63) == 0ull) v[u] = st[j0 + 32u * u];
                acc = acc + make_uint4((uint32_t)v[u] & 0x7FFFu, (uint32_t)(v[u] >> 15) & 0x7FFFu, (uint32_t)(v[u] >> 30) & 0x7FFFu, (uint32_t)(v[u] >> 45) & 0x7FFFu);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc = acc + make_uint4(__shfl_xor_sync(0xffffffffu, acc.x, o), __shfl_xor_sync(0xffffffffu, acc.y, o), __shfl_xor_sync(0xffffffffu, acc.z, o),
                                   __shfl_xor_sync(0xffffffffu, acc.w, o));
        if (lane == 0) off_s = acc;
    }
    __syncthreads();
    uint4 off = off_s;
    for (uint32_t w = 0; w < warp; ++w) off = off + wsum[w];
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (m[r] & 1u) { const uint32_t e = off.x + ex[r].x; if (e < cap) out0[e] = i; }
        if (m[r] & 2u) { const uint32_t e = off.y + ex[r].y; if (e < cap) out1[e] = i; }
        if (m[r] & 4u) { const uint32_t e = off.z + ex[r].z; if (e < cap) out2[e] = i; }
        if (m[r] & 8u) { const uint32_t e = off.w + ex[r].w; if (e < cap) out3[e] = i; }
    }
    if (threadIdx.x == SCAN_THREADS - 1 && (uint64_t)(b + 1) * SCAN_CHUNK >= n) {
        const uint4 t = off + carry;
        totals_out[0] = t.x;
        totals_out[1] = t.y;
        totals_out[2] = t.z;
        totals_out[3] = t.w;
    }
    // clean-up by the chunk that finishes last (as k_scan_fused): status[1 + nchunks] counts the chunks that are done
    __shared__ bool last_s;
    __syncthreads();
    const uint32_t nch = gridDim.x;
    if (threadIdx.x == 0) last_s = atomicAdd(reinterpret_cast<unsigned int*>(status + 1 + nch), 1u) == nch - 1u;
    __syncthreads();
    if (last_s)
        for (uint32_t q = threadIdx.x; q < nch + 2u; q += SCAN_THREADS) status[q] = 0ull;
}

// ---- particle records (migrants, ghosts): SoA block of `count` records, 8-byte arrays first ---------------------------
struct RecordArrays {
    float2* a2[3];
    float* a1[3];
    int n2, n1;
};
__host__ __device__ inline size_t record_bytes(int n2, int n1) { return (size_t)n2 * 8 + (size_t)n1 * 4; }
// buf <- arrays[idx(k)], k < na + nb, idx(k) = k < na ? idx_a[k] : idx_b[k - na]: one message = the migrants, then the ghost layer
__global__ void k_pack_records(RecordArrays arr, const uint32_t* __restrict__ idx_a, uint32_t na, const uint32_t* __restrict__ idx_b, uint32_t nb,
                               unsigned char* __restrict__ buf) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, count = na + nb;
    if (k >= count) return;
    const uint32_t s = k < na ? idx_a[k] : idx_b[k - na];
    float2* b2 = reinterpret_cast<float2*>(buf);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n2) b2[(size_t)q * count + k] = arr.a2[q][s];
    float* b1 = reinterpret_cast<float*>(buf + (size_t)arr.n2 * 8 * count);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n1) b1[(size_t)q * count + k] = arr.a1[q][s];
}
// arrays[first + k] <- record src_off + k of a message of `stride` records, k < count; the new particles are not ghosts of the
// current structure (pflag = 0)
__global__ void k_unpack_records(RecordArrays arr, uint32_t first, uint32_t count, const unsigned char* __restrict__ buf, uint32_t stride, uint32_t src_off,
                                 uint8_t* __restrict__ pflag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float2* b2 = reinterpret_cast<const float2*>(buf);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n2) arr.a2[q][first + k] = b2[(size_t)q * stride + src_off + k];
    const float* b1 = reinterpret_cast<const float*>(buf + (size_t)arr.n2 * 8 * stride);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < arr.n1) arr.a1[q][first + k] = b1[(size_t)q * stride + src_off + k];
    pflag[first + k] = 0;
}
// a migrant that arrived must lie inside the slab (particles move less than a cell per step; a slab is many cells wide)
__global__ void k_check_arrivals(const uint8_t* __restrict__ pflag, uint32_t first, uint32_t end, Control* ctl) {
    pdl_enter();
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

// ---- peer-memory transport (NVLink / NVSwitch, one process per GPU) ---------------------------------------------------------
// Every rank owns a MAILBOX in its device memory that its peers map (CUDA IPC) and write with plain stores over NVLink:
//   halo exchange    k_halo_exchange gathers the send lists and stores the values straight into the neighbours' mailboxes, the
//                    last CTA to finish publishes the sequence number (release, system scope); the same kernel then waits for
//                    the neighbours' numbers (acquire) and scatters what they stored here to the ghost slots -- one small kernel
//                    per exchange and rank, no staging copy, no collective call;
//   all-reduce       k_allreduce_peer stores the rank's scalar into every peer's mailbox, waits for every peer's scalar and
//                    combines them in rank order (every rank computes the same bits).
// Payload and all-reduce slots are double-buffered by the parity of the sequence number: the ranks run the same sequence of
// exchanges, every exchange moves data in both directions of a link, so a rank can be at most one exchange ahead of a peer
// that still reads the previous payload.  A wait that exceeds PEER_TIMEOUT_NS raises Control::err_comm instead of hanging.
constexpr int PEER_MAX_RANKS = 16;
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;
struct PeerBoxHeader {
    unsigned long long halo_flag[2];            // [side]: sequence number of the last complete halo message from that side
    unsigned long long ar_flag[PEER_MAX_RANKS]; // [source rank]: sequence number of its last all-reduce contribution
    double ar_val[2][PEER_MAX_RANKS];           // [parity][source rank]
};
constexpr size_t PEER_HEADER_BYTES = (sizeof(PeerBoxHeader) + 255) & ~(size_t)255;
constexpr size_t PEER_RECORD_MAX_BYTES = 3 * 8 + 3 * 4;  // the widest particle record (three float2 + three float arrays)
constexpr size_t PEER_MSG_HEADER_BYTES = 16;             // record messages: count of records, then the SoA block
// one payload slot holds the largest message: a halo field (8 bytes per particle) or a block of particle records
__host__ __device__ inline size_t peer_slot_bytes(uint32_t max_halo) { return (PEER_MSG_HEADER_BYTES + (size_t)max_halo * PEER_RECORD_MAX_BYTES + 255) & ~(size_t)255; }
__host__ __device__ inline size_t peer_box_bytes(uint32_t max_halo) { return PEER_HEADER_BYTES + 4 * peer_slot_bytes(max_halo); }
__host__ __device__ inline unsigned char* peer_payload(void* box, uint32_t max_halo, unsigned parity, int side) {
    return reinterpret_cast<unsigned char*>(box) + PEER_HEADER_BYTES + (size_t)(parity * 2 + side) * peer_slot_bytes(max_halo);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// waits until *flag >= seq; false on timeout
__device__ __forceinline__ bool peer_wait(const unsigned long long* flag, unsigned long long seq) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < seq) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > PEER_TIMEOUT_NS) return false;
    }
    return true;
}
// push and pull of one exchange in ONE launch: every CTA first stores its share of the send lists into the neighbours'
// mailboxes (the last one to finish publishes the sequence number), then waits for the neighbours' numbers and scatters its
// share of what they stored here.  Grid-stride loops over a grid of at most one CTA per SM: all CTAs are resident at once, so
// a CTA that already waits for the other rank can never keep a CTA that still has to push from running.
template <typename T>
__global__ void k_halo_exchange(T* __restrict__ field, const uint32_t* __restrict__ send_l, uint32_t nsl, const uint32_t* __restrict__ send_r, uint32_t nsr,
                                T* dst_l, T* dst_r, unsigned long long* pflag_l, unsigned long long* pflag_r, const uint32_t* __restrict__ ghost_l,
                                uint32_t ngl, const uint32_t* __restrict__ ghost_r, uint32_t ngr, const T* src_l, const T* src_r,
                                const unsigned long long* flag_l, const unsigned long long* flag_r, unsigned long long seq, unsigned int* ticket,
                                Control* ctl) {
    const uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t k = k0; k < nsl + nsr; k += stride) {
        if (k < nsl)
            dst_l[k] = field[send_l[k]];
        else
            dst_r[k - nsl] = field[send_r[k - nsl]];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence_system();
            if (pflag_l) st_release_sys(pflag_l, seq);
            if (pflag_r) st_release_sys(pflag_r, seq);
        }
        ok = 1;
        if (flag_l && !peer_wait(flag_l, seq)) ok = 0;
        if (flag_r && !peer_wait(flag_r, seq)) ok = 0;
        if (!ok) atomicOr(&ctl->err_comm, 1u);
    }
    __syncthreads();
    if (!ok) return;
    for (uint32_t k = k0; k < ngl + ngr; k += stride) {
        if (k < ngl)
            field[ghost_l[k]] = __ldcg(&src_l[k]);
        else
            field[ghost_r[k - ngl]] = __ldcg(&src_r[k - ngl]);
    }
}

// ---- particle records (migrants + ghost layers) through the mailboxes: the COUNTS stay on the device ------------------------
// One message per neighbour and update: [count of migrants, count of ghost-layer particles | SoA block of both, migrants first].
// k_records_exchange reads how many particles the ordered selection picked (device memory), packs them straight into the neighbours'
// mailboxes and publishes the sequence number, waits for both neighbours and appends their records behind the old local set:
// [migrants from the left | migrants from the right | ghosts from the left | ghosts from the right].  Its last part (k_slab_retain
// for the host-mediated transports) appends copies of this rank's own OUT-migrants that landed in a neighbour's first W columns: the neighbour owns them now, and they
// are part of the ghost layer this rank keeps of that neighbour (behind the received ghosts: the same order the neighbour's own sort
// gives them, after its stayers of the same cell).  It hands all counts to the host through mapped memory -- the ONE thing the host
// waits for in the exchange, to size the sort.
struct SlabCounts {  // device memory (Control::slab_cnt)
    uint32_t out_m[2], out_g[2];  // sent: migrants / ghost-layer particles to the left | right rank
    uint32_t in_m[2], in_g[2];    // received
    uint32_t retained[2];         // own out-migrants kept as ghosts of the left | right rank
};
struct PeerCounts {  // mapped host memory
    SlabCounts cnt;
    uint32_t seq;
};
// one warp: ordered copies of the out-migrants that stay visible as ghosts, then the counts for the host (see above).
// arr.a2[0] are the positions (already advanced): the migrant's new cell column decides.
__device__ __forceinline__ void slab_retain_warp(const RecordArrays& arr, const GridParams& g, const uint32_t* __restrict__ idx_ml, const uint32_t* __restrict__ idx_mr,
                                                 SlabCounts* dc, uint32_t n_old, uint32_t cap_n, uint32_t cap_halo, uint32_t col_lo, uint32_t col_hi, uint32_t W,
                                                 uint8_t* __restrict__ pflag, PeerCounts* host_counts, uint32_t host_seq) {
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned lt = lanemask_lt();
    uint32_t dst = n_old + dc->in_m[0] + dc->in_m[1] + dc->in_g[0] + dc->in_g[1];
    uint32_t kept[2] = {0u, 0u};
    const bool sane = dc->in_m[0] + dc->in_g[0] <= cap_halo && dc->in_m[1] + dc->in_g[1] <= cap_halo && dst <= cap_n;
    for (int side = 0; side < 2 && sane; ++side) {
        const uint32_t* idx = side ? idx_mr : idx_ml;
        const uint32_t count = min(dc->out_m[side], cap_halo);
        for (uint32_t base = 0; base < count; base += 32u) {
            const uint32_t k = base + lane;
            uint32_t s = 0;
            bool keep = false;
            if (k < count) {
                s = idx[k];
                const uint32_t col = compact_1by1(position_to_cidx(g, arr.a2[0][s]));
                keep = side ? (col >= col_hi && col - col_hi < W) : (col < col_lo && col_lo - col <= W);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            const uint32_t d = dst + (uint32_t)__popc(mask & lt);
            if (keep && d < cap_n) {
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (q < arr.n2) arr.a2[q][d] = arr.a2[q][s];
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (q < arr.n1) arr.a1[q][d] = arr.a1[q][s];
                pflag[d] = 0;
            }
            dst += (uint32_t)__popc(mask);
            kept[side] += (uint32_t)__popc(mask);
        }
    }
    __syncwarp();
    if (lane == 0) {
        dc->retained[0] = kept[0];
        dc->retained[1] = kept[1];
        volatile PeerCounts* h = host_counts;
        for (int q = 0; q < 2; ++q) {
            h->cnt.out_m[q] = dc->out_m[q];
            h->cnt.out_g[q] = dc->out_g[q];
            h->cnt.in_m[q] = dc->in_m[q];
            h->cnt.in_g[q] = dc->in_g[q];
            h->cnt.retained[q] = kept[q];
        }
        __threadfence_system();
        h->seq = host_seq;
        __threadfence_system();
    }
}
__global__ void k_slab_retain(RecordArrays arr, GridParams g, const uint32_t* __restrict__ idx_ml, const uint32_t* __restrict__ idx_mr, SlabCounts* dc, uint32_t n_old,
                              uint32_t cap_n, uint32_t cap_halo, uint32_t col_lo, uint32_t col_hi, uint32_t W, uint8_t* __restrict__ pflag, PeerCounts* host_counts,
                              uint32_t host_seq) {
    slab_retain_warp(arr, g, idx_ml, idx_mr, dc, n_old, cap_n, cap_halo, col_lo, col_hi, W, pflag, host_counts, host_seq);
}
// The whole exchange of the peer-memory transport in ONE launch (the scheme of k_halo_exchange): every CTA pushes its share of the
// outgoing records into the neighbours' mailboxes, the last one to finish publishes the sequence number; then every CTA waits for the
// neighbours' numbers and pulls its share of what they stored here; CTA 0 finally keeps the out-migrants that stay visible as ghosts
// and hands the counts to the host.  Grid of at most one CTA per SM: all CTAs are resident, so a CTA that already waits for the other
// rank never keeps a CTA that still has to push from running.
struct RecordExchangeArgs {
    RecordArrays arr;
    const uint32_t *idx_ml, *idx_mr, *idx_gl, *idx_gr;  // the four selections
    const uint32_t* counts4;                            // their sizes (device)
    uint32_t cap_halo, cap_n, n_old;
    unsigned char *dst_l, *dst_r;                       // the neighbours' mailbox slots for this exchange (null: no neighbour)
    unsigned long long *pflag_l, *pflag_r;              // ... and their flags
    const unsigned char *src_l, *src_r;                 // the own mailbox slots the neighbours fill
    const unsigned long long *flag_l, *flag_r;
    unsigned long long seq;
    unsigned int* ticket;
    uint8_t* pflag;
    SlabCounts* dc;
    Control* ctl;
    GridParams g;
    uint32_t col_lo, col_hi, W;
    PeerCounts* host_counts;
    uint32_t host_seq;
};
__global__ void __launch_bounds__(256) k_records_exchange(RecordExchangeArgs a) {
    pdl_enter();
    const RecordArrays& arr = a.arr;
    {   // ---- push
        const uint32_t nml = a.dst_l ? min(a.counts4[0], a.cap_halo) : 0u, ngl = a.dst_l ? min(a.counts4[2], a.cap_halo - nml) : 0u;
        const uint32_t nmr = a.dst_r ? min(a.counts4[1], a.cap_halo) : 0u, ngr = a.dst_r ? min(a.counts4[3], a.cap_halo - nmr) : 0u;
        const uint32_t tl = nml + ngl, tr = nmr + ngr;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < tl + tr; k += gridDim.x * blockDim.x) {
            const bool right = k >= tl;
            const uint32_t kk = right ? k - tl : k, cnt = right ? tr : tl, nm = right ? nmr : nml;
            const uint32_t s = kk < nm ? (right ? a.idx_mr : a.idx_ml)[kk] : (right ? a.idx_gr : a.idx_gl)[kk - nm];
            unsigned char* buf = (right ? a.dst_r : a.dst_l) + PEER_MSG_HEADER_BYTES;
            float2* b2 = reinterpret_cast<float2*>(buf);
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < arr.n2) b2[(size_t)q * cnt + kk] = arr.a2[q][s];
            float* b1 = reinterpret_cast<float*>(buf + (size_t)arr.n2 * 8 * cnt);
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < arr.n1) b1[(size_t)q * cnt + kk] = arr.a1[q][s];
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (a.dst_l) {
                reinterpret_cast<uint32_t*>(a.dst_l)[0] = nml;
                reinterpret_cast<uint32_t*>(a.dst_l)[1] = ngl;
            }
            if (a.dst_r) {
                reinterpret_cast<uint32_t*>(a.dst_r)[0] = nmr;
                reinterpret_cast<uint32_t*>(a.dst_r)[1] = ngr;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.ticket = 0u;
            __threadfence_system();
            if (a.pflag_l) st_release_sys(a.pflag_l, a.seq);
            if (a.pflag_r) st_release_sys(a.pflag_r, a.seq);
        }
        ok = 1;
        if (a.flag_l && !peer_wait(a.flag_l, a.seq)) ok = 0;
        if (a.flag_r && !peer_wait(a.flag_r, a.seq)) ok = 0;
        if (!ok) atomicOr(&a.ctl->err_comm, 4u);
    }
    __syncthreads();
    // ---- pull
    const uint32_t nml = (ok && a.flag_l) ? __ldcg(reinterpret_cast<const uint32_t*>(a.src_l)) : 0u, ngl = (ok && a.flag_l) ? __ldcg(reinterpret_cast<const uint32_t*>(a.src_l) + 1) : 0u;
    const uint32_t nmr = (ok && a.flag_r) ? __ldcg(reinterpret_cast<const uint32_t*>(a.src_r)) : 0u, ngr = (ok && a.flag_r) ? __ldcg(reinterpret_cast<const uint32_t*>(a.src_r) + 1) : 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.dc->out_m[0] = a.counts4[0];
        a.dc->out_m[1] = a.counts4[1];
        a.dc->out_g[0] = a.counts4[2];
        a.dc->out_g[1] = a.counts4[3];
        a.dc->in_m[0] = nml;
        a.dc->in_m[1] = nmr;
        a.dc->in_g[0] = ngl;
        a.dc->in_g[1] = ngr;
    }
    const uint32_t tl = nml + ngl, tr = nmr + ngr;
    const bool room = tl <= a.cap_halo && tr <= a.cap_halo && (unsigned long long)a.n_old + tl + tr <= a.cap_n;  // else the host fails the step on the counts
    if (room) {
        const uint32_t first = a.n_old, base_g = first + nml + nmr;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < tl + tr; k += gridDim.x * blockDim.x) {
            const bool right = k >= tl;
            const uint32_t kk = right ? k - tl : k, cnt = right ? tr : tl, nm = right ? nmr : nml;
            const uint32_t dst = kk < nm ? first + (right ? nml : 0u) + kk : base_g + (right ? ngl : 0u) + (kk - nm);
            const unsigned char* buf = (right ? a.src_r : a.src_l) + PEER_MSG_HEADER_BYTES;
            const float2* b2 = reinterpret_cast<const float2*>(buf);
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < arr.n2) arr.a2[q][dst] = __ldcg(&b2[(size_t)q * cnt + kk]);
            const float* b1 = reinterpret_cast<const float*>(buf + (size_t)arr.n2 * 8 * cnt);
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < arr.n1) arr.a1[q][dst] = __ldcg(&b1[(size_t)q * cnt + kk]);
            a.pflag[dst] = 0;
        }
    }
    // ---- the out-migrants that stay here as ghosts go behind everything received (their places depend on the counts only)
    if (blockIdx.x == 0) {
        __syncthreads();  // thread 0's counts in a.dc
        if (threadIdx.x < 32) slab_retain_warp(arr, a.g, a.idx_ml, a.idx_mr, a.dc, a.n_old, a.cap_n, a.cap_halo, a.col_lo, a.col_hi, a.W, a.pflag, a.host_counts, a.host_seq);
    }
}

// one warp; boxes[r] = rank r's mailbox as mapped here (boxes[rank] = the own one); value in place at dev_ptr.
// After: a functor the reduced value is handed to on the spot (the Jacobi loop decision, sweeps.cuh), saving its own launch.
struct NoAfter {
    __device__ __forceinline__ void operator()(Control*) const {}
};
template <class After>
__global__ void k_allreduce_peer(void* dev_ptr, int is_double_sum, PeerBoxHeader* const* __restrict__ boxes, int rank, int world, unsigned long long seq,
                                 Control* ctl, After after) {
    pdl_enter();
    const int lane = threadIdx.x;
    const unsigned parity = (unsigned)(seq & 1ull);
    const double mine = is_double_sum ? *reinterpret_cast<const double*>(dev_ptr) : (double)*reinterpret_cast<const float*>(dev_ptr);
    if (lane < world) {
        PeerBoxHeader* b = boxes[lane];
        reinterpret_cast<volatile double*>(b->ar_val[parity])[rank] = mine;
        __threadfence_system();
        st_release_sys(&b->ar_flag[rank], seq);
    }
    __syncwarp();
    PeerBoxHeader* me = boxes[rank];
    bool ok = true;
    if (lane < world) ok = peer_wait(&me->ar_flag[lane], seq);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) {
            atomicOr(&ctl->err_comm, 2u);
        } else {
            double acc = __ldcg(&me->ar_val[parity][0]);
            for (int r = 1; r < world; ++r) {
                const double v = __ldcg(&me->ar_val[parity][r]);
                acc = is_double_sum ? acc + v : fmax(acc, v);
            }
            if (is_double_sum)
                *reinterpret_cast<double*>(dev_ptr) = acc;
            else
                *reinterpret_cast<float*>(dev_ptr) = (float)acc;
            after(ctl);
        }
    }
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
