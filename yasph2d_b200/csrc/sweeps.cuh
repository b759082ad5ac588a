// sweeps.cuh -- every neighbour-dependent per-particle pass of the WCSPH / DFSPH solvers as one tile-staged kernel.
//
// Skeleton (k_sweep): persistent CTAs; a CTA starts with tile blockIdx.x and takes every further tile from a device-wide queue
// (one atomic per tile, fetched a tile ahead), so CTAs on slower SMs or with heavier tiles simply take fewer.  A tile's own particles
// plus the 1-cell apron are staged into shared memory (positions, up to two per-pass payload arrays, boundary positions)
// with cp.async, DOUBLE-BUFFERED: while the CTA computes tile k from buffer k&1 the copies for tile k+1 are in flight into
// the other buffer and the copy-run table of tile k+2 is on its way through registers, so the global-memory latency of the
// staging never sits between two barriers (one __syncthreads per tile).  One thread per particle then walks its compact
// neighbour list (u16 shared-memory slots, dynamic first, then static) in the reference's order; the four slots of an
// 8-byte list word are evaluated branch-free (select on the accumulator) so the compiler interleaves four independent
// pair evaluations.  Shared memory is sized from the largest tile of the current neighbourhood structure (Control::max_*).
// Reductions (Jacobi residual sum, CFL max) are fused: per-thread accumulation, block reduce, per-CTA partial, the
// last-arriving CTA combines the partials in fixed order and takes the convergence decision on the device (no host round
// trip inside a Jacobi iteration).
//
// Pass -> reference:
//   OpDensityAlpha   FluidParticleWorld::update_densities (fluidparticleworld.rs:197-231) fused with
//                    DFSPHSolver::compute_alpha_factors (dfsph.rs:68-97); WCSPH: + Tait pressure (wscsph.rs:52-57)
//   OpViscosity      non-pressure forces (dfsph.rs:436-469) + max |v + a dt|^2 (dfsph.rs:474-477)
//   OpJacobiA        compute_density_error (dfsph.rs:99-126) / compute_density_change (dfsph.rs:249-280) + residual sum
//                    and loop decision (dfsph.rs:219-245, 374-400); writes k_i = err_i * alpha_i (dfsph.rs:141 / 295)
//   OpJacobiB        correct_velocity_with_density_error (dfsph.rs:128-161), ..._divergence_error (dfsph.rs:282-314) and the
//                    two warm starts (dfsph.rs:163-193, 316-344) incl. the clamp (dfsph.rs:201-203, 356-358)
//   OpWcsphAccel     WCSPHSolver::update_accellerations (wscsph.rs:59-118) + CFL max (wscsph.rs:160-163)
#pragma once
#include "neighborhood.cuh"

namespace yasph {

#ifndef YASPH_SWEEP_PRODUCERS
#define YASPH_SWEEP_PRODUCERS 2
#endif
#ifndef YASPH_SWEEP_CONSUMERS
#define YASPH_SWEEP_CONSUMERS 8
#endif
constexpr int SW_PRODUCER_WARPS = YASPH_SWEEP_PRODUCERS;  // warps 0.. stage tiles into shared memory
constexpr int SW_CONSUMER_WARPS = YASPH_SWEEP_CONSUMERS;  // the other warps compute
constexpr int SW_THREADS = 32 * (SW_PRODUCER_WARPS + SW_CONSUMER_WARPS);
#ifndef YASPH_SWEEP_MIN_CTAS
#define YASPH_SWEEP_MIN_CTAS 3  // caps the registers so that three CTAs share an SM
#endif
#ifndef YASPH_SWEEP_PSLEEP
#define YASPH_SWEEP_PSLEEP 200  // ns between a producer's polls of an empty barrier
#endif
#ifndef YASPH_SWEEP_CSLEEP
#define YASPH_SWEEP_CSLEEP 40   // ns between a consumer's polls of a full barrier
#endif
#ifndef YASPH_SWEEP_STAGES
#define YASPH_SWEEP_STAGES 3
#endif
constexpr int SW_APRON_PER_LANE = (APRON_TABLE + 32 * SW_PRODUCER_WARPS - 1) / (32 * SW_PRODUCER_WARPS);  // table entries a producer lane prefetches
constexpr int SW_STAGES = YASPH_SWEEP_STAGES;  // stages of the shared-memory ring (fewer at run time when tiles are very large)
static_assert(2 * YASPH_SWEEP_STAGES * 8 <= 96, "barriers and the tile hand-over words share the 128-byte head");
constexpr uint32_t SW_NK_DONE = 0xFFFFFFFEu;  // stage marker: no more tiles
constexpr uint32_t SW_MAX_STAGED_WORDS = 8;  // list words per particle staged in shared memory (the rest, if any, is read from global memory)

struct SweepCommon {
    TileTables tt;
    const uint32_t* tile_nk;  // [tile] most list words of any particle of the tile
    const uint32_t* apron_idx;  // [tile][APRON_TABLE] global index of the tile's first apron slots (written by the list build)
    const unsigned long long* lists;
    const uint32_t* counts;   // per particle: count_dynamic | count_total << 8
    const float2* pos;
    const float2* bpos;
    Control* ctl;
    KernelConsts kc;
    uint32_t cap_dyn, cap_stat;  // staged candidates of the largest tile (multiples of 16)
    uint32_t cap_pc;             // particles of the largest tile (multiple of 16)
    uint32_t nk_stage;           // list words per particle that fit the stage (<= SW_MAX_STAGED_WORDS)
    uint32_t nstages;            // ring depth, 1..SW_STAGES
    uint32_t n;
    float mass, rho0;
    double* partials;  // [gridDim.x]
    double* partials_unstaged;      // [n_partials_unstaged] partial sums / maxima of k_sweep_unstaged (tiles too large to stage)
    uint32_t n_partials_unstaged;   // 0 unless that kernel ran ahead of this pass's k_sweep
    // slab mode (multi-GPU): ghosts are excluded from the reductions, the residual average runs over the global particle count
    // and the convergence decision is taken after the all-reduce (k_jacobi_decide)
    const uint8_t* ghost;  // null on a single GPU
#ifdef YASPH_SWEEP_TIMING
    unsigned long long* dbg;  // [8] cycle counters (profiling builds only)
#endif
    float n_avg;           // particle count of dfsph.rs:221,376 as f32
};

enum ReduceKind { REDUCE_NONE = 0, REDUCE_SUM = 1, REDUCE_MAX = 2 };

struct NoPay {};

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX) ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// the calling thread's earlier cp.async copies arrive on the barrier when they have landed (does not raise the pending count)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(void* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
// A waiting warp sleeps between polls so that its polling does not take issue slots from the warps that compute.
template <unsigned SLEEP_NS>
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(SLEEP_NS);
}
// 1-D bulk copy global -> shared through the TMA unit; completion is counted in bytes on the barrier.  16-byte aligned, 16 | bytes.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// ---- shared-memory layout -------------------------------------------------------------------------------------------------
// [ full[SW_STAGES] | empty[SW_STAGES] mbarriers ][ TileRuns x2 (producer) ][ stage 0 ][ stage 1 ] ...
// stage: TileHeader | nk | float2 pos[cap_dyn] | P0[cap_dyn] | P1[cap_dyn] | float2 stat[cap_stat] | u32 cnt[cap_pc] | O0[cap_pc] |
//        O1[cap_pc] | u64 lists[nk_stage * cap_pc]
template <class Op>
struct SweepLayout {
    static constexpr size_t P0 = Op::NPAY >= 1 ? sizeof(typename Op::P0) : 0;
    static constexpr size_t P1 = Op::NPAY >= 2 ? sizeof(typename Op::P1) : 0;
    static constexpr size_t O0 = Op::NOWN >= 1 ? sizeof(typename Op::O0) : 0;
    static constexpr size_t O1 = Op::NOWN >= 2 ? sizeof(typename Op::O1) : 0;
    static constexpr size_t HEAD = 128;                                   // barriers, then (offset 96) the producers' tile hand-over words
    static constexpr size_t RUNS = 2 * SW_PRODUCER_WARPS * sizeof(TileRuns);  // every producer warp's copy-run tables
    static constexpr size_t STAGE_HDR = 48;                               // TileHeader + nk, padded to 16
    __host__ __device__ static size_t stage_bytes(uint32_t cap_dyn, uint32_t cap_stat, uint32_t cap_pc, uint32_t nk_stage) {
        return STAGE_HDR + (size_t)cap_dyn * (sizeof(float2) + P0 + P1) + (Op::USES_STATIC ? (size_t)cap_stat * sizeof(float2) : 0) +
               (size_t)cap_pc * (4 + O0 + O1) + (size_t)nk_stage * cap_pc * 8;
    }
    static constexpr size_t TAIL = Op::WARP_TAIL ? sizeof(RadixHistSmem) : 0;  // behind the stages: the radix digit histograms of a pass that also generates sort keys
    __host__ __device__ static size_t total_bytes(uint32_t cap_dyn, uint32_t cap_stat, uint32_t cap_pc, uint32_t nk_stage, uint32_t nstages) {
        return HEAD + RUNS + nstages * stage_bytes(cap_dyn, cap_stat, cap_pc, nk_stage) + TAIL;
    }
};
template <class Op>
struct SweepStage {
    TileHeader* hdr;
    uint32_t* nk;
    float2* pos;
    typename Op::P0* p0;
    typename Op::P1* p1;
    float2* stat;
    uint32_t* cnt;
    typename Op::O0* o0;
    typename Op::O1* o1;
    unsigned long long* lists;
    __device__ __forceinline__ SweepStage(unsigned char* b, uint32_t cap_dyn, uint32_t cap_stat, uint32_t cap_pc) {
        typedef SweepLayout<Op> L;
        hdr = reinterpret_cast<TileHeader*>(b);
        nk = reinterpret_cast<uint32_t*>(b + sizeof(TileHeader));
        b += L::STAGE_HDR;
        pos = reinterpret_cast<float2*>(b);
        b += (size_t)cap_dyn * sizeof(float2);
        p0 = reinterpret_cast<typename Op::P0*>(b);
        b += (size_t)cap_dyn * L::P0;
        p1 = reinterpret_cast<typename Op::P1*>(b);
        b += (size_t)cap_dyn * L::P1;
        stat = reinterpret_cast<float2*>(b);
        b += Op::USES_STATIC ? (size_t)cap_stat * sizeof(float2) : 0;
        cnt = reinterpret_cast<uint32_t*>(b);
        b += (size_t)cap_pc * 4;
        o0 = reinterpret_cast<typename Op::O0*>(b);
        b += (size_t)cap_pc * L::O0;
        o1 = reinterpret_cast<typename Op::O1*>(b);
        b += (size_t)cap_pc * L::O1;
        lists = reinterpret_cast<unsigned long long*>(b);
    }
};
// list words per particle that fit a stage when the whole kernel may use `budget` bytes of shared memory
template <class Op>
inline uint32_t sweep_nk_stage(uint32_t cap_dyn, uint32_t cap_stat, uint32_t cap_pc, uint32_t nk_max, uint32_t nstages, size_t budget) {
    uint32_t nk = nk_max < SW_MAX_STAGED_WORDS ? nk_max : SW_MAX_STAGED_WORDS;
    while (nk > 0 && SweepLayout<Op>::total_bytes(cap_dyn, cap_stat, cap_pc, nk, nstages) > budget) --nk;
    return nk;
}

// slot of entry q of a list word, from constant shifts
__device__ __forceinline__ uint32_t word_slot(unsigned long long w, uint32_t q) {
    const uint32_t half = q < 2 ? (uint32_t)w : (uint32_t)(w >> 32);
    return (q & 1u) ? (half >> 16) : (half & 0xFFFFu);
}

// ---- producer warps: stage one tile ---------------------------------------------------------------------------------------
// Contiguous pieces travel as 16-byte aligned bulk copies (TMA unit, one instruction each, issued by one lane): the own
// particles' positions and payloads into their slots (TileHeader: pad slots absorb the alignment surplus), their counts and
// per-particle operands, and the tile's block of list words.  Only the apron (candidates from the eight surrounding tiles,
// ~36 short runs) and the boundary candidates are copied element-wise with cp.async, spread over the producer warps.
template <class Op>
__device__ __forceinline__ void sweep_stage_tile(const SweepCommon& c, const Op& op, const TileRuns& tr, uint32_t nk_tile, const SweepStage<Op>& st, void* full_bar,
                                                 uint32_t pw, const uint32_t (&pre_ap)[SW_APRON_PER_LANE]
#ifdef YASPH_SWEEP_TIMING
                                                 , long long* tsplit
#endif
                                                 ) {
#ifdef YASPH_SWEEP_TIMING
    long long z0 = clock64();
#endif
    typedef SweepLayout<Op> L;
    const uint32_t lane = lane_id();
    const uint32_t first = pw * 32u + lane, stride = 32u * SW_PRODUCER_WARPS;  // the producer warps interleave
    const TileHeader h = tr.hdr;
    const uint32_t d = h.pstart & 3u;
    const uint32_t nel = (d + h.pcount + 3u) & ~3u;  // elements of the aligned own range [pstart - d, roundup4(pstart + pcount))
    const bool fits = h.dyn_total <= c.cap_dyn && h.stat_total <= c.cap_stat && nel <= c.cap_pc;  // always: capacities are the maxima over all tiles
    const uint32_t nk = nk_tile < c.nk_stage ? nk_tile : c.nk_stage;
    if (lane == 0 && pw == 0) {
        *st.hdr = h;
        *st.nk = fits ? nk : 0xFFFFFFFFu;  // a tile too large to stage is left to k_sweep_unstaged
    }
    if (fits) {
        // apron: the first APRON_TABLE slots come with their global index from the list build's table (prefetched into
        // registers one tile ahead), the rest -- very full aprons only -- through the copy-run table
        const uint32_t na = tile_apron_count(h);
#pragma unroll
        for (int u = 0; u < SW_APRON_PER_LANE; ++u) {
            const uint32_t a = first + (uint32_t)u * stride;
            if (a < na && a < APRON_TABLE) {
                const uint32_t s = tile_apron_slot(h, a), g = pre_ap[u];
                cp_async<8>(&st.pos[s], &c.pos[g]);
                if constexpr (Op::NPAY >= 1) cp_async<sizeof(typename Op::P0)>(&st.p0[s], &op.pay0()[g]);
                if constexpr (Op::NPAY >= 2) cp_async<sizeof(typename Op::P1)>(&st.p1[s], &op.pay1()[g]);
            }
        }
        for (uint32_t a = (uint32_t)SW_APRON_PER_LANE * stride + first; a < na; a += stride) {
            const uint32_t s = tile_apron_slot(h, a);
            const uint32_t g = run_slot_to_global(tr.rd, s);
            cp_async<8>(&st.pos[s], &c.pos[g]);
            if constexpr (Op::NPAY >= 1) cp_async<sizeof(typename Op::P0)>(&st.p0[s], &op.pay0()[g]);
            if constexpr (Op::NPAY >= 2) cp_async<sizeof(typename Op::P1)>(&st.p1[s], &op.pay1()[g]);
        }
#ifdef YASPH_SWEEP_TIMING
        tsplit[0] += clock64() - z0;
        z0 = clock64();
#endif
        if (Op::USES_STATIC)
            for (uint32_t s = first; s < h.stat_total; s += stride) cp_async<8>(&st.stat[s], &c.bpos[run_slot_to_global(tr.rs, s)]);
    }
#ifdef YASPH_SWEEP_TIMING
    tsplit[1] += clock64() - z0;
    z0 = clock64();
#endif
    cp_async_mbar_arrive_noinc(full_bar);  // 32 arrivals per producer warp: this lane's copies have landed
    __syncwarp();                          // lane 0's header stores are ordered before its arrival below
    // The bulk copies: lane 0 of producer warp 0 announces the byte count with its arrival; the copies themselves are dealt to
    // the first lanes of ALL producer warps (each is a separately issued instruction, so only different warps overlap them).
    // The phase cannot complete before lane 0's arrival, hence a copy that lands before the byte count is announced is fine.
    if (fits) {
        const uint32_t lbytes = (nk * h.pcount * 8u + 15u) & ~15u;  // the first nk list words of every particle: one contiguous block
        const size_t g0 = h.pstart - d;
        const uint32_t s0 = h.own_lo - d;
        if (pw == 0 && lane == 0) mbar_arrive_expect_tx(full_bar, nel * (uint32_t)(sizeof(float2) + L::P0 + L::P1 + 4 + L::O0 + L::O1) + lbytes);
        // copy j goes to warp j % SW_PRODUCER_WARPS, lane j / SW_PRODUCER_WARPS
        auto mine = [&](uint32_t j) { return pw == j % SW_PRODUCER_WARPS && lane == j / SW_PRODUCER_WARPS; };
        if (mine(0)) bulk_copy_g2s(st.pos + s0, c.pos + g0, nel * (uint32_t)sizeof(float2), full_bar);
        if (mine(1) && lbytes) bulk_copy_g2s(st.lists, c.lists + (size_t)h.pstart * LIST_WORDS, lbytes, full_bar);
        if (mine(2)) bulk_copy_g2s(st.cnt, c.counts + g0, nel * 4u, full_bar);
        if constexpr (Op::NPAY >= 1)
            if (mine(3)) bulk_copy_g2s(st.p0 + s0, op.pay0() + g0, nel * (uint32_t)L::P0, full_bar);
        if constexpr (Op::NOWN >= 1)
            if (mine(4)) bulk_copy_g2s(st.o0, op.own0() + g0, nel * (uint32_t)L::O0, full_bar);
        if constexpr (Op::NPAY >= 2)
            if (mine(5)) bulk_copy_g2s(st.p1 + s0, op.pay1() + g0, nel * (uint32_t)L::P1, full_bar);
        if constexpr (Op::NOWN >= 2)
            if (mine(6)) bulk_copy_g2s(st.o1, op.own1() + g0, nel * (uint32_t)L::O1, full_bar);
    } else if (pw == 0 && lane == 0) {
        mbar_arrive(full_bar);
    }
}

// One particle of a tile that is too large to stage (more candidates than the shared memory of an SM holds -- a dense cluster, as
// a diverging run produces): everything comes from global memory, neighbour slots are translated to global indices through the
// tile's copy runs.  Slow and correct, in the same order as the staged path, so the results are the same bits.
template <class Op>
__device__ __forceinline__ double sweep_particle_unstaged(const SweepCommon& c, const Op& op, uint32_t t, const TileHeader& h, uint32_t tl) {
    typedef typename Op::P0 P0;
    typedef typename Op::P1 P1;
    const TileRuns* tr = c.tt.runs + t;
    const uint32_t i = h.pstart + tl;
    const float2 pi = c.pos[i];
    P0 s0;
    P1 s1;
    if constexpr (Op::NPAY >= 1) s0 = op.pay0()[i];
    if constexpr (Op::NPAY >= 2) s1 = op.pay1()[i];
    const uint32_t cnt = c.counts[i] & 0xFFFFu;
    const uint32_t cd = cnt & 0xFFu, ct = (cnt >> 8) & 0xFFu;
    typename Op::O0 w0;
    typename Op::O1 w1;
    if constexpr (Op::NOWN >= 1) w0 = op.own0()[i];
    if constexpr (Op::NOWN >= 2) w1 = op.own1()[i];
    typename Op::Acc acc;
    const bool active = op.init(c, acc, i, pi, s0, s1, ct);
    const uint32_t nkd = (cd + 3u) >> 2;
    if (active) {
        for (uint32_t e = 0; e < cd; ++e) {  // the padding entries of the last dynamic word are simply not visited
            const unsigned long long w = c.lists[list_word_index(h.pstart, h.pcount, e >> 2, tl)];
            const uint32_t g = dyn_slot_to_global(*tr, unpack_slot(w, e));
            P0 n0;
            P1 n1;
            if constexpr (Op::NPAY >= 1) n0 = op.pay0()[g];
            if constexpr (Op::NPAY >= 2) n1 = op.pay1()[g];
            op.dyn(c, acc, pi, s0, s1, c.pos[g], n0, n1);
        }
        if (Op::USES_STATIC) {
            for (uint32_t e = 0; e < ct - cd; ++e) {
                const unsigned long long w = c.lists[list_word_index(h.pstart, h.pcount, nkd + (e >> 2), tl)];
                op.stat(c, acc, pi, s0, s1, c.bpos[run_slot_to_global(tr->rs, unpack_slot(w, e))]);
            }
        }
    }
    double r = op.finish(c, acc, i, pi, s0, s1, active, w0, w1);
    if (Op::REDUCE != REDUCE_NONE && c.ghost != nullptr && c.ghost[i]) r = 0.0;
    return r;
}

// the staging capacity test of the tile kernels (sweep_stage_tile applies the same one)
__device__ __forceinline__ bool sweep_tile_fits(const SweepCommon& c, const TileHeader& h) {
    const uint32_t nel = ((h.pstart & 3u) + h.pcount + 3u) & ~3u;
    return h.dyn_total <= c.cap_dyn && h.stat_total <= c.cap_stat && nel <= c.cap_pc;
}
// The tiles k_sweep leaves out.  Launched BEFORE k_sweep<Op> of the same pass (only when such tiles exist): its per-CTA partial
// sums / maxima are combined with k_sweep's own by k_sweep's last CTA.
constexpr int SWU_THREADS = 256;
template <class Op>
__global__ void __launch_bounds__(SWU_THREADS) k_sweep_unstaged(SweepCommon c, Op op) {
    if (op.skip(c.ctl)) return;
    op.prepare(c);
    const uint32_t ntiles = c.ctl->num_tiles;
    double racc = 0.0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const TileHeader h = c.tt.runs[t].hdr;
        if (sweep_tile_fits(c, h)) continue;
        for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += SWU_THREADS) {
            const double r = sweep_particle_unstaged<Op>(c, op, t, h, tl);
            if (Op::REDUCE == REDUCE_SUM) racc += r;
            if (Op::REDUCE == REDUCE_MAX) racc = fmax(racc, r);
        }
    }
    if (Op::REDUCE != REDUCE_NONE) {
        __shared__ double wred[SWU_THREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, racc, o);
            racc = Op::REDUCE == REDUCE_SUM ? racc + u : fmax(racc, u);
        }
        if (lane_id() == 0) wred[threadIdx.x >> 5] = racc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = wred[0];
            for (int w = 1; w < SWU_THREADS / 32; ++w) tsum = Op::REDUCE == REDUCE_SUM ? tsum + wred[w] : fmax(tsum, wred[w]);
            c.partials_unstaged[blockIdx.x] = tsum;
        }
    }
}

#ifdef YASPH_SWEEP_TIMING
__device__ __forceinline__ unsigned long long global_timer_ns_sweep() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif
template <class Op>
__global__ void __launch_bounds__(SW_THREADS, YASPH_SWEEP_MIN_CTAS) k_sweep(SweepCommon c, Op op) {
    typedef typename Op::P0 P0;
    typedef typename Op::P1 P1;
    typedef SweepLayout<Op> L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* empty_bar = full_bar + SW_STAGES;
    const uint32_t NS = c.nstages;
    TileRuns* runs = reinterpret_cast<TileRuns*>(smem_raw + L::HEAD) + 2 * (threadIdx.x >> 5);  // [2] per producer warp
    const size_t sbytes = L::stage_bytes(c.cap_dyn, c.cap_stat, c.cap_pc, c.nk_stage);
    unsigned char* stage0 = smem_raw + L::HEAD + L::RUNS;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], 32 * SW_PRODUCER_WARPS + 1);  // cp.async arrivals of every producer lane + the bulk-copy / plain arrival
            mbar_init(&empty_bar[s], SW_CONSUMER_WARPS);  // every consumer warp releases the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    RadixHistSmem* const hist = Op::WARP_TAIL ? reinterpret_cast<RadixHistSmem*>(stage0 + NS * sbytes) : nullptr;
    if (Op::WARP_TAIL)
        for (uint32_t q = threadIdx.x; q < RS_PASSES * RS_BINS; q += SW_THREADS) (&hist->h[0][0])[q] = 0u;
    pdl_enter();  // the barriers above are set up while the previous kernel of the stream drains
    if (op.skip(c.ctl)) return;
#ifdef YASPH_SWEEP_TIMING
    if (threadIdx.x == 0 && c.dbg) atomicMin(&c.dbg[8], global_timer_ns_sweep());
#endif
    op.prepare(c);
    __syncthreads();
    const uint32_t ntiles = c.ctl->num_tiles;
    const uint32_t G = gridDim.x;
    double racc = 0.0;
    if (warp < SW_PRODUCER_WARPS) {
        // ---------------- producers ----------------
        uint4 pre[2];
        uint32_t pre_nk = 0;
        const uint32_t NV = sizeof(TileRuns) / 16;
        static_assert(sizeof(TileRuns) / 16 <= 64, "two uint4 per lane");
        uint32_t pre_ap[SW_APRON_PER_LANE], cur_ap[SW_APRON_PER_LANE];
        auto prefetch = [&](uint32_t t) {
            if (t < ntiles) {
                const uint4* src = reinterpret_cast<const uint4*>(c.tt.runs + t);
                if (lane < NV) pre[0] = src[lane];
                if (lane + 32 < NV) pre[1] = src[lane + 32];
                pre_nk = c.tile_nk[t];
#pragma unroll
                for (int u = 0; u < SW_APRON_PER_LANE; ++u) {
                    const uint32_t a = warp * 32u + lane + (uint32_t)u * (32u * SW_PRODUCER_WARPS);
                    pre_ap[u] = a < APRON_TABLE ? c.apron_idx[(size_t)t * APRON_TABLE + a] : 0u;
                }
            }
        };
        // tile queue: this CTA's first tile is blockIdx.x, the others come from ctl->tile_next.  The leader lane fetches the ticket
        // for the tile after next while the current one is staged (the atomic's round trip hides behind the staging) and hands
        // it to the other producer lanes through shared memory at the producers' own barrier.
        volatile uint32_t* next_tile = reinterpret_cast<volatile uint32_t*>(smem_raw + 96);
        const bool leader = threadIdx.x == 0;
        auto producers_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(32 * SW_PRODUCER_WARPS) : "memory"); };
        uint32_t t_cur = blockIdx.x, ticket = 0;
        prefetch(t_cur);  // in flight while the leader fetches the first ticket
        if (leader) next_tile[0] = G + atomicAdd(&c.ctl->tile_next, 1u);
        producers_sync();
        uint32_t t_nxt = next_tile[0];
        uint32_t k = 0, stage = 0, round = 0;  // round: completed passes over the ring
#ifdef YASPH_SWEEP_TIMING
        long long t_wait = 0, t_issue = 0, t_total = clock64(), n_tiles = 0, tsplit[2] = {0, 0};
#endif
        for (;; ++k) {
            if (t_cur >= ntiles) {  // the queue is empty: tell the consumers through one more stage
                if (round) mbar_wait<YASPH_SWEEP_PSLEEP>(&empty_bar[stage], (round - 1u) & 1u);
                const SweepStage<Op> st(stage0 + stage * sbytes, c.cap_dyn, c.cap_stat, c.cap_pc);
                if (leader) *st.nk = SW_NK_DONE;
                cp_async_mbar_arrive_noinc(&full_bar[stage]);
                __syncwarp();
                if (leader) mbar_arrive(&full_bar[stage]);
                break;
            }
            if (leader) ticket = G + atomicAdd(&c.ctl->tile_next, 1u);  // the tile after next; consumed at the end of this iteration
            TileRuns& tr = runs[k & 1u];
            uint4* dst = reinterpret_cast<uint4*>(&tr);
            if (lane < NV) dst[lane] = pre[0];
            if (lane + 32 < NV) dst[lane + 32] = pre[1];
            const uint32_t nk_tile = pre_nk;
#pragma unroll
            for (int u = 0; u < SW_APRON_PER_LANE; ++u) cur_ap[u] = pre_ap[u];
            __syncwarp();
            prefetch(t_nxt);  // in flight while this tile is staged
#ifdef YASPH_SWEEP_TIMING
            long long q0 = clock64();
#endif
            if (round) mbar_wait<YASPH_SWEEP_PSLEEP>(&empty_bar[stage], (round - 1u) & 1u);  // every consumer warp has released the stage
            const SweepStage<Op> st(stage0 + stage * sbytes, c.cap_dyn, c.cap_stat, c.cap_pc);
#ifdef YASPH_SWEEP_TIMING
            long long q1 = clock64();
#endif
#ifdef YASPH_SWEEP_TIMING
            sweep_stage_tile(c, op, tr, nk_tile, st, &full_bar[stage], warp, cur_ap, tsplit);
#else
            sweep_stage_tile(c, op, tr, nk_tile, st, &full_bar[stage], warp, cur_ap);
#endif
#ifdef YASPH_SWEEP_TIMING
            t_wait += q1 - q0;
            t_issue += clock64() - q1;
            ++n_tiles;
#endif
            if (++stage == NS) {
                stage = 0;
                ++round;
            }
            if (leader) next_tile[(k + 1u) & 1u] = ticket;
            producers_sync();
            t_cur = t_nxt;
            t_nxt = next_tile[(k + 1u) & 1u];
        }
#ifdef YASPH_SWEEP_TIMING
        if (lane == 0 && c.dbg) {
            atomicAdd(&c.dbg[2], (unsigned long long)(clock64() - t_total));
            atomicAdd(&c.dbg[3], (unsigned long long)t_wait);
            atomicAdd(&c.dbg[4], (unsigned long long)t_issue);
            atomicAdd(&c.dbg[5], (unsigned long long)n_tiles);
        }
#endif
    } else {
        // ---------------- consumers ----------------
        const uint32_t cw = warp - SW_PRODUCER_WARPS;
        uint32_t chunk_base = 0;  // chunks (32 particles) of all earlier tiles of this CTA: chunks are dealt round-robin to the warps
        uint32_t stage = 0, round = 0;
#ifdef YASPH_SWEEP_TIMING
        long long t_wait = 0, t_total = clock64(), n_chunks = 0;
#endif
        for (;;) {
            const SweepStage<Op> st(stage0 + stage * sbytes, c.cap_dyn, c.cap_stat, c.cap_pc);
#ifdef YASPH_SWEEP_TIMING
            long long q0 = clock64();
#endif
            mbar_wait<YASPH_SWEEP_CSLEEP>(&full_bar[stage], round & 1u);
#ifdef YASPH_SWEEP_TIMING
            t_wait += clock64() - q0;
#endif
            const uint32_t nk_st = *st.nk;
            if (nk_st == SW_NK_DONE) break;  // the producers found the tile queue empty
            const TileHeader h = *st.hdr;
            const uint32_t nchunks = (h.pcount + 31u) >> 5;
            const uint32_t od = h.pstart & 3u;  // the per-particle operand arrays are staged from the 4-element-aligned index below pstart
            if (nk_st != 0xFFFFFFFFu) {
                for (uint32_t j = (cw + SW_CONSUMER_WARPS - chunk_base % SW_CONSUMER_WARPS) % SW_CONSUMER_WARPS; j < nchunks; j += SW_CONSUMER_WARPS) {
                    const uint32_t wo = j * 32u + lane;  // particle of the tile (a per-tile order by list length was measured to change nothing)
                    uint32_t tail_key = 0u;  // Op::WARP_TAIL: what the particle hands to the warp-collective step behind the block below
                    if (wo < h.pcount) {
                        const uint32_t tl = wo;
                        const uint32_t i = h.pstart + tl;
                        const uint32_t own = h.own_lo + tl;
                        const float2 pi = st.pos[own];
                        P0 s0;
                        P1 s1;
                        if (Op::NPAY >= 1) s0 = st.p0[own];
                        if (Op::NPAY >= 2) s1 = st.p1[own];
                        const uint32_t cnt = st.cnt[od + tl] & 0xFFFFu;
                        const uint32_t cd = cnt & 0xFFu, ct = (cnt >> 8) & 0xFFu;
                        typename Op::O0 w0;
                        typename Op::O1 w1;
                        if (Op::NOWN >= 1) w0 = st.o0[od + tl];
                        if (Op::NOWN >= 2) w1 = st.o1[od + tl];
                        typename Op::Acc acc;
                        const bool active = op.init(c, acc, i, pi, s0, s1, ct);
                        const uint32_t nkd = (cd + 3u) >> 2;
                        if (active) {
                            // dynamic neighbours: nkd words of four slots; the last word is padded with the particle's own slot, whose
                            // pair contributes an exact zero unless the pass says otherwise (PAD_IS_ZERO == false: density)
                            for (uint32_t kb = 0; kb < nkd; ++kb) {
                                const unsigned long long w = kb < nk_st ? st.lists[kb * h.pcount + tl] : c.lists[list_word_index(h.pstart, h.pcount, kb, tl)];
#pragma unroll
                                for (uint32_t q = 0; q < 4; ++q) {
                                    const uint32_t slot = word_slot(w, q);
                                    P0 n0;
                                    P1 n1;
                                    if (Op::NPAY >= 1) n0 = st.p0[slot];
                                    if (Op::NPAY >= 2) n1 = st.p1[slot];
                                    if (Op::PAD_IS_ZERO) {
                                        op.dyn(c, acc, pi, s0, s1, st.pos[slot], n0, n1);
                                    } else {
                                        typename Op::Acc trial = acc;
                                        op.dyn(c, trial, pi, s0, s1, st.pos[slot], n0, n1);
                                        if (kb * 4 + q < cd) acc = trial;
                                    }
                                }
                            }
                            // static neighbours: ct - cd entries from word nkd on (only near boundaries)
                            if (Op::USES_STATIC) {
                                for (uint32_t e = 0; e < ct - cd; ++e) {
                                    const uint32_t kb = nkd + (e >> 2);
                                    const unsigned long long w = kb < nk_st ? st.lists[kb * h.pcount + tl] : c.lists[list_word_index(h.pstart, h.pcount, kb, tl)];
                                    op.stat(c, acc, pi, s0, s1, st.stat[unpack_slot(w, e)]);
                                }
                            }
                        }
                        double r;
                        if constexpr (Op::WARP_TAIL)
                            r = op.finish_tail(c, acc, i, pi, s0, s1, active, w0, w1, tail_key);
                        else
                            r = op.finish(c, acc, i, pi, s0, s1, active, w0, w1);
                        if (Op::REDUCE != REDUCE_NONE && c.ghost != nullptr && c.ghost[i]) r = 0.0;
                        if (Op::REDUCE == REDUCE_SUM) racc += r;
                        if (Op::REDUCE == REDUCE_MAX) racc = fmax(racc, r);
                    }
                    if constexpr (Op::WARP_TAIL)
                        if (op.tail_on()) radix_hist_add(*hist, tail_key, wo < h.pcount);  // all 32 lanes: the sort's digit histograms
                }
            }
            chunk_base += nchunks;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);  // this warp is done with the stage
#ifdef YASPH_SWEEP_TIMING
            n_chunks += (nchunks + SW_CONSUMER_WARPS - 1 - ((cw + SW_CONSUMER_WARPS - (chunk_base - nchunks) % SW_CONSUMER_WARPS) % SW_CONSUMER_WARPS)) / SW_CONSUMER_WARPS;
#endif
            if (++stage == NS) {
                stage = 0;
                ++round;
            }
        }
#ifdef YASPH_SWEEP_TIMING
        if (lane == 0 && c.dbg) {
            atomicAdd(&c.dbg[0], (unsigned long long)(clock64() - t_total));
            atomicAdd(&c.dbg[1], (unsigned long long)t_wait);
            atomicAdd(&c.dbg[6], (unsigned long long)n_chunks);
            atomicMax(&c.dbg[7], (unsigned long long)(clock64() - t_total));  // longest-lived consumer warp of any launch since the reset
        }
#endif
    }
    // the last CTA to get here re-arms the tile queue for the next sweep (nobody fetches a ticket any more: every CTA's
    // producers have seen the queue empty before their CTA counts itself done)
    // (a reducing sweep does this with the ticket of its reduction below: one atomic round trip less at the end of the launch)
    __syncthreads();
    if constexpr (Op::WARP_TAIL)
        if (op.tail_on()) radix_hist_flush(*hist, op.tail_scratch());
    if (Op::REDUCE == REDUCE_NONE && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&c.ctl->cta_done, 1u) == gridDim.x - 1) {
            c.ctl->tile_next = 0u;
            c.ctl->cta_done = 0u;
#ifdef YASPH_SWEEP_TIMING
            if (c.dbg) {  // span of this launch as seen from inside: first CTA's first instruction to the last CTA's last
                atomicAdd(&c.dbg[9], global_timer_ns_sweep() - c.dbg[8]);
                atomicAdd(&c.dbg[10], 1ull);
                c.dbg[8] = ~0ull;
            }
#endif
        }
    }
    if (Op::REDUCE != REDUCE_NONE) {
        __shared__ double wred[SW_THREADS / 32];
        __shared__ bool is_last;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double u = __shfl_xor_sync(0xffffffffu, racc, o);
            racc = Op::REDUCE == REDUCE_SUM ? racc + u : fmax(racc, u);
        }
        if (lane == 0) wred[warp] = racc;  // the producer warp contributes the neutral 0 (all reduced quantities are >= 0)
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = wred[0];
            for (int w = 1; w < SW_THREADS / 32; ++w) tsum = Op::REDUCE == REDUCE_SUM ? tsum + wred[w] : fmax(tsum, wred[w]);
            c.partials[blockIdx.x] = tsum;
            __threadfence();
            unsigned int ticket = atomicAdd(&c.ctl->ticket[Op::TICKET], 1u);
            is_last = ticket == gridDim.x - 1;
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            double v = 0.0;
            for (uint32_t b = threadIdx.x; b < gridDim.x + c.n_partials_unstaged; b += blockDim.x) {
                double pb = b < gridDim.x ? reinterpret_cast<volatile double*>(c.partials)[b] : c.partials_unstaged[b - gridDim.x];
                v = Op::REDUCE == REDUCE_SUM ? v + pb : fmax(v, pb);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double u = __shfl_xor_sync(0xffffffffu, v, o);
                v = Op::REDUCE == REDUCE_SUM ? v + u : fmax(v, u);
            }
            __syncthreads();
            if (lane == 0) wred[warp] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = wred[0];
                for (int w = 1; w < SW_THREADS / 32; ++w) tot = Op::REDUCE == REDUCE_SUM ? tot + wred[w] : fmax(tot, wred[w]);
                c.ctl->ticket[Op::TICKET] = 0u;
                c.ctl->tile_next = 0u;  // re-arm the tile queue (every CTA has left its tile loop)
                op.finalize(c, tot);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// density (+ alpha, + WCSPH pressure)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tait_pressure(float stiffness, float rho0, float rho) {  // wscsph.rs:52-57
    return stiffness * (powi_f(fmaxf(rho / rho0, 1.0f), 7) - 1.0f);
}
template <int KERNEL, bool WITH_ALPHA, bool WITH_PRESSURE = false>
struct OpDensityAlpha {
    typedef NoPay P0;
    typedef NoPay P1;
    static constexpr int NPAY = 0;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_NONE;
    static constexpr int TICKET = 0;
    static constexpr bool PAD_IS_ZERO = false;  // W(0) * m is not zero: padded entries are masked
    typedef NoPay O0;
    typedef NoPay O1;
    static constexpr int NOWN = 0;
    static constexpr bool WARP_TAIL = false;
    __device__ __forceinline__ const O0* own0() const { return nullptr; }
    __device__ __forceinline__ const O1* own1() const { return nullptr; }
    struct Acc {
        float dens;
        float2 gsum;
        float gsq;
    };
    float* dens;
    float* alpha;
    float2* rho_p;  // WCSPH: (rho, Tait pressure) per particle
    float stiffness;
    __device__ __forceinline__ const P0* pay0() const { return nullptr; }
    __device__ __forceinline__ const P1* pay1() const { return nullptr; }
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon&) {}
    __device__ __forceinline__ float w(const KernelConsts& k, float r_sq, float r) const {
        if (KERNEL == 0) return wendland_w(k, r);
        if (KERNEL == 1) return poly6_w(k, r_sq);
        if (KERNEL == 2) return spiky_w(k, r);
        return cubic_w(k, r);
    }
    __device__ __forceinline__ bool init(const SweepCommon& c, Acc& a, uint32_t, float2, P0, P1, uint32_t) const {
        a.dens = w(c.kc, 0.0f, 0.0f) * c.mass;  // self contribution, fluidparticleworld.rs:213
        a.gsum = f2(0.0f, 0.0f);
        a.gsq = 0.0f;
        return true;
    }
    __device__ __forceinline__ void pair(const SweepCommon& c, Acc& a, float2 pi, float2 pj) const {
        const float2 rij = pj - pi;
        const float r_sq = mag2(rij);
        const float r = (KERNEL == 1 && !WITH_ALPHA) ? 0.0f : sqrtf(r_sq);  // Poly6 needs r^2 only (poly6.rs:28-31)
        a.dens += w(c.kc, r_sq, r) * c.mass;
        if (WITH_ALPHA) {
            const float2 g = (wendland_grad_scalar(c.kc, r) * rij) * c.mass;
            a.gsum = a.gsum + g;
            a.gsq += mag2(g);
        }
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, P0, P1, float2 pj, P0, P1) const { pair(c, a, pi, pj); }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, P0, P1, float2 pb) const { pair(c, a, pi, pb); }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0, P1, bool, O0, O1) const {
        const float rho = fmaxf(a.dens, c.rho0);  // fluidparticleworld.rs:229
        dens[i] = rho;
        if (WITH_ALPHA) alpha[i] = 1.0f / fmaxf(mag2(a.gsum) + a.gsq, 1e-6f);  // dfsph.rs:94
        if (WITH_PRESSURE) rho_p[i] = f2(rho, tait_pressure(stiffness, c.rho0, rho));  // wscsph.rs:91-92, once per particle
        return 0.0;
    }
    __device__ __forceinline__ void finalize(const SweepCommon&, double) const {}
};

// alpha only (yasph_compute_alpha: dfsph.rs:68-97 on its own)
struct OpAlphaOnly : OpDensityAlpha<0, true> {
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, P0, P1, bool, O0, O1) const {
        alpha[i] = 1.0f / fmaxf(mag2(a.gsum) + a.gsq, 1e-6f);
        return 0.0;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// viscosity + gravity (DFSPH non-pressure forces) and the CFL maximum
// ---------------------------------------------------------------------------------------------------------------------
struct ViscParams {
    int kind;     // 0 XSPH, 1 physical
    float coeff;  // epsilon * m  (xsph.rs:22)  or  mu * m  (physical.rs:22), the first product of the reference's expression
};
__device__ __forceinline__ float visc_scalar(const KernelConsts& k, const ViscParams& v, float dt, float r_sq, float r, float rhoj) {
    if (v.kind == 0) return v.coeff * poly6_w(k, r_sq) / (rhoj * dt);
    return v.coeff * visc_laplacian(k, r) / rhoj;
}

struct OpViscosity {
    typedef float2 P0;  // velocity
    typedef float P1;   // density
    static constexpr int NPAY = 2;
    static constexpr bool USES_STATIC = false;
    static constexpr int REDUCE = REDUCE_MAX;
    static constexpr int TICKET = 1;
    static constexpr bool PAD_IS_ZERO = true;  // s * (v_i - v_i) == 0
    typedef NoPay O0;
    typedef NoPay O1;
    static constexpr int NOWN = 0;
    static constexpr bool WARP_TAIL = false;
    __device__ __forceinline__ const O0* own0() const { return nullptr; }
    __device__ __forceinline__ const O1* own1() const { return nullptr; }
    typedef float2 Acc;
    const float2* vel;
    const float* dens;
    float2* accel;
    float2 base_accel;  // (gravity * m) / m, dfsph.rs:442-444
    ViscParams vp;
    float dt;
    uint32_t need_token;  // != 0: launched ahead of the previous step's read-back; runs only if k_begin_step began this step
    __device__ __forceinline__ const P0* pay0() const { return vel; }
    __device__ __forceinline__ const P1* pay1() const { return dens; }
    __device__ __forceinline__ bool skip(const Control* ctl) const { return need_token != 0u && ctl->step_token != need_token; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt_prev; }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, P0, P1, uint32_t) const {
        a = base_accel;
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, P0 vi, P1, float2 pj, P0 vj, P1 rhoj) const {
        const float2 rij = pj - pi;
        const float r_sq = mag2(rij);
        const float r = vp.kind == 0 ? 0.0f : sqrtf(r_sq);  // XSPH needs r^2 only (xsph.rs:19-24)
        const float s = visc_scalar(c.kc, vp, dt, r_sq, r, rhoj);
        a = a + s * (vj - vi);
    }
    __device__ __forceinline__ void stat(const SweepCommon&, Acc&, float2, P0, P1, float2) const {}
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, P0 vi, P1, bool, O0, O1) const {
        accel[i] = a;
        return (double)mag2(vi + a * dt);  // dfsph.rs:476
    }
    __device__ __forceinline__ void finalize(const SweepCommon& c, double mx) const { c.ctl->max_v2_bits = __float_as_uint((float)mx); }
};

// ---------------------------------------------------------------------------------------------------------------------
// Jacobi A: density error (SOLVER 0) / density change (SOLVER 1) + residual and loop control
// ---------------------------------------------------------------------------------------------------------------------
struct SolverParams {
    float max_error;
    uint32_t max_iters;
};
// iteration bookkeeping and the loop decision of correct_density_error / correct_divergence_error (dfsph.rs:219-245, 374-400)
// from the residual sum in ctl->resid_sum
template <int SOLVER>
__device__ __forceinline__ void jacobi_decide(Control* ctl, const SolverParams& sp, uint32_t iter_index, float n_avg, float rho0) {
    const float dt = ctl->dt;
    const float s = (float)ctl->resid_sum;  // f64 accumulation rounded once (DESIGN.md "residual sums")
    const uint32_t it = iter_index + 1;
    ctl->iters[SOLVER] = it;
    bool conv;
    float avg;
    if (SOLVER == 0) {
        avg = s / n_avg;                      // dfsph.rs:221
        const float rel = avg / rho0;         // dfsph.rs:222
        conv = rel * dt < sp.max_error;       // dfsph.rs:226
    } else {
        avg = s / n_avg / rho0;               // dfsph.rs:376-377
        conv = avg * dt < sp.max_error;       // dfsph.rs:381
    }
    ctl->avg[SOLVER] = avg;
    if (!isfinite(avg)) {  // the reference asserts (dfsph.rs:223,378); stop and report
        ctl->nonfinite |= 1u << SOLVER;
        conv = true;
    }
    if (conv) {
        ctl->stop_iter[SOLVER] = it;
    } else if (it > sp.max_iters) {  // dfsph.rs:236,391
        ctl->stop_iter[SOLVER] = it;
        ctl->not_converged |= 1u << SOLVER;
    }
}
template <int SOLVER>
struct OpJacobiA {
    typedef float2 P0;  // predicted velocity
    typedef NoPay P1;
    static constexpr int NPAY = 1;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_SUM;
    static constexpr int TICKET = 2;
    static constexpr bool PAD_IS_ZERO = true;  // (v_i - v_i) . grad == 0
    typedef float O0;  // rho_i
    typedef float O1;  // alpha_i
    static constexpr int NOWN = 2;
    static constexpr bool WARP_TAIL = false;
    __device__ __forceinline__ const O0* own0() const { return dens; }
    __device__ __forceinline__ const O1* own1() const { return alpha; }
    typedef float Acc;
    const float2* vstar;
    const float* dens;
    const float* alpha;
    float* kfac;  // k_i = err_i * alpha_i, the only use of err_i after this pass (dfsph.rs:141,150 / 295,304)
    SolverParams sp;
    uint32_t iter_index;
    float dt;
    __device__ __forceinline__ const P0* pay0() const { return vstar; }
    __device__ __forceinline__ const P1* pay1() const { return nullptr; }
    __device__ __forceinline__ bool skip(const Control* ctl) const { return iter_index >= ctl->stop_iter[SOLVER]; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt; }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, P0, P1, uint32_t ct) const {
        a = 0.0f;
        return SOLVER == 0 ? true : ct >= 9u;  // particle deficiency, dfsph.rs:261
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, P0 vi, P1, float2 pj, P0 vj, P1) const {
        a += dot2(vi - vj, wendland_grad_from_positions(c.kc, pi, pj));
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, P0 vi, P1, float2 pb) const {
        a += dot2(vi, wendland_grad_from_positions(c.kc, pi, pb));
    }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0, P1, bool active, O0 rho_i, O1 alpha_i) const {
        float e;
        if (SOLVER == 0) {
            e = rho_i + a * c.mass * dt;         // dfsph.rs:121
            e = fmaxf(c.rho0, e) - c.rho0;       // dfsph.rs:124
        } else {
            e = active ? fmaxf(a * c.mass, 0.0f) : 0.0f;  // dfsph.rs:262,277-278
        }
        kfac[i] = e * alpha_i;
        return (double)e;
    }
    __device__ __forceinline__ void finalize(const SweepCommon& c, double sum) const {
        c.ctl->resid_sum = sum;
        if (c.ghost == nullptr) jacobi_decide<SOLVER>(c.ctl, sp, iter_index, c.n_avg, c.rho0);  // slab mode: after the all-reduce
    }
};
// the same decision as the tail of the residual all-reduce kernel (slab.cuh: k_allreduce_peer)
template <int SOLVER>
struct JacobiDecideAfter {
    SolverParams sp;
    uint32_t iter_index;
    float n_avg, rho0;
    __device__ __forceinline__ void operator()(Control* ctl) const {
        if (iter_index < ctl->stop_iter[SOLVER]) jacobi_decide<SOLVER>(ctl, sp, iter_index, n_avg, rho0);
    }
};
template <int SOLVER>
__global__ void k_jacobi_decide(Control* ctl, SolverParams sp, uint32_t iter_index, float n_avg, float rho0) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && iter_index < ctl->stop_iter[SOLVER]) jacobi_decide<SOLVER>(ctl, sp, iter_index, n_avg, rho0);
}

// ---------------------------------------------------------------------------------------------------------------------
// densities + alpha factors + the first density-change pass of the divergence solver in ONE sweep
// ---------------------------------------------------------------------------------------------------------------------
// update_densities (fluidparticleworld.rs:197-231), compute_alpha_factors (dfsph.rs:68-97) and iteration 0 of
// compute_density_change (dfsph.rs:249-280) walk the same neighbours with the same r_ij and the same gradient and need
// nothing from the neighbours that an earlier pass of this step has to produce (positions and v* only), so the step runs
// them as one sweep whenever the divergence warm start -- which would move v* first (dfsph.rs:354-360) -- does not run.
// Every accumulator keeps its own reference order, so each result is bit-identical to the three separate passes.
struct OpDensityAlphaDiv {
    typedef float2 P0;  // predicted velocity
    typedef NoPay P1;
    static constexpr int NPAY = 1;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_SUM;
    static constexpr int TICKET = 2;
    static constexpr bool PAD_IS_ZERO = false;
    typedef NoPay O0;
    typedef NoPay O1;
    static constexpr int NOWN = 0;
    static constexpr bool WARP_TAIL = false;
    __device__ __forceinline__ const O0* own0() const { return nullptr; }
    __device__ __forceinline__ const O1* own1() const { return nullptr; }
    struct Acc {
        float dens;
        float2 gsum;
        float gsq;
        float div;
        uint32_t ct;
    };
    const float2* vstar;
    float* dens;
    float* alpha;
    float* kfac;
    SolverParams sp;
    __device__ __forceinline__ const P0* pay0() const { return vstar; }
    __device__ __forceinline__ const P1* pay1() const { return nullptr; }
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon&) {}
    __device__ __forceinline__ bool init(const SweepCommon& c, Acc& a, uint32_t, float2, P0, P1, uint32_t ct) const {
        a.dens = wendland_w(c.kc, 0.0f) * c.mass;  // fluidparticleworld.rs:213
        a.gsum = f2(0.0f, 0.0f);
        a.gsq = 0.0f;
        a.div = 0.0f;
        a.ct = ct;
        return true;
    }
    __device__ __forceinline__ void pair(const SweepCommon& c, Acc& a, float2 pi, float2 vrel, float2 pj) const {
        const float2 rij = pj - pi;
        const float r = sqrtf(mag2(rij));
        a.dens += wendland_w(c.kc, r) * c.mass;                  // fluidparticleworld.rs:218-219 / 224-225
        const float2 grad = wendland_grad_scalar(c.kc, r) * rij;  // kernel.rs:22-28
        const float2 g = grad * c.mass;                          // dfsph.rs:81-83 / 87-89
        a.gsum = a.gsum + g;
        a.gsq += mag2(g);
        a.div += dot2(vrel, grad);                               // dfsph.rs:267 / 274
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, P0 vi, P1, float2 pj, P0 vj, P1) const { pair(c, a, pi, vi - vj, pj); }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, P0 vi, P1, float2 pb) const { pair(c, a, pi, vi, pb); }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0, P1, bool, O0, O1) const {
        dens[i] = fmaxf(a.dens, c.rho0);                                   // fluidparticleworld.rs:229
        const float al = 1.0f / fmaxf(mag2(a.gsum) + a.gsq, 1e-6f);       // dfsph.rs:94
        alpha[i] = al;
        const float e = a.ct >= 9u ? fmaxf(a.div * c.mass, 0.0f) : 0.0f;  // dfsph.rs:261-262,277-278
        kfac[i] = e * al;                                                 // dfsph.rs:295
        return (double)e;
    }
    __device__ __forceinline__ void finalize(const SweepCommon& c, double sum) const {
        c.ctl->resid_sum = sum;
        if (c.ghost == nullptr) jacobi_decide<1>(c.ctl, sp, 0u, c.n_avg, c.rho0);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Jacobi B and warm starts
// ---------------------------------------------------------------------------------------------------------------------
template <int SOLVER, bool WARM>
struct OpJacobiB {
    typedef float P0;  // k_j (Jacobi B) / raw warm-start value (warm start; the clamp is applied on use)
    typedef NoPay P1;
    static constexpr int NPAY = 1;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_NONE;
    static constexpr int TICKET = 0;
    static constexpr bool PAD_IS_ZERO = true;  // (k_i + k_i) * grad(r = 0) == (0, 0)
    typedef float2 O0;  // v*_i
    typedef float O1;   // accumulated warm-start value of i
    static constexpr int NOWN = 2;
    static constexpr bool WARP_TAIL = false;
    __device__ __forceinline__ const O0* own0() const { return vstar; }
    __device__ __forceinline__ const O1* own1() const { return warm; }
    typedef float2 Acc;
    float2* vstar;
    const float* kfac;
    float* warm;  // warmstart_kappa (SOLVER 0) / warmstart_stiffness (SOLVER 1)
    float clamp_min;  // -0.5 * rho0 * rho0
    uint32_t iter_index;
    float dt, inv_dt;
    __device__ __forceinline__ const P0* pay0() const { return WARM ? warm : kfac; }
    __device__ __forceinline__ const P1* pay1() const { return nullptr; }
    __device__ __forceinline__ bool skip(const Control* ctl) const {
        return WARM ? ctl->warm[SOLVER] == 0u : iter_index >= ctl->stop_iter[SOLVER];
    }
    __device__ __forceinline__ void prepare(const SweepCommon& c) {
        dt = c.ctl->dt;
        inv_dt = 1.0f / dt;  // dfsph.rs:132,167
    }
    __device__ __forceinline__ float kval(float raw) const {
        return WARM ? 0.5f * fmaxf(raw, clamp_min) : raw;  // dfsph.rs:201-203 / 356-358
    }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, P0, P1, uint32_t) const {
        a = f2(0.0f, 0.0f);
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, P0 ki, P1, float2 pj, P0 kj, P1) const {
        a = a + (kval(ki) + kval(kj)) * wendland_grad_from_positions(c.kc, pi, pj);
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, P0 ki, P1, float2 pb) const {
        a = a + kval(ki) * wendland_grad_from_positions(c.kc, pi, pb);
    }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0 ki, P1, bool, O0 v, O1 warm_i) const {
        if (SOLVER == 0)
            vstar[i] = sub_scalar(v, inv_dt * a * c.mass);  // dfsph.rs:159,191
        else
            vstar[i] = sub_scalar(v, a * c.mass);           // dfsph.rs:312,342
        // The warm start never stores its clamped values: other tiles are still reading the raw array, and iteration 0 of the
        // solve that always follows overwrites it (zeroing, dfsph.rs:206-208, fused into that iteration).
        if (!WARM) warm[i] = (iter_index == 0 ? 0.0f : warm_i) + ki;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(const SweepCommon&, double) const {}
};

// Jacobi B of the density solver that, when it is the solve's LAST correction (decided on the device by pass A of the same iteration),
// also advects the particles (dfsph.rs:502-509) and generates the keys and digit histograms of the re-sort that follows
// (neighborhood_search.rs:111-114): the advected position goes to a second array (neighbours of other tiles still read the old one), key
// and index to the sort's input.  Saves k_advect_keygen's pass over the particles; single GPU only (no slab classification here).
struct OpJacobiBAdvect : OpJacobiB<0, false> {
    static constexpr bool WARP_TAIL = true;
    float2* pos_out;
    uint32_t* keys;
    uint32_t* idx;
    uint32_t* sort_scratch;
    GridParams grid;
    uint32_t last;  // set by prepare(): this launch is the solve's last B
    __device__ __forceinline__ void prepare(const SweepCommon& c) {
        OpJacobiB<0, false>::prepare(c);
        last = c.ctl->stop_iter[0] == iter_index + 1u ? 1u : 0u;
    }
    __device__ __forceinline__ bool tail_on() const { return last != 0u; }
    __device__ __forceinline__ uint32_t* tail_scratch() const { return sort_scratch; }
    __device__ __forceinline__ double finish_tail(const SweepCommon& c, Acc& a, uint32_t i, float2 pi, P0 ki, P1, bool, O0 v, O1 warm_i, uint32_t& key) const {
        const float2 vn = sub_scalar(v, inv_dt * a * c.mass);  // dfsph.rs:159
        vstar[i] = vn;
        warm[i] = (iter_index == 0 ? 0.0f : warm_i) + ki;
        if (last) {
            const float2 p = pi + vn * dt;  // dfsph.rs:505
            pos_out[i] = p;
            key = position_to_cidx(grid, p);
            keys[i] = key;
        }
        return 0.0;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// WCSPH accelerations
// ---------------------------------------------------------------------------------------------------------------------
struct OpWcsphAccel {
    typedef float2 P0;  // velocity
    typedef float2 P1;  // (rho, p), written by the density pass
    static constexpr int NPAY = 2;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_MAX;
    static constexpr int TICKET = 1;
    static constexpr bool PAD_IS_ZERO = true;  // spiky gradient times r_ij and the velocity difference of a particle with itself vanish
    typedef NoPay O0;
    typedef NoPay O1;
    static constexpr int NOWN = 0;
    static constexpr bool WARP_TAIL = false;
    __device__ __forceinline__ const O0* own0() const { return nullptr; }
    __device__ __forceinline__ const O1* own1() const { return nullptr; }
    typedef float2 Acc;
    const float2* vel;
    const float2* rho_p;
    float2* accel;
    float2 gravity;
    ViscParams vp;
    float boundary_force_factor;
    float dt;
    __device__ __forceinline__ const P0* pay0() const { return vel; }
    __device__ __forceinline__ const P1* pay1() const { return rho_p; }
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt_prev; }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, P0, P1, uint32_t) const {
        a = gravity;  // wscsph.rs:84
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, P0 vi, P1 rpi, float2 pj, P0 vj, P1 rpj) const {
        const float2 rij = pj - pi;
        const float r_sq = mag2(rij);
        const float r = sqrtf(r_sq);
        const float pu = -c.mass * (rpi.y + rpj.y) / (2.0f * rpi.x * rpj.x);  // wscsph.rs:101
        a = a + pu * (spiky_grad_scalar(c.kc, r) * rij);                      // wscsph.rs:102
        a = a + visc_scalar(c.kc, vp, dt, r_sq, r, rpj.x) * (vj - vi);        // wscsph.rs:104-106
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, P0, P1, float2 pb) const {
        const float2 rij = pb - pi;
        const float r_sq = mag2(rij);
        a = sub_scalar(a, (boundary_force_factor * spiky_w(c.kc, sqrtf(r_sq)) / r_sq) * rij);  // wscsph.rs:113-115
    }
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, P0 vi, P1, bool, O0, O1) const {
        accel[i] = a;
        return (double)mag2(vi + a * dt);  // wscsph.rs:162
    }
    __device__ __forceinline__ void finalize(const SweepCommon& c, double mx) const { c.ctl->max_v2_bits = __float_as_uint((float)mx); }
};

// ---------------------------------------------------------------------------------------------------------------------
// element-wise passes
// ---------------------------------------------------------------------------------------------------------------------
// TimeManager::simulation_step at step entry (dfsph.rs:433 / wscsph.rs:133)
// guarded != 0: enqueued ahead of the read-back that ends the previous step (yasph_step_n) -- the step begins only if that step's
// divergence solve has finished; the kernels of the step's head then test the token.
// snap != null: the control block as the previous step left it is copied there first (its report; published to the host from a side
// stream while this step's first pass already runs).
__global__ void k_begin_step(Control* ctl, uint32_t token, uint32_t guarded, Control* snap) {
    pdl_enter();
    if (snap != nullptr) {
        for (uint32_t q = threadIdx.x; q < sizeof(Control) / 4; q += blockDim.x) reinterpret_cast<uint32_t*>(snap)[q] = reinterpret_cast<const uint32_t*>(ctl)[q];
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (guarded && ctl->stop_iter[1] == 0xFFFFFFFFu) return;
        ctl->step_token = token;
        ctl->step_prev_ns = ctl->step_ns;
        ctl->dt_prev = duration_as_secs_f32(ctl->step_ns);
        // the frame loop's `total_simulated_time += simulation_step` ahead of the step (timemanager.rs:246)
        if (ctl->total_is_current)
            ctl->total_is_current = 0u;
        else
            ctl->total_simulated_ns += ctl->step_ns;
        ctl->max_v2_bits = 0u;
        ctl->not_converged = 0u;
        ctl->nonfinite = 0u;
        // set up the divergence solver (dfsph.rs:354,368): its iteration count of the previous step does not change until it runs
        ctl->warm[1] = ctl->iters[1] > 1u ? 1u : 0u;
        ctl->stop_iter[1] = 0xFFFFFFFFu;
    }
}
// update_simulation_step (timemanager.rs:252-279) evaluated redundantly by every thread from the reduced maximum, then
// MODE 0: velocity prediction v* = v + a dt (dfsph.rs:486-491); MODE 1: second leap-frog kick v += 0.5 dt a (wscsph.rs:175-177)
template <int MODE>
__global__ void k_timestep_apply(Control* ctl, TimeParams tp, float particle_diameter, const float2* vel_in,
                                 const float2* __restrict__ accel, float2* vel_out, uint32_t n) {
    // one thread per CTA evaluates the rule (64-bit integer and f64 arithmetic); every CTA gets the same result
    __shared__ float dt_s;
    pdl_enter();
    if (threadIdx.x == 0) {
        const float max_velocity = sqrtf(__uint_as_float(ctl->max_v2_bits));
        const unsigned long long step = update_simulation_step(tp, ctl->step_prev_ns, particle_diameter, max_velocity, ctl->total_simulated_ns);
        const float dt0 = duration_as_secs_f32(step);
        dt_s = dt0;
        if (blockIdx.x == 0) {
            ctl->step_ns = step;
            ctl->dt = dt0;
            ctl->max_velocity = max_velocity;
            if (MODE == 0) {  // set up the density solver (dfsph.rs:199,213)
                ctl->warm[0] = ctl->iters[0] > 1u ? 1u : 0u;
                ctl->stop_iter[0] = 0xFFFFFFFFu;
            }
        }
    }
    __syncthreads();
    const float dt = dt_s;
    // two particles per thread (16-byte accesses)
    const uint32_t i = 2u * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 < n) {
        const float4 v = *reinterpret_cast<const float4*>(vel_in + i), a = *reinterpret_cast<const float4*>(accel + i);
        const float2 o0 = MODE == 0 ? f2(v.x, v.y) + f2(a.x, a.y) * dt : f2(v.x, v.y) + 0.5f * dt * f2(a.x, a.y);
        const float2 o1 = MODE == 0 ? f2(v.z, v.w) + f2(a.z, a.w) * dt : f2(v.z, v.w) + 0.5f * dt * f2(a.z, a.w);
        *reinterpret_cast<float4*>(vel_out + i) = make_float4(o0.x, o0.y, o1.x, o1.y);
    } else if (i < n) {
        vel_out[i] = MODE == 0 ? vel_in[i] + accel[i] * dt : vel_in[i] + 0.5f * dt * accel[i];
    }
}

}  // namespace yasph
