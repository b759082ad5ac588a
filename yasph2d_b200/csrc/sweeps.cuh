// sweeps.cuh -- every neighbour-dependent per-particle pass of the WCSPH / DFSPH solvers as one tile-staged kernel.
//
// Skeleton (k_sweep): persistent CTAs, each walking tiles blockIdx.x, blockIdx.x + gridDim.x, ...  A tile's own particles
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

constexpr int SW_THREADS = TILE_THREADS;

struct SweepCommon {
    TileTables tt;
    const unsigned long long* lists;
    const uchar2* counts;
    const float2* pos;
    const float2* bpos;
    Control* ctl;
    KernelConsts kc;
    uint32_t cap_dyn, cap_stat;
    uint32_t n;
    float mass, rho0;
    double* partials;  // [gridDim.x]
    // slab mode (multi-GPU): ghosts are excluded from the reductions, the residual average runs over the global particle count
    // and the convergence decision is taken after the all-reduce (k_jacobi_decide)
    const uint8_t* ghost;  // null on a single GPU
    float n_avg;           // particle count of dfsph.rs:221,376 as f32
};

enum ReduceKind { REDUCE_NONE = 0, REDUCE_SUM = 1, REDUCE_MAX = 2 };

struct NoPay {};

// per-buffer shared-memory layout of a sweep: float2 pos[cap_dyn] | P0[cap_dyn] | P1[cap_dyn] | float2 stat[cap_stat]
template <class Op>
struct SweepLayout {
    static constexpr size_t P0 = Op::NPAY >= 1 ? sizeof(typename Op::P0) : 0;
    static constexpr size_t P1 = Op::NPAY >= 2 ? sizeof(typename Op::P1) : 0;
    __host__ __device__ static size_t buffer_bytes(uint32_t cap_dyn, uint32_t cap_stat) {
        return (size_t)cap_dyn * (sizeof(float2) + P0 + P1) + (Op::USES_STATIC ? (size_t)cap_stat * sizeof(float2) : 0);
    }
    __host__ __device__ static size_t total_bytes(uint32_t cap_dyn, uint32_t cap_stat) { return 3 * sizeof(TileRuns) + 2 * buffer_bytes(cap_dyn, cap_stat); }
};
template <class Op>
inline size_t sweep_smem_bytes(uint32_t cap_dyn, uint32_t cap_stat) {
    return SweepLayout<Op>::total_bytes(cap_dyn, cap_stat);
}

template <class Op>
struct SweepBuffer {
    float2* pos;
    typename Op::P0* p0;
    typename Op::P1* p1;
    float2* stat;
    __device__ __forceinline__ SweepBuffer(unsigned char* base, uint32_t cap_dyn) {
        pos = reinterpret_cast<float2*>(base);
        p0 = reinterpret_cast<typename Op::P0*>(base + (size_t)cap_dyn * sizeof(float2));
        p1 = reinterpret_cast<typename Op::P1*>(base + (size_t)cap_dyn * (sizeof(float2) + SweepLayout<Op>::P0));
        stat = reinterpret_cast<float2*>(base + (size_t)cap_dyn * (sizeof(float2) + SweepLayout<Op>::P0 + SweepLayout<Op>::P1));
    }
};

template <class Op>
__device__ __forceinline__ void sweep_issue_stage(const SweepCommon& c, const Op& op, const TileRuns& tr, const SweepBuffer<Op>& b) {
    const TileHeader& h = tr.hdr;
    if (h.dyn_total > c.cap_dyn || h.stat_total > c.cap_stat) return;  // cannot happen: capacities are the maxima over all tiles
    for (uint32_t s = threadIdx.x; s < h.dyn_total; s += SW_THREADS) {
        const uint32_t g = dyn_slot_to_global(tr, s);
        cp_async<8>(&b.pos[s], &c.pos[g]);
        if constexpr (Op::NPAY >= 1) cp_async<sizeof(typename Op::P0)>(&b.p0[s], &op.pay0()[g]);
        if constexpr (Op::NPAY >= 2) cp_async<sizeof(typename Op::P1)>(&b.p1[s], &op.pay1()[g]);
    }
    if (Op::USES_STATIC)
        for (uint32_t s = threadIdx.x; s < h.stat_total; s += SW_THREADS) cp_async<8>(&b.stat[s], &c.bpos[run_slot_to_global(tr.rs, s)]);
}

template <class Op>
__global__ void __launch_bounds__(SW_THREADS) k_sweep(SweepCommon c, Op op) {
    typedef typename Op::P0 P0;
    typedef typename Op::P1 P1;
    if (op.skip(c.ctl)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileRuns* runs = reinterpret_cast<TileRuns*>(smem_raw);  // [3]
    const size_t bbytes = SweepLayout<Op>::buffer_bytes(c.cap_dyn, c.cap_stat);
    const SweepBuffer<Op> buf0(smem_raw + 3 * sizeof(TileRuns), c.cap_dyn), buf1(smem_raw + 3 * sizeof(TileRuns) + bbytes, c.cap_dyn);
    op.prepare(c);
    const uint32_t ntiles = c.ctl->num_tiles;
    const uint32_t G = gridDim.x;
    double racc = 0.0;
    // prologue: tables of the first two tiles, copies of the first
    {
        const uint32_t t0 = blockIdx.x, t1 = blockIdx.x + G;
        if (t0 < ntiles) load_tile_runs(runs[0], c.tt.runs + t0);
        if (t1 < ntiles) load_tile_runs(runs[1], c.tt.runs + t1);
        __syncthreads();
        if (t0 < ntiles) sweep_issue_stage(c, op, runs[0], buf0);
        cp_async_commit();
    }
    uint32_t k = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += G, ++k) {
        const SweepBuffer<Op>& cur = (k & 1u) ? buf1 : buf0;
        const SweepBuffer<Op>& nxt = (k & 1u) ? buf0 : buf1;
        const TileRuns& tr = runs[k % 3u];
        RunsPrefetch pre;
        const bool have2 = t + 2 * G < ntiles;
        pre.load(c.tt.runs + t + 2 * G, have2);
        cp_async_wait_all();
        __syncthreads();  // this tile's copies have landed; everybody is done with the previous tile's buffers
        if (t + G < ntiles) sweep_issue_stage(c, op, runs[(k + 1) % 3u], nxt);
        cp_async_commit();
        const TileHeader h = tr.hdr;
        if (h.dyn_total <= c.cap_dyn && h.stat_total <= c.cap_stat) {
            for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += SW_THREADS) {
                const uint32_t i = h.pstart + tl;
                const uint32_t own = h.own_lo + tl;
                const float2 pi = cur.pos[own];
                P0 s0;
                P1 s1;
                if (Op::NPAY >= 1) s0 = cur.p0[own];
                if (Op::NPAY >= 2) s1 = cur.p1[own];
                const uchar2 cnt = c.counts[i];
                const uint32_t cd = cnt.x, ct = Op::USES_STATIC ? cnt.y : cnt.x;
                typename Op::Acc acc;
                const bool active = op.init(c, acc, i, pi, s0, s1, cnt.y);
                if (active) {
                    // dynamic neighbours: entries [0, cd), four per 8-byte word, next word prefetched
                    const uint32_t nkb = (cd + 3u) >> 2;
                    unsigned long long w = nkb ? c.lists[list_word_index(h.pstart, h.pcount, 0, tl)] : 0ull;
                    for (uint32_t kb = 0; kb < nkb; ++kb) {
                        const unsigned long long wn = (kb + 1 < nkb) ? c.lists[list_word_index(h.pstart, h.pcount, kb + 1, tl)] : 0ull;
#pragma unroll
                        for (uint32_t q = 0; q < 4; ++q) {
                            const bool valid = kb * 4 + q < cd;
                            const uint32_t slot = valid ? ((uint32_t)(w >> (16 * q)) & 0xFFFFu) : own;
                            P0 n0;
                            P1 n1;
                            if (Op::NPAY >= 1) n0 = cur.p0[slot];
                            if (Op::NPAY >= 2) n1 = cur.p1[slot];
                            typename Op::Acc trial = acc;
                            op.dyn(c, trial, pi, s0, s1, cur.pos[slot], n0, n1);
                            if (valid) acc = trial;
                        }
                        w = wn;
                    }
                    // static neighbours: entries [cd, ct) of the same list (only near boundaries)
                    if (Op::USES_STATIC) {
                        for (uint32_t kk = cd; kk < ct; ++kk) {
                            const uint32_t slot = unpack_slot(c.lists[list_word_index(h.pstart, h.pcount, kk >> 2, tl)], kk);
                            op.stat(c, acc, pi, s0, s1, cur.stat[slot]);
                        }
                    }
                }
                double r = op.finish(c, acc, i, pi, s0, s1, active);
                if (Op::REDUCE != REDUCE_NONE && c.ghost != nullptr && c.ghost[i]) r = 0.0;
                if (Op::REDUCE == REDUCE_SUM) racc += r;
                if (Op::REDUCE == REDUCE_MAX) racc = fmax(racc, r);
            }
        }
        pre.store(runs[(k + 2) % 3u], have2);  // last read before this iteration's barrier; next read after the next one
    }
    cp_async_wait_all();
    if (Op::REDUCE != REDUCE_NONE) {
        __shared__ double wred[SW_THREADS / 32];
        __shared__ bool is_last;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double u = __shfl_xor_sync(0xffffffffu, racc, o);
            racc = Op::REDUCE == REDUCE_SUM ? racc + u : fmax(racc, u);
        }
        if (lane_id() == 0) wred[threadIdx.x >> 5] = racc;
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
            for (uint32_t b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
                double pb = reinterpret_cast<volatile double*>(c.partials)[b];
                v = Op::REDUCE == REDUCE_SUM ? v + pb : fmax(v, pb);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double u = __shfl_xor_sync(0xffffffffu, v, o);
                v = Op::REDUCE == REDUCE_SUM ? v + u : fmax(v, u);
            }
            __syncthreads();
            if (lane_id() == 0) wred[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = wred[0];
                for (int w = 1; w < SW_THREADS / 32; ++w) tot = Op::REDUCE == REDUCE_SUM ? tot + wred[w] : fmax(tot, wred[w]);
                c.ctl->ticket[Op::TICKET] = 0u;
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
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0, P1, bool) const {
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
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, P0, P1, bool) const {
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
    typedef float2 Acc;
    const float2* vel;
    const float* dens;
    float2* accel;
    float2 base_accel;  // (gravity * m) / m, dfsph.rs:442-444
    ViscParams vp;
    float dt;
    __device__ __forceinline__ const P0* pay0() const { return vel; }
    __device__ __forceinline__ const P1* pay1() const { return dens; }
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
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
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, P0 vi, P1, bool) const {
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
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0, P1, bool active) const {
        float e;
        if (SOLVER == 0) {
            e = dens[i] + a * c.mass * dt;       // dfsph.rs:121
            e = fmaxf(c.rho0, e) - c.rho0;       // dfsph.rs:124
        } else {
            e = active ? fmaxf(a * c.mass, 0.0f) : 0.0f;  // dfsph.rs:262,277-278
        }
        kfac[i] = e * alpha[i];
        return (double)e;
    }
    __device__ __forceinline__ void finalize(const SweepCommon& c, double sum) const {
        c.ctl->resid_sum = sum;
        if (c.ghost == nullptr) jacobi_decide<SOLVER>(c.ctl, sp, iter_index, c.n_avg, c.rho0);  // slab mode: after the all-reduce
    }
};
template <int SOLVER>
__global__ void k_jacobi_decide(Control* ctl, SolverParams sp, uint32_t iter_index, float n_avg, float rho0) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && iter_index < ctl->stop_iter[SOLVER]) jacobi_decide<SOLVER>(ctl, sp, iter_index, n_avg, rho0);
}

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
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, P0 ki, P1, bool) const {
        const float2 v = vstar[i];
        if (SOLVER == 0)
            vstar[i] = v - inv_dt * a * c.mass;  // dfsph.rs:159,191
        else
            vstar[i] = v - a * c.mass;           // dfsph.rs:312,342
        // The warm start never stores its clamped values: other tiles are still reading the raw array, and iteration 0 of the
        // solve that always follows overwrites it (zeroing, dfsph.rs:206-208, fused into that iteration).
        if (!WARM) warm[i] = (iter_index == 0 ? 0.0f : warm[i]) + ki;
        return 0.0;
    }
    __device__ __forceinline__ void finalize(const SweepCommon&, double) const {}
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
        a = a - (boundary_force_factor * spiky_w(c.kc, sqrtf(r_sq)) / r_sq) * rij;  // wscsph.rs:113-115
    }
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, P0 vi, P1, bool) const {
        accel[i] = a;
        return (double)mag2(vi + a * dt);  // wscsph.rs:162
    }
    __device__ __forceinline__ void finalize(const SweepCommon& c, double mx) const { c.ctl->max_v2_bits = __float_as_uint((float)mx); }
};

// ---------------------------------------------------------------------------------------------------------------------
// element-wise passes
// ---------------------------------------------------------------------------------------------------------------------
// TimeManager::simulation_step at step entry (dfsph.rs:433 / wscsph.rs:133)
__global__ void k_begin_step(Control* ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->step_prev_ns = ctl->step_ns;
        ctl->dt_prev = duration_as_secs_f32(ctl->step_ns);
        ctl->max_v2_bits = 0u;
        ctl->not_converged = 0u;
        ctl->nonfinite = 0u;
    }
}
// update_simulation_step (timemanager.rs:252-279) evaluated redundantly by every thread from the reduced maximum, then
// MODE 0: velocity prediction v* = v + a dt (dfsph.rs:486-491); MODE 1: second leap-frog kick v += 0.5 dt a (wscsph.rs:175-177)
template <int MODE>
__global__ void k_timestep_apply(Control* ctl, TimeParams tp, float particle_diameter, const float2* vel_in,
                                 const float2* __restrict__ accel, float2* vel_out, uint32_t n) {
    const float max_velocity = sqrtf(__uint_as_float(ctl->max_v2_bits));
    const unsigned long long step = update_simulation_step(tp, ctl->step_prev_ns, particle_diameter, max_velocity);
    const float dt = duration_as_secs_f32(step);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        ctl->step_ns = step;
        ctl->dt = dt;
        ctl->max_velocity = max_velocity;
        if (MODE == 0) {  // set up the density solver (dfsph.rs:199,213)
            ctl->warm[0] = ctl->iters[0] > 1u ? 1u : 0u;
            ctl->stop_iter[0] = 0xFFFFFFFFu;
        }
    }
    if (i < n) {
        if (MODE == 0)
            vel_out[i] = vel_in[i] + accel[i] * dt;
        else
            vel_out[i] = vel_in[i] + 0.5f * dt * accel[i];
    }
}
// set up the divergence solver (dfsph.rs:354,368)
__global__ void k_begin_divergence(Control* ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->warm[1] = ctl->iters[1] > 1u ? 1u : 0u;
        ctl->stop_iter[1] = 0xFFFFFFFFu;
    }
}

}  // namespace yasph
