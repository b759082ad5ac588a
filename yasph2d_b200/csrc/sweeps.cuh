// sweeps.cuh -- every neighbour-dependent per-particle pass of the WCSPH / DFSPH solvers as one tile-staged kernel.
//
// Skeleton (k_sweep): one CTA per 8x8-cell tile.  The tile's own particles plus the 1-cell apron are staged into shared
// memory once (a record per candidate: position + the per-pass neighbour payload; boundary positions), then one thread
// per particle walks its compact neighbour list (u16 shared-memory slots, dynamic first, then static) in the reference's
// order.  Shared memory is sized from the largest tile of the current neighbourhood structure (Control::max_*), a few KB,
// so 4-8 CTAs are resident per SM and the staging latency of one tile hides behind the arithmetic of the others.
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

constexpr int SW_THREADS = 256;

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
};

enum ReduceKind { REDUCE_NONE = 0, REDUCE_SUM = 1, REDUCE_MAX = 2 };

struct NoRec {};
__device__ __forceinline__ float2 rec_pos(float2 r) { return r; }
__device__ __forceinline__ float2 rec_pos(float4 r) { return f2(r.x, r.y); }

// shared memory of a sweep: TileRuns | RecA[cap_dyn] | RecB[cap_dyn] | float2[cap_stat]  (capacities are multiples of 2)
template <class Op>
inline size_t sweep_smem_bytes(uint32_t cap_dyn, uint32_t cap_stat) {
    return sizeof(TileRuns) + (size_t)cap_dyn * (sizeof(typename Op::RecA) + (Op::HAS_B ? sizeof(typename Op::RecB) : 0)) +
           (Op::USES_STATIC ? (size_t)cap_stat * sizeof(float2) : 0);
}

template <class Op>
__global__ void __launch_bounds__(SW_THREADS) k_sweep(SweepCommon c, Op op) {
    typedef typename Op::RecA RecA;
    typedef typename Op::RecB RecB;
    if (op.skip(c.ctl)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileRuns& tr = *reinterpret_cast<TileRuns*>(smem_raw);
    RecA* sa = reinterpret_cast<RecA*>(smem_raw + sizeof(TileRuns));
    RecB* sb = reinterpret_cast<RecB*>(smem_raw + sizeof(TileRuns) + sizeof(RecA) * (size_t)c.cap_dyn);
    float2* sstat = reinterpret_cast<float2*>(smem_raw + sizeof(TileRuns) + (sizeof(RecA) + (Op::HAS_B ? sizeof(RecB) : 0)) * (size_t)c.cap_dyn);
    op.prepare(c);
    const uint32_t ntiles = c.ctl->num_tiles;
    double racc = 0.0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        load_tile_runs(tr, c.tt.runs + t);
        __syncthreads();
        const TileHeader h = tr.hdr;
        if (h.dyn_total <= c.cap_dyn && h.stat_total <= c.cap_stat) {
            for (uint32_t s = threadIdx.x; s < h.dyn_total; s += SW_THREADS) {
                const uint32_t g = dyn_slot_to_global(tr, s);
                sa[s] = op.load_a(c, g);
                if (Op::HAS_B) sb[s] = op.load_b(c, g);
            }
            if (Op::USES_STATIC)
                for (uint32_t s = threadIdx.x; s < h.stat_total; s += SW_THREADS) sstat[s] = c.bpos[run_slot_to_global(tr.rs, s)];
            __syncthreads();
            for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += SW_THREADS) {
                const uint32_t i = h.pstart + tl;
                const uint32_t own = h.own_lo + tl;
                const RecA self_a = sa[own];
                RecB self_b;
                if (Op::HAS_B) self_b = sb[own];
                const uchar2 cnt = c.counts[i];
                const uint32_t cd = cnt.x, ct = Op::USES_STATIC ? cnt.y : cnt.x;
                typename Op::Acc acc;
                const bool active = op.init(c, acc, i, self_a, self_b, cnt.y);
                if (active) {
                    // dynamic neighbours: entries [0, cd), four per 8-byte word, next word prefetched
                    const uint32_t nkb = (cd + 3u) >> 2;
                    unsigned long long w = nkb ? c.lists[list_word_index(h.pstart, h.pcount, 0, tl)] : 0ull;
                    for (uint32_t kb = 0; kb < nkb; ++kb) {
                        const unsigned long long wn = (kb + 1 < nkb) ? c.lists[list_word_index(h.pstart, h.pcount, kb + 1, tl)] : 0ull;
#pragma unroll
                        for (uint32_t q = 0; q < 4; ++q) {
                            if (kb * 4 + q < cd) {
                                const uint32_t slot = (uint32_t)(w >> (16 * q)) & 0xFFFFu;
                                RecB nb_b;
                                if (Op::HAS_B) nb_b = sb[slot];
                                op.dyn(c, acc, self_a, self_b, sa[slot], nb_b);
                            }
                        }
                        w = wn;
                    }
                    // static neighbours: entries [cd, ct) of the same list (only near boundaries)
                    if (Op::USES_STATIC) {
                        for (uint32_t k = cd; k < ct; ++k) {
                            const uint32_t slot = unpack_slot(c.lists[list_word_index(h.pstart, h.pcount, k >> 2, tl)], k);
                            op.stat(c, acc, self_a, self_b, sstat[slot]);
                        }
                    }
                }
                const double r = op.finish(c, acc, i, self_a, self_b, active);
                if (Op::REDUCE == REDUCE_SUM) racc += r;
                if (Op::REDUCE == REDUCE_MAX) racc = fmax(racc, r);
            }
        }
        if (t + gridDim.x < ntiles) __syncthreads();
    }
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
    typedef float2 RecA;
    typedef NoRec RecB;
    static constexpr bool HAS_B = false;
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
    float* pressure;
    float stiffness;
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon&) {}
    __device__ __forceinline__ RecA load_a(const SweepCommon& c, uint32_t g) const { return c.pos[g]; }
    __device__ __forceinline__ RecB load_b(const SweepCommon&, uint32_t) const { return RecB(); }
    __device__ __forceinline__ float w(const KernelConsts& k, float r_sq, float r) const {
        if (KERNEL == 0) return wendland_w(k, r);
        if (KERNEL == 1) return poly6_w(k, r_sq);
        if (KERNEL == 2) return spiky_w(k, r);
        return cubic_w(k, r);
    }
    __device__ __forceinline__ bool init(const SweepCommon& c, Acc& a, uint32_t, RecA, RecB, uint32_t) const {
        a.dens = w(c.kc, 0.0f, 0.0f) * c.mass;  // self contribution, fluidparticleworld.rs:213
        a.gsum = f2(0.0f, 0.0f);
        a.gsq = 0.0f;
        return true;
    }
    __device__ __forceinline__ void pair(const SweepCommon& c, Acc& a, float2 pi, float2 pj) const {
        const float2 rij = pj - pi;
        const float r_sq = mag2(rij);
        const float r = sqrtf(r_sq);
        a.dens += w(c.kc, r_sq, r) * c.mass;
        if (WITH_ALPHA) {
            const float2 g = (wendland_grad_scalar(c.kc, r) * rij) * c.mass;
            a.gsum = a.gsum + g;
            a.gsq += mag2(g);
        }
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, RecA pi, RecB, RecA pj, RecB) const { pair(c, a, pi, pj); }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, RecA pi, RecB, float2 pb) const { pair(c, a, pi, pb); }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, RecA, RecB, bool) const {
        const float rho = fmaxf(a.dens, c.rho0);  // fluidparticleworld.rs:229
        dens[i] = rho;
        if (WITH_ALPHA) alpha[i] = 1.0f / fmaxf(mag2(a.gsum) + a.gsq, 1e-6f);  // dfsph.rs:94
        if (WITH_PRESSURE) pressure[i] = tait_pressure(stiffness, c.rho0, rho);  // wscsph.rs:91-92, once per particle
        return 0.0;
    }
    __device__ __forceinline__ void finalize(const SweepCommon&, double) const {}
};

// alpha only (yasph_compute_alpha: dfsph.rs:68-97 on its own)
struct OpAlphaOnly : OpDensityAlpha<0, true> {
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, RecA, RecB, bool) const {
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
    typedef float4 RecA;  // px, py, vx, vy
    typedef float RecB;   // rho
    static constexpr bool HAS_B = true;
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
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt_prev; }
    __device__ __forceinline__ RecA load_a(const SweepCommon& c, uint32_t g) const {
        const float2 p = c.pos[g], v = vel[g];
        return make_float4(p.x, p.y, v.x, v.y);
    }
    __device__ __forceinline__ RecB load_b(const SweepCommon&, uint32_t g) const { return dens[g]; }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, RecA, RecB, uint32_t) const {
        a = base_accel;
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, RecA self, RecB, RecA nb, RecB rhoj) const {
        const float2 rij = f2(nb.x - self.x, nb.y - self.y);
        const float r_sq = mag2(rij);
        const float r = vp.kind == 0 ? 0.0f : sqrtf(r_sq);  // XSPH needs r^2 only (xsph.rs:19-24)
        const float s = visc_scalar(c.kc, vp, dt, r_sq, r, rhoj);
        a = a + s * f2(nb.z - self.z, nb.w - self.w);
    }
    __device__ __forceinline__ void stat(const SweepCommon&, Acc&, RecA, RecB, float2) const {}
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, RecA self, RecB, bool) const {
        accel[i] = a;
        return (double)mag2(f2(self.z, self.w) + a * dt);  // dfsph.rs:476
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
template <int SOLVER>
struct OpJacobiA {
    typedef float4 RecA;  // px, py, predicted velocity
    typedef NoRec RecB;
    static constexpr bool HAS_B = false;
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
    __device__ __forceinline__ bool skip(const Control* ctl) const { return iter_index >= ctl->stop_iter[SOLVER]; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt; }
    __device__ __forceinline__ RecA load_a(const SweepCommon& c, uint32_t g) const {
        const float2 p = c.pos[g], v = vstar[g];
        return make_float4(p.x, p.y, v.x, v.y);
    }
    __device__ __forceinline__ RecB load_b(const SweepCommon&, uint32_t) const { return RecB(); }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, RecA, RecB, uint32_t ct) const {
        a = 0.0f;
        return SOLVER == 0 ? true : ct >= 9u;  // particle deficiency, dfsph.rs:261
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, RecA self, RecB, RecA nb, RecB) const {
        a += dot2(f2(self.z - nb.z, self.w - nb.w), wendland_grad_from_positions(c.kc, rec_pos(self), rec_pos(nb)));
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, RecA self, RecB, float2 pb) const {
        a += dot2(f2(self.z, self.w), wendland_grad_from_positions(c.kc, rec_pos(self), pb));
    }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, RecA, RecB, bool active) const {
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
        Control* ctl = c.ctl;
        const float s = (float)sum;  // f64 accumulation rounded once (DESIGN.md "residual sums")
        const uint32_t it = iter_index + 1;
        ctl->iters[SOLVER] = it;
        bool conv;
        float avg;
        if (SOLVER == 0) {
            avg = s / (float)c.n;                 // dfsph.rs:221
            const float rel = avg / c.rho0;       // dfsph.rs:222
            conv = rel * dt < sp.max_error;       // dfsph.rs:226
        } else {
            avg = s / (float)c.n / c.rho0;        // dfsph.rs:376-377
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
};

// ---------------------------------------------------------------------------------------------------------------------
// Jacobi B and warm starts
// ---------------------------------------------------------------------------------------------------------------------
template <int SOLVER, bool WARM>
struct OpJacobiB {
    typedef float2 RecA;  // position
    typedef float RecB;   // k_j
    static constexpr bool HAS_B = true;
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
    __device__ __forceinline__ bool skip(const Control* ctl) const {
        return WARM ? ctl->warm[SOLVER] == 0u : iter_index >= ctl->stop_iter[SOLVER];
    }
    __device__ __forceinline__ void prepare(const SweepCommon& c) {
        dt = c.ctl->dt;
        inv_dt = 1.0f / dt;  // dfsph.rs:132,167
    }
    __device__ __forceinline__ RecA load_a(const SweepCommon& c, uint32_t g) const { return c.pos[g]; }
    __device__ __forceinline__ RecB load_b(const SweepCommon&, uint32_t g) const {
        if (WARM) return 0.5f * fmaxf(warm[g], clamp_min);  // dfsph.rs:201-203 / 356-358
        return kfac[g];
    }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, RecA, RecB, uint32_t) const {
        a = f2(0.0f, 0.0f);
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, RecA pi, RecB ki, RecA pj, RecB kj) const {
        a = a + (ki + kj) * wendland_grad_from_positions(c.kc, pi, pj);
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, RecA pi, RecB ki, float2 pb) const {
        a = a + ki * wendland_grad_from_positions(c.kc, pi, pb);
    }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, RecA, RecB ki, bool) const {
        const float2 v = vstar[i];
        if (SOLVER == 0)
            vstar[i] = v - inv_dt * a * c.mass;  // dfsph.rs:159,191
        else
            vstar[i] = v - a * c.mass;           // dfsph.rs:312,342
        if (!WARM) warm[i] = (iter_index == 0 ? 0.0f : warm[i]) + ki;  // zeroing (dfsph.rs:206-208) fused into iteration 0
        return 0.0;
    }
    __device__ __forceinline__ void finalize(const SweepCommon&, double) const {}
};

// ---------------------------------------------------------------------------------------------------------------------
// WCSPH accelerations
// ---------------------------------------------------------------------------------------------------------------------
struct OpWcsphAccel {
    typedef float4 RecA;  // px, py, vx, vy
    typedef float2 RecB;  // rho, p
    static constexpr bool HAS_B = true;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_MAX;
    static constexpr int TICKET = 1;
    typedef float2 Acc;
    const float2* vel;
    const float* dens;
    const float* pressure;
    float2* accel;
    float2 gravity;
    ViscParams vp;
    float boundary_force_factor;
    float dt;
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt_prev; }
    __device__ __forceinline__ RecA load_a(const SweepCommon& c, uint32_t g) const {
        const float2 p = c.pos[g], v = vel[g];
        return make_float4(p.x, p.y, v.x, v.y);
    }
    __device__ __forceinline__ RecB load_b(const SweepCommon&, uint32_t g) const { return f2(dens[g], pressure[g]); }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, RecA, RecB, uint32_t) const {
        a = gravity;  // wscsph.rs:84
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, RecA self, RecB sb, RecA nb, RecB nbb) const {
        const float2 rij = f2(nb.x - self.x, nb.y - self.y);
        const float r_sq = mag2(rij);
        const float r = sqrtf(r_sq);
        const float pu = -c.mass * (sb.y + nbb.y) / (2.0f * sb.x * nbb.x);  // wscsph.rs:101
        a = a + pu * (spiky_grad_scalar(c.kc, r) * rij);                    // wscsph.rs:102
        a = a + visc_scalar(c.kc, vp, dt, r_sq, r, nbb.x) * f2(nb.z - self.z, nb.w - self.w);  // wscsph.rs:104-106
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, RecA self, RecB, float2 pb) const {
        const float2 rij = pb - rec_pos(self);
        const float r_sq = mag2(rij);
        a = a - (boundary_force_factor * spiky_w(c.kc, sqrtf(r_sq)) / r_sq) * rij;  // wscsph.rs:113-115
    }
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, RecA self, RecB, bool) const {
        accel[i] = a;
        return (double)mag2(f2(self.z, self.w) + a * dt);  // wscsph.rs:162
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
