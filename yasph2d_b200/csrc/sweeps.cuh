// sweeps.cuh -- every neighbour-dependent per-particle pass of the WCSPH / DFSPH solvers as one tile-staged kernel.
//
// Skeleton (k_sweep): one CTA per 8x8-cell tile (persistent grid-stride loop).  The tile's own particles plus the
// 1-cell apron are staged into shared memory once (positions + the per-pass neighbour payload, boundary positions),
// then one thread per particle walks its compact neighbour list (u16 shared-memory slots, dynamic first, then
// static) in the reference's order.  Reductions (Jacobi residual sum, CFL max) are fused: per-thread accumulation,
// block reduce, per-CTA partial, last-arriving CTA combines the partials in fixed order and takes the convergence
// decision on the device (no host round trip inside a Jacobi iteration).
//
// Pass -> reference:
//   OpDensityAlpha   FluidParticleWorld::update_densities (fluidparticleworld.rs:197-231) fused with
//                    DFSPHSolver::compute_alpha_factors (dfsph.rs:68-97)
//   OpViscosity      non-pressure forces (dfsph.rs:436-469) + max |v + a dt|^2 (dfsph.rs:474-477)
//   OpJacobiA        compute_density_error (dfsph.rs:99-126) / compute_density_change (dfsph.rs:249-280) + residual sum
//                    and loop decision (dfsph.rs:219-245, 374-400)
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

struct NoPayload {};

template <class Op>
__global__ void __launch_bounds__(SW_THREADS) k_sweep(SweepCommon c, Op op) {
    typedef typename Op::Payload Payload;
    if (op.skip(c.ctl)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& ts = *reinterpret_cast<TileSmem*>(smem_raw);
    Payload* spay = reinterpret_cast<Payload*>(smem_raw + sizeof(TileSmem));
    float2* sdyn = reinterpret_cast<float2*>(smem_raw + sizeof(TileSmem) + (Op::HAS_PAYLOAD ? sizeof(Payload) * (size_t)c.cap_dyn : 0));
    float2* sstat = sdyn + c.cap_dyn;
    op.prepare(c);
    const uint32_t ntiles = c.ctl->num_tiles;
    double racc = 0.0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        load_tile_tables(ts, c.tt, t);
        __syncthreads();
        const TileHeader h = ts.hdr;
        if (h.dyn_total <= c.cap_dyn && h.stat_total <= c.cap_stat) {
            for (uint32_t s = threadIdx.x; s < h.dyn_total; s += blockDim.x) {
                const uint32_t g = dyn_slot_to_global(ts, s);
                sdyn[s] = c.pos[g];
                if (Op::HAS_PAYLOAD) spay[s] = op.load(c, g);
            }
            if (Op::USES_STATIC)
                for (uint32_t s = threadIdx.x; s < h.stat_total; s += blockDim.x) sstat[s] = c.bpos[slot_to_global(ts.stat, s)];
            __syncthreads();
            for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += blockDim.x) {
                const uint32_t i = h.pstart + tl;
                const uint32_t own = h.own_lo + tl;
                const float2 pi = sdyn[own];
                Payload self;
                if (Op::HAS_PAYLOAD) self = spay[own];
                const uchar2 cnt = c.counts[i];
                const uint32_t cd = cnt.x, ct = Op::USES_STATIC ? cnt.y : cnt.x;
                typename Op::Acc acc;
                const bool active = op.init(c, acc, i, pi, self, cnt.y);
                if (active) {
                    const uint32_t nkb = (ct + 3u) >> 2;
                    unsigned long long w = nkb ? c.lists[list_word_index(h.pstart, h.pcount, 0, tl)] : 0ull;
                    for (uint32_t kb = 0; kb < nkb; ++kb) {
                        const unsigned long long wn = (kb + 1 < nkb) ? c.lists[list_word_index(h.pstart, h.pcount, kb + 1, tl)] : 0ull;
#pragma unroll
                        for (uint32_t q = 0; q < 4; ++q) {
                            const uint32_t k = kb * 4 + q;
                            const uint32_t slot = (uint32_t)(w >> (16 * q)) & 0xFFFFu;
                            if (k < cd) {
                                Payload pj;
                                if (Op::HAS_PAYLOAD) pj = spay[slot];
                                op.dyn(c, acc, pi, self, sdyn[slot], pj);
                            } else if (k < ct) {
                                op.stat(c, acc, pi, self, sstat[slot]);
                            }
                        }
                        w = wn;
                    }
                }
                const double r = op.finish(c, acc, i, pi, self, active);
                if (Op::REDUCE == REDUCE_SUM) racc += r;
                if (Op::REDUCE == REDUCE_MAX) racc = fmax(racc, r);
            }
        }
        __syncthreads();
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
// density (+ alpha)
// ---------------------------------------------------------------------------------------------------------------------
template <int KERNEL, bool WITH_ALPHA>
struct OpDensityAlpha {
    typedef NoPayload Payload;
    static constexpr bool HAS_PAYLOAD = false;
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
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon&) {}
    __device__ __forceinline__ Payload load(const SweepCommon&, uint32_t) const { return Payload(); }
    __device__ __forceinline__ float w(const KernelConsts& k, float r_sq, float r) const {
        if (KERNEL == 0) return wendland_w(k, r);
        if (KERNEL == 1) return poly6_w(k, r_sq);
        if (KERNEL == 2) return spiky_w(k, r);
        return cubic_w(k, r);
    }
    __device__ __forceinline__ bool init(const SweepCommon& c, Acc& a, uint32_t, float2, Payload, uint32_t) const {
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
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, Payload, float2 pj, Payload) const { pair(c, a, pi, pj); }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, Payload, float2 pb) const { pair(c, a, pi, pb); }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, Payload, bool) const {
        dens[i] = fmaxf(a.dens, c.rho0);  // fluidparticleworld.rs:229
        if (WITH_ALPHA) alpha[i] = 1.0f / fmaxf(mag2(a.gsum) + a.gsq, 1e-6f);  // dfsph.rs:94
        return 0.0;
    }
    __device__ __forceinline__ void finalize(const SweepCommon&, double) const {}
};

// alpha only (yasph_compute_alpha: dfsph.rs:68-97 on its own)
struct OpAlphaOnly : OpDensityAlpha<0, true> {
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, Payload, bool) const {
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
    typedef float4 Payload;  // vx, vy, rho, -
    static constexpr bool HAS_PAYLOAD = true;
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
    __device__ __forceinline__ Payload load(const SweepCommon&, uint32_t g) const {
        const float2 v = vel[g];
        return make_float4(v.x, v.y, dens[g], 0.0f);
    }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, Payload, uint32_t) const {
        a = base_accel;
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, Payload self, float2 pj, Payload nb) const {
        const float2 rij = pj - pi;
        const float r_sq = mag2(rij);
        const float r = sqrtf(r_sq);
        const float s = visc_scalar(c.kc, vp, dt, r_sq, r, nb.z);
        a = a + s * f2(nb.x - self.x, nb.y - self.y);
    }
    __device__ __forceinline__ void stat(const SweepCommon&, Acc&, float2, Payload, float2) const {}
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, Payload self, bool) const {
        accel[i] = a;
        return (double)mag2(f2(self.x, self.y) + a * dt);  // dfsph.rs:476
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
    typedef float2 Payload;  // predicted velocity
    static constexpr bool HAS_PAYLOAD = true;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_SUM;
    static constexpr int TICKET = 2;
    typedef float Acc;
    const float2* vstar;
    const float* dens;
    float* err;
    SolverParams sp;
    uint32_t iter_index;
    float dt;
    __device__ __forceinline__ bool skip(const Control* ctl) const { return iter_index >= ctl->stop_iter[SOLVER]; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt; }
    __device__ __forceinline__ Payload load(const SweepCommon&, uint32_t g) const { return vstar[g]; }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, Payload, uint32_t ct) const {
        a = 0.0f;
        return SOLVER == 0 ? true : ct >= 9u;  // particle deficiency, dfsph.rs:261
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, Payload vi, float2 pj, Payload vj) const {
        a += dot2(vi - vj, wendland_grad_from_positions(c.kc, pi, pj));
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, Payload vi, float2 pb) const {
        a += dot2(vi, wendland_grad_from_positions(c.kc, pi, pb));
    }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, Payload, bool active) const {
        float e;
        if (SOLVER == 0) {
            e = dens[i] + a * c.mass * dt;       // dfsph.rs:121
            e = fmaxf(c.rho0, e) - c.rho0;       // dfsph.rs:124
        } else {
            e = active ? fmaxf(a * c.mass, 0.0f) : 0.0f;  // dfsph.rs:262,277-278
        }
        err[i] = e;
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
    typedef float Payload;  // k_j
    static constexpr bool HAS_PAYLOAD = true;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_NONE;
    static constexpr int TICKET = 0;
    typedef float2 Acc;
    float2* vstar;
    const float* err;
    const float* alpha;
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
    __device__ __forceinline__ Payload load(const SweepCommon&, uint32_t g) const {
        if (WARM) return 0.5f * fmaxf(warm[g], clamp_min);  // dfsph.rs:201-203 / 356-358
        return err[g] * alpha[g];                           // dfsph.rs:141,150 / 295,304
    }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, Payload, uint32_t) const {
        a = f2(0.0f, 0.0f);
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, Payload ki, float2 pj, Payload kj) const {
        a = a + (ki + kj) * wendland_grad_from_positions(c.kc, pi, pj);
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, Payload ki, float2 pb) const {
        a = a + ki * wendland_grad_from_positions(c.kc, pi, pb);
    }
    __device__ __forceinline__ double finish(const SweepCommon& c, Acc& a, uint32_t i, float2, Payload ki, bool) const {
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
__device__ __forceinline__ float tait_pressure(float stiffness, float rho0, float rho) {  // wscsph.rs:52-57
    return stiffness * (powi_f(fmaxf(rho / rho0, 1.0f), 7) - 1.0f);
}
struct OpWcsphAccel {
    typedef float4 Payload;  // vx, vy, rho, p
    static constexpr bool HAS_PAYLOAD = true;
    static constexpr bool USES_STATIC = true;
    static constexpr int REDUCE = REDUCE_MAX;
    static constexpr int TICKET = 1;
    typedef float2 Acc;
    const float2* vel;
    const float* dens;
    float2* accel;
    float2 gravity;
    ViscParams vp;
    float stiffness, boundary_force_factor;
    float dt;
    __device__ __forceinline__ bool skip(const Control*) const { return false; }
    __device__ __forceinline__ void prepare(const SweepCommon& c) { dt = c.ctl->dt_prev; }
    __device__ __forceinline__ Payload load(const SweepCommon& c, uint32_t g) const {
        const float2 v = vel[g];
        const float rho = dens[g];
        return make_float4(v.x, v.y, rho, tait_pressure(stiffness, c.rho0, rho));
    }
    __device__ __forceinline__ bool init(const SweepCommon&, Acc& a, uint32_t, float2, Payload, uint32_t) const {
        a = gravity;  // wscsph.rs:84
        return true;
    }
    __device__ __forceinline__ void dyn(const SweepCommon& c, Acc& a, float2 pi, Payload self, float2 pj, Payload nb) const {
        const float2 rij = pj - pi;
        const float r_sq = mag2(rij);
        const float r = sqrtf(r_sq);
        const float pu = -c.mass * (self.w + nb.w) / (2.0f * self.z * nb.z);  // wscsph.rs:101
        a = a + pu * (spiky_grad_scalar(c.kc, r) * rij);                      // wscsph.rs:102
        a = a + visc_scalar(c.kc, vp, dt, r_sq, r, nb.z) * f2(nb.x - self.x, nb.y - self.y);  // wscsph.rs:104-106
    }
    __device__ __forceinline__ void stat(const SweepCommon& c, Acc& a, float2 pi, Payload, float2 pb) const {
        const float2 rij = pb - pi;
        const float r_sq = mag2(rij);
        a = a - (boundary_force_factor * spiky_w(c.kc, sqrtf(r_sq)) / r_sq) * rij;  // wscsph.rs:113-115
    }
    __device__ __forceinline__ double finish(const SweepCommon&, Acc& a, uint32_t i, float2, Payload self, bool) const {
        accel[i] = a;
        return (double)mag2(f2(self.x, self.y) + a * dt);  // wscsph.rs:162
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
