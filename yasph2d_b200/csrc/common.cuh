// common.cuh -- shared device/host definitions for libyasph_gpu (sm_100a).
//
// Arithmetic contract: the reference computes in f32 with no FMA contraction and no fast-math.
// This translation unit is compiled with -fmad=false and CUDA's default IEEE division / square root, and every
// expression below keeps the reference's left-to-right evaluation order, so each per-particle quantity is the
// bit pattern the CPU restatement (oracle/) produces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef YASPH_TILE_LOG2
#define YASPH_TILE_LOG2 3                    // a tile is an aligned 8x8 block of cells == 64 consecutive Morton codes
#endif
#define YASPH_MAXN 64                        // neighborhood_search.rs:322
#define YASPH_MIN_DISTANCE 1.0e-10f          // neighborhood_search.rs:323
#define YASPH_NUM_SMS_B200 148

namespace yasph {

// ---- float2 helpers mirroring cgmath (component-wise, no contraction) ---------------------------------------------
// On the device the two components travel through Blackwell's packed f32x2 pipe (add/sub/mul.rn.f32x2: one instruction,
// two independently rounded IEEE results -- bit-identical to the scalar pair, half the issue slots).
#ifndef YASPH_F32X2
#define YASPH_F32X2 6  // bit 0: add, bit 1: sub, bit 2: mul.  Packed add stays off and products that feed a subtraction go through
                       // sub_scalar(): ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2 into FFMA2 (one rounding), -fmad=false notwithstanding.
                       // build.py rejects a library whose SASS contains FFMA2.
#endif
__host__ __device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
#if defined(__CUDA_ARCH__) && YASPH_F32X2
#define YASPH_PACKED2(op, a, b)                                                                                   \
    unsigned long long r_;                                                                                        \
    asm(op ".rn.f32x2 %0, %1, %2;"                                                                                \
        : "=l"(r_)                                                                                                \
        : "l"(((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x)),                           \
          "l"(((unsigned long long)__float_as_uint(b.y) << 32) | __float_as_uint(b.x)));                          \
    return make_float2(__uint_as_float((uint32_t)r_), __uint_as_float((uint32_t)(r_ >> 32)));
#endif
__host__ __device__ __forceinline__ float2 operator+(float2 a, float2 b) {
#if defined(YASPH_PACKED2) && (YASPH_F32X2 & 1)
    YASPH_PACKED2("add", a, b)
#else
    return f2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 operator-(float2 a, float2 b) {
#if defined(YASPH_PACKED2) && (YASPH_F32X2 & 2)
    YASPH_PACKED2("sub", a, b)
#else
    return f2(a.x - b.x, a.y - b.y);
#endif
}
__host__ __device__ __forceinline__ float2 mul2(float2 a, float2 b) {
#if defined(YASPH_PACKED2) && (YASPH_F32X2 & 4)
    YASPH_PACKED2("mul", a, b)
#else
    return f2(a.x * b.x, a.y * b.y);
#endif
}
__host__ __device__ __forceinline__ float2 operator*(float2 a, float s) { return mul2(a, f2(s, s)); }
__host__ __device__ __forceinline__ float2 operator*(float s, float2 a) { return mul2(f2(s, s), a); }
__host__ __device__ __forceinline__ float dot2(float2 a, float2 b) {
    const float2 p = mul2(a, b);
    return p.x + p.y;
}
__host__ __device__ __forceinline__ float mag2(float2 a) {
    const float2 p = mul2(a, a);
    return p.x + p.y;
}
__host__ __device__ __forceinline__ float2 operator/(float2 a, float s) { return f2(a.x / s, a.y / s); }
// a - b with scalar instructions: for b = (packed product), see YASPH_F32X2
__host__ __device__ __forceinline__ float2 sub_scalar(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }

// Rust f32::powi == compiler-rt __powisf2
__host__ __device__ inline float powi_f(float a, int b) {
    float r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}

// ---- Morton codes (src/sph/morton.rs:38-77) ----------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t part_1by1(uint32_t x) {
    x &= 0xffffu;
    x = (x ^ (x << 8)) & 0x00ff00ffu;
    x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u;
    x = (x ^ (x << 1)) & 0x55555555u;
    return x;
}
__host__ __device__ __forceinline__ uint32_t compact_1by1(uint32_t x) {
    x &= 0x55555555u;
    x = (x ^ (x >> 1)) & 0x33333333u;
    x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
    x = (x ^ (x >> 4)) & 0x00ff00ffu;
    x = (x ^ (x >> 8)) & 0x0000ffffu;
    return x;
}
__host__ __device__ __forceinline__ uint32_t morton_encode(uint32_t x, uint32_t y) { return (part_1by1(y) << 1) | part_1by1(x); }
__host__ __device__ __forceinline__ uint32_t morton_x(uint32_t m) { return compact_1by1(m); }
__host__ __device__ __forceinline__ uint32_t morton_y(uint32_t m) { return compact_1by1(m >> 1); }

// ---- grid (neighborhood_search.rs:45-64) ---------------------------------------------------------------------------
struct GridParams {
    float radius, radius_sq, cell_size_inv;
    float2 grid_min;
};
// Rust `f32 as u16`: truncation toward zero, saturating, NaN -> 0
__host__ __device__ __forceinline__ uint32_t f32_as_u16(float f) {
    if (!(f == f)) return 0u;
    if (f <= 0.0f) return 0u;
    if (f >= 65535.0f) return 65535u;
    return (uint32_t)f;
}
__host__ __device__ __forceinline__ uint32_t position_to_cidx(const GridParams& g, float2 p) {
    float2 c = (p - g.grid_min) * g.cell_size_inv;
    return morton_encode(f32_as_u16(c.x), f32_as_u16(c.y));
}

// ---- smoothing kernels (src/sph/smoothing_kernel/*.rs); constants precomputed on the host with the same f32 steps ----
struct KernelConsts {
    float h, h_inv, hsq;
    float wendland_norm, wendland_norm_grad;  // wendland_quintic_c2.rs:23-29
    float poly6_norm, poly6_norm_grad;        // poly6.rs:18-23
    float spiky_norm, spiky_norm_grad;        // spiky.rs:18-23
    float cubic_norm, cubic_norm_grad;        // cubic.rs:17-21
    float visc_norm_laplacian;                // viscosity.rs:24
};
#define YASPH_PI_F 3.14159274101257324219f  // std::f64::consts::PI as f32
inline KernelConsts make_kernel_consts(float h) {
    KernelConsts k;
    k.h = h;
    k.h_inv = 1.0f / h;
    k.hsq = h * h;
    k.wendland_norm = 4.0f * 7.0f / (YASPH_PI_F * powi_f(h, 2));
    k.wendland_norm_grad = 140.0f / (YASPH_PI_F * powi_f(h, 4));
    k.poly6_norm = 4.0f / (YASPH_PI_F * powi_f(h, 8));
    k.poly6_norm_grad = 24.0f / (YASPH_PI_F * powi_f(h, 8));
    k.spiky_norm = 10.0f / (YASPH_PI_F * powi_f(h, 5));
    k.spiky_norm_grad = 30.0f / (YASPH_PI_F * powi_f(h, 5));
    k.cubic_norm = 6.0f * 40.0f / (7.0f * YASPH_PI_F * h * h);
    k.cubic_norm_grad = 6.0f * 40.0f / (7.0f * YASPH_PI_F * h * h * h);
    k.visc_norm_laplacian = 360.0f / (29.0f * YASPH_PI_F * powi_f(h, 5));
    return k;
}

// wendland_quintic_c2.rs:33-38
__device__ __forceinline__ float wendland_w(const KernelConsts& k, float r) {
    float q = fminf(k.h_inv * r, 1.0f);
    float omq = 1.0f - q;
    float omq2 = omq * omq;
    return k.wendland_norm * omq2 * omq2 * (q + 0.25f);
}
// wendland_quintic_c2.rs:41-46 : scalar s with gradient = s * (rj - ri)
__device__ __forceinline__ float wendland_grad_scalar(const KernelConsts& k, float r) {
    float q = fminf(r * k.h_inv, 1.0f);
    float omq = 1.0f - q;
    return k.wendland_norm_grad * omq * omq * omq;
}
// kernel.rs:22-28 with the Wendland kernel
__device__ __forceinline__ float2 wendland_grad_from_positions(const KernelConsts& k, float2 ri, float2 rj) {
    float2 rij = rj - ri;
    float r = sqrtf(mag2(rij));
    return wendland_grad_scalar(k, r) * rij;
}
// poly6.rs:28-31
__device__ __forceinline__ float poly6_w(const KernelConsts& k, float r_sq) {
    float d = fmaxf(k.hsq - r_sq, 0.0f);
    return k.poly6_norm * d * d * d;
}
// spiky.rs:28-31
__device__ __forceinline__ float spiky_w(const KernelConsts& k, float r) {
    float d = fmaxf(k.h - r, 0.0f);
    return k.spiky_norm * d * d * d;
}
// spiky.rs:34-37
__device__ __forceinline__ float spiky_grad_scalar(const KernelConsts& k, float r) {
    float d = fmaxf(k.h - r, 0.0f);
    return k.spiky_norm_grad * d * d / (r + 1.0e-10f);
}
// cubic.rs:24-37
__device__ __forceinline__ float cubic_w(const KernelConsts& k, float r) {
    float q = r * k.h_inv;
    if (q <= 0.5f) {
        float q2 = q * q;
        return k.cubic_norm * ((1.0f / 6.0f) + q2 * q - q2);
    } else if (q <= 1.0f) {
        float omq = 1.0f - q;
        return k.cubic_norm * omq * omq * omq * (2.0f / 6.0f);
    }
    return 0.0f;
}
// viscosity.rs:45-47
__device__ __forceinline__ float visc_laplacian(const KernelConsts& k, float r) { return k.visc_norm_laplacian * (k.h - r); }

// ---- TimeManager arithmetic (timemanager.rs:252-279) on integer nanoseconds -------------------------------------------
// std::time::Duration::from_secs_f32, round-to-nearest (Rust >= 1.67)
__host__ __device__ inline uint64_t duration_from_secs_f32(float s) {
    if (!(s >= 0.0f)) return 0ull;
    double ns = rint((double)s * 1e9);
    if (ns > 1.8e19) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)ns;
}
// std::time::Duration::as_secs_f32
__host__ __device__ inline float duration_as_secs_f32(uint64_t ns) {
    uint64_t secs = ns / 1000000000ull;
    uint32_t nanos = (uint32_t)(ns % 1000000000ull);
    return (float)secs + (float)nanos / 1.0e9f;
}
struct TimeParams {
    int adaptive;
    uint64_t fixed_ns, min_ns, max_ns;
    uint64_t target_ns;  // AdaptiveTimeStepTarget::TargetFrameLength, 0 = None
    float cfl_factor;
};
// total_ns: TimeManager::total_simulated_time at the time of the call (it already includes the running step, timemanager.rs:246)
__host__ __device__ inline uint64_t update_simulation_step(const TimeParams& t, uint64_t prev_ns, float particle_diameter, float max_velocity,
                                                           uint64_t total_ns) {
    if (!t.adaptive) return t.fixed_ns;
    uint64_t time_cfl = duration_from_secs_f32(t.cfl_factor * 0.4f * particle_diameter / (max_velocity + 0.00001f));
    uint64_t upper = t.max_ns < prev_ns * 2 ? t.max_ns : prev_ns * 2;
    uint64_t lower = t.min_ns;
    if (t.target_ns) {  // timemanager.rs:268-272: total - target * ((total / target) as u32), then min with timestep_min
        const uint64_t time_to_target = total_ns - t.target_ns * (uint64_t)(uint32_t)(total_ns / t.target_ns);
        lower = lower < time_to_target ? lower : time_to_target;
    }
    uint64_t m = upper < time_cfl ? upper : time_cfl;
    return lower > m ? lower : m;
}

// ---- device-resident control block: everything the step decides on the device ------------------------------------------
struct Control {
    // time
    unsigned long long step_ns;       // TimeManager::simulation_step
    unsigned long long step_prev_ns;  // value at step entry
    double resid_sum;                 // Jacobi residual sum of the running iteration (slab mode: all-reduced in place over the ranks)
    float dt;                         // step_ns as_secs_f32 (after update) -- the dt of this step
    float dt_prev;                    // dt at step entry (viscosity, CFL estimate)
    unsigned int max_v2_bits;         // max |v + a dt|^2 as uint bits (value >= 0 so uint order == float order)
    float max_velocity;
    // counts produced by the neighbourhood update
    unsigned int num_cells, num_tiles;
    unsigned int num_cells_static, num_tiles_static;
    // per-tile maxima of the last tile-table build: they size the shared memory of every tile kernel
    unsigned int max_dyn_total, max_stat_total, max_pcount;
    unsigned int max_nk;              // most list words of any particle (written by the list build; sizes the list staging of the NEXT sweeps)
    // list-build statistics: 16 contiguous bytes, zeroed together before every list build
    unsigned long long total_neighbors;
    unsigned int capped, dropped;
    // Jacobi loop state, index 0 = density solver, 1 = divergence solver
    unsigned int iters[2];            // num_*_correction_iterations (persist across steps: decide the warm start)
    unsigned int stop_iter[2];        // iteration count at which the running solve stops (0xFFFFFFFF while undecided)
    unsigned int warm[2];             // 1 if the warm start of the running solve is enabled
    float avg[2];
    unsigned int not_converged;
    unsigned int nonfinite;
    // error flags
    unsigned int err_tile_capacity;   // a tile stages more than 65535 candidates (slots are 16 bit) or a cell holds more than 65535 particles
    unsigned int err_tile_count;      // more tiles than max_tiles
    // last-block tickets
    unsigned int ticket[4];
    // tile queue of the persistent sweeps (sweeps.cuh): tiles handed out beyond the first one per CTA, CTAs that have finished
    unsigned int tile_next, cta_done;
    // slab mode (multi-GPU): counts of the ordered selections of a neighbourhood update
    unsigned long long slab_migrants;    // owned particles counted by a mid-step selection (yasph_step_host_slab: checked with the step's last control block)
    unsigned long long slab_ghost_send;  // own particles in the first / last owned column (sent as ghosts)
    unsigned long long slab_send;        // per-pass halo send lists (left | right), in sorted order
    unsigned long long slab_ghost;       // ghosts from the left | right rank, in sorted order
    unsigned long long slab_own;         // owned particles (low word)
    unsigned int err_slab;               // a particle arrived that this rank does not own (moved more than one slab in a step)
    unsigned long long total_simulated_ns;  // TimeManager::total_simulated_time before the frame loop's addition for the running step
    unsigned int total_is_current;          // 1: the host supplied the total incl. the running step (k_begin_step must not add it again)
    unsigned int step_token;                // set by k_begin_step: the step whose head (k_begin_step + first pass) may run (yasph_step_n)
    unsigned int err_comm;               // peer-memory transport: bit 0 a halo message, bit 1 an all-reduce contribution did not arrive in time
    unsigned int slab_sel4[4];           // slab mode: counts of the four-way selection of an update (migrants left | right, ghost layer left | right)
    unsigned int slab_cnt[10];           // SlabCounts (slab.cuh) of the running particle exchange
};

// slab decomposition parameters handed to the key-generating kernels (active == 0: single GPU)
struct SlabParams {
    uint32_t col_lo, col_hi;
    uint8_t* pflag;  // in: 1 = ghost of the current structure; out: 0 stays, 1 / 2 migrates left / right, 3 dropped ghost
    int active;
};
enum { SLAB_STAY = 0, SLAB_MIG_LEFT = 1, SLAB_MIG_RIGHT = 2, SLAB_DROP_GHOST = 3 };
#define YASPH_KEY_DROPPED 0xFFFFFFFFu  // sort key of particles that leave the local set (they sort to the end and are cut off)
// classification of a local particle with (new) cell key `key`; returns the key to sort by
__device__ __forceinline__ uint32_t slab_classify(const SlabParams& sp, uint32_t i, uint32_t key) {
    if (!sp.active) return key;
    const uint32_t col = compact_1by1(key);
    uint8_t out = SLAB_STAY;
    if (sp.pflag[i])
        out = SLAB_DROP_GHOST;
    else if (col < sp.col_lo)
        out = SLAB_MIG_LEFT;
    else if (col >= sp.col_hi)
        out = SLAB_MIG_RIGHT;
    sp.pflag[i] = out;
    return out ? YASPH_KEY_DROPPED : key;
}

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its predecessor in the stream still runs; pdl_enter() blocks until that predecessor has completed and its memory
// is visible, then lets the kernel's own successor become resident the same way.  Everything ahead of pdl_enter() in a kernel
// touches registers and shared memory only.  Launch latency, CTA placement and the shared-memory set-up of kernel N + 1 thus
// overlap the tail of kernel N.  In a kernel launched without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// warp helpers
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace yasph
