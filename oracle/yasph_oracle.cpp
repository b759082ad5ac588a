// yasph_oracle.cpp -- CPU restatement of the yasph2d per-step SPH hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
// yasph2d_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  The product path never calls into it.
//
// It restates, function by function and in f32 with FMA contraction disabled
// (-ffp-contract=off), the algorithm of the reference (paths relative to /root/reference):
//   src/sph/morton.rs, src/sph/neighborhood_search.rs, src/sph/fluidparticleworld.rs,
//   src/sph/solver/dfsph.rs, src/sph/solver/wscsph.rs, src/sph/smoothing_kernel/*.rs,
//   src/sph/viscositymodel/*.rs, src/sph/timemanager.rs:104-138,252-279.
// Each function cites the file:line it follows.
//
// PINNING.  The reference cannot be compiled here (no Rust toolchain).  The reference's own
// tests pin: Morton encode/decode/BIGMIN known answers (morton.rs:189-251), the brute-force
// neighbour-list equality property (neighborhood_search.rs:529-556) and the smoothing-kernel
// properties (smoothing_kernel/kernel.rs:40-164).  tests/test_oracle_*.py restate all of them
// against this file.  The reference has NO tests for update_densities, the DFSPH / WCSPH solvers or
// the TimeManager: for those passes parity is UNPINNED by the reference (this oracle is the pin).
//
// Documented deviations from the reference (all forced by third-party behaviour that is not pinned):
//   D1  the particle sort is STABLE (ties keep their previous relative order); the reference uses
//       rayon's unstable par_sort_unstable_by_key (neighborhood_search.rs:116-118).
//   D2  the Jacobi residual sums (dfsph.rs:221,376) are accumulated in f64 and rounded to f32;
//       the reference uses rayon's order-nondeterministic f32 sum.
//   D3  jitter PRNG = xoshiro256++ seeded by SplitMix64 (what rand 0.8 SmallRng is on 64-bit
//       targets, restated from its published algorithm; unverifiable offline).
//   D4  when the dynamic neighbours already fill all 64 slots and a static candidate passes the
//       distance test the reference indexes neighbor_set[64] and panics
//       (neighborhood_search.rs:373); here the candidate is dropped and counted.
//   D5  Duration::from_secs_f32 rounds to the nearest nanosecond (Rust >= 1.67 semantics).
//
// Parallelism: OpenMP `parallel for schedule(static)` exactly where the reference uses rayon
// (SURVEY.md 2.3), sequential where the reference is sequential.  Per-particle results do not
// depend on the thread count.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef float Real;
struct V2 {
    Real x, y;
};
static inline V2 v2(Real x, Real y) { return V2{x, y}; }
static inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
static inline V2 operator-(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
static inline V2 operator*(V2 a, Real s) { return V2{a.x * s, a.y * s}; }
static inline V2 operator*(Real s, V2 a) { return V2{s * a.x, s * a.y}; }
static inline V2 operator/(V2 a, Real s) { return V2{a.x / s, a.y / s}; }
static inline V2 operator-(V2 a) { return V2{-a.x, -a.y}; }
static inline Real dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline Real magnitude2(V2 a) { return a.x * a.x + a.y * a.y; }
// cgmath: distance2(a, b) = (b - a).magnitude2()
static inline Real distance2(V2 a, V2 b) { return magnitude2(b - a); }
// Rust f32::max / f32::min ignore NaN (return the other operand) == fmaxf / fminf.
static inline Real rmax(Real a, Real b) { return fmaxf(a, b); }
static inline Real rmin(Real a, Real b) { return fminf(a, b); }

// Rust f32::powi -> llvm.powi -> compiler-rt __powisf2 (square-and-multiply, LSB first).
static Real powi(Real a, int b) {
    const bool recip = b < 0;
    Real r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}
static const Real PI_F = (Real)3.14159265358979323846;  // std::f64::consts::PI as f32

// ---------------------------------------------------------------------------------------------
// morton.rs
// ---------------------------------------------------------------------------------------------
static const uint32_t MORTON_XBITS = 0x55555555u;  // morton.rs:1
static const uint32_t MORTON_YBITS = 0xAAAAAAAAu;  // morton.rs:2

// morton.rs:38-45
static inline uint32_t part_1by1(uint16_t x16) {
    uint32_t x = x16;
    x = (x ^ (x << 8)) & 0x00ff00ffu;
    x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u;
    x = (x ^ (x << 1)) & 0x55555555u;
    return x;
}
// morton.rs:49-51 (encode_bitfiddle) == morton.rs:85-110 (encode_lookup, the one that ships)
static inline uint32_t morton_encode(uint16_t x, uint16_t y) { return (part_1by1(y) << 1) + part_1by1(x); }
// morton.rs:85-110 restated with a table built from part_1by1 of a byte (== MORTON_TABLE256)
static uint32_t morton_encode_lookup(uint16_t x, uint16_t y) {
    static uint16_t table[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; ++i) table[i] = (uint16_t)part_1by1((uint16_t)i);
        init = true;
    }
    return ((uint32_t)table[y >> 8] << 17) | ((uint32_t)table[x >> 8] << 16) | ((uint32_t)table[y & 0xFF] << 1) |
           (uint32_t)table[x & 0xFF];
}
// morton.rs:57-65
static inline uint32_t compact_1by1(uint32_t x) {
    x &= 0x55555555u;
    x = (x ^ (x >> 1)) & 0x33333333u;
    x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
    x = (x ^ (x >> 4)) & 0x00ff00ffu;
    x = (x ^ (x >> 8)) & 0x0000ffffu;
    return x;
}
static inline uint32_t morton_decode_x(uint32_t m) { return compact_1by1(m); }       // morton.rs:69-71
static inline uint32_t morton_decode_y(uint32_t m) { return compact_1by1(m >> 1); }  // morton.rs:75-77

// morton.rs:123-128
static inline bool is_in_rect_presplit(uint32_t m, uint32_t minx, uint32_t miny, uint32_t maxx, uint32_t maxy) {
    uint32_t cx = m & MORTON_XBITS, cy = m & MORTON_YBITS;
    return cx >= minx && cy >= miny && cx <= maxx && cy <= maxy;
}
// morton.rs:137-141
static inline uint32_t load_bits(uint32_t pattern, uint32_t patternlen, uint32_t value, uint32_t dim) {
    uint32_t wipe_mask = ~(part_1by1((uint16_t)(0xffffu >> (16 - (patternlen / 2 + 1)))) << dim);
    uint32_t p = part_1by1((uint16_t)pattern) << dim;
    return (value & wipe_mask) | p;
}
// morton.rs:151-182 (Tropf-Herzog BIGMIN decision table)
static uint32_t find_bigmin(uint32_t m_cur, uint32_t min_morton, uint32_t max_morton) {
    uint32_t bigmin = 0;
    for (int bitpos = 31; bitpos >= 0; --bitpos) {
        uint32_t setbit = 1u << bitpos;
        bool curbit = (m_cur & setbit) != 0, minbit = (min_morton & setbit) != 0, maxbit = (max_morton & setbit) != 0;
        uint32_t dim = (uint32_t)bitpos % 2, mask = 1u << (bitpos / 2);
        if (!curbit && !minbit && !maxbit) {
        } else if (!curbit && !minbit && maxbit) {
            bigmin = load_bits(mask, (uint32_t)bitpos, min_morton, dim);
            max_morton = load_bits(mask - 1, (uint32_t)bitpos, max_morton, dim);
        } else if (!curbit && minbit && maxbit) {
            return min_morton;
        } else if (curbit && !minbit && !maxbit) {
            return bigmin;
        } else if (curbit && !minbit && maxbit) {
            min_morton = load_bits(mask, (uint32_t)bitpos, min_morton, dim);
        } else if (curbit && minbit && maxbit) {
        } else {
            // (false,true,false) and (true,true,false): unreachable_unchecked in the reference
            return bigmin;
        }
    }
    return bigmin;
}

// ---------------------------------------------------------------------------------------------
// smoothing_kernel/*.rs
// ---------------------------------------------------------------------------------------------
enum KernelId { K_WENDLAND = 0, K_POLY6 = 1, K_SPIKY = 2, K_CUBIC = 3, K_VISCOSITY = 4 };
static const Real DIVISION_EPSILON = 1.0e-10f;  // kernel.rs:9

struct Wendland {  // wendland_quintic_c2.rs:16-46
    Real h_inv, normalizer, normalizer_grad;
    explicit Wendland(Real h) {
        h_inv = 1.0f / h;
        normalizer = 4.0f * 7.0f / (PI_F * powi(h, 2));
        normalizer_grad = 140.0f / (PI_F * powi(h, 4));
    }
    inline Real evaluate(Real, Real r) const {
        Real q = rmin(h_inv * r, 1.0f);
        Real omq = 1.0f - q;
        Real omq2 = omq * omq;
        return normalizer * omq2 * omq2 * (q + 0.25f);
    }
    inline V2 gradient(V2 rij, Real, Real r) const {
        Real q = rmin(r * h_inv, 1.0f);
        Real omq = 1.0f - q;
        return (normalizer_grad * omq * omq * omq) * rij;
    }
};
struct Poly6 {  // poly6.rs:15-37
    Real hsq, normalizer, normalizer_grad;
    explicit Poly6(Real h) {
        hsq = h * h;
        normalizer = 4.0f / (PI_F * powi(h, 8));
        normalizer_grad = 24.0f / (PI_F * powi(h, 8));
    }
    inline Real evaluate(Real r_sq, Real) const {
        Real d = rmax(hsq - r_sq, 0.0f);
        return normalizer * d * d * d;
    }
    inline V2 gradient(V2 rij, Real r_sq, Real) const {
        Real d = rmax(hsq - r_sq, 0.0f);
        return normalizer_grad * d * d * rij;
    }
};
struct Spiky {  // spiky.rs:15-37
    Real h, normalizer, normalizer_grad;
    explicit Spiky(Real h_) {
        h = h_;
        normalizer = 10.0f / (PI_F * powi(h, 5));
        normalizer_grad = 30.0f / (PI_F * powi(h, 5));
    }
    inline Real evaluate(Real, Real r) const {
        Real d = rmax(h - r, 0.0f);
        return normalizer * d * d * d;
    }
    inline V2 gradient(V2 rij, Real, Real r) const {
        Real d = rmax(h - r, 0.0f);
        return (normalizer_grad * d * d / (r + DIVISION_EPSILON)) * rij;
    }
};
struct CubicSpline {  // cubic.rs:15-51
    Real h_inv, normalizer, normalizer_grad;
    explicit CubicSpline(Real h) {
        h_inv = 1.0f / h;
        normalizer = 6.0f * 40.0f / (7.0f * PI_F * h * h);
        normalizer_grad = 6.0f * 40.0f / (7.0f * PI_F * h * h * h);
    }
    inline Real evaluate(Real, Real r) const {
        Real q = r * h_inv;
        if (q <= 0.5f) {
            Real q2 = q * q;
            return normalizer * ((1.0f / 6.0f) + q2 * q - q2);
        } else if (q <= 1.0f) {
            Real omq = 1.0f - q;
            return normalizer * omq * omq * omq * (2.0f / 6.0f);
        }
        return 0.0f;
    }
    inline V2 gradient(V2 rij, Real, Real r) const {
        Real q = r * h_inv;
        if (q <= 0.5f) return normalizer_grad * q * (2.0f - q * 3.0f) / r * rij;
        if (q < 1.0f) {
            Real f = 1.0f - q;
            return normalizer_grad * f * f / r * rij;
        }
        return v2(0, 0);
    }
};
struct ViscosityKernel {  // viscosity.rs:11-47
    Real h, hsq, normalizer, normalizer_laplacian;
    explicit ViscosityKernel(Real h_) {
        h = h_;
        hsq = h * h;
        normalizer = 90.0f / (29.0f * PI_F * h * h);
        normalizer_laplacian = 360.0f / (29.0f * PI_F * powi(h, 5));
    }
    inline Real evaluate(Real r_sq, Real r) const {
        if (r < h) return normalizer * (4.0f * r_sq * r / (9.0f * h) + r_sq) / hsq;
        return 0.0f;
    }
    inline Real laplacian(Real, Real r) const { return normalizer_laplacian * (h - r); }
};
// kernel.rs:22-28
template <class K>
static inline V2 gradient_from_positions(const K& k, V2 ri, V2 rj) {
    V2 rij = rj - ri;
    Real r_sq = magnitude2(rij);
    Real r = sqrtf(r_sq);
    return k.gradient(rij, r_sq, r);
}

// ---------------------------------------------------------------------------------------------
// viscositymodel/*.rs
// ---------------------------------------------------------------------------------------------
enum ViscKind { VISC_XSPH = 0, VISC_PHYSICAL = 1 };
struct ViscosityModel {
    int kind;
    Real param;  // epsilon (xsph.rs:14, default 0.05) or fluid_viscosity mu (physical.rs:15)
    Poly6 poly6;
    ViscosityKernel visc;
    ViscosityModel(int kind_, Real param_, Real h) : kind(kind_), param(param_), poly6(h), visc(h) {}
    inline V2 accel(Real dt, Real r_sq, Real r, Real massj, Real rhoj, V2 vdiff) const {
        if (kind == VISC_XSPH)  // xsph.rs:21-23
            return param * massj * poly6.evaluate(r_sq, r) / (rhoj * dt) * vdiff;
        // physical.rs:21-23
        return param * massj * visc.laplacian(r_sq, r) / rhoj * vdiff;
    }
};

// ---------------------------------------------------------------------------------------------
// timemanager.rs (only the parts the solvers touch)
// ---------------------------------------------------------------------------------------------
static uint64_t duration_from_secs_f32(Real s) {  // std Duration::from_secs_f32 (deviation D5)
    if (!(s >= 0.0f)) return 0;                   // the reference would panic on negative / NaN
    double ns = nearbyint((double)s * 1e9);
    if (ns > 1.8e19) return UINT64_MAX;
    return (uint64_t)ns;
}
static Real duration_as_secs_f32(uint64_t ns) {  // std Duration::as_secs_f32
    uint64_t secs = ns / 1000000000ull;
    uint32_t nanos = (uint32_t)(ns % 1000000000ull);
    return (Real)secs + (Real)nanos / 1.0e9f;
}
struct TimeManager {
    int adaptive;  // 0 = FixedTimeStep, 1 = AdaptiveTimeStep
    uint64_t fixed_ns, min_ns, max_ns;
    Real cfl_factor;
    uint64_t simulation_step_ns;
    uint64_t target_ns = 0;             // AdaptiveTimeStepTarget: 0 = None, else TargetFrameLength (timemanager.rs:23-36)
    uint64_t total_simulated_ns = 0;    // timemanager.rs:92
    // timemanager.rs:105-109, 131-133
    void reset() {
        simulation_step_ns = adaptive ? min_ns : fixed_ns;
        total_simulated_ns = 0;
    }
    // timemanager.rs:136-138
    uint64_t simulation_step() const { return simulation_step_ns; }
    // the bookkeeping of simulation_frame_loop when it decides to perform a step (timemanager.rs:243-247); the frame pacing
    // around it (render-time comparison, step dropping) belongs to the application
    void perform_step() { total_simulated_ns += simulation_step_ns; }
    // timemanager.rs:252-279
    uint64_t update_simulation_step(Real particle_diameter, Real max_velocity) {
        if (!adaptive) {
            simulation_step_ns = fixed_ns;
        } else {
            const Real VELOCITY_EPSILON = 0.00001f;
            uint64_t time_cfl = duration_from_secs_f32(cfl_factor * 0.4f * particle_diameter / (max_velocity + VELOCITY_EPSILON));
            uint64_t upper = std::min(max_ns, simulation_step_ns * 2);
            uint64_t lower = min_ns;
            if (target_ns) {  // timemanager.rs:268-272
                uint64_t time_to_target = total_simulated_ns - target_ns * (uint64_t)(uint32_t)(total_simulated_ns / target_ns);
                lower = std::min(min_ns, time_to_target);
            }
            simulation_step_ns = std::max(lower, std::min(upper, time_cfl));
        }
        return simulation_step_ns;
    }
};

// ---------------------------------------------------------------------------------------------
// neighborhood_search.rs
// ---------------------------------------------------------------------------------------------
struct MortonCell {  // neighborhood_search.rs:33-37
    uint32_t first_particle, cidx;
};
struct GridProperties {  // neighborhood_search.rs:45-64
    Real radius, cell_size_inv;
    V2 grid_min;
    static inline uint16_t as_u16(Real f) {  // Rust `as u16`: saturating, NaN -> 0
        if (!(f == f)) return 0;
        if (f <= 0.0f) return 0;
        if (f >= 65535.0f) return 65535;
        return (uint16_t)f;
    }
    inline void cellpos(V2 p, uint16_t& cx, uint16_t& cy) const {
        V2 c = (p - grid_min) * cell_size_inv;
        cx = as_u16(c.x);
        cy = as_u16(c.y);
    }
    inline uint32_t cidx(V2 p) const {
        uint16_t cx, cy;
        cellpos(p, cx, cy);
        return morton_encode(cx, cy);
    }
};
struct Runs {  // neighborhood_search.rs:40-43
    uint32_t r[5][2];
};
struct CellGrid {  // neighborhood_search.rs:66-260
    std::vector<MortonCell> cells;
    CellGrid() { cells.push_back(MortonCell{0, UINT32_MAX}); }  // :80-87

    // neighborhood_search.rs:90-166.  `sorting` receives the applied permutation (out[k] = in[sorting[k]]).
    void update(const GridProperties& grid, std::vector<V2>& positions, std::vector<std::vector<V2>*>& vattrs,
                std::vector<std::vector<Real>*>& rattrs, std::vector<uint32_t>* sorting_out) {
        const size_t n = positions.size();
        std::vector<uint32_t> idx(n), key(n);
        for (size_t i = 0; i < n; ++i) {  // :111-114 (sequential in the reference)
            idx[i] = (uint32_t)i;
            key[i] = grid.cidx(positions[i]);
        }
        // :116-118 -- deviation D1: stable
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        {  // :122-140 apply_sorting (:71-78)
            std::vector<V2> tmp(n);
            for (size_t k = 0; k < n; ++k) tmp[k] = positions[idx[k]];
            positions.swap(tmp);
            for (auto* a : vattrs) {
                for (size_t k = 0; k < n; ++k) tmp[k] = (*a)[idx[k]];
                a->swap(tmp);
            }
            std::vector<Real> tr(n);
            for (auto* a : rattrs) {
                for (size_t k = 0; k < n; ++k) tr[k] = (*a)[idx[k]];
                a->swap(tr);
            }
        }
        // :146-165 create cells
        cells.clear();
        uint16_t px = UINT16_MAX, py = UINT16_MAX;
        for (size_t p = 0; p < n; ++p) {
            uint16_t cx, cy;
            grid.cellpos(positions[p], cx, cy);
            if (cx != px || cy != py) {
                cells.push_back(MortonCell{(uint32_t)p, morton_encode(cx, cy)});
                px = cx;
                py = cy;
            }
        }
        cells.push_back(MortonCell{(uint32_t)n, UINT32_MAX});
        if (sorting_out) sorting_out->swap(idx);
    }

    // neighborhood_search.rs:169-189
    static size_t find_next_cell(const MortonCell* cells, size_t len, uint32_t cidx) {
        const size_t LINEAR = 16;
        size_t mn = 0, mx = len, range = mx - mn;
        while (range > LINEAR) {
            range /= 2;
            size_t mid = mn + range;
            uint32_t c = cells[mid].cidx;
            if (c > cidx)
                mx = mid;
            else if (c < cidx)
                mn = mid;
            else
                return mid;
        }
        for (size_t p = mn; p < mx; ++p)
            if (cells[p].cidx >= cidx) return p;
        return mx;
    }

    // neighborhood_search.rs:191-259
    Runs get_particle_runs_in_neighborbox(uint32_t cidx) const {
        uint16_t px = (uint16_t)morton_decode_x(cidx), py = (uint16_t)morton_decode_y(cidx);
        // u16 arithmetic: release-mode wrap-around, as `pos.x - 1` / `pos.x + 1` (:193-194)
        uint32_t cidx_min = morton_encode((uint16_t)(px - 1), (uint16_t)(py - 1));
        uint32_t cidx_max = morton_encode((uint16_t)(px + 1), (uint16_t)(py + 1));
        uint32_t minx = cidx_min & MORTON_XBITS, miny = cidx_min & MORTON_YBITS;
        uint32_t maxx = cidx_max & MORTON_XBITS, maxy = cidx_max & MORTON_YBITS;
        const uint32_t MAX_MISSES = 8;
        size_t ai = find_next_cell(cells.data(), cells.size(), cidx_min);
        MortonCell cell = cells[ai];
        Runs runs;
        memset(&runs, 0, sizeof(runs));
        int run_idx = 0;
        while (cell.cidx <= cidx_max) {
            uint32_t misses = 0;
            while (!is_in_rect_presplit(cell.cidx, minx, miny, maxx, maxy)) {
                misses += 1;
                if (misses > MAX_MISSES) {
                    uint32_t expect = find_bigmin(cell.cidx, cidx_min, cidx_max);
                    ai += find_next_cell(cells.data() + ai, cells.size() - ai, expect);
                } else {
                    ai += 1;
                }
                cell = cells[ai];
                if (cell.cidx > cidx_max) return runs;
            }
            runs.r[run_idx][0] = cell.first_particle;
            while (true) {
                ai += 1;
                cell = cells[ai];
                if (!is_in_rect_presplit(cell.cidx, minx, miny, maxx, maxy)) break;
            }
            runs.r[run_idx][1] = cell.first_particle;
            run_idx += 1;
            if (run_idx == 5) break;
            ai += 1;
            if (ai >= cells.size()) break;
            cell = cells[ai];
        }
        return runs;
    }
};

struct NeighborRange {  // neighborhood_search.rs:268-273
    uint32_t start_index;
    uint16_t count_dynamic, count_total;
};
static const uint16_t MAX_NUM_NEIGHBORS = 64;  // neighborhood_search.rs:322
static const Real MIN_DISTANCE = 1.0e-10f;     // neighborhood_search.rs:323

struct NeighborLists {  // neighborhood_search.rs:297-450
    std::vector<NeighborRange> ranges;
    std::vector<uint32_t> lists;  // fixed stride of 64 per particle (placement in the reference is
                                  // an atomic bump, i.e. arbitrary; content per particle is what matters)
    uint64_t static_overflow_drops = 0, capped = 0;

    // neighborhood_search.rs:312-397
    void update(const GridProperties& grid, const CellGrid& gd, const CellGrid& gs, const std::vector<V2>& pd,
                const std::vector<V2>& ps) {
        const size_t n = pd.size();
        ranges.assign(n, NeighborRange{0, 0, 0});
        lists.resize(n * MAX_NUM_NEIGHBORS);
        const Real radius_sq = grid.radius * grid.radius;
        const long ncellpairs = (long)gd.cells.size() - 1;
        uint64_t drops = 0, cap = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : drops, cap)
        for (long c = 0; c < ncellpairs; ++c) {  // cells.par_windows(2) :337
            MortonCell cur = gd.cells[c], next = gd.cells[c + 1];
            Runs rd = gd.get_particle_runs_in_neighborbox(cur.cidx);
            Runs rs = gs.get_particle_runs_in_neighborbox(cur.cidx);
            for (uint32_t i = cur.first_particle; i < next.first_particle; ++i) {
                V2 q = pd[i];
                uint32_t* set = &lists[(size_t)i * MAX_NUM_NEIGHBORS];
                uint16_t cd = 0;
                bool full = false;
                for (int r = 0; r < 5 && !full; ++r) {
                    for (uint32_t j = rd.r[r][0]; j < rd.r[r][1]; ++j) {
                        Real d = distance2(q, pd[j]);
                        if (d <= radius_sq && d > MIN_DISTANCE) {
                            set[cd++] = j;
                            if (cd == MAX_NUM_NEIGHBORS) {
                                full = true;
                                cap++;
                                break;
                            }
                        }
                    }
                }
                uint16_t ct = cd;
                full = false;
                for (int r = 0; r < 5 && !full; ++r) {
                    for (uint32_t j = rs.r[r][0]; j < rs.r[r][1]; ++j) {
                        Real d = distance2(q, ps[j]);
                        if (d <= radius_sq && d > MIN_DISTANCE) {
                            if (ct >= MAX_NUM_NEIGHBORS) {  // deviation D4 (reference panics here)
                                drops++;
                                full = true;
                                break;
                            }
                            set[ct++] = j;
                            if (ct == MAX_NUM_NEIGHBORS) {
                                full = true;
                                cap++;
                                break;
                            }
                        }
                    }
                }
                ranges[i] = NeighborRange{(uint32_t)(i * MAX_NUM_NEIGHBORS), cd, ct};
            }
        }
        static_overflow_drops = drops;
        capped = cap;
    }
    inline const uint32_t* dyn(uint32_t i, uint32_t& n) const {  // :433-438
        n = ranges[i].count_dynamic;
        return &lists[ranges[i].start_index];
    }
    inline const uint32_t* stat(uint32_t i, uint32_t& n) const {  // :440-445
        n = (uint32_t)ranges[i].count_total - ranges[i].count_dynamic;
        return &lists[ranges[i].start_index + ranges[i].count_dynamic];
    }
};

struct NeighborhoodSearch {  // neighborhood_search.rs:452-522
    GridProperties grid;
    CellGrid dyn, stat;
    NeighborLists lists;
    explicit NeighborhoodSearch(Real radius) {  // :464-486
        grid.radius = radius;
        grid.cell_size_inv = 1.0f / radius;
        grid.grid_min = v2(-100.0f, -100.0f);
    }
    void update_static(std::vector<V2>& positions) {  // :488-491
        std::vector<std::vector<V2>*> va;
        std::vector<std::vector<Real>*> ra;
        stat.update(grid, positions, va, ra, nullptr);
    }
    void update_dynamic(std::vector<V2>& pd, std::vector<std::vector<V2>*>& va, std::vector<std::vector<Real>*>& ra,
                        const std::vector<V2>& ps, std::vector<uint32_t>* sorting_out) {  // :493-516
        dyn.update(grid, pd, va, ra, sorting_out);
        lists.update(grid, dyn, stat, pd, ps);
    }
};

// ---------------------------------------------------------------------------------------------
// rand 0.8 SmallRng (64-bit targets) = Xoshiro256PlusPlus seeded with SplitMix64 (deviation D3)
// ---------------------------------------------------------------------------------------------
struct SmallRng {
    uint64_t s[4];
    explicit SmallRng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) {
            seed += 0x9e3779b97f4a7c15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            s[i] = z ^ (z >> 31);
        }
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // rand Standard for f32: 24 random bits * 2^-24, in [0, 1)
    Real gen_f32() { return (Real)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};

// ---------------------------------------------------------------------------------------------
// fluidparticleworld.rs
// ---------------------------------------------------------------------------------------------
struct World {
    // ConstantFluidProperties :46-90
    Real smoothing_length, particle_density, fluid_density;
    std::vector<V2> positions, velocities, boundary;
    std::vector<Real> densities;
    NeighborhoodSearch ns;
    V2 gravity;
    bool boundary_changed;
    std::vector<uint32_t> last_sorting;  // permutation applied by the last update_dynamic (test tap)

    static Real radius_from_density(Real pd) { return 0.5f / sqrtf(pd); }  // :82-85
    World(Real smoothing_factor, Real pd, Real fd)
        : smoothing_length(2.0f * radius_from_density(pd) * smoothing_factor),  // :58
          particle_density(pd),
          fluid_density(fd),
          ns(smoothing_length),
          gravity(v2(0.0f, -9.81f)),
          boundary_changed(true) {}
    Real particle_mass() const { return fluid_density / particle_density; }  // :74-76
    Real num_particles_per_meter() const { return sqrtf(particle_density); }  // :78-80
    Real particle_radius() const { return radius_from_density(particle_density); }

    // :140-166
    void add_fluid_rect(Real rx, Real ry, Real rw, Real rh, Real jitter_amount) {
        Real nppm = num_particles_per_meter() * 0.9f;
        size_t nx = std::max<size_t>(1, (size_t)(rw * nppm));
        size_t ny = std::max<size_t>(1, (size_t)(rh * nppm));
        size_t total = positions.size() + nx * ny;
        velocities.resize(total, v2(0, 0));
        densities.resize(total, 0.0f);
        SmallRng rng((uint64_t)positions.size());
        V2 bl = v2(rx, ry);
        Real step = rmin(rw / (Real)nx, rh / (Real)ny);
        Real jf = step * jitter_amount;
        for (size_t y = 0; y < ny; ++y)
            for (size_t x = 0; x < nx; ++x) {
                Real jx = rng.gen_f32(), jy = rng.gen_f32();
                V2 jitter = (v2(jx, jy) * 0.5f + v2(0.5f, 0.5f)) * jf;
                positions.push_back(bl + jitter + v2(step * (Real)x, step * (Real)y));
            }
    }
    // :181-195
    void add_boundary_line(V2 start, V2 end) {
        Real distance = sqrtf(distance2(start, end));
        Real nppm = num_particles_per_meter();
        size_t n = std::max<size_t>(1, (size_t)ceilf(distance * nppm));
        V2 step = (end - start) / distance / nppm;
        V2 pos = start;
        for (size_t i = 0; i < n; ++i) {
            boundary.push_back(pos);
            pos = pos + step;
        }
        boundary_changed = true;
    }
    // :168-179
    void add_boundary_thick_line(V2 start, V2 end, uint32_t thickness) {
        V2 d = end - start;
        V2 dir = d * (1.0f / sqrtf(magnitude2(d)));  // cgmath normalize = v * (1 / |v|)
        V2 perp = v2(-dir.y, dir.x);
        Real tw = (Real)thickness / num_particles_per_meter();
        V2 elongation = dir * tw;
        V2 offset = (-perp) * tw;
        V2 step = perp * tw / (Real)thickness;
        for (uint32_t i = 0; i < thickness; ++i) {
            add_boundary_line(start + offset, end + offset + elongation);
            offset = offset + step;
        }
    }
    // :235-261.  velocities are always appended to the vector attributes (:242-243).
    void update_neighborhood_datastructure(std::vector<std::vector<V2>*> va, std::vector<std::vector<Real>*> ra) {
        va.push_back(&velocities);
        if (boundary_changed) {
            ns.update_static(boundary);
            boundary_changed = false;
        }
        ns.update_dynamic(positions, va, ra, boundary, &last_sorting);
    }
    // :197-231
    template <class K>
    void update_densities(const K& kernel) {
        const Real mass = particle_mass();
        const long n = (long)positions.size();
        densities.resize(n);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            V2 ri = positions[i];
            Real density = kernel.evaluate(0.0f, 0.0f) * mass;
            uint32_t cnt;
            const uint32_t* l = ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                Real r_sq = distance2(ri, positions[l[k]]);
                density += kernel.evaluate(r_sq, sqrtf(r_sq)) * mass;
            }
            l = ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                Real r_sq = distance2(ri, boundary[l[k]]);
                density += kernel.evaluate(r_sq, sqrtf(r_sq)) * mass;
            }
            densities[i] = rmax(density, fluid_density);
        }
    }
    void update_densities_id(int kernel) {
        switch (kernel) {
            case K_WENDLAND: update_densities(Wendland(smoothing_length)); break;
            case K_POLY6: update_densities(Poly6(smoothing_length)); break;
            case K_SPIKY: update_densities(Spiky(smoothing_length)); break;
            case K_CUBIC: update_densities(CubicSpline(smoothing_length)); break;
            default: break;
        }
    }
};

// Residual sum, deviation D2: f64 accumulation, rounded once to f32.
static Real sum_f64(const std::vector<Real>& a) {
    const long n = (long)a.size();
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long i = 0; i < n; ++i) s += (double)a[i];
    return (Real)s;
}

// ---------------------------------------------------------------------------------------------
// solver/dfsph.rs
// ---------------------------------------------------------------------------------------------
struct StepReport {
    uint64_t dt_ns;
    Real dt, max_velocity;
    uint32_t iters_density, iters_divergence;
    Real avg_density_error, avg_divergence;
    uint32_t warm_density, warm_divergence;
    uint32_t not_converged;  // bit 0: the density solver, bit 1: the divergence solver left its loop through the iteration cap (dfsph.rs:236-245, 391-400)
    uint32_t pad;
};

struct DFSPH {
    ViscosityModel visc;
    Wendland kernel;
    Real max_avg_density_error = 0.01f / 100.0f;  // dfsph.rs:49
    size_t max_iters_density = 200;               // :50
    size_t iters_density = 1;                     // :51
    Real max_divergence_error = 0.1f / 100.0f;    // :53
    size_t max_iters_divergence = 400;            // :54
    size_t iters_divergence = 0;                  // :55
    std::vector<Real> alpha, kappa, stiffness;    // :36-40
    std::vector<V2> predicted, accel;
    std::vector<Real> scratch;
    StepReport rep;

    DFSPH(int vk, Real vp, Real h) : visc(vk, vp, h), kernel(h) { memset(&rep, 0, sizeof(rep)); }

    void clear_cached_data() {  // :406-412
        alpha.clear();
        stiffness.clear();
        kappa.clear();
        iters_divergence = 0;
        iters_density = 0;
    }

    void compute_alpha_factors(const World& w) {  // :68-97
        const Real EPS = 1e-6f;
        const Real m = w.particle_mass();
        const long n = (long)w.positions.size();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            V2 ri = w.positions[i];
            Real gsq = 0.0f;
            V2 gsum = v2(0, 0);
            uint32_t cnt;
            const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                V2 g = gradient_from_positions(kernel, ri, w.positions[l[k]]) * m;
                gsum = gsum + g;
                gsq += magnitude2(g);
            }
            l = w.ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                V2 g = gradient_from_positions(kernel, ri, w.boundary[l[k]]) * m;
                gsum = gsum + g;
                gsq += magnitude2(g);
            }
            alpha[i] = 1.0f / rmax(magnitude2(gsum) + gsq, EPS);
        }
    }
    void compute_density_error(Real dt, const World& w, const std::vector<V2>& vel, std::vector<Real>& err) {  // :99-126
        const Real m = w.particle_mass(), rho0 = w.fluid_density;
        const long n = (long)w.positions.size();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            V2 pi = w.positions[i], vi = vel[i];
            Real delta = 0.0f;
            uint32_t cnt;
            const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                uint32_t j = l[k];
                delta += dot(vi - vel[j], gradient_from_positions(kernel, pi, w.positions[j]));
            }
            l = w.ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) delta += dot(vi, gradient_from_positions(kernel, pi, w.boundary[l[k]]));
            Real e = w.densities[i] + delta * m * dt;
            err[i] = rmax(rho0, e) - rho0;
        }
    }
    // :128-161 (accumulate=true) and :163-193 (warm start, accumulate=false, k = kappa)
    void correct_velocity_density(Real dt, const World& w, std::vector<V2>& vel, const std::vector<Real>* err) {
        const Real m = w.particle_mass();
        const Real inv_dt = 1.0f / dt;
        const long n = (long)w.positions.size();
        std::vector<V2> out(vel);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            V2 ri = w.positions[i];
            V2 delta = v2(0, 0);
            Real ki = err ? (*err)[i] * alpha[i] : kappa[i];
            uint32_t cnt;
            const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                uint32_t j = l[k];
                Real kj = err ? (*err)[j] * alpha[j] : kappa[j];
                delta = delta + (ki + kj) * gradient_from_positions(kernel, ri, w.positions[j]);
            }
            l = w.ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) delta = delta + ki * gradient_from_positions(kernel, ri, w.boundary[l[k]]);
            out[i] = vel[i] - inv_dt * delta * m;
        }
        if (err)
            for (long i = 0; i < n; ++i) kappa[i] += (*err)[i] * alpha[i];  // :142 (same product as ki)
        vel.swap(out);
    }
    void correct_density_error(Real dt, World& w, std::vector<V2>& vel) {  // :195-247
        rep.warm_density = 0;
        rep.not_converged = 0;
        if (iters_density > 1) {
            for (auto& k : kappa) k = 0.5f * rmax(k, -0.5f * w.fluid_density * w.fluid_density);  // :201-203
            correct_velocity_density(dt, w, vel, nullptr);
            rep.warm_density = 1;
        }
        for (auto& k : kappa) k = 0.0f;
        scratch.resize(w.positions.size());
        iters_density = 0;
        while (true) {
            compute_density_error(dt, w, vel, scratch);
            correct_velocity_density(dt, w, vel, &scratch);
            iters_density += 1;
            Real avg = sum_f64(scratch) / (Real)scratch.size();
            Real rel = avg / w.fluid_density;
            rep.avg_density_error = avg;
            if (rel * dt < max_avg_density_error) break;
            if (iters_density > max_iters_density) {  // :236-245 (the reference prints a warning and carries on)
                rep.not_converged |= 1u;
                break;
            }
        }
    }
    void compute_density_change(const World& w, const std::vector<V2>& vel, std::vector<Real>& chg) {  // :249-280
        const Real m = w.particle_mass();
        const long n = (long)w.positions.size();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            if (w.ns.lists.ranges[i].count_total < 9) {  // :261
                chg[i] = 0.0f;
                continue;
            }
            V2 ri = w.positions[i], vi = vel[i];
            Real delta = 0.0f;
            uint32_t cnt;
            const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                uint32_t j = l[k];
                delta += dot(vi - vel[j], gradient_from_positions(kernel, ri, w.positions[j]));
            }
            l = w.ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) delta += dot(vi, gradient_from_positions(kernel, ri, w.boundary[l[k]]));
            chg[i] = rmax(delta * m, 0.0f);
        }
    }
    // :282-314 (chg != null) and :316-344 (warm start, k = stiffness)
    void correct_velocity_divergence(const World& w, std::vector<V2>& vel, const std::vector<Real>* chg) {
        const Real m = w.particle_mass();
        const long n = (long)w.positions.size();
        std::vector<V2> out(vel);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            V2 ri = w.positions[i];
            V2 delta = v2(0, 0);
            Real ki = chg ? (*chg)[i] * alpha[i] : stiffness[i];
            uint32_t cnt;
            const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                uint32_t j = l[k];
                Real kj = chg ? (*chg)[j] * alpha[j] : stiffness[j];
                delta = delta + (ki + kj) * gradient_from_positions(kernel, ri, w.positions[j]);
            }
            l = w.ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) delta = delta + ki * gradient_from_positions(kernel, ri, w.boundary[l[k]]);
            out[i] = vel[i] - delta * m;
        }
        if (chg)
            for (long i = 0; i < n; ++i) stiffness[i] += (*chg)[i] * alpha[i];  // :296
        vel.swap(out);
    }
    void correct_divergence_error(Real dt, World& w, std::vector<V2>& vel) {  // :346-402
        rep.warm_divergence = 0;
        if (iters_divergence > 1) {
            for (auto& s : stiffness) s = 0.5f * rmax(s, -0.5f * w.fluid_density * w.fluid_density);
            correct_velocity_divergence(w, vel, nullptr);
            rep.warm_divergence = 1;
        }
        for (auto& s : stiffness) s = 0.0f;
        scratch.resize(w.positions.size());
        iters_divergence = 0;
        while (true) {
            compute_density_change(w, vel, scratch);
            correct_velocity_divergence(w, vel, &scratch);
            iters_divergence += 1;
            Real avg = sum_f64(scratch) / (Real)scratch.size() / w.fluid_density;
            rep.avg_divergence = avg;
            if (avg * dt < max_divergence_error) break;
            if (iters_divergence > max_iters_divergence) {  // :391-400
                rep.not_converged |= 2u;
                break;
            }
        }
    }
    void simulation_step(World& w, TimeManager& tm) {  // :414-525
        const size_t n = w.positions.size();
        if (alpha.size() != n) {  // :419-428
            alpha.resize(n, 0.0f);
            stiffness.resize(n, 0.0f);
            kappa.resize(n, 0.0f);
            w.update_neighborhood_datastructure({}, {&alpha});
            w.update_densities(kernel);
            compute_alpha_factors(w);
        }
        predicted.resize(n);
        accel.resize(n);
        Real dt = duration_as_secs_f32(tm.simulation_step());  // :433
        {                                                     // :436-469 non-pressure forces
            const Real m = w.particle_mass();
            V2 force = w.gravity * m;
            V2 npa = force / m;
            const long nn = (long)n;
#pragma omp parallel for schedule(static)
            for (long i = 0; i < nn; ++i) {
                V2 ri = w.positions[i], vi = w.velocities[i];
                V2 a = npa;
                uint32_t cnt;
                const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
                for (uint32_t k = 0; k < cnt; ++k) {
                    uint32_t j = l[k];
                    Real r_sq = distance2(ri, w.positions[j]);
                    a = a + visc.accel(dt, r_sq, sqrtf(r_sq), m, w.densities[j], w.velocities[j] - vi);
                }
                accel[i] = a;
            }
        }
        {  // :472-481 update timestep (sequential in the reference)
            Real mx = 0.0f;
            for (size_t i = 0; i < n; ++i) mx = rmax(mx, magnitude2(w.velocities[i] + accel[i] * dt));
            rep.max_velocity = sqrtf(mx);
            rep.dt_ns = tm.update_simulation_step(w.particle_radius() * 2.0f, sqrtf(mx));
            dt = duration_as_secs_f32(rep.dt_ns);
            rep.dt = dt;
        }
        for (size_t i = 0; i < n; ++i) predicted[i] = w.velocities[i] + accel[i] * dt;  // :486-491
        correct_density_error(dt, w, predicted);                                         // :496
        {                                                                                // :502-509 advect
            const long nn = (long)n;
#pragma omp parallel for schedule(static)
            for (long i = 0; i < nn; ++i) w.positions[i] = w.positions[i] + predicted[i] * dt;
        }
        w.update_neighborhood_datastructure({&predicted}, {});  // :512
        w.update_densities(kernel);                              // :516
        compute_alpha_factors(w);                                // :518
        correct_divergence_error(dt, w, predicted);              // :521
        w.velocities.swap(predicted);                            // :524
        rep.iters_density = (uint32_t)iters_density;
        rep.iters_divergence = (uint32_t)iters_divergence;
    }
};

// ---------------------------------------------------------------------------------------------
// solver/wscsph.rs
// ---------------------------------------------------------------------------------------------
struct WCSPH {
    ViscosityModel visc;
    Poly6 density_kernel;
    Spiky pressure_kernel;
    Real boundary_force_factor = 1.0f;  // wscsph.rs:34
    Real stiffness = 0.0f;
    std::vector<V2> accel;
    StepReport rep;
    WCSPH(int vk, Real vp, const World& w) : visc(vk, vp, w.smoothing_length), density_kernel(w.smoothing_length), pressure_kernel(w.smoothing_length) {
        memset(&rep, 0, sizeof(rep));
        set_compressibility(w, 0.01f, 1.0f);  // :39
    }
    void set_compressibility(const World& w, Real variation, Real max_flow_speed) {  // :45-49
        Real c = max_flow_speed / sqrtf(variation);
        stiffness = w.fluid_density * c * c / (Real)7;
    }
    static inline Real pressure(Real B, Real rho0, Real rho) {  // :52-57
        return B * (powi(rmax(rho / rho0, 1.0f), 7) - 1.0f);
    }
    void clear_cached_data() { accel.clear(); }  // :122-124
    void update_accellerations(const World& w, Real dt) {  // :59-118
        const Real mass = w.particle_mass(), rho0 = w.fluid_density;
        const V2 g = w.gravity;
        const long n = (long)w.positions.size();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) {
            V2 vi = w.velocities[i], ri = w.positions[i];
            Real rhoi = w.densities[i];
            V2 a = g;
            Real pi = pressure(stiffness, rho0, rhoi);
            uint32_t cnt;
            const uint32_t* l = w.ns.lists.dyn((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                uint32_t j = l[k];
                Real rhoj = w.densities[j];
                Real pj = pressure(stiffness, rho0, rhoj);
                V2 rij = w.positions[j] - ri;
                Real r_sq = magnitude2(rij);
                Real r = sqrtf(r_sq);
                Real pu = -mass * (pi + pj) / (2.0f * rhoi * rhoj);
                a = a + pu * pressure_kernel.gradient(rij, r_sq, r);
                a = a + visc.accel(dt, r_sq, r, mass, rhoj, w.velocities[j] - vi);
            }
            l = w.ns.lists.stat((uint32_t)i, cnt);
            for (uint32_t k = 0; k < cnt; ++k) {
                V2 rij = w.boundary[l[k]] - ri;
                Real r_sq = magnitude2(rij);
                a = a - boundary_force_factor * pressure_kernel.evaluate(r_sq, sqrtf(r_sq)) / r_sq * rij;
            }
            accel[i] = a;
        }
    }
    void simulation_step(World& w, TimeManager& tm) {  // :126-179
        const size_t n = w.positions.size();
        accel.resize(n, v2(0, 0));
        Real dt = duration_as_secs_f32(tm.simulation_step());
        for (size_t i = 0; i < n; ++i) {  // :141-150 (sequential in the reference)
            w.velocities[i] = w.velocities[i] + 0.5f * dt * accel[i];
            w.positions[i] = w.positions[i] + w.velocities[i] * dt;
        }
        w.update_neighborhood_datastructure({}, {});  // :153
        w.update_densities(density_kernel);           // :154
        update_accellerations(w, dt);                 // :155
        Real mx = 0.0f;                               // :160-166
        for (size_t i = 0; i < n; ++i) mx = rmax(mx, magnitude2(w.velocities[i] + accel[i] * dt));
        rep.max_velocity = sqrtf(mx);
        rep.dt_ns = tm.update_simulation_step(w.particle_radius() * 2.0f, sqrtf(mx));
        dt = duration_as_secs_f32(rep.dt_ns);
        rep.dt = dt;
        for (size_t i = 0; i < n; ++i) w.velocities[i] = w.velocities[i] + 0.5f * dt * accel[i];  // :175-177
    }
};

// ---------------------------------------------------------------------------------------------
// C interface (ctypes)
// ---------------------------------------------------------------------------------------------
extern "C" {
int yo_num_threads(int set) {
#ifdef _OPENMP
    if (set > 0) omp_set_num_threads(set);
    return omp_get_max_threads();
#else
    (void)set;
    return 1;
#endif
}
uint32_t yo_morton_encode(uint32_t x, uint32_t y) { return morton_encode((uint16_t)x, (uint16_t)y); }
uint32_t yo_morton_encode_lookup(uint32_t x, uint32_t y) { return morton_encode_lookup((uint16_t)x, (uint16_t)y); }
uint32_t yo_morton_decode_x(uint32_t m) { return morton_decode_x(m); }
uint32_t yo_morton_decode_y(uint32_t m) { return morton_decode_y(m); }
uint32_t yo_find_bigmin(uint32_t cur, uint32_t mn, uint32_t mx) { return find_bigmin(cur, mn, mx); }
int yo_is_in_rect(uint32_t m, uint32_t mn, uint32_t mx) {
    return is_in_rect_presplit(m, mn & MORTON_XBITS, mn & MORTON_YBITS, mx & MORTON_XBITS, mx & MORTON_YBITS);
}
uint32_t yo_position_to_cidx(float radius, float x, float y) {
    NeighborhoodSearch ns(radius);
    return ns.grid.cidx(v2(x, y));
}
float yo_kernel_evaluate(int kernel, float h, float r_sq, float r) {
    switch (kernel) {
        case K_WENDLAND: return Wendland(h).evaluate(r_sq, r);
        case K_POLY6: return Poly6(h).evaluate(r_sq, r);
        case K_SPIKY: return Spiky(h).evaluate(r_sq, r);
        case K_CUBIC: return CubicSpline(h).evaluate(r_sq, r);
        case K_VISCOSITY: return ViscosityKernel(h).evaluate(r_sq, r);
    }
    return 0.0f;
}
void yo_kernel_gradient(int kernel, float h, float dx, float dy, float* out) {
    V2 rij = v2(dx, dy);
    Real r_sq = magnitude2(rij), r = sqrtf(r_sq);
    V2 g = v2(0, 0);
    switch (kernel) {
        case K_WENDLAND: g = Wendland(h).gradient(rij, r_sq, r); break;
        case K_POLY6: g = Poly6(h).gradient(rij, r_sq, r); break;
        case K_SPIKY: g = Spiky(h).gradient(rij, r_sq, r); break;
        case K_CUBIC: g = CubicSpline(h).gradient(rij, r_sq, r); break;
    }
    out[0] = g.x;
    out[1] = g.y;
}
float yo_kernel_laplacian(float h, float r) { return ViscosityKernel(h).laplacian(r * r, r); }
uint64_t yo_duration_from_secs_f32(float s) { return duration_from_secs_f32(s); }
float yo_duration_as_secs_f32(uint64_t ns) { return duration_as_secs_f32(ns); }
void yo_rng_fill(uint64_t seed, float* out, uint32_t n) {
    SmallRng r(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.gen_f32();
}

// ---- standalone neighbourhood search (mirrors NeighborhoodSearch::new / update_static / update_dynamic) ----
void* yo_ns_new(float radius) { return new NeighborhoodSearch(radius); }
void yo_ns_free(void* h) { delete (NeighborhoodSearch*)h; }

// ---- world ----
void* yo_world_new(float smoothing_factor, float particle_density, float fluid_density) {
    return new World(smoothing_factor, particle_density, fluid_density);
}
// a world whose smoothing length (== search radius == cell size) is given directly
void* yo_world_new_h(float h, float particle_density, float fluid_density) {
    World* w = new World(1.0f, particle_density, fluid_density);
    w->smoothing_length = h;
    w->ns = NeighborhoodSearch(h);
    return w;
}
void yo_world_free(void* h) { delete (World*)h; }
void yo_world_add_fluid_rect(void* h, float x, float y, float w, float hh, float jitter) { ((World*)h)->add_fluid_rect(x, y, w, hh, jitter); }
void yo_world_add_boundary_line(void* h, float sx, float sy, float ex, float ey) { ((World*)h)->add_boundary_line(v2(sx, sy), v2(ex, ey)); }
void yo_world_add_boundary_thick_line(void* h, float sx, float sy, float ex, float ey, uint32_t t) {
    ((World*)h)->add_boundary_thick_line(v2(sx, sy), v2(ex, ey), t);
}
uint32_t yo_world_num_particles(void* h) { return (uint32_t)((World*)h)->positions.size(); }
uint32_t yo_world_num_boundary(void* h) { return (uint32_t)((World*)h)->boundary.size(); }
void yo_world_props(void* h, float* out) {  // h, mass, radius, rho0, gx, gy
    World* w = (World*)h;
    out[0] = w->smoothing_length;
    out[1] = w->particle_mass();
    out[2] = w->particle_radius();
    out[3] = w->fluid_density;
    out[4] = w->gravity.x;
    out[5] = w->gravity.y;
}
void yo_world_set_gravity(void* h, float gx, float gy) { ((World*)h)->gravity = v2(gx, gy); }
void yo_world_get(void* h, float* pos, float* vel, float* dens, float* boundary) {
    World* w = (World*)h;
    if (pos) memcpy(pos, w->positions.data(), w->positions.size() * sizeof(V2));
    if (vel) memcpy(vel, w->velocities.data(), w->velocities.size() * sizeof(V2));
    if (dens) memcpy(dens, w->densities.data(), w->densities.size() * sizeof(Real));
    if (boundary) memcpy(boundary, w->boundary.data(), w->boundary.size() * sizeof(V2));
}
void yo_world_set_particles(void* h, const float* pos, const float* vel, uint32_t n) {
    World* w = (World*)h;
    w->positions.resize(n);
    w->velocities.assign(n, v2(0, 0));
    w->densities.assign(n, 0.0f);
    memcpy(w->positions.data(), pos, n * sizeof(V2));
    if (vel) memcpy(w->velocities.data(), vel, n * sizeof(V2));
}
void yo_world_set_boundary(void* h, const float* b, uint32_t m) {
    World* w = (World*)h;
    w->boundary.resize(m);
    if (m) memcpy(w->boundary.data(), b, m * sizeof(V2));
    w->boundary_changed = true;
}
void yo_world_update_neighborhood(void* h) { ((World*)h)->update_neighborhood_datastructure({}, {}); }
void yo_world_update_densities(void* h, int kernel) { ((World*)h)->update_densities_id(kernel); }
void yo_world_last_sorting(void* h, uint32_t* out) {
    World* w = (World*)h;
    memcpy(out, w->last_sorting.data(), w->last_sorting.size() * sizeof(uint32_t));
}
uint32_t yo_world_num_cells(void* h, int is_static) {
    World* w = (World*)h;
    return (uint32_t)(is_static ? w->ns.stat.cells.size() : w->ns.dyn.cells.size());
}
void yo_world_cells(void* h, int is_static, uint32_t* first_particle, uint32_t* cidx) {
    World* w = (World*)h;
    const auto& c = is_static ? w->ns.stat.cells : w->ns.dyn.cells;
    for (size_t i = 0; i < c.size(); ++i) {
        first_particle[i] = c[i].first_particle;
        cidx[i] = c[i].cidx;
    }
}
void yo_world_runs(void* h, int is_static, uint32_t cidx, uint32_t* out10) {
    World* w = (World*)h;
    Runs r = (is_static ? w->ns.stat : w->ns.dyn).get_particle_runs_in_neighborbox(cidx);
    memcpy(out10, r.r, sizeof(r.r));
}
// neighbour lists in the reference's layout: per particle (count_dynamic, count_total) and a
// fixed-stride-64 list array (dynamic neighbours first, then static).
void yo_world_neighbors(void* h, uint16_t* count_dynamic, uint16_t* count_total, uint32_t* lists64) {
    World* w = (World*)h;
    const size_t n = w->positions.size();
    for (size_t i = 0; i < n; ++i) {
        count_dynamic[i] = w->ns.lists.ranges[i].count_dynamic;
        count_total[i] = w->ns.lists.ranges[i].count_total;
    }
    if (lists64) memcpy(lists64, w->ns.lists.lists.data(), n * MAX_NUM_NEIGHBORS * sizeof(uint32_t));
}
uint64_t yo_world_neighbor_stats(void* h, uint64_t* capped, uint64_t* static_drops) {
    World* w = (World*)h;
    uint64_t total = 0;
    for (auto& r : w->ns.lists.ranges) total += r.count_total;
    if (capped) *capped = w->ns.lists.capped;
    if (static_drops) *static_drops = w->ns.lists.static_overflow_drops;
    return total;
}

// ---- time manager ----
void* yo_time_new(int adaptive, uint64_t fixed_ns, uint64_t min_ns, uint64_t max_ns, float cfl_factor) {
    TimeManager* t = new TimeManager();
    t->adaptive = adaptive;
    t->fixed_ns = fixed_ns;
    t->min_ns = min_ns;
    t->max_ns = max_ns;
    t->cfl_factor = cfl_factor;
    t->reset();
    return t;
}
void yo_time_free(void* t) { delete (TimeManager*)t; }
uint64_t yo_time_simulation_step(void* t) { return ((TimeManager*)t)->simulation_step(); }
uint64_t yo_time_update(void* t, float diameter, float max_velocity) { return ((TimeManager*)t)->update_simulation_step(diameter, max_velocity); }
void yo_time_set_step(void* t, uint64_t ns) { ((TimeManager*)t)->simulation_step_ns = ns; }
void yo_time_set_target(void* t, uint64_t ns) { ((TimeManager*)t)->target_ns = ns; }
void yo_time_perform_step(void* t) { ((TimeManager*)t)->perform_step(); }
uint64_t yo_time_total(void* t) { return ((TimeManager*)t)->total_simulated_ns; }

// ---- solvers ----
void* yo_dfsph_new(void* world, int visc_kind, float visc_param) {
    World* w = (World*)world;
    return new DFSPH(visc_kind, visc_param, w->smoothing_length);
}
void yo_dfsph_free(void* s) { delete (DFSPH*)s; }
void yo_dfsph_clear(void* s) { ((DFSPH*)s)->clear_cached_data(); }
void yo_dfsph_step(void* s, void* w, void* t, StepReport* rep) {
    DFSPH* d = (DFSPH*)s;
    d->simulation_step(*(World*)w, *(TimeManager*)t);
    if (rep) *rep = d->rep;
}
// tolerances and iteration caps (dfsph.rs:49-50,53-54 are plain fields of DFSPHSolver); a value <= 0 / 0 keeps the current one
void yo_dfsph_set_params(void* s, float max_avg_density_error, uint32_t max_density_iters, float max_divergence_error, uint32_t max_divergence_iters) {
    DFSPH* d = (DFSPH*)s;
    if (max_avg_density_error > 0.0f) d->max_avg_density_error = max_avg_density_error;
    if (max_density_iters) d->max_iters_density = max_density_iters;
    if (max_divergence_error > 0.0f) d->max_divergence_error = max_divergence_error;
    if (max_divergence_iters) d->max_iters_divergence = max_divergence_iters;
}
void yo_dfsph_get(void* s, float* alpha, float* kappa, float* stiffness) {
    DFSPH* d = (DFSPH*)s;
    if (alpha) memcpy(alpha, d->alpha.data(), d->alpha.size() * sizeof(Real));
    if (kappa) memcpy(kappa, d->kappa.data(), d->kappa.size() * sizeof(Real));
    if (stiffness) memcpy(stiffness, d->stiffness.data(), d->stiffness.size() * sizeof(Real));
}
// individual passes for pass-level parity tests
void yo_dfsph_alpha(void* s, void* w, float* out) {
    DFSPH* d = (DFSPH*)s;
    World* ww = (World*)w;
    d->alpha.resize(ww->positions.size());
    d->compute_alpha_factors(*ww);
    memcpy(out, d->alpha.data(), d->alpha.size() * sizeof(Real));
}
void* yo_wcsph_new(void* world, int visc_kind, float visc_param) { return new WCSPH(visc_kind, visc_param, *(World*)world); }
void yo_wcsph_free(void* s) { delete (WCSPH*)s; }
void yo_wcsph_clear(void* s) { ((WCSPH*)s)->clear_cached_data(); }
void yo_wcsph_step(void* s, void* w, void* t, StepReport* rep) {
    WCSPH* d = (WCSPH*)s;
    d->simulation_step(*(World*)w, *(TimeManager*)t);
    if (rep) *rep = d->rep;
}
void yo_wcsph_get(void* s, float* accel, float* stiffness) {
    WCSPH* d = (WCSPH*)s;
    if (accel) memcpy(accel, d->accel.data(), d->accel.size() * sizeof(V2));
    if (stiffness) *stiffness = d->stiffness;
}
}  // extern "C"
