/* yasph_gpu.h -- C ABI of libyasph_gpu.so: the B200 (sm_100a) implementation of yasph2d's per-step SPH hot path.
 *
 * This is the drop-in boundary.  Everything above it (the Rust host of the reference, the C++/Python host
 * mirrors in this repo, the tests) talks to the GPU path only through these functions: plain pointers and
 * sizes, no C++ / torch types, never unwinds.  Paths below are relative to the reference repository.
 *
 * What each group replaces:
 *   yasph_create / yasph_destroy      FluidParticleWorld::new + ConstantFluidProperties (src/sph/fluidparticleworld.rs:46-127),
 *                                     DFSPHSolver::new (src/sph/solver/dfsph.rs:43-61), WCSPHSolver::new (src/sph/solver/wscsph.rs:29-41),
 *                                     NeighborhoodSearch::new (src/sph/neighborhood_search.rs:464-486), TimeManager::new (src/sph/timemanager.rs:105-129)
 *   yasph_set_boundary                NeighborhoodSearch::update_static (neighborhood_search.rs:488-491) via fluidparticleworld.rs:247-252
 *   yasph_upload/download_particles   the Particles SoA the Rust side owns (fluidparticleworld.rs:11-23)
 *   yasph_step                        Solver::simulation_step (src/sph/solver/mod.rs:17; dfsph.rs:414-525, wscsph.rs:126-179)
 *   yasph_clear_cached                Solver::clear_cached_data (solver/mod.rs:14; dfsph.rs:406-412, wscsph.rs:122-124)
 *   yasph_neighborhood_update         FluidParticleWorld::update_neighborhood_datastructure(vec![], vec![]) (fluidparticleworld.rs:235-261)
 *   yasph_neighbors_download          NeighborLists::{neighbors_dynamic, neighbors_static, num_neighbors} (neighborhood_search.rs:433-449)
 *   yasph_update_densities            FluidParticleWorld::update_densities (fluidparticleworld.rs:197-231)
 *   yasph_time_*                      TimeManager::simulation_step / update_simulation_step (timemanager.rs:136-138,252-279)
 *
 * Threading: a context is not thread-safe; one caller thread (the reference's world is !Send, scratch_buffer.rs:51-52).
 * Calls are synchronous: host arrays passed in may be reused on return, host arrays passed out are complete on return.
 * Errors: every function returns a yasph_status (0 = ok); yasph_last_error() gives the message.  Soft conditions the
 * reference reports with println!/assert! (neighbour cap hit, solver not converged, non-finite residual) are returned in
 * yasph_step_report so the caller can replicate them.
 */
#ifndef YASPH_GPU_H
#define YASPH_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YASPH_ABI_VERSION 4u
#define YASPH_MAX_NEIGHBORS 64u /* neighborhood_search.rs:322 */

typedef struct yasph_ctx yasph_ctx;

typedef enum yasph_status {
    YASPH_OK = 0,
    YASPH_ERR_INVALID_ARGUMENT = 1,
    YASPH_ERR_CUDA = 2,          /* a CUDA runtime call failed (message has the CUDA error string) */
    YASPH_ERR_CAPACITY = 3,      /* more particles / tiles / staged candidates than the context was created for */
    YASPH_ERR_STATE = 4,         /* call order violated (e.g. step before upload) */
    YASPH_ERR_NONFINITE = 5,     /* non-finite Jacobi residual (the reference asserts: dfsph.rs:223,378) */
    YASPH_ERR_NO_DEVICE = 6,     /* no CUDA device / wrong architecture: there is no CPU fallback */
    YASPH_ERR_COMM = 7           /* NCCL failure */
} yasph_status;

typedef enum yasph_solver_kind { YASPH_SOLVER_DFSPH = 0, YASPH_SOLVER_WCSPH = 1 } yasph_solver_kind;
typedef enum yasph_viscosity_kind { YASPH_VISCOSITY_XSPH = 0, YASPH_VISCOSITY_PHYSICAL = 1 } yasph_viscosity_kind;
/* smoothing kernels accepted by yasph_update_densities (src/sph/smoothing_kernel/mod.rs) */
typedef enum yasph_kernel_kind { YASPH_KERNEL_WENDLAND_C2 = 0, YASPH_KERNEL_POLY6 = 1, YASPH_KERNEL_SPIKY = 2, YASPH_KERNEL_CUBIC = 3 } yasph_kernel_kind;

/* per-particle arrays that can be read back for parity checks / checkpointing */
typedef enum yasph_field {
    YASPH_FIELD_POSITION = 0,     /* float2[N]  */
    YASPH_FIELD_VELOCITY = 1,     /* float2[N]  */
    YASPH_FIELD_DENSITY = 2,      /* float[N]   */
    YASPH_FIELD_ALPHA = 3,        /* float[N]   DFSPH alpha_values (dfsph.rs:36) */
    YASPH_FIELD_KAPPA = 4,        /* float[N]   DFSPH warmstart_kappa (dfsph.rs:40) */
    YASPH_FIELD_STIFFNESS = 5,    /* float[N]   DFSPH warmstart_stiffness (dfsph.rs:39) */
    YASPH_FIELD_ACCELERATION = 6, /* float2[N]  WCSPH accellerations (wscsph.rs:22) / DFSPH non-pressure accel of the last step */
    YASPH_FIELD_CELL_KEY = 7,     /* uint32[N]  Morton cell index of each sorted particle */
    YASPH_FIELD_SORT_PERMUTATION = 8, /* uint32[N] permutation applied by the last re-sort: new[k] = old[perm[k]] */
    YASPH_FIELD_BOUNDARY = 9,     /* float2[M]  boundary particles in their sorted order (fluidparticleworld.rs:247-252) */
    YASPH_FIELD_ID = 10,          /* uint32[N]  caller-visible particle id: index in the last yasph_upload_particles (+ id_base); needs YASPH_FLAG_TRACK_IDS */
    YASPH_FIELD_GHOST = 11        /* uint8[N_local] 1 = ghost copy of a particle another rank owns (slab mode; only with YASPH_FIELD_LOCAL_BIT) */
} yasph_field;
/* Slab mode: OR into the field id to read the rank's whole local array (owned + ghost particles, sorted order, the
 * indexing yasph_neighbors_download uses) instead of the owned particles only; *n_local of yasph_slab_get sizes it. */
#define YASPH_FIELD_LOCAL_BIT 0x100

typedef struct yasph_config {
    uint32_t abi_version;         /* = YASPH_ABI_VERSION */
    int32_t device;               /* CUDA device ordinal */
    uint32_t max_particles;       /* capacity for dynamic particles (slab mode: owned + ghosts + one step's migrants) */
    uint32_t max_boundary;        /* capacity for boundary particles */
    /* ConstantFluidProperties, fluidparticleworld.rs:46-90 */
    float smoothing_length;       /* h == neighbour search radius == cell size (neighborhood_search.rs:466) */
    float particle_density;       /* particles per m^2 at rest */
    float fluid_density;          /* rho0 */
    float gravity[2];             /* fluidparticleworld.rs:123: (0, -9.81) */
    float grid_min[2];            /* neighborhood_search.rs:478: (-100, -100) */
    int32_t solver;               /* yasph_solver_kind */
    int32_t viscosity;            /* yasph_viscosity_kind */
    float viscosity_param;        /* XSPH epsilon (xsph.rs:14: 0.05) or physical mu (physical.rs:15: 1.0016e-3) */
    /* DFSPH, dfsph.rs:49-55 */
    float dfsph_max_avg_density_error;   /* 0.01/100 */
    uint32_t dfsph_max_density_iters;    /* 200 */
    float dfsph_max_divergence_error;    /* 0.1/100 */
    uint32_t dfsph_max_divergence_iters; /* 400 */
    /* WCSPH, wscsph.rs:29-49 */
    float wcsph_stiffness;               /* B = rho0 * c^2 / 7, c = 1/sqrt(0.01) */
    float wcsph_boundary_force_factor;   /* 1.0 */
    /* TimeManager step policy, timemanager.rs:38-59 / main.rs:115-127 */
    int32_t adaptive_timestep;           /* 0 = FixedTimeStep, 1 = AdaptiveTimeStep */
    uint64_t timestep_fixed_ns;
    uint64_t timestep_min_ns;            /* from_secs_f32(1/60/400) = 41667 */
    uint64_t timestep_max_ns;            /* from_secs_f32(1/120/3)  = 2777778 */
    uint64_t timestep_target_frame_ns;   /* AdaptiveTimeStepTarget: 0 = None, else TargetFrameLength (timemanager.rs:23-36, 268-274) */
    float cfl_factor;                    /* 1.5 (DFSPH) / 0.2 (WCSPH), main.rs:115-118 */
    /* implementation knobs (0 = default) */
    uint32_t max_tiles;                  /* capacity for 8x8-cell tiles; default max_particles/32 + 4096 */
    uint32_t tile_dynamic_capacity;      /* most dynamic candidates (tile + apron) a tile may STAGE in shared memory; 0 = whatever the SM holds.  Larger
                                          * tiles (dense clusters) are processed unstaged from global memory: slow, same results */
    uint32_t tile_static_capacity;       /* the same for boundary candidates */
    uint32_t speculative_iterations;     /* Jacobi iterations launched between two convergence read-backs (default 2) */
    uint32_t flags;                      /* YASPH_FLAG_* */
    uint32_t max_halo;                   /* slab mode: capacity (particles per side) of the ghost / migrant buffers; default max(65536, max_particles / 8) */
    uint32_t ghost_columns;              /* slab mode: width W (cell columns) of the ghost layer a rank keeps of each neighbour's slab; 0 = 1.  With W > 1 a rank
                                          * recomputes its ghosts' per-pass values itself (same arithmetic, same order: same bits) and a pass needs a halo
                                          * exchange only once the values it gathers are stale in the innermost ghost column -- none at all in a step of
                                          * 2 + 2 Jacobi iterations with W = 8.  Must not exceed the width of the narrowest slab; the same on all ranks. */
} yasph_config;

#define YASPH_FLAG_PERMUTE_WARMSTART 1u /* permute kappa/stiffness with the particles; default off = reference behaviour (quirk Q1: dfsph.rs:512 passes only v*) */
#define YASPH_FLAG_PROFILE_PASSES 2u    /* record a CUDA event pair per pass (see yasph_pass_times) */
#define YASPH_FLAG_TRACK_IDS 4u         /* carry a uint32 id with every particle through re-sorts and migration (YASPH_FIELD_ID) */
#define YASPH_FLAG_NO_PEER_TRANSPORT 8u /* slab mode: keep every exchange on NCCL send/recv + all-reduce instead of stores into the peers' mapped mailboxes */

/* Fills every field with the reference's defaults for the given world parameters
 * (FluidParticleWorld::new(smoothing_factor, particle_density, fluid_density), main.rs:85-89). */
int32_t yasph_config_default(yasph_config* cfg, float smoothing_factor, float particle_density, float fluid_density, int32_t solver);

typedef struct yasph_step_report {
    uint64_t dt_prev_ns;          /* TimeManager::simulation_step() at entry (dfsph.rs:433) */
    uint64_t dt_ns;               /* result of update_simulation_step (dfsph.rs:478-480): the dt this step integrated with */
    float dt;                     /* dt_ns as_secs_f32 */
    float max_velocity;           /* sqrt(max |v + a dt|^2), dfsph.rs:474-479 */
    uint32_t iters_density;       /* num_density_correction_iterations (dfsph.rs:219) */
    uint32_t iters_divergence;    /* num_divergence_correction_iterations (dfsph.rs:374) */
    float avg_density_error;      /* last avg_density_error (dfsph.rs:221) */
    float avg_divergence;         /* last avg_divergence (dfsph.rs:376-377) */
    uint32_t warm_density;        /* 1 if the density warm start ran (dfsph.rs:199) */
    uint32_t warm_divergence;     /* 1 if the divergence warm start ran (dfsph.rs:354) */
    uint32_t neighbors_capped;    /* particles that hit the 64-neighbour cap ("particle has too many neighbors", neighborhood_search.rs:361,376) */
    uint32_t neighbors_dropped;   /* static candidates dropped because the dynamic ones filled all 64 slots (the reference panics, ns.rs:373) */
    uint32_t not_converged;       /* bit 0: density solver hit its cap (dfsph.rs:236), bit 1: divergence solver (dfsph.rs:391) */
    uint32_t num_cells;           /* non-empty cells of the dynamic grid */
    uint32_t num_tiles;           /* non-empty 8x8-cell tiles */
    uint32_t list_rebuilds;       /* list builds launched ahead of the tile-size read-back that had to be repeated, since creation */
    uint64_t total_neighbors;     /* sum over particles of count_total after the step's list build */
} yasph_step_report;

/* ---- lifecycle ---------------------------------------------------------------------------------------------- */
int32_t yasph_create(const yasph_config* cfg, yasph_ctx** out);
int32_t yasph_destroy(yasph_ctx* ctx);
const char* yasph_last_error(const yasph_ctx* ctx); /* ctx may be NULL: error of the last failed yasph_create */
int32_t yasph_get_config(const yasph_ctx* ctx, yasph_config* out);
/* replaces cfg.flags (YASPH_FLAG_*) at run time */
int32_t yasph_set_flags(yasph_ctx* ctx, uint32_t flags);
/* derived ConstantFluidProperties: out[0]=particle_mass, out[1]=particle_radius (fluidparticleworld.rs:74-89) */
int32_t yasph_get_properties(const yasph_ctx* ctx, float* out2);

/* ---- particle state ------------------------------------------------------------------------------------------ */
/* Boundary ("shadow") particles: copied to the device and sorted into Morton cell order once, like update_static.
 * xy = M interleaved (x, y) pairs.  yasph_download_field(YASPH_FIELD_BOUNDARY) returns the sorted order. */
int32_t yasph_set_boundary(yasph_ctx* ctx, const float* xy, uint32_t m);
/* Dynamic particles.  vel may be NULL (zeros).  Does not touch solver caches: follow the reference and call
 * yasph_clear_cached when the scene is reset (main.rs:292-298). */
int32_t yasph_upload_particles(yasph_ctx* ctx, const float* pos_xy, const float* vel_xy, uint32_t n);
/* Any pointer may be NULL.  Arrays are in the current sorted order, as the reference leaves its Vecs. */
int32_t yasph_download_particles(yasph_ctx* ctx, float* pos_xy, float* vel_xy, float* densities);
int32_t yasph_download_field(yasph_ctx* ctx, int32_t field, void* out, uint64_t out_bytes);

/* ---- checkpoint / resume (the reference has no counterpart: its solver state dies with the process) ----
 * What a solver carries from one simulation_step to the next besides positions and velocities: DFSPH the two warm-start
 * arrays and the iteration counts that switch the warm starts on (dfsph.rs:36-41, 199, 354), WCSPH the accelerations of the
 * previous step (wscsph.rs:22, 141-150); plus TimeManager's current step (timemanager.rs:136-138).  Densities, alpha factors
 * and the neighbour lists are functions of the (sorted) positions and are rebuilt.
 * Resume: yasph_set_boundary, yasph_upload_particles(sorted positions, velocities of the checkpoint), yasph_solver_state_set
 * (runs the solver's first-call initialisation, dfsph.rs:419-428, right away), then yasph_upload_field for KAPPA / STIFFNESS
 * (DFSPH) or ACCELERATION (WCSPH).  The following steps are bit-identical to the uninterrupted run. */
typedef struct yasph_solver_state {
    uint64_t step_ns;             /* TimeManager::simulation_step */
    uint32_t iters_density;       /* num_density_correction_iterations of the last solve */
    uint32_t iters_divergence;    /* num_divergence_correction_iterations of the last solve */
    uint32_t initialized;         /* get: the solver has run its first-call initialisation; set: run it now */
    uint32_t reserved;
    uint64_t total_simulated_ns;  /* TimeManager::total_simulated_time between steps (timemanager.rs:246): the lower bound of the
                                   * TargetFrameLength rule depends on it (timemanager.rs:268-272) */
} yasph_solver_state;
int32_t yasph_solver_state_get(yasph_ctx* ctx, yasph_solver_state* out);
int32_t yasph_solver_state_set(yasph_ctx* ctx, const yasph_solver_state* in);
/* overwrite a per-particle solver array (same order as yasph_download_field): YASPH_FIELD_KAPPA, _STIFFNESS, _ACCELERATION */
int32_t yasph_upload_field(yasph_ctx* ctx, int32_t field, const void* data, uint64_t bytes);
int32_t yasph_num_particles(const yasph_ctx* ctx, uint32_t* n, uint32_t* m);

/* ---- solver --------------------------------------------------------------------------------------------------- */
int32_t yasph_clear_cached(yasph_ctx* ctx);
/* One Solver::simulation_step on the device-resident state.  report may be NULL. */
int32_t yasph_step(yasph_ctx* ctx, yasph_step_report* report);
/* `steps` simulation steps in one call -- the application's frame loop (main.rs:339-360 runs several simulation steps per rendered
 * frame).  Same results as `steps` calls of yasph_step; reports (NULL or [steps]) receives every step's report.  On one GPU the head of
 * each following step is enqueued ahead of the read-back that ends the running one, so the GPU does not idle between steps. */
int32_t yasph_step_n(yasph_ctx* ctx, uint32_t steps, yasph_step_report* reports);
/* The reference-facing call with HOST buffers: upload pos/vel (N particles), one step, download pos/vel/densities
 * into the same arrays (new sorted order) -- what `solver.simulation_step(&mut world, &mut time)` does to the Vecs. */
int32_t yasph_step_host(yasph_ctx* ctx, float* pos_xy, float* vel_xy, float* densities, uint32_t n, yasph_step_report* report);
/* The same with options.  YASPH_HOST_INPUT_UNCHANGED: the caller has not written to pos_xy / vel_xy since the previous
 * yasph_step_host* call on this context handed them back (the reference's application only READS the particle arrays between steps,
 * main.rs:242-258), so the device still holds their content and the upload is skipped; the arrays are outputs only. */
#define YASPH_HOST_INPUT_UNCHANGED 1u
int32_t yasph_step_host_ex(yasph_ctx* ctx, float* pos_xy, float* vel_xy, float* densities, uint32_t n, uint32_t options, yasph_step_report* report);

/* ---- TimeManager mirror ------------------------------------------------------------------------------------- */
int32_t yasph_time_get_step_ns(const yasph_ctx* ctx, uint64_t* step_ns);
int32_t yasph_time_set_step_ns(yasph_ctx* ctx, uint64_t step_ns);
int32_t yasph_time_restart(yasph_ctx* ctx);
/* TimeManager::total_simulated_time as it stands when simulation_step is called, i.e. after the frame loop has added the
 * current step (timemanager.rs:246); only the TargetFrameLength rule reads it (timemanager.rs:268-274).  Without this call the
 * context keeps the sum itself (every yasph_step adds its entry step, as the frame loop does before each step). */
int32_t yasph_time_set_total_simulated_ns(yasph_ctx* ctx, uint64_t total_ns);
int32_t yasph_time_get_total_simulated_ns(const yasph_ctx* ctx, uint64_t* total_ns); /* TimeManager::restart (timemanager.rs:131-133) */

/* ---- neighbourhood-only surface ------------------------------------------------------------------------------ */
/* Re-sorts positions and velocities and rebuilds cells, tiles and the neighbour lists. */
int32_t yasph_neighborhood_update(yasph_ctx* ctx, yasph_step_report* report);
/* Neighbour lists in the reference's shape: per particle count_dynamic / count_total (u16, ns.rs:268-273) and the
 * u32 neighbour indices at a fixed stride of 64 per particle (dynamic first, then static; ascending index order).
 * lists64 may be NULL. */
int32_t yasph_neighbors_download(yasph_ctx* ctx, uint16_t* count_dynamic, uint16_t* count_total, uint32_t* lists64);
int32_t yasph_update_densities(yasph_ctx* ctx, int32_t kernel);
/* DFSPH compute_alpha_factors (dfsph.rs:68-97) on the current lists; result readable via YASPH_FIELD_ALPHA */
int32_t yasph_compute_alpha(yasph_ctx* ctx);

/* ---- multi-GPU: 1-D slab decomposition over cell columns (one process per GPU, NCCL over NVLink) ----------------- */
/* The reference is single-address-space; this group has no counterpart there.  Every rank owns the
 * particles whose cell column (x index of neighborhood_search.rs:45-64) lies in [col_lo, col_hi); the slabs of ranks
 * 0..world-1 are adjacent and ascending in x.  Each neighbourhood update migrates particles that left the slab to the
 * adjacent rank and receives the neighbours' boundary columns as ghost particles; every neighbour-dependent pass is
 * followed by a halo exchange of the field the next pass gathers, the Jacobi residual and the CFL maximum are
 * all-reduced.  Boundary particles are replicated on every rank.  Slab mode implies YASPH_FLAG_PERMUTE_WARMSTART (the
 * reference's index-stale warm-start arrays, quirk Q1, have no meaning across address spaces).
 * In slab mode yasph_upload_particles takes this rank's particles only, and yasph_download_particles /
 * yasph_download_field / yasph_num_particles report the owned particles only (ghosts are internal). */
#define YASPH_COMM_ID_BYTES 128u
/* ncclGetUniqueId: call on one rank, hand the bytes to every rank over any host channel */
int32_t yasph_comm_unique_id(void* out_id, uint64_t bytes);
/* ncclCommInitRank on the context's device; collective over all ranks */
int32_t yasph_comm_init(yasph_ctx* ctx, int32_t rank, int32_t world, const void* id, uint64_t bytes);
/* Loopback transport: all ranks are contexts of ONE process (one host thread per rank, on one or several devices); the
 * messages of the NCCL transport travel as device-to-device copies.  It exists so that the complete slab logic can be
 * tested on a single GPU.  Every rank's calls must come from its own thread (they rendezvous). */
int32_t yasph_loopback_create(int32_t world, void** fabric);
int32_t yasph_loopback_destroy(void* fabric);
int32_t yasph_comm_init_loopback(yasph_ctx* ctx, void* fabric, int32_t rank);
/* Owned cell-column range of this rank and the global number of dynamic particles (the N of dfsph.rs:221,376).
 * id_base is added to the upload index to form YASPH_FIELD_ID.  Call before yasph_upload_particles. */
int32_t yasph_slab_set(yasph_ctx* ctx, uint32_t col_lo, uint32_t col_hi, uint64_t n_global, uint32_t id_base);
typedef struct yasph_slab_info {
    int32_t rank, world;
    uint32_t col_lo, col_hi;
    uint32_t n_own;                 /* particles this rank owns after the last neighbourhood update */
    uint32_t n_local;               /* owned + ghost particles */
    uint32_t n_ghost_left, n_ghost_right;
    uint32_t migrated_out_left, migrated_out_right, migrated_in;  /* of the last neighbourhood update */
    uint32_t peer_transport;        /* 1: per-pass halo exchanges and all-reduces run as stores into the peers' mapped mailboxes (NVLink); 0: NCCL / loopback */
    uint64_t n_global;
    uint64_t halo_exchanges;        /* halo exchanges since creation */
    uint64_t allreduces;            /* all-reduces since creation */
} yasph_slab_info;
int32_t yasph_slab_get(yasph_ctx* ctx, yasph_slab_info* out);
/* cell column of an x coordinate under the context's grid (neighborhood_search.rs:52-58), for host-side partitioning */
int32_t yasph_cell_column(const yasph_config* cfg, float x, uint32_t* column);
/* yasph_step_host for a slab: n_in particles in (this rank's, in the order of the last download), one step, the owned
 * particles after migration out (*n_out <= capacity). */
int32_t yasph_step_host_slab(yasph_ctx* ctx, float* pos_xy, float* vel_xy, float* densities, uint32_t n_in, uint32_t capacity,
                             uint32_t* n_out, yasph_step_report* report);
/* The same with options (YASPH_HOST_INPUT_UNCHANGED as for yasph_step_host_ex: the arrays still hold what the previous call on this
 * context handed back, so nothing is uploaded -- the arrays are outputs only). */
int32_t yasph_step_host_slab_ex(yasph_ctx* ctx, float* pos_xy, float* vel_xy, float* densities, uint32_t n_in, uint32_t capacity,
                                uint32_t options, uint32_t* n_out, yasph_step_report* report);

/* ---- measurement ----------------------------------------------------------------------------------------------- */
#define YASPH_NUM_PASSES 17
/* Per-pass device time of the last step in microseconds (needs YASPH_FLAG_PROFILE_PASSES), index = yasph_pass. */
typedef enum yasph_pass {
    YASPH_PASS_VISCOSITY = 0, YASPH_PASS_PREDICT = 1, YASPH_PASS_DENSITY_WARM = 2, YASPH_PASS_DENSITY_SOLVE = 3,
    YASPH_PASS_ADVECT_KEYGEN = 4, YASPH_PASS_SORT = 5, YASPH_PASS_GATHER = 6, YASPH_PASS_CELLS_TILES = 7,
    YASPH_PASS_LISTS = 8, YASPH_PASS_DENSITY_ALPHA = 9, YASPH_PASS_DIVERGENCE_WARM = 10, YASPH_PASS_DIVERGENCE_SOLVE = 11,
    YASPH_PASS_WCSPH_ACCEL = 12, YASPH_PASS_WCSPH_KICK = 13, YASPH_PASS_HALO = 14 /* per-pass halo exchanges */,
    YASPH_PASS_MIGRATE = 15 /* slab mode: migrant + ghost particle exchange of a neighbourhood update */, YASPH_PASS_TOTAL = 16
} yasph_pass;
int32_t yasph_pass_times(yasph_ctx* ctx, float* out_us /* [YASPH_NUM_PASSES] */);
/* Device timeline of the last yasph_step_host call in microseconds from its first upload (needs YASPH_FLAG_PROFILE_PASSES):
 * [1] both uploads done, [2] positions on the host, [3] densities on the host, [4] last kernel done, [5] velocities on the host
 * (= end of the call's device work); [0] is 0. */
int32_t yasph_host_step_times(yasph_ctx* ctx, float* out_us /* [6] */);
/* number of kernel launches issued by this context since creation */
int32_t yasph_launch_count(const yasph_ctx* ctx, uint64_t* launches);
/* raw CUDA stream (cudaStream_t) the context launches on, for event timing by the caller */
int32_t yasph_stream(const yasph_ctx* ctx, void** stream);

/* ---- host-side scene builders (no device work) ---------------------------------------------------------------- */
/* FluidParticleWorld::add_fluid_rect (fluidparticleworld.rs:140-166): lattice at 0.9x rest spacing with jitter drawn from
 * rand 0.8's SmallRng (xoshiro256++ seeded with SplitMix64) seeded with `seed` = number of particles already in the
 * world (fluidparticleworld.rs:153).  out_xy == NULL only counts.  *count receives the number of particles. */
int32_t yasph_scene_fluid_rect(float particle_density, float x, float y, float w, float h, float jitter, uint64_t seed,
                               float* out_xy, uint32_t capacity, uint32_t* count);
/* FluidParticleWorld::add_boundary_line (fluidparticleworld.rs:181-195) */
int32_t yasph_scene_boundary_line(float particle_density, float sx, float sy, float ex, float ey, float* out_xy,
                                  uint32_t capacity, uint32_t* count);
/* FluidParticleWorld::add_boundary_thick_line (fluidparticleworld.rs:168-179) */
int32_t yasph_scene_boundary_thick_line(float particle_density, float sx, float sy, float ex, float ey, uint32_t thickness,
                                        float* out_xy, uint32_t capacity, uint32_t* count);
/* std::time::Duration::from_secs_f32 / as_secs_f32 as the TimeManager mirror uses them */
uint64_t yasph_duration_from_secs_f32(float secs);
float yasph_duration_as_secs_f32(uint64_t ns);

#ifdef __cplusplus
}
#endif
#endif /* YASPH_GPU_H */
