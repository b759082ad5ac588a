//! Raw bindings to `libyasph_gpu.so` -- hand-written from `include/yasph_gpu.h` (ABI version 4); `bindgen` on that header
//! produces the same declarations.  NOT COMPILED in the repository that ships it (no Rust toolchain in that image): the same
//! ABI is exercised there by `tests/cabi_driver.c` and `yasph2d_b200/_capi.py`; `tests/test_host_abi.py` pins the struct sizes
//! asserted at the bottom of this file.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const YASPH_ABI_VERSION: u32 = 4;
pub const YASPH_MAX_NEIGHBORS: usize = 64; // neighborhood_search.rs:322

// yasph_status
pub const YASPH_OK: i32 = 0;
pub const YASPH_ERR_INVALID_ARGUMENT: i32 = 1;
pub const YASPH_ERR_CUDA: i32 = 2;
pub const YASPH_ERR_CAPACITY: i32 = 3;
pub const YASPH_ERR_STATE: i32 = 4;
pub const YASPH_ERR_NONFINITE: i32 = 5; // the reference asserts: dfsph.rs:223,378
pub const YASPH_ERR_NO_DEVICE: i32 = 6; // there is no CPU fallback
pub const YASPH_ERR_COMM: i32 = 7;

pub const YASPH_SOLVER_DFSPH: i32 = 0;
pub const YASPH_SOLVER_WCSPH: i32 = 1;
pub const YASPH_VISCOSITY_XSPH: i32 = 0;
pub const YASPH_VISCOSITY_PHYSICAL: i32 = 1;
pub const YASPH_KERNEL_WENDLAND_C2: i32 = 0;
pub const YASPH_KERNEL_POLY6: i32 = 1;
pub const YASPH_KERNEL_SPIKY: i32 = 2;
pub const YASPH_KERNEL_CUBIC: i32 = 3;

// yasph_field
pub const YASPH_FIELD_POSITION: i32 = 0;
pub const YASPH_FIELD_VELOCITY: i32 = 1;
pub const YASPH_FIELD_DENSITY: i32 = 2;
pub const YASPH_FIELD_ALPHA: i32 = 3;
pub const YASPH_FIELD_KAPPA: i32 = 4;
pub const YASPH_FIELD_STIFFNESS: i32 = 5;
pub const YASPH_FIELD_ACCELERATION: i32 = 6;
pub const YASPH_FIELD_CELL_KEY: i32 = 7;
pub const YASPH_FIELD_SORT_PERMUTATION: i32 = 8;
pub const YASPH_FIELD_BOUNDARY: i32 = 9;
pub const YASPH_FIELD_ID: i32 = 10;

pub const YASPH_FLAG_PERMUTE_WARMSTART: u32 = 1;
pub const YASPH_FLAG_PROFILE_PASSES: u32 = 2;
pub const YASPH_FLAG_TRACK_IDS: u32 = 4;
pub const YASPH_FLAG_NO_PEER_TRANSPORT: u32 = 8;

#[repr(C)]
pub struct yasph_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct yasph_config {
    pub abi_version: u32,
    pub device: i32,
    pub max_particles: u32,
    pub max_boundary: u32,
    pub smoothing_length: f32, // h == search radius == cell size (neighborhood_search.rs:466)
    pub particle_density: f32,
    pub fluid_density: f32,
    pub gravity: [f32; 2],  // fluidparticleworld.rs:123
    pub grid_min: [f32; 2], // neighborhood_search.rs:478
    pub solver: i32,
    pub viscosity: i32,
    pub viscosity_param: f32, // XSPH epsilon (xsph.rs:14) or physical mu (physical.rs:15)
    pub dfsph_max_avg_density_error: f32, // dfsph.rs:49
    pub dfsph_max_density_iters: u32,     // dfsph.rs:50
    pub dfsph_max_divergence_error: f32,  // dfsph.rs:53
    pub dfsph_max_divergence_iters: u32,  // dfsph.rs:54
    pub wcsph_stiffness: f32,             // wscsph.rs:48
    pub wcsph_boundary_force_factor: f32, // wscsph.rs:34
    pub adaptive_timestep: i32,           // SimulationStepConfig (timemanager.rs:38-59)
    pub timestep_fixed_ns: u64,
    pub timestep_min_ns: u64,
    pub timestep_max_ns: u64,
    pub timestep_target_frame_ns: u64, // AdaptiveTimeStepTarget: 0 = None, else TargetFrameLength
    pub cfl_factor: f32,
    pub max_tiles: u32,
    pub tile_dynamic_capacity: u32,
    pub tile_static_capacity: u32,
    pub speculative_iterations: u32,
    pub flags: u32,
    pub max_halo: u32,
    pub ghost_columns: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct yasph_step_report {
    pub dt_prev_ns: u64, // TimeManager::simulation_step() at entry (dfsph.rs:433)
    pub dt_ns: u64,      // result of update_simulation_step (dfsph.rs:478-480)
    pub dt: f32,
    pub max_velocity: f32,
    pub iters_density: u32,
    pub iters_divergence: u32,
    pub avg_density_error: f32,
    pub avg_divergence: f32,
    pub warm_density: u32,
    pub warm_divergence: u32,
    pub neighbors_capped: u32,  // "particle has too many neighbors" (neighborhood_search.rs:361,376)
    pub neighbors_dropped: u32, // static hits beyond a full list (the reference panics there, :373)
    pub not_converged: u32,     // bit 0 density, bit 1 divergence solver hit its cap (dfsph.rs:236,391)
    pub num_cells: u32,
    pub num_tiles: u32,
    pub list_rebuilds: u32,
    pub total_neighbors: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct yasph_solver_state {
    pub step_ns: u64,
    pub iters_density: u32,
    pub iters_divergence: u32,
    pub initialized: u32,
    pub reserved: u32,
    pub total_simulated_ns: u64,
}

#[link(name = "yasph_gpu")]
extern "C" {
    pub fn yasph_config_default(cfg: *mut yasph_config, smoothing_factor: f32, particle_density: f32, fluid_density: f32, solver: i32) -> i32;
    pub fn yasph_create(cfg: *const yasph_config, out: *mut *mut yasph_ctx) -> i32;
    pub fn yasph_destroy(ctx: *mut yasph_ctx) -> i32;
    pub fn yasph_last_error(ctx: *const yasph_ctx) -> *const c_char;
    pub fn yasph_get_config(ctx: *const yasph_ctx, out: *mut yasph_config) -> i32;
    pub fn yasph_set_flags(ctx: *mut yasph_ctx, flags: u32) -> i32;
    pub fn yasph_get_properties(ctx: *const yasph_ctx, out2: *mut f32) -> i32;
    pub fn yasph_num_particles(ctx: *const yasph_ctx, n: *mut u32, m: *mut u32) -> i32;
    // Particles SoA (fluidparticleworld.rs:11-23); Vec<Point2<f32>> / Vec<Vector2<f32>> are interleaved (x, y) pairs
    pub fn yasph_set_boundary(ctx: *mut yasph_ctx, xy: *const f32, m: u32) -> i32;
    pub fn yasph_upload_particles(ctx: *mut yasph_ctx, pos: *const f32, vel: *const f32, n: u32) -> i32;
    pub fn yasph_download_particles(ctx: *mut yasph_ctx, pos: *mut f32, vel: *mut f32, dens: *mut f32) -> i32;
    pub fn yasph_download_field(ctx: *mut yasph_ctx, field: i32, out: *mut c_void, bytes: u64) -> i32;
    pub fn yasph_upload_field(ctx: *mut yasph_ctx, field: i32, data: *const c_void, bytes: u64) -> i32;
    pub fn yasph_solver_state_get(ctx: *mut yasph_ctx, out: *mut yasph_solver_state) -> i32;
    pub fn yasph_solver_state_set(ctx: *mut yasph_ctx, state: *const yasph_solver_state) -> i32;
    // trait Solver (solver/mod.rs:12-18)
    pub fn yasph_clear_cached(ctx: *mut yasph_ctx) -> i32;
    pub fn yasph_step(ctx: *mut yasph_ctx, report: *mut yasph_step_report) -> i32;
    pub fn yasph_step_host_slab_ex(ctx: *mut yasph_ctx, pos: *mut f32, vel: *mut f32, dens: *mut f32, n_in: u32, capacity: u32, options: u32, n_out: *mut u32, report: *mut yasph_step_report) -> i32;
    pub fn yasph_step_n(ctx: *mut yasph_ctx, steps: u32, reports: *mut yasph_step_report) -> i32;
    pub fn yasph_step_host(ctx: *mut yasph_ctx, pos: *mut f32, vel: *mut f32, dens: *mut f32, n: u32, report: *mut yasph_step_report) -> i32;
    // TimeManager mirror (timemanager.rs:131-138, 252-279)
    pub fn yasph_time_get_step_ns(ctx: *const yasph_ctx, ns: *mut u64) -> i32;
    pub fn yasph_time_set_step_ns(ctx: *mut yasph_ctx, ns: u64) -> i32;
    pub fn yasph_time_restart(ctx: *mut yasph_ctx) -> i32;
    pub fn yasph_time_set_total_simulated_ns(ctx: *mut yasph_ctx, ns: u64) -> i32;
    pub fn yasph_time_get_total_simulated_ns(ctx: *const yasph_ctx, ns: *mut u64) -> i32;
    // NeighborhoodSearch / NeighborLists / update_densities (neighborhood_search.rs:433-449, 461-522; fluidparticleworld.rs:197-231)
    pub fn yasph_neighborhood_update(ctx: *mut yasph_ctx, report: *mut yasph_step_report) -> i32;
    pub fn yasph_neighbors_download(ctx: *mut yasph_ctx, count_dynamic: *mut u16, count_total: *mut u16, lists64: *mut u32) -> i32;
    pub fn yasph_update_densities(ctx: *mut yasph_ctx, kernel: i32) -> i32;
    pub fn yasph_compute_alpha(ctx: *mut yasph_ctx) -> i32;
    // measurement taps
    pub fn yasph_pass_times(ctx: *mut yasph_ctx, out_us: *mut f32 /* [17] */) -> i32;
    pub fn yasph_host_step_times(ctx: *mut yasph_ctx, out_us: *mut f32 /* [6] */) -> i32;
    pub fn yasph_launch_count(ctx: *const yasph_ctx, launches: *mut u64) -> i32;
}

// CUDA runtime: page-locking the particle Vecs lets yasph_step_host hand results back while the step still computes
#[link(name = "cudart")]
extern "C" {
    pub fn cudaHostRegister(ptr: *mut c_void, size: usize, flags: u32) -> i32;
    pub fn cudaHostUnregister(ptr: *mut c_void) -> i32;
}

/// `yasph_last_error` as an owned string (`ctx` may be null: the error of a failed `yasph_create`).
pub fn last_error(ctx: *const yasph_ctx) -> String {
    unsafe {
        let p = yasph_last_error(ctx);
        if p.is_null() {
            String::new()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

// the sizes tests/test_host_abi.py pins for the ctypes mirror of the same structs
const _: () = assert!(std::mem::size_of::<yasph_config>() == 152);
const _: () = assert!(std::mem::size_of::<yasph_step_report>() == 80);
const _: () = assert!(std::mem::size_of::<yasph_solver_state>() == 32);
