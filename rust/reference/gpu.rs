// src/sph/solver/gpu.rs -- `GpuSolver: sph::Solver`, the drop-in for DFSPHSolver / WCSPHSolver that runs the step on a B200
// through libyasph_gpu.so (crate yasph2d-gpu-sys).  NOT COMPILED where it was written (no Rust toolchain there); it needs
// visibility.patch (two `pub(super)` accessors).  It lives inside the `sph` module tree because it calls those accessors and
// `Particles` is reached through `FluidParticleWorld` (fluidparticleworld.rs:92-101).
//
// Construction site: src/main.rs:98-101 --
//     Solver::GPU => Box::new(sph::GpuSolver::new_dfsph(&fluid_world, &time_manager, xsph_epsilon, 1 << 20, 1 << 16)),
// nothing else in the application changes: after `simulation_step` the world's `positions`, `velocities` and `densities` hold
// the post-step state in the new sorted order, as the CPU solvers leave them (SURVEY.md 8b "Ownership").
use std::os::raw::c_void;
use std::time::Duration;

use yasph2d_gpu_sys as sys;

use super::super::fluidparticleworld::FluidParticleWorld;
use super::super::timemanager::{AdaptiveTimeStepTarget, SimulationStepConfig, TimeManager};
use super::Solver;

pub struct GpuSolver {
    ctx: *mut sys::yasph_ctx, // !Send / !Sync like the world itself (scratch_buffer.rs:51-52): one caller thread
    registered: Vec<(*mut c_void, usize)>, // page-locked particle Vecs (pointer, bytes)
    device_step: Duration,    // the step length the device holds
}

fn check(ctx: *const sys::yasph_ctx, rc: i32, what: &str) {
    // the reference panics on failure (assert!, unwrap); so does the wrapper.  YASPH_ERR_NONFINITE == assert!(avg.is_finite()) dfsph.rs:223,378
    assert!(rc == sys::YASPH_OK, "{}: status {}: {}", what, rc, sys::last_error(ctx));
}

impl GpuSolver {
    fn create(world: &FluidParticleWorld, time: &TimeManager, solver: i32, viscosity: i32, viscosity_param: f32, max_particles: u32, max_boundary: u32) -> Self {
        let p = &world.properties;
        let mut cfg = unsafe { std::mem::zeroed::<sys::yasph_config>() };
        // every field gets the reference's default for this world; smoothing_length etc. are then overwritten with the world's own
        // values so that both sides use bit-identical constants
        let particle_density = p.fluid_density() / p.particle_mass(); // fluidparticleworld.rs:74-76
        unsafe { sys::yasph_config_default(&mut cfg, 2.0, particle_density, p.fluid_density(), solver) };
        cfg.smoothing_length = p.smoothing_length();
        cfg.gravity = [world.gravity.x, world.gravity.y];
        cfg.viscosity = viscosity;
        cfg.viscosity_param = viscosity_param;
        cfg.max_particles = max_particles;
        cfg.max_boundary = max_boundary;
        match &time.config().step_config {
            // timemanager.rs:38-59
            SimulationStepConfig::FixedTimeStep(step) => {
                cfg.adaptive_timestep = 0;
                cfg.timestep_fixed_ns = step.as_nanos() as u64;
            }
            SimulationStepConfig::AdaptiveTimeStep { timestep_max, timestep_min, timestep_target_frame, cfl_factor } => {
                cfg.adaptive_timestep = 1;
                cfg.timestep_min_ns = timestep_min.as_nanos() as u64;
                cfg.timestep_max_ns = timestep_max.as_nanos() as u64;
                cfg.cfl_factor = *cfl_factor;
                cfg.timestep_target_frame_ns = match timestep_target_frame {
                    AdaptiveTimeStepTarget::None => 0,
                    AdaptiveTimeStepTarget::TargetFrameLength(d) => d.as_nanos() as u64,
                };
            }
        }
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::yasph_create(&cfg, &mut ctx) };
        check(std::ptr::null(), rc, "yasph_create");
        GpuSolver { ctx, registered: Vec::new(), device_step: time.simulation_step() }
    }

    /// DFSPHSolver::new(XSPHViscosityModel, smoothing_length) (dfsph.rs:43-61, xsph.rs:12-16)
    pub fn new_dfsph(world: &FluidParticleWorld, time: &TimeManager, xsph_epsilon: f32, max_particles: u32, max_boundary: u32) -> Self {
        Self::create(world, time, sys::YASPH_SOLVER_DFSPH, sys::YASPH_VISCOSITY_XSPH, xsph_epsilon, max_particles, max_boundary)
    }
    /// WCSPHSolver::new(XSPHViscosityModel, &ConstantFluidProperties) (wscsph.rs:29-41)
    pub fn new_wcsph(world: &FluidParticleWorld, time: &TimeManager, xsph_epsilon: f32, max_particles: u32, max_boundary: u32) -> Self {
        Self::create(world, time, sys::YASPH_SOLVER_WCSPH, sys::YASPH_VISCOSITY_XSPH, xsph_epsilon, max_particles, max_boundary)
    }

    /// Page-locks a particle Vec's allocation once (and again after it re-allocated): with pinned arrays `yasph_step_host` hands
    /// positions and densities back while the divergence solve still runs.  Pageable arrays work too, just later.
    fn pin<T>(&mut self, v: &mut Vec<T>) {
        let (ptr, bytes) = (v.as_mut_ptr() as *mut c_void, v.capacity() * std::mem::size_of::<T>());
        if bytes == 0 || self.registered.iter().any(|r| r.0 == ptr && r.1 == bytes) {
            return;
        }
        self.registered.retain(|r| {
            // a re-allocated Vec left its old registration behind
            if r.0 == ptr {
                unsafe { sys::cudaHostUnregister(r.0) };
                false
            } else {
                true
            }
        });
        if unsafe { sys::cudaHostRegister(ptr, bytes, 0) } == 0 {
            self.registered.push((ptr, bytes));
        }
    }

    pub fn last_report(&self) -> sys::yasph_step_report {
        sys::yasph_step_report::default()
    }
}

impl GpuSolver {
    /// The frame loop's inner part in ONE call (`main.rs:339-360`: `update()` performs simulation steps while
    /// `simulation_frame_loop()` answers `PerformStepAndCallAgain`): `steps` simulation steps on the device-resident state, no
    /// particle array crosses PCIe.  `yasph_step_n` starts each step's first pass ahead of the read-back that ends the previous one;
    /// the reports carry every step's `dt_ns`, from which the caller advances `TimeManager` exactly as `steps` calls of
    /// `simulation_step` would.  The host `Vec`s are stale afterwards: `download()` refreshes them when the renderer wants a frame.
    pub fn simulation_steps_resident(&mut self, steps: u32, time_manager: &mut TimeManager) -> Vec<sys::yasph_step_report> {
        let mut reps = vec![sys::yasph_step_report::default(); steps as usize];
        check(self.ctx, unsafe { sys::yasph_step_n(self.ctx, steps, reps.as_mut_ptr()) }, "yasph_step_n");
        if let Some(last) = reps.last() {
            self.device_step = Duration::from_nanos(last.dt_ns);
            time_manager.set_simulation_step(self.device_step);
        }
        reps
    }

    /// Device -> host `Vec`s (positions, velocities, densities in the current sorted order).
    pub fn download(&mut self, fluid_world: &mut FluidParticleWorld) {
        let parts = &mut fluid_world.particles;
        check(
            self.ctx,
            unsafe {
                sys::yasph_download_particles(self.ctx, parts.positions.as_mut_ptr() as *mut f32, parts.velocities.as_mut_ptr() as *mut f32, parts.densities.as_mut_ptr())
            },
            "yasph_download_particles",
        );
    }
}

impl Solver for GpuSolver {
    fn clear_cached_data(&mut self) {
        // dfsph.rs:406-412 / wscsph.rs:122-124
        check(self.ctx, unsafe { sys::yasph_clear_cached(self.ctx) }, "yasph_clear_cached");
    }

    fn simulation_step(&mut self, fluid_world: &mut FluidParticleWorld, time_manager: &mut TimeManager) {
        // update_neighborhood_datastructure's boundary branch (fluidparticleworld.rs:247-252): the boundary particles are sorted
        // once; the reference sorts them in place, so the sorted order is read back for the renderer (main.rs:250-258)
        if fluid_world.take_boundary_changed() {
            let b = &mut fluid_world.particles.boundary_particles;
            check(self.ctx, unsafe { sys::yasph_set_boundary(self.ctx, b.as_ptr() as *const f32, b.len() as u32) }, "yasph_set_boundary");
            check(
                self.ctx,
                unsafe { sys::yasph_download_field(self.ctx, sys::YASPH_FIELD_BOUNDARY, b.as_mut_ptr() as *mut c_void, (b.len() * 8) as u64) },
                "yasph_download_field(boundary)",
            );
        }
        // TimeManager -> device: the previous step's length (dfsph.rs:433) if somebody else changed it (restart), and for
        // TargetFrameLength stepping the total the frame loop has accumulated (timemanager.rs:246, read by :268-274)
        if time_manager.simulation_step() != self.device_step {
            check(self.ctx, unsafe { sys::yasph_time_set_step_ns(self.ctx, time_manager.simulation_step().as_nanos() as u64) }, "yasph_time_set_step_ns");
        }
        if let SimulationStepConfig::AdaptiveTimeStep { timestep_target_frame: AdaptiveTimeStepTarget::TargetFrameLength(_), .. } = &time_manager.config().step_config {
            check(
                self.ctx,
                unsafe { sys::yasph_time_set_total_simulated_ns(self.ctx, time_manager.total_simulated_time().as_nanos() as u64) },
                "yasph_time_set_total_simulated_ns",
            );
        }
        let n = fluid_world.particles.positions.len();
        fluid_world.particles.velocities.resize(n, cgmath::Zero::zero());
        fluid_world.particles.densities.resize(n, 0.0);
        {
            let parts = &mut fluid_world.particles;
            // (the borrow checker wants the three Vecs pinned one after the other)
            let (mut p, mut v, mut d) = (std::mem::take(&mut parts.positions), std::mem::take(&mut parts.velocities), std::mem::take(&mut parts.densities));
            self.pin(&mut p);
            self.pin(&mut v);
            self.pin(&mut d);
            parts.positions = p;
            parts.velocities = v;
            parts.densities = d;
        }
        let parts = &mut fluid_world.particles;
        let mut rep = sys::yasph_step_report::default();
        // cgmath::Point2<f32> / Vector2<f32> are #[repr(C)] {x, y}: the Vecs are the interleaved arrays the ABI takes
        let rc = unsafe {
            sys::yasph_step_host(self.ctx, parts.positions.as_mut_ptr() as *mut f32, parts.velocities.as_mut_ptr() as *mut f32, parts.densities.as_mut_ptr(), n as u32, &mut rep)
        };
        check(self.ctx, rc, "yasph_step_host");
        // device -> TimeManager: update_simulation_step (timemanager.rs:252-279) was evaluated on the device
        self.device_step = Duration::from_nanos(rep.dt_ns);
        time_manager.set_simulation_step(self.device_step);
        // the reference's println! diagnostics
        if rep.neighbors_capped > 0 {
            println!("particle has too many neighbors ({} particles)", rep.neighbors_capped); // neighborhood_search.rs:361,376
        }
        if rep.not_converged & 1 != 0 {
            println!("density correction did not converge: {} iterations, avg error {}", rep.iters_density, rep.avg_density_error); // dfsph.rs:237
        }
        if rep.not_converged & 2 != 0 {
            println!("divergence correction did not converge: {} iterations, avg {}", rep.iters_divergence, rep.avg_divergence); // dfsph.rs:392
        }
    }
}

impl Drop for GpuSolver {
    fn drop(&mut self) {
        for r in &self.registered {
            unsafe { sys::cudaHostUnregister(r.0) };
        }
        unsafe { sys::yasph_destroy(self.ctx) };
    }
}
