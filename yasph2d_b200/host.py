"""Host-side mirror of the reference's solver / particle / neighbourhood surface, driving the CUDA path through the C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to the reference repository):
  FluidParticleWorld, Particles, ConstantFluidProperties   src/sph/fluidparticleworld.rs:11-262
  Solver, DFSPHSolver, WCSPHSolver                         src/sph/solver/mod.rs:12-18, dfsph.rs:43-61,405-525, wscsph.rs:29-41,121-179
  XSPHViscosityModel, PhysicalViscosityModel               src/sph/viscositymodel/xsph.rs, physical.rs
  TimeManager, SimulationStepConfig                        src/sph/timemanager.rs:38-59,104-138,252-279
  NeighborhoodSearch, NeighborLists                        src/sph/neighborhood_search.rs:297-522

The reference is Rust; no Rust toolchain exists in this image, so the host side above the C ABI is mirrored here (and in
INTEGRATION.md as the Rust FFI stub a maintainer would add).  All compute happens in libyasph_gpu.so; nothing in this
module computes particle physics on the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi as capi

f32 = np.float32


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Rect:
    """ggez::graphics::Rect as used by add_fluid_rect (fluidparticleworld.rs:3,140)."""

    def __init__(self, x, y, w, h):
        self.x, self.y, self.w, self.h = float(x), float(y), float(w), float(h)


class ConstantFluidProperties:
    """fluidparticleworld.rs:46-90 (all arithmetic in f32 like the reference)."""

    def __init__(self, smoothing_factor, particle_density, fluid_density):
        self.particle_density = f32(particle_density)
        self._fluid_density = f32(fluid_density)
        self._smoothing_length = f32(2.0) * self.particle_radius_from_particle_density(self.particle_density) * f32(smoothing_factor)

    @staticmethod
    def particle_radius_from_particle_density(particle_density):
        return f32(0.5) / np.sqrt(f32(particle_density))

    def smoothing_length(self):
        return self._smoothing_length

    def fluid_density(self):
        return self._fluid_density

    def particle_mass(self):
        return self._fluid_density / self.particle_density

    def num_particles_per_meter(self):
        return np.sqrt(self.particle_density)

    def particle_radius(self):
        return self.particle_radius_from_particle_density(self.particle_density)


class GpuContext:
    """RAII wrapper of a yasph_ctx."""

    def __init__(self, cfg):
        self.cfg = cfg
        h = C.c_void_p()
        capi.check(capi.lib().yasph_create(C.byref(cfg), C.byref(h)))
        self.h = h
        got = capi.Config()
        capi.check(capi.lib().yasph_get_config(self.h, C.byref(got)), self.h)
        self.cfg = got

    def close(self):
        if getattr(self, "h", None):
            capi.lib().yasph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, status):
        capi.check(status, self.h)

    def counts(self):
        n, m = C.c_uint32(0), C.c_uint32(0)
        self._ck(capi.lib().yasph_num_particles(self.h, C.byref(n), C.byref(m)))
        return n.value, m.value

    def set_boundary(self, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        self._ck(capi.lib().yasph_set_boundary(self.h, _f32p(xy), len(xy)))

    def upload_particles(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 2) if vel is not None else None
        self._ck(capi.lib().yasph_upload_particles(self.h, _f32p(pos), _f32p(vel), len(pos)))

    def download_particles(self, want_pos=True, want_vel=True, want_dens=True):
        n, _ = self.counts()
        pos = np.empty((n, 2), np.float32) if want_pos else None
        vel = np.empty((n, 2), np.float32) if want_vel else None
        dens = np.empty(n, np.float32) if want_dens else None
        self._ck(capi.lib().yasph_download_particles(self.h, _f32p(pos), _f32p(vel), _f32p(dens)))
        return pos, vel, dens

    def field(self, field):
        n, m = self.counts()
        if field in (capi.FIELD_POSITION, capi.FIELD_VELOCITY, capi.FIELD_ACCELERATION):
            out = np.empty((n, 2), np.float32)
        elif field == capi.FIELD_BOUNDARY:
            out = np.empty((m, 2), np.float32)
        elif field in (capi.FIELD_CELL_KEY, capi.FIELD_SORT_PERMUTATION, capi.FIELD_ID):
            out = np.empty(n, np.uint32)
        else:
            out = np.empty(n, np.float32)
        self._ck(capi.lib().yasph_download_field(self.h, field, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def solver_state(self):
        st = capi.SolverState()
        self._ck(capi.lib().yasph_solver_state_get(self.h, C.byref(st)))
        return st

    def set_solver_state(self, step_ns, iters_density=0, iters_divergence=0, initialized=True, total_simulated_ns=0):
        st = capi.SolverState(int(step_ns), int(iters_density), int(iters_divergence), 1 if initialized else 0, 0, int(total_simulated_ns))
        self._ck(capi.lib().yasph_solver_state_set(self.h, C.byref(st)))

    def upload_field(self, field, data):
        data = np.ascontiguousarray(data, np.float32)
        self._ck(capi.lib().yasph_upload_field(self.h, field, data.ctypes.data_as(C.c_void_p), data.nbytes))

    def clear_cached(self):
        self._ck(capi.lib().yasph_clear_cached(self.h))

    def set_flags(self, flags):
        self._ck(capi.lib().yasph_set_flags(self.h, flags))

    def step(self):
        rep = capi.StepReport()
        self._ck(capi.lib().yasph_step(self.h, C.byref(rep)))
        return rep

    def step_n(self, steps):
        """`steps` simulation steps in one call (the application's frame loop); returns the list of their reports."""
        reps = (capi.StepReport * max(steps, 1))()
        self._ck(capi.lib().yasph_step_n(self.h, steps, reps))
        return list(reps)[:steps]

    def step_host(self, pos, vel, dens=None, input_unchanged=False):
        """The reference-facing call: HOST arrays in, one simulation_step, HOST arrays out (in place).  input_unchanged: the arrays
        still hold what the previous call handed back (the application only read them), so their upload is skipped."""
        assert pos.dtype == np.float32 and vel.dtype == np.float32 and pos.flags.c_contiguous and vel.flags.c_contiguous
        rep = capi.StepReport()
        self._ck(capi.lib().yasph_step_host_ex(self.h, _f32p(pos), _f32p(vel), _f32p(dens), len(pos), capi.HOST_INPUT_UNCHANGED if input_unchanged else 0,
                                               C.byref(rep)))
        return rep

    def time_step_ns(self):
        v = C.c_uint64(0)
        self._ck(capi.lib().yasph_time_get_step_ns(self.h, C.byref(v)))
        return v.value

    def set_total_simulated_ns(self, ns):
        self._ck(capi.lib().yasph_time_set_total_simulated_ns(self.h, int(ns)))

    def total_simulated_ns(self):
        v = C.c_uint64(0)
        self._ck(capi.lib().yasph_time_get_total_simulated_ns(self.h, C.byref(v)))
        return v.value

    def set_time_step_ns(self, ns):
        self._ck(capi.lib().yasph_time_set_step_ns(self.h, int(ns)))

    def neighborhood_update(self):
        rep = capi.StepReport()
        self._ck(capi.lib().yasph_neighborhood_update(self.h, C.byref(rep)))
        return rep

    def neighbors(self, with_lists=True):
        n, _ = self.counts()
        cd = np.zeros(n, np.uint16)
        ct = np.zeros(n, np.uint16)
        lists = np.zeros((n, capi.MAX_NEIGHBORS), np.uint32) if with_lists else None
        self._ck(capi.lib().yasph_neighbors_download(
            self.h, cd.ctypes.data_as(C.POINTER(C.c_uint16)), ct.ctypes.data_as(C.POINTER(C.c_uint16)),
            lists.ctypes.data_as(C.POINTER(C.c_uint32)) if with_lists else None))
        return cd, ct, lists

    def update_densities(self, kernel=capi.KERNEL_WENDLAND_C2):
        self._ck(capi.lib().yasph_update_densities(self.h, kernel))

    def compute_alpha(self):
        self._ck(capi.lib().yasph_compute_alpha(self.h))

    def pass_times_us(self):
        out = np.zeros(capi.NUM_PASSES, np.float32)
        self._ck(capi.lib().yasph_pass_times(self.h, _f32p(out)))
        return dict(zip(capi.PASS_NAMES, out.tolist()))

    def host_step_times_us(self):
        """Device timeline of the last step_host call (see yasph_host_step_times)."""
        out = np.zeros(6, np.float32)
        self._ck(capi.lib().yasph_host_step_times(self.h, _f32p(out)))
        return dict(zip(("start", "uploaded", "positions_on_host", "densities_on_host", "kernels_done", "velocities_on_host"), out.tolist()))

    def launch_count(self):
        v = C.c_uint64(0)
        self._ck(capi.lib().yasph_launch_count(self.h, C.byref(v)))
        return v.value

    def stream_handle(self):
        v = C.c_void_p()
        self._ck(capi.lib().yasph_stream(self.h, C.byref(v)))
        return v.value


# ----------------------------------------------------------------------------------------------------------------------
# scene builders (host only; fluidparticleworld.rs:140-195)
# ----------------------------------------------------------------------------------------------------------------------
def scene_fluid_rect(particle_density, rect, jitter, seed):
    L = capi.lib()
    cnt = C.c_uint32(0)
    capi.check(L.yasph_scene_fluid_rect(particle_density, rect.x, rect.y, rect.w, rect.h, jitter, seed, None, 0, C.byref(cnt)))
    out = np.empty((cnt.value, 2), np.float32)
    capi.check(L.yasph_scene_fluid_rect(particle_density, rect.x, rect.y, rect.w, rect.h, jitter, seed, _f32p(out), cnt.value, C.byref(cnt)))
    return out


def scene_boundary_line(particle_density, start, end):
    L = capi.lib()
    cnt = C.c_uint32(0)
    capi.check(L.yasph_scene_boundary_line(particle_density, start[0], start[1], end[0], end[1], None, 0, C.byref(cnt)))
    out = np.empty((cnt.value, 2), np.float32)
    capi.check(L.yasph_scene_boundary_line(particle_density, start[0], start[1], end[0], end[1], _f32p(out), cnt.value, C.byref(cnt)))
    return out


def scene_boundary_thick_line(particle_density, start, end, thickness):
    L = capi.lib()
    cnt = C.c_uint32(0)
    capi.check(L.yasph_scene_boundary_thick_line(particle_density, start[0], start[1], end[0], end[1], thickness, None, 0, C.byref(cnt)))
    out = np.empty((cnt.value, 2), np.float32)
    capi.check(L.yasph_scene_boundary_thick_line(particle_density, start[0], start[1], end[0], end[1], thickness, _f32p(out), cnt.value,
                                                 C.byref(cnt)))
    return out


class Particles:
    """fluidparticleworld.rs:11-44.  Host arrays, like the Vecs the Rust side owns."""

    def __init__(self):
        self.positions = np.zeros((0, 2), np.float32)
        self.velocities = np.zeros((0, 2), np.float32)
        self.densities = np.zeros(0, np.float32)
        self.boundary_particles = np.zeros((0, 2), np.float32)

    def num_dynamic_particles(self):
        return len(self.positions)

    def num_boundary_particles(self):
        return len(self.boundary_particles)


class FluidParticleWorld:
    """fluidparticleworld.rs:92-262."""

    def __init__(self, smoothing_factor, particle_density, fluid_density):
        self.smoothing_factor = float(smoothing_factor)
        self.properties = ConstantFluidProperties(smoothing_factor, particle_density, fluid_density)
        self.particles = Particles()
        self.gravity = (0.0, -9.81)  # fluidparticleworld.rs:123
        self.boundary_changed = True
        self.host_dirty = True  # host arrays changed since the last upload

    def remove_all_fluid_particles(self):  # :129-132
        self.particles.positions = np.zeros((0, 2), np.float32)
        self.particles.velocities = np.zeros((0, 2), np.float32)
        self.host_dirty = True

    def remove_all_boundary_particles(self):  # :134-137 (also clears velocities, quirk Q9)
        self.particles.boundary_particles = np.zeros((0, 2), np.float32)
        self.particles.velocities = np.zeros((0, 2), np.float32)
        self.boundary_changed = True

    def add_fluid_rect(self, fluid_rect, jitter_amount):  # :140-166
        p = self.particles
        new = scene_fluid_rect(float(self.properties.particle_density), fluid_rect, jitter_amount, len(p.positions))
        total = len(p.positions) + len(new)
        p.positions = np.concatenate([p.positions, new]).astype(np.float32)
        vel = np.zeros((total, 2), np.float32)
        vel[: min(len(p.velocities), total)] = p.velocities[:total]
        p.velocities = vel
        dens = np.zeros(total, np.float32)
        dens[: min(len(p.densities), total)] = p.densities[:total]
        p.densities = dens
        self.host_dirty = True

    def add_boundary_thick_line(self, start, end, thickness_in_particles):  # :168-179
        new = scene_boundary_thick_line(float(self.properties.particle_density), start, end, thickness_in_particles)
        self.particles.boundary_particles = np.concatenate([self.particles.boundary_particles, new]).astype(np.float32)
        self.boundary_changed = True

    def add_boundary_line(self, start, end):  # :181-195
        new = scene_boundary_line(float(self.properties.particle_density), start, end)
        self.particles.boundary_particles = np.concatenate([self.particles.boundary_particles, new]).astype(np.float32)
        self.boundary_changed = True


def dam_break_scene(world):
    """The application's scene, main.rs:177-196."""
    world.remove_all_fluid_particles()
    world.remove_all_boundary_particles()
    world.add_fluid_rect(Rect(0.1, 0.7, 0.5, 1.0), 0.05)
    world.add_boundary_thick_line((0.0, 2.5), (2.0, 2.5), 4)
    world.add_boundary_thick_line((0.0, 0.0), (2.0, 0.0), 4)
    world.add_boundary_thick_line((0.0, 0.0), (0.0, 2.5), 4)
    world.add_boundary_thick_line((2.0, 0.0), (2.0, 2.5), 4)
    world.add_boundary_thick_line((0.0, 0.6), (1.75, 0.5), 2)
    world.add_boundary_thick_line((0.0, 2.5), (2.0, 2.5), 2)
    world.add_boundary_thick_line((-2.0, -0.5), (4.0, -0.5), 4)
    return world


def tank_scene(world, columns, rows, x0=1.0, y0=0.2, wall_thickness=4, jitter=0.05, tank_height=None, obstacle=True):
    """A dam-break tank scaled to columns x rows fluid particles (BASELINE.json configs 3 and 4).

    Fluid lattice of `columns` x `rows` particles at the reference's 0.9x rest spacing starting at (x0, y0); a closed tank
    around it with one third of free space to the right of the column for the break, and (optionally) the slanted
    obstacle of the application's scene scaled to the tank.
    """
    nppm = float(world.properties.num_particles_per_meter()) * 0.9
    w = (columns + 0.2) / nppm
    h = (rows + 0.2) / nppm
    world.remove_all_fluid_particles()
    world.remove_all_boundary_particles()
    world.add_fluid_rect(Rect(x0, y0, w, h), jitter)
    assert world.particles.num_dynamic_particles() == columns * rows, (world.particles.num_dynamic_particles(), columns, rows)
    x1 = x0 + w * 1.34 + 1.0
    top = tank_height if tank_height is not None else y0 + h * 1.45
    world.add_boundary_thick_line((0.0, 0.0), (x1, 0.0), wall_thickness)
    world.add_boundary_thick_line((0.0, top), (x1, top), wall_thickness)
    world.add_boundary_thick_line((0.0, 0.0), (0.0, top), wall_thickness)
    world.add_boundary_thick_line((x1, 0.0), (x1, top), wall_thickness)
    if obstacle:
        world.add_boundary_thick_line((x0 + w + 0.3, 0.0), (x0 + w + 0.3 + 0.25 * h, 0.2 * h), 2)
    return world


# ----------------------------------------------------------------------------------------------------------------------
# time manager (timemanager.rs)
# ----------------------------------------------------------------------------------------------------------------------
class SimulationStepConfig:
    """timemanager.rs:38-59.  Durations are integer nanoseconds."""

    def __init__(self, adaptive, fixed_ns=0, timestep_min_ns=0, timestep_max_ns=0, cfl_factor=1.0, target_frame_ns=0):
        self.adaptive, self.fixed_ns = bool(adaptive), int(fixed_ns)
        self.timestep_min_ns, self.timestep_max_ns, self.cfl_factor = int(timestep_min_ns), int(timestep_max_ns), float(cfl_factor)
        self.target_frame_ns = int(target_frame_ns)  # AdaptiveTimeStepTarget: 0 = None, else TargetFrameLength (timemanager.rs:23-36)

    @classmethod
    def FixedTimeStep(cls, step_ns):
        return cls(False, fixed_ns=step_ns)

    @classmethod
    def AdaptiveTimeStep(cls, timestep_max_ns=None, timestep_min_ns=None, cfl_factor=1.5, target_frame_ns=0):
        L = capi.lib()
        if timestep_max_ns is None:
            timestep_max_ns = L.yasph_duration_from_secs_f32(f32(1.0) / f32(120.0) / f32(3.0))  # main.rs:123
        if timestep_min_ns is None:
            timestep_min_ns = L.yasph_duration_from_secs_f32(f32(1.0) / f32(60.0) / f32(400.0))  # main.rs:124
        return cls(True, timestep_min_ns=timestep_min_ns, timestep_max_ns=timestep_max_ns, cfl_factor=cfl_factor, target_frame_ns=target_frame_ns)


class TimeManager:
    """The part of TimeManager the solvers touch (timemanager.rs:104-138,252-279); frame pacing stays with the application."""

    def __init__(self, step_config):
        self.step_config = step_config
        self.restart()

    def restart(self):  # :131-133 with :105-109
        c = self.step_config
        self._simulation_step_ns = c.timestep_min_ns if c.adaptive else c.fixed_ns
        self.num_simulation_steps = 0
        self.total_simulated_time_ns = 0

    def simulation_step(self):  # :136-138, nanoseconds
        return self._simulation_step_ns

    def perform_step(self):
        """The bookkeeping simulation_frame_loop does when it lets a step happen (timemanager.rs:243-247); the application
        calls it before Solver.simulation_step.  Only the TargetFrameLength rule reads the total (:268-274)."""
        self.total_simulated_time_ns += self._simulation_step_ns

    def simulation_step_secs(self):
        return capi.lib().yasph_duration_as_secs_f32(self._simulation_step_ns)

    def _set_from_device(self, step_ns):
        """update_simulation_step is evaluated on the device (timemanager.rs:252-279); its result is mirrored here."""
        self._simulation_step_ns = int(step_ns)


# ----------------------------------------------------------------------------------------------------------------------
# viscosity models and solvers
# ----------------------------------------------------------------------------------------------------------------------
class XSPHViscosityModel:
    kind = capi.VISCOSITY_XSPH

    def __init__(self, smoothing_length):
        self.epsilon = 0.05  # xsph.rs:14
        self.smoothing_length = smoothing_length

    @property
    def param(self):
        return self.epsilon


class PhysicalViscosityModel:
    kind = capi.VISCOSITY_PHYSICAL

    def __init__(self, smoothing_length):
        self.fluid_viscosity = float(f32(1.0016) / f32(1000.0))  # physical.rs:15
        self.smoothing_length = smoothing_length

    @property
    def param(self):
        return self.fluid_viscosity


class Solver:
    """trait Solver (solver/mod.rs:12-18), backed by a yasph_ctx."""

    solver_kind = None

    def __init__(self, viscosity_model, device=0, flags=0, max_particles=None, max_boundary=None, **knobs):
        self.viscosity_model = viscosity_model
        self.device, self.flags, self.knobs = device, flags, knobs
        self.max_particles, self.max_boundary = max_particles, max_boundary
        self.ctx = None
        self._ctx_step_ns = None
        self.last_report = None

    def _configure(self, cfg):
        pass

    def _ensure_ctx(self, world, time_manager):
        if self.ctx is not None:
            n, m = world.particles.num_dynamic_particles(), world.particles.num_boundary_particles()
            if n <= self.ctx.cfg.max_particles and m <= self.ctx.cfg.max_boundary:
                return
            self._grow_ctx(world, time_manager, n, m)
            return
        self._create_ctx(world, time_manager)

    def _grow_ctx(self, world, time_manager, n, m):
        """The reference accepts add_fluid_rect / add_boundary_* at any time (fluidparticleworld.rs:140-195); a context has fixed
        capacities, so a world that outgrew them gets a new context with headroom.  The solver's carried state moves over: the
        time step and iteration counts, and the per-particle arrays in their current order (DFSPH warm starts: Vec::resize keeps
        them, dfsph.rs:420-423; WCSPH accelerations, wscsph.rs:128)."""
        old = self.ctx
        st = old.solver_state()
        n_old = old.counts()[0]
        carried = {}
        if n_old:
            if self.solver_kind == capi.SOLVER_DFSPH:
                if st.initialized:
                    carried = {capi.FIELD_KAPPA: old.field(capi.FIELD_KAPPA), capi.FIELD_STIFFNESS: old.field(capi.FIELD_STIFFNESS)}
            else:
                carried = {capi.FIELD_ACCELERATION: old.field(capi.FIELD_ACCELERATION)}
        old.close()
        self.ctx = None
        self.max_particles = max(int(self.max_particles or 0), n + n // 2 + 1024)
        self.max_boundary = max(int(self.max_boundary or 0), m + m // 2 + 1024)
        self._create_ctx(world, time_manager)
        world.boundary_changed = True  # _sync_inputs uploads the boundary into the new context
        if n_old and n_old <= n:
            # the first n_old host particles are the old set in the order of the last step (new fluid is appended, fluidparticleworld.rs:150)
            self.ctx.set_boundary(world.particles.boundary_particles)
            self.ctx.upload_particles(world.particles.positions[:n_old], world.particles.velocities[:n_old])
            self.ctx.set_solver_state(st.step_ns, st.iters_density, st.iters_divergence, bool(st.initialized), st.total_simulated_ns)
            for f, a in carried.items():
                self.ctx.upload_field(f, a)
            self._ctx_step_ns = st.step_ns

    def _create_ctx(self, world, time_manager):
        p = world.properties
        cfg = capi.default_config(world.smoothing_factor, float(p.particle_density), float(p.fluid_density()), self.solver_kind)
        assert f32(cfg.smoothing_length) == p.smoothing_length()
        cfg.device = self.device
        cfg.flags = self.flags
        cfg.gravity[0], cfg.gravity[1] = world.gravity
        cfg.viscosity = self.viscosity_model.kind
        cfg.viscosity_param = self.viscosity_model.param
        n, m = world.particles.num_dynamic_particles(), world.particles.num_boundary_particles()
        cfg.max_particles = self.max_particles or max(n, 1)
        cfg.max_boundary = self.max_boundary or max(m, 1)
        sc = time_manager.step_config
        cfg.adaptive_timestep = int(sc.adaptive)
        cfg.timestep_fixed_ns, cfg.timestep_min_ns, cfg.timestep_max_ns, cfg.cfl_factor = sc.fixed_ns, sc.timestep_min_ns, sc.timestep_max_ns, sc.cfl_factor
        cfg.timestep_target_frame_ns = sc.target_frame_ns
        for k, v in self.knobs.items():
            setattr(cfg, k, v)
        self._configure(cfg)
        self.ctx = GpuContext(cfg)
        self._ctx_step_ns = self.ctx.cfg.timestep_min_ns if sc.adaptive else sc.fixed_ns

    def clear_cached_data(self):
        if self.ctx is not None:
            self.ctx.clear_cached()

    def _sync_inputs(self, world, time_manager):
        self._ensure_ctx(world, time_manager)
        if world.boundary_changed:
            self.ctx.set_boundary(world.particles.boundary_particles)
            world.boundary_changed = False
        if time_manager.simulation_step() != self._ctx_step_ns:
            self.ctx.set_time_step_ns(time_manager.simulation_step())
            self._ctx_step_ns = time_manager.simulation_step()
        if time_manager.step_config.target_frame_ns:  # the rule reads TimeManager's total, which the application advances
            self.ctx.set_total_simulated_ns(time_manager.total_simulated_time_ns)

    def _after(self, rep, time_manager):
        self._ctx_step_ns = rep.dt_ns
        time_manager._set_from_device(rep.dt_ns)
        time_manager.num_simulation_steps += 1
        self.last_report = rep

    def simulation_step(self, fluid_world, time_manager):
        """Solver::simulation_step(&mut world, &mut time_manager): host arrays in, host arrays out (new sorted order)."""
        self._sync_inputs(fluid_world, time_manager)
        p = fluid_world.particles
        if len(p.densities) != len(p.positions):
            p.densities = np.zeros(len(p.positions), np.float32)
        rep = self.ctx.step_host(p.positions, p.velocities, p.densities)
        fluid_world.host_dirty = False
        self._after(rep, time_manager)
        return rep

    def simulation_step_resident(self, fluid_world, time_manager):
        """Same step on the device-resident state; host arrays are refreshed only by download(world)."""
        self._sync_inputs(fluid_world, time_manager)
        if fluid_world.host_dirty:
            self.ctx.upload_particles(fluid_world.particles.positions, fluid_world.particles.velocities)
            fluid_world.host_dirty = False
        rep = self.ctx.step()
        self._after(rep, time_manager)
        return rep

    def download(self, fluid_world):
        p = fluid_world.particles
        p.positions, p.velocities, p.densities = self.ctx.download_particles()
        p.boundary_particles = self.ctx.field(capi.FIELD_BOUNDARY)


class DFSPHSolver(Solver):
    """DFSPHSolver::new(viscosity_model, smoothing_length) (dfsph.rs:43-61)."""

    solver_kind = capi.SOLVER_DFSPH

    def __init__(self, viscosity_model, smoothing_length=None, **kw):
        super().__init__(viscosity_model, **kw)
        self.max_avg_density_error = None
        self.max_divergence_error = None

    def _configure(self, cfg):
        if self.max_avg_density_error is not None:
            cfg.dfsph_max_avg_density_error = self.max_avg_density_error
        if self.max_divergence_error is not None:
            cfg.dfsph_max_divergence_error = self.max_divergence_error


class WCSPHSolver(Solver):
    """WCSPHSolver::new(viscosity_model, &ConstantFluidProperties) (wscsph.rs:29-41)."""

    solver_kind = capi.SOLVER_WCSPH

    def __init__(self, viscosity_model, fluid_properties=None, **kw):
        super().__init__(viscosity_model, **kw)


class NeighborLists:
    """neighborhood_search.rs:297-450: neighbors_dynamic / neighbors_static / num_neighbors."""

    def __init__(self, count_dynamic, count_total, lists):
        self.count_dynamic, self.count_total, self.lists = count_dynamic, count_total, lists

    def neighbors_dynamic(self, particle):
        return self.lists[particle, : self.count_dynamic[particle]]

    def neighbors_static(self, particle):
        return self.lists[particle, self.count_dynamic[particle] : self.count_total[particle]]

    def num_neighbors(self, particle):
        return int(self.count_total[particle])


class NeighborhoodSearch:
    """NeighborhoodSearch::{new, update_static, update_dynamic, neighbor_lists} (neighborhood_search.rs:461-522)."""

    def __init__(self, radius, max_particles=1 << 20, max_boundary=1 << 16, device=0, **knobs):
        cfg = capi.default_config()
        cfg.smoothing_length = radius  # cell_size = radius (neighborhood_search.rs:466)
        cfg.device = device
        cfg.max_particles, cfg.max_boundary = max_particles, max(max_boundary, 1)
        for k, v in knobs.items():
            setattr(cfg, k, v)
        self.ctx = GpuContext(cfg)
        self.last_report = None

    def update_static(self, positions):
        """Sorts `positions` (returned) and builds the static cell grid."""
        self.ctx.set_boundary(positions)
        return self.ctx.field(capi.FIELD_BOUNDARY)

    def update_dynamic(self, positions_dynamic, velocities=None):
        """Re-sorts the particles and rebuilds the neighbour lists; returns (sorted positions, sorted velocities)."""
        self.ctx.upload_particles(positions_dynamic, velocities)
        self.last_report = self.ctx.neighborhood_update()
        pos, vel, _ = self.ctx.download_particles(True, velocities is not None, False)
        return pos, vel

    def neighbor_lists(self):
        return NeighborLists(*self.ctx.neighbors(True))
