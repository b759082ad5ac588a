"""ctypes binding of oracle/libyasph_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It is the parity checker for the CUDA path, never part of the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libyasph_oracle.so")

K_WENDLAND, K_POLY6, K_SPIKY, K_CUBIC, K_VISCOSITY = 0, 1, 2, 3, 4
VISC_XSPH, VISC_PHYSICAL = 0, 1
MAX_NUM_NEIGHBORS = 64


class StepReport(C.Structure):
    _fields_ = [
        ("dt_ns", C.c_uint64),
        ("dt", C.c_float),
        ("max_velocity", C.c_float),
        ("iters_density", C.c_uint32),
        ("iters_divergence", C.c_uint32),
        ("avg_density_error", C.c_float),
        ("avg_divergence", C.c_float),
        ("warm_density", C.c_uint32),
        ("warm_divergence", C.c_uint32),
        ("not_converged", C.c_uint32),
        ("pad", C.c_uint32),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, -ffp-contract=off)."""
    src = os.path.join(_HERE, "yasph_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    f32p = C.POINTER(C.c_float)
    u32p = C.POINTER(C.c_uint32)
    u16p = C.POINTER(C.c_uint16)
    u64p = C.POINTER(C.c_uint64)
    vp = C.c_void_p

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("yo_num_threads", C.c_int, C.c_int)
    for n in ("yo_morton_encode", "yo_morton_encode_lookup"):
        sig(n, C.c_uint32, C.c_uint32, C.c_uint32)
    sig("yo_morton_decode_x", C.c_uint32, C.c_uint32)
    sig("yo_morton_decode_y", C.c_uint32, C.c_uint32)
    sig("yo_find_bigmin", C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32)
    sig("yo_is_in_rect", C.c_int, C.c_uint32, C.c_uint32, C.c_uint32)
    sig("yo_position_to_cidx", C.c_uint32, C.c_float, C.c_float, C.c_float)
    sig("yo_kernel_evaluate", C.c_float, C.c_int, C.c_float, C.c_float, C.c_float)
    sig("yo_kernel_gradient", None, C.c_int, C.c_float, C.c_float, C.c_float, f32p)
    sig("yo_kernel_laplacian", C.c_float, C.c_float, C.c_float)
    sig("yo_duration_from_secs_f32", C.c_uint64, C.c_float)
    sig("yo_duration_as_secs_f32", C.c_float, C.c_uint64)
    sig("yo_rng_fill", None, C.c_uint64, f32p, C.c_uint32)
    sig("yo_world_new", vp, C.c_float, C.c_float, C.c_float)
    sig("yo_world_new_h", vp, C.c_float, C.c_float, C.c_float)
    sig("yo_world_free", None, vp)
    sig("yo_world_add_fluid_rect", None, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float)
    sig("yo_world_add_boundary_line", None, vp, C.c_float, C.c_float, C.c_float, C.c_float)
    sig("yo_world_add_boundary_thick_line", None, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32)
    sig("yo_world_num_particles", C.c_uint32, vp)
    sig("yo_world_num_boundary", C.c_uint32, vp)
    sig("yo_world_props", None, vp, f32p)
    sig("yo_world_set_gravity", None, vp, C.c_float, C.c_float)
    sig("yo_world_get", None, vp, f32p, f32p, f32p, f32p)
    sig("yo_world_set_particles", None, vp, f32p, f32p, C.c_uint32)
    sig("yo_world_set_boundary", None, vp, f32p, C.c_uint32)
    sig("yo_world_update_neighborhood", None, vp)
    sig("yo_world_update_densities", None, vp, C.c_int)
    sig("yo_world_last_sorting", None, vp, u32p)
    sig("yo_world_num_cells", C.c_uint32, vp, C.c_int)
    sig("yo_world_cells", None, vp, C.c_int, u32p, u32p)
    sig("yo_world_runs", None, vp, C.c_int, C.c_uint32, u32p)
    sig("yo_world_neighbors", None, vp, u16p, u16p, u32p)
    sig("yo_world_neighbor_stats", C.c_uint64, vp, u64p, u64p)
    sig("yo_time_new", vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float)
    sig("yo_time_free", None, vp)
    sig("yo_time_simulation_step", C.c_uint64, vp)
    sig("yo_time_update", C.c_uint64, vp, C.c_float, C.c_float)
    sig("yo_time_set_step", None, vp, C.c_uint64)
    sig("yo_time_set_target", None, vp, C.c_uint64)
    sig("yo_time_perform_step", None, vp)
    sig("yo_time_total", C.c_uint64, vp)
    sig("yo_dfsph_new", vp, vp, C.c_int, C.c_float)
    sig("yo_dfsph_free", None, vp)
    sig("yo_dfsph_clear", None, vp)
    sig("yo_dfsph_step", None, vp, vp, vp, C.POINTER(StepReport))
    sig("yo_dfsph_set_params", None, vp, C.c_float, C.c_uint32, C.c_float, C.c_uint32)
    sig("yo_dfsph_get", None, vp, f32p, f32p, f32p)
    sig("yo_dfsph_alpha", None, vp, vp, f32p)
    sig("yo_wcsph_new", vp, vp, C.c_int, C.c_float)
    sig("yo_wcsph_free", None, vp)
    sig("yo_wcsph_clear", None, vp)
    sig("yo_wcsph_step", None, vp, vp, vp, C.POINTER(StepReport))
    sig("yo_wcsph_get", None, vp, f32p, f32p)
    _lib = L
    return L


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def _u16p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


def num_threads(set_to=0):
    return lib().yo_num_threads(int(set_to))


class World:
    """Restated FluidParticleWorld (fluidparticleworld.rs:92-262)."""

    def __init__(self, smoothing_factor=2.0, particle_density=10000.0, fluid_density=100.0, h=None):
        L = lib()
        if h is None:
            self.h_ = L.yo_world_new(smoothing_factor, particle_density, fluid_density)
        else:
            self.h_ = L.yo_world_new_h(h, particle_density, fluid_density)

    def __del__(self):
        if getattr(self, "h_", None) and _lib is not None:
            _lib.yo_world_free(self.h_)
            self.h_ = None

    # scene builders
    def add_fluid_rect(self, x, y, w, h, jitter):
        lib().yo_world_add_fluid_rect(self.h_, x, y, w, h, jitter)

    def add_boundary_line(self, s, e):
        lib().yo_world_add_boundary_line(self.h_, s[0], s[1], e[0], e[1])

    def add_boundary_thick_line(self, s, e, thickness):
        lib().yo_world_add_boundary_thick_line(self.h_, s[0], s[1], e[0], e[1], thickness)

    @property
    def n(self):
        return lib().yo_world_num_particles(self.h_)

    @property
    def m(self):
        return lib().yo_world_num_boundary(self.h_)

    def props(self):
        out = np.zeros(6, np.float32)
        lib().yo_world_props(self.h_, _f32p(out))
        return dict(h=out[0], mass=out[1], radius=out[2], rho0=out[3], gravity=(out[4], out[5]))

    def set_gravity(self, gx, gy):
        lib().yo_world_set_gravity(self.h_, gx, gy)

    def positions(self):
        a = np.zeros((self.n, 2), np.float32)
        lib().yo_world_get(self.h_, _f32p(a), None, None, None)
        return a

    def velocities(self):
        a = np.zeros((self.n, 2), np.float32)
        lib().yo_world_get(self.h_, None, _f32p(a), None, None)
        return a

    def densities(self):
        a = np.zeros(self.n, np.float32)
        lib().yo_world_get(self.h_, None, None, _f32p(a), None)
        return a

    def boundary(self):
        a = np.zeros((self.m, 2), np.float32)
        lib().yo_world_get(self.h_, None, None, None, _f32p(a))
        return a

    def set_particles(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32) if vel is not None else None
        lib().yo_world_set_particles(self.h_, _f32p(pos), _f32p(vel), len(pos))

    def set_boundary(self, b):
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 2)
        lib().yo_world_set_boundary(self.h_, _f32p(b), len(b))

    def update_neighborhood(self):
        lib().yo_world_update_neighborhood(self.h_)

    def update_densities(self, kernel=K_WENDLAND):
        lib().yo_world_update_densities(self.h_, kernel)

    def last_sorting(self):
        a = np.zeros(self.n, np.uint32)
        lib().yo_world_last_sorting(self.h_, _u32p(a))
        return a

    def cells(self, static=False):
        c = lib().yo_world_num_cells(self.h_, int(static))
        fp = np.zeros(c, np.uint32)
        ci = np.zeros(c, np.uint32)
        lib().yo_world_cells(self.h_, int(static), _u32p(fp), _u32p(ci))
        return fp, ci

    def runs(self, cidx, static=False):
        out = np.zeros((5, 2), np.uint32)
        lib().yo_world_runs(self.h_, int(static), int(cidx), _u32p(out))
        return out

    def neighbors(self, with_lists=True):
        n = self.n
        cd = np.zeros(n, np.uint16)
        ct = np.zeros(n, np.uint16)
        lists = np.zeros((n, MAX_NUM_NEIGHBORS), np.uint32) if with_lists else None
        lib().yo_world_neighbors(self.h_, _u16p(cd), _u16p(ct), _u32p(lists))
        return cd, ct, lists

    def neighbor_stats(self):
        capped = C.c_uint64(0)
        drops = C.c_uint64(0)
        total = lib().yo_world_neighbor_stats(self.h_, C.byref(capped), C.byref(drops))
        return dict(total=total, capped=capped.value, static_drops=drops.value)


class TimeManager:
    """Restated TimeManager step logic (timemanager.rs:104-138,252-279)."""

    def __init__(self, adaptive=True, fixed_ns=0, min_ns=None, max_ns=None, cfl_factor=1.5, target_frame_ns=0):
        L = lib()
        if min_ns is None:
            min_ns = L.yo_duration_from_secs_f32(np.float32(1.0) / np.float32(60.0) / np.float32(400.0))  # main.rs:124
        if max_ns is None:
            max_ns = L.yo_duration_from_secs_f32(np.float32(1.0) / np.float32(120.0) / np.float32(3.0))  # main.rs:123
        self.min_ns, self.max_ns, self.cfl_factor, self.adaptive, self.fixed_ns = min_ns, max_ns, cfl_factor, adaptive, fixed_ns
        self.h_ = L.yo_time_new(int(adaptive), fixed_ns, min_ns, max_ns, cfl_factor)
        self.target_frame_ns = int(target_frame_ns)
        if target_frame_ns:
            L.yo_time_set_target(self.h_, int(target_frame_ns))

    def __del__(self):
        if getattr(self, "h_", None) and _lib is not None:
            _lib.yo_time_free(self.h_)
            self.h_ = None

    def simulation_step_ns(self):
        return lib().yo_time_simulation_step(self.h_)

    def update_simulation_step(self, diameter, max_velocity):
        return lib().yo_time_update(self.h_, diameter, max_velocity)

    def set_step_ns(self, ns):
        lib().yo_time_set_step(self.h_, ns)

    def perform_step(self):
        """The frame loop's bookkeeping ahead of a step (timemanager.rs:243-247): total_simulated_time += simulation_step."""
        lib().yo_time_perform_step(self.h_)

    def total_simulated_ns(self):
        return lib().yo_time_total(self.h_)


class DFSPHSolver:
    def __init__(self, world, visc_kind=VISC_XSPH, visc_param=0.05):
        self.h_ = lib().yo_dfsph_new(world.h_, visc_kind, visc_param)

    def __del__(self):
        if getattr(self, "h_", None) and _lib is not None:
            _lib.yo_dfsph_free(self.h_)
            self.h_ = None

    def clear_cached_data(self):
        lib().yo_dfsph_clear(self.h_)

    def set_params(self, max_avg_density_error=0.0, max_density_iters=0, max_divergence_error=0.0, max_divergence_iters=0):
        """dfsph.rs:49-50,53-54 (plain fields of DFSPHSolver); 0 keeps the current value."""
        lib().yo_dfsph_set_params(self.h_, max_avg_density_error, int(max_density_iters), max_divergence_error, int(max_divergence_iters))

    def simulation_step(self, world, time):
        rep = StepReport()
        lib().yo_dfsph_step(self.h_, world.h_, time.h_, C.byref(rep))
        return rep

    def state(self, n):
        a = np.zeros(n, np.float32)
        k = np.zeros(n, np.float32)
        s = np.zeros(n, np.float32)
        lib().yo_dfsph_get(self.h_, _f32p(a), _f32p(k), _f32p(s))
        return a, k, s

    def alpha_factors(self, world):
        a = np.zeros(world.n, np.float32)
        lib().yo_dfsph_alpha(self.h_, world.h_, _f32p(a))
        return a


class WCSPHSolver:
    def __init__(self, world, visc_kind=VISC_XSPH, visc_param=0.05):
        self.h_ = lib().yo_wcsph_new(world.h_, visc_kind, visc_param)

    def __del__(self):
        if getattr(self, "h_", None) and _lib is not None:
            _lib.yo_wcsph_free(self.h_)
            self.h_ = None

    def clear_cached_data(self):
        lib().yo_wcsph_clear(self.h_)

    def simulation_step(self, world, time):
        rep = StepReport()
        lib().yo_wcsph_step(self.h_, world.h_, time.h_, C.byref(rep))
        return rep

    def accelerations(self, n):
        a = np.zeros((n, 2), np.float32)
        lib().yo_wcsph_get(self.h_, _f32p(a), None)
        return a

    def stiffness(self):
        s = C.c_float(0)
        lib().yo_wcsph_get(self.h_, None, C.byref(s))
        return s.value


def dam_break_scene(world):
    """The app's scene (main.rs:177-196)."""
    world.add_fluid_rect(0.1, 0.7, 0.5, 1.0, 0.05)
    world.add_boundary_thick_line((0.0, 2.5), (2.0, 2.5), 4)
    world.add_boundary_thick_line((0.0, 0.0), (2.0, 0.0), 4)
    world.add_boundary_thick_line((0.0, 0.0), (0.0, 2.5), 4)
    world.add_boundary_thick_line((2.0, 0.0), (2.0, 2.5), 4)
    world.add_boundary_thick_line((0.0, 0.6), (1.75, 0.5), 2)
    world.add_boundary_thick_line((0.0, 2.5), (2.0, 2.5), 2)
    world.add_boundary_thick_line((-2.0, -0.5), (4.0, -0.5), 4)
    return world


def tank_scene(world, columns, rows, x0=1.0, y0=0.2, wall_thickness=4, jitter=0.05, tank_height=None, obstacle=True):
    """The dam-break tank of BASELINE.json configs 3 / 4, built with the oracle's own restated scene builders
    (fluidparticleworld.rs:140-195) -- the same geometry as yasph2d_b200.tank_scene, so that bench.py's reference arm needs nothing
    from the product library.  `world` must be empty."""
    nppm = 100.0 * 0.9  # num_particles_per_meter = sqrt(particle_density) of the application's world (main.rs:85-89), 0.9x packing
    w = (columns + 0.2) / nppm
    h = (rows + 0.2) / nppm
    world.add_fluid_rect(x0, y0, w, h, jitter)
    assert world.n == columns * rows, (world.n, columns, rows)
    x1 = x0 + w * 1.34 + 1.0
    top = tank_height if tank_height is not None else y0 + h * 1.45
    world.add_boundary_thick_line((0.0, 0.0), (x1, 0.0), wall_thickness)
    world.add_boundary_thick_line((0.0, top), (x1, top), wall_thickness)
    world.add_boundary_thick_line((0.0, 0.0), (0.0, top), wall_thickness)
    world.add_boundary_thick_line((x1, 0.0), (x1, top), wall_thickness)
    if obstacle:
        world.add_boundary_thick_line((x0 + w + 0.3, 0.0), (x0 + w + 0.3 + 0.25 * h, 0.2 * h), 2)
    return world
