"""ctypes binding of libyasph_gpu.so (include/yasph_gpu.h).  No torch types cross this boundary.

The library is built in-tree by ``__graft_entry__.build()`` / ``yasph2d_b200.build.build()``.  There is no fallback:
if the shared library is missing or no sm_100a device is present the calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyasph_gpu.so")

ABI_VERSION = 4
HOST_INPUT_UNCHANGED = 1
MAX_NEIGHBORS = 64
SOLVER_DFSPH, SOLVER_WCSPH = 0, 1
VISCOSITY_XSPH, VISCOSITY_PHYSICAL = 0, 1
KERNEL_WENDLAND_C2, KERNEL_POLY6, KERNEL_SPIKY, KERNEL_CUBIC = 0, 1, 2, 3
(FIELD_POSITION, FIELD_VELOCITY, FIELD_DENSITY, FIELD_ALPHA, FIELD_KAPPA, FIELD_STIFFNESS, FIELD_ACCELERATION, FIELD_CELL_KEY,
 FIELD_SORT_PERMUTATION, FIELD_BOUNDARY, FIELD_ID, FIELD_GHOST) = range(12)
FIELD_LOCAL_BIT = 0x100
FLAG_PERMUTE_WARMSTART, FLAG_PROFILE_PASSES, FLAG_TRACK_IDS, FLAG_NO_PEER_TRANSPORT = 1, 2, 4, 8
COMM_ID_BYTES = 128
NUM_PASSES = 17
PASS_NAMES = ["viscosity", "predict", "density_warm", "density_solve", "advect_keygen", "sort", "gather", "cells_tiles", "lists",
              "density_alpha", "divergence_warm", "divergence_solve", "wcsph_accel", "wcsph_kick", "halo", "migrate", "total"]
STATUS_NAMES = {0: "OK", 1: "INVALID_ARGUMENT", 2: "CUDA", 3: "CAPACITY", 4: "STATE", 5: "NONFINITE", 6: "NO_DEVICE", 7: "COMM"}

# every symbol include/yasph_gpu.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "yasph_config_default", "yasph_create", "yasph_destroy", "yasph_last_error", "yasph_get_config", "yasph_set_flags", "yasph_get_properties",
    "yasph_set_boundary", "yasph_upload_particles", "yasph_download_particles", "yasph_download_field", "yasph_num_particles",
    "yasph_clear_cached", "yasph_step", "yasph_step_n", "yasph_step_host", "yasph_step_host_ex", "yasph_time_get_step_ns", "yasph_time_set_step_ns", "yasph_time_restart", "yasph_time_set_total_simulated_ns", "yasph_time_get_total_simulated_ns",
    "yasph_neighborhood_update", "yasph_neighbors_download", "yasph_update_densities", "yasph_compute_alpha", "yasph_pass_times", "yasph_host_step_times",
    "yasph_solver_state_get", "yasph_solver_state_set", "yasph_upload_field",
    "yasph_launch_count", "yasph_stream", "yasph_scene_fluid_rect", "yasph_scene_boundary_line", "yasph_scene_boundary_thick_line",
    "yasph_duration_from_secs_f32", "yasph_duration_as_secs_f32",
    "yasph_comm_unique_id", "yasph_comm_init", "yasph_slab_set", "yasph_slab_get", "yasph_cell_column", "yasph_step_host_slab", "yasph_step_host_slab_ex",
    "yasph_loopback_create", "yasph_loopback_destroy", "yasph_comm_init_loopback",
]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32), ("max_particles", C.c_uint32), ("max_boundary", C.c_uint32),
        ("smoothing_length", C.c_float), ("particle_density", C.c_float), ("fluid_density", C.c_float),
        ("gravity", C.c_float * 2), ("grid_min", C.c_float * 2),
        ("solver", C.c_int32), ("viscosity", C.c_int32), ("viscosity_param", C.c_float),
        ("dfsph_max_avg_density_error", C.c_float), ("dfsph_max_density_iters", C.c_uint32),
        ("dfsph_max_divergence_error", C.c_float), ("dfsph_max_divergence_iters", C.c_uint32),
        ("wcsph_stiffness", C.c_float), ("wcsph_boundary_force_factor", C.c_float),
        ("adaptive_timestep", C.c_int32), ("timestep_fixed_ns", C.c_uint64), ("timestep_min_ns", C.c_uint64),
        ("timestep_max_ns", C.c_uint64), ("timestep_target_frame_ns", C.c_uint64), ("cfl_factor", C.c_float),
        ("max_tiles", C.c_uint32), ("tile_dynamic_capacity", C.c_uint32), ("tile_static_capacity", C.c_uint32),
        ("speculative_iterations", C.c_uint32), ("flags", C.c_uint32), ("max_halo", C.c_uint32), ("ghost_columns", C.c_uint32),
    ]


class StepReport(C.Structure):
    _fields_ = [
        ("dt_prev_ns", C.c_uint64), ("dt_ns", C.c_uint64), ("dt", C.c_float), ("max_velocity", C.c_float),
        ("iters_density", C.c_uint32), ("iters_divergence", C.c_uint32), ("avg_density_error", C.c_float),
        ("avg_divergence", C.c_float), ("warm_density", C.c_uint32), ("warm_divergence", C.c_uint32),
        ("neighbors_capped", C.c_uint32), ("neighbors_dropped", C.c_uint32), ("not_converged", C.c_uint32),
        ("num_cells", C.c_uint32), ("num_tiles", C.c_uint32), ("list_rebuilds", C.c_uint32), ("total_neighbors", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class SlabInfo(C.Structure):
    _fields_ = [
        ("rank", C.c_int32), ("world", C.c_int32), ("col_lo", C.c_uint32), ("col_hi", C.c_uint32), ("n_own", C.c_uint32), ("n_local", C.c_uint32),
        ("n_ghost_left", C.c_uint32), ("n_ghost_right", C.c_uint32), ("migrated_out_left", C.c_uint32),
        ("migrated_out_right", C.c_uint32), ("migrated_in", C.c_uint32), ("peer_transport", C.c_uint32), ("n_global", C.c_uint64),
        ("halo_exchanges", C.c_uint64), ("allreduces", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SolverState(C.Structure):
    _fields_ = [("step_ns", C.c_uint64), ("iters_density", C.c_uint32), ("iters_divergence", C.c_uint32), ("initialized", C.c_uint32),
                ("reserved", C.c_uint32), ("total_simulated_ns", C.c_uint64)]


class YasphError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("yasph_gpu: %s (%d): %s" % (STATUS_NAMES.get(status, "?"), status, message))
        self.status = status


_lib = None


def lib():
    """Load libyasph_gpu.so.  Raises if it has not been built -- there is no pure-Python / CPU substitute."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libyasph_gpu.so is missing at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "yasph2d_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(os.environ.get("YASPH_GPU_LIB", LIB_PATH))  # the override serves A/B timing of kernel variants (still this library)
    vp, f32p, u32p, u16p, u64p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_uint64)
    rp = C.POINTER(StepReport)

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("yasph_config_default", C.c_int32, C.POINTER(Config), C.c_float, C.c_float, C.c_float, C.c_int32)
    sig("yasph_create", C.c_int32, C.POINTER(Config), C.POINTER(vp))
    sig("yasph_destroy", C.c_int32, vp)
    sig("yasph_last_error", C.c_char_p, vp)
    sig("yasph_get_config", C.c_int32, vp, C.POINTER(Config))
    sig("yasph_set_flags", C.c_int32, vp, C.c_uint32)
    sig("yasph_get_properties", C.c_int32, vp, f32p)
    sig("yasph_set_boundary", C.c_int32, vp, f32p, C.c_uint32)
    sig("yasph_upload_particles", C.c_int32, vp, f32p, f32p, C.c_uint32)
    sig("yasph_download_particles", C.c_int32, vp, f32p, f32p, f32p)
    sig("yasph_download_field", C.c_int32, vp, C.c_int32, vp, C.c_uint64)
    sig("yasph_num_particles", C.c_int32, vp, u32p, u32p)
    sig("yasph_clear_cached", C.c_int32, vp)
    sig("yasph_step", C.c_int32, vp, rp)
    sig("yasph_step_n", C.c_int32, vp, C.c_uint32, rp)
    sig("yasph_step_host", C.c_int32, vp, f32p, f32p, f32p, C.c_uint32, rp)
    sig("yasph_step_host_ex", C.c_int32, vp, f32p, f32p, f32p, C.c_uint32, C.c_uint32, rp)
    sig("yasph_time_get_step_ns", C.c_int32, vp, u64p)
    sig("yasph_time_set_step_ns", C.c_int32, vp, C.c_uint64)
    sig("yasph_time_restart", C.c_int32, vp)
    sig("yasph_time_set_total_simulated_ns", C.c_int32, vp, C.c_uint64)
    sig("yasph_time_get_total_simulated_ns", C.c_int32, vp, u64p)
    sig("yasph_neighborhood_update", C.c_int32, vp, rp)
    sig("yasph_neighbors_download", C.c_int32, vp, u16p, u16p, u32p)
    sig("yasph_update_densities", C.c_int32, vp, C.c_int32)
    sig("yasph_compute_alpha", C.c_int32, vp)
    sig("yasph_pass_times", C.c_int32, vp, f32p)
    sig("yasph_host_step_times", C.c_int32, vp, f32p)
    sig("yasph_solver_state_get", C.c_int32, vp, C.POINTER(SolverState))
    sig("yasph_solver_state_set", C.c_int32, vp, C.POINTER(SolverState))
    sig("yasph_upload_field", C.c_int32, vp, C.c_int32, C.c_void_p, C.c_uint64)
    sig("yasph_launch_count", C.c_int32, vp, u64p)
    sig("yasph_stream", C.c_int32, vp, C.POINTER(vp))
    sig("yasph_scene_fluid_rect", C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64, f32p,
        C.c_uint32, u32p)
    sig("yasph_scene_boundary_line", C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_uint32, u32p)
    sig("yasph_scene_boundary_thick_line", C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32, f32p,
        C.c_uint32, u32p)
    sig("yasph_duration_from_secs_f32", C.c_uint64, C.c_float)
    sig("yasph_duration_as_secs_f32", C.c_float, C.c_uint64)
    sig("yasph_comm_unique_id", C.c_int32, vp, C.c_uint64)
    sig("yasph_comm_init", C.c_int32, vp, C.c_int32, C.c_int32, vp, C.c_uint64)
    sig("yasph_slab_set", C.c_int32, vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32)
    sig("yasph_slab_get", C.c_int32, vp, C.POINTER(SlabInfo))
    sig("yasph_cell_column", C.c_int32, C.POINTER(Config), C.c_float, u32p)
    sig("yasph_loopback_create", C.c_int32, C.c_int32, C.POINTER(vp))
    sig("yasph_loopback_destroy", C.c_int32, vp)
    sig("yasph_comm_init_loopback", C.c_int32, vp, vp, C.c_int32)
    sig("yasph_step_host_slab", C.c_int32, vp, f32p, f32p, f32p, C.c_uint32, C.c_uint32, u32p, rp)
    sig("yasph_step_host_slab_ex", C.c_int32, vp, f32p, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p, rp)
    _lib = L
    return L


def check(status, ctx=None):
    if status != 0:
        msg = lib().yasph_last_error(ctx)
        raise YasphError(status, msg.decode() if msg else "")


def default_config(smoothing_factor=2.0, particle_density=10000.0, fluid_density=100.0, solver=SOLVER_DFSPH):
    cfg = Config()
    check(lib().yasph_config_default(C.byref(cfg), smoothing_factor, particle_density, fluid_density, solver))
    return cfg
