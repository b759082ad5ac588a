"""CPU-side checks: the C-ABI library loads, exports every symbol include/yasph_gpu.h declares, fails loudly without a
GPU, and the host-side scene builders / duration helpers agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import yasph2d_b200 as y
from oracle import pyoracle as po

capi = y.capi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "yasph_gpu.h")).read()
    declared = sorted(set(re.findall(r"^(?:int32_t|uint64_t|float|const char\*)\s+(yasph_[a-z0-9_]+)\(", hdr, re.M)))
    assert declared, "no declarations parsed"
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), "libyasph_gpu.so does not export %s" % name
    assert sorted(capi.EXPORTED_SYMBOLS) == declared


def test_struct_sizes_match_header():
    # the C structs are plain PODs with natural alignment; sizes are part of the ABI
    assert C.sizeof(capi.StepReport) == 80
    assert C.sizeof(capi.Config) == 152
    assert C.sizeof(capi.SolverState) == 32
    assert C.sizeof(capi.SlabInfo) == 72


def test_create_without_gpu_fails_loudly():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(capi.YasphError) as e:
        y.GpuContext(capi.default_config())
    assert e.value.status == 6 and "no CPU fallback" in str(e.value)


def test_default_config_matches_reference_literals():
    cfg = capi.default_config(2.0, 10000.0, 100.0, capi.SOLVER_DFSPH)
    assert np.float32(cfg.smoothing_length) == np.float32(0.02)
    assert (cfg.timestep_min_ns, cfg.timestep_max_ns) == (41667, 2777778)  # main.rs:123-124
    assert np.float32(cfg.cfl_factor) == np.float32(1.5) and cfg.dfsph_max_density_iters == 200 and cfg.dfsph_max_divergence_iters == 400
    assert np.float32(cfg.dfsph_max_avg_density_error) == np.float32(0.01) / np.float32(100.0)
    w = capi.default_config(2.0, 10000.0, 100.0, capi.SOLVER_WCSPH)
    assert np.float32(w.cfl_factor) == np.float32(0.2)
    ow = po.World()
    assert np.float32(w.wcsph_stiffness) == np.float32(po.WCSPHSolver(ow).stiffness())


def test_scene_builders_equal_oracle():
    w = y.dam_break_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0))
    ow = po.dam_break_scene(po.World())
    assert np.array_equal(w.particles.positions, ow.positions())
    assert np.array_equal(w.particles.boundary_particles, ow.boundary())
    w2 = y.FluidParticleWorld(2.0, 10000.0, 100.0)
    w2.add_fluid_rect(y.Rect(0.0, 0.0, 1.0, 1.0), 0.5)
    w2.add_fluid_rect(y.Rect(2.0, 0.0, 0.3, 0.2), 0.1)  # second rect: RNG seeded with the particle count (fpw.rs:153)
    w2.add_boundary_line((0.0, -0.1), (3.0, -0.2))
    ow2 = po.World()
    ow2.add_fluid_rect(0.0, 0.0, 1.0, 1.0, 0.5)
    ow2.add_fluid_rect(2.0, 0.0, 0.3, 0.2, 0.1)
    ow2.add_boundary_line((0.0, -0.1), (3.0, -0.2))
    assert np.array_equal(w2.particles.positions, ow2.positions()) and np.array_equal(w2.particles.boundary_particles, ow2.boundary())
    p = w.properties
    op = ow.props()
    assert (p.smoothing_length(), p.particle_mass(), p.particle_radius()) == (op["h"], op["mass"], op["radius"])


def test_tank_scene_counts():
    w = y.tank_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0), 300, 100)
    assert w.particles.num_dynamic_particles() == 30000
    assert w.particles.num_boundary_particles() > 1000
    # the oracle's own builder of the same tank (bench.py's reference arm builds its scene with it, without the product library)
    ow = po.tank_scene(po.World(), 300, 100)
    assert np.array_equal(w.particles.positions, ow.positions()) and np.array_equal(w.particles.boundary_particles, ow.boundary())


def test_duration_helpers_equal_oracle():
    L, O = capi.lib(), po.lib()
    for s in [0.0, 1e-9, 4.1667e-5, 1.0 / 360.0, 0.5, 1.75, 600.0, 1e-4 / 3.0]:
        assert L.yasph_duration_from_secs_f32(np.float32(s)) == O.yo_duration_from_secs_f32(np.float32(s))
    for ns in [0, 1, 41667, 2777778, 999999999, 1000000000, 3000000123]:
        assert L.yasph_duration_as_secs_f32(ns) == O.yo_duration_as_secs_f32(ns)
