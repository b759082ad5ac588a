"""The C ABI from a plain C program (tests/cabi_driver.c): the header compiles as C99, the library links without Python, and a
run through yasph_step_host from C gives exactly what the ctypes mirror gives."""
import os
import subprocess

import numpy as np
import pytest

import yasph2d_b200 as y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "yasph2d_b200")
FNV0, FNVP, MASK = 14695981039346656037, 1099511628211, (1 << 64) - 1


def fnv1a(arrays):
    h = FNV0
    for a in arrays:
        for b in np.ascontiguousarray(a).tobytes():
            h = ((h ^ b) * FNVP) & MASK
    return h


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    y.capi.lib()  # built?
    exe = str(tmp_path_factory.mktemp("cabi") / "cabi_driver")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_driver.c"),
           "-L" + LIBDIR, "-lyasph_gpu", "-Wl,-rpath," + LIBDIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_header_is_c99_and_scene_builders_work_without_a_device(driver):
    out = subprocess.run([driver, "scene"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    w = y.dam_break_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0))
    want = "scene fluid=%d boundary=%d hash=%016x" % (len(w.particles.positions), len(w.particles.boundary_particles),
                                                      fnv1a([w.particles.positions, w.particles.boundary_particles]))
    assert out.stdout.strip() == want


@pytest.mark.gpu
def test_c_host_equals_ctypes_mirror(driver):
    steps = 25
    out = subprocess.run([driver, "run", str(steps)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.stdout[-500:], out.stderr[-500:])
    lines = out.stdout.strip().splitlines()
    w = y.dam_break_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0))
    tm = y.TimeManager(y.SimulationStepConfig.AdaptiveTimeStep(cfl_factor=1.5))
    solver = y.DFSPHSolver(y.XSPHViscosityModel(w.properties.smoothing_length()), w.properties.smoothing_length())
    want = []
    for s in range(steps):
        rep = solver.simulation_step(w, tm)
        want.append("step %d dt_ns=%d iters=%d/%d" % (s, rep.dt_ns, rep.iters_density, rep.iters_divergence))
    assert lines[1 : 1 + steps] == want
    assert lines[-1] == "state hash=%016x" % fnv1a([w.particles.positions, w.particles.velocities, w.particles.densities])
