"""Golden vectors: the reference's own known answers (tests/golden/reference_kats.json, values from the reference's unit
tests) against the oracle, and the frozen oracle outputs (tests/golden/oracle_dam_break.json) against the oracle (CPU) and
against the CUDA path through the C ABI (GPU)."""
import json
import os

import numpy as np
import pytest

import sys

from oracle import pyoracle as po

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, HERE)
from make_oracle_fixture import digest, run  # noqa: E402

KATS = json.load(open(os.path.join(HERE, "reference_kats.json")))
FIX = json.load(open(os.path.join(HERE, "oracle_dam_break.json")))


def test_morton_known_answers():  # morton.rs:189-251
    L = po.lib()
    for x, y, m in KATS["morton_encode"]["vectors"]:
        assert L.yo_morton_encode(x, y) == m and L.yo_morton_encode_lookup(x, y) == m
    for m, x, y in KATS["morton_decode"]["vectors"]:
        assert (L.yo_morton_decode_x(m), L.yo_morton_decode_y(m)) == (x, y)
    for cur, lo, hi, want in KATS["find_bigmin"]["vectors"]:
        assert L.yo_find_bigmin(cur, lo, hi) == want


@pytest.mark.parametrize("kind", ["dfsph", "wcsph"])
def test_oracle_matches_frozen_fixture(kind):
    got = run(kind, len(FIX[kind]["steps"]))
    assert got["n"] == FIX[kind]["n"] and got["m"] == FIX[kind]["m"]
    assert got["scene_positions"] == FIX[kind]["scene_positions"] and got["scene_boundary"] == FIX[kind]["scene_boundary"]
    for a, b in zip(got["steps"], FIX[kind]["steps"]):
        for k in ("dt_ns", "iters_density", "iters_divergence", "positions", "velocities", "densities"):
            assert a[k] == b[k], k


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dfsph", "wcsph"])
def test_gpu_matches_frozen_fixture(kind):
    """The CUDA path reproduces the frozen oracle trajectory of the application's dam-break scene: dt and iteration counts
    exactly; kinetic energy and mean density within the north-star tolerance (1e-4 relative); in the default strict
    arithmetic mode the per-particle arrays are bit-identical (digest equality)."""
    import yasph2d_b200 as y

    world = y.dam_break_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0))
    h = world.properties.smoothing_length()
    if kind == "dfsph":
        tm = y.TimeManager(y.SimulationStepConfig.AdaptiveTimeStep(cfl_factor=1.5))
        solver = y.DFSPHSolver(y.XSPHViscosityModel(h), h)
    else:
        tm = y.TimeManager(y.SimulationStepConfig.AdaptiveTimeStep(cfl_factor=0.2))
        solver = y.WCSPHSolver(y.XSPHViscosityModel(h), world.properties)
    assert digest(world.particles.positions) == FIX[kind]["scene_positions"]
    for want in FIX[kind]["steps"]:
        rep = solver.simulation_step(world, tm)
        assert rep.dt_ns == want["dt_ns"]
        if kind == "dfsph":
            assert rep.iters_density == want["iters_density"] and rep.iters_divergence == want["iters_divergence"]
        kin = float(0.5 * 0.01 * np.sum(world.particles.velocities.astype(np.float64) ** 2))
        assert abs(kin - want["kinetic"]) <= 1e-4 * abs(want["kinetic"])
        assert abs(float(np.mean(world.particles.densities.astype(np.float64))) - want["mean_density"]) <= 1e-4 * want["mean_density"]
        assert digest(world.particles.positions) == want["positions"]
        assert digest(world.particles.velocities) == want["velocities"]
        assert digest(world.particles.densities) == want["densities"]
