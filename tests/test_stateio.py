"""State files and trajectory logs (yasph2d_b200/stateio.py): host-side round trips, the recorder on an oracle run, and
-- on the GPU -- checkpoint / resume that continues bit-exactly."""
import numpy as np
import pytest

import yasph2d_b200 as y
from oracle import pyoracle as po
from yasph2d_b200 import stateio

capi = y.capi


def test_state_file_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    arrays = {"positions": rng.random((1001, 2), np.float32), "velocities": rng.random((1001, 2), np.float32),
              "boundary": rng.random((77, 2), np.float32), "kappa": rng.random(1001, np.float32), "empty": np.zeros((0, 2), np.float32)}
    params = {"smoothing_factor": 2.0, "particle_density": 10000.0, "fluid_density": 100.0}
    solver = {"kind": 0, "step_ns": 123456, "iters_density": 3, "iters_divergence": 2, "initialized": 1}
    path = tmp_path / "s.ysph"
    stateio.save_state(path, arrays, params, solver)
    a2, p2, s2 = stateio.load_state(path)
    assert p2 == params and s2 == solver and list(a2) == list(arrays)
    for k in arrays:
        assert a2[k].dtype == np.float32 and np.array_equal(a2[k], arrays[k]), k
    raw = open(path, "rb").read()
    assert raw[:8] == b"YSPH2D01" and len(raw) % 8 == 0
    with open(path, "wb") as f:
        f.write(b"garbage!" + raw[8:])
    with pytest.raises(ValueError):
        stateio.load_state(path)


def test_recorder_on_oracle_run(tmp_path):
    """The recorder takes the oracle's step reports as they are; two identical runs compare clean, a perturbed one does not."""
    def run(path, nudge=0.0):
        w = po.dam_break_scene(po.World())
        if nudge:  # a block of particles gets an initial sideways velocity
            v = w.velocities()
            v[:500, 0] += np.float32(nudge)
            w.set_particles(w.positions(), v)
        tm, s = po.TimeManager(cfl_factor=1.5), po.DFSPHSolver(w)
        rec = stateio.TrajectoryRecorder(path, header={"scene": "dam_break"}, particle_mass=0.01)
        for _ in range(40):
            rec.record(s.simulation_step(w, tm), w.velocities())
        rec.close()
        return rec.rows

    a = run(tmp_path / "a.jsonl")
    b = run(tmp_path / "b.jsonl")
    assert stateio.compare_trajectories(a, b, rel=0.0, iters_slack=0, abs_tol={"avg_density_error": 0.0}) == []
    hdr, rows = stateio.load_trajectory(tmp_path / "a.jsonl")
    assert hdr["header"]["scene"] == "dam_break" and rows == a and rows[-1]["step"] == 39
    assert rows[-1]["time_ns"] == sum(r["dt_ns"] for r in rows) and rows[-1]["kinetic_energy"] > 0.0
    c = run(None, nudge=0.05)
    assert stateio.compare_trajectories(a, c, rel=1e-7, iters_slack=0) != []
    assert stateio.compare_trajectories(a, c, rel=5e-2, iters_slack=1, abs_tol={"kinetic_energy": 0.01}) == []  # the north star's kind of bound


def _ctx(world, solver, **kw):
    cfg = capi.default_config(2.0, 10000.0, 100.0, solver)
    cfg.max_particles = len(world.particles.positions)
    cfg.max_boundary = len(world.particles.boundary_particles)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return y.GpuContext(cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("solver,knobs", [(capi.SOLVER_DFSPH, {}), (capi.SOLVER_DFSPH, dict(dfsph_max_avg_density_error=1e-6, dfsph_max_divergence_error=1e-5)),
                                          (capi.SOLVER_WCSPH, {})])
def test_checkpoint_resume_is_bit_exact(tmp_path, solver, knobs):
    """N steps in one go == a first part, checkpoint to a file, fresh context, resume, the rest (positions, velocities,
    densities, dt and iteration counts of every later step).  The tight-tolerance variant runs with active warm starts."""
    w = y.dam_break_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0))
    a = _ctx(w, solver, **knobs)
    a.set_boundary(w.particles.boundary_particles)
    a.upload_particles(w.particles.positions, w.particles.velocities)
    b = _ctx(w, solver, **knobs)
    b.set_boundary(w.particles.boundary_particles)
    b.upload_particles(w.particles.positions, w.particles.velocities)
    total, cut = (110, 90) if knobs else (40, 20)  # the tight-tolerance run checkpoints where the Jacobi loops iterate
    reps_a = [a.step() for _ in range(total)]
    for _ in range(cut):
        b.step()
    path = tmp_path / "ck.ysph"
    stateio.checkpoint(b, path, params={"scene": "dam_break"})
    b.close()
    c = _ctx(w, solver, **knobs)
    arrays, params, sol = stateio.resume(c, path)
    assert params == {"scene": "dam_break"} and sol["kind"] == solver
    warm = 0
    for s in range(cut, total):
        r = c.step()
        ra = reps_a[s]
        assert (r.dt_ns, r.iters_density, r.iters_divergence, r.warm_density, r.warm_divergence) == (
            ra.dt_ns, ra.iters_density, ra.iters_divergence, ra.warm_density, ra.warm_divergence), s
        assert r.avg_density_error == ra.avg_density_error and r.avg_divergence == ra.avg_divergence, s
        warm += r.warm_density + r.warm_divergence
    if knobs:
        assert warm > 0  # the resumed run did use the restored warm-start arrays
    pa, va, da = a.download_particles()
    pc, vc, dc = c.download_particles()
    assert np.array_equal(pa, pc) and np.array_equal(va, vc) and np.array_equal(da, dc)


@pytest.mark.gpu
def test_checkpoint_resume_with_target_frame_length(tmp_path):
    """TargetFrameLength stepping (timemanager.rs:268-274): the lower bound of dt depends on TimeManager::total_simulated_time, so
    the checkpoint carries it (yasph_solver_state.total_simulated_ns); the resumed run repeats the uninterrupted one's dt."""
    target = 1_000_000
    w = y.dam_break_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0))
    w.particles.velocities[:10, 0] = 300.0  # the CFL time drops below timestep_min: the lower bound (the target rule) decides dt

    def fresh():
        c = _ctx(w, capi.SOLVER_DFSPH, timestep_target_frame_ns=target)
        c.set_boundary(w.particles.boundary_particles)
        c.upload_particles(w.particles.positions, w.particles.velocities)
        return c

    a, b = fresh(), fresh()
    total, cut = 40, 17
    reps_a = [a.step() for _ in range(total)]
    assert min(r.dt_ns for r in reps_a[cut:]) < capi.default_config(2.0, 10000.0, 100.0, capi.SOLVER_DFSPH).timestep_min_ns  # the rule is active after the cut
    for _ in range(cut):
        b.step()
    assert b.solver_state().total_simulated_ns == sum(r.dt_prev_ns for r in reps_a[:cut])
    path = tmp_path / "ck_target.ysph"
    stateio.checkpoint(b, path)
    b.close()
    c = _ctx(w, capi.SOLVER_DFSPH, timestep_target_frame_ns=target)
    _, _, sol = stateio.resume(c, path)
    assert sol["total_simulated_ns"] > 0
    for s in range(cut, total):
        r = c.step()
        assert (r.dt_ns, r.iters_density, r.iters_divergence) == (reps_a[s].dt_ns, reps_a[s].iters_density, reps_a[s].iters_divergence), s
    assert c.total_simulated_ns() == a.total_simulated_ns()
    pa, va, _ = a.download_particles()
    pc, vc, _ = c.download_particles()
    assert np.array_equal(pa, pc) and np.array_equal(va, vc)
