"""TimeManager step logic (src/sph/timemanager.rs:104-138,252-279) -- the reference has no tests for it; these
pin the restated integer-nanosecond semantics."""
import numpy as np

from oracle import pyoracle as po

f32 = np.float32


def test_duration_roundtrip():
    L = po.lib()
    assert L.yo_duration_from_secs_f32(f32(1.0) / f32(60.0) / f32(400.0)) == 41667  # main.rs:124
    assert L.yo_duration_from_secs_f32(f32(1.0) / f32(120.0) / f32(3.0)) == 2777778  # main.rs:123
    assert L.yo_duration_as_secs_f32(41667) == f32(41667) / f32(1e9)
    assert L.yo_duration_as_secs_f32(3_000_000_123) == f32(3.0) + f32(123) / f32(1e9)


def test_adaptive_step_rule():
    tm = po.TimeManager(cfl_factor=1.5)
    assert tm.simulation_step_ns() == tm.min_ns  # initial step = timestep_min (timemanager.rs:106-109)
    d = f32(0.01)
    # at rest: CFL time is huge -> limited by 2 * previous (timemanager.rs:267)
    assert tm.update_simulation_step(d, f32(0.0)) == 2 * 41667
    assert tm.update_simulation_step(d, f32(0.0)) == 4 * 41667
    for _ in range(10):
        tm.update_simulation_step(d, f32(0.0))
    assert tm.simulation_step_ns() == tm.max_ns
    # fast: cfl = 1.5*0.4*0.01/(v+1e-5)
    v = f32(8.0)
    expect = po.lib().yo_duration_from_secs_f32(f32(1.5) * f32(0.4) * d / (v + f32(0.00001)))
    assert tm.update_simulation_step(d, v) == expect
    # very fast: clamped to timestep_min
    assert tm.update_simulation_step(d, f32(1e6)) == tm.min_ns


def test_fixed_step():
    tm = po.TimeManager(adaptive=False, fixed_ns=250000)
    assert tm.simulation_step_ns() == 250000
    assert tm.update_simulation_step(f32(0.01), f32(123.0)) == 250000


def test_target_frame_length_rule():
    """AdaptiveTimeStepTarget::TargetFrameLength (timemanager.rs:23-36, 268-274): the lower bound becomes
    min(timestep_min, total_simulated_time - target * floor(total / target))."""
    target = 1_000_000
    tm = po.TimeManager(cfl_factor=1.5, target_frame_ns=target)
    d = f32(0.01)
    # total = 0: time_to_target = 0 -> lower bound 0 -> a very fast particle can push the step below timestep_min
    expect = po.lib().yo_duration_from_secs_f32(f32(1.5) * f32(0.4) * d / (f32(1e6) + f32(0.00001)))
    assert expect < tm.min_ns
    assert tm.update_simulation_step(d, f32(1e6)) == expect
    # total just past a multiple of the target by less than timestep_min: that remainder is the lower bound
    tm2 = po.TimeManager(cfl_factor=1.5, target_frame_ns=target)
    tm2.set_step_ns(target + 1234)
    tm2.perform_step()
    assert tm2.total_simulated_ns() == target + 1234
    assert tm2.update_simulation_step(d, f32(1e9)) == 1234
    # remainder above timestep_min: timestep_min rules, as without a target
    tm3 = po.TimeManager(cfl_factor=1.5, target_frame_ns=target)
    tm3.set_step_ns(target + 500_000)
    tm3.perform_step()
    assert tm3.update_simulation_step(d, f32(1e9)) == tm3.min_ns
    # without a target nothing changes
    tm4 = po.TimeManager(cfl_factor=1.5)
    tm4.perform_step()
    assert tm4.update_simulation_step(d, f32(1e9)) == tm4.min_ns
