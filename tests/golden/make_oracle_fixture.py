"""Generates tests/golden/oracle_dam_break.json from the CPU oracle (oracle/yasph_oracle.cpp).

The reference is Rust and cannot be imported or built in this image, so this fixture does not come from the reference
itself: it freezes the ORACLE's output on the application's dam-break scene (main.rs:177-196) so that an accidental change
of the oracle (or of the scene builders it shares with the product path) shows up as a test failure on CPU.

    python tests/golden/make_oracle_fixture.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run(solver_kind, steps):
    w = po.dam_break_scene(po.World())
    out = {"n": int(w.n), "m": int(w.m), "scene_positions": digest(w.positions()), "scene_boundary": digest(w.boundary()), "steps": []}
    if solver_kind == "dfsph":
        tm, s = po.TimeManager(cfl_factor=1.5), po.DFSPHSolver(w)
    else:
        tm, s = po.TimeManager(cfl_factor=0.2), po.WCSPHSolver(w)
    for _ in range(steps):
        r = s.simulation_step(w, tm)
        out["steps"].append({"dt_ns": int(r.dt_ns), "iters_density": int(r.iters_density), "iters_divergence": int(r.iters_divergence),
                             "positions": digest(w.positions()), "velocities": digest(w.velocities()), "densities": digest(w.densities()),
                             "kinetic": float(0.5 * 0.01 * np.sum(w.velocities().astype(np.float64) ** 2)),
                             "mean_density": float(np.mean(w.densities().astype(np.float64)))})
    return out


if __name__ == "__main__":
    fx = {"_comment": "frozen oracle outputs; regenerate with tests/golden/make_oracle_fixture.py", "dfsph": run("dfsph", 5), "wcsph": run("wcsph", 5)}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_dam_break.json"), "w") as f:
        json.dump(fx, f, indent=1)
    print("ok")
