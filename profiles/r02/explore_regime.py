"""Where does the 2000 x 1000 tank of bench.py enter the regime in which the Jacobi solvers iterate?  Logs per-step reports."""
import json, sys, time
import numpy as np
import yasph2d_b200 as y
capi = y.capi
cols, rows, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hw = y.tank_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0), cols, rows)
cfg = capi.default_config(2.0, 10000.0, 100.0, capi.SOLVER_DFSPH)
cfg.max_particles, cfg.max_boundary = len(hw.particles.positions), len(hw.particles.boundary_particles)
ctx = y.GpuContext(cfg)
ctx.set_boundary(hw.particles.boundary_particles)
ctx.upload_particles(hw.particles.positions, hw.particles.velocities)
rows_ = []
t0 = time.time()
for s in range(steps):
    r = ctx.step()
    if s % 25 == 0 or r.iters_density > 1 and s % 5 == 0:
        rows_.append((s, r.dt_ns, r.iters_density, r.iters_divergence, r.warm_density, r.warm_divergence, round(r.total_neighbors / cfg.max_particles, 2), r.neighbors_capped))
print("wall", time.time() - t0)
for r in rows_:
    print(*r)
