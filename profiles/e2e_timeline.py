"""Wall time and device timeline of yasph_step_host on the bench workload (2 M-particle tank, pinned host arrays), plus the per-pass
times inside the call next to those of a device-resident step.  Run on a GPU box: python profiles/e2e_timeline.py"""
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import yasph2d_b200 as y
capi = y.capi
w = y.tank_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0), 2000, 1000)
cfg = capi.default_config(2.0, 10000.0, 100.0, capi.SOLVER_DFSPH)
cfg.max_particles = len(w.particles.positions); cfg.max_boundary = len(w.particles.boundary_particles)
ctx = y.GpuContext(cfg); ctx.set_boundary(w.particles.boundary_particles); ctx.upload_particles(w.particles.positions, w.particles.velocities)
for _ in range(100): ctx.step()
n = cfg.max_particles
pos_t = torch.empty((n, 2), dtype=torch.float32, pin_memory=True); vel_t = torch.empty((n, 2), dtype=torch.float32, pin_memory=True); den_t = torch.empty((n,), dtype=torch.float32, pin_memory=True)
pos, vel, den = pos_t.numpy(), vel_t.numpy(), den_t.numpy()
p0, v0, _ = ctx.download_particles(); pos[:] = p0; vel[:] = v0
import time
for _ in range(5): ctx.step_host(pos, vel, den)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): ctx.step_host(pos, vel, den)
ms = (time.perf_counter() - t) / 20 * 1e3
ctx.set_flags(capi.FLAG_PROFILE_PASSES)
tl = []
for _ in range(5):
    ctx.step_host(pos, vel, den); tl.append(ctx.host_step_times_us())
print("ms/call %.3f" % ms, {k: round(float(np.mean([t[k] for t in tl[2:]]))) for k in tl[0]})
pt = ctx.pass_times_us()
print("e2e passes", {k: round(v) for k, v in pt.items() if v > 0.5})
ctx.step(); print("resident passes", {k: round(v) for k, v in ctx.pass_times_us().items() if v > 0.5})
