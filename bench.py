#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the DFSPH hot path on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload tank|dam_break]

A "step" is one Solver::simulation_step (dfsph.rs:414-525) over the whole particle set.  Default workload: the DFSPH
dam-break tank of BASELINE.json configs[3] at 2 000 x 1 000 = 2 M fluid particles per GPU (8 GPUs = the 16 M-particle
tank), ONE tank slab-decomposed over the ranks along x (migration, ghost columns and per-pass halo exchange over NCCL, all-
reduced Jacobi residual and CFL maximum), advanced `--presteps` steps before anything is timed so the column is collapsing.
`value` is timed with the state resident in HBM; `e2e` goes through yasph_step_host with pinned HOST buffers
(upload pos+vel, step, download pos+vel+densities every step) -- what the reference's `simulation_step(&mut world, ..)`
does to the Vecs the Rust host owns.  `--impl reference` times the CPU restatement of the reference (oracle/, C++/OpenMP,
all host cores) on a bounded sample of the same workload; the Rust binary itself cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (DFSPH, 2D dam-break)"
UNIT = "particle-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tank", choices=["tank", "dam_break", "neighbors"])
    ap.add_argument("--neighbors-n", default="100000,1000000,4000000,16000000,64000000", help="particle counts of the config-5 sweep")
    ap.add_argument("--neighbors-radius", default="0.5,0.75,1.0,1.25", help="search radii of the config-5 sweep")
    ap.add_argument("--columns-per-gpu", type=int, default=2000)
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--presteps", type=int, default=200)
    ap.add_argument("--solver", default="dfsph", choices=["dfsph", "wcsph"], help="wcsph: BASELINE configs[0] (cfl 0.2, main.rs:115-118)")
    ap.add_argument("--total-columns", type=int, default=0, help="strong scaling: ONE tank of this many fluid columns split over the ranks")
    ap.add_argument("--collapse-presteps", type=int, default=1000, help="second timed regime: this many steps into the run (0 = skip)")
    ap.add_argument("--cpu-collapse-presteps", type=int, default=2000, help="presteps of the CPU sample of the second regime (0 = skip)")
    ap.add_argument("--cpu-columns", type=int, default=250, help="fluid columns of the bounded CPU sample (rows as the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="multi-GPU: skip the N-rank-vs-one-context equivalence check that precedes the timing")
    ap.add_argument("--no-peer-transport", action="store_true", help="multi-GPU: keep halo exchanges and all-reduces on NCCL (A/B of the peer-memory transport)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
# algorithmic bytes per particle and pass; K = mean total neighbours, L = 8 + 4K
# ----------------------------------------------------------------------------------------------------------------------
def pass_bytes(K):
    L = 8.0 + 4.0 * K
    return {
        "viscosity": 28 + L,
        "predict": 24,
        "density_warm": 36 + L,
        "density_iter": 64 + 2 * L,   # A (24+L) + B (40+L)
        "advect_keygen": 24,
        "neighborhood": 140 + 4 * K,  # key-gen, 4-pass radix sort, gathers, cells, list build
        "lists": 16 + 4 * K,
        "density_alpha": 40 + 8 * K,
        "divergence_warm": 36 + L,
        "divergence_iter": 60 + 2 * L,
    }


def clocks_sampler(stop, out, device_index, errors=None):
    """SM clock, power and throttle reasons sampled during the timed regions: NVML in-process (a sample per ~2 ms, so that even
    a timed region of a few milliseconds is covered), else the nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(device_index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [getattr(nv, "nvmlClocksThrottleReason" + n) for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap")]
        while True:
            r = reasons(h)
            try:
                watts = "%.2f" % (nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                watts = "0"
            out.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), watts] + [("Active" if (r & b) else "Not Active") for b in bits]
                       + [time.perf_counter()])
            if stop.wait(0.002):
                return
    except Exception as e:  # no NVML binding on this box: fall back to the command-line tool
        if errors is not None:
            errors.append("nvml: %r" % (e,))
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while True:
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 7:
                out.append(f[:7] + [time.perf_counter()])
        except Exception as e:
            if errors is not None:
                errors.append("nvidia-smi: %r" % (e,))
        if stop.wait(0.2):
            return


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    reasons = []
    for idx, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
        if any(s[idx].lower().startswith("active") for s in samples):
            reasons.append(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "power_w_max": max(float(s[2]) for s in samples),
            "samples": len(samples), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++/OpenMP restatement of the reference) on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------------------------------
def run_cpu(args, columns, steps, warmup, presteps=None):
    """The reference's CPU path (its C++/OpenMP restatement, oracle/) on all host cores.  The scene comes from the oracle's own
    restated builders (pyoracle.tank_scene / dam_break_scene): nothing of the product library is loaded by this arm."""
    from oracle import pyoracle as po

    cores = po.num_threads(os.cpu_count() or 1)
    presteps = args.presteps if presteps is None else presteps
    if args.workload == "tank":
        w = po.tank_scene(po.World(), columns, args.rows)
        sample = "tank %d x %d = %d fluid particles (+%d boundary), %s, %d presteps, %d timed steps" % (columns, args.rows, w.n, w.m, args.solver.upper(), presteps, steps)
    else:
        w = po.dam_break_scene(po.World())
        sample = "application dam-break scene (main.rs:177-196), %d fluid + %d boundary particles, %s, %d presteps, %d timed steps" % (w.n, w.m, args.solver.upper(), presteps, steps)
    if args.solver == "wcsph":
        tm, s = po.TimeManager(cfl_factor=0.2), po.WCSPHSolver(w)  # main.rs:115-118
    else:
        tm, s = po.TimeManager(cfl_factor=1.5), po.DFSPHSolver(w)
    for _ in range(presteps + warmup):
        s.simulation_step(w, tm)
    its, warm = [0, 0], [0, 0]
    t0 = time.perf_counter()
    for _ in range(steps):
        r = s.simulation_step(w, tm)
        its[0] += r.iters_density
        its[1] += r.iters_divergence
        warm[0] += r.warm_density
        warm[1] += r.warm_divergence
    dt = time.perf_counter() - t0
    return {"value": w.n * steps / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": 1e3 * dt / steps, "n": w.n, "iters_density_mean": its[0] / steps, "iters_divergence_mean": its[1] / steps,
            "warm_density_mean": warm[0] / steps, "warm_divergence_mean": warm[1] / steps}


def kernel_source_hash():
    """sha256 over the CUDA sources: ties a committed ncu capture (profiles/r02/traffic.json) to the kernels it was taken from."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "yasph2d_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def neighbor_sweep(args):
    """BASELINE.json configs[4]: neighbour search only, uniform points at density 10 / unit^2, seed
    123456789 (neighborhood_search.rs:531-538, benches/benchmarks/neighborhood_search.rs:9-29).  Warm = the points are already
    in sorted order (what the reference's bench measures after its first update), cold = shuffled.  GB/s are algorithmic bytes
    (B_ns = 140 + 4K per particle for the whole update, B_list = 16 + 4K for the list build) over device time."""
    import torch
    import yasph2d_b200 as y
    from oracle import pyoracle as po  # input generator only (the reference's SmallRng stream restated): not on the timed path

    capi = y.capi
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rows = []
    for n in [int(x) for x in args.neighbors_n.split(",")]:
        r = np.zeros(2 * n, np.float32)
        po.lib().yo_rng_fill(123456789, r.ctypes.data_as(po.C.POINTER(po.C.c_float)), 2 * n)
        pos = (r.reshape(n, 2) * np.float32(np.sqrt(np.float32(n) / np.float32(10.0)))).astype(np.float32)
        del r
        for radius in [float(x) for x in args.neighbors_radius.split(",")]:
            ns = y.NeighborhoodSearch(radius, max_particles=n, max_boundary=1)
            ctx = ns.ctx
            ctx.upload_particles(pos)
            ctx.neighborhood_update()  # cold-start allocation effects out of the way
            ctx.set_flags(capi.FLAG_PROFILE_PASSES)
            res = {"n": n, "radius": radius}
            for kind in ("cold", "warm"):
                acc, reps = {}, 3
                for _ in range(reps):
                    if kind == "cold":
                        ctx.upload_particles(pos)  # the generator's order: random in space
                    rep = ctx.neighborhood_update()
                    for k, v in ctx.pass_times_us().items():
                        acc[k] = acc.get(k, 0.0) + v / reps
                K = rep.total_neighbors / n
                t_all = acc["sort"] + acc["gather"] + acc["cells_tiles"] + acc["lists"]
                res.update({"mean_neighbors": round(K, 2), "capped": int(rep.neighbors_capped),
                            kind + "_us": round(t_all, 1), kind + "_lists_us": round(acc["lists"], 1),
                            kind + "_GBps": round((140 + 4 * K) * n / (t_all * 1e-6) / 1e9, 1),
                            kind + "_lists_GBps": round((16 + 4 * K) * n / (acc["lists"] * 1e-6) / 1e9, 1),
                            kind + "_Mparticles_per_s": round(n / t_all, 1)})
            res["warm_frac_of_hbm_peak"] = round(res["warm_GBps"] / peak, 4)
            rows.append(res)
            print("# " + json.dumps(res), file=sys.stderr, flush=True)
            ctx.close()
            del ns, ctx
            torch.cuda.empty_cache()
    print(json.dumps({"metric": "neighbour search only (BASELINE configs[4]): algorithmic GB/s of update_dynamic", "unit": "GB/s", "n_gpus": 1,
                      "peak": peak, "data": "synthetic uniform points, density 10, seed 123456789", "sweep": rows}))
    return 0


def bind_to_gpu_numa_node(device_index):
    """Pins this process to the CPU cores next to its GPU (the GPU's NUMA node) BEFORE any pinned host memory is allocated, so that the
    e2e arm's host buffers are first-touched on that node: with one process per GPU the uploads / downloads of all ranks then do not
    funnel through one socket's memory controllers.  Best effort: returns a note for the JSON line."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return "no NUMA information for %s" % bus
        os.sched_setaffinity(0, cpus)
        return "bound to NUMA node %d (%d cores) of GPU %s" % (node, len(cpus), bus)
    except Exception as e:  # noqa: BLE001
        return "not bound (%r)" % (e,)


def verify_slabs(args, dist, rank, world, local_rank):
    """N-rank equivalence on the hardware of this very run: a small tank stepped by the `world` slab contexts of
    this job (peer-memory transport, ghost layers, migration) against ONE context on rank 0.  Bars as in tests/test_gpu_slab.py: every
    particle is owned by exactly one rank at every step; bit-identical states, dt, iteration counts and residuals until the first
    particle migrates (same arithmetic, same neighbour order); afterwards only the summation order inside cells that received a
    migrant differs -- fields within 1e-4 relative for the next three steps, then global quantities (kinetic energy 2 %, mean density 0.1 %)."""
    import yasph2d_b200 as y
    from yasph2d_b200 import slab

    capi = y.capi
    steps, columns, rows = 50, 128 * world, 160
    hw = y.tank_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0), columns, rows)
    pos, vel, bnd = hw.particles.positions, hw.particles.velocities.copy(), hw.particles.boundary_particles
    vel[:, 0] = 0.4  # a drift towards +x: particles cross the slab boundaries within the first steps
    n = len(pos)
    cfg = capi.default_config(2.0, 10000.0, 100.0, capi.SOLVER_DFSPH if args.solver == "dfsph" else capi.SOLVER_WCSPH)
    cfg.device = local_rank
    cfg.max_particles, cfg.max_boundary = 2 * n // world + 65536, len(bnd)
    cfg.flags = capi.FLAG_PERMUTE_WARMSTART | capi.FLAG_TRACK_IDS | (capi.FLAG_NO_PEER_TRANSPORT if args.no_peer_transport else 0)
    uid = slab.broadcast_unique_id(dist)
    ctx, ranges, id_map = slab.make_slab_context(cfg, rank, world, uid, pos, vel, bnd)
    snaps, reps, migs = [], [], []
    for _ in range(steps):
        r = ctx.step()
        info = ctx.info()
        p, v, d = ctx.download_particles()
        snaps.append((id_map[ctx.ids()].astype(np.int64), p, v, d))
        reps.append((int(r.dt_ns), int(r.iters_density), int(r.iters_divergence), float(r.avg_density_error), float(r.avg_divergence)))
        migs.append(int(info.migrated_in + info.migrated_out_left + info.migrated_out_right))
    peer = int(ctx.info().peer_transport)
    ghost_cols = int(ctx.cfg.ghost_columns)
    halos = int(ctx.info().halo_exchanges)
    ctx.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((snaps, reps, migs), gathered, dst=0)
    if rank != 0:
        return None
    cfg1 = capi.default_config(2.0, 10000.0, 100.0, cfg.solver)
    cfg1.device = local_rank
    cfg1.max_particles, cfg1.max_boundary = n, len(bnd)
    cfg1.flags = capi.FLAG_PERMUTE_WARMSTART | capi.FLAG_TRACK_IDS
    one = y.GpuContext(cfg1)
    one.set_boundary(bnd)
    one.upload_particles(pos, vel)
    first_mig = next((s for s in range(steps) if any(g[2][s] for g in gathered)), steps)
    res = {"ok": True, "steps": steps, "particles": n, "ranks": world, "first_migration_step": first_mig, "bit_identical_steps": 0, "max_rel_dev_after_migration": 0.0,
           "peer_transport": peer, "ghost_columns": ghost_cols, "halo_exchanges_rank0": halos, "problems": []}
    for s in range(steps):
        r1 = one.step()
        ids1 = one.field(capi.FIELD_ID).astype(np.int64)
        p1, v1, d1 = one.download_particles()
        ref = {"pos": np.empty_like(p1), "vel": np.empty_like(v1), "dens": np.empty_like(d1)}
        ref["pos"][ids1], ref["vel"][ids1], ref["dens"][ids1] = p1, v1, d1
        got = {"pos": np.full_like(p1, np.nan), "vel": np.full_like(v1, np.nan), "dens": np.full_like(d1, np.nan)}
        seen = np.zeros(n, np.int32)
        for g in gathered:
            ids, p, v, d = g[0][s]
            np.add.at(seen, ids, 1)
            got["pos"][ids], got["vel"][ids], got["dens"][ids] = p, v, d
        if not (seen == 1).all():
            res["problems"].append("step %d: %d particles unowned, %d owned twice" % (s, int((seen == 0).sum()), int((seen > 1).sum())))
            break
        rep1 = (int(r1.dt_ns), int(r1.iters_density), int(r1.iters_divergence), float(r1.avg_density_error), float(r1.avg_divergence))
        if len({g[1][s] for g in gathered}) != 1:
            res["problems"].append("step %d: the ranks disagree on dt / iteration counts / residuals" % s)
        if s < first_mig:
            same = all(np.array_equal(got[k], ref[k]) for k in ref) and gathered[0][1][s] == rep1
            if same:
                res["bit_identical_steps"] += 1
            else:
                res["problems"].append("step %d: differs from the single context before any migration" % s)
        else:
            dev = 0.0
            for k in ("vel", "dens"):
                a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
                dev = max(dev, float((np.abs(a - b) / np.maximum(np.abs(b), 0.1 * np.abs(b).mean() + 1e-30)).max()))
            if s < first_mig + 3:
                res["max_rel_dev_after_migration"] = max(res["max_rel_dev_after_migration"], dev)
                if dev > 1e-4:
                    res["problems"].append("step %d: relative deviation %.2e > 1e-4 within three steps of the first migration" % (s, dev))
            else:
                ea, eb = float((got["vel"].astype(np.float64) ** 2).sum()), float((ref["vel"].astype(np.float64) ** 2).sum())
                if abs(ea - eb) > 0.02 * max(eb, 1e-12) or abs(float(got["dens"].mean()) - float(ref["dens"].mean())) > 1e-3 * float(ref["dens"].mean()):
                    res["problems"].append("step %d: kinetic energy / mean density left the bounds" % s)
            if abs(gathered[0][1][s][0] - rep1[0]) > 2e-3 * rep1[0]:
                res["problems"].append("step %d: dt %d vs %d" % (s, gathered[0][1][s][0], rep1[0]))
    one.close()
    res["ok"] = not res["problems"]
    res["problems"] = res["problems"][:5]
    return res


def pass_groups(solver, pt, B, K, it_rho, it_div, w_rho, w_div):
    """(us per step, algorithmic bytes per particle and step) of every timed pass group.  The density / alpha sweep also carries
    iteration 0's density-change pass of the divergence solver whenever that solver's warm start does not run (OpDensityAlphaDiv):
    its bytes (B_A,div = 20 + L) are credited to the group that spends the time."""
    L = 8.0 + 4.0 * K
    if solver == "wcsph":
        return {
            "kickdrift_keygen": (pt["advect_keygen"], 40.0),
            "density": (pt["density_alpha"], 12 + L),
            "wcsph_accel": (pt["wcsph_accel"], 28 + L),
            "cfl_kick": (pt["wcsph_kick"], 40.0),
            "lists": (pt["lists"], B["lists"]),
            "neighborhood(sort+gather+cells+lists)": (pt["sort"] + pt["gather"] + pt["cells_tiles"] + pt["lists"], B["neighborhood"]),
        }
    fused = 1.0 - w_div  # fraction of steps whose iteration-0 pass A ran inside the density / alpha sweep
    return {
        "viscosity": (pt["viscosity"], B["viscosity"]),
        # one GPU: the solve's last Jacobi B also advects and generates the sort keys (OpJacobiBAdvect) -- no separate pass then, its
        # bytes (24: position + predicted velocity in, position out) are credited to the group that spends the time
        "density_solve": (pt["density_solve"], B["density_iter"] * it_rho + (B["advect_keygen"] if pt["advect_keygen"] < 1.0 else 0.0)),
        "density_warm": (pt["density_warm"], B["density_warm"] * w_rho),
        "divergence_solve": (pt["divergence_solve"], B["divergence_iter"] * it_div - fused * (20 + L)),
        "divergence_warm": (pt["divergence_warm"], B["divergence_warm"] * w_div),
        "density_alpha": (pt["density_alpha"], B["density_alpha"] + fused * (20 + L)),
        "density_alpha+divergence_solve": (pt["density_alpha"] + pt["divergence_solve"], B["density_alpha"] + B["divergence_iter"] * it_div),
        "lists": (pt["lists"], B["lists"]),
        "neighborhood(sort+gather+cells+lists)": (pt["sort"] + pt["gather"] + pt["cells_tiles"] + pt["lists"], B["neighborhood"]),
    }


PASS_KERNELS = {  # the kernel(s) behind each sweep group, as ncu names them
    "viscosity": ["k_sweep<OpViscosity>"],
    "density_solve": ["k_sweep<OpJacobiA<0>>", "k_sweep<OpJacobiB<0, 0>>", "k_sweep<OpJacobiBAdvect>"],
    "divergence_solve": ["k_sweep<OpJacobiB<1, 0>>"],
    "density_alpha": ["k_sweep<OpDensityAlphaDiv>"],
    "lists": ["k_build_lists"],
    "density": ["k_sweep<OpDensityAlpha<1, 0, 1>>"],
    "wcsph_accel": ["k_sweep<OpWcsphAccel>"],
}


def main():
    args = parse()
    if args.workload == "neighbors":
        return neighbor_sweep(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = METRIC if args.solver == "dfsph" else "particle-steps/sec (WCSPH, 2D dam-break)"
    total_columns = args.total_columns if args.total_columns else args.columns_per_gpu * world
    scaling = "strong" if args.total_columns else "weak"

    if args.impl == "reference":
        if rank != 0:
            return 0
        # N = 1: exactly the GPU arm's workload (tank, presteps, warm-up and step count).  N > 1: one GPU's share of the tank -- the
        # CPU's particle-steps/s barely depend on the tank's width, and the full 16 M tank would take ~1.5 s per step here.
        columns = args.columns_per_gpu if not args.total_columns else max(1, args.total_columns // world)
        cb = run_cpu(args, columns, args.steps, args.warmup)
        same = world == 1
        line = {
            "impl": "reference", "metric": metric, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, CPU restatement of the reference (C++/OpenMP, oracle/) -- not the Rust binary (no cargo in this image)" % workload_name(args, columns, 1),
                       "sample": cb["sample"], "same_config_as_gpu_arm": same, "presteps": args.presteps,
                       "iters_density": cb["iters_density_mean"], "iters_divergence": cb["iters_divergence_mean"],
                       "warm_density": cb["warm_density_mean"], "warm_divergence": cb["warm_divergence_mean"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import yasph2d_b200 as y

    capi = y.capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: not bound"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    verify = verify_slabs(args, dist, rank, world, local_rank) if (world > 1 and not args.no_verify) else None

    # ---- scene (host): every rank builds the whole scene (the jitter stream is sequential), then keeps its slab ----
    hw = y.FluidParticleWorld(2.0, 10000.0, 100.0)
    if args.workload == "tank":
        y.tank_scene(hw, total_columns, args.rows)
    else:
        y.dam_break_scene(hw)
    workload = workload_name(args, total_columns // world if args.workload == "tank" else 0, world)
    n_global, m = hw.particles.num_dynamic_particles(), hw.particles.num_boundary_particles()

    solver_kind = capi.SOLVER_DFSPH if args.solver == "dfsph" else capi.SOLVER_WCSPH
    cfg = capi.default_config(2.0, 10000.0, 100.0, solver_kind)
    cfg.device = local_rank
    if args.no_peer_transport:
        cfg.flags |= capi.FLAG_NO_PEER_TRANSPORT
    if world == 1:
        n = n_global
        cfg.max_particles, cfg.max_boundary = n, m
        ctx = y.GpuContext(cfg)
        ctx.set_boundary(hw.particles.boundary_particles)
        ctx.upload_particles(hw.particles.positions, hw.particles.velocities)
        ranges = None
    else:
        from yasph2d_b200 import slab

        # 1-D slab decomposition over cell columns: migration + ghost columns + per-pass halo exchange, all-reduced residual
        cfg.max_particles, cfg.max_boundary = int(n_global / world * 1.3) + 65536, m
        uid = slab.broadcast_unique_id(dist)
        ctx, ranges, _ = slab.make_slab_context(cfg, rank, world, uid, hw.particles.positions, hw.particles.velocities, hw.particles.boundary_particles,
                                                track_ids=False)
        n = ctx.counts()[0]
    del hw
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", local_rank))

    # clocks sampler: started before the pre-steps (NVML initialisation takes longer than a short timed region); only the
    # samples stamped inside the timed region are reported
    stop, samples, clock_errors, window = threading.Event(), [], [], {}
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank, clock_errors), daemon=True)
    th.start()
    steps_done = [0]

    def dev_step():
        steps_done[0] += 1
        return ctx.step()

    for _ in range(args.presteps):
        dev_step()

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = []
        l0 = ctx.launch_count()
        window.setdefault("t0", time.perf_counter())  # the first timed region (device-resident steps) is the one the clocks belong to
        e0.record(stream)
        if step_fn is dev_step and not os.environ.get("YASPH_BENCH_SINGLE_STEPS"):  # device-resident steps: ONE call for the whole timed region (yasph_step_n, the application's frame loop)
            steps_done[0] += steps
            reps = ctx.step_n(steps)
        else:
            for _ in range(steps):
                reps.append(step_fn())
        e1.record(stream)
        torch.cuda.synchronize()
        window.setdefault("t1", time.perf_counter())
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count() - l0
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, reps, launches

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic_file = None
    try:
        traffic_file = json.load(open(os.path.join(ROOT, "profiles", "r02", "traffic.json")))
    except Exception:
        pass

    def total_particles():
        if world == 1:
            return ctx.counts()[0]
        t = torch.tensor([ctx.counts()[0]], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    def measure(label):
        """Device-resident throughput of `--steps` steps, then per-pass device times (CUDA event pairs around every pass, recorded
        on the library's launching stream) of a few more steps -> the roofline of every pass group."""
        ms, reps, launches = timed(dev_step, args.steps, args.warmup)
        n_total = total_particles()
        assert n_total == n_global, (n_total, n_global)
        n_loc = ctx.counts()[0]
        res = {"regime": label, "first_timed_step": steps_done[0] - args.steps, "value": n_total * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
               "gpu_launches": int(launches)}
        it_rho = float(np.mean([r.iters_density for r in reps]))
        it_div = float(np.mean([r.iters_divergence for r in reps]))
        w_rho = float(np.mean([r.warm_density for r in reps]))
        w_div = float(np.mean([r.warm_divergence for r in reps]))
        K = float(np.mean([r.total_neighbors for r in reps])) / max(n_loc, 1)
        ctx.set_flags(cfg.flags | capi.FLAG_PROFILE_PASSES)
        acc, psteps, pit = {}, max(3, min(args.steps, 10)), [0.0, 0.0, 0.0, 0.0]
        for _ in range(psteps):
            r = dev_step()
            for q, v in enumerate((r.iters_density, r.iters_divergence, r.warm_density, r.warm_divergence)):
                pit[q] += v / psteps
            for k, v in ctx.pass_times_us().items():
                acc[k] = acc.get(k, 0.0) + v
        ctx.set_flags(cfg.flags)
        pt = {k: v / psteps for k, v in acc.items()}  # us per step
        B = pass_bytes(K)
        groups = pass_groups(args.solver, pt, B, K, *pit)
        passes = {}
        for k, (us, bpp) in groups.items():
            gbs = (bpp * n_loc) / (us * 1e-6) / 1e9 if us > 0 else 0.0
            passes[k] = {"us_per_step": round(us, 2), "alg_bytes_per_particle": round(bpp, 1), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
        cands = [k for k in groups if k in PASS_KERNELS]
        dominant = max(cands, key=lambda k: groups[k][0])
        d_us, d_bpp = groups[dominant]
        achieved = (d_bpp * n_loc) / (d_us * 1e-6) / 1e9
        if args.solver == "wcsph":
            step_bytes = 276 + 12 * K
        else:
            step_bytes = (B["viscosity"] + B["predict"] + it_rho * B["density_iter"] + w_rho * B["density_warm"] + 4 + B["advect_keygen"] + B["neighborhood"]
                          + B["density_alpha"] + it_div * B["divergence_iter"] + w_div * B["divergence_warm"] + 4)
        # DRAM traffic ncu counted per launch of the dominant group's kernels: from the committed capture of THESE kernel sources
        traffic, traffic_src = None, None
        if traffic_file and world == 1 and n_loc == traffic_file.get("particles") and args.solver == traffic_file.get("solver", "dfsph"):
            if traffic_file.get("kernel_source_hash") == kernel_source_hash():
                def entry(kname):  # ncu prints template arguments its own way (k_build_lists<0>): match by prefix
                    hits = [v for k, v in traffic_file["kernels"].items() if k == kname or k.startswith(kname + "<")]
                    if not hits:
                        raise KeyError(kname)
                    return max(hits, key=lambda v: v["us"])

                try:
                    traffic = sum(entry(k)["dram_bytes_read"] + entry(k)["dram_bytes_write"] for k in PASS_KERNELS[dominant])
                    traffic_src = traffic_file.get("source")
                except KeyError:
                    traffic_src = "profiles/r02/traffic.json has no entry for %s" % PASS_KERNELS[dominant]
            else:
                traffic_src = "profiles/r02/traffic.json was captured from other kernel sources (hash %s, now %s): not used" % (
                    traffic_file.get("kernel_source_hash"), kernel_source_hash())
        res.update({"mean_neighbors": round(K, 2), "iters_density": it_rho, "iters_divergence": it_div, "warm_density": w_rho, "warm_divergence": w_div})
        res["roofline"] = {
            "bound": "hbm", "kernel": "%s (%s)" % (dominant, " + ".join(PASS_KERNELS[dominant])),
            "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "frac_of_nominal_8000_GBps": round(achieved / 8000.0, 4), "traffic": traffic,
            "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
            "alg_bytes_per_launch": round(d_bpp * n_loc, 0), "peak_source": peak_kind,
            "whole_step": {"alg_bytes_per_particle_step": round(step_bytes, 1), "GBps": round(step_bytes * n_total * args.steps / (ms * 1e-3) / 1e9, 1),
                           "frac_per_gpu": round(step_bytes * n_total / world * args.steps / (ms * 1e-3) / 1e9 / peak, 4)},
            "passes": passes, "pass_us_per_step": {k: round(v, 2) for k, v in pt.items()},
        }
        return res

    # ---- regime "early": `--presteps` steps into the run (the headline: both arms run exactly this) ----
    early = measure("early")
    stop.set()
    th.join(timeout=3)
    in_window = [s_ for s_ in samples if window["t0"] <= s_[7] <= window["t1"]]
    clocks = summarize_clocks(in_window if in_window else samples[-1:])
    clocks["window"] = "timed region" if in_window else "nearest sample (the timed region is shorter than one sampling period)"
    if clock_errors:
        clocks["sampler_notes"] = clock_errors[:2]
    n_total = n_global

    # ---- end to end through the reference-facing call with pinned HOST buffers ----
    e2e = None
    if not args.no_e2e:
        cap = int(cfg.max_particles)
        pos_t = torch.empty((cap, 2), dtype=torch.float32, pin_memory=True)
        vel_t = torch.empty((cap, 2), dtype=torch.float32, pin_memory=True)
        den_t = torch.empty((cap,), dtype=torch.float32, pin_memory=True)
        pos, vel, den = pos_t.numpy(), vel_t.numpy(), den_t.numpy()
        p0, v0, _ = ctx.download_particles()
        n_cur = [len(p0)]
        pos[: n_cur[0]], vel[: n_cur[0]] = p0, v0
        moved = [0, 0]

        unchanged = [False]

        def host_step():
            steps_done[0] += 1
            if world == 1:
                return ctx.step_host(pos[: n_cur[0]], vel[: n_cur[0]], den[: n_cur[0]], input_unchanged=unchanged[0])
            moved[0] += 0 if unchanged[0] else n_cur[0] * 16
            rep, n_out = ctx.step_host_slab(pos, vel, den, n_cur[0], input_unchanged=unchanged[0])
            n_cur[0] = n_out
            moved[1] += n_out * 20
            return rep

        ems, ereps, _ = timed(host_step, args.steps, args.warmup)
        timeline = None
        if world == 1:  # where the call's time goes: device timeline of a few more calls (events on the library's streams; the event
            # records themselves queue behind the link traffic, so this run is slower than the timed one -- read it as an order)
            ctx.set_flags(cfg.flags | capi.FLAG_PROFILE_PASSES)
            tl = [dict(host_step() and ctx.host_step_times_us()) for _ in range(5)][2:]
            ctx.set_flags(cfg.flags)
            timeline = {k: round(float(np.mean([t[k] for t in tl])), 1) for k in tl[0]}
        # the same call for a host that only READS the particle arrays between steps (the reference's application, main.rs:242-258):
        # yasph_step_host_ex(YASPH_HOST_INPUT_UNCHANGED) skips the upload; reported beside the headline, never instead of it
        unchanged[0] = True
        ums, _, _ = timed(host_step, args.steps, args.warmup)
        unchanged[0] = False
        no_upload = {"value": n_total * args.steps / (ums * 1e-3), "ms_per_step": ums / args.steps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(n_total * 20),
                     "api": "yasph_step_host%s_ex(YASPH_HOST_INPUT_UNCHANGED)" % ("" if world == 1 else "_slab")}
        e2e = {"value": n_total * args.steps / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n_total * 16), "d2h_bytes_per_step": int(n_total * 20),
               "ms_per_step": ems / args.steps, "device_timeline_us": timeline, "input_unchanged": no_upload, "host_affinity": numa_note,
               "api": "yasph_step_host%s (upload pos+vel, simulation_step, download pos+vel+densities; bytes summed over ranks)" % ("" if world == 1 else "_slab")}

    # ---- regime "collapse": further into the run, where the divergence solver iterates and warm-starts ----
    collapse = None
    if args.collapse_presteps and args.solver == "dfsph" and args.workload == "tank":
        while steps_done[0] < args.collapse_presteps:
            dev_step()
        collapse = measure("collapse")

    # ---- CPU baseline (rank 0, single-GPU run only): bounded samples of the same tank, one per regime ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        keys = ("value", "unit", "cores", "kind", "sample", "ms_per_step", "iters_density_mean", "iters_divergence_mean", "warm_divergence_mean")
        c1 = run_cpu(args, args.cpu_columns, max(3, min(args.steps, 10)), 1)
        cpu = {k: c1[k] for k in keys}
        if collapse is not None and args.cpu_collapse_presteps:
            # the narrow sample tank reaches the iterating regime later than the wide one (its waves reflect sooner): the sample is
            # matched by solver work (iterations, warm start), not by step number
            c2 = run_cpu(args, args.cpu_columns, max(3, min(args.steps, 10)), 1, presteps=args.cpu_collapse_presteps)
            cpu["collapse"] = {k: c2[k] for k in keys}

    if rank == 0:
        line = {
            "metric": metric, "value": early["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": early["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload, "solver": args.solver, "regime": "early", "particles_per_gpu": n, "particles_total": n_global, "boundary_particles": m,
                "presteps": args.presteps,
                "parallelism": "1 GPU" if world == 1 else "1-D slab decomposition over cell columns, %d ranks, peer-memory (NVLink) migration / ghost / halo exchange + all-reduced residual" % world,
                "slab_ranges": ranges, "slab_info_rank0": (ctx.info().as_dict() if world > 1 else None),
                "l2": "working set %.0f MB per GPU > 126 MB L2 (no explicit flush)" % (n * 240 / 1e6) if n * 240 > 126e6 else "working set fits L2 (small scene)",
                "mean_neighbors": early["mean_neighbors"], "iters_density": early["iters_density"], "iters_divergence": early["iters_divergence"],
                "warm_density": early["warm_density"], "warm_divergence": early["warm_divergence"],
                "arithmetic": "strict f32, no FMA contraction, IEEE div/sqrt (bit-exact vs oracle)",
                "api": ("yasph_step (one call per step)" if os.environ.get("YASPH_BENCH_SINGLE_STEPS") else
                        "yasph_step_n: the timed steps in ONE call (the application's frame loop, main.rs:339-360); per-step reports as from yasph_step"),
            },
            "gpu_launches": early["gpu_launches"], "clocks": clocks, "roofline": early["roofline"],
        }
        if collapse is not None:
            line["regimes"] = {"early": {k: early[k] for k in ("first_timed_step", "value", "ms_per_step", "iters_density", "iters_divergence", "warm_density", "warm_divergence", "mean_neighbors")},
                               "collapse": collapse}
        if verify is not None:
            line["verify"] = verify
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def workload_name(args, columns_per_gpu, world):
    if args.workload != "tank":
        return "%s application dam-break scene (BASELINE configs[0]/[1], main.rs:177-196)" % args.solver.upper()
    return "%s dam-break tank (BASELINE configs[3]): %d x %d fluid particles per GPU, %d in total" % (
        args.solver.upper(), columns_per_gpu, args.rows, columns_per_gpu * world * args.rows)


if __name__ == "__main__":
    sys.exit(main())
