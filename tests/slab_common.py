"""Helpers shared by the slab (multi-GPU) tests: run a scene on `world` slabs and collect per-id state."""
import threading

import numpy as np

import yasph2d_b200 as y
from yasph2d_b200 import slab

capi = y.capi


def scene_arrays(kind="dam"):
    w = y.FluidParticleWorld(2.0, 10000.0, 100.0)
    if kind == "dam":
        y.dam_break_scene(w)
    else:  # a wide shallow pool: many columns, splashes towards +x
        w.add_fluid_rect(y.Rect(0.1, 0.1, 1.2, 0.4), 0.05)
        w.add_boundary_thick_line((0.0, 0.0), (3.0, 0.0), 4)
        w.add_boundary_thick_line((0.0, 0.0), (0.0, 1.5), 4)
        w.add_boundary_thick_line((3.0, 0.0), (3.0, 1.5), 4)
    return w.particles.positions.copy(), w.particles.velocities.copy(), w.particles.boundary_particles.copy()


def base_config(n, m, solver=capi.SOLVER_DFSPH, **kw):
    cfg = capi.default_config(2.0, 10000.0, 100.0, solver)
    cfg.max_particles, cfg.max_boundary = n, max(m, 1)
    cfg.flags = capi.FLAG_PERMUTE_WARMSTART | capi.FLAG_TRACK_IDS
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def snapshot(ctx, id_map=None):
    """State of the owned particles keyed by global particle index."""
    pos, vel, dens = ctx.download_particles()
    ids = ctx.field(capi.FIELD_ID) if not hasattr(ctx, "ids") else ctx.ids()
    gid = ids if id_map is None else id_map[ids]
    return {"id": gid.astype(np.int64), "pos": pos, "vel": vel, "dens": dens}


def merge(snaps, n):
    out = {"pos": np.full((n, 2), np.nan, np.float32), "vel": np.full((n, 2), np.nan, np.float32), "dens": np.full(n, np.nan, np.float32)}
    seen = np.zeros(n, np.int32)
    for s in snaps:
        np.add.at(seen, s["id"], 1)
        for k in ("pos", "vel", "dens"):
            out[k][s["id"]] = s[k]
    assert (seen == 1).all(), "every particle must be owned by exactly one rank: %d missing, %d duplicated" % ((seen == 0).sum(), (seen > 1).sum())
    return out


def run_single(pos, vel, boundary, steps, checkpoints, solver=capi.SOLVER_DFSPH, **kw):
    cfg = base_config(len(pos), len(boundary), solver, **kw)
    ctx = y.GpuContext(cfg)
    ctx.set_boundary(boundary)
    ctx.upload_particles(pos, vel)
    reps, snaps = [], {}
    for s in range(steps):
        reps.append(ctx.step().as_dict())
        if s in checkpoints:
            snaps[s] = merge([snapshot(ctx)], len(pos))
    return reps, snaps, ctx


def neighbor_sets_global(ctx, id_map=None):
    """{global id of an owned particle: (frozenset of dynamic neighbour global ids, tuple of static neighbours)}."""
    if isinstance(ctx, slab.SlabContext):
        cd, ct, lists = ctx.local_neighbors()
        ids = ctx.local_field(capi.FIELD_ID, np.uint32)
        ghost = ctx.local_field(capi.FIELD_GHOST, np.uint8)
    else:
        cd, ct, lists = ctx.neighbors()
        ids = ctx.field(capi.FIELD_ID)
        ghost = np.zeros(len(ids), np.uint8)
    gid = ids.astype(np.int64) if id_map is None else id_map[ids].astype(np.int64)
    out = {}
    for i in np.nonzero(ghost == 0)[0]:
        out[int(gid[i])] = (frozenset(gid[lists[i, : cd[i]]].tolist()), int(ct[i]) - int(cd[i]))
    return out


def run_slabs_loopback(world, pos, vel, boundary, steps, checkpoints, solver=capi.SOLVER_DFSPH, neighbor_step=None, ranges=None, frame=None, **kw):
    """`world` slabs of one scene on cuda:0, one thread per rank, loopback transport.  frame: step through yasph_step_n in frames of
    that many steps (checkpoints must then fall on the last step of a frame)."""
    fabric = slab.LoopbackFabric(world)
    cfg = base_config(2 * len(pos), len(boundary), solver, **kw)  # room for the ghost layers
    results = [None] * world
    errors = []

    def worker(rank):
        try:
            ctx, rngs, id_map = slab.make_slab_context(cfg, rank, world, fabric, pos, vel, boundary, ranges)
            reps, snaps, infos, nsets = [], {}, [], None
            for s in range(steps):
                if frame is None:
                    reps.append(ctx.step().as_dict())
                elif s % frame == 0:
                    reps.extend(r.as_dict() for r in ctx.step_n(min(frame, steps - s)))
                if frame is not None and (s + 1) % frame != 0 and s != steps - 1:
                    assert s not in checkpoints
                    infos.append(None)
                    continue
                infos.append(ctx.info().as_dict())
                if s in checkpoints:
                    snaps[s] = snapshot(ctx, id_map)
                if neighbor_step is not None and s == neighbor_step:
                    nsets = neighbor_sets_global(ctx, id_map)
            results[rank] = {"reps": reps, "snaps": snaps, "infos": infos, "ranges": rngs, "nsets": nsets}
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "slab ranks hung (errors so far: %s)" % errors
    assert not errors, errors
    fabric.close()
    merged = {s: merge([results[r]["snaps"][s] for r in range(world)], len(pos)) for s in checkpoints if s < steps}
    return results, merged
