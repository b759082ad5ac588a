"""Multi-GPU equivalence check, one process per GPU (launch with torchrun / torch.distributed.run, world >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/slab_nccl_check.py

Every rank owns one slab of the dam-break scene and steps it through the NCCL transport of libyasph_gpu.so; rank 0 also
runs the same scene on a single context.  Bars as in tests/test_gpu_slab.py (which runs the same comparison through the
in-process loopback transport): bit-identical until the first migration, 1e-4 relative right after, global quantities later.
Prints one JSON line on rank 0 and exits non-zero on failure.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import yasph2d_b200 as y
    from slab_common import base_config, merge, run_single, scene_arrays, snapshot
    from test_gpu_slab import compare_runs, first_migration
    from yasph2d_b200 import slab

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # host channel only: the NCCL communicator lives inside libyasph_gpu.so
    steps = int(os.environ.get("SLAB_STEPS", "120"))
    pos, vel, boundary = scene_arrays("dam")
    uid = slab.broadcast_unique_id(dist)
    cfg = base_config(2 * len(pos), len(boundary))  # room for the ghost layers
    cfg.device = local
    if os.environ.get("SLAB_TRANSPORT", "peer") == "nccl":
        cfg.flags |= y.capi.FLAG_NO_PEER_TRANSPORT
    ctx, ranges, id_map = slab.make_slab_context(cfg, rank, world, uid, pos, vel, boundary)
    reps, snaps, infos = [], {}, []
    for s in range(steps):
        reps.append(ctx.step().as_dict())
        infos.append(ctx.info().as_dict())
        snaps[s] = snapshot(ctx, id_map)
    mine = {"reps": reps, "snaps": snaps, "infos": infos, "ranges": ranges}
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        reps1, snaps1, _ = run_single(pos, vel, boundary, steps, range(steps))
        merged = {s: merge([g["snaps"][s] for g in gathered], len(pos)) for s in range(steps)}
        fm = first_migration(gathered, steps)
        try:
            compare_runs(snaps1, merged, reps1, gathered, steps, fm)
            msg = "ok"
        except AssertionError as e:
            ok, msg = False, str(e)[:400]
        print(json.dumps({"check": "slab_nccl", "world": world, "steps": steps, "first_migration": fm, "ranges": ranges, "result": msg,
                          "n_own_final": [g["infos"][-1]["n_own"] for g in gathered],
                          "peer_transport": min(g["infos"][-1]["peer_transport"] for g in gathered),
                          "halo_exchanges": gathered[0]["infos"][-1]["halo_exchanges"], "allreduces": gathered[0]["infos"][-1]["allreduces"]}))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    ctx.close()
    dist.destroy_process_group()
    return 0 if flag[0] else 1


if __name__ == "__main__":
    sys.exit(main())
