"""Host-side slab logic under a real world_size-2 process group (gloo, CPU only; no GPU, no NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_host_check.py

Every rank builds the same scene, partitions it with yasph2d_b200.slab exactly as bench.py / make_slab_context do, and the
ranks then check over the process group that they agree: identical column ranges, a disjoint cover of the particles, balanced
slabs, id bases that concatenate, edge columns (the ghost sets each side expects from the other) that match pairwise, and a
byte blob broadcast the way the NCCL unique id travels.  Rank 0 prints one JSON line; non-zero exit on any disagreement.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    import yasph2d_b200 as y
    from yasph2d_b200 import slab

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    hw = y.tank_scene(y.FluidParticleWorld(2.0, 10000.0, 100.0), 300 * world, 120)
    pos = hw.particles.positions
    h = float(hw.properties.smoothing_length())
    ranges, own = slab.scatter_scene(pos, h, -100.0, world)
    cols = slab.cell_columns(pos[:, 0], h, -100.0)
    lo, hi = ranges[rank]
    mine = own[rank]
    # what this rank would upload, and what it expects as ghosts from each side / sends to each side (one cell column)
    report = {
        "rank": rank, "ranges": [list(r) for r in ranges], "n_own": int(len(mine)), "id_base": int(sum(len(o) for o in own[:rank])),
        "checksum": int(mine.astype(np.int64).sum()),
        "send_left": int(np.count_nonzero(cols[mine] == lo)) if rank > 0 else 0,
        "send_right": int(np.count_nonzero(cols[mine] == hi - 1)) if rank + 1 < world else 0,
        "expect_from_left": int(np.count_nonzero(cols == lo - 1)) if rank > 0 else 0,
        "expect_from_right": int(np.count_nonzero(cols == hi)) if rank + 1 < world else 0,
    }
    # the C ABI's column function must agree with the numpy restatement the partitioning uses
    cfg = y.capi.default_config(2.0, 10000.0, 100.0, y.capi.SOLVER_DFSPH)
    import ctypes as C

    col = C.c_uint32(0)
    for x in pos[mine[:: max(1, len(mine) // 50)], 0]:
        y.capi.check(y.capi.lib().yasph_cell_column(C.byref(cfg), C.c_float(float(x)), C.byref(col)))
        assert lo <= col.value < hi, (float(x), col.value, lo, hi)
    # a blob broadcast like the NCCL unique id (slab.broadcast_unique_id uses the same call)
    blob = [os.urandom(y.capi.COMM_ID_BYTES) if rank == 0 else None]
    dist.broadcast_object_list(blob, src=0)
    report["blob"] = blob[0].hex()
    gathered = [None] * world
    dist.all_gather_object(gathered, report)
    ok, why = True, "ok"
    try:
        assert all(g["ranges"] == gathered[0]["ranges"] for g in gathered), "ranks disagree on the column ranges"
        assert all(g["blob"] == gathered[0]["blob"] for g in gathered), "broadcast blob differs"
        assert gathered[0]["ranges"][0][0] == 0 and gathered[0]["ranges"][-1][1] == 65536
        for a, b in zip(gathered[0]["ranges"][:-1], gathered[0]["ranges"][1:]):
            assert a[1] == b[0], "ranges are not adjacent"
        assert sum(g["n_own"] for g in gathered) == len(pos), "the slabs do not cover the scene"
        assert sum(g["checksum"] for g in gathered) == len(pos) * (len(pos) - 1) // 2, "a particle is owned twice or not at all"
        for r, g in enumerate(gathered):
            assert g["id_base"] == sum(q["n_own"] for q in gathered[:r])
            assert abs(g["n_own"] - len(pos) / world) < 0.1 * len(pos) / world, "unbalanced slabs"
            if r + 1 < world:
                assert g["send_right"] == gathered[r + 1]["expect_from_left"] and gathered[r + 1]["send_left"] == g["expect_from_right"], "ghost columns disagree"
                assert g["send_right"] > 0 and gathered[r + 1]["send_left"] > 0
    except AssertionError as e:
        ok, why = False, str(e)
    if rank == 0:
        print(json.dumps({"check": "slab_host", "world": world, "result": why, "ranges": gathered[0]["ranges"], "n_own": [g["n_own"] for g in gathered]}))
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
