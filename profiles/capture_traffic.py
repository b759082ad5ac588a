#!/usr/bin/env python
"""DRAM traffic per launch of every kernel of the default bench workload, for bench.py's `roofline.traffic`.

Run ON THE GPU BOX (through gpurun) from the repository root:

    python profiles/capture_traffic.py [--solver dfsph] [--skip 4400] [--count 120]

It profiles a short `bench.py` run under ncu (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum; `--clock-control
none`), averages the launches of the steady-state steps per kernel and writes gpurun_out/traffic.json.  The file records a hash of the
CUDA sources it was taken from; bench.py uses profiles/r02/traffic.json only if that hash matches the sources it runs (a number taken
from other kernels is not a measurement of these).  Copy gpurun_out/traffic.json to profiles/r02/ and commit it.
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solver", default="dfsph")
    ap.add_argument("--skip", type=int, default=3600)
    ap.add_argument("--count", type=int, default=120)
    args = ap.parse_args()
    import bench

    out_csv = os.path.join(ROOT, "gpurun_out", "traffic_launches.csv")
    os.makedirs(os.path.dirname(out_csv), exist_ok=True)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none", "--launch-skip", str(args.skip),
           "-c", str(args.count), "--csv", "--log-file", out_csv, sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3",
           "--no-cpu-baseline", "--no-e2e", "--collapse-presteps", "0", "--solver", args.solver]
    r = subprocess.run(cmd, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    particles = json.loads(line[-1])["config"]["particles_per_gpu"] if line else None
    rows = list(csv.reader(open(out_csv)))
    hi = [i for i, row in enumerate(rows) if "Kernel Name" in row][0]
    col = {h: i for i, h in enumerate(rows[hi])}
    acc = {}
    for row in rows[hi + 2:]:
        if len(row) < len(col):
            continue
        name = re.sub(r"^void ", "", row[col["Kernel Name"]])
        name = re.sub(r"\(.*$", "", name)  # drop the argument list
        key, metric, unit = (row[col["ID"]], name), row[col["Metric Name"]], row[col["Metric Unit"]]
        v = float(row[col["Metric Value"]].replace(",", ""))
        if metric.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        elif metric.startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        acc.setdefault(key, {})[metric] = v
    kernels = {}
    for (_, name), m in acc.items():
        k = kernels.setdefault(name, {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "us": 0.0})
        k["launches"] += 1
        k["dram_bytes_read"] += m.get("dram__bytes_read.sum", 0.0)
        k["dram_bytes_write"] += m.get("dram__bytes_write.sum", 0.0)
        k["us"] += m.get("gpu__time_duration.sum", 0.0)
    for k in kernels.values():
        for f in ("dram_bytes_read", "dram_bytes_write", "us"):
            k[f] = round(k[f] / k["launches"], 1)
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip %d -c %d of `bench.py --steps 4 --warmup 3 "
                     "--no-cpu-baseline --no-e2e --collapse-presteps 0 --solver %s` (profiles/capture_traffic.py); per-launch averages" % (args.skip, args.count, args.solver),
           "kernel_source_hash": bench.kernel_source_hash(), "particles": particles, "solver": args.solver, "kernels": kernels}
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "traffic.json"), "w"), indent=1)
    for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["us"] * kv[1]["launches"]):
        print("%-60s %3d x %8.2f us  read %8.2f MB  write %8.2f MB" % (name[:60], k["launches"], k["us"], k["dram_bytes_read"] / 1e6, k["dram_bytes_write"] / 1e6))


if __name__ == "__main__":
    main()
