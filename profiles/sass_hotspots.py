#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples for one kernel of `ncu --page source --csv --print-source sass` output."""
import csv
import sys


def kernels(path):
    out, cur = [], None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    return out


def main(path, pattern, top=25):
    best = {}
    for k in kernels(path):
        if pattern not in k["name"]:
            continue
        h = {n: i for i, n in enumerate(k["hdr"])}
        k["inst"] = sum(int(r[h["Instructions Executed"]]) for r in k["rows"])
        if k["name"] not in best or k["inst"] > best[k["name"]]["inst"]:
            best[k["name"]] = k  # the busiest instance (skips early-exit launches)
    for k in best.values():
        h = {n: i for i, n in enumerate(k["hdr"])}
        tot = sum(int(r[h["# Samples"]]) for r in k["rows"])
        inst = sum(int(r[h["Instructions Executed"]]) for r in k["rows"])
        print("=====", k["name"][:120])
        print("  sass lines %d, warp-instructions %d, samples %d" % (len(k["rows"]), inst, tot))
        stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        agg = {n: sum(int(r[h[n]]) for r in k["rows"]) for n in stall_cols}
        print("  stall mix:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(tot, 1)) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        rows = sorted(k["rows"], key=lambda r: -int(r[h["# Samples"]]))[:top]
        for r in rows:
            s = int(r[h["# Samples"]])
            main_stall = max(stall_cols, key=lambda n: int(r[h[n]]))
            print("  %5.1f%%  exec %9s  thr %5s  %-14s %s" % (100.0 * s / max(tot, 1), r[h["Instructions Executed"]], r[h["Avg. Threads Executed"]], main_stall[6:],
                                                            r[h["Source"]].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
