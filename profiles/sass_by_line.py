#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of `ncu --page source --csv --print-source sass` to CUDA source lines.

ncu cannot show the CUDA source here (the profile was taken on another box), so the line table comes from
`nvdisasm -g` of the cubin inside the .so that was profiled: both listings enumerate a kernel's instructions in the same
order, which is the join key.

    python profiles/sass_by_line.py <source.csv> <libyasph_gpu.so> <kernel substring (demangled)> <mangled substring> [top]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_hotspots as sh  # noqa: E402


def line_table(so, mangled):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, cur, inside = [], None, False
    for ln in txt:
        if ln.startswith("//---------------------"):
            inside = (".text." in ln) and (mangled in ln)
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            out.append(cur)
    return out


def main(csv_path, so, pattern, mangled, top=40):
    best = None
    for k in sh.kernels(csv_path):
        if pattern not in k["name"]:
            continue
        h = {n: i for i, n in enumerate(k["hdr"])}
        inst = sum(int(r[h["Instructions Executed"]]) for r in k["rows"])
        if best is None or inst > best[0]:
            best = (inst, k, h)
    inst, k, h = best
    lt = line_table(so, mangled)
    rows = k["rows"]
    print("kernel %s: %d sass rows, %d lines from nvdisasm, %d warp instructions" % (k["name"][:80], len(rows), len(lt), inst))
    n = min(len(rows), len(lt))
    agg = collections.defaultdict(lambda: [0, 0])
    for i in range(n):
        key = lt[i] or ("?", 0)
        agg[key][0] += int(rows[i][h["Instructions Executed"]])
        agg[key][1] += int(rows[i][h["# Samples"]])
    tot_s = sum(v[1] for v in agg.values())
    src_cache = {}
    for (f, l), (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        for base in ("yasph2d_b200/csrc",):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), base, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < l <= len(src_cache[p]):
                    text = src_cache[p][l - 1].strip()[:100]
        print("  %5.1f%% inst  %5.1f%% samples  %s:%d  %s" % (100.0 * c / inst, 100.0 * s / max(tot_s, 1), f, l, text))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 40)
