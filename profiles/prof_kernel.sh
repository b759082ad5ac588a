#!/bin/bash
# ncu --set full of ONE kernel of the default bench workload, in steady state: prof_kernel.sh <kernel name regex> <tag> [launch-skip] [extra bench args]
# Writes gpurun_out/<tag>.ncu-rep, <tag>_raw.csv (--page raw) and <tag>_src.csv (--page source, SASS).
K="$1"; TAG="$2"; SKIP="${3:-230}"; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip "$SKIP" -c 1 -f -o "gpurun_out/$TAG" \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --collapse-presteps 0 "$@" > "gpurun_out/$TAG.log" 2>&1 || tail -5 "gpurun_out/$TAG.log"
ncu -i "gpurun_out/$TAG.ncu-rep" --page raw --csv > "gpurun_out/${TAG}_raw.csv" 2>/dev/null
ncu -i "gpurun_out/$TAG.ncu-rep" --page source --csv --print-source sass > "gpurun_out/${TAG}_src.csv" 2>/dev/null
ls -la gpurun_out/$TAG*
