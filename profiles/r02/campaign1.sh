#!/bin/bash
# round-2 single-GPU evidence: final bench line, reference arm, launch list, ncu --set full of the list build and the sweeps, DRAM traffic
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
bash profiles/launch_list.sh r02_final 3600 200 > gpurun_out/launches_r02_final_summary.txt 2>&1
python profiles/capture_traffic.py > gpurun_out/traffic_r02_summary.txt 2>&1
bash profiles/prof_kernel.sh k_build_lists ncu_r02_lists 200 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_sweep" --launch-skip 1000 -c 5 -f -o gpurun_out/ncu_r02_sweeps \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --collapse-presteps 0 > gpurun_out/ncu_r02_sweeps.log 2>&1
ncu -i gpurun_out/ncu_r02_sweeps.ncu-rep --page raw --csv > gpurun_out/ncu_r02_sweeps_raw.csv 2>/dev/null
ls -la gpurun_out/*r02*
