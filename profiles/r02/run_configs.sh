#!/bin/bash
# BASELINE.json configs on one GPU (results -> gpurun_out/r02_cfg*.json)
python bench.py --workload dam_break --solver wcsph --steps 200 --presteps 200 > gpurun_out/r02_cfg1_wcsph_dam.json 2> gpurun_out/r02_cfg1.err
python bench.py --workload dam_break --steps 200 --presteps 200 > gpurun_out/r02_cfg2_dfsph_dam.json 2> gpurun_out/r02_cfg2.err
python bench.py --columns-per-gpu 1000 --rows 1000 --cpu-columns 125 > gpurun_out/r02_cfg3_1m.json 2> gpurun_out/r02_cfg3.err
python bench.py --columns-per-gpu 16000 --no-cpu-baseline --collapse-presteps 0 > gpurun_out/r02_cfg4_16m_1gpu.json 2> gpurun_out/r02_cfg4.err
python bench.py --solver wcsph > gpurun_out/r02_wcsph_tank.json 2> gpurun_out/r02_wcsph_tank.err
tail -2 gpurun_out/r02_cfg*.err gpurun_out/r02_wcsph_tank.err
