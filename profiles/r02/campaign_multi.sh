#!/bin/bash
# usage: campaign_multi.sh N  -- weak (2 M per GPU) and strong (16 M in total) scaling lines at N GPUs
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_r02_weak_n$N.json 2> gpurun_out/bench_r02_weak_n$N.err
if [ "$N" != "8" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --total-columns 16000 --collapse-presteps 0 > gpurun_out/bench_r02_strong_n$N.json 2> gpurun_out/bench_r02_strong_n$N.err
fi
tail -n 2 gpurun_out/bench_r02_*_n$N.err
