timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -15
for extra in "" "--no-peer-transport"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e $extra > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -20 gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=2 value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), 'peer', d['config']['slab_info_rank0']['peer_transport'])
print({k:v for k,v in d['roofline']['pass_us_per_step'].items() if v>0})
"
done
