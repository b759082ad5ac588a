for so in yasph2d_b200/libyasph_gpu.so variants/skippad.so; do
  for solver in dfsph wcsph; do
  YASPH_GPU_LIB=$PWD/$so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --collapse-presteps 0 --solver $solver > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
  python - "$so $solver" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
p=d['roofline']['pass_us_per_step']
print("%-40s ms/step %.4f  " % (sys.argv[1], d['ms_per_step']) + " ".join("%s=%.0f"%(k[:9],v) for k,v in p.items() if v>1))
PY
  done
done
