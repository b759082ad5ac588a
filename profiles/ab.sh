#!/bin/bash
# A/B timing of library variants: ./scratch_ab.sh variants/a.so variants/b.so ...
for so in "$@"; do
  YASPH_GPU_LIB=$PWD/$so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --collapse-presteps 0 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
  python - "$so" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
p=d['roofline']['pass_us_per_step']
print("%-28s ms/step %.4f  " % (sys.argv[1], d['ms_per_step']) + " ".join("%s=%.0f"%(k[:9],v) for k,v in p.items() if v>1))
PY
done
