#!/bin/bash
# quick GPU check: parity tests then the device-resident bench line (per-pass times)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --collapse-presteps 0 > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick.json'))
print("value %.4g  ms/step %.4f" % (d['value'], d['ms_per_step']))
print({k:v for k,v in d['roofline']['pass_us_per_step'].items() if v>0})
print({k:v['frac'] for k,v in d['roofline']['passes'].items()})
PY
