#!/bin/bash
# per-launch durations of a few steady-state steps: launch_list.sh <tag> [launch-skip] [count] [extra bench args]
TAG="$1"; SKIP="${2:-3600}"; CNT="${3:-160}"; shift 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip "$SKIP" -c "$CNT" --csv --log-file "gpurun_out/launches_$TAG.csv" \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --collapse-presteps 0 "$@" > "gpurun_out/launches_$TAG.log" 2>&1 || tail -5 "gpurun_out/launches_$TAG.log"
python - "$TAG" <<'PY'
import csv,collections,sys
rows=list(csv.reader(open('gpurun_out/launches_%s.csv'%sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; col={h:i for i,h in enumerate(hdr)}
acc={}; cnt={}
for r in rows[hi+2:]:
    if len(r)<len(hdr) or r[col['Metric Name']]!='gpu__time_duration.sum': continue
    n=r[col['Kernel Name']][:70]; v=float(r[col['Metric Value']].replace(',',''))
    acc[n]=acc.get(n,0)+v; cnt[n]=cnt.get(n,0)+1
tot=sum(acc.values())
for n,v in sorted(acc.items(), key=lambda kv:-kv[1]): print("%-72s %8.1f us %3d x %7.2f us %5.1f%%"%(n,v/1000,cnt[n],v/1000/cnt[n],100*v/tot))
PY
