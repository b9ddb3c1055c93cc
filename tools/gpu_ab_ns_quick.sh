#!/bin/bash
# timing-only A/B of NS library variants (no parity tests): tools/gpu_ab_ns_quick.sh libA.so libB.so ...
mkdir -p gpurun_out
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file gpurun_out/l_ab.csv python bench.py --model ns --steps 1 --warmup 3 --no-cpu --no-e2e --cells 64 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/l_ab.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows: agg.setdefault(r[4][:28],[]).append(float(r[-1])/1e3)
print('lib=${lib:-default}', {k: round(sum(v[-3:])/len(v[-3:]),1) for k,v in agg.items() if 'ns' in k})
PY
done
