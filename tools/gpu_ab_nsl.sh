#!/bin/bash
# per-kernel times (ncu launch list, 64^3 NS bench) of the default library and of variant libraries: tools/gpu_ab_nsl.sh [libA.so ...]
mkdir -p gpurun_out
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/l_ab.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/l_ab.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows: agg.setdefault(r[4][:34],[]).append(float(r[-1])/1e3)
print('lib=${lib:-default}', {k: [round(x,1) for x in v[-6:]] for k,v in agg.items() if 'nsl' in k})
PY
done
