#!/bin/bash
# per-kernel times of the default library and of variant libraries on both headline workloads (NS 64^3 launch list; Euler 96^3 launch list)
mkdir -p gpurun_out
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  for w in "--model ns --cells 64" "--cells 96 --no-ns-target"; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/l_ab.csv python bench.py $w --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
    python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/l_ab.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows: agg.setdefault(r[4][:34],[]).append(float(r[-1])/1e3)
print('lib=${lib:-default} [$w]', {k: [round(x,1) for x in v[-6:]] for k,v in agg.items() if 'nsl' in k and 'Trace' not in k})
PY
  done
done
