#!/bin/bash
# round-2 ncu evidence: launch lists of the bench commands + one `--set full` capture per dominant kernel, raw CSV summaries
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_ns_64cube.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_euler_96cube.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-ns-target --cells 96 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsl -s 12 -c 6 -o gpurun_out/r02_prof_ns -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_full_ns.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nslStage -s 9 -c 3 -o gpurun_out/r02_prof_euler -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-ns-target --cells 96 > gpurun_out/ncu_full_euler.log 2>&1
for n in ns euler; do ncu -i gpurun_out/r02_prof_$n.ncu-rep --page raw --csv > gpurun_out/r02_prof_${n}_raw.csv 2>/dev/null; python tools/ncu_raw_summary.py gpurun_out/r02_prof_${n}_raw.csv > gpurun_out/r02_prof_${n}_summary.txt; done
cat gpurun_out/r02_prof_ns_summary.txt gpurun_out/r02_prof_euler_summary.txt | grep -E "^-----|gpu__time|dram__bytes|fp64|issue_active|stalls"
