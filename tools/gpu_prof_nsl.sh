#!/bin/bash
# ncu evidence for the trace-based NS line kernels: launch list of the bench command + one `--set full` capture of pass G and pass R
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_nsl.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsl -s 12 -c 3 -o gpurun_out/prof_nsl -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_full_nsl.log 2>&1
tail -2 gpurun_out/ncu_full_nsl.log
ncu -i gpurun_out/prof_nsl.ncu-rep --page raw --csv > gpurun_out/prof_nsl_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/prof_nsl_raw.csv
