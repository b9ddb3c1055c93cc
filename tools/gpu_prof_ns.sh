#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_ns_q.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns -s 8 -c 2 -o gpurun_out/prof_ns_q -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_full_ns_q.log 2>&1
tail -2 gpurun_out/ncu_full_ns_q.log
