#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --model ns --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ns.json 2> gpurun_out/bench_ns.err; cat gpurun_out/bench_ns.json; tail -3 gpurun_out/bench_ns.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_ns.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_launch_ns.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns -s 8 -c 2 -o gpurun_out/prof_ns -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_full_ns.log 2>&1
ls -la gpurun_out
