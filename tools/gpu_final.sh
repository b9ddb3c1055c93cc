#!/bin/bash
# round-end evidence: parity tests, bench lines (Euler = the metric's config, NS = north_star target), reference arm, launch lists, ncu --set full
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_euler.json 2> gpurun_out/bench_euler.err; cat gpurun_out/bench_euler.json; tail -2 gpurun_out/bench_euler.err
timeout 900 python bench.py --model ns --steps 5 --warmup 3 > gpurun_out/bench_ns.json 2> gpurun_out/bench_ns.err; cat gpurun_out/bench_ns.json; tail -2 gpurun_out/bench_ns.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_euler.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_ns.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eulerLine -s 7 -c 1 -o gpurun_out/prof_euler -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_euler.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns -s 8 -c 2 -o gpurun_out/prof_ns -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_ns.log 2>&1
ls -la gpurun_out | head -30
