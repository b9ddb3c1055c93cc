#!/bin/bash
# round-end evidence: parity tests, bench lines (Euler = the metric's config, NS = north_star target), reference arm, launch lists, hybrid side bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_euler.json 2> gpurun_out/bench_euler.err; cat gpurun_out/bench_euler.json; tail -2 gpurun_out/bench_euler.err
timeout 900 python bench.py --model ns --steps 5 --warmup 3 > gpurun_out/bench_ns.json 2> gpurun_out/bench_ns.err; cat gpurun_out/bench_ns.json; tail -2 gpurun_out/bench_ns.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 300 python tools/bench_hybrid.py 4.0 20 > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err; cat gpurun_out/bench_hybrid.json; tail -2 gpurun_out/bench_hybrid.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_euler.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_ns.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_hybrid.csv python tools/bench_hybrid.py 4.0 2 > /dev/null 2>&1
ls -la gpurun_out | head -30
