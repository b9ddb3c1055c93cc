#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list, ncu --set full of the stage kernels. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_euler.json 2> gpurun_out/bench_euler.err; cat gpurun_out/bench_euler.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_euler.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --cells 96 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eulerStageKernel -s 6 -c 3 -o gpurun_out/prof_euler -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --cells 96 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
