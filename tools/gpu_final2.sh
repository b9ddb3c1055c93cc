#!/bin/bash
# round-2 closing evidence on one B200: full GPU suite, smoke, the default bench line exactly as the driver runs it, the reference arm,
# launch lists of the same bench commands (per-launch times under ncu are cold-cache and serialised: shares only)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/bench_default.err; cat gpurun_out/r02_bench_default_1gpu.json; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err; cat gpurun_out/r02_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_default_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
tail -3 gpurun_out/r02_launches_default_bench.csv
