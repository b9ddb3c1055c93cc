#!/bin/bash
# whole GPU suite + the driver's bench invocation (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
/usr/bin/time -v timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-3000; grep -E "Elapsed|Maximum resident" gpurun_out/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
