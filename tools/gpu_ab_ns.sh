#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  timeout 600 python bench.py --model ns --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab.json'));print('NS lib=${lib:-default}', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_ab.err
done
