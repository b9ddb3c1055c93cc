#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "periodic_3d_cns or sphere_3d or ns_target or two_contexts" 2>&1 | tail -2
for ah in 444 148 1000000; do
echo "== SDG_AHEAD=$ah"
SDG_AHEAD=$ah bash tools/gpu_ab_nsl.sh
SDG_AHEAD=$ah SDG_EULER_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_trace.json 2> gpurun_out/bench_euler_trace.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_trace.json'));print('EULER trace', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_trace.err
done
