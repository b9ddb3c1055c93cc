#!/bin/bash
# Euler through the trace-based residual pass (SDG_EULER_TRACE=1) vs eulerLineKernel: parity subset + bench at 128^3
mkdir -p gpurun_out
SDG_EULER_TRACE=1 timeout 600 python -m pytest tests -m gpu -q -k "periodic_3d_ceuler or curved_p3_hexahedra or affine_p3_hexahedra or weak_eos or config4 or travelling or two_contexts" 2>&1 | tail -5
SDG_EULER_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_trace.json 2> gpurun_out/bench_euler_trace.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_trace.json'));print('EULER trace', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_trace.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_line.json 2> gpurun_out/bench_euler_line.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_line.json'));print('EULER line', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_line.err
