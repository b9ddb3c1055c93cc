#!/bin/bash
mkdir -p gpurun_out
for lib in "" libeul4.so; do
if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
SDG_EULER_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_trace.json 2> gpurun_out/bench_euler_trace.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_trace.json'));print('EULER trace lib=$lib', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_trace.err
done
