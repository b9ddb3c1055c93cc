#!/bin/bash
mkdir -p gpurun_out
for lib in librel3.so librel4.so; do
export SDG_LIB=$PWD/subrosadg_b200/$lib
for k in link trace; do
SDG_EULER_KERNEL=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_$k.json 2> gpurun_out/bench_euler_$k.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_$k.json'));print('$lib EULER $k', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_$k.err
done
done
unset SDG_LIB
bash tools/gpu_ab_nsl.sh librel3.so librel4.so
