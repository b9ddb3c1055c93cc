#!/bin/bash
mkdir -p gpurun_out
SDG_EULER_KERNEL=link timeout 600 python -m pytest tests -m gpu -q -k "periodic_3d_ceuler or curved_p3_hexahedra or affine_p3_hexahedra or weak_eos or config4 or travelling or two_contexts or conservative" 2>&1 | tail -5
for k in link line trace; do
SDG_EULER_KERNEL=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_$k.json 2> gpurun_out/bench_euler_$k.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_$k.json'));print('EULER $k', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_$k.err
done
