#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "periodic_3d_cns or sphere_3d or ns_target" 2>&1 | tail -2
bash tools/gpu_ab_nsl.sh
SDG_EULER_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_trace.json 2> gpurun_out/bench_euler_trace.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_trace.json'));print('EULER trace', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_euler_trace.err
SDG_EULER_TRACE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nslStage -s 8 -c 2 -o gpurun_out/prof_eulertrace -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --cells 96 > gpurun_out/ncu_full_et.log 2>&1
ncu -i gpurun_out/prof_eulertrace.ncu-rep --page raw --csv > gpurun_out/prof_et_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/prof_et_raw.csv
