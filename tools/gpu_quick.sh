#!/bin/bash
# quick GPU check: parity tests + both benches (no CPU legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_euler_q.json 2> gpurun_out/bench_euler_q.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_q.json'));print('EULER', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_euler_q.err
timeout 600 python bench.py --model ns --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ns_q.json 2> gpurun_out/bench_ns_q.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ns_q.json'));print('NS', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_ns_q.err
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_ns_q.csv python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns -s 8 -c 2 -o gpurun_out/prof_ns_q -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_full_ns_q.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eulerStage -s 7 -c 1 -o gpurun_out/prof_euler_q -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --cells 96 > gpurun_out/ncu_full_euler_q.log 2>&1
fi
