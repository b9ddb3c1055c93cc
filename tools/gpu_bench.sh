#!/bin/bash
# the driver's bench invocation (both arms), timed
mkdir -p gpurun_out
t0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "ours: rc=$? $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/bench_default.json | cut -c1-3500; tail -3 gpurun_out/bench_default.err
t0=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference: rc=$? $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/bench_reference.json | cut -c1-1500
