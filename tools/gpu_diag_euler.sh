#!/bin/bash
# Where the time of the inviscid residual pass goes: the default library against the SDG_NSL_DIAG builds (wrong numbers, timing only)
# usage: tools/gpu_diag_euler.sh libdiag1.so libdiag2.so ...   (built with tools/build_variant.sh diagN "-DSDG_NSL_DIAG=N")
mkdir -p gpurun_out
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-ns-target > gpurun_out/bench_diag.json 2> gpurun_out/bench_diag.err; python -c "
import json;d=json.load(open('gpurun_out/bench_diag.json'));print('EULER lib=${lib:-default}', round(d['value'],2), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4))" | tee -a gpurun_out/diag_euler.txt; tail -2 gpurun_out/bench_diag.err
done
