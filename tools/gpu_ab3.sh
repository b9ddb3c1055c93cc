#!/bin/bash
# A/B of variant libraries on the Euler 128^3 headline (same box): parity of the P3-hexahedron cases first, then bench numbers
# usage: tools/gpu_ab3.sh libX.so ...   (built with tools/build_variant.sh)
mkdir -p gpurun_out
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_step_host.py tests/test_reference_sweeps.py -m gpu -q -x -k "32cube or p3_hex or streamed or hex" 2>&1 | tail -2 | tee -a gpurun_out/ab3.txt
  for rep in 1 2; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-ns-target > gpurun_out/bench_ab3.json 2> gpurun_out/bench_ab3.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab3.json'));print('lib=${lib:-default} EULER', round(d['value'],2), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4))" | tee -a gpurun_out/ab3.txt; tail -2 gpurun_out/bench_ab3.err
  done
done
