#!/bin/bash
# A/B of variant libraries on both headline workloads (Euler 128^3 and the NS 96^3 target), bench numbers only
# usage: tools/gpu_ab2.sh libX.so ...   (built with tools/build_variant.sh)
mkdir -p gpurun_out
for lib in "" $@; do
  if [ -n "$lib" ]; then export SDG_LIB=$PWD/subrosadg_b200/$lib; else unset SDG_LIB; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ab2.json 2> gpurun_out/bench_ab2.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab2.json'));print('lib=${lib:-default} EULER', round(d['value'],2), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), 'NS', round(d['ns_target']['value'],2), round(d['ns_target']['ms_per_stage'],3), round(d['ns_target']['roofline']['frac'],4))" | tee -a gpurun_out/ab2.txt; tail -2 gpurun_out/bench_ab2.err
done
