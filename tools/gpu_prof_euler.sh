#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:euler -s 7 -c 1 -o gpurun_out/prof_euler_q -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --cells 96 > gpurun_out/ncu_full_euler_q.log 2>&1
tail -3 gpurun_out/ncu_full_euler_q.log
