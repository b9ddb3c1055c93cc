#!/bin/bash
# ncu evidence of the Navier-Stokes kernel pair at HEAD: one --set full capture of six launches (gradient + residual pass of three stages) at 64^3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsl -s 12 -c 6 -o gpurun_out/r02_prof_ns_final -f python bench.py --model ns --steps 2 --warmup 3 --no-cpu --no-e2e --cells 64 > gpurun_out/ncu_full_ns_final.log 2>&1
ncu -i gpurun_out/r02_prof_ns_final.ncu-rep --page raw --csv > gpurun_out/r02_prof_ns_final_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r02_prof_ns_final_raw.csv > gpurun_out/r02_prof_ns_final_summary.txt
grep -E "^-----|gpu__time|dram__bytes|fp64|issue_active|stalls|wavefronts_mem_shared" gpurun_out/r02_prof_ns_final_summary.txt | head -60
