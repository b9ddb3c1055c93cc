#!/bin/bash
# ncu evidence of the Euler residual pass at HEAD (relative-error transform by transposition): one --set full capture of three launches
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nslStage -s 9 -c 3 -o gpurun_out/r02_prof_euler_final -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-ns-target --cells 96 > gpurun_out/ncu_full_euler_final.log 2>&1
ncu -i gpurun_out/r02_prof_euler_final.ncu-rep --page raw --csv > gpurun_out/r02_prof_euler_final_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r02_prof_euler_final_raw.csv > gpurun_out/r02_prof_euler_final_summary.txt
grep -E "^-----|gpu__time|dram__bytes|fp64|issue_active|stalls|lsu_wavefronts|shared" gpurun_out/r02_prof_euler_final_summary.txt | head -60
