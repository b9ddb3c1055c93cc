#!/bin/bash
# upload-group sweep of sdg_step_host at 128^3 (the default is n / 16384 clamped to 2..64)
mkdir -p gpurun_out
for g in 48 96 128 192; do
  export SDG_HOST_PIPE_GROUPS=$g
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-ns-target > gpurun_out/step_host_bench_$g.json 2> gpurun_out/step_host_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/step_host_bench_$g.json'));e=d['e2e'];print('groups=$g', 'e2e', round(e['value'],2), round(e['ms_per_step'],1), 'ms; phases', round(e['phase_after_phase']['ms_per_step'],1), 'ms; early', e['groups_downloaded_during_upload'])" | tee -a gpurun_out/step_host_sweep.txt
done
