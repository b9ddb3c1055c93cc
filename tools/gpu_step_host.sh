#!/bin/bash
# sdg_step_host on the B200: bit-identity tests, then the default bench line's e2e part (streamed vs phase after phase)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/step_host_smi.txt 2>&1
timeout 600 python -m pytest tests/test_step_host.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/step_host_pytest.log
for g in "" 16 64; do
  if [ -n "$g" ]; then export SDG_HOST_PIPE_GROUPS=$g; else unset SDG_HOST_PIPE_GROUPS; fi
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-ns-target > gpurun_out/step_host_bench_${g:-default}.json 2> gpurun_out/step_host_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/step_host_bench_${g:-default}.json'));e=d['e2e'];print('groups=${g:-default}', 'value', round(d['value'],2), 'e2e', round(e['value'],2), round(e['ms_per_step'],1), 'ms; phases', round(e['phase_after_phase']['value'],2), round(e['phase_after_phase']['ms_per_step'],1), 'ms; groups', e['upload_groups'], e['groups_downloaded_during_upload'])" | tee -a gpurun_out/step_host_ab.txt
  tail -3 gpurun_out/step_host_bench.err
done
