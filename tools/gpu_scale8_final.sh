#!/bin/bash
# 8-GPU closing run: parity of the multi-rank path incl. the node-viscosity all-reduce (peer-memory transport), then the default bench line
mkdir -p gpurun_out
SDG_HALO=ipc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/r02_mgpu_check_av_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_mgpu_check_av_8gpu.log; grep -E "mgpu_check|rc=|Error|error" gpurun_out/r02_mgpu_check_av_8gpu.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_default_8gpu.json 2> gpurun_out/bench_8gpu.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_default_8gpu.json'));print('N=8 EULER', d['value'], d['ms_per_step'], 'NS', d.get('ns_target',{}).get('value'), 'e2e', d.get('e2e',{}).get('value'))"; tail -2 gpurun_out/bench_8gpu.err
