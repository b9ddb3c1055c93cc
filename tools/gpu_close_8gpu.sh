#!/bin/bash
# closing multi-rank evidence on 8 GPUs at HEAD: bit-identity to one GPU (peer-memory transport), the default bench line at N = 8
mkdir -p gpurun_out
SDG_HALO=ipc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/r02_mgpu_check_8gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_mgpu_check_8gpu_final.log
grep -E "mgpu_check|rc=|Error|error" gpurun_out/r02_mgpu_check_8gpu_final.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02_bench_default_8gpu.json 2> gpurun_out/bench_8gpu.err
cat gpurun_out/r02_bench_default_8gpu.json; tail -2 gpurun_out/bench_8gpu.err
