#!/bin/bash
# usage: gpu_mgpu_av.sh N — multi-GPU parity incl. shock capturing (node maximum all-reduced), both halo transports
N=${1:-2}
mkdir -p gpurun_out
for tr in ipc nccl; do
SDG_HALO=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/r02_mgpu_check_av_${N}gpu_${tr}.log 2>&1; echo "rc=$?" >> gpurun_out/r02_mgpu_check_av_${N}gpu_${tr}.log; echo "== $tr"; grep -E "mgpu_check|rc=|Error|error" gpurun_out/r02_mgpu_check_av_${N}gpu_${tr}.log | tail -12
done
