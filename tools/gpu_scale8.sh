#!/bin/bash
# 8-GPU check: bit-exactness against one GPU (peer-memory transport), then the driver's bench invocation at N = 8 (Euler + ns_target)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_$N.log; grep -E "mgpu_check|rc=|Error|error" gpurun_out/mgpu_check_$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_n$N.json | cut -c1-3000; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
