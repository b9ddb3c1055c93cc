#!/bin/bash
# usage: gpu_mgpu.sh N  — multi-GPU parity + scaling bench on N GPUs of one box
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_contexts" > gpurun_out/pytest_part.log 2>&1; tail -3 gpurun_out/pytest_part.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_$N.log; grep -E "mgpu_check|rc=|Error|error" gpurun_out/mgpu_check_$N.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_euler_n$N.json 2> gpurun_out/bench_euler_n$N.err; cat gpurun_out/bench_euler_n$N.json; tail -5 gpurun_out/bench_euler_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --model ns --no-e2e > gpurun_out/bench_ns_n$N.json 2> gpurun_out/bench_ns_n$N.err; cat gpurun_out/bench_ns_n$N.json; tail -5 gpurun_out/bench_ns_n$N.err
