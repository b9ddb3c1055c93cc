#!/bin/bash
# usage: gpu_scale.sh N — scaling bench lines only (Euler 128^3 and NS 96^3) on N GPUs of one box
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_euler_n$N.json 2> gpurun_out/bench_euler_n$N.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_euler_n$N.json') if l.startswith('{')][0]);print('EULER n=$N', d['value'], d['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --model ns --no-e2e > gpurun_out/bench_ns_n$N.json 2> gpurun_out/bench_ns_n$N.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_ns_n$N.json') if l.startswith('{')][0]);print('NS n=$N', d['value'], d['ms_per_step'])"
