#!/bin/bash
# usage: gpu_ipc.sh N — bit-exactness check and scaling bench with the peer-memory (ipc) and the NCCL halo transport
N=${1:-2}
mkdir -p gpurun_out
for tr in ipc nccl; do
  export SDG_HALO=$tr
  if [ $tr == ipc ]; then timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_${tr}_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_${tr}_$N.log; grep -E "mgpu_check|rc=|rror" gpurun_out/mgpu_check_${tr}_$N.log | tail -6 | cut -c1-200; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_euler_${tr}_n$N.json 2> gpurun_out/bench_euler_${tr}_n$N.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_euler_${tr}_n$N.json') if l.startswith('{')][0]);print('EULER $tr n=$N', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/bench_euler_${tr}_n$N.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --model ns --no-e2e > gpurun_out/bench_ns_${tr}_n$N.json 2> gpurun_out/bench_ns_${tr}_n$N.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_ns_${tr}_n$N.json') if l.startswith('{')][0]);print('NS $tr n=$N', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/bench_ns_${tr}_n$N.err
done
