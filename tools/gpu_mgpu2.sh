#!/bin/bash
# usage: gpu_mgpu2.sh N — trace-row halo: partition tests on one GPU, bit-exactness over N GPUs (both transports), scaling benches
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_contexts or state_setters" 2>&1 | tail -2
SDG_EULER_KERNEL=trace timeout 600 python -m pytest tests -m gpu -x -q -k "two_contexts" 2>&1 | tail -2
for tr in ipc nccl; do
for ek in line trace; do
SDG_HALO=$tr SDG_EULER_KERNEL=$ek timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_${N}_${tr}_${ek}.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_${N}_${tr}_${ek}.log; echo "== $tr $ek"; grep -E "mgpu_check|rc=|Error|error" gpurun_out/mgpu_check_${N}_${tr}_${ek}.log | tail -8
done
done
for ek in line trace; do
SDG_EULER_KERNEL=$ek timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_euler_${ek}_n$N.json 2> gpurun_out/bench_euler_${ek}_n$N.err; python -c "
import json;d=json.load(open('gpurun_out/bench_euler_${ek}_n$N.json'));print('EULER $ek N=$N', d['value'], d['ms_per_step'], d['config']['halo'])"; tail -3 gpurun_out/bench_euler_${ek}_n$N.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --model ns --no-e2e > gpurun_out/bench_ns_n$N.json 2> gpurun_out/bench_ns_n$N.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ns_n$N.json'));print('NS N=$N', d['value'], d['ms_per_step'], d['config']['halo'])"; tail -3 gpurun_out/bench_ns_n$N.err
