#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
export SDG_HALO=ipc
for nb in 32 96 296; do
  export SDG_PUSH_BLOCKS=$nb
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/b.json') if l.startswith('{')][0]);print('EULER ipc blocks=$nb n=$N', d['value'], d['ms_per_step'])" || tail -3 gpurun_out/b.err
done
