#!/bin/bash
# NS parity tests + NS bench line (quick check of a kernel change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cns or ns or two_contexts or walls or sphere or karman or models" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --model ns --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab.json'));print('NS', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['relative_error'][:2])"; tail -2 gpurun_out/bench_ab.err
