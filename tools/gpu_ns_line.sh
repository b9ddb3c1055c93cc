#!/bin/bash
# first contact of the trace-based NS line kernels: NS parity tests, then the NS bench (new vs node-per-thread kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "cns or sphere or ns_target or two_contexts or conservative or walls" --durations=8 > gpurun_out/pytest_ns.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ns.log; tail -30 gpurun_out/pytest_ns.log
timeout 600 python bench.py --model ns --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ns_line.json 2> gpurun_out/bench_ns_line.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ns_line.json'));print('NS line', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_ns_line.err
SDG_NS_NODE_KERNEL=1 timeout 600 python bench.py --model ns --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ns_node.json 2> gpurun_out/bench_ns_node.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ns_node.json'));print('NS node', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_ns_node.err
