#!/bin/bash
# GPU parity suite only (no -x: collect every failure in one call), durations of the slow tests included
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
