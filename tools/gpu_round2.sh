#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py tc wattn dec perf > gpurun_out/check2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.log; head -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
