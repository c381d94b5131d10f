#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python tools/gpu_check.py trace > gpurun_out/trace.log 2>&1
grep -A22 "\[decode\]" gpurun_out/trace.log; grep "encode " gpurun_out/trace.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
head -c 1800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
