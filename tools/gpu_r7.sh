#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
PROBE_SKIP=1 PROBE_MASKS=0,1 timeout 200 python tools/concurrency_probe.py 4 2>&1 | tail -3
STEPS=8 bash tools/gpu_sweep.sh 4: 4:xattn_ctas_per_sm=1 6: 2>&1 | grep -v "^+"
