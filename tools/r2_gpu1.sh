#!/bin/bash
# Round 2, GPU call 1: run every previously gated GPU test (never executed on a B200 before), one group per process.
mkdir -p gpurun_out
export ALM_RUN_UNVALIDATED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
timeout 200 python -m pytest tests/test_preprocess.py -m gpu -q -rA --timeout 120 > gpurun_out/r2_pre.log 2>&1; echo "pre rc=$?"; tail -25 gpurun_out/r2_pre.log
timeout 300 python -m pytest tests/test_gpu_omniparser.py -m gpu -q -rA --timeout 200 -k "more_than_32" > gpurun_out/r2_gt32.log 2>&1; echo "gt32 rc=$?"; tail -25 gpurun_out/r2_gt32.log
timeout 300 python -m pytest tests/test_gpu_omniparser.py -m gpu -q -rA --timeout 200 -k "config5" > gpurun_out/r2_c5.log 2>&1; echo "c5 rc=$?"; tail -25 gpurun_out/r2_c5.log
timeout 300 python -m pytest tests/test_gpu_omniparser.py -m gpu -q -rA --timeout 200 -k "tma_cross" > gpurun_out/r2_tma.log 2>&1; echo "tma rc=$?"; tail -25 gpurun_out/r2_tma.log
timeout 600 python -m pytest tests -m gpu -q -rA --timeout 300 > gpurun_out/r2_all.log 2>&1; echo "all rc=$?"; tail -40 gpurun_out/r2_all.log
