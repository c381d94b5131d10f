#!/bin/bash
# One GPU-box session: tests, bring-up battery, bench (both arms), ncu launch list + one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_check.py mgp > gpurun_out/check_mgp.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list of the same bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --batch 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# one full capture of the dominant kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 3 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 0 --batch 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | head -c 3000; tail -5 gpurun_out/bench.err
