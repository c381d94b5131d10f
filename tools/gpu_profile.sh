#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the dominant GEMM shapes
set -x
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/r01_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --inflight 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_launches.py gpurun_out/r01_launches_bench.csv > gpurun_out/r01_launches_bench_summary.txt 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 1 -o gpurun_out/r01_prof_gemm_fc1 \
    python tools/bench_one_gemm.py 65536 2048 512 1 1 1 > gpurun_out/ncu_fc1.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 1 -o gpurun_out/r01_prof_gemm_fc2 \
    python tools/bench_one_gemm.py 65536 512 2048 1 0 0 > gpurun_out/ncu_fc2.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:window_attention_mma -c 1 -o gpurun_out/r01_prof_wattn \
    python tools/gpu_check.py --run wattn > gpurun_out/ncu_wattn.log 2>&1
tail -2 gpurun_out/ncu_fc1.log gpurun_out/ncu_fc2.log; head -30 gpurun_out/r01_launches_bench_summary.txt
