#!/bin/bash
# checkpoint with the final defaults: full GPU suite, smoke, default bench (with CPU baseline), the other workloads, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu -rA --timeout 900 > gpurun_out/r02_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r02_gpu_tests_final.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r02_gpu_tests_final.log | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_omni_n1.json 2> gpurun_out/r02_bench_omni_n1.err; echo "bench omni rc=$?"; head -c 3000 gpurun_out/r02_bench_omni_n1.json; echo; tail -2 gpurun_out/r02_bench_omni_n1.err
for wl in mgpstr table platypus; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r02_bench_${wl}_n1.json 2> gpurun_out/r02_bench_${wl}_n1.err; echo "bench $wl rc=$?"; head -c 700 gpurun_out/r02_bench_${wl}_n1.json; echo
done
timeout 600 python bench.py --workload mgpstr --nsplit 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_mgpstr_split_n1.json 2> gpurun_out/r02_bench_mgpstr_split_n1.err; head -c 400 gpurun_out/r02_bench_mgpstr_split_n1.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err; echo "reference rc=$?"; head -c 900 gpurun_out/r02_bench_reference_n1.json; echo
