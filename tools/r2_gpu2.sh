#!/bin/bash
# Round 2, GPU call 2: full GPU suite (new parity tests), smoke, every bench workload on 1 GPU, ncu launch lists.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > gpurun_out/r2_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_all.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r2_all.log | head -20
grep -h "config-2\|single-pass" gpurun_out/r2_all.log | head -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in omni mgpstr table platypus; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "bench $wl rc=$?"; head -c 1500 gpurun_out/r2_bench_$wl.json; echo; tail -3 gpurun_out/r2_bench_$wl.err
done
timeout 600 python bench.py --workload mgpstr --nsplit 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mgpstr_split.json 2> gpurun_out/r2_bench_mgpstr_split.err; head -c 600 gpurun_out/r2_bench_mgpstr_split.json; echo
# ncu launch lists (cold-cache, serialised: shares only)
for wl in mgpstr table; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_$wl.csv python tools/one_step.py $wl > gpurun_out/ncu_$wl.log 2>&1
  python tools/ncu_launches.py gpurun_out/r02_launches_$wl.csv > gpurun_out/r02_launches_${wl}_summary.txt 2>&1
  gzip -f gpurun_out/r02_launches_$wl.csv
  head -25 gpurun_out/r02_launches_${wl}_summary.txt
done
