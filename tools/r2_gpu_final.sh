#!/bin/bash
# final validation of HEAD: full GPU suite (as the driver runs it), smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_gpu_tests_head.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_tests_head.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_omni_head.json 2> gpurun_out/r02_bench_omni_head.err; echo "bench rc=$?"; head -c 400 gpurun_out/r02_bench_omni_head.json; echo
timeout 600 python bench.py --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_mgpstr_head.json 2> gpurun_out/r02_bench_mgpstr_head.err; head -c 300 gpurun_out/r02_bench_mgpstr_head.json; echo
