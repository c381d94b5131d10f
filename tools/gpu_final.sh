#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 2600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 900 gpurun_out/bench_ref.json
