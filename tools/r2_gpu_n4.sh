#!/bin/bash
# 4-GPU sanity run of the default bench (weight broadcast, 5 contexts per rank, ordered per-step gathers)
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --watchdog-s 150 > gpurun_out/r02_bench_omni_n4.json 2> gpurun_out/r02_bench_omni_n4.err; echo "bench n4 rc=$?"; tail -1 gpurun_out/r02_bench_omni_n4.json | cut -c1-1800; grep -v -i "warn" gpurun_out/r02_bench_omni_n4.err | tail -5 | cut -c1-220
