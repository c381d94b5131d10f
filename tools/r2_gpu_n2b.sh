#!/bin/bash
# 2-GPU check of the OmniParser bench after the collective-order fix (short limits: a hang must not burn the box)
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --watchdog-s 150 > gpurun_out/r2_bench_omni_n2.json 2> gpurun_out/r2_bench_omni_n2.err; echo "bench n2 rc=$?"; tail -1 gpurun_out/r2_bench_omni_n2.json | cut -c1-2500; grep -v -i "warn" gpurun_out/r2_bench_omni_n2.err | tail -40 | cut -c1-220
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload table --steps 6 --warmup 3 --no-cpu-baseline --watchdog-s 150 > gpurun_out/r2_bench_table_n2.json 2> gpurun_out/r2_bench_table_n2.err; echo "table n2 rc=$?"; tail -1 gpurun_out/r2_bench_table_n2.json | cut -c1-600
