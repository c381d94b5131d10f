#!/bin/bash
# 4-GPU lines of the other BASELINE configs: config-4 stand-in (platypus) and config 3 (MGP-STR bf16)
mkdir -p gpurun_out
for wl in platypus mgpstr; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --watchdog-s 100 > gpurun_out/r02_bench_${wl}_n4.json 2> gpurun_out/r02_bench_${wl}_n4.err; echo "$wl n4 rc=$?"; tail -1 gpurun_out/r02_bench_${wl}_n4.json | cut -c1-500
done
