#!/bin/bash
# 2-GPU validation of the C-ABI NCCL path + bench at N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/r2_comm2.py > gpurun_out/r2_comm2.log 2>&1; echo "comm2 rc=$?"; grep -v Warning gpurun_out/r2_comm2.log | tail -8 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_omni_n2.json 2> gpurun_out/r2_bench_omni_n2.err; echo "bench n2 rc=$?"; tail -1 gpurun_out/r2_bench_omni_n2.json | cut -c1-2200; tail -3 gpurun_out/r2_bench_omni_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mgpstr_n2.json 2> gpurun_out/r2_bench_mgpstr_n2.err; echo "mgp n2 rc=$?"; tail -1 gpurun_out/r2_bench_mgpstr_n2.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --workload mgpstr --steps 2 --warmup 1 | tail -1 | cut -c1-300
