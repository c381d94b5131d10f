#!/bin/bash
# MGP-STR at 2 GPUs: contexts in flight vs the per-step gather
mkdir -p gpurun_out
for i in 2 3 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload mgpstr --steps 12 --warmup 3 --no-cpu-baseline --watchdog-s 100 --inflight $i > gpurun_out/r2_m2.json 2> gpurun_out/r2_m2.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_m2.json').read().strip().splitlines()[-1]); print('mgpstr N=2 inflight $i:', round(d['value'], 1), 'crops/s  e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 2))
except Exception as e:
    print('failed', e, open('gpurun_out/r2_m2.err').read()[-600:])
PY
done
