#!/bin/bash
# 2-GPU A/B of who issues the per-batch gathers (lane vs order) + the 1-GPU number of the same box
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d = json.loads(open('$1').read().strip().splitlines()[-1])
    print('$2:', round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 1), 'selfcheck', d['self_check'])
except Exception as e:
    print('$2 failed', e, open('$1'.replace('.json', '.err')).read()[-1500:])
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n1.json 2> gpurun_out/r2_n1.err; show gpurun_out/r2_n1.json "N=1"
for mode in lane order lane order; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --watchdog-s 120 --collectives $mode > gpurun_out/r2_n2_$mode.json 2> gpurun_out/r2_n2_$mode.err; echo "rc=$?"; show gpurun_out/r2_n2_$mode.json "N=2 $mode"
done
grep -v -i warn gpurun_out/r2_n2_lane.err | tail -5 | cut -c1-200
