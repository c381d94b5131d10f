#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for cfg in "3:" "4:" "3:--opt xattn_impl=1"; do
  n=${cfg%%:*}; o=${cfg#*:}; tag=$(echo "i${n}${o}" | tr -d ' =-')
  timeout 300 python bench.py --inflight $n --steps 6 --no-cpu-baseline $o > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_{t}.json').read().strip().splitlines()[-1])
    print(t, 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms', round(d['ms_per_step'],1), d['phase_ms'], 'launches', d['gpu_launches'])
except Exception as e:
    print(t, 'ERR', e); print(open(f'gpurun_out/bench_{t}.err').read()[-1500:])
PY
done
