#!/bin/bash
# small sweeps with the final build: MGP-STR contexts in flight, cross-attention CTAs per SM
mkdir -p gpurun_out
for i in 2 3 1 4 2 3; do
timeout 300 python bench.py --workload mgpstr --steps 12 --warmup 3 --no-cpu-baseline --inflight $i > gpurun_out/r2_sw.json 2> gpurun_out/r2_sw.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_sw.json')); print('mgpstr inflight $i:', round(d['value'], 1), 'crops/s  e2e', round(d['e2e']['value'], 1))
except Exception as e:
    print('failed', e, open('gpurun_out/r2_sw.err').read()[-600:])
PY
done
for o in "" "--opt xattn_ctas_per_sm=3" "" "--opt xattn_ctas_per_sm=3"; do
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $o > gpurun_out/r2_sw.json 2> gpurun_out/r2_sw.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_sw.json')); print('omni [$o]:', round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d['phase_ms'], d['self_check']['ok'])
except Exception as e:
    print('failed', e, open('gpurun_out/r2_sw.err').read()[-600:])
PY
done
