#!/bin/bash
# attn_tc two-threads-per-row validation, CUDA_DEVICE_MAX_CONNECTIONS A/B, full GPU suite
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mgpstr.py -m gpu -q -rA --timeout 300 > gpurun_out/r2_t9_mgp.log 2>&1; echo "mgp tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t9_mgp.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t9_mgp.log | head
for ns in 1 3; do
timeout 400 python bench.py --workload mgpstr --nsplit $ns --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mgpstr_v2_$ns.json 2> gpurun_out/r2_bench_mgpstr_v2_$ns.err; echo "mgp bench nsplit=$ns rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_bench_mgpstr_v2_$ns.json'))
    print('mgpstr nsplit $ns:', round(d['value'], 1), 'crops/s  e2e', round(d['e2e']['value'], 1), 'iso ms', round(d['forward_ms_per_batch_isolated'], 2), 'all gemms', d['roofline']['all_gemms_per_step'])
except Exception as e:
    print('mgp bench failed', e, open('gpurun_out/r2_bench_mgpstr_v2_$ns.err').read()[-1500:])
PY
done
for mc in "" 32; do
  if [ -n "$mc" ]; then export CUDA_DEVICE_MAX_CONNECTIONS=$mc; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_mc$mc.json 2> gpurun_out/r2_ab_mc$mc.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab_mc$mc.json'))
    print("omni max_connections='$mc':", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'selfcheck', d['self_check']['ok'])
except Exception as e:
    print("omni A/B '$mc' failed", e, open('gpurun_out/r2_ab_mc$mc.err').read()[-800:])
PY
done
unset CUDA_DEVICE_MAX_CONNECTIONS
timeout 900 python -m pytest tests -x -q -m gpu -rA --timeout 600 > gpurun_out/r2_all2.log 2>&1; echo "full suite rc=$?"; grep -E "passed|failed" gpurun_out/r2_all2.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_all2.log | head
