#!/bin/bash
# base-2 softmax in wattn_ms, unpredicated fast path in attn_tc: tests + benches
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -q -m gpu -rA --timeout 300 -k "window_attention_core or mgpstr" > gpurun_out/r2_t12.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t12.log | tail -2; grep -E "^(FAILED|ERROR)|assert|Error" gpurun_out/r2_t12.log | head
for o in "" "--opt wattn_impl=3"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $o > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab.json'))
    print("A/B '$o':", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'enc', round(d['encoder_ms_per_batch'], 2), 'selfcheck', d['self_check']['ok'])
except Exception as e:
    print("A/B '$o' failed", e, open('gpurun_out/r2_ab.err').read()[-800:])
PY
done
for ns in 1 3; do
timeout 400 python bench.py --workload mgpstr --nsplit $ns --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mgpstr_v3_$ns.json 2> gpurun_out/r2_bench_mgpstr_v3_$ns.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_bench_mgpstr_v3_$ns.json'))
    print('mgpstr nsplit $ns:', round(d['value'], 1), 'crops/s  e2e', round(d['e2e']['value'], 1), 'iso ms', round(d['forward_ms_per_batch_isolated'], 2))
except Exception as e:
    print('mgp bench failed', e, open('gpurun_out/r2_bench_mgpstr_v3_$ns.err').read()[-1500:])
PY
done
