#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -m gpu -q -rA --timeout 400 \
  -k "tma_cross or fused_cross or fused_tcgen05 or window_attention_core or fused_attention_model" > gpurun_out/r2_t5.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t5.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t5.log | head; grep -E "Error|assert " gpurun_out/r2_t5.log | head -8 | cut -c1-250
for o in "" "--opt xattn_impl=3"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $o > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab.json'))
    print("A/B '$o':", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'selfcheck', d['self_check']['ok'], d['self_check'].get('poly_sequences_differing'), d['self_check'].get('rec_sequences_differing'))
except Exception as e:
    print("A/B '$o' failed", e, open('gpurun_out/r2_ab.err').read()[-800:])
PY
done
timeout 300 python bench.py --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mgpstr bf16', round(d['value']), 'crops/s', d['forward_ms_per_batch_isolated'], d['roofline']['all_gemms_per_step'])"
timeout 300 python bench.py --workload mgpstr --nsplit 3 --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mgpstr split', round(d['value']), 'crops/s', d['forward_ms_per_batch_isolated'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref_omni.json 2> gpurun_out/r2_ref_omni.err; tail -1 gpurun_out/r2_ref_omni.json | cut -c1-1500
