#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -m gpu -q -rA --timeout 600 \
  -k "linear_tcgen05 or tma_cross or fused_cross or layernorm or forward_matches or fresh_batch or config3 or fused or variants or encoder_matches" > gpurun_out/r2_t7.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t7.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t7.log | head; grep -E "Error|assert |single-pass" gpurun_out/r2_t7.log | head -8 | cut -c1-250
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
timeout 300 python bench.py --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mgpstr bf16', round(d['value']), 'crops/s', d['forward_ms_per_batch_isolated'], d['roofline']['all_gemms_per_step'], d['roofline']['achieved'])"
timeout 300 python bench.py --workload mgpstr --nsplit 3 --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mgpstr split', round(d['value']), 'crops/s', d['forward_ms_per_batch_isolated'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_mgpstr.csv python tools/one_step.py mgpstr > gpurun_out/ncu_mgp2.log 2>&1
python tools/ncu_launches.py gpurun_out/r02b_launches_mgpstr.csv > gpurun_out/r02b_launches_mgpstr_summary.txt 2>&1; gzip -f gpurun_out/r02b_launches_mgpstr.csv; head -12 gpurun_out/r02b_launches_mgpstr_summary.txt
