#!/bin/bash
# wattn_impl 3 (persistent TMA-fed mma.sync window attention): op test, A/B bench with self-check, ncu captures
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_omniparser.py -q -m gpu -rA --timeout 200 -k "window_attention_core" > gpurun_out/r2_t11.log 2>&1; echo "wattn tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t11.log | tail -2; grep -E "^(FAILED|ERROR)|assert|Error" gpurun_out/r2_t11.log | head
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:window_attention_ms --launch-skip 8 --launch-count 1 \
    -o gpurun_out/r02_prof_wattn_ms python tools/one_step.py omni wattn_impl=3 > gpurun_out/ncu_wattn_ms.log 2>&1; tail -2 gpurun_out/ncu_wattn_ms.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc --launch-skip 2 --launch-count 1 \
    -o gpurun_out/r02_prof_attn_tc python tools/one_step.py mgpstr > gpurun_out/ncu_attn_tc.log 2>&1; tail -2 gpurun_out/ncu_attn_tc.log
ls -la gpurun_out/*.ncu-rep | tail -3
