#!/bin/bash
# A/B of builds on one box (ALM_OCR_LIB): P = before attn_tc fast path / wattn_ms exp2, A = with them, current = A + base-2 cross-attention
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omniparser.py -x -q -m gpu -rA --timeout 600 > gpurun_out/r2_t13.log 2>&1; echo "omni tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t13.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t13.log | head
run_omni() { # label lib opts
  ALM_OCR_LIB=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab.json'))
    print("omni $1 $3:", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'enc', round(d['encoder_ms_per_batch'], 2), 'selfcheck', d['self_check']['ok'])
except Exception as e:
    print("omni $1 failed", e, open('gpurun_out/r2_ab.err').read()[-800:])
PY
}
run_mgp() {
  ALM_OCR_LIB=$2 timeout 400 python bench.py --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_abm.json 2> gpurun_out/r2_abm.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_abm.json'))
    print('mgpstr $1:', round(d['value'], 1), 'crops/s  iso ms', round(d['forward_ms_per_batch_isolated'], 2))
except Exception as e:
    print('mgp $1 failed', e, open('gpurun_out/r2_abm.err').read()[-800:])
PY
}
CUR=$PWD/advancedliteratemachinery_b200/libalm_ocr.so
run_mgp P $PWD/ab/libalm_P.so
run_mgp A $PWD/ab/libalm_A.so
run_mgp P $PWD/ab/libalm_P.so
run_mgp A $PWD/ab/libalm_A.so
run_omni A $PWD/ab/libalm_A.so ""
run_omni cur $CUR ""
run_omni A $PWD/ab/libalm_A.so ""
run_omni cur $CUR ""
run_omni P $PWD/ab/libalm_P.so "--opt wattn_impl=3"
run_omni cur $CUR "--opt wattn_impl=3"
