#!/bin/bash
# A/B on one box: A = checkpoint build, B1 = A + hot mbarrier waits in attn_tc / wattn_ms, cur = B1 + pipelined TMEM loads in the TMA-store epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -x -q -m gpu --timeout 600 > gpurun_out/r2_t16.log 2>&1; echo "tests (cur) rc=$?"; grep -E "passed|failed" gpurun_out/r2_t16.log | tail -1; grep -E "^(FAILED|ERROR)|Error" gpurun_out/r2_t16.log | head -5
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
    print('mgpstr $1:', round(d['value'], 1), 'crops/s  iso ms', round(d['forward_ms_per_batch_isolated'], 2), 'gemms', round(d['roofline']['all_gemms_per_step']['ms'], 2), 'fc1 TF/s', round(d['roofline']['achieved'], 1))
except Exception as e:
    print('mgp $1 failed', e, open('gpurun_out/r2_abm.err').read()[-800:])
PY
}
CUR=$PWD/advancedliteratemachinery_b200/libalm_ocr.so
for rep in 1 2; do
run_mgp A $PWD/ab/libalm_A.so
run_mgp B1 $PWD/ab/libalm_B1.so
run_mgp cur $CUR
done
for rep in 1 2; do
run_omni A $PWD/ab/libalm_A.so ""
run_omni B1 $PWD/ab/libalm_B1.so ""
run_omni cur $CUR ""
done
