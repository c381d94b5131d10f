#!/bin/bash
# TMA-store GEMM epilogue (gemm_plain_epilogue=2) A/B, in-flight sweep with the new defaults (wattn_impl 3, base-2 cross-attention)
mkdir -p gpurun_out
run_omni() { # label opts
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab.json'))
    print("omni $1 [$2]:", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'enc', round(d['encoder_ms_per_batch'], 2), 'selfcheck', d['self_check']['ok'], 'gemms', round(d['roofline']['all_gemms_per_step']['ms'], 1))
except Exception as e:
    print("omni $1 failed", e, open('gpurun_out/r2_ab.err').read()[-800:])
PY
}
run_mgp() {
  timeout 400 python bench.py --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/r2_abm.json 2> gpurun_out/r2_abm.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_abm.json'))
    print('mgpstr $1 [$2]:', round(d['value'], 1), 'crops/s  iso ms', round(d['forward_ms_per_batch_isolated'], 2), 'gemms', round(d['roofline']['all_gemms_per_step']['ms'], 2))
except Exception as e:
    print('mgp $1 failed', e, open('gpurun_out/r2_abm.err').read()[-800:])
PY
}
timeout 600 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -x -q -m gpu --timeout 400 -k "encoder_matches or full_size_page or forward_matches or variants or window_attention_core" > gpurun_out/r2_t14a.log 2>&1; echo "default tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t14a.log | tail -1
ALM_OCR_OPTIONS="gemm_plain_epilogue=2" timeout 600 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -x -q -m gpu --timeout 400 -k "encoder_matches or full_size_page or forward_matches or variants or config3" > gpurun_out/r2_t14b.log 2>&1; echo "tma-store tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t14b.log | tail -1; grep -E "^(FAILED|ERROR)|Error" gpurun_out/r2_t14b.log | head -5
run_mgp base ""
run_mgp tma "--opt gemm_plain_epilogue=2"
run_mgp base ""
run_mgp tma "--opt gemm_plain_epilogue=2"
run_omni base ""
run_omni tma "--opt gemm_plain_epilogue=2"
run_omni base ""
run_omni tma "--opt gemm_plain_epilogue=2"
run_omni if4 "--inflight 4"
run_omni if6 "--inflight 6"
run_omni if7 "--inflight 7"
