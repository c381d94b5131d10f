#!/bin/bash
# wide K/V projection GEMMs (column-group scatter): parity tests + bench, then the round-2 ncu launch list of one config-2 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omniparser.py -x -q -m gpu -rA --timeout 600 > gpurun_out/r2_t10.log 2>&1; echo "omni tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t10.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t10.log | head
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_omni_v3.json 2> gpurun_out/r2_bench_omni_v3.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_bench_omni_v3.json'))
    print("omni:", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'enc', round(d['encoder_ms_per_batch'], 2), 'selfcheck', d['self_check']['ok'], 'gemms', d['roofline']['all_gemms_per_step'])
except Exception as e:
    print("omni bench failed", e, open('gpurun_out/r2_bench_omni_v3.err').read()[-800:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_omni.csv python tools/one_step.py omni > gpurun_out/ncu_omni.log 2>&1; echo "ncu rc=$?"
python tools/ncu_launches.py gpurun_out/r02_launches_omni.csv > gpurun_out/r02_launches_omni_summary.txt 2>&1; gzip -f gpurun_out/r02_launches_omni.csv; head -45 gpurun_out/r02_launches_omni_summary.txt
