#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -m gpu -q -rA --timeout 600 \
  -k "linear_tcgen05 or window_attention_core or encoder_matches or decode_matches or forward_matches or variants or full_size_page or config3" > gpurun_out/r2_t6.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t6.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t6.log | head; grep -E "Error|assert " gpurun_out/r2_t6.log | head -8 | cut -c1-250
python - <<'PY'
import sys
sys.path.insert(0, '.')
from advancedliteratemachinery_b200 import _lib
c = _lib.Context(0)
shapes = [('vit fc1', 131584, 3072, 768, 1, 1), ('vit qkv', 131584, 2304, 768, 1, 0), ('vit fc2', 131584, 768, 3072, 0, 0),
          ('s2 fc1', 65536, 2048, 512, 1, 1), ('s2 fc2', 65536, 512, 2048, 0, 0), ('s0 fc1', 1048576, 512, 128, 1, 1), ('s0 qkv', 1073296, 384, 128, 1, 0), ('s1 fc1', 262144, 1024, 256, 1, 1)]
for ns in (1, 3):
    c.set_option('nsplit', ns)
    for name, M, N, K, split, act in shapes:
        r = []
        for plain in (0, 1):
            c.set_option('gemm_plain_epilogue', plain)
            r.append(c.bench_gemm_ex(M, N, K, 1, split, act, iters=10)[0])
        print(f'nsplit {ns} {name:8s} {M}x{N}x{K}: generic(8 warps) {r[0]*1e3:8.1f} us  plain(16 warps) {r[1]*1e3:8.1f} us  -> {2*M*N*K/r[1]/1e9:7.1f} TFLOP/s', flush=True)
PY
for o in "" "--opt wattn_impl=2" "--opt gemm_plain_epilogue=0"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $o > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab.json'))
    print("A/B '$o':", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), 'selfcheck', d['self_check']['ok'], d['roofline']['all_gemms_per_step'])
except Exception as e:
    print("A/B '$o' failed", e, open('gpurun_out/r2_ab.err').read()[-800:])
PY
done
timeout 300 python bench.py --workload mgpstr --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mgpstr bf16', round(d['value']), 'crops/s', d['forward_ms_per_batch_isolated'], d['roofline']['all_gemms_per_step'], d['roofline']['achieved'])"
timeout 300 python bench.py --workload mgpstr --nsplit 3 --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mgpstr split', round(d['value']), 'crops/s', d['forward_ms_per_batch_isolated'])"
