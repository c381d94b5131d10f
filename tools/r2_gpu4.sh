#!/bin/bash
# Round 2, GPU call 4: slim GEMM epilogue + mbarrier hint, tcgen05 cross-attention, window attention with staged bias
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omniparser.py tests/test_gpu_mgpstr.py -m gpu -q -rA --timeout 600 -x \
  -k "linear_tcgen05 or window_attention_core or tma_cross or fused_cross or encoder_matches or decode_matches or forward_matches or fused_tcgen05 or variants or ln_fused or shared_weights" > gpurun_out/r2_t4.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t4.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t4.log | head; grep -E "Error|assert " gpurun_out/r2_t4.log | head -8 | cut -c1-250
python - <<'PY'
import sys
sys.path.insert(0, '.')
from advancedliteratemachinery_b200 import _lib
c = _lib.Context(0)
shapes = [('vit fc1', 131584, 3072, 768, 1, 1), ('vit qkv', 131584, 2304, 768, 1, 0), ('vit fc2', 131584, 768, 3072, 0, 0),
          ('s2 fc1', 65536, 2048, 512, 1, 1), ('s0 fc1', 1048576, 512, 128, 1, 1), ('s0 qkv', 1073296, 384, 128, 1, 0), ('s1 fc1', 262144, 1024, 256, 1, 1)]
for ns in (1, 3):
    c.set_option('nsplit', ns)
    for name, M, N, K, split, act in shapes:
        r = []
        for plain in (0, 1):
            c.set_option('gemm_plain_epilogue', plain)
            r.append(c.bench_gemm_ex(M, N, K, 1, split, act, iters=10)[0])
        print(f'nsplit {ns} {name:8s} {M}x{N}x{K}: generic {r[0]*1e3:8.1f} us  plain {r[1]*1e3:8.1f} us  -> {2*M*N*K/r[1]/1e9:7.1f} TFLOP/s', flush=True)
PY
for o in "" "--opt xattn_impl=3" "--opt wattn_impl=2" "--opt xattn_impl=3 --opt wattn_impl=2"; do
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
timeout 300 python bench.py --workload table --steps 6 --warmup 3 --no-cpu-baseline --opt xattn_impl=3 --opt kv_decoders=1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('table xattn3', round(d['value'],1), 'img/s', d['phase_ms'])"
# source-level capture of the tcgen05 window attention (stage-2 shape)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:window_attention_tc --launch-skip 6 --launch-count 1 \
    -o gpurun_out/r02_prof_wattn_tc python tools/one_step.py omni wattn_impl=2 > gpurun_out/ncu_wattn_tc.log 2>&1; tail -2 gpurun_out/ncu_wattn_tc.log
