#!/bin/bash
# Round 2, GPU call 3: tcgen05 window attention bring-up, A/B runs, source-level ncu capture of the GEMM epilogue
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_omniparser.py -m gpu -q -rA --timeout 300 -k "window_attention_core" > gpurun_out/r2_wattn.log 2>&1; echo "wattn rc=$?"; tail -15 gpurun_out/r2_wattn.log | cut -c1-250
# encoder with the tcgen05 window attention: parity + timing
timeout 600 python - <<'PY' 2>&1 | tail -12
import torch, time
from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab, synthetic as W
from oracle import omniparser_ref as O
torch.set_grad_enabled(False)
sd = W.omniparser_state_dict(seed=0, pt_eos_bias=0.45)
m = OmniParserB200(sd, OmniVocab(pt_seq_length=6), workspace_mb=20480)
g = torch.Generator().manual_seed(5)
img = torch.randn(1, 3, 320, 448, generator=g)
mem, pos, kpm, _ = O.encode(img, torch.zeros(1, 320, 448, dtype=torch.bool), sd)
for impl in (0, 2):
    m.ctx.set_option('wattn_impl', impl)
    m.encode(img.cuda(), None)
    got = m.memory(0)
    print(f'wattn_impl {impl}: encoder memory rel err vs oracle {float((got - mem).norm() / mem.norm()):.2e}')
pages = torch.randn(16, 3, 1024, 1024, generator=g).cuda()
for impl in (0, 2):
    m.ctx.set_option('wattn_impl', impl)
    for _ in range(2): m.encode(pages, None)
    m.ctx.synchronize(); t = time.time()
    for _ in range(5): m.encode(pages, None)
    m.ctx.synchronize()
    print(f'wattn_impl {impl}: encoder 16 x 1024^2 = {(time.time() - t) / 5 * 1e3:.2f} ms', m.ctx.omni_last_timing())
PY
# A/B at the benchmark configuration
for o in "" "--opt xattn_impl=2" "--opt wattn_impl=2"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $o > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_ab.json'))
    print("A/B '$o':", round(d['value'], 1), 'img/s  e2e', round(d['e2e']['value'], 1), d.get('phase_ms'), d.get('self_check'))
except Exception as e:
    print("A/B '$o' failed", e, open('gpurun_out/r2_ab.err').read()[-500:])
PY
done
# source-level capture of the epilogue-bound GEMM (ViT fc1 shape, single-pass bf16, GELU + split output)
cat > /tmp/one_gemm.py <<'PY'
import sys
sys.path.insert(0, '.')
from advancedliteratemachinery_b200 import _lib
c = _lib.Context(0)
c.set_option('nsplit', int(sys.argv[1]))
print(c.bench_gemm_ex(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), 1, 1, 1, iters=3)[0], 'ms')
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 3 --launch-count 1 \
    -o gpurun_out/r02_prof_gemm_vit_fc1_bf16 python /tmp/one_gemm.py 1 131584 3072 768 > gpurun_out/ncu_gemm1.log 2>&1; tail -2 gpurun_out/ncu_gemm1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 3 --launch-count 1 \
    -o gpurun_out/r02_prof_gemm_s0_fc1_split python /tmp/one_gemm.py 3 1048576 512 128 > gpurun_out/ncu_gemm2.log 2>&1; tail -2 gpurun_out/ncu_gemm2.log
ls -la gpurun_out/*.ncu-rep | tail -4
