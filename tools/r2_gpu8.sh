#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | grep -v Warning | tail -30
import torch, sys
sys.path.insert(0, '.')
from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab, synthetic as W
from oracle import omniparser_ref as O
torch.set_grad_enabled(False)
sd = W.omniparser_state_dict(seed=0, pt_eos_bias=0.45)
g = torch.Generator().manual_seed(1000)
img = torch.randn(1, 3, 1024, 1024, generator=g)
mask = torch.zeros(1, 1024, 1024, dtype=torch.bool)
mem, pos, kpm, hw = O.encode(img, mask, sd)
for plain in (1, 0, 1):
    m = OmniParserB200(sd, OmniVocab(), workspace_mb=8192)
    m.ctx.set_option('gemm_plain_epilogue', plain)
    for rep in range(2):
        m.encode(img.cuda(), None)
        got = m.memory(1)
        d = (got - pos).abs()
        i = int(d.argmax())
        tok, ch = (i // 512) % 4096, i % 512
        print(f'plain={plain} rep={rep}: pos max abs diff {float(d.max()):.3e} at token {tok} (y={tok//64}, x={tok%64}) channel {ch}; got {float(got.reshape(-1)[i]):.6f} ref {float(pos.reshape(-1)[i]):.6f}; n>1e-5: {int((d>1e-5).sum())}; memory rel {float((m.memory(0)-mem).norm()/mem.norm()):.2e}')
    m.ctx.close()
PY
timeout 600 python -m pytest tests/test_gpu_omniparser.py -m gpu -q -rA --timeout 400 -k "tma_cross or ln_fused or full_size_page" > gpurun_out/r2_t8.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r2_t8.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2_t8.log | head
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
