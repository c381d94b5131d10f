#!/bin/bash
# fused tcgen05 attention bring-up: unit op, model path, then timing A/B at B = 512
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mgpstr.py -m gpu -q -rA --timeout 200 -k "fused" > gpurun_out/r2_attn.log 2>&1; echo "attn rc=$?"; tail -30 gpurun_out/r2_attn.log | cut -c1-300
for impl in 1 0; do
timeout 300 python - <<PY 2>&1 | tail -3
import time, torch
from advancedliteratemachinery_b200 import MGPSTRB200, synthetic as W
m = MGPSTRB200(W.mgpstr_state_dict(seed=0))
x = torch.rand(512, 3, 32, 128, generator=torch.Generator().manual_seed(1)).cuda()
for ns in (1, 3):
    m.ctx.set_option('nsplit', ns); m.ctx.set_option('attn_impl', $impl)
    for _ in range(3): m.recognize(x)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10): m.recognize(x)
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print(f'attn_impl $impl nsplit {ns}: {dt*1e3:.1f} ms / 512 crops = {512/dt:.0f} crops/s')
PY
done
