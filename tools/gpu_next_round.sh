#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent.
#   1. the gated GPU tests (pre-processing kernels, > 32 pages per call)
#   2. the regular GPU suite + smoke + default bench (blocking host waits were only smoke-tested)
#   3. a timing of the pre-processing on page-sized inputs
#   4. A/B of the experimental TMA + mbarrier cross-attention (xattn_impl 2) if its test passed:
#        STEPS=8 bash tools/gpu_sweep.sh 4: 4:xattn_impl=2
mkdir -p gpurun_out
ALM_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_preprocess.py tests/test_gpu_omniparser.py -m gpu -q \
    -k "preprocess or gpu_omni_pages or gpu_mgp_crops or page_scale or more_than_32 or tma_cross_attention or config5" > gpurun_out/pytest_unvalidated.log 2>&1
tail -15 gpurun_out/pytest_unvalidated.log
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 700 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 120 python - <<'PY'
import time, numpy as np, torch
from advancedliteratemachinery_b200 import _lib, preprocess as P
ctx = _lib.Context(0)
rng = np.random.default_rng(0)
pages = [rng.integers(0, 256, (1500, 1100, 3), dtype=np.uint8) for _ in range(16)]
dev = [torch.from_numpy(p).cuda() for p in pages]
for src, name in ((pages, 'host uint8 pages (H2D inside)'), (dev, 'device uint8 pages')):
    P.omni_pages(ctx, src, 1024, 1824)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(5):
        nt = P.omni_pages(ctx, src, 1024, 1824)
    torch.cuda.synchronize()
    print(f'pre-processing, 16 pages 1500x1100 -> {tuple(nt.tensors.shape)}, {name}: {(time.time() - t) / 5 * 1e3:.2f} ms')
PY
