"""Run one GEMM shape a few times (for ncu): python tools/bench_one_gemm.py M N K batch split act"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advancedliteratemachinery_b200 import _lib
M, N, K, batch, split, act = [int(x) for x in sys.argv[1:7]]
c = _lib.Context(0)
ms, _ = c.bench_gemm_ex(M, N, K, batch, split, act, iters=3)
print(f'M={M} N={N} K={K} batch={batch} split={split} act={act}: {ms * 1e3:.1f} us/launch')
