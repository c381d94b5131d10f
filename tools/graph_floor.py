import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advancedliteratemachinery_b200 import _lib
c = _lib.Context(0)
for nodes in (10, 50, 200):
    us = C.c_float()
    c.check(c.lib.alm_bench_graph_floor(c.h, nodes, 20, C.byref(us)))
    print(f'graph of {nodes} trivial dependent kernels: {us.value:.2f} us per node')
