"""One bench-shaped step of a workload on ONE context without CUDA graphs: the target of the ncu launch lists / captures.

    python tools/one_step.py [omni|mgpstr|table|platypus] [nsplit=1|3] [option=value ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from advancedliteratemachinery_b200 import MGPSTRB200, OmniParserB200, OmniVocab, _lib
from advancedliteratemachinery_b200 import synthetic as W

torch.set_grad_enabled(False)
args = sys.argv[1:]
name = args.pop(0) if args and args[0] in bench.WORKLOADS else 'omni'
w = bench.WORKLOADS[name]
cx = _lib.Context(0)
opts = dict(kv.split('=') for kv in args)
cx.set_option('nsplit', int(opts.pop('nsplit', 1 if w['kind'] == 'mgp' else 3)))
if w['kind'] == 'mgp':
    m = MGPSTRB200(W.mgpstr_state_dict(seed=0), ctx=cx)
    for k, v in opts.items():
        cx.set_option(k, int(v))
    x = torch.rand(w['batch'], 3, 32, 128, generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(2):
        ids, prob = m.recognize(x)
    print('mgpstr', tuple(ids.shape), 'decoded chars', bench.mgp_decoded_chars(ids))
else:
    cx.set_option('workspace_mb', w['workspace_mb'])
    cx.set_option('use_graphs', 0)
    for k, v in opts.items():
        cx.set_option(k, int(v))
    m = OmniParserB200(W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0), OmniVocab(pt_seq_length=w['pt_len'], rec_length=25), ctx=cx)
    pages = torch.randn(w['batch'], 3, w['page'], w['page'], generator=torch.Generator().manual_seed(1)).cuda()
    m.encode(pages, None)
    out = m.decode_points() if w['points_only'] else m.decode()
    print(name, 'units decoded', len(out), cx.omni_last_timing())
