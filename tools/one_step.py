"""One bench-shaped step (16 x 1024^2 pages, N = 64 pinned) on one context, no graphs: the target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab, _lib
from advancedliteratemachinery_b200 import synthetic as W

torch.set_grad_enabled(False)
sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0)
cx = _lib.Context(0)
cx.set_option('workspace_mb', 20480)
cx.set_option('use_graphs', 0)
for kv in sys.argv[1:]:
    cx.set_option(kv.split('=')[0], int(kv.split('=')[1]))
m = OmniParserB200(sd, OmniVocab(pt_seq_length=128, rec_length=25), ctx=cx)
pages = torch.randn(16, 3, 1024, 1024, generator=torch.Generator().manual_seed(1)).cuda()
m.encode(pages, None)
out = m.decode()
print('instances', sum(0 if o is None else o[0][0].numel() // 2 for o in out), cx.omni_last_timing())
