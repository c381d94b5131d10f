"""encode 16 x 1024^2 once, then decode with a short pt loop (for ncu launch lists of the decode kernels)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab, synthetic as W
sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = OmniParserB200(sd, OmniVocab(pt_seq_length=n))
img = torch.randn(16, 3, 1024, 1024, device='cuda')
m.encode(img, None)
for _ in range(2):
    outs = m.decode()
print('done', outs[0][0][2].shape)
