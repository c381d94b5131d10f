import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from advancedliteratemachinery_b200 import NestedTensor, OmniParserB200, OmniVocab, synthetic as W
from oracle import omniparser_ref as O
from oracle.gen_golden import OMNI_CASES, omni_inputs
torch.set_grad_enabled(False)
case = OMNI_CASES['odd']; gold = np.load('tests/golden/omni_odd.npz')
sd = W.omniparser_state_dict(seed=case['wseed'], pt_eos_bias=case['pt_eos_bias'])
m = OmniParserB200(sd, OmniVocab(pt_seq_length=case['pt_seq_length']), workspace_mb=8192)
img, mask = omni_inputs(case)
def run(tag):
    m.encode(img, mask)
    out = m.decode()[0]
    (pt, poly, rec), (pr,) = out
    bad = np.argwhere(rec.numpy() != gold['rec'])
    print(tag, 'pt ok', np.array_equal(pt.numpy(), gold['pt']), 'poly ok', np.array_equal(poly.numpy(), gold['poly']),
          'rec mismatches at', bad[:4].tolist(), flush=True)
for streams in (2, 1):
    m.ctx.set_option('decode_streams', streams)
    for graphs in (1, 0):
        m.ctx.set_option('use_graphs', graphs)
        for i in range(3):
            run(f'streams={streams} graphs={graphs} #{i}')
# teacher-forced logits at every position vs the oracle
mem, pos, kpm, _ = O.encode(img, mask, sd)
m.encode(img, mask)
pt = torch.from_numpy(gold['pt']); n = pt.numel() // 2
rec_full = torch.cat([pt.reshape(-1, 2), torch.full((n, 1), 1102), torch.from_numpy(gold['rec'])[0]], 1)
ref = O.decode_logits(rec_full, mem[0], kpm[0], pos[0], sd, 'rec')
got = m.decode_logits(0, 'rec', rec_full)
err = (got - ref).abs().amax(-1)
print('teacher-forced rec max abs err per position (seq0):', [f'{e:.1e}' for e in err[0].tolist()])
print('teacher-forced rec max abs err per position (seq1):', [f'{e:.1e}' for e in err[1].tolist()])
