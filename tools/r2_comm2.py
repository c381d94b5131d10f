"""torchrun --nproc-per-node N tools/r2_comm2.py: the C-ABI multi-GPU path end to end.
  1. alm_comm_init on every rank (id shipped through torch.distributed), rank 0 loads the checkpoint, the others lay out
     placeholders, ONE alm_broadcast_weights;
  2. EVERY rank decodes the benchmark page (seed 1000) and compares with the reference fixture -> the broadcast weights
     are intact on every GPU;
  3. uneven sharding (7 pages over N ranks) + alm_gather_sequences: every rank gets all pages back in page order, equal to
     what a single rank decodes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from advancedliteratemachinery_b200 import NestedTensor, OmniParserB200, OmniVocab, _lib
from advancedliteratemachinery_b200 import synthetic as W
from advancedliteratemachinery_b200.dist import gather_sequences, init_comm, load_weights_broadcast, shard_pages
from oracle.gen_golden import config2_page

torch.set_grad_enabled(False)
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
ctx = _lib.Context(local)
ctx.set_option('workspace_mb', 20480)
init_comm(ctx, device=dev)
sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0) if rank == 0 else None
load_weights_broadcast(ctx, _lib.MODEL_OMNI_SPOT, sd, src=0, device=dev)
v = OmniVocab(pt_seq_length=128, rec_length=25)
m = OmniParserB200(None, v, ctx=ctx)
gold = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'omni_config2_page0.npz'))
out = m.forward_batch(NestedTensor(config2_page(1000)[0], None))[0]
(pt, poly, rec), _ = out
ok = np.array_equal(pt.numpy(), gold['pt']) and np.array_equal(poly.numpy(), gold['poly']) and np.array_equal(rec.numpy(), gold['rec'])
print(f'rank {rank}: benchmark page after the weight broadcast equals the reference fixture: {ok}', flush=True)
assert ok
# uneven shards, small pages
v2 = OmniVocab(pt_seq_length=8, rec_length=25)
m.vocab = v2
N = 7
pages = [torch.randn(1, 3, 96, 128, generator=torch.Generator().manual_seed(50 + p)) for p in range(N)]
mine = shard_pages(N, rank, world)
outs = m.forward_batch(NestedTensor(torch.cat([pages[p] for p in mine]), None)) if mine else []
allp = gather_sequences(outs, v2, n_pages=N, ctx=ctx)
ref = m.forward_batch(NestedTensor(torch.cat(pages), None))
same = len(allp) == N
for a, b in zip(allp, ref):
    same = same and ((a is None) == (b is None)) and (a is None or all(torch.equal(x, y) for x, y in zip(a[0], b[0])))
print(f'rank {rank}: {N} pages over {world} ranks gathered in page order and equal to the single-rank decode: {same}', flush=True)
assert same
dist.barrier()
dist.destroy_process_group()
