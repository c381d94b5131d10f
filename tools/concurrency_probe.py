"""How well do the phases of independent in-flight batches overlap on one GPU?
Times encode-only, decode-only and full steps with K concurrent contexts (one host thread + stream each)."""
import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab, _lib
from advancedliteratemachinery_b200 import synthetic as W

torch.set_grad_enabled(False)
KMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 6
OPTS = [a for a in sys.argv[2:] if '=' in a]
sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0)
vocab = OmniVocab(pt_seq_length=128, rec_length=25)
streams = [torch.cuda.Stream() for _ in range(KMAX)]
models = []
for j in range(KMAX):
    cx = _lib.Context(0, streams[j].cuda_stream)
    cx.set_option('workspace_mb', 20480)
    for kv in OPTS:
        cx.set_option(kv.split('=')[0], int(kv.split('=')[1]))
    models.append(OmniParserB200(sd, vocab, ctx=cx))
pages = torch.randn(16, 3, 1024, 1024, generator=torch.Generator().manual_seed(1)).cuda()
for m in models:
    m.encode(pages, None); m.decode()
torch.cuda.synchronize()


def run(kind, K, reps):
    def work(j):
        for _ in range(reps):
            if kind in ('enc', 'all'):
                models[j].encode(pages, None)
            if kind in ('dec', 'all'):
                models[j].decode()
    torch.cuda.synchronize()
    t = time.time()
    ts = [threading.Thread(target=work, args=(j,)) for j in range(K)]
    [x.start() for x in ts]; [x.join() for x in ts]
    torch.cuda.synchronize()
    return (time.time() - t) * 1e3 / (K * reps)


if os.environ.get('PROBE_SKIP'):
    # which kernel class holds the GPU?  decode-only, K = 1 and 4, with one class of layer kernels dropped at a time
    for mask in [int(x) for x in os.environ.get('PROBE_MASKS', '0,1,2,4,8,3,15').split(',')]:
        for m in models:
            m.ctx.set_option('debug_skip', mask)
            m.decode()
        print(f'skip={mask:2d}: dec K=1 {run("dec", 1, 2):7.1f}   K=4 {run("dec", 4, 2):7.1f} ms per batch', flush=True)
    sys.exit(0)
if os.environ.get('PROBE_OPTS'):
    # decode-only scaling under alternative launch structures
    for opts in ([], ['use_graphs=0'], ['decode_streams=1'], ['use_graphs=0', 'decode_streams=1']):
        for m in models:
            for kv in ('use_graphs=1', 'decode_streams=2'):
                m.ctx.set_option(kv.split('=')[0], int(kv.split('=')[1]))
            for kv in opts:
                m.ctx.set_option(kv.split('=')[0], int(kv.split('=')[1]))
            m.decode()
        print(f'{str(opts):40s}: dec K=1 {run("dec", 1, 2):7.1f}   K=2 {run("dec", 2, 2):7.1f}   K=4 {run("dec", 4, 2):7.1f} ms per batch', flush=True)
    sys.exit(0)
for kind in ('enc', 'dec', 'all'):
    for K in (1, 2, 3, 4, 6):
        if K > KMAX:
            continue
        print(f'{kind} K={K}: {run(kind, K, 2):7.1f} ms per batch', flush=True)
