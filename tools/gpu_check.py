"""GPU bring-up battery (run on the B200 box): each group runs in its own subprocess under `timeout`
so that a trapped kernel cannot poison the next group.  Prints max-abs / relative errors against the CPU
oracle.  Not a test (tests/ has the asserts) -- this is the verbose diagnostic used while debugging.

    python tools/gpu_check.py [group ...]     groups: simt tc tcbig ln wattn enc_simt enc dec perf
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _err(a, b):
    import torch
    a, b = a.double(), b.double()
    d = (a - b).abs()
    return f'max_abs={d.max().item():.3e} rel_fro={(d.norm() / (b.norm() + 1e-30)).item():.3e} ref_absmax={b.abs().max().item():.3e}'


def _ctx(**opts):
    from advancedliteratemachinery_b200 import _lib
    c = _lib.Context(0)
    for k, v in opts.items():
        c.set_option(k, v)
    return c


def _linear(c, M, N, K, act=0, batch=1, bias=True, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(batch, M, K, generator=g)
    W = torch.randn(batch, N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    ref = torch.einsum('bmk,bnk->bmn', A.double(), W.double())
    if bias:
        ref = ref + b.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    Ad, Wd = A.cuda(), W.cuda()
    bd = b.cuda() if bias else None
    out = torch.full((batch, M, N), float('nan'), device='cuda')
    c.check(c.lib.alm_op_linear(c.h, Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr() if bias else None, out.data_ptr(),
                                M, N, K, act, batch))
    torch.cuda.synchronize()
    return out.cpu(), ref


def g_simt():
    c = _ctx(gemm_impl=1)
    for (M, N, K, act, batch) in [(64, 48, 32, 0, 1), (130, 70, 96, 1, 1), (33, 200, 64, 2, 2)]:
        out, ref = _linear(c, M, N, K, act, batch)
        print(f'simt linear M={M} N={N} K={K} act={act} batch={batch}: {_err(out, ref)}')


def g_tc():
    for ns in (3, 1):
        c = _ctx(nsplit=ns)
        for (M, N, K, act, batch) in [(128, 128, 64, 0, 1), (128, 128, 256, 0, 1), (256, 384, 128, 0, 1),
                                      (300, 200, 96, 1, 1), (77, 1104, 512, 2, 1), (49, 27, 96, 0, 3),
                                      (1000, 130, 2048, 0, 2)]:
            t = time.time()
            out, ref = _linear(c, M, N, K, act, batch)
            print(f'tc nsplit={ns} M={M} N={N} K={K} act={act} batch={batch}: {_err(out, ref)}  ({time.time() - t:.2f}s)',
                  flush=True)


def g_tcbig():
    import torch
    c = _ctx(nsplit=3)
    for (M, N, K) in [(20000, 1536, 512), (65536, 512, 128), (8192, 8192, 1024)]:
        out, ref = _linear(c, M, N, K, 0, 1)
        print(f'tc big M={M} N={N} K={K}: {_err(out, ref)}', flush=True)
        A = torch.randn(M, K, device='cuda')
        W = torch.randn(N, K, device='cuda')
        o = torch.empty(M, N, device='cuda')
        for ns in (3, 1):
            c.set_option('nsplit', ns)
            for _ in range(2):
                c.check(c.lib.alm_op_linear(c.h, A.data_ptr(), W.data_ptr(), None, o.data_ptr(), M, N, K, 0, 1))
            torch.cuda.synchronize()
            t = time.time()
            for _ in range(5):
                c.check(c.lib.alm_op_linear(c.h, A.data_ptr(), W.data_ptr(), None, o.data_ptr(), M, N, K, 0, 1))
            torch.cuda.synchronize()
            dt = (time.time() - t) / 5
            print(f'   nsplit={ns}: {dt * 1e3:.3f} ms incl. operand split -> {2 * M * N * K / dt / 1e12:.1f} TFLOP/s', flush=True)
        c.set_option('nsplit', 3)


def g_ln():
    import torch
    c = _ctx()
    for C_ in (128, 256, 512, 768, 1024, 2048):
        x = torch.randn(1000, C_) * 3 + 1
        g, b = torch.randn(C_), torch.randn(C_)
        ref = torch.nn.functional.layer_norm(x.double(), (C_,), g.double(), b.double(), 1e-5)
        xd, gd, bd = x.cuda(), g.cuda(), b.cuda()
        y = torch.empty_like(xd)
        c.check(c.lib.alm_op_layernorm(c.h, xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), 1e-5, y.data_ptr(), 1000, C_))
        print(f'layernorm C={C_}: {_err(y.cpu(), ref)}')


def g_wattn():
    import torch
    from oracle import omniparser_ref as O
    c = _ctx()
    for (B, nWh, nWw, heads, shift) in [(1, 1, 1, 4, 0), (2, 2, 3, 4, 3), (1, 3, 2, 8, 3)]:
        Cc = heads * 32
        g = torch.Generator().manual_seed(1)
        rows = B * nWh * nWw * 49
        qkv = torch.randn(rows, 3 * Cc, generator=g)
        tab = torch.randn(169, heads, generator=g) * 0.5
        # oracle core: same arithmetic as swin_transformer.py:127-148 without the projections
        q, k, v = qkv.view(-1, 49, 3, heads, 32).permute(2, 0, 3, 1, 4)
        attn = (q * 32 ** -0.5) @ k.transpose(-2, -1)
        from oracle.weights import relative_position_index
        bias = tab[relative_position_index().view(-1)].view(49, 49, -1).permute(2, 0, 1)
        attn = attn + bias.unsqueeze(0)
        if shift:
            mask = O.shift_mask(nWh * 7, nWw * 7)
            attn = attn.view(B, nWh * nWw, heads, 49, 49) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, heads, 49, 49)
        ref = (attn.softmax(-1) @ v).transpose(1, 2).reshape(rows, Cc)
        qd, td = qkv.cuda(), tab.cuda()
        for impl in (0, 1):
            c.set_option('wattn_impl', impl)
            out = torch.full((rows, Cc), float('nan'), device='cuda')
            c.check(c.lib.alm_op_window_attention(c.h, qd.data_ptr(), td.data_ptr(), out.data_ptr(), B, nWh, nWw, Cc, heads, shift))
            print(f'window_attention impl={impl} B={B} nW={nWh}x{nWw} heads={heads} shift={shift}: {_err(out.cpu(), ref)}')
    # timing at the stage-0 / stage-2 sizes of config 2 (batch 16, 1024^2)
    for (B, nW, heads) in [(16, 37, 4), (16, 10, 16)]:
        Cc = heads * 32
        rows = B * nW * nW * 49
        qd = torch.randn(rows, 3 * Cc, device='cuda')
        td = torch.randn(169, heads, device='cuda')
        out = torch.empty(rows, Cc, device='cuda')
        for impl in (0, 1):
            c.set_option('wattn_impl', impl)
            for _ in range(2):
                c.check(c.lib.alm_op_window_attention(c.h, qd.data_ptr(), td.data_ptr(), out.data_ptr(), B, nW, nW, Cc, heads, 3))
            torch.cuda.synchronize(); t0 = time.time()
            for _ in range(5):
                c.check(c.lib.alm_op_window_attention(c.h, qd.data_ptr(), td.data_ptr(), out.data_ptr(), B, nW, nW, Cc, heads, 3))
            torch.cuda.synchronize()
            print(f'window_attention timing impl={impl} B={B} nW={nW}^2 heads={heads}: {(time.time() - t0) / 5 * 1e3:.3f} ms', flush=True)


def _enc(impl):
    import numpy as np
    import torch
    from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab
    from oracle import omniparser_ref as O
    from oracle import weights as W
    from oracle.gen_golden import OMNI_CASES, omni_inputs
    torch.set_grad_enabled(False)
    sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0)
    m = OmniParserB200(sd, OmniVocab(pt_seq_length=8))
    m.ctx.set_option('gemm_impl', impl)
    for name in ('full', 'masked'):
        case = OMNI_CASES[name]
        img, mask = omni_inputs(case)
        feats, raws = O.swin_backbone(img, sd, return_raw=True)
        mem, pos, kpm, hw = O.encode(img, mask, sd)
        m.encode(img, mask)
        for lvl in range(4):
            f = m.features(lvl).permute(0, 2, 3, 1)
            print(f'[{name}] impl={impl} feature{lvl}: {_err(f, feats[lvl])}')
        print(f'[{name}] impl={impl} memory  : {_err(m.memory(0), mem)}')
        print(f'[{name}] impl={impl} pos     : {_err(m.memory(1), pos)}', flush=True)
    return m, sd


def g_enc_simt():
    _enc(1)


def g_enc():
    _enc(0)


def g_dec():
    import numpy as np
    import torch
    from oracle import omniparser_ref as O
    from oracle.gen_golden import OMNI_CASES, omni_inputs
    m, sd = _enc(0)
    for name in ('full', 'masked'):
        case = OMNI_CASES[name]
        gold = np.load(os.path.join(REPO, 'tests', 'golden', f'omni_{name}.npz'))
        img, mask = omni_inputs(case)
        m.vocab.pt_seq_length = case['pt_seq_length']
        m.encode(img, mask)
        pt = torch.from_numpy(gold['pt'])
        n = pt.numel() // 2
        pt_full = torch.cat([m.vocab.pt_prompt(), pt], dim=1)
        lg = m.decode_logits(0, 'pt', pt_full)
        print(f'[{name}] teacher-forced pt logits: {_err(lg[0, 6:], torch.from_numpy(gold["tf_pt"]))}')
        poly_full = torch.cat([pt.reshape(-1, 2), torch.full((n, 1), 1101), torch.from_numpy(gold['poly']).reshape(n, 32)], 1)
        lg = m.decode_logits(0, 'poly', poly_full)
        print(f'[{name}] teacher-forced poly logits: {_err(lg[:, [2, 17, 33]], torch.from_numpy(gold["tf_poly"]))}')
        rec_full = torch.cat([pt.reshape(-1, 2), torch.full((n, 1), 1102), torch.from_numpy(gold['rec'])[0]], 1)
        lg = m.decode_logits(0, 'rec', rec_full)
        print(f'[{name}] teacher-forced rec logits: {_err(lg[:, [2, 14, 26]], torch.from_numpy(gold["tf_rec"]))}')
        out = m.decode()[0]
        if out is None:
            print(f'[{name}] decode -> None (gold none={gold["none"][0]})')
            continue
        (p, po, r), (pr,) = out
        print(f'[{name}] greedy pt match={np.array_equal(p.numpy(), gold["pt"])} poly match={np.array_equal(po.numpy(), gold["poly"])} '
              f'rec match={np.array_equal(r.numpy(), gold["rec"])} probs: {_err(pr, torch.from_numpy(gold["probs"]))}', flush=True)
        if not np.array_equal(p.numpy(), gold['pt']):
            print('   pt  got ', p.numpy().tolist(), '\n   pt  gold', gold['pt'].tolist())


def g_perf():
    import torch
    from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab
    from oracle import weights as W
    sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0)
    for ns in (3, 1):
        m = OmniParserB200(sd, OmniVocab(pt_seq_length=16))
        m.ctx.set_option('nsplit', ns)
        B = 4
        img = torch.randn(B, 3, 1024, 1024, device='cuda')
        for it in range(3):
            torch.cuda.synchronize()
            t = time.time()
            m.encode(img, None)
            m.ctx.lib.alm_omni_vocab(m.ctx.h)
            m.memory_shape()
            mem = m.memory(0)
            t1 = time.time()
            outs = m.decode()
            t2 = time.time()
            print(f'nsplit={ns} B={B} 1024^2: encode {1e3 * (t1 - t):.1f} ms  decode(8 inst) {1e3 * (t2 - t1):.1f} ms  '
                  f'launches={m.ctx.launch_count(True)} mem_absmax={mem.abs().max().item():.3f}', flush=True)
        del m


def g_mgp():
    import numpy as np
    import torch
    from advancedliteratemachinery_b200 import MGPSTRB200
    from advancedliteratemachinery_b200 import synthetic as W
    from oracle import mgpstr_ref as M
    torch.set_grad_enabled(False)
    sd = W.mgpstr_state_dict(seed=0)
    m = MGPSTRB200(sd)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(3, 3, 32, 128, generator=g)
    ref = M.forward(img, sd)
    out = m(img, is_eval=True)
    for i, nm in enumerate(('char', 'bpe', 'wp')):
        print(f'mgp attn[{nm}]: {_err(out[0][i], ref[0][i])}')
        print(f'mgp logits[{nm}]: {_err(out[1 + i], ref[1 + i])}  ids match={torch.equal(m.last_ids[i].long(), ref[1 + i].argmax(-1))}')
    for B in (64, 512):
        img = torch.rand(B, 3, 32, 128, device='cuda')
        for ns in (3, 1):
            m.ctx.set_option('nsplit', ns)
            for it in range(3):
                torch.cuda.synchronize(); t = time.time()
                m.forward(img, is_eval=True, want_logits=False)
                dt = time.time() - t
            print(f'mgp B={B} nsplit={ns}: {dt * 1e3:.1f} ms -> {B / dt:.0f} crops/s ({49.75e9 * B / dt / 1e12:.1f} algorithmic TFLOP/s)', flush=True)
        m.ctx.set_option('nsplit', 3)


def g_trace():
    """In-kernel timeline of the GEMMs inside one bench step (B=16, N=64): durations by shape + gaps."""
    import collections
    import numpy as np
    import torch
    from advancedliteratemachinery_b200 import OmniParserB200, OmniVocab
    from advancedliteratemachinery_b200 import synthetic as W
    sd = W.omniparser_state_dict(seed=0, pt_eos_bias=-30.0)
    m = OmniParserB200(sd, OmniVocab(pt_seq_length=128))
    img = torch.randn(16, 3, 1024, 1024, device='cuda')
    for it in range(2):
        m.encode(img, None); m.decode()
    m.ctx.set_option('trace_gemm', 20000)
    torch.cuda.synchronize(); t0 = time.time()
    m.encode(img, None)
    m.memory_shape(); rec_enc = m.ctx.trace_read()
    t1 = time.time()
    m.decode()
    t2 = time.time()
    rec = m.ctx.trace_read()
    print(f'encode {1e3 * (t1 - t0):.1f} ms ({len(rec_enc)} gemms)  decode {1e3 * (t2 - t1):.1f} ms ({len(rec)} gemms)')
    for name, r in (('encode', rec_enc), ('decode', rec)):
        if len(r) < 2:
            continue
        r = r[np.argsort(r[:, 0])]
        dur = (r[:, 1] - r[:, 0]).astype(np.float64) / 1e3
        gap = (r[1:, 0] - r[:-1, 1]).astype(np.float64) / 1e3
        span = (r[-1, 1] - r[0, 0]) / 1e6
        print(f'[{name}] span {span:.2f} ms, sum(gemm) {dur.sum() / 1e3:.2f} ms, sum(gaps) {gap.sum() / 1e3:.2f} ms')
        agg = collections.defaultdict(list)
        for i in range(len(r)):
            agg[(int(r[i, 2]), int(r[i, 3]), int(r[i, 4]), int(r[i, 5]))].append(dur[i])
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:28]:
            print(f'   M={k[0]:8d} N={k[1]:6d} K={k[2]:5d} tiles={k[3] // 1000:6d} BN={k[3] % 1000:3d} n={len(v):5d} avg={np.mean(v):8.2f} us total={sum(v) / 1e3:8.2f} ms')
        gs = np.sort(gap)
        print(f'   gaps: median {np.median(gap):.2f} us  p90 {gs[int(0.9 * len(gs))]:.2f} us  max {gs[-1]:.1f} us')


def g_detail():
    """Per-role timeline of CTA 0 for a few GEMM shapes (where does a tile's time go?)."""
    import numpy as np
    c = _ctx()
    c.set_option('trace_detail', 1)
    for (M, N, K, batch, split, act, tag) in [(64, 4096, 64, 128, 0, 0, 'decode scores'), (16, 512, 512, 1, 0, 0, 'pt linear'),
                                              (16, 512, 2048, 1, 0, 0, 'pt ffn2'), (65536, 2048, 512, 1, 1, 1, 'stage2 fc1 gelu split'),
                                              (65536, 512, 2048, 1, 0, 0, 'stage2 fc2'), (1048576, 128, 128, 1, 0, 0, 'stage0 proj-like'),
                                              (1048576, 512, 128, 1, 1, 1, 'stage0 fc1')]:
        ms, d = c.bench_gemm_ex(M, N, K, batch, split, act, iters=5, detail=True)
        d = d.astype(np.int64)
        t0 = d[0, 0]
        print(f'--- {tag}: M={M} N={N} K={K} batch={batch} split={split} act={act}: {ms * 1e3:.1f} us/launch')
        print('   tile:  tma_issue  mma_start  operands   committed  epi_start  epi_end   (us since first TMA issue)')
        for t in range(8):
            if d[t, 0] == 0:
                break
            print('   %3d: ' % t + ' '.join('%9.2f' % ((d[t, j] - t0) / 1e3) for j in range(6)))


GROUPS = {'detail': g_detail, 'trace': g_trace, 'mgp': g_mgp, 'simt': g_simt, 'tc': g_tc, 'tcbig': g_tcbig, 'ln': g_ln, 'wattn': g_wattn, 'enc_simt': g_enc_simt,
          'enc': g_enc, 'dec': g_dec, 'perf': g_perf}

if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--run':
        GROUPS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(GROUPS)
    for n in names:
        print(f'===== {n} =====', flush=True)
        t = time.time()
        try:
            r = subprocess.run(['timeout', '600', sys.executable, os.path.abspath(__file__), '--run', n], cwd=REPO,
                               capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f'[group {n} exit {r.returncode}]\n' + r.stderr[-3000:])
        except Exception as e:  # noqa
            print('group failed to run:', e)
        print(f'----- {n} done in {time.time() - t:.1f}s', flush=True)
