"""Post-processing fixtures from the UNMODIFIED reference -> tests/golden/post_omni.json, post_mgp.json.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Usage:
    python -m oracle.gen_golden_post omni
    python -m oracle.gen_golden_post kie
    python -m oracle.gen_golden_post mgp
(two interpreters: both reference sub-projects define a top-level `utils`).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, 'tests', 'golden')
REF = '/root/reference/OCR'
CHARS = ' !"#$%&\'()*+,-./0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz{|}~'


def omni_args(chars=CHARS, rec_length=25):
    a = types.SimpleNamespace(chars=chars, num_bins=1000, rec_length=rec_length)
    a.recog_pad_index = a.num_bins + len(chars) + 1          # utils/parser.py:91-96
    a.pt_eos_index = a.recog_pad_index + 1
    a.poly_eos_index = a.pt_eos_index + 1
    a.rec_eos_index = a.poly_eos_index + 1
    return a


def omni_case(seed, n, rec_length, orig, file_name, chars=CHARS, odd_pt=False):
    g = torch.Generator().manual_seed(seed)
    a = omni_args(chars, rec_length)
    pt = torch.randint(0, a.num_bins, (1, 2 * n + (1 if odd_pt else 0)), generator=g)
    poly = torch.randint(0, a.num_bins, (1, 32 * n), generator=g)
    rec = torch.randint(a.num_bins, a.recog_pad_index - 1, (1, n, rec_length), generator=g)
    for i in range(n):                                      # sprinkle unknown / pad / eos tokens
        r = torch.rand(3, generator=g)
        if r[0] < 0.5:
            rec[0, i, int(torch.randint(0, rec_length, (1,), generator=g))] = a.recog_pad_index - 1
        cut = int(torch.randint(0, rec_length + 1, (1,), generator=g))
        if cut < rec_length:
            rec[0, i, cut] = a.rec_eos_index if r[1] < 0.5 else a.recog_pad_index
    prob = torch.rand(n, rec_length, generator=g)
    prob[prob < 0.05] *= 1e-4                               # small values: exponent notation in the JSON
    return dict(pt=pt, poly=poly, rec=rec, prob=prob, orig=orig, file_name=file_name, args=a)


def gen_omni():
    sys.path[:0] = [os.path.join(REPO, 'oracle', 'shim'), os.path.join(REF, 'OmniParser')]
    sys.modules.setdefault('bezier', types.ModuleType('bezier'))   # utils/misc.py:2 (only the training-side fit uses it)
    from engine.val import decode_pred_seq                   # the reference function, unmodified
    from oracle.postprocess_ref import omni_results
    cases = [omni_case(1, 5, 25, (480, 640), 'img_0001.jpg'),
             omni_case(2, 3, 25, (750, 1333), 'dir/with "quotes"\\and\ttab.png'),
             omni_case(3, 64, 25, (1024, 1024), 'p\u00e4ge_\u4e2d\u6587_\U0001F600.jpg'),
             omni_case(4, 4, 7, (1, 3), 'tiny.jpg'),
             omni_case(5, 0, 25, (100, 100), 'empty.jpg'),
             omni_case(6, 6, 25, (2160, 3840), 'greek.jpg', chars=CHARS[:60] + '\u03b1\u03b2\u03b3\u00e9\u00fc' + CHARS[65:])]
    # An odd-length point sequence never reaches decode_pred_seq (transformer.py:138-139 strips the tail).  The
    # reference's own guard for it (val.py:73-74) reads `[:-len(seq) % 2]`, i.e. `[:1]`, and crashes in decode_seq:
    odd = omni_case(9, 3, 25, (480, 640), 'odd.jpg', odd_pt=True)
    try:
        decode_pred_seq([odd['pt'][0], odd['poly'][0], odd['rec'][0]], odd['prob'],
                        {'file_name': 'odd.jpg', 'orig_size': torch.tensor(odd['orig'])}, odd['args'])
        raise AssertionError('the reference was expected to fail on an odd-length point sequence')
    except RuntimeError:
        pass
    out = []
    for c in cases:
        target = {'file_name': c['file_name'], 'orig_size': torch.tensor(c['orig'])}
        res = decode_pred_seq([c['pt'][0], c['poly'][0], c['rec'][0]], c['prob'], target, c['args'])
        mine = omni_results([c['pt'], c['poly'], c['rec']], c['prob'], c['file_name'], c['orig'], c['args'])
        assert json.dumps(res) == json.dumps(mine), 'restatement differs from the reference'
        out.append({'pt': c['pt'].tolist(), 'poly': c['poly'].tolist(), 'rec': c['rec'].tolist(),
                    'prob_f32_hex': c['prob'].numpy().astype('<f4').tobytes().hex(), 'orig': list(c['orig']),
                    'file_name': c['file_name'], 'chars': c['args'].chars, 'rec_length': c['args'].rec_length,
                    'results': res, 'json': json.dumps(res, indent=4)})
    with open(os.path.join(GOLD, 'post_omni.json'), 'w') as f:
        json.dump(out, f)
    print('post_omni.json:', len(out), 'cases,', sum(len(o['results']) for o in out), 'instances')


def kie_case(seed, n_tok, rec_length, orig, chars=CHARS, vie=4):
    """A random KIE point-token stream (bins, lone bins, class tokens) + the polygon / transcription ids of every
    (x, y) pair in it."""
    g = torch.Generator().manual_seed(seed)
    a = omni_args(chars, rec_length)
    a.pt_sos_index = a.rec_eos_index + 1
    a.padding_index = a.pt_sos_index + 3                      # utils/parser.py:98-103
    a.vie_categories = vie
    base = a.padding_index + 1
    toks, i = [], 0
    while len(toks) < n_tok:
        r = float(torch.rand(1, generator=g))
        if r < 0.55 and len(toks) + 2 <= n_tok:
            toks += torch.randint(0, a.num_bins, (2,), generator=g).tolist()
        elif r < 0.7:
            toks.append(int(torch.randint(0, a.num_bins, (1,), generator=g)))
            toks.append(base + int(torch.randint(0, vie, (1,), generator=g)))   # lone bin, then a class token
        else:
            toks.append(base + int(torch.randint(0, vie, (1,), generator=g)))
    toks = toks[:n_tok]
    pos = []
    i = 0
    while i < len(toks):
        if toks[i] < a.num_bins and i + 1 <= len(toks) - 1 and toks[i + 1] < a.num_bins:
            pos.append(i)
            i += 2
        else:
            i += 1
    n = len(pos)
    poly = torch.randint(0, a.num_bins, (n, 32), generator=g)
    rec = torch.randint(a.num_bins, a.recog_pad_index - 1, (n, rec_length), generator=g)
    for k in range(n):
        cut = int(torch.randint(0, rec_length + 1, (1,), generator=g))
        if cut < rec_length:
            rec[k, cut] = a.rec_eos_index if k % 2 else a.recog_pad_index
        if k % 3 == 0:
            rec[k, int(torch.randint(0, rec_length, (1,), generator=g))] = a.recog_pad_index - 1
    probs = torch.rand(len(toks), generator=g)
    return dict(tokens=toks, probs=probs, pos=pos, poly=poly, rec=rec, orig=orig, args=a, base=base)


def gen_kie():
    """Pins the KIE entity walk against the reference's own decode_vie_pt_poly_rec_seq: the method is called unbound on
    a stub `self` whose `decode` is scripted to emit the prepared polygon / transcription ids, so every line of the
    reference walk (pair detection, extents, transcription, class lookup, entity closing) runs unmodified."""
    sys.path[:0] = [os.path.join(REPO, 'oracle', 'shim'), os.path.join(REF, 'OmniParser')]
    sys.modules.setdefault('bezier', types.ModuleType('bezier'))
    from model.transformer import Transformer                # the reference class, unmodified
    from oracle.postprocess_ref import kie_walk
    classes = ['company', 'date', 'address', 'total']         # SROIE (transformer.py:58-61)
    cases = [kie_case(21, 30, 25, (480, 640)), kie_case(22, 9, 25, (1000, 333)), kie_case(23, 64, 7, (2160, 3840)),
             kie_case(24, 1, 25, (10, 10)), kie_case(25, 40, 25, (777, 555), chars=CHARS[:50] + '\u00e9\u4e2d' + CHARS[52:])]
    out = []
    for c in cases:
        a = c['args']
        V = a.padding_index + 1 + a.vie_categories
        state = {'inst': -1}

        def decode(seq, memory, mask, pos_embed, kind, c=c, state=state, V=V):
            step = seq.shape[1] - 3
            if kind == 'poly' and step == 0:
                state['inst'] += 1
            target = int((c['poly'] if kind == 'poly' else c['rec'])[state['inst'], step])
            lg = torch.full((1, seq.shape[1], V), -30.0)
            lg[:, :, target] = 30.0
            return lg

        stub = types.SimpleNamespace(args=a, decode=decode,
                                     index2class={c['base'] + i: n for i, n in enumerate(classes)})
        pt_seq = torch.tensor(c['tokens'], dtype=torch.long)
        ref = Transformer.decode_vie_pt_poly_rec_seq(stub, pt_seq, c['probs'], torch.tensor([[a.rec_eos_index - 1]]),
                                                     torch.tensor([[a.rec_eos_index]]), torch.tensor(c['orig']), None, None, None)
        mine = kie_walk(c['tokens'], c['probs'], c['pos'], c['poly'], c['rec'], c['orig'], a, classes, c['base'])
        assert json.dumps(ref) == json.dumps(mine), 'restatement differs from the reference walk'
        out.append({'tokens': c['tokens'], 'probs_f32_hex': c['probs'].numpy().astype('<f4').tobytes().hex(), 'pos': c['pos'],
                    'poly': c['poly'].tolist(), 'rec': c['rec'].tolist(), 'orig': list(c['orig']), 'chars': a.chars,
                    'rec_length': a.rec_length, 'classes': classes, 'class_base': c['base'], 'json': json.dumps(ref)})
    with open(os.path.join(GOLD, 'post_kie.json'), 'w') as f:
        json.dump(out, f)
    print('post_kie.json:', len(out), 'cases,', sum(len(json.loads(o['json'])) for o in out), 'entities')


def bytes_to_unicode():
    """GPT-2's published byte <-> printable-unicode table (the tokenizer's vocab strings use it)."""
    bs = list(range(ord('!'), ord('~') + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def gen_mgp():
    sys.path[:0] = [os.path.join(REPO, 'oracle', 'shim'), os.path.join(REF, 'MGP-STR')]
    for name in ('strsimpy', 'strsimpy.normalized_levenshtein'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['strsimpy.normalized_levenshtein'].NormalizedLevenshtein = object
    from transformers import BertTokenizer, GPT2Tokenizer
    import utils as ref_utils                                # OCR/MGP-STR/utils.py, unmodified
    from oracle.postprocess_ref import mgp_fuse_ref
    d = tempfile.mkdtemp()
    # synthetic GPT-2 vocabulary: id 1 = '"' and id 2 = '#' like the real one (the reference hard-codes them)
    u = bytes_to_unicode()
    sp = u[ord(' ')]
    bpe_tokens = ['!', '"', '#', 'a', 'b', 'c', sp, sp + 'a', 'ab', '.', ',', sp + '.', "'s", sp + "'s", 'n', "'t",
                  sp + 'n', sp + "n't", 'the', sp + 'the', 'ing', sp + ',', sp + '!', '?', sp + '?', "'", sp + "'",
                  've', sp + "'ve", 'x', 'y', 'z', '0', '1', '2', u[0xC3] + u[0xA9], sp + 'z']
    json.dump({t: i for i, t in enumerate(bpe_tokens)}, open(d + '/vocab.json', 'w'))
    open(d + '/merges.txt', 'w').write('#version: 0.2\n')
    bpe_tok = GPT2Tokenizer(d + '/vocab.json', d + '/merges.txt')
    wp_tokens = ['[PAD]'] + ['[unused%d]' % i for i in range(99)] + ['[UNK]', '[CLS]', '[SEP]', '[MASK]', 'a', 'b', '##c',
                 '##ing', '.', ',', "'", 's', 'hello', '##s', 'x', '##y', '!', '?', 'n', "##'", 't', 'the', '##e', '1',
                 '##2']
    open(d + '/vocab.txt', 'w').write('\n'.join(wp_tokens))
    wp_tok = BertTokenizer(d + '/vocab.txt')
    for t in (bpe_tok, wp_tok):
        t.clean_up_tokenization_spaces = True                # transformers==4.2.1 (the reference's pin) default
    character = '0123456789abcdefghijklmnopqrstuvwxyz'
    conv = object.__new__(ref_utils.TokenLabelConverter)     # skip __init__: it downloads the real tokenizers
    conv.SPACE, conv.GO = '[s]', '[GO]'
    conv.list_token = [conv.GO, conv.SPACE]
    conv.character = conv.list_token + list(character)
    conv.batch_max_length = 25 + 2
    conv.bpe_tokenizer, conv.wp_tokenizer = bpe_tok, wp_tok
    T, B = 27, 48
    g = torch.Generator().manual_seed(11)
    sizes = [len(conv.character), len(bpe_tokens), len(wp_tokens)]
    ids = torch.stack([torch.randint(0, sizes[h], (B, T), generator=g) for h in range(3)])
    ids[2][ids[2] < 100] += 100                              # keep wp ids on real tokens mostly
    ids[2].clamp_(max=sizes[2] - 1)
    eos = [1, 2, 102]
    for h in range(3):                                       # EOS somewhere in most rows, none in a few
        for b in range(B):
            ids[h, b][ids[h, b] == eos[h]] = 3 if h < 2 else 104
            if b % 7 != 3:
                ids[h, b, int(torch.randint(1, T, (1,), generator=g))] = eos[h]
    ids[0, 5, 1:4] = 0                                       # '[GO]' tokens before '[s]': string index != token index
    ids[0, 5, 4] = 1
    prob = torch.rand(3, B, T, generator=g) * 0.6 + 0.4
    length = torch.IntTensor([T - 1] * B)
    strings = [conv.char_decode(ids[0][:, 1:], length), conv.bpe_decode(ids[1][:, 1:], length),
               conv.wp_decode(ids[2][:, 1:], length)]       # the reference's converters
    texts, conf, fused, source = mgp_fuse_ref(strings, ids, prob)
    bdec = {v: k for k, v in u.items()}
    bpe_table_hex = [bytes(bdec[ch] for ch in t).hex() for t in bpe_tokens]
    out = {'ids': ids.tolist(), 'prob_f32_hex': prob.numpy().astype('<f4').tobytes().hex(), 'B': B, 'T': T,
           'char_table': conv.character, 'bpe_table_hex': bpe_table_hex, 'wp_table': wp_tokens,
           'strings': strings, 'texts': texts, 'conf': conf, 'fused': fused, 'source': source}
    with open(os.path.join(GOLD, 'post_mgp.json'), 'w') as f:
        json.dump(out, f)
    print('post_mgp.json:', B, 'crops; sources', {s: source.count(s) for s in set(source)})


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which == 'all':
        import subprocess
        for w in ('omni', 'kie', 'mgp'):
            subprocess.check_call([sys.executable, '-m', 'oracle.gen_golden_post', w], cwd=REPO)
    elif which == 'omni':
        gen_omni()
    elif which == 'kie':
        gen_kie()
    else:
        gen_mgp()
