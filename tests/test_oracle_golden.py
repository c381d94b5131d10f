"""CPU: the restated oracle reproduces the fixtures written from the UNMODIFIED reference
(oracle/gen_golden.py).  This is what pins the oracle on a box without /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import mgpstr_ref as M
from oracle import omniparser_ref as O
from oracle.gen_golden import MGP_CASES, OMNI_CASES, omni_inputs
from tests.conftest import mgp_sd, omni_sd

torch.set_grad_enabled(False)


@pytest.mark.parametrize('name', ['full', 'masked', 'odd', 'eos', 'oddlen', 'empty'])
def test_omniparser_oracle_matches_reference_fixture(name, golden_dir):
    case = OMNI_CASES[name]
    gold = np.load(os.path.join(golden_dir, f'omni_{name}.npz'))
    sd = omni_sd(case['wseed'], case['pt_eos_bias'])
    img, mask = omni_inputs(case)
    feats = O.swin_backbone(img, sd)
    for lvl in range(4):
        np.testing.assert_allclose(feats[lvl].reshape(-1)[::7].numpy(), gold[f'feat{lvl}_s'], atol=3e-5, rtol=0)
    mem, pos, kpm, (h, w) = O.encode(img, mask, sd)
    assert (h, w) == tuple(gold['hw'])
    np.testing.assert_allclose(mem[0].numpy(), gold['memory'], atol=3e-5, rtol=0)
    np.testing.assert_allclose(pos[0].numpy(), gold['pos'], atol=2e-6, rtol=0)
    assert np.array_equal(kpm[0].numpy(), gold['kpm'])
    pt_prompt, _, _ = O.default_prompts(True)
    res = O.greedy_text_spotting(mem[0], kpm[0], pos[0], sd, pt_prompt, case['pt_seq_length'], case['rec_length'])
    if gold['none'][0]:
        assert res is None
        return
    (pt, poly, rec), (probs,) = res
    assert np.array_equal(pt.numpy(), gold['pt'])
    assert np.array_equal(poly.numpy(), gold['poly'])
    assert np.array_equal(rec.numpy(), gold['rec'])
    np.testing.assert_allclose(probs.numpy(), gold['probs'], atol=1e-5, rtol=0)
    texts, confs = O.decode_rec_strings(rec[0], probs)
    assert texts == gold['texts'].tolist()
    np.testing.assert_allclose(np.array(confs), gold['confs'], rtol=1e-4)


@pytest.mark.parametrize('name', ['b1', 'b3'])
def test_mgpstr_oracle_matches_reference_fixture(name, golden_dir):
    case = MGP_CASES[name]
    gold = np.load(os.path.join(golden_dir, f'mgp_{name}.npz'))
    g = torch.Generator().manual_seed(case['seed'])
    img = torch.rand(case['batch'], 3, 32, 128, generator=g)
    attns, char, bpe, wp = M.forward(img, mgp_sd(case['wseed']))
    np.testing.assert_allclose(char.numpy(), gold['char'], atol=3e-5, rtol=0)
    np.testing.assert_allclose(attns[0].numpy(), gold['char_attn'], atol=1e-6, rtol=0)
    for nm, lg in (('bpe', bpe), ('wp', wp)):
        assert np.array_equal(lg.argmax(-1).numpy(), gold[nm + '_ids'])
        np.testing.assert_allclose(lg.max(-1)[0].numpy(), gold[nm + '_max'], atol=3e-5, rtol=0)
        np.testing.assert_allclose(lg.reshape(-1)[::997].numpy(), gold[nm + '_s'], atol=3e-5, rtol=0)


def test_vocab_layout():
    """Token-id layout of OCR/OmniParser/utils/parser.py:91-103."""
    from oracle import weights as W
    assert (W.RECOG_PAD, W.PT_EOS, W.POLY_EOS, W.REC_EOS, W.PT_SOS, W.POLY_SOS, W.REC_SOS, W.PADDING) == \
           (1096, 1097, 1098, 1099, 1100, 1101, 1102, 1103)
    assert len(W.CHARS) == 95


def test_kie_oracle_matches_reference_fixture(golden_dir):
    """decode_vie_pt_poly_rec_seq restatement vs the fixture written from the reference (SROIE classes)."""
    from advancedliteratemachinery_b200.omniparser import CLASSES_SROIE
    from advancedliteratemachinery_b200 import synthetic as W
    from oracle.gen_golden import KIE_CASES
    case = KIE_CASES['kie']
    gold = np.load(os.path.join(golden_dir, 'omni_kie.npz'))
    sd = W.omniparser_state_dict(seed=case['wseed'], vie_categories=case['vie'], pt_eos_bias=case['pt_eos_bias'])
    img, mask = omni_inputs(case)
    mem, pos, kpm, _ = O.encode(img, mask, sd)
    res, (pt_seq, pt_probs) = O.greedy_kie(mem[0], kpm[0], pos[0], sd, O.default_prompts(True)[0], case['pt_seq_length'],
                                           case['rec_length'], case['vie'], torch.tensor(case['canvas']), CLASSES_SROIE)
    assert [r[0] for r in res] == gold['texts'].tolist()
    assert [r[1] for r in res] == gold['classes'].tolist()
    assert np.array_equal(pt_seq.numpy(), gold['pt_seq'])
    np.testing.assert_allclose(np.array([r[3] for r in res]), gold['rects'])


@pytest.mark.parametrize('name', ['tiny', 'small', 'large', 'charstr'])
def test_mgpstr_variant_oracle_matches_reference_fixture(name, golden_dir):
    """tiny / small / large MGP-STR (mgp_str.py:176-230) and the char-only CHAR-STR (char_str.py:43-81): restatement vs
    fixtures written from the reference classes."""
    from oracle import weights as W
    from oracle.gen_golden import MGP_VARIANT_CASES
    case = MGP_VARIANT_CASES[name]
    gold = np.load(os.path.join(golden_dir, f'mgp_{name}.npz'))
    char_only = name == 'charstr'
    dim, depth, heads = W.MGP_VARIANTS['base' if char_only else name]
    assert (dim, depth, heads) == tuple(gold['dims'])
    sd = W.mgpstr_state_dict(seed=case['seed'], dim=dim, depth=depth, heads=heads, char_only=char_only)
    g = torch.Generator().manual_seed(case['seed'])
    img = torch.rand(2, 3, 32, 128, generator=g)
    out = M.forward(img, sd, depth=depth, heads=heads)
    np.testing.assert_allclose(out[1].numpy(), gold['char'], atol=3e-5, rtol=0)
    np.testing.assert_allclose(out[0][0].numpy(), gold['char_attn'], atol=1e-6, rtol=0)
    if not char_only:
        for nm, lg in (('bpe', out[2]), ('wp', out[3])):
            assert np.array_equal(lg.argmax(-1).numpy(), gold[nm + '_ids'])
            np.testing.assert_allclose(lg.reshape(-1)[::997].numpy(), gold[nm + '_s'], atol=3e-5, rtol=0)


def test_restated_timm_vit_matches_the_independent_hf_port(golden_dir):
    """The pin of the timm-0.4.12 layer: oracle/mgpstr_ref.py (and with it oracle/shim/timm) against
    HuggingFace transformers' MGP-STR port built with the reference's LayerNorm epsilons (oracle/pin_timm_hf.py), on the
    synthetic base checkpoint and on the tiny geometry; plus the committed HF fixture."""
    pytest.importorskip('transformers')
    from oracle import pin_timm_hf as P
    from oracle import weights as W
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 32, 128, generator=g)
    d, out = P.cross_check(mgp_sd(0), img, W.VIT_DIM, W.VIT_DEPTH, W.VIT_HEADS)
    assert max(d.values()) < 2e-5, d
    gold = np.load(os.path.join(golden_dir, 'mgp_hf_b2.npz'))
    np.testing.assert_allclose(out.logits[0].numpy(), gold['char'], atol=3e-5, rtol=0)
    assert np.array_equal(out.logits[1].argmax(-1).numpy(), gold['bpe_ids'])
    dim, depth, heads = W.MGP_VARIANTS['tiny']
    d2, _ = P.cross_check(W.mgpstr_state_dict(seed=21, dim=dim, depth=depth, heads=heads), img, dim, depth, heads)
    assert max(d2.values()) < 2e-5, d2
