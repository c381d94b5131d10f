"""Post-processing parity (host-only C ABI entry points; runs without a GPU).

The fixtures hold outputs of the UNMODIFIED reference functions (oracle/gen_golden_post.py):
`decode_pred_seq` for OmniParser (structured results and the `json.dumps(results, indent=4)` text) and
`TokenLabelConverter.{char,bpe,wp}_decode` + the restated fusion block for MGP-STR.  Bar: exact equality.
"""
import json
import os
import types

import numpy as np
import pytest
import torch

from advancedliteratemachinery_b200 import postprocess as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _args(case):
    a = types.SimpleNamespace(chars=case['chars'], num_bins=1000, rec_length=case['rec_length'])
    a.recog_pad_index = a.num_bins + len(a.chars) + 1
    a.rec_eos_index = a.recog_pad_index + 3
    return a


def _omni_cases():
    return json.load(open(os.path.join(GOLD, 'post_omni.json')))


@pytest.mark.parametrize('i', range(6))
def test_decode_pred_seq_equals_reference(i):
    c = _omni_cases()[i]
    a = _args(c)
    prob = torch.from_numpy(np.frombuffer(bytes.fromhex(c['prob_f32_hex']), dtype='<f4').copy()).reshape(-1, a.rec_length)
    seqs = [torch.tensor(c['pt']), torch.tensor(c['poly']), torch.tensor(c['rec'])]
    target = {'file_name': c['file_name'], 'orig_size': torch.tensor(c['orig'])}
    got = P.decode_pred_seq(seqs, [prob], target, a)
    assert got == c['results']                       # floats compare exactly: same float32 / double arithmetic
    assert P.results_json(seqs, [prob], target, a) == c['json']   # byte-identical JSON text


def test_decode_pred_seq_matches_the_oracle_on_random_inputs():
    from oracle.postprocess_ref import omni_results
    g = torch.Generator().manual_seed(3)
    a = _args({'chars': _omni_cases()[0]['chars'], 'rec_length': 25})
    for _ in range(20):
        n = int(torch.randint(0, 9, (1,), generator=g))
        pt = torch.randint(0, 1000, (1, 2 * n), generator=g)
        poly = torch.randint(0, 1000, (1, 32 * n), generator=g)
        rec = torch.randint(1000, a.recog_pad_index + 1, (1, n, 25), generator=g)
        prob = torch.rand(n, 25, generator=g) ** 6     # many tiny probabilities -> exponent-notation floats
        orig = (int(torch.randint(1, 5000, (1,), generator=g)), int(torch.randint(1, 5000, (1,), generator=g)))
        ref = omni_results([pt, poly, rec], prob, 'x.jpg', orig, a)
        target = {'file_name': 'x.jpg', 'orig_size': torch.tensor(orig)}
        got = P.decode_pred_seq([pt, poly, rec], [prob], target, a)
        text = P.results_json([pt, poly, rec], [prob], target, a)
        assert json.dumps(json.loads(text), indent=4) == text      # exactly the text json.dumps would write for it
        assert json.loads(text) == got
        assert len(got) == len(ref)
        for x, y in zip(got, ref):
            # `score` is sum(probabilities) / (count + 1e-5) in double: the library adds left to right like the
            # interpreters the reference is pinned to (torch 1.7 => Python <= 3.8); Python >= 3.12 `sum()` is
            # compensated (Neumaier), so the oracle running here may differ in the last bit
            assert x['score'] == pytest.approx(y['score'], rel=1e-14, abs=0)
            assert {k: v for k, v in x.items() if k != 'score'} == {k: v for k, v in y.items() if k != 'score'}


def test_postprocess_error_paths():
    from advancedliteratemachinery_b200 import AlmError
    a = _args({'chars': 'ab', 'rec_length': 3})
    target = {'file_name': 'x', 'orig_size': torch.tensor([10, 10])}
    with pytest.raises(ValueError):                    # odd-length point sequence (the reference crashes on it)
        P.decode_pred_seq([torch.zeros(1, 3, dtype=torch.long), torch.zeros(1, 32, dtype=torch.long),
                           torch.zeros(1, 1, 3, dtype=torch.long)], [torch.zeros(1, 3)], target, a)
    bad = torch.full((1, 1, 3), 1000 + 7)              # id with no character: IndexError in the reference
    with pytest.raises(AlmError):
        P.decode_pred_seq([torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 32, dtype=torch.long), bad],
                          [torch.zeros(1, 3)], target, a)


def test_mgp_fusion_equals_reference():
    c = json.load(open(os.path.join(GOLD, 'post_mgp.json')))
    ids = torch.tensor(c['ids'])
    prob = torch.from_numpy(np.frombuffer(bytes.fromhex(c['prob_f32_hex']), dtype='<f4').copy()).reshape(3, c['B'], c['T'])
    bpe_table = [bytes.fromhex(h) for h in c['bpe_table_hex']]
    out = P.mgp_fuse(ids, prob, c['char_table'], bpe_table, c['wp_table'])
    for hd in range(3):
        assert out['texts'][hd] == c['texts'][hd], hd
        np.testing.assert_array_equal(out['conf'][hd], np.asarray(c['conf'][hd], dtype=np.float32))
    assert out['fused'] == c['fused']
    assert out['source'].tolist() == c['source']
    assert -1 in c['source'] and {0, 1, 2} <= set(c['source'])   # the fixture exercises every branch


def test_kie_walk_json_equals_reference():
    """alm_post_omni_kie_json == json.dumps of the reference's own decode_vie_pt_poly_rec_seq (run unbound with a scripted
    decode stub, oracle/gen_golden_post.py): pair detection, lone bins, polygon extents, transcriptions, class lookup,
    entity closing, an unfinished trailing entity, an entity without words, unicode characters."""
    cases = json.load(open(os.path.join(GOLD, 'post_kie.json')))
    assert len(cases) == 5
    seen_empty_entity = False
    for c in cases:
        a = types.SimpleNamespace(chars=c['chars'], num_bins=1000, rec_length=c['rec_length'])
        a.recog_pad_index = a.num_bins + len(a.chars) + 1
        a.rec_eos_index = a.recog_pad_index + 3
        probs = np.frombuffer(bytes.fromhex(c['probs_f32_hex']), dtype='<f4').copy()
        text = P.kie_json(c['tokens'], probs, c['pos'], np.asarray(c['poly'], dtype=np.int64).reshape(-1, 32),
                          np.asarray(c['rec'], dtype=np.int64).reshape(-1, a.rec_length), c['orig'], a, c['classes'],
                          c['class_base'])
        assert text == c['json']
        seen_empty_entity |= any(e[0] == '' and e[3] == [] for e in json.loads(text))
    assert seen_empty_entity


def test_kie_walk_rejects_inconsistent_input():
    from advancedliteratemachinery_b200 import AlmError
    a = types.SimpleNamespace(chars='ab', num_bins=1000, rec_length=2, recog_pad_index=1003, rec_eos_index=1006)
    with pytest.raises(AlmError):          # an (x, y) pair in the stream but no decoded instance for it
        P.kie_json([1, 2, 1011], np.ones(3, dtype=np.float32), [], np.zeros((0, 32)), np.zeros((0, 2)), (10, 10), a,
                   ['c0', 'c1', 'c2', 'c3'], 1011)
    with pytest.raises(AlmError):          # a token that is neither a bin nor a class (KeyError in the reference)
        P.kie_json([1005], np.ones(1, dtype=np.float32), [], np.zeros((0, 32)), np.zeros((0, 2)), (10, 10), a,
                   ['c0'], 1011)
