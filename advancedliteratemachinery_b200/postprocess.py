"""Sequence -> structured-result post-processing through the C ABI (host-only entry points of libalm_ocr.so).

Mirrors of the reference functions, same names and argument meaning:

    decode_pred_seq(index_seqs, prob_seqs, target, args)   OCR/OmniParser/engine/val.py:70-100
    results_json(...)                                       the text `json.dumps(results, indent=4)` writes (val.py:63-67)
    mgp_fuse(ids, prob, char_table, bpe_table, wp_table)   OCR/MGP-STR/test_final.py:176-240 (per-head text +
                                                            cumprod confidence + 3-way fusion)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib


def _check(rc):
    if rc != 0:
        raise _lib.AlmError(rc, (_lib.load().alm_post_last_error() or b'').decode())


def _seqs(index_seqs, prob_seqs, rec_length):
    pt = np.ascontiguousarray(torch.as_tensor(index_seqs[0]).reshape(-1).numpy(), dtype=np.int64)
    if pt.size % 2:
        # transformer.py:138-139 already strips an odd tail, so the model never returns one; the reference's own guard
        # (val.py:73-74, `[:-len(seq) % 2]` == `[:1]`) crashes on such input -- reject it with a clear message instead
        raise ValueError('odd-length point sequence: not a model output (the reference fails on it in decode_seq)')
    n = pt.size // 2
    poly = np.ascontiguousarray(torch.as_tensor(index_seqs[1]).reshape(-1).numpy(), dtype=np.int64)
    rec = np.ascontiguousarray(torch.as_tensor(index_seqs[2]).reshape(-1).numpy(), dtype=np.int64)
    prob = np.ascontiguousarray(torch.as_tensor(prob_seqs).reshape(-1).numpy(), dtype=np.float32)
    assert poly.size == 32 * n and rec.size == rec_length * n and prob.size == rec_length * n, 'sequence shapes'
    return pt, poly, rec, prob, n


def _orig(target):
    h, w = target['orig_size']
    return int(h), int(w)


def decode_pred_seq(index_seqs, prob_seqs, target, args) -> List[dict]:
    """Reference `decode_pred_seq`: ids -> [{'image_id', 'pts', 'score', 'polys', 'rec'}, ...] for one image.
    `args` needs num_bins, rec_length, recog_pad_index, rec_eos_index, chars (utils/parser.py:16-96)."""
    lib = _lib.load()
    pt, poly, rec, prob, n = _seqs(index_seqs, prob_seqs[0] if isinstance(prob_seqs, (list, tuple)) else prob_seqs,
                                   args.rec_length)
    h, w = _orig(target)
    pts = np.zeros((n, 2)); polys = np.zeros((n, 32)); scores = np.zeros(n)
    stride = 4 * args.rec_length + 8
    texts = C.create_string_buffer(max(1, n) * stride)
    _check(lib.alm_post_omni_spotting(pt.ctypes.data, poly.ctypes.data, rec.ctypes.data, prob.ctypes.data, n,
                                      args.rec_length, args.num_bins, args.recog_pad_index, args.rec_eos_index,
                                      args.chars.encode('utf-8'), h, w, pts.ctypes.data, polys.ctypes.data,
                                      scores.ctypes.data, C.addressof(texts), stride))
    out = []
    for i in range(n):
        out.append({'image_id': target['file_name'], 'pts': [[float(pts[i, 0]), float(pts[i, 1])]],
                    'score': float(scores[i]), 'polys': polys[i].reshape(-1, 2).tolist(),
                    'rec': C.string_at(C.addressof(texts) + i * stride).decode('utf-8')})
    return out


def results_json(index_seqs, prob_seqs, target, args) -> str:
    """`json.dumps(decode_pred_seq(...), indent=4)` produced by the library (byte-identical text)."""
    lib = _lib.load()
    pt, poly, rec, prob, n = _seqs(index_seqs, prob_seqs[0] if isinstance(prob_seqs, (list, tuple)) else prob_seqs,
                                   args.rec_length)
    h, w = _orig(target)
    need = C.c_size_t(0)
    cap = 4096 + n * 4096
    buf = C.create_string_buffer(cap)
    rc = lib.alm_post_omni_json(pt.ctypes.data, poly.ctypes.data, rec.ctypes.data, prob.ctypes.data, n, args.rec_length,
                                args.num_bins, args.recog_pad_index, args.rec_eos_index, args.chars.encode('utf-8'), h, w,
                                str(target['file_name']).encode('utf-8'), C.addressof(buf), cap, C.byref(need))
    if rc != 0 and need.value > cap:
        cap = need.value
        buf = C.create_string_buffer(cap)
        rc = lib.alm_post_omni_json(pt.ctypes.data, poly.ctypes.data, rec.ctypes.data, prob.ctypes.data, n,
                                    args.rec_length, args.num_bins, args.recog_pad_index, args.rec_eos_index,
                                    args.chars.encode('utf-8'), h, w, str(target['file_name']).encode('utf-8'),
                                    C.addressof(buf), cap, C.byref(need))
    _check(rc)
    return buf.value.decode('utf-8')


def kie_json(tokens, probs, inst_pos, poly, rec, image_size, args, classes: Sequence[str], class_base: int) -> str:
    """`json.dumps(output)` of the reference's KIE result for one image (model/transformer.py:148-215,
    engine/val.py:38-42) from the raw outputs of `alm_omni_decode_kie` (`OmniParserB200.last_kie_raw`, one image):
    tokens / probs [n_tok], inst_pos [n_inst], poly [n_inst, 32], rec [n_inst, rec_length]; image_size = (h, w)."""
    lib = _lib.load()
    tokens = np.ascontiguousarray(tokens, dtype=np.int64)
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    inst_pos = np.ascontiguousarray(inst_pos, dtype=np.int32)
    poly = np.ascontiguousarray(poly, dtype=np.int64)
    rec = np.ascontiguousarray(rec, dtype=np.int64)
    table = _table(classes)
    need = C.c_size_t(0)
    cap = 1 << 16
    for _ in range(2):
        buf = C.create_string_buffer(cap)
        rc = lib.alm_post_omni_kie_json(tokens.ctypes.data, probs.ctypes.data, tokens.size, inst_pos.ctypes.data, inst_pos.size,
                                        poly.ctypes.data, rec.ctypes.data, args.rec_length, args.num_bins, args.recog_pad_index,
                                        args.rec_eos_index, args.chars.encode('utf-8'), table, len(classes), class_base,
                                        int(image_size[0]), int(image_size[1]), C.addressof(buf), cap, C.byref(need))
        if rc == 0 or need.value <= cap:
            break
        cap = need.value
    _check(rc)
    return buf.value.decode('utf-8')


def _table(tokens: Sequence) -> "C.Array":
    enc = [t if isinstance(t, bytes) else str(t).encode('utf-8') for t in tokens]
    return (C.c_char_p * len(enc))(*enc)


def mgp_fuse(ids, prob, char_table: Sequence, bpe_table: Sequence, wp_table: Sequence):
    """ids / prob: [3, B, T] top-1 ids and max-softmax probabilities of the char / bpe / wp heads incl. position 0
    (`MGPSTRB200.last_ids`, `.last_prob`).  Tables: token strings per id (char: ['[GO]', '[s]'] + opt.character;
    bpe: byte-decoded GPT-2 tokens (bytes allowed); wp: WordPiece tokens).
    Returns dict(texts=[3][B] str, conf=f32 [3,B], fused=[B] str, source=int32 [B] (0 char / 1 bpe / 2 wp / -1 none))."""
    lib = _lib.load()
    ids = np.ascontiguousarray(torch.as_tensor(ids).numpy(), dtype=np.int32)
    prob = np.ascontiguousarray(torch.as_tensor(prob).numpy(), dtype=np.float32)
    assert ids.ndim == 3 and ids.shape[0] == 3 and prob.shape == ids.shape
    _, B, T = ids.shape
    ct, bt, wt = _table(char_table), _table(bpe_table), _table(wp_table)
    stride = 64 * T
    texts = C.create_string_buffer(max(1, 3 * B) * stride)
    fused = C.create_string_buffer(max(1, B) * stride)
    conf = np.zeros((3, B), dtype=np.float32)
    source = np.zeros(B, dtype=np.int32)
    _check(lib.alm_post_mgp_fuse(ids.ctypes.data, prob.ctypes.data, B, T, ct, len(char_table), bt, len(bpe_table), wt,
                                 len(wp_table), C.addressof(texts), C.addressof(fused), stride, conf.ctypes.data,
                                 source.ctypes.data))
    get = lambda base, i: C.string_at(C.addressof(base) + i * stride).decode('utf-8', errors='replace')
    return {'texts': [[get(texts, hd * B + b) for b in range(B)] for hd in range(3)], 'conf': conf,
            'fused': [get(fused, b) for b in range(B)], 'source': source}
