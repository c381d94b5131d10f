"""Deterministic synthetic checkpoints in the reference's two state-dict layouts.

The reference ships no weights and there is no network, so the benchmark and the parity tests run on
synthetic checkpoints (data only: no model arithmetic lives here).  The key set / shapes are those of

  * OmniParser  : ``torch.load(path)['model']`` (OCR/OmniParser/utils/checkpointer.py:20,44-47);
                  610 tensors, 144.15 M parameters, keys as listed in SURVEY.md section 8b.
  * MGP-STR     : flat state_dict with the ``module.mgp_str.`` prefix
                  (OCR/MGP-STR/test_final.py:348,356; models.py:36).

The golden-fixture generator of the test suite (run in the build container only) proves the key sets are exact by
``load_state_dict(strict=True)`` into the unmodified reference modules.

Values are drawn from a seeded CPU ``torch.Generator`` so the GPU box regenerates the very
same tensors without reading /root/reference.  Every parameter is randomised (LayerNorm affine
terms and biases included) so that a kernel that drops one of them fails parity.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

# ---- OmniParser vocabulary (OCR/OmniParser/utils/parser.py:16,91-103) -------------------------
NUM_BINS = 1000
CHARS = ' !"#$%&\'()*+,-./0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz{|}~'
RECOG_PAD = NUM_BINS + len(CHARS) + 1  # 1096
PT_EOS, POLY_EOS, REC_EOS = RECOG_PAD + 1, RECOG_PAD + 2, RECOG_PAD + 3  # 1097..1099
PT_SOS, POLY_SOS, REC_SOS = REC_EOS + 1, REC_EOS + 2, REC_EOS + 3  # 1100..1102
PADDING = REC_SOS + 1  # 1103

SWIN_EMBED = 128
SWIN_DEPTHS = (2, 2, 18, 2)
SWIN_HEADS = (4, 8, 16, 32)
WINDOW = 7
D_MODEL = 512
DEC_LAYERS = 4
DEC_FFN = 2048
MAX_POS = 1024


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator(device='cpu')
        self.g.manual_seed(seed)

    def uniform(self, shape, bound):
        return (torch.rand(shape, generator=self.g, dtype=torch.float32) * 2 - 1) * bound

    def normal(self, shape, std):
        return torch.randn(shape, generator=self.g, dtype=torch.float32) * std


def _linear(sd, g, name, out_f, in_f, bias=True, shape=None, gain=1.0):
    b = gain / math.sqrt(in_f)
    sd[name + '.weight'] = g.uniform(shape or (out_f, in_f), b)
    if bias:
        sd[name + '.bias'] = g.uniform((out_f,), b)


def _ln(sd, g, name, dim):
    sd[name + '.weight'] = 1.0 + g.uniform((dim,), 0.1)
    sd[name + '.bias'] = g.uniform((dim,), 0.1)


def relative_position_index(ws: int = WINDOW) -> torch.Tensor:
    """index[i, j] = (dy + ws-1) * (2 ws - 1) + (dx + ws-1)  (swin_transformer.py:98-109)."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing='ij')).flatten(1)
    rel = coords[:, :, None] - coords[:, None, :]
    return ((rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)).contiguous()


def omniparser_state_dict(seed: int = 0, vie_categories: int = 0, pt_eos_bias: float = 0.0):
    """Synthetic ``['model']`` dict for the shipped inference graph (--tfm_pre_norm --use_fpn).

    ``pt_eos_bias`` is added to the pt head's final bias at ``pt_eos`` -- a strongly negative
    value pins the number of decoded points at pt_seq_length/2,
    a positive one makes an early EOS likely (exercises the stop path).
    """
    g = _Gen(seed)
    sd = OrderedDict()
    p = 'backbone.0.'
    _linear(sd, g, p + 'patch_embed.proj', SWIN_EMBED, 48, shape=(SWIN_EMBED, 3, 4, 4))
    _ln(sd, g, p + 'patch_embed.norm', SWIN_EMBED)
    for s, (depth, heads) in enumerate(zip(SWIN_DEPTHS, SWIN_HEADS)):
        c = SWIN_EMBED << s
        for b in range(depth):
            q = f'{p}layers.{s}.blocks.{b}.'
            _ln(sd, g, q + 'norm1', c)
            sd[q + 'attn.relative_position_bias_table'] = g.normal(((2 * WINDOW - 1) ** 2, heads), 0.3)
            sd[q + 'attn.relative_position_index'] = relative_position_index()
            _linear(sd, g, q + 'attn.qkv', 3 * c, c)
            _linear(sd, g, q + 'attn.proj', c, c)
            _ln(sd, g, q + 'norm2', c)
            _linear(sd, g, q + 'mlp.fc1', 4 * c, c)
            _linear(sd, g, q + 'mlp.fc2', c, 4 * c)
        if s < 3:
            q = f'{p}layers.{s}.downsample.'
            _linear(sd, g, q + 'reduction', 2 * c, 4 * c, bias=False)
            _ln(sd, g, q + 'norm', 4 * c)
    for s in range(4):
        _ln(sd, g, f'{p}norm{s}', SWIN_EMBED << s)

    v = PADDING + 1 + vie_categories
    t = 'transformer.'
    sd[t + 'embedding.word_embeddings.weight'] = g.normal((v, D_MODEL), 1.0)
    for kind in ('pt', 'poly', 'rec', 'other'):
        sd[t + f'embedding.{kind}_position_embeddings.weight'] = g.normal((MAX_POS, D_MODEL), 1.0)
    _ln(sd, g, t + 'embedding.LayerNorm', D_MODEL)
    shared_norm = OrderedDict()
    _ln(shared_norm, g, 'n', D_MODEL)
    for dec in ('pt', 'poly', 'rec'):
        for l in range(DEC_LAYERS):
            q = f'{t}{dec}_decoder.layers.{l}.'
            for attn in ('self_attn', 'multihead_attn'):
                b = 1.0 / math.sqrt(D_MODEL)
                sd[q + attn + '.in_proj_weight'] = g.uniform((3 * D_MODEL, D_MODEL), 2.0 * b)
                sd[q + attn + '.in_proj_bias'] = g.uniform((3 * D_MODEL,), b)
                _linear(sd, g, q + attn + '.out_proj', D_MODEL, D_MODEL)
            _linear(sd, g, q + 'linear1', DEC_FFN, D_MODEL)
            _linear(sd, g, q + 'linear2', D_MODEL, DEC_FFN)
            for n in ('norm3', 'norm1', 'norm2'):
                _ln(sd, g, q + n, D_MODEL)
        # one nn.LayerNorm object is shared by the three decoders (transformer.py:24-33)
        sd[f'{t}{dec}_decoder.norm.weight'] = shared_norm['n.weight'].clone()
        sd[f'{t}{dec}_decoder.norm.bias'] = shared_norm['n.bias'].clone()
    for dec in ('pt', 'poly', 'rec'):
        q = f'{t}{dec}_pred_layer.layers.'
        _linear(sd, g, q + '0', D_MODEL, D_MODEL)
        _linear(sd, g, q + '1', D_MODEL, D_MODEL)
        _linear(sd, g, q + '2', v, D_MODEL, gain=4.0)
    sd[t + 'pt_pred_layer.layers.2.bias'][PT_EOS] += pt_eos_bias
    for i, cin in enumerate((1024, 512, 256, 128)):
        _linear(sd, g, f'fpn.fpn_in.{i}', 256, cin, bias=False, shape=(256, cin, 1, 1))
    _linear(sd, g, 'input_proj', D_MODEL, 1024, shape=(D_MODEL, 1024, 1, 1))
    return sd


# ---- MGP-STR-base (OCR/MGP-STR/modules/mgp_str.py:195-206) -------------------------------------
VIT_DIM, VIT_DEPTH, VIT_HEADS, VIT_TOKENS = 768, 12, 12, 257
MGP_MAXLEN = 27
MGP_CHAR, MGP_BPE, MGP_WP = 38, 50257, 30522


# the four released sizes (OCR/MGP-STR/modules/mgp_str.py:176-230): embed dim, depth, heads
MGP_VARIANTS = {'tiny': (192, 12, 3), 'small': (384, 12, 6), 'base': (768, 12, 12), 'large': (1024, 24, 16)}


def mgpstr_state_dict(seed: int = 0, prefix: str = 'module.mgp_str.', dim: int = VIT_DIM,
                      depth: int = VIT_DEPTH, heads: int = VIT_HEADS, char_only: bool = False):
    """char_only: the CHAR-STR ablation (modules/char_str.py:43-81): one A^3 module; the logits come from timm's
    `head` (char_str.py:70), `char_head` exists after reset_classifier but is never called."""
    g = _Gen(seed + 7919)
    sd = OrderedDict()
    sd['cls_token'] = g.normal((1, 1, dim), 0.5)
    sd['pos_embed'] = g.normal((1, VIT_TOKENS, dim), 0.5)
    _linear(sd, g, 'patch_embed.proj', dim, 48, shape=(dim, 3, 4, 4), gain=3.0)
    for b in range(depth):
        q = f'blocks.{b}.'
        _ln(sd, g, q + 'norm1', dim)
        _linear(sd, g, q + 'attn.qkv', 3 * dim, dim, gain=2.0)
        _linear(sd, g, q + 'attn.proj', dim, dim)
        _ln(sd, g, q + 'norm2', dim)
        _linear(sd, g, q + 'mlp.fc1', 4 * dim, dim)
        _linear(sd, g, q + 'mlp.fc2', dim, 4 * dim)
    _ln(sd, g, 'norm', dim)          # present in the checkpoints, never applied (mgp_str.py:64-94)
    _linear(sd, g, 'head', MGP_CHAR, dim)  # timm's classifier, unused as well
    for a in (('char',) if char_only else ('char', 'bpe', 'wp')):
        q = f'{a}_tokenLearner.'
        _ln(sd, g, q + 'token_norm', dim)
        sd[q + 'tokenLearner.0.weight'] = g.uniform((dim, dim // 8, 1, 1), 1.0 / math.sqrt(dim // 8))
        sd[q + 'tokenLearner.1.weight'] = g.uniform((MGP_MAXLEN, dim, 1, 1), 4.0 / math.sqrt(dim))
        sd[q + 'feat.weight'] = g.uniform((dim, dim // 8, 1, 1), 1.0 / math.sqrt(dim // 8))
        _ln(sd, g, q + 'norm', dim)
    _linear(sd, g, 'char_head', MGP_CHAR, dim, gain=4.0)
    if not char_only:
        _linear(sd, g, 'bpe_head', MGP_BPE, dim, gain=4.0)
        _linear(sd, g, 'wp_head', MGP_WP, dim, gain=4.0)
    return OrderedDict((prefix + k, v) for k, v in sd.items())
