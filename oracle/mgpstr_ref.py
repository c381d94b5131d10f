"""CPU fp32 restatement of the MGP-STR forward -- TEST INFRASTRUCTURE ONLY.

Restates OCR/MGP-STR/modules/mgp_str.py:64-101 (MGPSTR.forward_features), the A^3 module
OCR/MGP-STR/modules/token_learner.py:21-32 and the timm==0.4.12 ViT blocks it inherits
(third-party, pinned in OCR/MGP-STR/requirements.txt:4; block LayerNorm eps = 1e-6, scores
scaled AFTER q@k^T, timm's final ``norm`` never applied).  Pinned by oracle/gen_golden.py against
the unmodified reference MGPSTR class running over the timm shim (oracle/shim) -- the timm layer
itself is a restatement of the published 0.4.12 source ("parity unpinned" at the timm boundary,
SURVEY.md section 8c).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import weights as W


def _ln(x, sd, p, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], eps)


def vit_block(x, sd, p, heads):
    B, N, C = x.shape
    y = _ln(x, sd, p + 'norm1', 1e-6)
    qkv = F.linear(y, sd[p + 'attn.qkv.weight'], sd[p + 'attn.qkv.bias'])
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(y, sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias'])
    y = _ln(x, sd, p + 'norm2', 1e-6)
    y = F.gelu(F.linear(y, sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias']))
    return x + F.linear(y, sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])


def token_learner(x, sd, p):  # token_learner.py:21-32
    x = _ln(x, sd, p + 'token_norm', 1e-5)
    xc = x.transpose(1, 2).unsqueeze(-1)
    sel = F.conv2d(F.conv2d(xc, sd[p + 'tokenLearner.0.weight'], groups=8), sd[p + 'tokenLearner.1.weight'])
    sel = F.softmax(sel.flatten(2), dim=-1)
    feat = F.conv2d(xc, sd[p + 'feat.weight'], groups=8).flatten(2).transpose(1, 2)
    out = torch.einsum('...si,...id->...sd', sel, feat)
    return sel, _ln(out, sd, p + 'norm', 1e-5)


def backbone(img, sd, prefix='module.mgp_str.', depth=W.VIT_DEPTH, heads=W.VIT_HEADS):
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    B = img.shape[0]
    x = F.conv2d(img, sd['patch_embed.proj.weight'], sd['patch_embed.proj.bias'], stride=4)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd['cls_token'].expand(B, -1, -1), x), dim=1) + sd['pos_embed']
    for b in range(depth):
        x = vit_block(x, sd, f'blocks.{b}.', heads)
    return x, sd


def forward(img, sd, prefix='module.mgp_str.', depth=W.VIT_DEPTH, heads=W.VIT_HEADS):
    """-> [[char_attn, bpe_attn, wp_attn] each [B,27,257], char [B,27,38], bpe [B,27,50257],
    wp [B,27,30522]]   (mgp_str.py:96-101 with is_eval=True)."""
    x, sd = backbone(img, sd, prefix, depth, heads)
    attns, outs = [], []
    if 'bpe_tokenLearner.token_norm.weight' not in sd:  # CHAR-STR (modules/char_str.py:56-81): logits through `head`
        sel, y = token_learner(x, sd, 'char_tokenLearner.')
        return [[sel], F.linear(y, sd['head.weight'], sd['head.bias'])]
    for a in ('char', 'bpe', 'wp'):
        sel, y = token_learner(x, sd, f'{a}_tokenLearner.')
        attns.append(sel)
        outs.append(F.linear(y, sd[f'{a}_head.weight'], sd[f'{a}_head.bias']))
    return [attns] + outs


def greedy_ids(outs):
    """demo.py:36-60: top-1 id and max softmax prob per position for the three heads."""
    res = []
    for lg in outs:
        p = F.softmax(lg, dim=2)
        res.append((lg.argmax(-1), p.max(dim=2)[0]))
    return res
