"""Pin of the timm layer -- TEST INFRASTRUCTURE ONLY.

The MGP-STR encoder arithmetic lives in timm==0.4.12 (OCR/MGP-STR/requirements.txt:4; modules/mgp_str.py:19-21,46),
which is neither vendored under /root/reference nor installable here, so oracle/shim/timm/models/vision_transformer.py
and oracle/mgpstr_ref.py restate it.  This module checks that restatement against an INDEPENDENT implementation that
IS installed: HuggingFace ``transformers.models.mgp_str`` (a port of the released MGP-STR, written by other people from
the released code and validated by them against the released checkpoints).  The port uses one ``layer_norm_eps`` for
every LayerNorm (default 1e-5), whereas the reference has eps 1e-6 in the 24 timm block norms and 1e-5 in the six A^3
norms (modules/token_learner.py:15,20): the HF model is therefore built with layer_norm_eps=1e-6 and the A^3 norms
are put back to 1e-5 on the instantiated modules.

    python -m oracle.pin_timm_hf      # regenerates tests/golden/mgp_hf_b2.npz (HF outputs on the synthetic checkpoint)

tests/test_oracle_golden.py runs the same comparison on every CPU test run (transformers travels with the image).
"""
from __future__ import annotations

import os

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, 'tests', 'golden')


def hf_key(k: str) -> str:
    """reference / timm state-dict key (without the ``module.mgp_str.`` prefix) -> HF MgpstrForSceneTextRecognition key."""
    if k in ('cls_token', 'pos_embed'):
        return 'mgp_str.embeddings.' + k
    if k.startswith('patch_embed.proj.'):
        return 'mgp_str.embeddings.proj.' + k[len('patch_embed.proj.'):]
    if k.startswith('blocks.'):
        return 'mgp_str.encoder.' + k
    for a in ('char', 'bpe', 'wp'):
        if k.startswith(a + '_tokenLearner.'):
            return a + '_a3_module.' + k[len(a + '_tokenLearner.'):]
    return k  # char_head / bpe_head / wp_head


def build_hf(sd, dim, depth, heads, prefix='module.mgp_str.'):
    from transformers import MgpstrConfig, MgpstrForSceneTextRecognition
    cfg = MgpstrConfig(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, layer_norm_eps=1e-6)
    model = MgpstrForSceneTextRecognition(cfg).eval()
    for a in ('char', 'bpe', 'wp'):  # nn.LayerNorm default eps in token_learner.py:15,20
        m = getattr(model, a + '_a3_module')
        m.token_norm.eps = 1e-5
        m.norm.eps = 1e-5
    mapped = {}
    for k, v in sd.items():
        k = k[len(prefix):]
        if k.startswith('norm.') or k.startswith('head.'):
            continue  # timm's final norm / classifier: present in the checkpoints, never applied (mgp_str.py:64-94)
        mapped[hf_key(k)] = v
    r = model.load_state_dict(mapped, strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    return model


def cross_check(sd, img, dim, depth, heads):
    """max |restatement - HF| for the encoder output, the three A^3 attention maps and the three logit tensors."""
    from . import mgpstr_ref as M
    model = build_hf(sd, dim, depth, heads)
    with torch.no_grad():
        hf_tokens = model.mgp_str(img).last_hidden_state
        out = model(img, output_a3_attentions=True)
        tokens, _ = M.backbone(img, sd, depth=depth, heads=heads)
        ours = M.forward(img, sd, depth=depth, heads=heads)
    d = {'encoder': float((hf_tokens - tokens).abs().max())}
    for i, a in enumerate(('char', 'bpe', 'wp')):
        d[a + '_attn'] = float((out.a3_attentions[i] - ours[0][i]).abs().max())
        d[a + '_logits'] = float((out.logits[i] - ours[1 + i]).abs().max())
    return d, out


if __name__ == '__main__':
    from . import weights as W
    torch.set_grad_enabled(False)
    sd = W.mgpstr_state_dict(seed=0)
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 32, 128, generator=g)
    d, out = cross_check(sd, img, W.VIT_DIM, W.VIT_DEPTH, W.VIT_HEADS)
    print('restated timm-0.4.12 ViT + A^3 vs HF transformers port:', {k: f'{v:.2e}' for k, v in d.items()})
    assert max(d.values()) < 2e-5
    np.savez_compressed(os.path.join(GOLD, 'mgp_hf_b2.npz'), char=out.logits[0].numpy(),
                        bpe_s=out.logits[1].reshape(-1)[::997].numpy(), wp_s=out.logits[2].reshape(-1)[::997].numpy(),
                        bpe_ids=out.logits[1].argmax(-1).numpy(), wp_ids=out.logits[2].argmax(-1).numpy(),
                        char_attn=out.a3_attentions[0].numpy())
    print('wrote tests/golden/mgp_hf_b2.npz')
