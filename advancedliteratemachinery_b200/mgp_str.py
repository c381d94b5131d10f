"""MGPSTRB200: the reference's ``model(image, is_eval=True)`` call site served by libalm_ocr.so.

Reference boundary (relative to /root/reference/OCR/MGP-STR/): demo.py:33, test_final.py:140,
modules/mgp_str.py:96-101; checkpoint keys carry the ``module.mgp_str.`` prefix (test_final.py:348,356).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class MGPSTRB200:
    def __init__(self, state_dict, device: int = 0, stream: Optional[int] = None, ctx: Optional[_lib.Context] = None):
        self.ctx = ctx or _lib.Context(device, stream)
        self.ctx.load_state_dict(_lib.MODEL_MGPSTR, state_dict)
        self.lib = self.ctx.lib

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def __call__(self, image, is_eval=False):
        return self.forward(image, is_eval)

    def forward(self, image: torch.Tensor, is_eval: bool = False, want_logits: bool = True):
        """-> [[char_attn, bpe_attn, wp_attn], char, bpe, wp] when is_eval else [char, bpe, wp]."""
        assert image.dim() == 4 and tuple(image.shape[1:]) == (3, 32, 128) and image.dtype == torch.float32
        img = image.contiguous()
        B = img.shape[0]
        attn = torch.empty(3, B, 27, 257, dtype=torch.float32)
        char = torch.empty(B, 27, 38, dtype=torch.float32)
        bpe = torch.empty(B, 27, 50257, dtype=torch.float32) if want_logits else None
        wp = torch.empty(B, 27, 30522, dtype=torch.float32) if want_logits else None
        ids = torch.empty(3, B, 27, dtype=torch.int32)
        prob = torch.empty(3, B, 27, dtype=torch.float32)
        self.ctx.check(self.lib.alm_mgpstr_forward(
            self.ctx.h, img.data_ptr(), B, attn.data_ptr(), char.data_ptr(),
            bpe.data_ptr() if bpe is not None else None, wp.data_ptr() if wp is not None else None,
            ids.data_ptr(), prob.data_ptr()))
        self.last_ids, self.last_prob = ids, prob
        if is_eval:
            return [[attn[0], attn[1], attn[2]], char, bpe, wp]
        return [char, bpe, wp]
