"""MGPSTRB200: the reference's ``model(image, is_eval=True)`` call site served by libalm_ocr.so.

Reference boundary (relative to /root/reference/OCR/MGP-STR/): demo.py:33, test_final.py:140,
modules/mgp_str.py:96-101; checkpoint keys carry the ``module.mgp_str.`` prefix (test_final.py:348,356).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class MGPSTRB200:
    def __init__(self, state_dict, device: int = 0, stream: Optional[int] = None, ctx: Optional[_lib.Context] = None,
                 share_from: Optional['MGPSTRB200'] = None):
        """state_dict: the reference checkpoint (``module.mgp_str.*`` keys); None with ``share_from=other`` makes this
        a second execution context over the same device weights; None alone when ``ctx`` already holds the weights."""
        self.ctx = ctx or _lib.Context(device, stream)
        if share_from is not None:
            assert state_dict is None
            self.ctx.share_weights(share_from.ctx)
        elif state_dict is not None:
            self.ctx.load_state_dict(_lib.MODEL_MGPSTR, state_dict)
        self.lib = self.ctx.lib
        self._info = None

    def info(self):
        """dict(dim, depth, heads, n_a3, vocab): the variant the checkpoint defines (tiny / small / base / large, or the
        char-only CHAR-STR with n_a3 == 1)."""
        if self._info is None:
            import ctypes as C
            d, dp, h, n = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            v = (C.c_int * 3)()
            self.ctx.check(self.lib.alm_mgpstr_info(self.ctx.h, C.byref(d), C.byref(dp), C.byref(h), C.byref(n), v))
            self._info = dict(dim=d.value, depth=dp.value, heads=h.value, n_a3=n.value, vocab=[int(x) for x in v])
        return self._info

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def __call__(self, image, is_eval=False):
        return self.forward(image, is_eval)

    def recognize(self, image: torch.Tensor):
        """What the reference drivers keep of the forward (demo.py:36-60, test_final.py:146-172): per head the top-1 id
        and its softmax probability at each of the 27 positions -> (ids int32 [3,B,27], probs f32 [3,B,27]).  The
        argmax / max-softmax are fused on the device; neither the logits (2.8 GB for the BPE head at B = 512) nor the
        A^3 maps cross PCIe."""
        assert image.dim() == 4 and tuple(image.shape[1:]) == (3, 32, 128) and image.dtype == torch.float32
        img = image.contiguous()
        B = img.shape[0]
        self.ctx.wait_torch(img)
        ids = torch.empty(3, B, 27, dtype=torch.int32)
        prob = torch.empty(3, B, 27, dtype=torch.float32)
        self.ctx.check(self.lib.alm_mgpstr_forward(self.ctx.h, img.data_ptr(), B, None, None, None, None, ids.data_ptr(),
                                                   prob.data_ptr()))
        self.last_ids, self.last_prob = ids, prob
        return ids, prob

    def forward(self, image: torch.Tensor, is_eval: bool = False, want_logits: bool = True):
        """-> [[char_attn, bpe_attn, wp_attn], char, bpe, wp] when is_eval else [char, bpe, wp]."""
        assert image.dim() == 4 and tuple(image.shape[1:]) == (3, 32, 128) and image.dtype == torch.float32
        img = image.contiguous()
        B = img.shape[0]
        self.ctx.wait_torch(img)  # a CUDA input is ordered after torch's current stream
        inf = self.info()
        vc, vb, vw = inf['vocab']
        char_only = inf['n_a3'] == 1
        attn = torch.empty(3, B, 27, 257, dtype=torch.float32)
        char = torch.empty(B, 27, vc, dtype=torch.float32)
        bpe = torch.empty(B, 27, vb, dtype=torch.float32) if want_logits and not char_only else None
        wp = torch.empty(B, 27, vw, dtype=torch.float32) if want_logits and not char_only else None
        ids = torch.empty(3, B, 27, dtype=torch.int32)
        prob = torch.empty(3, B, 27, dtype=torch.float32)
        self.ctx.check(self.lib.alm_mgpstr_forward(
            self.ctx.h, img.data_ptr(), B, attn.data_ptr(), char.data_ptr(),
            bpe.data_ptr() if bpe is not None else None, wp.data_ptr() if wp is not None else None,
            ids.data_ptr(), prob.data_ptr()))
        self.last_ids, self.last_prob = ids, prob
        if char_only:  # CHARSTR.forward (modules/char_str.py:76-81)
            return [[attn[0]], char] if is_eval else [char]
        if is_eval:
            return [[attn[0], attn[1], attn[2]], char, bpe, wp]
        return [char, bpe, wp]
