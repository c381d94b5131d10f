"""NestedTensor: zero-padded image batch + boolean pad mask.

Mirror of the reference container so callers of ``model(samples, seqs)`` keep their code
(OCR/OmniParser/utils/nested_tensor.py:7-54); written fresh, same fields and meaning:
``tensors`` f32 [B,3,H,W], ``mask`` bool [B,H,W] with True = padding.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class NestedTensor:
    def __init__(self, tensors: torch.Tensor, mask: Optional[torch.Tensor]):
        self.tensors = tensors
        self.mask = mask

    def to(self, device):
        return NestedTensor(self.tensors.to(device), None if self.mask is None else self.mask.to(device))

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return f'NestedTensor(tensors={tuple(self.tensors.shape)}, mask={None if self.mask is None else tuple(self.mask.shape)})'


def nested_tensor_from_tensor_list(tensor_list: List[torch.Tensor]) -> NestedTensor:
    """Pad [3,h,w] images to the batch maximum (top-left aligned) and mark the padding."""
    if not tensor_list or tensor_list[0].dim() != 3:
        raise ValueError('expected a non-empty list of [C,H,W] tensors')
    c = tensor_list[0].shape[0]
    H = max(int(t.shape[1]) for t in tensor_list)
    W = max(int(t.shape[2]) for t in tensor_list)
    out = torch.zeros((len(tensor_list), c, H, W), dtype=tensor_list[0].dtype, device=tensor_list[0].device)
    mask = torch.ones((len(tensor_list), H, W), dtype=torch.bool, device=tensor_list[0].device)
    for i, img in enumerate(tensor_list):
        out[i, :, :img.shape[1], :img.shape[2]] = img
        mask[i, :img.shape[1], :img.shape[2]] = False
    return NestedTensor(out, mask)
