"""Test-time image pipeline on the GPU through the C ABI.

    omni_plan(sizes, test_min_size, test_max_size)   RandomResize.get_size_with_aspect_ratio per page + batch canvas
                                                     (OCR/OmniParser/dataset/transforms.py:275-296; host only)
    omni_pages(ctx, images, ...) -> NestedTensor     RandomResize -> ToTensor -> Normalize -> nested_tensor_from_tensor_list
                                                     (dataset/__init__.py:109-113, transforms.py:249-298,312-322,
                                                      utils/nested_tensor.py:37-54), on the device
    mgp_crops(ctx, images, imgH, imgW) -> tensor     PIL bicubic resize + ToTensor (OCR/MGP-STR/demo.py:126-132)

`images`: decoded 8-bit RGB pages, each a uint8 array / tensor [h, w, 3] (host numpy / torch, or CUDA torch).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .nested_tensor import NestedTensor


def omni_plan(sizes: Sequence[Tuple[int, int]], test_min_size: int, test_max_size: int):
    """sizes: (h, w) per page -> (resized [(h, w), ...], (Hmax, Wmax))."""
    lib = _lib.load()
    n = len(sizes)
    hs = np.asarray([s[0] for s in sizes], dtype=np.int32)
    ws = np.asarray([s[1] for s in sizes], dtype=np.int32)
    out = np.zeros((max(n, 1), 2), dtype=np.int32)
    hm, wm = C.c_int(0), C.c_int(0)
    rc = lib.alm_pre_omni_plan(hs.ctypes.data, ws.ctypes.data, n, test_min_size, test_max_size, out.ctypes.data,
                               C.byref(hm), C.byref(wm))
    if rc != 0:
        raise _lib.AlmError(rc, 'alm_pre_omni_plan: invalid sizes')
    return [tuple(int(v) for v in out[i]) for i in range(n)], (hm.value, wm.value)


def resample_coeffs(in_size: int, out_size: int, filter: str = 'bilinear'):
    """(bounds [out,2], coefs [out,ksize]) int32: the fixed-point weights one resampling pass of the kernels uses."""
    lib = _lib.load()
    f = {'bilinear': 0, 'bicubic': 1}[filter]
    k = C.c_int(0)
    lib.alm_pre_coeffs(in_size, out_size, f, C.byref(k), None, None, 0)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coefs = np.zeros((out_size, k.value), dtype=np.int32)
    rc = lib.alm_pre_coeffs(in_size, out_size, f, C.byref(k), bounds.ctypes.data, coefs.ctypes.data, coefs.size)
    if rc != 0:
        raise _lib.AlmError(rc, 'alm_pre_coeffs')
    return bounds, coefs


def _image_args(images):
    keep, ptrs, hs, ws = [], [], [], []
    for im in images:
        t = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im))
        assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, 'pages must be uint8 [h, w, 3]'
        t = t.contiguous()
        keep.append(t)
        ptrs.append(t.data_ptr())
        hs.append(t.shape[0])
        ws.append(t.shape[1])
    n = len(keep)
    return keep, (C.c_void_p * n)(*ptrs), np.asarray(hs, dtype=np.int32), np.asarray(ws, dtype=np.int32), n


def omni_pages(ctx: _lib.Context, images: List, test_min_size: int, test_max_size: int) -> NestedTensor:
    """The reference's validation transform + batching, on the GPU: returns the NestedTensor `model(samples, seqs)` takes
    (tensors f32 [n,3,Hmax,Wmax] normalised and zero padded, mask bool [n,Hmax,Wmax], both on the context's device)."""
    keep, ptrs, hs, ws, n = _image_args(images)
    _, (Hc, Wc) = omni_plan(list(zip(hs.tolist(), ws.tolist())), test_min_size, test_max_size)
    dev = torch.device('cuda', ctx.device)
    tensors = torch.empty(n, 3, Hc, Wc, dtype=torch.float32, device=dev)
    mask = torch.empty(n, Hc, Wc, dtype=torch.uint8, device=dev)
    ctx.wait_torch(*keep)
    ctx.check(ctx.lib.alm_pre_omni_pages(ctx.h, ptrs, hs.ctypes.data, ws.ctypes.data, n, test_min_size, test_max_size,
                                         tensors.data_ptr(), mask.data_ptr()))
    ctx.synchronize()
    return NestedTensor(tensors, mask.bool())


def mgp_crops(ctx: _lib.Context, images: List, imgH: int = 32, imgW: int = 128) -> torch.Tensor:
    """[n,3,imgH,imgW] f32 in [0,1] on the device: `img.resize((imgW, imgH), Image.BICUBIC)` + ToTensor per crop."""
    keep, ptrs, hs, ws, n = _image_args(images)
    out = torch.empty(n, 3, imgH, imgW, dtype=torch.float32, device=torch.device('cuda', ctx.device))
    ctx.wait_torch(*keep)
    ctx.check(ctx.lib.alm_pre_mgp_crops(ctx.h, ptrs, hs.ctypes.data, ws.ctypes.data, n, imgH, imgW, out.data_ptr()))
    ctx.synchronize()
    return out
