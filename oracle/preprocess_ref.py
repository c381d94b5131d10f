"""CPU restatement of the reference's test-time image pipeline -- TEST INFRASTRUCTURE ONLY (the product is
advancedliteratemachinery_b200/csrc/preproc.cu; nothing under advancedliteratemachinery_b200/ imports this).

  resize(img, oh, ow, name)        Pillow's two-pass 8-bit resampling (`Image.resize` with BILINEAR / BICUBIC).  Pillow is
                                   a third-party dependency of the reference (pillow==8.1.0, OCR/MGP-STR/requirements.txt:6;
                                   torchvision's F.resize on PIL images for OmniParser) and not under /root/reference: this
                                   restates the published algorithm of src/libImaging/Resample.c (precompute_coeffs,
                                   normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc) and is PINNED bit for bit
                                   against the Pillow installed in the build container (oracle/gen_golden_pre.py).
  omni_size(h, w, min, max)        RandomResize.get_size_with_aspect_ratio (OCR/OmniParser/dataset/transforms.py:275-296)
  omni_pages(images, min, max)     RandomResize -> ToTensor -> Normalize -> nested_tensor_from_tensor_list
                                   (dataset/__init__.py:109-113, transforms.py:249-298,312-322, utils/nested_tensor.py:37-54)
  mgp_crops(images, H, W)          bicubic resize + ToTensor (OCR/MGP-STR/demo.py:126-132)
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


FILTERS = {'bilinear': (bilinear, 1.0), 'bicubic': (bicubic, 2.0)}


def precompute(in_size, out_size, name):
    f, support0 = FILTERS[name]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        for x in range(xmax):
            w = f((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                kk[xx, x] /= ww
        bounds[xx] = (xmin, xmax)
    ik = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64))
    return bounds, ik


def resample_axis(img, out_size, name, axis):
    """img uint8 [H, W, C]; resample along axis 0 (vertical) or 1 (horizontal)."""
    in_size = img.shape[axis]
    bounds, ik = precompute(in_size, out_size, name)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], dtype=np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * ik[xx, x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize(img, out_h, out_w, name):
    h, w = img.shape[:2]
    x = img
    if out_w != w:
        x = resample_axis(x, out_w, name, 1)
    if out_h != h:
        x = resample_axis(x, out_h, name, 0)
    return x


def omni_size(h, w, size, max_size=None):
    if max_size is not None:
        mn, mx = float(min((w, h))), float(max((w, h)))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


MEAN = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
STD = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)


def to_tensor(img_u8):
    """F.to_tensor on an 8-bit RGB PIL image: HWC uint8 -> CHW float32 / 255."""
    return torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)


def omni_pages(images, test_min_size, test_max_size):
    """images: list of uint8 [h, w, 3] arrays -> (tensors [n,3,H,W] f32, mask [n,H,W] bool)."""
    outs = []
    for im in images:
        oh, ow = omni_size(im.shape[0], im.shape[1], test_min_size, test_max_size)
        t = to_tensor(resize(im, oh, ow, 'bilinear'))
        outs.append(t.sub(MEAN).div(STD))
    H, W = max(t.shape[1] for t in outs), max(t.shape[2] for t in outs)
    tensors = torch.zeros(len(outs), 3, H, W)
    mask = torch.ones(len(outs), H, W, dtype=torch.bool)
    for b, t in enumerate(outs):
        tensors[b, :, :t.shape[1], :t.shape[2]] = t
        mask[b, :t.shape[1], :t.shape[2]] = False
    return tensors, mask


def mgp_crops(images, imgH=32, imgW=128):
    return torch.stack([to_tensor(resize(im, imgH, imgW, 'bicubic')) for im in images])
