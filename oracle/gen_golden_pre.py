"""Pre-processing fixtures from the UNMODIFIED reference transforms + Pillow -> tests/golden/pre_omni.npz, pre_mgp.npz.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Usage:  python -m oracle.gen_golden_pre
Pins oracle/preprocess_ref.py: (1) `resize` against `PIL.Image.resize` bit for bit over random sizes and both
filters, (2) `omni_pages` against the reference's own RandomResize / ToTensor / Normalize classes and
nested_tensor_from_tensor_list, (3) `omni_size` against RandomResize.get_size_with_aspect_ratio over a size sweep.
"""
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, 'tests', 'golden')
REF = '/root/reference/OCR'


def main():
    sys.path[:0] = [os.path.join(REPO, 'oracle', 'shim'), os.path.join(REF, 'OmniParser')]
    for name in ('bezier',):
        sys.modules.setdefault(name, types.ModuleType(name))
    from dataset import transforms as T                      # the reference transforms, unmodified
    from utils.nested_tensor import nested_tensor_from_tensor_list
    from oracle import preprocess_ref as O
    rng = np.random.default_rng(7)
    # (1) resample vs Pillow
    n_checked = 0
    for _ in range(120):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        oh, ow = int(rng.integers(1, 100)), int(rng.integers(1, 100))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for name, flt in (('bilinear', Image.BILINEAR), ('bicubic', Image.BICUBIC)):
            ref = np.asarray(Image.fromarray(img).resize((ow, oh), flt))
            assert np.array_equal(ref, O.resize(img, oh, ow, name)), (h, w, oh, ow, name)
            n_checked += 1
    # (3) size rule sweep
    rr = T.RandomResize([64], 96)
    sizes_in, sizes_out = [], []
    for _ in range(400):
        h, w = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        mn, mx = int(rng.integers(8, 200)), int(rng.integers(8, 400))
        ref = T.RandomResize([mn], mx).get_size_with_aspect_ratio((w, h), mn, mx)
        assert tuple(ref) == O.omni_size(h, w, mn, mx), (h, w, mn, mx, ref)
        sizes_in.append((h, w, mn, mx))
        sizes_out.append(tuple(ref))
    # (2) full OmniParser validation transform on small pages of different shapes
    shapes = [(37, 53), (64, 48), (90, 30), (20, 200), (64, 64), (5, 7), (131, 77)]
    pages = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    pages[4][:] = 255                                         # a saturated page
    pages[5][:] = 0
    tlist = []
    for p in pages:
        img, target = rr(Image.fromarray(p), {})
        img, target = T.ToTensor()(img, target)
        img, target = T.Normalize()(img, target)
        tlist.append(img)
    nt = nested_tensor_from_tensor_list(tlist)
    mine_t, mine_m = O.omni_pages(pages, 64, 96)
    assert torch.equal(nt.tensors, mine_t) and torch.equal(nt.mask, mine_m), 'restatement differs from the reference'
    np.savez_compressed(os.path.join(GOLD, 'pre_omni.npz'), n_pages=len(pages), min_size=64, max_size=96,
                        tensors=nt.tensors.numpy(), mask=nt.mask.numpy(),
                        sizes_in=np.asarray(sizes_in, dtype=np.int32), sizes_out=np.asarray(sizes_out, dtype=np.int32),
                        **{f'page{i}': p for i, p in enumerate(pages)})
    # MGP-STR crops: demo.py:126-132 (PIL bicubic to imgW x imgH, ToTensor)
    import torchvision.transforms as tvt
    cshapes = [(17, 60), (32, 128), (48, 200), (100, 31), (9, 9), (64, 257)]
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in cshapes]
    ref = torch.stack([tvt.ToTensor()(Image.fromarray(c).resize((128, 32), Image.BICUBIC)) for c in crops])
    assert torch.equal(ref, O.mgp_crops(crops, 32, 128))
    np.savez_compressed(os.path.join(GOLD, 'pre_mgp.npz'), n=len(crops), out=ref.numpy(),
                        **{f'crop{i}': c for i, c in enumerate(crops)})
    print(f'resample pinned on {n_checked} cases; pre_omni.npz canvas {tuple(nt.tensors.shape)}; pre_mgp.npz {tuple(ref.shape)}')


if __name__ == '__main__':
    main()
