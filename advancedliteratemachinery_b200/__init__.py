"""B200-native OCR forward path (OmniParser Swin-B + point-conditioned decoder, MGP-STR ViT-Base).

Drop-in adapters over libalm_ocr.so (hand-written sm_100a CUDA behind a C ABI, include/alm_ocr.h):

    OmniParserB200  -- replaces reference ``OmniParser.forward``   (OCR/OmniParser/model/omniparser.py:19-32)
    MGPSTRB200      -- replaces reference ``MGPSTR.forward``       (OCR/MGP-STR/modules/mgp_str.py:96-101)

There is no CPU / PyTorch fallback: importing the adapters without the built library raises.
"""
from ._lib import AlmError, Context, LIB_PATH  # noqa: F401
from .nested_tensor import NestedTensor, nested_tensor_from_tensor_list  # noqa: F401
from .omniparser import OmniParserB200, OmniVocab  # noqa: F401
from .mgp_str import MGPSTRB200  # noqa: F401
from . import postprocess, preprocess  # noqa: F401
