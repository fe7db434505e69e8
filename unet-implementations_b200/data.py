"""Device-side input conversion of the reference's dataset class (`PetSegmentationDataset.__getitem__`,
Our_UNet/src/train.py:299-311) for a whole batch: uint8 HWC images and uint8 masks cross PCIe as they are stored
(4 bytes per pixel instead of 12 + 8) and one kernel produces what the trainer feeds the model and the loss -- fp32
NCHW `(x / 255 - mean) / std` and int64 masks with stray labels cleaned.  SURVEY.md section 8f, row 2.  Bit-exact with
the reference's torch CPU arithmetic (tests/test_gpu_aux.py)."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # train.py:306
IMAGENET_STD = (0.229, 0.224, 0.225)   # train.py:307


def preprocess_batch(images_u8: Optional[torch.Tensor], masks_u8: Optional[torch.Tensor] = None,
                     mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
    """images_u8: CUDA uint8 [B,H,W,3]; masks_u8: CUDA uint8 [B,H,W] -> (fp32 [B,3,H,W] or None, int64 [B,H,W] or None)."""
    ref = images_u8 if images_u8 is not None else masks_u8
    if ref is None or not ref.is_cuda:
        raise RuntimeError("b200unet: preprocess_batch needs CUDA uint8 tensors; there is no CPU path")
    img_out = mask_out = None
    if images_u8 is not None:
        assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.size(3) == 3 and images_u8.is_contiguous()
        b, h, w, _ = images_u8.shape
        img_out = torch.empty((b, 3, h, w), dtype=torch.float32, device=images_u8.device)
    if masks_u8 is not None:
        assert masks_u8.dtype == torch.uint8 and masks_u8.dim() == 3 and masks_u8.is_contiguous()
        b, h, w = masks_u8.shape
        mask_out = torch.empty((b, h, w), dtype=torch.int64, device=masks_u8.device)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)  # noqa: E731
    with torch.cuda.device(ref.device):
        _lib.call("b200unet_preprocess_u8", p(images_u8), p(masks_u8), p(img_out), p(mask_out), m3, s3, b, h * w,
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return img_out, mask_out
