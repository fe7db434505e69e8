"""Validation metric of the reference's trainer (`validate`, Our_UNet/src/train.py:536-585) as one kernel:
argmax over the 3 logits + per-class intersection / prediction / target counts over valid pixels, integer counters,
no host synchronisation (the reference does nine `.item()` syncs per batch).  SURVEY.md section 8f, row 5."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def argmax_counts(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = 255, want_pred: bool = True):
    """logits fp32 [B,3,H,W], target int64 [B,H,W] -> (pred int64 [B,H,W] or None, counts int64 [3,3]) with
    counts[c] = (#pred==c & target==c, #pred==c, #target==c) over pixels whose target is not `ignore_index`."""
    if not logits.is_cuda:
        raise RuntimeError("b200unet: argmax_counts needs CUDA tensors; there is no CPU path")
    lg = logits.detach()
    if lg.dtype != torch.float32 or not lg.is_contiguous():
        lg = lg.float().contiguous()
    n, k, h, w = lg.shape
    assert k == 3, "the trainer's metric is built for 3 classes (train.py:557)"
    tg = target.contiguous()
    assert tg.dtype == torch.int64 and tuple(tg.shape) == (n, h, w)
    pred = torch.empty((n, h, w), dtype=torch.int64, device=lg.device) if want_pred else None
    counts = torch.empty((3, 3), dtype=torch.int64, device=lg.device)
    with torch.cuda.device(lg.device):
        _lib.call("b200unet_argmax_counts", ctypes.c_void_p(lg.data_ptr()), ctypes.c_void_p(tg.data_ptr()), int(ignore_index),
                  ctypes.c_void_p(pred.data_ptr()) if pred is not None else ctypes.c_void_p(0),
                  ctypes.c_void_p(counts.data_ptr()), n, h * w, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return pred, counts


def dice_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """Per-class Dice as validate() computes it (train.py:566-572): 2*I / (union + 1e-5), 1.0 when union == 0.
    Stays on the device (one sync per epoch instead of nine per batch)."""
    inter = counts[:, 0].to(torch.float32)
    union = (counts[:, 1] + counts[:, 2]).to(torch.float32)
    return torch.where(union > 0, 2.0 * inter / (union + 1e-5), torch.ones_like(union))
