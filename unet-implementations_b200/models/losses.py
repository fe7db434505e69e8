"""B200-native drop-in for the reference's `models.losses.SimpleLoss` (Our_UNet/models/losses.py:5-121).

Same constructor and `forward(input, target)` contract: weight_ce * CrossEntropy(weight=w, ignore_index) +
weight_dice * Dice, with `w` recomputed per batch from the inverse class frequency when `dynamic_weights` is set
(losses.py:24-62).  The ~45 small ATen launches and 3 host syncs of the reference become one reduction kernel over
(logits, target) plus a tiny finalize in the forward, and one elementwise kernel in the backward
(b200unet_loss_fwd / b200unet_loss_bwd, include/b200unet.h).  No host synchronisation happens anywhere.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

try:
    from .. import ops
except ImportError:
    from unet_implementations_b200 import ops


class _SimpleLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, class_weights, dynamic, weight_ce, weight_dice, ignore_index, smooth):
        ops.require_device()
        with torch.cuda.device(logits.device):
            lg = logits.detach()
            if lg.dtype != torch.float32 or not lg.is_contiguous():
                lg = lg.float().contiguous()  # the reference's CE/softmax run in fp32 under autocast
            tg = target.detach()
            if tg.dtype == torch.uint8:  # masks as the dataset stores them: read as they are
                tg = tg.contiguous()
            elif tg.dtype != torch.int64 or not tg.is_contiguous():
                tg = tg.long().contiguous()
            out, tables = ops.loss_forward(lg, tg, class_weights, dynamic, weight_ce, weight_dice, ignore_index, smooth)
        ctx.save_for_backward(lg, tg, tables)
        ctx.cfg = (weight_ce, weight_dice, ignore_index)
        ctx.in_dtype = logits.dtype
        ctx.parts = out  # [total, ce, dice] (kept for logging; no sync)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        lg, tg, tables = ctx.saved_tensors
        weight_ce, weight_dice, ignore_index = ctx.cfg
        with torch.cuda.device(lg.device):
            dl = ops.loss_backward(lg, tg, tables, grad_out.float(), weight_ce, weight_dice, ignore_index)
        if ctx.in_dtype != torch.float32:
            dl = dl.to(ctx.in_dtype)
        return dl, None, None, None, None, None, None, None


class SimpleLoss(nn.Module):
    """Combined Dice + (dynamically) class-weighted cross entropy with ignore_index handling (losses.py:5-22)."""

    def __init__(self, weight_dice=1.0, weight_ce=1.0, ignore_index=255, smooth=1e-5, class_weights=None,
                 dynamic_weights=True):
        super().__init__()
        self.weight_dice = weight_dice
        self.weight_ce = weight_ce
        self.ignore_index = ignore_index
        self.smooth = smooth
        self.class_weights = class_weights
        self.dynamic_weights = dynamic_weights
        # kept for surface compatibility (the reference exposes `.ce`, losses.py:22); the kernels do not call it
        self.ce = nn.CrossEntropyLoss(weight=class_weights, ignore_index=ignore_index)

    def forward(self, input, target):
        if not input.is_cuda:
            raise RuntimeError("b200unet: SimpleLoss needs CUDA tensors on an sm_100 device; there is no CPU path")
        if input.shape[-2:] != target.shape[-2:]:
            # resize guard of the reference (losses.py:66-68); never taken on the training path (the model returns
            # logits at the mask's resolution), so it stays a library call at the boundary
            input = F.interpolate(input, size=target.shape[-2:], mode="bilinear", align_corners=False)
        if input.size(1) != 3:
            raise NotImplementedError("b200unet: SimpleLoss kernels are built for the reference's 3 classes (losses.py:40)")
        dynamic = bool(self.dynamic_weights) and target.size(0) > 0
        cw = None
        if not dynamic and self.class_weights is not None:
            cw = torch.as_tensor(self.class_weights, dtype=torch.float32, device=input.device).contiguous()
        return _SimpleLossFunction.apply(input, target, cw, dynamic, float(self.weight_ce), float(self.weight_dice),
                                         int(self.ignore_index), float(self.smooth))


class _MSEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, target):
        ops.require_device()
        with torch.cuda.device(input.device):
            a = input.detach()
            if a.dtype != torch.float32 or not a.is_contiguous():
                a = a.float().contiguous()
            b = target.detach()
            if b.dtype != torch.float32 or not b.is_contiguous():
                b = b.float().contiguous()
            out = ops.mse_forward(a, b)
        ctx.save_for_backward(a, b)
        ctx.in_dtype = input.dtype
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        with torch.cuda.device(a.device):
            da = ops.mse_backward(a, b, grad_out.float())
        if ctx.in_dtype != torch.float32:
            da = da.to(ctx.in_dtype)
        return da, None


class MSELoss(nn.Module):
    """`nn.MSELoss()` (mean reduction) as the autoencoder trainer uses it (AE_pretrained/reconstruction/src/train.py:431):
    one reduction kernel forward, one elementwise kernel backward (b200unet_mse_fwd / b200unet_mse_bwd)."""

    def forward(self, input, target):
        if not input.is_cuda:
            raise RuntimeError("b200unet: MSELoss needs CUDA tensors on an sm_100 device; there is no CPU path")
        if input.shape != target.shape:
            raise ValueError(f"MSELoss: shapes differ: {tuple(input.shape)} vs {tuple(target.shape)}")
        return _MSEFunction.apply(input, target)

