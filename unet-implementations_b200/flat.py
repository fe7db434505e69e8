"""Flat gradient buffer of the UNet step: every parameter gradient at a fixed 16-byte-aligned offset of ONE fp32
tensor, laid out in the order `UNet`'s fused backward produces them (head, decoder 4..0, bottleneck, encoder 4..0).

The fused backward asks the sink for the destination of a gradient (`dest(p)`) and has the producing kernel -- the
weight-gradient finalize, the norm-backward parameter sums, the head backward -- write it THERE; nothing is copied.
The 22 conv biases that feed an InstanceNorm have an exactly-zero gradient: their slots are zeroed
once at construction and never written again.  `delivered(p)` tells the sink that the kernels producing p's gradient
have been enqueued (the data-parallel subclass in ddp.py launches a bucket's all-reduce when its last gradient has
been delivered); `finish()` runs at the end of backward.  The flat layout is also what the fused optimizer step
walks (optim.FusedSGD with `model=`): master weights, gradients and momentum at the same offsets.

Reference: the reference has neither a flat buffer nor a fused step -- `optimizer.zero_grad()` / `loss.backward()` /
`optimizer.step()` over 94 separate tensors (Our_UNet/src/train.py:634, :649-650, :663-664).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn


def backward_param_order(model) -> List[nn.Parameter]:
    """Parameters in the order UNet's backward produces their gradients (models/unet.py:_backward_impl)."""
    order: List[nn.Parameter] = []
    head = model._head_conv() if hasattr(model, "_head_conv") else model.segmentation_output
    order += [head.weight] + ([head.bias] if head.bias is not None else [])
    layers = model._layers()
    fusion = model._fusion_unit() if hasattr(model, "_fusion_unit") else None
    if fusion is not None:  # between the encoder and the decoder (models/clip_unet.py); needs the extra features every step
        layers = [L for L in layers if L["kind"] == "enc"] + [dict(kind="fusion", unit=fusion)] + \
                 [L for L in layers if L["kind"] == "dec"]
    for L in reversed(layers):
        conv, norm, _, _ = L["unit"]
        order += [norm.weight, norm.bias]
        if conv.bias is not None:
            order.append(conv.bias)
        order.append(conv.weight)
    seen, uniq = set(), []
    for p in order:
        if p is not None and id(p) not in seen:
            seen.add(id(p))
            uniq.append(p)
    # anything the walk above does not know (a subclass's extra parameters) goes last
    for p in model.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            uniq.append(p)
    return uniq


class FlatGradSink:
    """Gradient sink over one flat fp32 buffer (no communication).  Attach with `FlatGradSink(model)`."""

    def __init__(self, model, device: Optional[torch.device] = None, include_frozen: bool = False):
        params = [p for p in backward_param_order(model) if p.requires_grad or include_frozen]
        if not params:
            raise ValueError("FlatGradSink: the model has no trainable parameter")
        dev = device or params[0].device
        self.params: List[nn.Parameter] = params
        self.offsets: Dict[int, int] = {}
        off = 0
        for p in params:
            self.offsets[id(p)] = off
            off += (p.numel() + 3) // 4 * 4  # keep every gradient 16-byte aligned
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self._views: Dict[int, torch.Tensor] = {}
        model._grad_sink = self

    # ------------------------------------------------------------------------------------------------ protocol
    def has(self, p: nn.Parameter) -> bool:
        return id(p) in self.offsets

    def dest(self, p: nn.Parameter) -> Optional[torch.Tensor]:
        """Where p's gradient lives: a view of the flat buffer shaped like p (None for a parameter outside the layout)."""
        v = self._views.get(id(p))
        if v is None:
            off = self.offsets.get(id(p))
            if off is None:
                return None
            v = self._views[id(p)] = self.flat[off:off + p.numel()].view(p.shape)
        g = p.grad
        if g is not None and g.data_ptr() == v.data_ptr():
            # autograd adopted the previous step's view as p.grad and the caller kept it (gradient accumulation, or
            # zero_grad(set_to_none=False)): writing the new gradient in place and then accumulating it onto itself
            # would give 2 * g_new and lose the old value
            raise RuntimeError("b200unet: p.grad still aliases the flat gradient buffer from the previous backward; "
                               "call optimizer.zero_grad(set_to_none=True) (the trainer does, train.py:634) or clone "
                               "the gradients you want to accumulate")
        return v

    def __call__(self, p: nn.Parameter, g: torch.Tensor) -> torch.Tensor:
        """Hand over a gradient.  If it was not produced in place (`dest`), it is copied into its slot."""
        v = self.dest(p)
        if v is None:
            return g
        if g.data_ptr() != v.data_ptr():
            v.copy_(g)
        self.delivered(p)
        return v

    def delivered(self, p: nn.Parameter) -> None:
        pass

    def finish(self) -> None:
        pass


def sink_of(model) -> Optional[FlatGradSink]:
    s = getattr(model, "_grad_sink", None)
    return s if isinstance(s, FlatGradSink) else None
