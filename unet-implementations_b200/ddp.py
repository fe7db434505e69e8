"""Batch-sharded data parallelism for the UNet step: one process per GPU, identical replicas, one logical
all-reduce (mean) of the 19.66 M-element gradient per step, bucketed and overlapped with backward.

The reference has no distributed code; InstanceNorm and SpatialDropout are per-(sample, channel),
so batch sharding needs no collective in the forward.  `SimpleLoss` normalises over the LOCAL batch
(losses.py:44-60, :118), so the all-reduced gradient is the mean over ranks of the per-rank oracle gradients --
exactly what stock DistributedDataParallel around the reference would compute.

Mechanism: `UNet`'s fused backward hands every parameter gradient to `model._grad_sink(param, grad)` in the order
it produces them (head, decoder 4..0, bottleneck, encoder 4..0).  The sink copies it into one flat fp32 buffer laid
out in that order; when the last gradient of a bucket has arrived the bucket is all-reduced asynchronously
(NCCL runs it on its own stream, ordered after the producing kernels) while backward keeps launching dgrad/wgrad
kernels.  `finish()` -- called at the end of backward -- makes the compute stream wait for the outstanding buckets.
The gradients autograd returns are views of the flat buffer.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


def backward_param_order(model) -> List[nn.Parameter]:
    """Parameters in the order UNet's backward produces their gradients (models/unet.py:_backward_impl)."""
    order: List[nn.Parameter] = []
    head = model.segmentation_output
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
    return order


class BucketedGradAllReduce:
    """Gradient sink: flat fp32 buffer + bucketed asynchronous all-reduce (mean) over `group`."""

    def __init__(self, model, group=None, bucket_bytes: int = 16 << 20, device: Optional[torch.device] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else "none"
        params = [p for p in backward_param_order(model) if p.requires_grad]
        dev = device or params[0].device
        self.offsets: Dict[int, int] = {}
        self.bucket_of: Dict[int, int] = {}
        self.buckets: List[List[int]] = []  # [start, end) element ranges
        off, start, last_ids = 0, 0, []
        self.last_param_of_bucket: Dict[int, int] = {}
        for p in params:
            self.offsets[id(p)] = off
            self.bucket_of[id(p)] = len(self.buckets)
            off += (p.numel() + 3) // 4 * 4  # keep every gradient 16-byte aligned
            last = id(p)
            if (off - start) * 4 >= bucket_bytes:
                self.buckets.append([start, off])
                self.last_param_of_bucket[last] = len(self.buckets) - 1
                start = off
        if off > start:
            self.buckets.append([start, off])
            self.last_param_of_bucket[id(params[-1])] = len(self.buckets) - 1
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        # gradients may arrive slightly out of production order (weight gradients run on a side stream and are
        # delivered one layer later): a bucket is reduced when ALL of its parameters have arrived
        self.bucket_size: List[int] = [0] * len(self.buckets)
        for p in params:
            b = 0
            while not (self.buckets[b][0] <= self.offsets[id(p)] < self.buckets[b][1]):
                b += 1
            self.bucket_of[id(p)] = b
            self.bucket_size[b] += 1
        self.arrived: List[int] = [0] * len(self.buckets)
        self.works: list = []
        self.numel = off
        model._grad_sink = self

    def __call__(self, p: nn.Parameter, g: torch.Tensor) -> torch.Tensor:
        off = self.offsets[id(p)]
        view = self.flat[off:off + p.numel()].view_as(p)
        view.copy_(g)
        b = self.bucket_of[id(p)]
        self.arrived[b] += 1
        if self.arrived[b] == self.bucket_size[b] and self.world > 1:
            lo, hi = self.buckets[b]
            chunk = self.flat[lo:hi]
            if self.backend == "nccl":
                self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return view

    def finish(self):
        """Block the current stream (not the host) until every outstanding bucket has been reduced."""
        for w in self.works:
            w.wait()
        if self.works and self.backend != "nccl":
            self.flat.mul_(1.0 / self.world)
        self.works = []
        self.arrived = [0] * len(self.buckets)


def broadcast_parameters(model, src: int = 0, group=None):
    """Make every replica start from rank `src`'s weights (identical seeds already give identical weights; this is
    the belt-and-braces equivalent of DDP's constructor broadcast)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)
