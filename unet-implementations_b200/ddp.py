"""Batch-sharded data parallelism for the UNet step: one process per GPU, identical replicas, one logical
all-reduce (mean) of the 19.66 M-element gradient per step, bucketed and overlapped with backward.

The reference has no distributed code; InstanceNorm and SpatialDropout are per-(sample, channel),
so batch sharding needs no collective in the forward.  `SimpleLoss` normalises over the LOCAL batch
(losses.py:44-60, :118), so the all-reduced gradient is the mean over ranks of the per-rank oracle gradients --
exactly what stock DistributedDataParallel around the reference would compute.

Mechanism: `UNet`'s fused backward has every gradient-producing kernel write straight into one flat fp32 buffer laid
out in production order (flat.FlatGradSink: head, decoder 4..0, bottleneck, encoder 4..0 -- no per-gradient copy);
when the last gradient of a bucket has been delivered the bucket is all-reduced asynchronously (NCCL runs it on its
own stream, ordered after the producing kernels) while backward keeps launching dgrad/wgrad kernels.  `finish()` --
called at the end of backward -- reduces whatever is left and makes the compute stream wait for the outstanding
buckets.  The gradients autograd returns are views of the flat buffer.

NCCL and the persistent conv kernels.  The tensor-core conv kernels run one CTA per SM with ~226 KB of shared memory;
an NCCL kernel that lands on an SM delays that SM's CTA, and a persistent kernel is as slow as its slowest CTA
(round 1: dgrad +11 %, wgrad +4 % at 8 GPUs with NCCL's default channel count).  Two knobs, measured on 8 x B200
(round 2, ms per step, the slowest rank's compute-only step 21.8 ms): communicator capped to 16 CTAs and no SMs
reserved 22.29; 4 CTAs + 4 SMs reserved in backward 22.44; 8 + 8: 22.47; 2 + 2: 22.58.  Reserving SMs costs the conv
kernels of backward that share of the device (2.7 % of 9 ms) and buys nothing once the communicator is small, so the
default is `DEFAULT_NCCL_CTAS` CTAs (`nccl_options`, NCCL_MAX_CTAS) and no reservation; `B200UNET_RESERVED_SMS=n`
sizes the conv grids of BACKWARD -- the only phase with an all-reduce in flight -- for n fewer SMs
(b200unet_set_reserved_sms).  The last bucket, whose all-reduce cannot overlap anything, holds only the small
gradients produced last (`tail_bytes`).  What remains at 8 GPUs (~0.45 ms over the slowest rank's own step) is the
exposed tail plus the per-step lock-step jitter of eight power-capped GPUs; the spread between the ranks' own step
times (21.3 .. 21.9 ms) is the other half of the distance to 8 x one GPU.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .flat import FlatGradSink, backward_param_order  # noqa: F401  (re-exported)

DEFAULT_NCCL_CTAS = 16


def nccl_options(max_ctas: int = DEFAULT_NCCL_CTAS):
    """ProcessGroupNCCL options that cap the communicator at `max_ctas` CTAs (ncclConfig.maxCTAs)."""
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True  # the few NCCL CTAs go first when a persistent conv kernel's CTAs retire
    try:
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = 1
    except AttributeError:  # a torch build without ncclConfig plumbing: the environment variable does the same
        pass
    return opts


def init_process_group(device: torch.device, max_ctas: Optional[int] = None, reserve_sms: Optional[int] = None):
    """`dist.init_process_group("nccl")` for one process per GPU with the communicator capped to `max_ctas` CTAs and
    (optionally) `reserve_sms` SMs kept free by the conv kernels of backward.  Environment overrides:
    B200UNET_NCCL_CTAS, B200UNET_RESERVED_SMS."""
    ctas = int(os.environ.get("B200UNET_NCCL_CTAS", max_ctas if max_ctas is not None else DEFAULT_NCCL_CTAS))
    os.environ.setdefault("NCCL_MAX_CTAS", str(ctas))
    os.environ.setdefault("NCCL_MIN_CTAS", "1")
    if "B200UNET_RESERVED_SMS" not in os.environ:
        os.environ["B200UNET_RESERVED_SMS"] = str(reserve_sms if reserve_sms is not None else 0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=device, pg_options=nccl_options(ctas))
    return ctas


class BucketedGradAllReduce(FlatGradSink):
    """Gradient sink: flat fp32 buffer + bucketed asynchronous all-reduce (mean) over `group`."""

    def __init__(self, model, group=None, bucket_bytes: int = 16 << 20, device: Optional[torch.device] = None,
                 tail_bytes: int = 2 << 20):
        super().__init__(model, device=device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else "none"
        # SMs the backward's conv grids leave free for the NCCL kernels (read by models/unet.py:_UNetFunction.backward)
        self.reserve_sms = int(os.environ.get("B200UNET_RESERVED_SMS", "0")) if (self.world > 1 and self.backend == "nccl") else 0
        # The LAST bucket cannot start before backward ends, so its all-reduce is exposed: it is kept small (the
        # gradients produced last are the small ones -- encoder stages 2..0, 0.3 M of the 19.7 M elements), everything
        # before it goes in buckets of `bucket_bytes`.
        ends = [self.offsets[id(p)] + (p.numel() + 3) // 4 * 4 for p in self.params]
        tail_start_idx = len(self.params)
        while tail_start_idx > 1 and (self.numel - ends[tail_start_idx - 2]) * 4 <= tail_bytes:
            tail_start_idx -= 1
        self.buckets: List[List[int]] = []  # [start, end) element ranges
        self.bucket_of: Dict[int, int] = {}
        start = 0
        for i, p in enumerate(self.params):
            end = ends[i]
            self.bucket_of[id(p)] = len(self.buckets)
            if (end - start) * 4 >= bucket_bytes or i == tail_start_idx - 1:
                self.buckets.append([start, end])
                start = end
        if self.numel > start:
            self.buckets.append([start, self.numel])
        # gradients may arrive slightly out of production order (weight gradients run on a side stream and are
        # delivered one layer later): a bucket is reduced when ALL of its parameters have arrived
        self.bucket_size: List[int] = [0] * len(self.buckets)
        for p in self.params:
            self.bucket_size[self.bucket_of[id(p)]] += 1
        self.arrived: List[int] = [0] * len(self.buckets)
        self.reduced: List[bool] = [False] * len(self.buckets)
        self.works: list = []
        self.enabled = True  # False = accumulate locally (no_sync)

    def _reduce(self, b: int) -> None:
        self.reduced[b] = True
        if self.world <= 1 or not self.enabled:
            return
        lo, hi = self.buckets[b]
        chunk = self.flat[lo:hi]
        op = dist.ReduceOp.AVG if self.backend == "nccl" else dist.ReduceOp.SUM
        self.works.append(dist.all_reduce(chunk, op=op, group=self.group, async_op=True))

    def delivered(self, p: nn.Parameter) -> None:
        b = self.bucket_of[id(p)]
        self.arrived[b] += 1
        if self.arrived[b] == self.bucket_size[b]:
            self._reduce(b)

    def finish(self) -> None:
        """Reduce the buckets that did not fill up (a parameter without a gradient this step -- a frozen layer, an
        unused fusion conv -- must not leave its bucket mates un-reduced), then block the current stream (not the
        host) until every outstanding bucket has been reduced."""
        for b in range(len(self.buckets)):
            if not self.reduced[b] and self.arrived[b] > 0:
                self._reduce(b)
        for w in self.works:
            w.wait()
        if self.works and self.backend != "nccl":
            self.flat.mul_(1.0 / self.world)
        self.works = []
        self.arrived = [0] * len(self.buckets)
        self.reduced = [False] * len(self.buckets)


def broadcast_parameters(model, src: int = 0, group=None):
    """Make every replica start from rank `src`'s weights (identical seeds already give identical weights; this is
    the belt-and-braces equivalent of DDP's constructor broadcast)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)
