"""Fused optimizer step for the reference's trainer: `torch.optim.SGD(lr, momentum, nesterov=True, weight_decay)`
(Our_UNet/src/train.py:431-451) as ONE multi-tensor CUDA launch over all parameters.

`FusedSGD` is a `torch.optim.Optimizer`: `param_groups[i]["lr"]` is what `LambdaLR` (train.py:454-477) drives, the
per-parameter state is `momentum_buffer` as in torch, so `state_dict()` is interchangeable with `torch.optim.SGD`.
The arithmetic mirrors torch's foreach implementation on CUDA operation by operation (bit-exact, tests/test_gpu_aux.py).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if dampening != 0.0:
            raise ValueError("FusedSGD: dampening is not supported (the trainer uses 0)")
        if nesterov and momentum <= 0:
            raise ValueError("Nesterov momentum requires a momentum")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            mom = float(group["momentum"])
            # parameters seen for the first time get buf = g (torch: clone of the decayed gradient); they go in their own launch
            if mom != 0:
                batches = [(True, [p for p in ps if "momentum_buffer" not in self.state[p]]),
                           (False, [p for p in ps if "momentum_buffer" in self.state[p]])]
            else:
                batches = [(True, ps)]
            for first, sel in batches:
                if not sel:
                    continue
                for p in sel:
                    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                            and p.grad.dtype == torch.float32 and not p.grad.is_sparse):
                        raise RuntimeError("FusedSGD: parameters and gradients must be contiguous fp32 CUDA tensors")
                    if mom != 0 and first:
                        self.state[p]["momentum_buffer"] = torch.empty_like(p)
                n = len(sel)
                PT = ctypes.c_void_p * n
                params = PT(*[p.data_ptr() for p in sel])
                grads = PT(*[p.grad.data_ptr() for p in sel])
                bufs = PT(*[self.state[p]["momentum_buffer"].data_ptr() for p in sel]) if mom != 0 else None
                numels = (ctypes.c_int64 * n)(*[p.numel() for p in sel])
                with torch.cuda.device(sel[0].device):
                    _lib.call("b200unet_sgd_nesterov_step", params, grads, bufs, numels, n, float(group["lr"]), mom,
                              float(group["weight_decay"]), int(bool(group["nesterov"])), int(first and mom != 0),
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return loss
