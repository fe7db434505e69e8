"""Developer tool: GPU time of the norm backward at the small levels, measured by replaying a CUDA graph of 20 calls
(the per-call host overhead of ~45 us hides the kernels in a plain loop).  B200UNET_NO_NORM_FUSED=1 = three-kernel form."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

B = 32
for hw, c in [(64, 256), (32, 512), (16, 512)]:
    y = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    dz = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda")
    drop = (torch.rand(B, c, device="cuda") > 0.2).float() / 0.8
    yf = y.float()
    stats = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], -1).unsqueeze(1).contiguous()
    mean, rstd, a, b = ops.in_finalize(stats, gamma, beta, drop, 1e-5, hw * hw)
    res = []
    for dz2 in (None, dz):
        ops.in_backward(dz, dz2, y, a, b, mean, rstd, drop, gamma, 0.01)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.in_backward(dz, dz2, y, a, b, mean, rstd, drop, gamma, 0.01)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 100 * 1e3)
    print(f"{hw:3d}^2 x {c:3d}: bwd {res[0]:6.1f} us | bwd+dz2 {res[1]:6.1f} us")
