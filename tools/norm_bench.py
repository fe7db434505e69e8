"""Developer tool: time the InstanceNorm/LeakyReLU/dropout kernels per activation shape of Our_UNet (batch 32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

SHAPES = [(512, 32), (256, 64), (128, 128), (64, 256), (32, 512), (16, 512)]


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


B = 32
tf = tb = 0.0
for hw, c in SHAPES:
    y = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    dz = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    z = torch.empty_like(y)
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda")
    drop = (torch.rand(B, c, device="cuda") > 0.2).float() / 0.8
    yf = y.float()
    stats = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], -1).unsqueeze(1).contiguous()
    mean, rstd, a, b = ops.in_finalize(stats, gamma, beta, drop, 1e-5, hw * hw)
    nbytes = y.numel() * 2
    t1 = timeit(lambda: ops.in_apply(y, a, b, 0.01, out=z))
    t2 = timeit(lambda: ops.in_backward(dz, None, y, a, b, mean, rstd, drop, gamma, 0.01))
    t3 = timeit(lambda: ops.in_backward(dz, dz, y, a, b, mean, rstd, drop, gamma, 0.01))
    tf += t1
    tb += t2
    print(f"{hw:3d}^2 x {c:3d}: apply {t1 * 1e3:7.1f} us {2 * nbytes / t1 / 1e9:6.2f} TB/s | bwd {t2 * 1e3:7.1f} us "
          f"{5 * nbytes / t2 / 1e9:6.2f} TB/s | bwd+dz2 {t3 * 1e3:7.1f} us {7 * nbytes / t3 / 1e9:6.2f} TB/s")
print(f"sum apply {tf:.3f} ms, bwd {tb:.3f} ms (x the number of layers per shape)")
