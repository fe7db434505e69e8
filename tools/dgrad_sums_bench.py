"""Developer tool: data gradient with / without the producer-side norm-backward sums, at the real layer shapes.
    python tools/dgrad_sums_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

B = 32
CASES = [("e0c2/d4c2 dgrad 32->32 @512 (pconv)", 32, 512), ("e1c2/d3c2 dgrad 64->64 @256 (nconv)", 64, 256)]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, c, hw in CASES:
    g = torch.Generator(device="cuda").manual_seed(1)
    dy = torch.randn(B, hw, hw, c, device="cuda", generator=g).bfloat16()
    y = torch.randn(B, hw, hw, c, device="cuda", generator=g).bfloat16()
    wt = torch.randn(c, c, 3, 3, device="cuda", generator=g) * 0.05
    _, wd = ops.pack_conv_weights(wt)
    a = torch.rand(B, c, device="cuda") + 0.5
    b = torch.randn(B, c, device="cuda") * 0.1
    dx = torch.empty(B, hw, hw, c, dtype=torch.bfloat16, device="cuda")
    t0 = timeit(lambda: ops.conv_dgrad(dy, wd, (hw, hw), 1, out=dx))
    t1 = timeit(lambda: ops.conv_dgrad(dy, wd, (hw, hw), 1, out=dx, bwd_sums=(y, a, b, 0.01)))
    mean = torch.zeros(B, c, device="cuda")
    rstd = torch.ones(B, c, device="cuda")
    gamma = torch.ones(c, device="cuda")
    _, part = ops.conv_dgrad(dy, wd, (hw, hw), 1, out=dx, bwd_sums=(y, a, b, 0.01))
    t2 = timeit(lambda: ops.in_backward(dx, None, y, a, b, mean, rstd, None, gamma, 0.01))
    t3 = timeit(lambda: ops.in_backward(dx, None, y, a, b, mean, rstd, None, gamma, 0.01, ext_part=part))
    gb = B * hw * hw * c * 2 / 1e9
    print(f"{name}: dgrad {t0:.1f} us -> with sums {t1:.1f} us (+{t1 - t0:.1f}); norm backward {t2:.1f} -> {t3:.1f} us "
          f"(-{t2 - t3:.1f}); tensor = {gb:.3f} GB, one pass at 6.4 TB/s = {gb / 6.4e-3:.1f} us")
