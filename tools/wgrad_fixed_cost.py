"""Developer tool: fixed cost against per-image cost of the weight-gradient kernels (time over batch size).
    python tools/wgrad_fixed_cost.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from unet_implementations_b200 import ops
from wgrad_pairs_bench import timeit  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)
CASES = ((32, 32, 512, 1), (64, 64, 256, 1), (192, 64, 256, 1), (96, 32, 512, 1), (128, 128, 128, 1), (384, 128, 128, 1),
         (256, 256, 64, 1), (768, 256, 64, 1), (512, 512, 32, 1), (1024, 512, 32, 1), (512, 512, 16, 1), (128, 256, 128, 2))
for cin, cout, hw, stride in CASES:
    ts = []
    for B in (8, 16, 32, 64):
        if B * hw * hw * max(cin, cout) * 2 > (6 << 30):
            continue
        x = torch.randn(B, hw, hw, cin, device="cuda", generator=g).bfloat16()
        dy = torch.randn(B, hw // stride, hw // stride, cout, device="cuda", generator=g).bfloat16()
        ts.append((B, timeit(lambda: ops.conv_wgrad(x, dy, stride))))
        del x, dy
    slope = (ts[-1][1] - ts[-2][1]) / (ts[-1][0] - ts[-2][0])
    t32 = dict(ts).get(32)
    print(f"{cin}->{cout} s{stride} @{hw}: " + "  ".join(f"B={b}: {t:.1f}" for b, t in ts) +
          f" us  => {slope:.2f} us/image, fixed {ts[-1][1] - slope * ts[-1][0]:.1f} us ({100 * (ts[-1][1] - slope * ts[-1][0]) / t32:.0f} % at B=32)")
