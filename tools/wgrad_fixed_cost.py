"""Developer tool: fixed cost against per-pixel cost of the narrow weight-gradient kernels (time over batch size).
    python tools/wgrad_fixed_cost.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops
from wgrad_pairs_bench import timeit  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)
for cin, cout, hw, modes in ((32, 32, 512, ("00", "11")), (64, 64, 256, ("01",)), (192, 64, 256, ("01",)), (96, 32, 512, ("00",))):
    for m in modes:
        os.environ["B200UNET_WGRAD_PAIRS"], os.environ["B200UNET_WGRAD_ONEDY"] = m
        ts = []
        for B in (4, 8, 16, 32):
            x = torch.randn(B, hw, hw, cin, device="cuda", generator=g).bfloat16()
            dy = torch.randn(B, hw, hw, cout, device="cuda", generator=g).bfloat16()
            ts.append((B, timeit(lambda: ops.conv_wgrad(x, dy, 1))))
            del x, dy
        slope = (ts[-1][1] - ts[-2][1]) / (ts[-1][0] - ts[-2][0])
        print(f"{cin}->{cout} @{hw} mode {m}: " + "  ".join(f"B={b}: {t:.1f} us" for b, t in ts) +
              f"   => {slope:.2f} us/image, fixed {ts[-1][1] - slope * ts[-1][0]:.1f} us")
