"""Developer tool: head forward / backward (with the fused norm apply and the norm-backward sums) at batch 32, 512^2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

B, hw, c = 32, 512, 32
g = torch.Generator(device="cuda").manual_seed(1)
y = torch.randn(B, hw, hw, c, device="cuda", generator=g).bfloat16()
a = torch.rand(B, c, device="cuda") + 0.5
b = torch.randn(B, c, device="cuda") * 0.1
w = torch.randn(3, c, 1, 1, device="cuda") * 0.1
bias = torch.zeros(3, device="cuda")
dl = torch.randn(B, 3, hw, hw, device="cuda", generator=g)


def timeit(fn, n=15):
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


print(f"head_norm_fwd {timeit(lambda: ops.head_forward(y, w, bias, norm=(a, b, 0.01))):.1f} us (637 MB)")
print(f"head_norm_bwd {timeit(lambda: ops.head_backward(dl, y, w, norm=(a, b, 0.01))):.1f} us (1174 MB)")
print(f"head_norm_bwd + norm-backward sums {timeit(lambda: ops.head_backward(dl, y, w, norm=(a, b, 0.01), want_bwd_part=True)):.1f} us")
