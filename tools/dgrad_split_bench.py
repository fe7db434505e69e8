"""Developer tool: the level-0 concat gradient as one [.,96] tensor (two channel slices) against two dense tensors.
    python tools/dgrad_split_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

B, hw = 32, 512


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


g = torch.Generator(device="cuda").manual_seed(1)
dy = torch.randn(B, hw, hw, 32, device="cuda", generator=g).bfloat16()
wt = torch.randn(32, 96, 3, 3, device="cuda", generator=g) * 0.05
_, wd = ops.pack_conv_weights(wt)
dx = torch.empty(B, hw, hw, 96, dtype=torch.bfloat16, device="cuda")
t_one = timeit(lambda: ops.conv_dgrad(dy, wd, (hw, hw), 1, out=dx))
t_two = timeit(lambda: ops.conv_dgrad_split(dy, wd, (hw, hw), 64))
d1, d2 = ops.conv_dgrad_split(dy, wd, (hw, hw), 64)
t_up_slice = timeit(lambda: ops.upsample2x_backward(dx[..., :64]))
t_up_dense = timeit(lambda: ops.upsample2x_backward(d1))
y = torch.randn(B, hw, hw, 32, device="cuda", generator=g).bfloat16()
dz = torch.randn(B, hw, hw, 32, device="cuda", generator=g).bfloat16()
a = torch.rand(B, 32, device="cuda") + 0.5
b = torch.randn(B, 32, device="cuda") * 0.1
mean, rstd, gamma = torch.zeros(B, 32, device="cuda"), torch.ones(B, 32, device="cuda"), torch.ones(32, device="cuda")
t_nb_slice = timeit(lambda: ops.in_backward(dz, dx[..., 64:], y, a, b, mean, rstd, None, gamma, 0.01))
t_nb_dense = timeit(lambda: ops.in_backward(dz, d2, y, a, b, mean, rstd, None, gamma, 0.01))
print(f"dgrad 32->96 @512: one tensor {t_one:.1f} us, two dense tensors {t_two:.1f} us")
print(f"upsample backward of the 64-channel half: slice {t_up_slice:.1f} us, dense {t_up_dense:.1f} us")
print(f"norm backward with the 32-channel skip operand: slice {t_nb_slice:.1f} us, dense {t_nb_dense:.1f} us")
print(f"sum: {t_one + t_up_slice + t_nb_slice:.1f} -> {t_two + t_up_dense + t_nb_dense:.1f} us")
