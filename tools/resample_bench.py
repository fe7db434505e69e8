"""Developer tool: time the bilinear-2x kernels (forward with the fused norm apply, backward) at the five decoder
shapes of the default model, batch 32, writing into / reading from the channel slice of a concat buffer as the model
does.  Prints time, algorithmic bytes and GB/s.   python tools/resample_bench.py [--batch 32] [--iters 20]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

SHAPES = [(256, 64, 32), (128, 128, 64), (64, 256, 128), (32, 512, 256), (16, 512, 512)]  # (hw, C, skip channels)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    B = args.batch
    tot = {"fwd": 0.0, "bwd": 0.0}
    for hw, c, skip in SHAPES:
        x = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
        cat = torch.empty(B, 2 * hw, 2 * hw, c + skip, device="cuda", dtype=torch.bfloat16)
        a = torch.rand(B, c, device="cuda") + 0.5
        b = torch.randn(B, c, device="cuda") * 0.1
        dx = torch.empty_like(x)
        nbytes = x.numel() * 2 + B * 4 * hw * hw * c * 2
        tf = timed(lambda: ops.upsample2x(x, cat[..., :c], norm=(a, b, 0.01)), args.iters)
        cat.normal_()
        tb = timed(lambda: ops.upsample2x_backward(cat[..., :c], out=dx), args.iters)
        tot["fwd"] += tf
        tot["bwd"] += tb
        print(f"{hw:4d}^2 x{c:4d} -> cat pitch {c + skip:5d} | fwd {tf:7.1f} us {nbytes / tf / 1e3:6.0f} GB/s | "
              f"bwd {tb:7.1f} us {nbytes / tb / 1e3:6.0f} GB/s")
    print("total", {k: f"{v / 1e3:.3f} ms" for k, v in tot.items()})


if __name__ == "__main__":
    main()
