"""Developer tool: time the conv kernels of every distinct layer shape of Our_UNet at batch B, 512^2 input.
    python tools/conv_bench.py [--batch 32] [--only fprop,dgrad,wgrad] [--layers 1,20] [--nostats]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

# (name, Cin, Cout, stride, Hin)
LAYERS = [
    ("e0c2", 32, 32, 1, 512), ("e1c1", 32, 64, 2, 512), ("e1c2", 64, 64, 1, 256), ("e2c1", 64, 128, 2, 256),
    ("e2c2", 128, 128, 1, 128), ("e3c1", 128, 256, 2, 128), ("e3c2", 256, 256, 1, 64), ("e4c1", 256, 512, 2, 64),
    ("e4c2", 512, 512, 1, 32), ("e5c1", 512, 512, 2, 32), ("e5c2", 512, 512, 1, 16), ("d0c1", 1024, 512, 1, 32),
    ("d0c2", 512, 512, 1, 32), ("d1c1", 768, 256, 1, 64), ("d1c2", 256, 256, 1, 64), ("d2c1", 384, 128, 1, 128),
    ("d2c2", 128, 128, 1, 128), ("d3c1", 192, 64, 1, 256), ("d3c2", 64, 64, 1, 256), ("d4c1", 96, 32, 1, 512),
    ("d4c2", 32, 32, 1, 512),
]


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--only", default="fprop,dgrad,wgrad")
    ap.add_argument("--layers", default="")
    ap.add_argument("--nostats", action="store_true")
    args = ap.parse_args()
    only = args.only.split(",")
    sel = set(args.layers.split(",")) if args.layers else None
    B = args.batch
    tot = {k: 0.0 for k in only}
    totf = 0.0
    for name, cin, cout, s, h in LAYERS:
        if sel and name not in sel:
            continue
        x = torch.randn(B, h, h, cin, device="cuda").bfloat16()
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        wf, wd = ops.pack_conv_weights(w)
        oh = (h - 1) // s + 1
        dy = torch.randn(B, oh, oh, cout, device="cuda").bfloat16()
        y = torch.empty(B, oh, oh, cout, device="cuda", dtype=torch.bfloat16)
        dx = torch.empty(B, h, h, cin, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * B * oh * oh * cout * cin * 9
        totf += flops
        line = f"{name:5s} {cin:4d}->{cout:3d} s{s} {h:3d}^2  {flops / 1e9:7.1f} GF "
        if "fprop" in only:
            t = timeit(lambda: ops.conv_fprop(x, wf, s, out=y, want_stats=not args.nostats))
            tot["fprop"] += t
            line += f"| fprop {t * 1e3:7.1f} us {flops / t / 1e9:6.0f} TF/s "
        if "dgrad" in only:
            ws2 = ops.pack_s2_dgrad_weights(wd) if s == 2 else None  # parity-stacked single launch (Cin <= 64)
            if ws2 is not None:
                t = timeit(lambda: ops.conv_dgrad_s2(dy, ws2, (h, h), out=dx))
            else:
                t = timeit(lambda: ops.conv_dgrad(dy, wd, (h, h), s, out=dx))
            tot["dgrad"] += t
            line += f"| dgrad {t * 1e3:7.1f} us {flops / t / 1e9:6.0f} TF/s "
        if "wgrad" in only:
            t = timeit(lambda: ops.conv_wgrad(x, dy, s))
            tot["wgrad"] += t
            line += f"| wgrad {t * 1e3:7.1f} us {flops / t / 1e9:6.0f} TF/s "
        print(line, flush=True)
        del x, dy, y, dx
    print("total", {k: f"{v:.3f} ms = {totf / v / 1e9:.0f} TF/s" for k, v in tot.items() if v > 0})


if __name__ == "__main__":
    main()
