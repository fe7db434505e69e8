"""Developer tool for ncu: launch each kernel family ONCE at a real layer shape of Our_UNet (batch 32, 512^2 input),
so that `ncu --set full` only replays a dozen kernels with a small memory footprint.

    python tools/prof_ops.py [--batch 32] [--cases e0c2,d1c1,...]
Cases are conv layers (fprop + dgrad + wgrad each) or one of: norm512 norm128 up256 stem head loss
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

LAYERS = {
    "e0c2": (32, 32, 1, 512), "e1c1": (32, 64, 2, 512), "e1c2": (64, 64, 1, 256), "e2c1": (64, 128, 2, 256),
    "e2c2": (128, 128, 1, 128), "e3c1": (128, 256, 2, 128), "e3c2": (256, 256, 1, 64), "e4c1": (256, 512, 2, 64),
    "e4c2": (512, 512, 1, 32), "e5c1": (512, 512, 2, 32), "e5c2": (512, 512, 1, 16), "d0c1": (1024, 512, 1, 32),
    "d1c1": (768, 256, 1, 64), "d2c1": (384, 128, 1, 128), "d3c1": (192, 64, 1, 256), "d4c1": (96, 32, 1, 512),
}
DEFAULT = "e0c2,d3c1,d4c1,d1c1,e2c2,e3c1,norm512,up256,stem,head,loss"


def conv_case(name, B):
    cin, cout, s, h = LAYERS[name]
    x = torch.randn(B, h, h, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    wf, wd = ops.pack_conv_weights(w)
    oh = (h - 1) // s + 1
    dy = torch.randn(B, oh, oh, cout, device="cuda").bfloat16()
    ops.conv_fprop(x, wf, s, want_stats=True)
    ops.conv_dgrad(dy, wd, (h, h), s)
    ops.conv_wgrad(x, dy, s)
    torch.cuda.synchronize()


def norm_case(hw, c, B):
    y = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    dz = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda")
    drop = (torch.rand(B, c, device="cuda") > 0.2).float() / 0.8
    yf = y.float()
    stats = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], -1).unsqueeze(1).contiguous()
    mean, rstd, a, b = ops.in_finalize(stats, gamma, beta, drop, 1e-5, hw * hw)
    ops.in_apply(y, a, b, 0.01)
    ops.in_backward(dz, None, y, a, b, mean, rstd, drop, gamma, 0.01)
    torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cases", default=DEFAULT)
    args = ap.parse_args()
    B = args.batch
    for case in args.cases.split(","):
        if case in LAYERS:
            conv_case(case, B)
        elif case.startswith("norm"):
            hw = int(case[4:])
            norm_case(hw, 32 * 512 // hw if hw > 16 else 512, B)
        elif case.startswith("up"):
            hw = int(case[2:])
            c = 32 * 512 // hw
            x = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
            out = torch.empty(B, 2 * hw, 2 * hw, c, device="cuda", dtype=torch.bfloat16)
            a = torch.rand(B, c, device="cuda") + 0.5
            b = torch.randn(B, c, device="cuda") * 0.1
            ops.upsample2x(x, out, norm=(a, b, 0.01))  # as in the model: the producer's norm apply fused in
            ops.upsample2x_backward(out)
        elif case == "stem":
            img = torch.randn(B, 3, 512, 512, device="cuda")
            w = torch.randn(32, 3, 3, 3, device="cuda") * 0.1
            y, _ = ops.stem_fprop(img, w)
            ops.stem_wgrad(img, y)
        elif case == "head":
            z = torch.randn(B, 512, 512, 32, device="cuda").bfloat16()
            w = torch.randn(3, 32, 1, 1, device="cuda")
            bias = torch.zeros(3, device="cuda")
            lg = ops.head_forward(z, w, bias)
            ops.head_backward(lg, z, w)
        elif case == "loss":
            lg = torch.randn(B, 3, 512, 512, device="cuda")
            t = torch.randint(0, 3, (B, 512, 512), device="cuda")
            out, tables = ops.loss_forward(lg, t, None, True, 1.0, 1.0, 255, 1e-5)
            ops.loss_backward(lg, t, tables, None, 1.0, 1.0, 255)
        else:
            raise SystemExit(f"unknown case {case}")
        torch.cuda.synchronize()
        print("done", case, flush=True)


if __name__ == "__main__":
    main()
