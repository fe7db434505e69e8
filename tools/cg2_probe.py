"""Developer tool: the CTA-pair (cta_group::2) configurations of gconv_kernel against the direct CUDA-core kernels, with
the error broken down by M tile of the pair and by half of the output channels (= which CTA's rows of B), so that a
wrong operand mapping can be read off one run.    python tools/cg2_probe.py --case N   (one case per process)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops

# (batch, H, W, Cin, Cout, what)
CASES = [
    (1, 16, 16, 256, 256, "fprop"),   # one pair tile, N = 256
    (2, 32, 32, 256, 256, "fprop"),   # 16 pair tiles, one per cluster
    (32, 64, 64, 256, 256, "fprop"),  # 7 pair tiles per cluster: rings and TMEM buffers wrap
    (32, 64, 64, 768, 256, "dgrad"),  # N = 192 x 4
    (4, 128, 128, 128, 128, "fprop"),  # N = 128, two M tiles per CTA
    (2, 256, 256, 192, 64, "fprop"),  # N = 64, four M tiles per CTA
    (2, 256, 256, 192, 64, "dgrad"),  # N = 192
    (32, 16, 16, 512, 512, "fprop"),  # two N tiles
]


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, required=True)
    a = ap.parse_args()
    n, h, w, cin, cout, what = CASES[a.case]
    g = torch.Generator(device="cuda").manual_seed(5 + a.case)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    wf, wd = ops.pack_conv_weights(wt)
    if what == "fprop":
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        y, stats = ops.conv_fprop(x, wf, 1)
        torch.cuda.synchronize()
        ref, _ = ops.conv_fprop(x, wf, 1, want_stats=False, simt=True)
        yf = y.float()
        s_ref = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1)
        serr = rel(stats.sum(dim=1), s_ref)
        nch = cout
    else:
        dy = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
        y = ops.conv_dgrad(dy, wd, (h, w), 1)
        torch.cuda.synchronize()
        ref = ops.conv_dgrad(dy, wd, (h, w), 1, simt=True)
        serr = 0.0
        nch = cin
    e = rel(y.float(), ref.float())
    print(f"case {a.case} {what} {n}x{h}x{w} {cin}->{cout}: rel-L2 {e:.3e}, stats {serr:.3e} -> {'OK' if e < 2e-3 and serr < 1e-4 else 'MISMATCH'}")
    if e >= 2e-3:
        bn = 256 if nch % 256 == 0 else (192 if nch % 192 == 0 else (128 if nch % 128 == 0 else 64))
        mt = {256: 1, 192: 1, 128: 2, 64: 4}[bn]
        th = 16 * mt
        for r in range(2):       # tiles of a pair are vertically adjacent when the image has >= 2 tile rows
            for half in range(2):
                rows = [i for i in range(h) if ((i // th) % 2 == r)] if h > th else None
                if rows is not None:
                    ys, rs = y[:, rows], ref[:, rows]
                else:            # one tile row: the pair is two horizontally adjacent 8-pixel tiles
                    cols = [j for j in range(w) if ((j // 8) % 2 == r)]
                    ys, rs = y[:, :, cols], ref[:, :, cols]
                c0 = [c for c in range(nch) if ((c % bn) // (bn // 2)) == half]
                print(f"   M tile of rank {r}, channel half {half}: rel-L2 {rel(ys[..., c0].float(), rs[..., c0].float()):.3e}"
                      f"  |y| {float(ys[..., c0].float().abs().mean()):.3e} |ref| {float(rs[..., c0].float().abs().mean()):.3e}")
        # the same against the reference with the channel halves swapped
        sw = torch.cat([ref[..., bn // 2:bn], ref[..., :bn // 2]], dim=-1) if nch == bn else None
        if sw is not None:
            print(f"   against the reference with its channel halves swapped: {rel(y.float(), sw.float()):.3e}")


if __name__ == "__main__":
    main()
