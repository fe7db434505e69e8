"""Developer tool: variants of the narrow-output weight gradient, same process, alternating, against each other and
(small shapes) against F.conv2d's weight gradient in fp64.
  B200UNET_WGRAD_PAIRS=1  32 -> 32 on pixel-pair rows (wgradn<64,64> on the (W/2, 64) views) instead of wgradn<32,32>
  B200UNET_WGRAD_ONEDY=1  one (16+2)-pixel dY tile per stage, the three kw shifts as N units one row apart (=2: with the
                          descriptor's base-offset field set)
    python tools/wgrad_pairs_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200 import ops


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


def run(mode, x, dy):
    os.environ["B200UNET_WGRAD_PAIRS"], os.environ["B200UNET_WGRAD_ONEDY"] = mode
    return ops.conv_wgrad(x, dy, 1)


if __name__ == "__main__":
    g = torch.Generator(device="cuda").manual_seed(3)
    only = sys.argv[1:]  # e.g. "10 11" to restrict the modes
    for cin, cout, B, H, W in ((32, 32, 2, 24, 40), (32, 32, 1, 9, 34), (64, 64, 2, 24, 40), (64, 64, 1, 9, 33),
                               (192, 64, 1, 20, 24), (32, 32, 32, 512, 512), (64, 64, 32, 256, 256), (192, 64, 32, 256, 256)):
        modes = ["00", "10", "11", "20"] if cin == 32 else ["00", "01"]
        modes = [m for m in modes if not only or m in only or m == "00"]
        x = torch.randn(B, H, W, cin, device="cuda", generator=g).bfloat16()
        dy = torch.randn(B, H, W, cout, device="cuda", generator=g).bfloat16()
        res = {m: run(m, x, dy) for m in modes}
        ref = None
        if B * H * W <= 1 << 16:
            xr = x.float().permute(0, 3, 1, 2).double()
            dyr = dy.float().permute(0, 3, 1, 2).double()
            w = torch.zeros(cout, cin, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
            (torch.nn.functional.conv2d(xr, w, padding=1) * dyr).sum().backward()
            ref = w.grad
        for m in modes:
            line = f"{cin}->{cout} {B}x{H}x{W} mode {m}: vs mode 00 {float((res[m] - res['00']).abs().max() / res['00'].abs().max()):.2e}"
            if ref is not None:
                line += f", vs fp64 {float((res[m] - ref).abs().max() / ref.abs().max()):.2e}"
            line += f", deterministic {bool(torch.equal(res[m], run(m, x, dy)))}"
            print(line)
        if B == 32:
            for rep in range(2):
                print("   " + "   ".join(f"mode {m}: {timeit(lambda: run(m, x, dy)):.1f} us" for m in modes))
