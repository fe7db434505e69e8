"""Checker-side tool (uses oracle/, so it lives under tests/): the SAME-BOX competitor of SURVEY.md 8(d) -- the
reference's op sequence (oracle/unet_oracle.py: F.conv2d / F.instance_norm / F.leaky_relu / F.interpolate / the
SimpleLoss restatement, i.e. stock PyTorch + cuDNN kernels) run on the B200 itself, fp32 eager and under
torch.autocast(bf16), forward + loss + backward at the benchmark configuration (batch 32, 512^2, default UNet).
Not part of bench.py: it reports what unmodified PyTorch gets out of the same GPU.
    python tests/tools/stock_torch_gpu.py [--batch 32] [--steps 5]  -> gpurun_out/stock_torch_gpu.json"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import unet_oracle as O
from unet_implementations_b200.models.unet import UNet


def run(sd, x, target, cfg, steps, autocast, channels_last):
    masks = O.draw_dropout_masks(cfg, x.shape[0], x)

    def step():
        xx = x.contiguous(memory_format=torch.channels_last) if channels_last else x
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return O.training_step(sd, xx, target, cfg, masks=masks)
        return O.training_step(sd, xx, target, cfg, masks=masks)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(ms_per_step=ms, img_per_s=x.shape[0] / ms * 1e3, loss=float(out["loss"]),
                peak_gib=torch.cuda.max_memory_allocated() / 2 ** 30)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    torch.manual_seed(1234)
    model = UNet()
    cfg = O.config_of(model)
    sd = {k: v.detach().clone().cuda() for k, v in model.state_dict().items()}
    x, target = O.synthetic_batch(args.batch, args.size, seed=0)
    x, target = x.cuda(), target.cuda()
    torch.backends.cudnn.benchmark = True
    res = {"config": dict(batch=args.batch, size=args.size, steps=args.steps, torch=torch.__version__,
                          cudnn=torch.backends.cudnn.version())}
    for name, ac, cl in (("fp32_eager", False, False), ("bf16_autocast", True, False),
                         ("bf16_autocast_channels_last", True, True)):
        try:
            torch.cuda.reset_peak_memory_stats()
            res[name] = run(sd, x, target, cfg, args.steps, ac, cl)
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": repr(e)[:300]}
        print(name, res[name], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/stock_torch_gpu.json", "w"), indent=1)


if __name__ == "__main__":
    main()
