"""Developer tool: time the model variants of BASELINE.json configs[3] and configs[4] on one B200.
    python tools/variant_bench.py ae   [--batch 64]   Autoencoder + MSELoss reconstruction step, bf16, 512x512
    python tools/variant_bench.py clip [--batch 32]   CLIP-conditioned UNet + SimpleLoss, bf16, 512x512, with a seeded
                                                      random [B,512,16,16] patch-feature tensor (the frozen CLIP ViT-B/16
                                                      encoder is a third-party model and not part of the timed step)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200.models.losses import MSELoss, SimpleLoss

ap = argparse.ArgumentParser()
ap.add_argument("variant", choices=["ae", "clip"])
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
B = args.batch or (64 if args.variant == "ae" else 32)
torch.manual_seed(1234)
g = torch.Generator().manual_seed(0)
if args.variant == "ae":
    from unet_implementations_b200.models.autoencoder import Autoencoder
    model = Autoencoder(encoder_dropout_rates=[0.0, 0.0, 0.05, 0.1, 0.15, 0.15],
                        decoder_dropout_rates=[0.15, 0.1, 0.1, 0.05, 0.0]).cuda().train()
    loss_fn = MSELoss()
    x = torch.rand(B, 3, 512, 512, generator=g).cuda()

    def fwd_loss():
        return loss_fn(model(x), x)
else:
    from unet_implementations_b200.models.clip_unet import UNet
    model = UNet().cuda().train()
    loss_fn = SimpleLoss()
    x = torch.randn(B, 3, 512, 512, generator=g).cuda()
    clip = torch.randn(B, 512, 16, 16, generator=g).cuda()
    t = torch.randint(0, 3, (B, 512, 512), generator=g).cuda()

    def fwd_loss():
        return loss_fn(model(x, clip), t)


def step():
    for p in model.parameters():
        p.grad = None
    loss = fwd_loss()
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f"{args.variant} step: batch {B}, 512x512, bf16: {ms:.2f} ms/step = {B / ms * 1e3:.0f} img/s, loss {loss.item():.5f}, "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
