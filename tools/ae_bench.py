"""Developer tool: time the autoencoder reconstruction step (BASELINE.json configs[3]: Autoencoder + MSELoss, bf16,
batch 64, 512x512) on one B200.    python tools/ae_bench.py [--batch 64] [--steps 10]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200.models.autoencoder import Autoencoder
from unet_implementations_b200.models.losses import MSELoss

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
torch.manual_seed(1234)
model = Autoencoder(encoder_dropout_rates=[0.0, 0.0, 0.05, 0.1, 0.15, 0.15],
                    decoder_dropout_rates=[0.15, 0.1, 0.1, 0.05, 0.0]).cuda().train()
loss_fn = MSELoss()
g = torch.Generator().manual_seed(0)
x = torch.rand(args.batch, 3, 512, 512, generator=g).cuda()


def step():
    for p in model.parameters():
        p.grad = None
    loss = loss_fn(model(x), x)
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
print(f"autoencoder step: batch {args.batch}, 512x512, bf16: {ms:.2f} ms/step = {args.batch / ms * 1e3:.0f} img/s, loss {loss.item():.5f}, "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
