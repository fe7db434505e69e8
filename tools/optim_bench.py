"""Developer tool: the flat fused optimizer step (FusedSGD(model=...): one launch, emits the bf16 conv packs) of the
default UNet, timed alone.   python tools/optim_bench.py     (B200UNET_LIB=<other build> for an A/B)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200.models.unet import UNet
from unet_implementations_b200.flat import sink_of
from unet_implementations_b200.optim import FusedSGD

torch.manual_seed(0)
model = UNet().cuda().train()
opt = FusedSGD(model.parameters(), lr=0.01, momentum=0.99, nesterov=True, weight_decay=1e-4, model=model)
sink = sink_of(model)
sink.flat.normal_(0, 1e-3)  # a gradient for every parameter, in place in the flat buffer (what backward leaves)
for p in sink.params:
    p.grad = sink.dest(p)
n = sum(p.numel() for p in sink.params)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
ts = []
for _ in range(15):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opt.step()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"FusedSGD(model=...).step(): {ts[len(ts) // 2]:.1f} us for {n / 1e6:.2f} M parameters "
      f"({n * 20 / 1e6:.0f} MB of fp32 traffic + {n * 4.5 / 1e6:.0f} MB of bf16 packs)")
