"""Developer probe: is there anything to gain from running two half-batch chains side by side, so
that tensor-bound and HBM-bound kernels overlap?  Times fwd + loss + bwd of (a) one model at batch 32 and (b) two model
replicas at batch 16 each on two CUDA streams, launches interleaved by the host.  Same total work; (b) is exact for
this network (InstanceNorm and dropout are per-sample).    python tools/overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_implementations_b200.models.losses import SimpleLoss
from unet_implementations_b200.models.unet import UNet


def main():
    torch.manual_seed(0)
    B, S, steps = 32, 512, 8
    x = torch.randn(B, 3, S, S, device="cuda")
    t = torch.randint(0, 3, (B, S, S), device="cuda")
    loss_fn = SimpleLoss()
    m = UNet().cuda().train()
    m2 = [UNet().cuda().train() for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]

    def full():
        m.zero_grad(set_to_none=True)
        loss_fn(m(x), t).backward()

    def halves():
        cur = torch.cuda.current_stream()
        for i in range(2):
            streams[i].wait_stream(cur)
        outs = []
        for i in range(2):  # forward of both halves queued first, then both backwards
            with torch.cuda.stream(streams[i]):
                m2[i].zero_grad(set_to_none=True)
                outs.append(loss_fn(m2[i](x[16 * i:16 * i + 16]), t[16 * i:16 * i + 16]))
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                outs[i].backward()
        for i in range(2):
            cur.wait_stream(streams[i])

    for name, fn in (("one chain, batch 32", full), ("two chains, batch 16 + 16 on two streams", halves),
                     ("one chain, batch 32", full)):
        for _ in range(25):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:45s} {e0.elapsed_time(e1) / steps:7.2f} ms/step", flush=True)


if __name__ == "__main__":
    main()
