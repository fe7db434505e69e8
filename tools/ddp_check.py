"""2+ rank NCCL check of the data-parallel path, run under torchrun on N GPUs of one box:
the gradient every rank holds after the bucketed, overlapped all-reduce == mean over ranks of the gradients the SAME
ranks compute locally (no reducer) on their own shards, bit-for-bit across ranks and within fp32 rounding of the mean.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from unet_implementations_b200 import ddp
from unet_implementations_b200.models.losses import SimpleLoss
from unet_implementations_b200.models.unet import UNet


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    model = UNet().to(dev).train()
    ddp.broadcast_parameters(model)
    g = torch.Generator().manual_seed(100 + rank)  # every rank has its own shard
    x = torch.randn(4, 3, 128, 128, generator=g).to(dev)
    t = torch.randint(0, 3, (4, 128, 128), generator=g).to(dev)
    loss_fn = SimpleLoss()

    def step():
        for p in model.parameters():
            p.grad = None
        torch.manual_seed(7 + rank)  # same dropout masks in both runs
        loss_fn(model(x), t).backward()
        return [p.grad.clone() for p in model.parameters()]

    local = step()  # no reducer: this rank's own gradient
    reducer = ddp.BucketedGradAllReduce(model, bucket_bytes=4 << 20)
    reduced = step()
    reduced2 = step()  # a second step through the same reducer (bucket counters reset)
    worst, worst_x = 0.0, 0.0
    for lg, rg, rg2, (name, p) in zip(local, reduced, reduced2, model.named_parameters()):
        gathered = [torch.empty_like(lg) for _ in range(world)]
        dist.all_gather(gathered, lg)
        mean = torch.stack(gathered).double().mean(0)
        den = mean.norm().item()
        if den > 0:
            worst = max(worst, (rg.double() - mean).norm().item() / den)
        # every rank holds the same bits, and the reducer is repeatable
        same = [torch.empty_like(rg) for _ in range(world)]
        dist.all_gather(same, rg)
        worst_x = max(worst_x, max((s - same[0]).abs().max().item() for s in same), (rg2 - rg).abs().max().item())
    if rank == 0:
        print(f"ddp_check: world {world}, {len(reducer.buckets)} buckets; all-reduced gradient vs mean of the ranks' local "
              f"gradients: worst rel-L2 {worst:.2e}; max difference between ranks / between two steps: {worst_x:.1e}")
    assert worst <= 1e-6 and worst_x == 0.0
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
