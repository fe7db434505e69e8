"""CPU, world_size 2, gloo: the bucketed gradient all-reduce of unet-implementations_b200/ddp.py.  The gradient
sink is driven by hand in the order UNet's backward produces gradients (the CUDA kernels themselves need a GPU)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet_implementations_b200 import ddp
    from unet_implementations_b200.models.unet import UNet
    torch.manual_seed(1234 + rank)  # deliberately different: broadcast must make the replicas identical
    model = UNet(n_stages=3, features_per_stage=[32, 64, 64], encoder_dropout_rates=[0, 0.1, 0.2],
                 decoder_dropout_rates=[0.2, 0])
    ddp.broadcast_parameters(model)
    w0 = model.segmentation_output.weight.detach().clone()
    red = ddp.BucketedGradAllReduce(model, bucket_bytes=64 << 10, device=torch.device("cpu"), tail_bytes=16 << 10)
    order = ddp.backward_param_order(model)
    assert len(order) == len(list(model.parameters())) and len({id(p) for p in order}) == len(order)
    assert len(red.buckets) >= 3
    # the last bucket (its all-reduce cannot overlap anything) holds only the small gradients produced last
    assert (red.buckets[-1][1] - red.buckets[-1][0]) * 4 <= 16 << 10 and red.buckets[-1][1] == red.numel
    assert all(a[1] == b[0] for a, b in zip(red.buckets, red.buckets[1:])) and red.buckets[0][0] == 0
    views = {}
    g = torch.Generator().manual_seed(7)  # same base on both ranks; rank enters as a known offset
    expect = {}
    for p in order:
        base = torch.randn(p.shape, generator=g)
        views[id(p)] = red(p, base + rank)           # rank r contributes base + r
        expect[id(p)] = base + (world - 1) / 2.0      # mean over ranks
    red.finish()
    ok = all(torch.allclose(views[id(p)], expect[id(p)], atol=1e-6) for p in order)
    ok &= all(views[id(p)].shape == p.shape and views[id(p)].data_ptr() % 16 == 0 for p in order)
    # second step through the same reducer, with every conv weight delivered ONE LAYER LATE (what the side-stream weight
    # gradients of models/unet.py do): a bucket must wait for all of its gradients, in whatever order they arrive
    late, held = [], None
    for p in order:
        if p.dim() == 4 and p.shape[-1] == 3:   # a 3x3 conv weight: hold it back until the next one shows up
            if held is not None:
                late.append(held)
            held = p
        else:
            late.append(p)
    late.append(held)
    assert len(late) == len(order) and [id(p) for p in late] != [id(p) for p in order]
    g2 = torch.Generator().manual_seed(8)
    vals = {id(p): torch.randn(p.shape, generator=g2) for p in order}
    for p in late:
        views[id(p)] = red(p, vals[id(p)] * (rank + 1))     # mean over ranks = vals * (1 + 2) / 2
    red.finish()
    ok &= all(torch.allclose(views[id(p)], vals[id(p)] * (world + 1) / 2.0, atol=1e-5) for p in order)
    # third step: one parameter gets NO gradient (a frozen layer, an unused fusion conv): its bucket never fills up, but
    # finish() must still reduce the gradients that did arrive in it
    skipped = order[len(order) // 2]
    g3 = torch.Generator().manual_seed(9)
    vals3 = {id(p): torch.randn(p.shape, generator=g3) for p in order}
    for p in order:
        if p is not skipped:
            views[id(p)] = red(p, vals3[id(p)] * (rank + 1))
    red.finish()
    ok &= all(torch.allclose(views[id(p)], vals3[id(p)] * (world + 1) / 2.0, atol=1e-5) for p in order if p is not skipped)
    # a gradient that autograd adopted as p.grad and the caller kept must not be overwritten in place
    w = order[0]
    w.grad = red.dest(w)
    try:
        red.dest(w)
        ok = False
    except RuntimeError:
        pass
    w.grad = None
    # bytes, not a tensor: a tensor travels as a shared-memory handle that dies with this process if the parent is slow
    q.put((rank, bool(ok), w0.numpy().tobytes()))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]  # identical replicas after the broadcast
