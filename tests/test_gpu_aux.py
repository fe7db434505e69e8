"""GPU parity of the two steps either side of the hot path: the fused SGD-Nesterov optimizer step
against torch.optim.SGD (bit-exact) and the fused argmax + Dice counters against the reference's validate() arithmetic
(Our_UNet/src/train.py:554-572; integer counts exact, argmax bit-exact)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(32, 3, 3, 3), (32,), (64, 32, 3, 3), (3, 32, 1, 1), (3,), (512, 256, 3, 3), (1,), (1025,)]
    return [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]


@pytest.mark.parametrize("momentum,nesterov,wd", [(0.99, True, 1e-4), (0.9, False, 0.0), (0.0, False, 1e-4)])
def test_fused_sgd_bit_exact_with_torch(momentum, nesterov, wd):
    from unet_implementations_b200.optim import FusedSGD
    a, b = _params(1), _params(1)
    oa = torch.optim.SGD(a, lr=0.01, momentum=momentum, nesterov=nesterov, weight_decay=wd)
    ob = FusedSGD(b, lr=0.01, momentum=momentum, nesterov=nesterov, weight_decay=wd)
    sched_a = torch.optim.lr_scheduler.LambdaLR(oa, lambda e: (1 - e / 10) ** 0.9)  # train.py:466-475
    sched_b = torch.optim.lr_scheduler.LambdaLR(ob, lambda e: (1 - e / 10) ** 0.9)
    g = torch.Generator().manual_seed(2)
    for step in range(4):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).cuda()
            pa.grad = gr.clone()
            pb.grad = gr.clone() if not (step == 1 and pa.numel() == 1025) else None  # a parameter that skips a step
            if pb.grad is None:
                pa.grad = None
        oa.step()
        ob.step()
        sched_a.step()
        sched_b.step()
        for pa, pb in zip(a, b):
            assert torch.equal(pa, pb), (step, tuple(pa.shape), (pa - pb).abs().max().item())
    if momentum:
        sa, sb = oa.state_dict(), ob.state_dict()
        for k in sa["state"]:
            assert torch.equal(sa["state"][k]["momentum_buffer"], sb["state"][k]["momentum_buffer"])


def test_fused_sgd_trains_the_model():
    from unet_implementations_b200.models.losses import SimpleLoss
    from unet_implementations_b200.models.unet import UNet
    from unet_implementations_b200.optim import FusedSGD
    torch.manual_seed(5)
    model = UNet(n_stages=3, features_per_stage=[32, 64, 64], encoder_dropout_rates=[0, 0.1, 0.2],
                 decoder_dropout_rates=[0.2, 0]).cuda().train()
    x = torch.randn(2, 3, 32, 32, device="cuda")
    t = torch.randint(0, 3, (2, 32, 32), device="cuda")
    opt = FusedSGD(model.parameters(), lr=0.005, momentum=0.99, nesterov=True, weight_decay=1e-4)
    loss_fn = SimpleLoss()
    losses = []
    for _ in range(10):
        opt.zero_grad()
        loss = loss_fn(model(x), t)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_argmax_counts_match_reference_validate():
    from unet_implementations_b200 import metrics
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(3, 3, 37, 53, generator=g)
    logits[0, :, :4, :4] = 0.5          # ties -> lowest index
    logits[1, 1, 5, 5] = logits[1, 2, 5, 5] = 9.0
    target = torch.randint(0, 3, (3, 37, 53), generator=g)
    target[torch.rand(3, 37, 53, generator=g) < 0.15] = 255
    target[2][target[2] == 2] = 1        # a class that is absent from one image
    pred, counts = metrics.argmax_counts(logits.cuda(), target.cuda())
    ref_pred = torch.argmax(logits, dim=1)
    assert torch.equal(pred.cpu(), ref_pred)
    valid = target != 255
    for c in range(3):  # train.py:557-572
        pc, mc = (ref_pred == c) & valid, (target == c) & valid
        assert counts[c, 0].item() == int((pc & mc).sum())
        assert counts[c, 1].item() == int(pc.sum())
        assert counts[c, 2].item() == int(mc.sum())
        inter, union = (pc & mc).float().sum(), pc.float().sum() + mc.float().sum()
        ref_dice = (2.0 * inter) / (union + 1e-5) if union > 0 else torch.tensor(1.0)
        assert abs(metrics.dice_from_counts(counts)[c].item() - ref_dice.item()) <= 1e-6
    # all-ignored batch: union == 0 -> 1.0 (train.py:571-572)
    _, c0 = metrics.argmax_counts(logits.cuda(), torch.full((3, 37, 53), 255, dtype=torch.int64, device="cuda"), want_pred=False)
    assert int(c0.sum()) == 0 and torch.equal(metrics.dice_from_counts(c0).cpu(), torch.ones(3))


def test_preprocess_u8_bit_exact_with_reference_dataset_arithmetic():
    """train.py:299-311: image.float().permute(2,0,1) / 255.0, (image - mean) / std, mask clean-up, .long()."""
    import numpy as np
    from unet_implementations_b200 import data
    g = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (3, 40, 56, 3), generator=g, dtype=torch.uint8)
    mask = torch.randint(0, 3, (3, 40, 56), generator=g, dtype=torch.uint8)
    mask[torch.rand(3, 40, 56, generator=g) < 0.1] = 255
    mask[torch.rand(3, 40, 56, generator=g) < 0.05] = 7   # stray label -> 0
    out, mout = data.preprocess_batch(img.cuda(), mask.cuda())
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    for b in range(3):
        ref = (torch.from_numpy(img[b].numpy()).float().permute(2, 0, 1) / 255.0 - mean) / std
        assert torch.equal(out[b].cpu(), ref)
        m = mask[b].numpy()
        refm = torch.from_numpy(np.where((m > 2) & (m != 255), 0, m)).long()
        assert torch.equal(mout[b].cpu(), refm)
    assert mout.dtype == torch.int64 and out.dtype == torch.float32
