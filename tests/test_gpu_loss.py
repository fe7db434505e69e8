"""GPU parity of SimpleLoss (fp32 path): the fused reduction / elementwise-gradient kernels against the reference's
committed outputs (tests/golden/loss_cases.pt) and the float64 numpy oracle.  fp32 mode tolerance: 1e-4 relative."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def test_loss_cases_match_reference():
    from unet_implementations_b200.models.losses import SimpleLoss
    cases = load_golden("loss_cases.pt")
    for name, c in cases.items():
        kw = dict(c["kwargs"])
        if kw["class_weights"] is not None:
            kw["class_weights"] = kw["class_weights"].cuda()
        fn = SimpleLoss(**kw)
        lg = c["logits"].cuda().requires_grad_(True)
        val = fn(lg, c["target"].cuda())
        val.backward()
        assert val.dtype == torch.float32 and val.dim() == 0
        assert abs(val.item() - c["loss"].item()) <= 1e-5 * abs(c["loss"].item()), name
        assert O.rel_l2(lg.grad, c["dlogits"]) <= 1e-5, name
        assert (lg.grad.cpu() - c["dlogits"]).abs().max().item() <= 1e-8 + 1e-4 * c["dlogits"].abs().max().item(), name


@pytest.mark.parametrize("variant", ["uniform", "pets", "cats"])
def test_loss_full_size_against_numpy_oracle(variant):
    """512x512 (the BASELINE size), including the absent-class clamp (losses.py:53-54) for 'cats'."""
    from unet_implementations_b200.models.losses import SimpleLoss
    _, target = O.synthetic_batch(2, 512, seed=4, variant=variant)
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(2, 3, 512, 512, generator=g) * 3
    tot, ce, dice, dz = O.simple_loss_numpy(logits.numpy(), target.numpy())
    lg = logits.cuda().requires_grad_(True)
    val = SimpleLoss()(lg, target.cuda())
    (val * 3.0).backward()  # upstream gradient is honoured
    assert abs(val.item() - tot) <= 1e-5 * abs(tot)
    assert O.rel_l2(lg.grad, torch.from_numpy(dz * 3.0)) <= 1e-5


def test_loss_counts_are_exact_integers_at_full_size():
    """Class counts are integer work: the dynamic weights must equal the oracle's bit for bit in fp32."""
    from unet_implementations_b200 import ops
    _, target = O.synthetic_batch(4, 512, seed=11)
    w_ref = O.class_weights(target)
    logits = torch.zeros(4, 3, 512, 512, device="cuda")
    out, tables = ops.loss_forward(logits, target.cuda(), None, True, 1.0, 1.0, 255, 1e-5)
    w_got = tables[:3].cpu()  # w_c / sum_valid w_t: normalise back to sum 3 (losses.py:60)
    w_got = w_got * (3.0 / w_got.sum())
    assert torch.allclose(w_got, w_ref, rtol=2e-6, atol=0)
    # with all-zero logits CE = sum_c w_c n_c log 3 / sum_c w_c n_c = log 3 exactly up to rounding
    assert abs(out[1].item() - float(np.log(3.0))) < 1e-5


def test_loss_accepts_half_logits_and_all_ignored_image():
    from unet_implementations_b200.models.losses import SimpleLoss
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(2, 3, 16, 24, generator=g)
    target = torch.randint(0, 3, (2, 16, 24), generator=g)
    target[1] = 255
    ref = O.simple_loss(logits.bfloat16().float(), target)
    lg = logits.cuda().bfloat16().requires_grad_(True)
    val = SimpleLoss()(lg, target.cuda())
    val.backward()
    assert val.dtype == torch.float32 and lg.grad.dtype == torch.bfloat16
    assert abs(val.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert float(lg.grad[1].abs().max()) == 0.0  # ignored pixels receive no gradient
