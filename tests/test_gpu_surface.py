"""GPU: behaviour of the module surface at its edges -- input layouts and dtypes the trainer can hand over, sizes the
kernels do not cover (must raise, never fall back), repeated backward, inference with frozen parameters."""
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _small():
    from unet_implementations_b200.models.unet import UNet
    torch.manual_seed(3)
    return UNet(n_stages=3, features_per_stage=[32, 64, 64], encoder_dropout_rates=[0, 0, 0],
                decoder_dropout_rates=[0, 0]).cuda().eval()


def test_input_layouts_and_dtypes_give_the_same_logits():
    model = _small()
    x = torch.randn(2, 3, 64, 96, device="cuda")
    with torch.no_grad():
        ref = model(x)
        assert ref.shape == (2, 3, 64, 96) and ref.dtype == torch.float32
        assert torch.equal(model(x.to(memory_format=torch.channels_last)), ref)           # channels-last strides
        assert torch.equal(model(x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)), ref)  # non-contiguous view
        assert torch.equal(model(x.double()), ref)                                        # other float dtypes are cast to fp32
        h = model(x.half())
        assert O.rel_l2(h, ref) < 5e-2                                                    # fp16 input: its rounding (5e-4), amplified by the net
        one = model(x[:1])
        assert torch.equal(one, ref[:1])                                                  # batch 1; samples are independent


def test_unsupported_sizes_and_devices_raise():
    model = _small()
    with pytest.raises(NotImplementedError):
        model(torch.randn(1, 3, 30, 32, device="cuda"))      # 30 does not halve twice exactly (F.interpolate to skip size)
    with pytest.raises(ValueError):
        model(torch.randn(1, 4, 32, 32, device="cuda"))      # wrong channel count
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.randn(1, 3, 32, 32))


def test_backward_twice_raises_and_frozen_model_runs_without_saving():
    from unet_implementations_b200.models.losses import SimpleLoss
    model = _small().train()
    x = torch.randn(1, 3, 32, 32, device="cuda")
    t = torch.randint(0, 3, (1, 32, 32), device="cuda")
    loss = SimpleLoss()(model(x), t)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        loss.backward()                                       # the fused node frees its arena after the first backward
    for p in model.parameters():
        p.requires_grad = False
    out = model(x)                                            # grad mode on, nothing trainable: plain inference
    assert not out.requires_grad


def test_simple_loss_accepts_bf16_and_fp16_logits():
    from unet_implementations_b200.models.losses import SimpleLoss
    g = torch.Generator().manual_seed(1)
    lg = torch.randn(2, 3, 16, 24, generator=g).cuda().requires_grad_(True)
    t = torch.randint(0, 3, (2, 16, 24), generator=g).cuda()
    ref = SimpleLoss()(lg, t)
    for dt in (torch.bfloat16, torch.float16):
        l2 = lg.detach().to(dt).requires_grad_(True)
        v = SimpleLoss()(l2, t)
        v.backward()
        assert v.dtype == torch.float32 and abs(v.item() - ref.item()) < 2e-2 and l2.grad.dtype == dt
