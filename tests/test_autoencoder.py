"""Autoencoder variant (BASELINE.json configs[3]; AE_pretrained/reconstruction): CPU tests of the module surface and of
the oracle against the unmodified reference's outputs (tests/golden/small_ae.pt), GPU parity of the fused step in
fp32 mode (1e-4) and in bf16 mode (loss 1e-2; output/gradients bounded by the oracle's own bf16-autocast deviation)."""
import hashlib

import pytest
import torch

from conftest import load_golden
from oracle import unet_oracle as O


def _sha(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def test_autoencoder_surface_matches_reference():
    from unet_implementations_b200.models.autoencoder import Autoencoder
    g = load_golden("small_ae.pt")
    torch.manual_seed(1234)
    m = Autoencoder()
    assert list(m.state_dict().keys()) == g["default_keys"]
    assert _sha(m.state_dict()) == g["default_sha256"]          # same seed => the reference's weights, bit for bit
    assert isinstance(m.reconstruction_output[0], torch.nn.Conv2d) and isinstance(m.reconstruction_output[1], torch.nn.Sigmoid)
    assert m.get_encoder() is m.encoder_stages and m.get_decoder()[1] is m.reconstruction_output
    assert m.out_channels == 3 and m.in_channels == 3
    small = Autoencoder(**g["cfg"])
    small.load_state_dict(g["state_dict"])                       # reference checkpoints load


def test_oracle_autoencoder_matches_reference_fixture():
    from unet_implementations_b200.models.autoencoder import Autoencoder
    g = load_golden("small_ae.pt")
    cfg = O.config_of(Autoencoder(**g["cfg"]))
    torch.manual_seed(g["dropout_seed"])
    masks = O.draw_dropout_masks(cfg, g["x"].shape[0], g["x"])
    r = O.autoencoder_training_step(g["state_dict"], g["x"], g["target"], cfg, masks)
    assert O.rel_l2(r["output"], g["output_train"]) <= 1e-5
    assert abs(r["loss"].item() - g["loss"].item()) <= 1e-6 * abs(g["loss"].item())
    for k, v in g["grads"].items():
        if v.abs().max() > 1e-6:
            assert O.rel_l2(r["grads"][k], v) <= 1e-4, k
    ev = O.unet_forward(g["state_dict"], g["x"], cfg, None, training=False)
    assert O.rel_l2(ev, g["output_eval"]) <= 1e-5


def _build(g, precision):
    from unet_implementations_b200.models.autoencoder import Autoencoder
    m = Autoencoder(**g["cfg"])
    m.load_state_dict(g["state_dict"])
    m.precision = precision
    return m.cuda().train()


def _dead_bias(name, model):
    if not name.endswith(".bias") or name.startswith("reconstruction_output"):
        return False
    return isinstance(model.get_submodule(name.rsplit(".", 1)[0]), torch.nn.Conv2d)


@pytest.mark.gpu
def test_autoencoder_fp32_mode_against_reference_golden():
    from unet_implementations_b200.models.losses import MSELoss
    g = load_golden("small_ae.pt")
    model = _build(g, "fp32")
    cfg = O.config_of(model)
    torch.manual_seed(g["dropout_seed"])
    model._mask_override = O.draw_dropout_masks(cfg, g["x"].shape[0], g["x"])
    out = model(g["x"].cuda())
    loss = MSELoss()(out, g["target"].cuda())
    loss.backward()
    assert O.rel_l2(out, g["output_train"]) <= 1e-4
    assert abs(loss.item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    worst = 0.0
    for k, p in model.named_parameters():
        if _dead_bias(k, model):
            assert p.grad.abs().max().item() == 0.0, k
            continue
        worst = max(worst, O.rel_l2(p.grad, g["grads"][k]))
    assert worst <= 1e-4, worst
    model.eval()
    with torch.no_grad():
        assert O.rel_l2(model(g["x"].cuda()), g["output_eval"]) <= 1e-4
    # torch's own nn.MSELoss on the module output gives the same loss and drives the same backward
    model.train()
    model.zero_grad(set_to_none=True)
    torch.manual_seed(g["dropout_seed"])
    model._mask_override = O.draw_dropout_masks(cfg, g["x"].shape[0], g["x"])
    loss2 = torch.nn.MSELoss()(model(g["x"].cuda()), g["target"].cuda())
    assert abs(loss2.item() - loss.item()) <= 1e-6


@pytest.mark.gpu
def test_autoencoder_bf16_mode_against_oracle_yardstick():
    from unet_implementations_b200.models.losses import MSELoss
    g = load_golden("small_ae.pt")
    model = _build(g, "bf16")
    cfg = O.config_of(model)
    torch.manual_seed(g["dropout_seed"])
    masks = O.draw_dropout_masks(cfg, g["x"].shape[0], g["x"])
    model._mask_override = masks
    out = model(g["x"].cuda())
    loss = MSELoss()(out, g["target"].cuda())
    loss.backward()
    with torch.autocast("cpu", dtype=torch.bfloat16):  # what bf16 costs the reference's own ops
        ref16 = O.autoencoder_training_step(g["state_dict"], g["x"], g["target"], cfg, masks)
    y_out = O.rel_l2(ref16["output"].float(), g["output_train"])
    assert abs(loss.item() - g["loss"].item()) <= 1e-2 * abs(g["loss"].item())
    assert O.rel_l2(out, g["output_train"]) <= 1.25 * y_out + 1e-3
    ours, yard = 0.0, 0.0
    for k, p in model.named_parameters():
        if _dead_bias(k, model):
            continue
        ours += O.rel_l2(p.grad, g["grads"][k])
        yard += O.rel_l2(ref16["grads"][k].float(), g["grads"][k])
    assert ours <= 1.15 * yard, (ours, yard)


@pytest.mark.gpu
def test_autoencoder_default_model_fp32_against_fp64_oracle():
    """The trainer's full 6-stage Autoencoder (train.py:351-370 dropout rates) at 128x128 in fp32 mode: output, loss and
    every gradient within 1e-4 of the fp64 oracle."""
    from unet_implementations_b200.models.autoencoder import Autoencoder
    from unet_implementations_b200.models.losses import MSELoss
    torch.manual_seed(1234)
    model = Autoencoder(encoder_dropout_rates=[0.0, 0.0, 0.05, 0.1, 0.15, 0.15], decoder_dropout_rates=[0.15, 0.1, 0.1, 0.05, 0.0])
    cfg = O.config_of(model)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.precision = "fp32"
    model = model.cuda().train()
    gen = torch.Generator().manual_seed(0)
    x, target = torch.rand(1, 3, 128, 128, generator=gen), torch.rand(1, 3, 128, 128, generator=gen)
    torch.manual_seed(99)
    masks = O.draw_dropout_masks(cfg, 1, x)
    model._mask_override = masks
    out = model(x.cuda())
    loss = MSELoss()(out, target.cuda())
    loss.backward()
    ref = O.autoencoder_training_step(sd, x, target, cfg, masks, dtype=torch.float64)
    assert O.rel_l2(out.double().cpu(), ref["output"]) <= 1e-4
    assert abs(loss.item() - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
    worst = max(O.rel_l2(p.grad.double().cpu(), ref["grads"][k]) for k, p in model.named_parameters() if not _dead_bias(k, model))
    assert worst <= 1e-4, worst
