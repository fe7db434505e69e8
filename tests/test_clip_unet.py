"""CLIP-conditioned UNet (BASELINE.json configs[4]; CLIP_UNet/models/unet.py): surface and oracle on CPU against the
unmodified reference's outputs (tests/golden/small_clip_unet.pt), and GPU parity of the fused step with the fusion
layer -- fp32 mode within 1e-4 of the reference; bf16 mode bounded by the oracle's bf16-autocast deviation."""
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


def test_clip_unet_surface_matches_reference():
    from unet_implementations_b200.models.clip_unet import UNet
    g = load_golden("small_clip_unet.pt")
    torch.manual_seed(1234)
    m = UNet()
    assert list(m.state_dict().keys()) == g["default_keys"] and len(g["default_keys"]) == 94
    assert _sha(m.state_dict()) == g["default_sha256"]
    assert m.with_clip_features and m.clip_dim == 512 and m.clip_fusion_conv[0].in_channels == 1024
    UNet(**g["cfg"]).load_state_dict(g["state_dict"])
    assert not hasattr(UNet(with_clip_features=False), "clip_fusion_conv")


def test_oracle_clip_unet_matches_reference_fixture():
    from unet_implementations_b200.models.clip_unet import UNet
    g = load_golden("small_clip_unet.pt")
    cfg = O.config_of(UNet(**g["cfg"]))
    torch.manual_seed(g["dropout_seed"])
    masks = O.draw_dropout_masks(cfg, 2, g["x"])
    r = O.training_step(g["state_dict"], g["x"], g["target"], cfg, masks, clip_features=g["clip"])
    assert O.rel_l2(r["logits"], g["logits_train"]) <= 1e-5
    assert abs(r["loss"].item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    for k, v in g["grads"].items():
        if v.abs().max() > 1e-6:
            assert O.rel_l2(r["grads"][k], v) <= 2e-4, k
    for key, cf in (("logits_eval", g["clip"]), ("logits_eval_resized", g["clip_other"]), ("logits_eval_noclip", None)):
        ev = O.unet_forward(g["state_dict"], g["x"], cfg, None, training=False, clip_features=cf)
        assert O.rel_l2(ev, g[key]) <= 1e-5, key


def _build(g, precision):
    from unet_implementations_b200.models.clip_unet import UNet
    m = UNet(**g["cfg"])
    m.load_state_dict(g["state_dict"])
    m.precision = precision
    return m.cuda()


def _dead_bias(name, model):
    if not name.endswith(".bias") or name.startswith("segmentation_output"):
        return False
    return isinstance(model.get_submodule(name.rsplit(".", 1)[0]), torch.nn.Conv2d)


@pytest.mark.gpu
def test_clip_unet_fp32_mode_against_reference_golden():
    from unet_implementations_b200.models.losses import SimpleLoss
    g = load_golden("small_clip_unet.pt")
    model = _build(g, "fp32").train()
    cfg = O.config_of(model)
    torch.manual_seed(g["dropout_seed"])
    model._mask_override = O.draw_dropout_masks(cfg, 2, g["x"])
    logits = model(g["x"].cuda(), g["clip"].cuda())
    loss = SimpleLoss()(logits, g["target"].cuda())
    loss.backward()
    assert O.rel_l2(logits, g["logits_train"]) <= 1e-4
    assert abs(loss.item() - g["loss"].item()) <= 1e-4 * abs(g["loss"].item())
    assert torch.equal(logits.argmax(1).cpu(), g["logits_train"].argmax(1))
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        if _dead_bias(k, model):
            assert p.grad.abs().max().item() == 0.0, k
            continue
        e = O.rel_l2(p.grad, g["grads"][k])
        worst = max(worst, (k, e), key=lambda t: t[1])
    assert worst[1] <= 1e-4, worst
    model.eval()
    with torch.no_grad():
        assert O.rel_l2(model(g["x"].cuda(), g["clip"].cuda()), g["logits_eval"]) <= 1e-4
        assert O.rel_l2(model(g["x"].cuda(), g["clip_other"].cuda()), g["logits_eval_resized"]) <= 1e-4  # resized patch grid
        assert O.rel_l2(model(g["x"].cuda()), g["logits_eval_noclip"]) <= 1e-4                          # no features: plain UNet


@pytest.mark.gpu
def test_clip_unet_bf16_mode_against_oracle_yardstick():
    from unet_implementations_b200.models.losses import SimpleLoss
    g = load_golden("small_clip_unet.pt")
    model = _build(g, "bf16").train()
    cfg = O.config_of(model)
    torch.manual_seed(g["dropout_seed"])
    masks = O.draw_dropout_masks(cfg, 2, g["x"])
    model._mask_override = masks
    logits = model(g["x"].cuda(), g["clip"].cuda())
    loss = SimpleLoss()(logits, g["target"].cuda())
    loss.backward()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ref16 = O.training_step(g["state_dict"], g["x"], g["target"], cfg, masks, clip_features=g["clip"])
    assert abs(loss.item() - g["loss"].item()) <= 1e-2 * abs(g["loss"].item())
    assert O.rel_l2(logits, g["logits_train"]) <= 1.25 * O.rel_l2(ref16["logits"].float(), g["logits_train"]) + 1e-3
    ours = yard = 0.0
    for k, p in model.named_parameters():
        if not _dead_bias(k, model):
            ours += O.rel_l2(p.grad, g["grads"][k])
            yard += O.rel_l2(ref16["grads"][k].float(), g["grads"][k])
    assert ours <= 1.15 * yard, (ours, yard)


@pytest.mark.gpu
def test_clip_unet_default_model_with_patch_grid():
    """The trainer's configuration (512-d patch features on the 16x16 grid of a 512x512 image) at batch 1, bf16: the
    fusion layer runs on the tensor-core kernels (1024 -> 512 channels) and the step trains."""
    from unet_implementations_b200.models.clip_unet import UNet
    from unet_implementations_b200.models.losses import SimpleLoss
    torch.manual_seed(1234)
    model = UNet().cuda().train()
    x = torch.randn(1, 3, 512, 512, device="cuda")
    clip = torch.randn(1, 512, 16, 16, device="cuda")
    t = torch.randint(0, 3, (1, 512, 512), device="cuda")
    loss = SimpleLoss()(model(x, clip), t)
    loss.backward()
    assert torch.isfinite(loss)
    gw = model.clip_fusion_conv[0].weight.grad
    assert gw.shape == (512, 1024, 1, 1) and torch.isfinite(gw).all() and gw.abs().max() > 0
    # the clip half of the fusion weight matters: different features, different logits
    with torch.no_grad():
        a, b = model.eval()(x, clip), model(x, clip * 0.5)
    assert not torch.equal(a, b)


@pytest.mark.gpu
def test_pooled_embedding_broadcast_over_the_grid_cancels_in_the_instance_norm():
    """What the reference's ClipPatchExtractor actually feeds the model is ONE pooled embedding per image expanded over
    the 16x16 grid (CLIP_UNet/models/unet.py:611-612).  A plane-constant input of the 1x1 fusion conv adds a plane-
    constant to its output, which the following InstanceNorm removes: the fused path skips that half of the conv.
    Against the oracle on the SAME expanded features (fp32 mode, 1e-4): identical outputs, and a clip-half weight
    gradient that is zero up to the reference's own rounding noise."""
    from unet_implementations_b200.models.losses import SimpleLoss
    g = load_golden("small_clip_unet.pt")
    model = _build(g, "fp32").train()
    cfg = O.config_of(model)
    pooled = g["clip"].mean((2, 3), keepdim=True)
    expanded = pooled.expand(-1, -1, g["clip"].shape[2], g["clip"].shape[3])
    assert expanded.stride(2) == 0 and expanded.stride(3) == 0
    torch.manual_seed(g["dropout_seed"])
    masks = O.draw_dropout_masks(cfg, 2, g["x"])
    ref = O.training_step(g["state_dict"], g["x"], g["target"], cfg, masks, clip_features=expanded.contiguous())
    res = {}
    for mode in (True, False):  # the shortcut, and the full concat path on the same values
        model.skip_constant_features = mode
        model.zero_grad(set_to_none=True)
        model._mask_override = [m.clone() for m in masks]
        feats = pooled.cuda().expand(-1, -1, g["clip"].shape[2], g["clip"].shape[3])  # expanded ON the device (a .cuda() of
        assert feats.stride(2) == 0                                                    # an expanded tensor materialises it)
        logits = model(g["x"].cuda(), feats if mode else feats.contiguous())
        loss = SimpleLoss()(logits, g["target"].cuda())
        loss.backward()
        res[mode] = (logits.detach().cpu(), loss.item(), {k: p.grad.detach().cpu() for k, p in model.named_parameters()})
        assert O.rel_l2(res[mode][0], ref["logits"]) <= 1e-4
        assert abs(res[mode][1] - ref["loss"].item()) <= 1e-4 * abs(ref["loss"].item())
    gw = res[True][2]["clip_fusion_conv.0.weight"]
    enc_c = gw.shape[1] - g["clip"].shape[1]
    assert float(gw[:, enc_c:].abs().max()) == 0.0
    ref_gw = ref["grads"]["clip_fusion_conv.0.weight"]
    assert float(ref_gw[:, enc_c:].abs().max()) <= 1e-4 * float(ref_gw[:, :enc_c].abs().max())  # the reference's is noise
    assert O.rel_l2(gw[:, :enc_c], ref_gw[:, :enc_c]) <= 1e-4
    for k, v in res[True][2].items():
        if k != "clip_fusion_conv.0.weight" and not _dead_bias(k, model):
            assert O.rel_l2(v, ref["grads"][k]) <= 1e-4, k
