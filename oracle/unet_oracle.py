"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's hot path (Ulixes-8/UNet-Implementations, Our_UNet): the 6-stage UNet forward
and SimpleLoss, written as plain functional torch ops on fp32 CPU tensors plus an independent float64 numpy
restatement of the loss.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may
import this module, and only as the checker or the reported CPU baseline -- never as something the product path
calls (the product raises if its CUDA library is missing).

Parity pin: the reference ships no tests or golden vectors, so this restatement is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by tests/golden/make_golden.py (which
imports /root/reference/Our_UNet) and committed under tests/golden/.  tests/test_oracle.py checks the oracle
against those fixtures on CPU.

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """Constructor arguments of models.unet.UNet that shape the computation (Our_UNet/models/unet.py:238-311;
    the trainer's values at Our_UNet/src/train.py:776-795 are the defaults)."""
    in_channels: int = 3
    num_classes: int = 3
    features: Sequence[int] = (32, 64, 128, 256, 512, 512)
    strides: Sequence[int] = (1, 2, 2, 2, 2, 2)
    encoder_dropout: Sequence[float] = (0.0, 0.0, 0.1, 0.2, 0.3, 0.3)
    decoder_dropout: Sequence[float] = (0.3, 0.2, 0.2, 0.1, 0.0)
    eps: float = 1e-5
    negative_slope: float = 0.01
    n_convs: int = 2

    @property
    def n_stages(self):
        return len(self.features)


def block_keys(prefix: str, rate: float, n_convs: int = 2):
    """state_dict key stems of one ConvBlock: the nn.Sequential index of conv i and norm i depends on whether a
    SpatialDropout2d module sits in the block (unet.py:101-134)."""
    per = 4 if rate > 0 else 3  # conv, norm, lrelu, [dropout]
    return [(f"{prefix}.block.{i * per}", f"{prefix}.block.{i * per + 1}") for i in range(n_convs)]


def dropout_sites(cfg: UNetConfig):
    """(channels, rate) of every SpatialDropout2d call in forward order: encoder stages then
    decoder stages, two per block, only where rate > 0 (unet.py:126-127)."""
    sites = []
    for s in range(cfg.n_stages):
        if cfg.encoder_dropout[s] > 0:
            sites += [(cfg.features[s], cfg.encoder_dropout[s])] * cfg.n_convs
    for j in range(cfg.n_stages - 1):
        d = cfg.n_stages - 2 - j
        if cfg.decoder_dropout[j] > 0:
            sites += [(cfg.features[d], cfg.decoder_dropout[j])] * cfg.n_convs
    return sites


def draw_dropout_masks(cfg: UNetConfig, batch: int, like: torch.Tensor) -> List[torch.Tensor]:
    """The reference's mask draw, verbatim in call and order (SpatialDropout2d.forward, unet.py:30-31):
    `x.new_empty(B, C, 1, 1).bernoulli_(1 - p).div_(1 - p)` for each site, on `like`'s device/dtype generator."""
    return [like.new_empty(batch, c, 1, 1).bernoulli_(1 - p).div_(1 - p) for (c, p) in dropout_sites(cfg)]


class _RoundBF16(torch.autograd.Function):
    """Storage emulation for the `bf16_storage` mode: the value is rounded to bf16 on the way forward and its
    gradient is rounded to bf16 on the way back -- the two places where the CUDA path keeps a tensor in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def rb(x: torch.Tensor, on: bool) -> torch.Tensor:
    return _RoundBF16.apply(x) if on else x


TRACE: Optional[list] = None  # debug: set to a list to collect (raw conv output, activation) per conv unit


def conv_block(x, sd: Dict[str, torch.Tensor], prefix: str, stride: int, rate: float, cfg: UNetConfig, masks, training,
               bf16_storage: bool = False, fp32_first_weight: bool = False):
    """ConvBlock.forward (unet.py:97-141): [Conv2d 3x3 pad 1 (stride on the first conv only, :103) ->
    InstanceNorm2d(eps, affine) (:118-119) -> LeakyReLU (:122-123) -> SpatialDropout2d (:126-127)] x n_convs.
    `bf16_storage` mirrors the CUDA path's precision policy on top of the same arithmetic: conv operands (input
    activation, weights) and the raw conv output are bf16-rounded, everything else stays fp32."""
    for i, (ck, nk) in enumerate(block_keys(prefix, rate, cfg.n_convs)):
        w = sd[ck + ".weight"]
        if bf16_storage:
            if not (fp32_first_weight and i == 0):  # the Cin=3 stem reads the fp32 image with fp32 weights
                w = w + (w.detach().bfloat16().float() - w.detach())  # bf16 value, straight-through gradient
                x = rb(x, True)
        bias = sd.get(ck + ".bias")
        if bf16_storage:
            # the CUDA path stores the bias-free conv output: the bias feeds an InstanceNorm and cancels exactly
            # (SURVEY.md 8a), so it is added after the rounding, where it changes nothing but its own (zero) gradient
            x = rb(F.conv2d(x, w, None, stride=stride if i == 0 else 1, padding=1), True)
            if bias is not None:
                x = x + bias.view(1, -1, 1, 1)
        else:
            x = F.conv2d(x, w, bias, stride=stride if i == 0 else 1, padding=1)
        y_raw = x
        x = F.instance_norm(x, weight=sd[nk + ".weight"], bias=sd[nk + ".bias"], eps=cfg.eps)
        x = F.leaky_relu(x, cfg.negative_slope)
        if rate > 0 and training:
            x = x * masks.pop(0).expand_as(x)  # unet.py:34
        if TRACE is not None:
            TRACE.append((y_raw.detach(), x.detach()))
    return x


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: UNetConfig = UNetConfig(),
                 masks: Optional[List[torch.Tensor]] = None, training: bool = True,
                 bf16_storage: bool = False, clip_features: Optional[torch.Tensor] = None) -> torch.Tensor:
    """UNet.forward (unet.py:399-432): encoder with 5 skips, bottleneck, 5 UpBlocks (bilinear to the skip's size,
    cat([x, skip], 1), ConvBlock -- unet.py:203-231), 1x1 head (:430).  `masks` = draw_dropout_masks(...) when
    training (consumed in order); eval mode ignores dropout (unet.py:23-24)."""
    masks = list(masks) if (masks is not None and training) else []
    if training and not masks and dropout_sites(cfg):
        raise ValueError("training forward needs the dropout masks (draw_dropout_masks)")
    q = bf16_storage
    skips = []
    for s in range(cfg.n_stages - 1):
        x = conv_block(x, sd, f"encoder_stages.{s}", cfg.strides[s], cfg.encoder_dropout[s], cfg, masks, training, q,
                       fp32_first_weight=(s == 0))
        skips.append(rb(x, q))  # the activated tensor is stored once (bf16); each consumer's gradient is stored separately
    s = cfg.n_stages - 1
    x = conv_block(x, sd, f"encoder_stages.{s}", cfg.strides[s], cfg.encoder_dropout[s], cfg, masks, training, q)
    if clip_features is not None and "clip_fusion_conv.0.weight" in sd:
        # CLIP_UNet/models/unet.py:441-478: resize the patch features to the bottleneck, cat([x, clip], 1), then
        # clip_fusion_conv = Conv2d(1x1, bias) -> InstanceNorm2d -> LeakyReLU (unet.py:356-364)
        cf = clip_features.to(x.dtype)
        if cf.shape[2:] != x.shape[2:]:
            cf = F.interpolate(cf, size=x.shape[2:], mode="bilinear", align_corners=False)
        w = sd["clip_fusion_conv.0.weight"]
        if q:
            w = w + (w.detach().bfloat16().float() - w.detach())
        x = torch.cat([rb(x, q), rb(cf, q)], dim=1)
        x = rb(F.conv2d(x, w, None), q)
        if sd.get("clip_fusion_conv.0.bias") is not None:
            x = x + sd["clip_fusion_conv.0.bias"].view(1, -1, 1, 1)
        x = F.instance_norm(x, weight=sd["clip_fusion_conv.1.weight"], bias=sd["clip_fusion_conv.1.bias"], eps=cfg.eps)
        x = F.leaky_relu(x, cfg.negative_slope)
    for j in range(cfg.n_stages - 1):
        skip = skips[len(skips) - 1 - j]
        x = rb(x, q)
        if x.shape[2:] != skip.shape[2:]:
            x = F.interpolate(x, size=skip.shape[2:], mode="bilinear", align_corners=False)  # unet.py:220-225
        x = torch.cat([rb(x, q), skip], dim=1)  # unet.py:228 -- upsampled first, skip second
        x = conv_block(x, sd, f"decoder_stages.{j}.conv_block", 1, cfg.decoder_dropout[j], cfg, masks, training, q)
    if "reconstruction_output.0.weight" in sd:
        # the autoencoder variant (AE_pretrained/reconstruction/models/autoencoder.py:374-387, :436):
        # Conv2d(32 -> 3, 3x3, pad 1, bias) + Sigmoid instead of the 1x1 segmentation head
        pre = F.conv2d(rb(x, q), rb(sd["reconstruction_output.0.weight"], q), None, padding=1)
        return torch.sigmoid(rb(pre, q) + sd["reconstruction_output.0.bias"].view(1, -1, 1, 1))
    return F.conv2d(rb(x, q), sd["segmentation_output.weight"], sd["segmentation_output.bias"])


def autoencoder_training_step(sd: Dict[str, torch.Tensor], x: torch.Tensor, target: torch.Tensor,
                              cfg: "UNetConfig", masks=None, training: bool = True, bf16_storage: bool = False,
                              dtype: torch.dtype = torch.float32):
    """One reference autoencoder training step on CPU (AE_pretrained/reconstruction/src/train.py:527-549: forward,
    nn.MSELoss (train.py:431), backward).  Returns dict(output, loss, grads)."""
    leaves = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    if masks is not None:
        masks = [m.to(dtype) for m in masks]
    out = unet_forward(leaves, x.to(dtype), cfg, masks, training, bf16_storage)
    loss = F.mse_loss(out, target.to(dtype))
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return dict(output=out.detach(), loss=loss.detach(), grads=grads)


def class_weights(target: torch.Tensor, ignore_index: int = 255) -> torch.Tensor:
    """SimpleLoss._compute_class_weights (Our_UNet/models/losses.py:24-62): inverse frequency over valid pixels,
    exactly-zero counts set to 1 (:53-54), normalised to sum to 3 (:60).  Hard-coded 3 classes (:40)."""
    mask = target != ignore_index
    total = mask.sum().float()
    cnt = torch.stack([((target == c) & mask).sum().float() for c in range(3)])
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    w = total / cnt
    return w * (3 / w.sum())


def simple_loss(logits: torch.Tensor, target: torch.Tensor, weight_dice=1.0, weight_ce=1.0, ignore_index=255,
                smooth=1e-5, weights: Optional[torch.Tensor] = None, dynamic=True, parts=False):
    """SimpleLoss.forward (losses.py:64-82): weight_ce * CrossEntropy(weight=w, ignore_index) + weight_dice * Dice
    (losses.py:84-121).  `weights` = static class weights used when dynamic is False (train.py:851-858)."""
    if logits.shape[-2:] != target.shape[-2:]:
        logits = F.interpolate(logits, size=target.shape[-2:], mode="bilinear", align_corners=False)  # :66-68
    if dynamic and target.size(0) > 0:
        weights = class_weights(target, ignore_index)  # :71-73
    ce = F.cross_entropy(logits, target, weight=weights, ignore_index=ignore_index)  # :76
    mask = (target != ignore_index).float()
    p = F.softmax(logits, dim=1)  # :92
    dice = 0
    for c in range(logits.shape[1]):  # :98-118
        t_c = (target == c).float() * mask
        p_c = p[:, c] * mask
        inter = (p_c * t_c).reshape(p_c.size(0), -1).sum(1)
        union = p_c.reshape(p_c.size(0), -1).sum(1) + t_c.reshape(t_c.size(0), -1).sum(1)
        dice = dice + (1.0 - ((2.0 * inter + smooth) / (union + smooth)).mean())
    dice = dice / logits.shape[1]  # :121
    total = weight_ce * ce + weight_dice * dice
    return (total, ce, dice) if parts else total


def simple_loss_numpy(logits: np.ndarray, target: np.ndarray, weight_dice=1.0, weight_ce=1.0, ignore_index=255,
                      smooth=1e-5):
    """Independent float64 restatement of SimpleLoss with dynamic weights (losses.py:24-121) using only numpy:
    returns (total, ce, dice, dlogits) with the analytic gradient of SURVEY.md A.4."""
    z = logits.astype(np.float64)
    B, C = z.shape[:2]
    t = target
    valid = t != ignore_index
    cnt_raw = np.array([np.sum((t == c) & valid) for c in range(3)], dtype=np.float64)
    cnt = np.where(cnt_raw == 0, 1.0, cnt_raw)
    w = valid.sum() / cnt
    w = w * (3.0 / w.sum())
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    p = e / e.sum(axis=1, keepdims=True)
    logp = (z - m) - np.log(e.sum(axis=1, keepdims=True))
    onehot = np.stack([(t == c) & valid for c in range(C)], axis=1).astype(np.float64)
    wpix = sum(w[c] * onehot[:, c] for c in range(3))
    W = wpix.sum()
    ce = -(wpix[:, None] * onehot * logp).sum() / W
    vm = valid[:, None].astype(np.float64)
    I = (p * onehot).sum(axis=(2, 3))
    U = (p * vm).sum(axis=(2, 3)) + onehot.sum(axis=(2, 3))
    d = (2 * I + smooth) / (U + smooth)
    dice = np.mean(1.0 - d.mean(axis=0))
    # gradient
    g_ce = wpix[:, None] * (p * valid[:, None] - onehot) / W
    G = -(1.0 / (C * B)) * (2 * onehot * (U + smooth)[:, :, None, None] - (2 * I + smooth)[:, :, None, None]) \
        / ((U + smooth) ** 2)[:, :, None, None] * vm
    gp = (G * p).sum(axis=1, keepdims=True)
    g_dice = p * (G - gp)
    dz = weight_ce * g_ce + weight_dice * g_dice
    return weight_ce * ce + weight_dice * dice, ce, dice, dz


def synthetic_batch(batch: int, size: int = 512, seed: int = 0, variant: str = "uniform", width: Optional[int] = None):
    """BASELINE.md section 3 inputs: image ~ N(0,1) fp32 [B,3,H,W]; mask in {0,1,2} with ~10 % 255.
    variant "pets": each image has background + ONE foreground class (cat=1 or dog=2), like the real dataset;
    variant "cats": class 2 absent from the whole batch (exercises the n_c == 0 -> 1 clamp, losses.py:53-54)."""
    g = torch.Generator().manual_seed(seed)
    width = width or size
    image = torch.randn(batch, 3, size, width, generator=g)
    mask = torch.randint(0, 3, (batch, size, width), generator=g)
    mask[torch.rand(batch, size, width, generator=g) < 0.1] = 255
    if variant in ("pets", "cats"):
        for b in range(batch):
            fg = 1 if (variant == "cats" or b % 2 == 0) else 2
            m = mask[b]
            m[(m != 255) & (m != 0)] = fg
    elif variant != "uniform":
        raise ValueError(variant)
    return image, mask


def training_step(sd: Dict[str, torch.Tensor], x: torch.Tensor, target: torch.Tensor, cfg: UNetConfig = UNetConfig(),
                  masks: Optional[List[torch.Tensor]] = None, training: bool = True, loss_kwargs: Optional[dict] = None,
                  bf16_storage: bool = False, dtype: torch.dtype = torch.float32,
                  clip_features: Optional[torch.Tensor] = None):
    """One reference training step on CPU fp32 (train.py:654-663: forward, SimpleLoss, backward) through the
    restatement above, differentiated by torch autograd.  Returns dict(logits, loss, ce, dice, grads).
    `bf16_storage=True` keeps the arithmetic but rounds to bf16 exactly where the CUDA path stores bf16 (conv
    operands, raw conv outputs, activations, upsampled tensors and their gradients): the matched-precision oracle
    that separates kernel error from the cost of bf16 storage.  The default is the reference's fp32.
    `dtype=torch.float64` runs the same ops in double: the "exact" answer that tells how much of a difference between
    two fp32 runs is the conditioning of the problem (at random init this network amplifies rounding noise ~1e5 x:
    the reference's own fp32 gradients sit up to 1e-2 from the fp64 ones, tests/test_gpu_fp32_mode.py)."""
    leaves = {}
    for k, v in sd.items():
        leaves[k] = v.detach().clone().to(dtype).requires_grad_(True)
    if masks is not None:
        masks = [m.to(dtype) for m in masks]
    logits = unet_forward(leaves, x.to(dtype), cfg, masks, training, bf16_storage, clip_features=clip_features)
    kw = dict(loss_kwargs or {})
    if dtype != torch.float32 and kw.get("dynamic", True) and kw.get("weights") is None:
        # the reference computes the class weights in fp32 from integer counts (losses.py:44-60); keep that and only
        # lift them to the working precision
        kw.update(weights=class_weights(target, kw.get("ignore_index", 255)).to(dtype), dynamic=False)
    total, ce, dice = simple_loss(logits, target, parts=True, **kw)
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return dict(logits=logits.detach(), loss=total.detach(), ce=ce.detach(), dice=dice.detach(), grads=grads)


def config_of(model) -> UNetConfig:
    """UNetConfig describing a constructed UNet module (reference or drop-in): read off the module tree."""
    feats = list(model.features_per_stage)
    strides, enc_p, dec_p = [], [], []

    def rate(block):
        for m in block.block:
            if hasattr(m, "drop_prob"):
                return float(m.drop_prob)
        return 0.0

    for st in model.encoder_stages:
        strides.append(int(st.block[0].stride[0]))
        enc_p.append(rate(st))
    for up in model.decoder_stages:
        dec_p.append(rate(up.conv_block))
    norm = model.encoder_stages[0].block[1]
    act = model.encoder_stages[0].block[2]
    return UNetConfig(in_channels=model.in_channels, num_classes=model.num_classes, features=feats, strides=strides,
                      encoder_dropout=enc_p, decoder_dropout=dec_p, eps=float(norm.eps),
                      negative_slope=float(act.negative_slope))


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((got - ref).norm() / (ref.norm() + 1e-30))
