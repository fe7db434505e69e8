"""Size-independent properties of the CUDA path at BASELINE.json's FULL configuration (default `UNet()`, batch 32,
512x512, bf16) -- the size at which the CPU oracle takes minutes, so parity is established through what must hold at
any size: run-to-run determinism (bit-exact), exact linearity of backward in the upstream gradient (a power-of-two
loss scale must scale every gradient bit-exactly), exactly-zero gradients of the conv biases that feed an InstanceNorm
(SURVEY.md 8a), per-sample independence (InstanceNorm and SpatialDropout2d are per (sample, channel), unet.py:13-35,
:118-119: an image's logits do not depend on its batch mates), the fused argmax + Dice counters against torch.argmax
(bit-exact, train.py:554-572), and the loss' invariance to a per-pixel logit shift (softmax, losses.py:73-118).
The small-size parity against the reference's own outputs is tests/test_gpu_model.py / test_gpu_fp32_mode.py."""
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu

B, S = 32, 512


@pytest.fixture(scope="module")
def full():
    from unet_implementations_b200.models.unet import UNet
    torch.manual_seed(1234)
    model = UNet().cuda().train()
    cfg = O.config_of(model)
    x, target = O.synthetic_batch(B, S, seed=0)
    torch.manual_seed(99)
    masks = O.draw_dropout_masks(cfg, B, x)
    return dict(model=model, x=x.cuda(), target=target.cuda(), masks=masks)


def _step(model, x, target, masks, scale=1.0):
    from unet_implementations_b200.models.losses import SimpleLoss
    model._mask_override = masks
    model.zero_grad(set_to_none=True)
    logits = model(x)
    loss = SimpleLoss()(logits, target)
    (loss * scale if scale != 1.0 else loss).backward()
    return logits.detach(), loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def test_full_size_step_is_deterministic_and_linear_in_the_loss_scale(full):
    m = full["model"]
    torch.cuda.reset_peak_memory_stats()
    l1, loss1, g1 = _step(m, full["x"], full["target"], full["masks"])
    l2, loss2, g2 = _step(m, full["x"], full["target"], full["masks"])
    assert torch.isfinite(loss1) and 0.5 < float(loss1) < 10.0
    assert torch.equal(l1, l2) and torch.equal(loss1, loss2)          # no atomics anywhere on the path
    assert all(torch.equal(g1[k], g2[k]) for k in g1)
    assert torch.cuda.max_memory_allocated() < 20 * 2 ** 30           # 11-12 GiB at this size; 180 GB per GPU
    _, _, g4 = _step(m, full["x"], full["target"], full["masks"], scale=4.0)
    bad = [k for k in g1 if not torch.equal(g4[k], g1[k] * 4.0)]      # bf16 / fp32 rounding commutes with a scale of 4
    assert not bad, bad[:5]
    for k, p in m.named_parameters():                                  # biases that feed an InstanceNorm: exact zeros
        if k.endswith(".bias") and not k.startswith("segmentation_output") and \
                isinstance(m.get_submodule(k.rsplit(".", 1)[0]), torch.nn.Conv2d):
            assert float(g1[k].abs().max()) == 0.0, k
        else:
            assert torch.isfinite(g1[k]).all() and float(g1[k].abs().max()) > 0.0, k


def test_full_size_logits_do_not_depend_on_batch_mates(full):
    m = full["model"]
    sel = [3, 17, 30]
    out = {}
    try:
        for prec in ("fp32", "bf16"):
            m.precision = prec
            with torch.no_grad():
                m._mask_override = full["masks"]
                big = m(full["x"])[sel].float()
                m._mask_override = [mk[sel] for mk in full["masks"]]
                small = m(full["x"][sel]).float()
            out[prec] = (O.rel_l2(big, small), (big.argmax(1) == small.argmax(1)).float().mean().item())
    finally:
        m.precision = "bf16"
        m._mask_override = None
    # A different batch size changes the tile-to-CTA partition, i.e. only the ORDER of the statistics partial sums.
    # fp32 verification mode (same kernels' algorithm, double accumulation): the images are independent to rounding.
    assert out["fp32"][0] <= 1e-5 and out["fp32"][1] >= 0.9999, out
    # bf16: a 1e-7 change of a mean flips a few bf16 roundings, and at random init this network amplifies that like any
    # other rounding difference -- bounded by the distance between two bf16 evaluations of the same function (the
    # reference under bf16 autocast sits 7.6e-2 from its own fp32 logits at this size, __graft_entry__.smoke())
    assert out["bf16"][0] <= 8e-2 and out["bf16"][1] >= 0.97, out


def test_full_size_argmax_counters_and_loss_shift_invariance(full):
    from unet_implementations_b200.metrics import argmax_counts
    from unet_implementations_b200.models.losses import SimpleLoss
    m = full["model"]
    with torch.no_grad():
        m._mask_override = full["masks"]
        logits = m(full["x"]).float()
    m._mask_override = None
    target = full["target"]
    pred, counts = argmax_counts(logits, target)
    ref = torch.argmax(logits, dim=1)
    assert torch.equal(pred, ref)                                      # bit-exact, ties to the lowest index
    valid = target != 255
    for c in range(3):
        assert int(counts[c, 0]) == int(((ref == c) & (target == c) & valid).sum())
        assert int(counts[c, 1]) == int(((ref == c) & valid).sum())
        assert int(counts[c, 2]) == int(((target == c) & valid).sum())
    assert int(counts[:, 2].sum()) == int(valid.sum())
    loss_fn = SimpleLoss()
    base = loss_fn(logits, target)
    shift = torch.randn(B, 1, S, S, device="cuda")                     # softmax(z + s) == softmax(z) per pixel
    moved = loss_fn(logits + shift, target)
    assert abs(float(moved) - float(base)) <= 1e-5 * abs(float(base)) + 1e-6


# (Cin, Cout, stride, H = W) of the benchmark step at batch 32: one layer per tensor-core kernel family / dispatch path
FULL_LAYERS = [
    ("e0c2", 32, 32, 1, 512),    # pair-row kernel fprop + dgrad, wgradn<32,32>
    ("d4c1", 96, 32, 1, 512),    # nconv with the zero-filled K tail, gconv<32,96> dgrad, wgradn over three chunks
    ("e1c1", 32, 64, 2, 512),    # stride 2: parity sub-lattices, parity-stacked dgrad
    ("d3c1", 192, 64, 1, 256),   # four M tiles per weight tile, N = 192 dgrad, wgradn<64,64>
    ("e2c2", 128, 128, 1, 128),  # two M tiles per weight tile, wgrad<64,128>
    ("d1c1", 768, 256, 1, 64),   # streamed weights N = 256, wgrad<64,256>
    ("e5c2", 512, 512, 1, 16),   # two N tiles, fewer tiles than SMs
]


@pytest.mark.parametrize("name,cin,cout,stride,hw", FULL_LAYERS)
def test_full_size_tensor_core_convs_against_the_direct_kernels(name, cin, cout, stride, hw):
    """At the real layer shapes (batch 32) the CPU oracle is out of reach; the independent implementation is the
    direct CUDA-core convolution (conv_simt.cu: one thread per output, plain loops, no tiling / TMA / swizzle), itself
    checked against F.conv2d at small sizes (tests/test_gpu_ops.py).  Same bf16 operands, fp32 accumulation in both:
    outputs agree to accumulation-order noise; the InstanceNorm partial sums equal the sums of the stored output."""
    from unet_implementations_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, hw, hw, cin, device="cuda", generator=g).bfloat16()
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    wf, wd = ops.pack_conv_weights(wt)
    y, stats = ops.conv_fprop(x, wf, stride)
    y_ref, _ = ops.conv_fprop(x, wf, stride, want_stats=False, simt=True)
    assert O.rel_l2(y.float(), y_ref.float()) <= 2e-3
    yf = y.float()
    s_ref = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1)
    assert O.rel_l2(stats.sum(dim=1), s_ref) <= 1e-4
    oh = y.shape[1]
    dy = torch.randn(B, oh, oh, cout, device="cuda", generator=g).bfloat16()
    dx_ref = ops.conv_dgrad(dy, wd, (hw, hw), stride, simt=True)
    assert O.rel_l2(ops.conv_dgrad(dy, wd, (hw, hw), stride).float(), dx_ref.float()) <= 2e-3
    if stride == 2 and cin <= 64:  # the form the model uses for these layers
        dx2 = torch.empty_like(dx_ref)
        ops.conv_dgrad_s2(dy, ops.pack_s2_dgrad_weights(wd), (hw, hw), out=dx2)
        assert O.rel_l2(dx2.float(), dx_ref.float()) <= 2e-3
    assert O.rel_l2(ops.conv_wgrad(x, dy, stride), ops.conv_wgrad(x, dy, stride, simt=True)) <= 1e-3


@pytest.mark.parametrize("name,cin,cout,stride,hw", FULL_LAYERS)
def test_full_size_tensor_core_convs_against_torch_conv2d_on_the_gpu(name, cin, cout, stride, hw):
    """The independent oracle at the benchmark's layer shapes (batch 32): the reference's own operator -- `F.conv2d`
    and its autograd (Our_UNet/models/unet.py:106-115) -- in fp32 on the same GPU with TF32 off, fed the same
    bf16-rounded operands.  Test infrastructure only; nothing on the product path calls torch's convolution."""
    import torch.nn.functional as F
    from unet_implementations_b200 import ops
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda").manual_seed(12)
        x = torch.randn(B, hw, hw, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5).bfloat16().float()
        wf, wd = ops.pack_conv_weights(wt)
        y, _ = ops.conv_fprop(x, wf, stride)
        xn = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        wr = wt.clone().requires_grad_(True)
        y_ref = F.conv2d(xn, wr, None, stride, 1)
        assert O.rel_l2(y.float().permute(0, 3, 1, 2), y_ref.detach()) <= 4e-3
        oh = y.shape[1]
        dy = torch.randn(B, oh, oh, cout, device="cuda", generator=g).bfloat16()
        y_ref.backward(dy.float().permute(0, 3, 1, 2))
        del y_ref
        if stride == 2 and cin <= 64:
            dx = torch.empty((B, hw, hw, cin), dtype=torch.bfloat16, device="cuda")
            ops.conv_dgrad_s2(dy, ops.pack_s2_dgrad_weights(wd), (hw, hw), out=dx)
        else:
            dx = ops.conv_dgrad(dy, wd, (hw, hw), stride)
        assert O.rel_l2(dx.float().permute(0, 3, 1, 2), xn.grad) <= 4e-3
        assert O.rel_l2(ops.conv_wgrad(x, dy, stride), wr.grad) <= 1e-3
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _act(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).bfloat16()


def test_full_size_norm_kernels_against_torch_fp32():
    """InstanceNorm + LeakyReLU + channel dropout forward / backward (unet.py:118-127, :22-35) at the largest tensor of the
    step ([32, 512, 512, 32] bf16, with the skip-connection's second gradient operand) against torch's own fp32 ops and
    autograd run on the same GPU from the same bf16-rounded operands."""
    import torch.nn.functional as F

    from unet_implementations_b200 import ops
    c, p = 32, 0.2
    y = _act((B, S, S, c), 1, 1.7)
    g = torch.Generator(device="cuda").manual_seed(2)
    gamma = (torch.rand(c, device="cuda", generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, device="cuda", generator=g) * 0.2).requires_grad_(True)
    drop = (torch.rand(B, c, device="cuda", generator=g) > p).float().div(1 - p)
    yf = y.float()
    stats = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], -1).unsqueeze(1).contiguous()
    mean, rstd, a, b = ops.in_finalize(stats, gamma.detach(), beta.detach(), drop, 1e-5, S * S)
    z = ops.in_apply(y, a, b, 0.01)
    yr = yf.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    del yf
    zr = F.leaky_relu(F.instance_norm(yr, weight=gamma, bias=beta, eps=1e-5), 0.01) * drop[:, :, None, None]
    assert O.rel_l2(z.float(), zr.detach().permute(0, 2, 3, 1)) <= 4e-3
    dz, dz2 = _act((B, S, S, c), 3), _act((B, S, S, c), 4)
    dy, dg, db = ops.in_backward(dz, dz2, y, a, b, mean, rstd, drop, gamma.detach(), 0.01)
    zr.backward((dz.float() + dz2.float()).permute(0, 3, 1, 2))
    assert O.rel_l2(dy.float(), yr.grad.permute(0, 2, 3, 1)) <= 6e-3
    assert O.rel_l2(dg, gamma.grad) <= 1e-3 and O.rel_l2(db, beta.grad) <= 1e-3


def test_full_size_upsample_and_head_against_torch_fp32():
    """The 256^2 -> 512^2 upsample into the 96-channel concat buffer (with the producer's norm apply fused in) and the
    1x1 head forward / backward with the fused apply ([32, 512, 512, 32] -> fp32 NCHW logits), against
    F.interpolate / F.conv2d and autograd in fp32 on the same GPU."""
    import torch.nn.functional as F

    from unet_implementations_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    # ---- upsample2x with (a, b, slope)
    c, h = 64, S // 2
    y = _act((B, h, h, c), 6, 1.5)
    a = (torch.rand(B, c, device="cuda", generator=g) + 0.3) * (torch.rand(B, c, device="cuda", generator=g) > 0.2)
    b = torch.randn(B, c, device="cuda", generator=g) * (a != 0)
    cat = torch.full((B, S, S, c + 32), 5.0, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x(y, cat[..., :c], norm=(a.contiguous(), b.contiguous(), 0.01))
    zr = F.leaky_relu(y.float() * a[:, None, None, :] + b[:, None, None, :], 0.01).permute(0, 3, 1, 2).requires_grad_(True)
    ur = F.interpolate(zr, size=(S, S), mode="bilinear", align_corners=False)
    assert O.rel_l2(cat[..., :c].float(), ur.detach().permute(0, 2, 3, 1)) <= 4e-3
    assert float((cat[..., c:].float() - 5.0).abs().max()) == 0.0     # the skip half of the concat buffer is untouched
    dcat = _act((B, S, S, c + 32), 7)
    dx = ops.upsample2x_backward(dcat[..., :c])
    ur.backward(dcat[..., :c].float().permute(0, 3, 1, 2))
    assert O.rel_l2(dx.float(), zr.grad.permute(0, 2, 3, 1)) <= 4e-3
    del cat, dcat, ur, zr, dx
    # ---- head with (a, b, slope)
    c = 32
    y = _act((B, S, S, c), 8, 1.5)
    a = torch.rand(B, c, device="cuda", generator=g) + 0.3
    b = torch.randn(B, c, device="cuda", generator=g)
    wt = (torch.randn(3, c, 1, 1, device="cuda", generator=g) * 0.3).requires_grad_(True)
    bias = torch.randn(3, device="cuda", generator=g).requires_grad_(True)
    zr = F.leaky_relu(y.float() * a[:, None, None, :] + b[:, None, None, :], 0.01).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # the checker must be real fp32: cuDNN's default TF32 conv is 2e-4 off
    try:
        lr = F.conv2d(zr, wt, bias)
        logits = ops.head_forward(y, wt.detach(), bias.detach(), norm=(a, b, 0.01))
        assert O.rel_l2(logits, lr.detach()) <= 1e-4
        dl = torch.randn(B, 3, S, S, device="cuda", generator=g) * 1e-3
        lr.backward(dl)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    dz, dw, db = ops.head_backward(dl, y, wt.detach(), norm=(a, b, 0.01))
    assert O.rel_l2(dz.float(), zr.grad.permute(0, 2, 3, 1)) <= 4e-3
    assert O.rel_l2(dw, wt.grad) <= 1e-3 and O.rel_l2(db, bias.grad) <= 1e-3


def test_full_size_loss_against_the_restated_reference_loss():
    """SimpleLoss forward / backward on [32, 3, 512, 512] logits + int64 targets with ignored pixels against the
    restatement of losses.py:24-121 (oracle.simple_loss, itself pinned to the reference's committed values at small
    sizes) evaluated in float64 on the same GPU."""
    from unet_implementations_b200.models.losses import SimpleLoss
    g = torch.Generator(device="cuda").manual_seed(9)
    logits = (torch.randn(B, 3, S, S, device="cuda", generator=g) * 2.0).requires_grad_(True)
    _, target = O.synthetic_batch(B, S, seed=3)
    target = target.cuda()
    loss = SimpleLoss()(logits, target)
    loss.backward()
    ref_in = logits.detach().double().requires_grad_(True)
    w = O.class_weights(target).double()
    ref = O.simple_loss(ref_in, target, weights=w, dynamic=False)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert O.rel_l2(logits.grad.double(), ref_in.grad) <= 1e-4
