"""GPU parity of each kernel family through the C ABI against plain CPU fp32 restatements of the reference ops
(F.conv2d / F.instance_norm / F.leaky_relu / F.interpolate and their autograd), on identical bf16-rounded operands.
Tolerance: outputs stored in bf16 -> rel-L2 <= 4e-3 (bf16 rounding is 2^-9 per element); fp32 outputs -> 1e-4."""
import pytest
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3
F32_TOL = 1e-4


def rand_act(n, h, w, c, pitch=None, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    pitch = pitch or c
    buf = (torch.randn(n, h, w, pitch, generator=g) * scale).bfloat16()
    dev = buf.cuda()
    return (dev[..., :c] if pitch != c else dev), buf[..., :c].float()


# every distinct (Cin, Cout, stride) of the model at reduced spatial size, plus ragged sizes
CONV_CASES = [
    (2, 16, 16, 32, 32, 1), (2, 16, 16, 32, 64, 2), (2, 16, 16, 64, 64, 1), (2, 16, 16, 64, 128, 2),
    (1, 16, 16, 128, 128, 1), (1, 16, 16, 128, 256, 2), (1, 8, 16, 256, 256, 1), (1, 16, 16, 256, 512, 2),
    (1, 8, 8, 512, 512, 1), (1, 8, 8, 512, 512, 2), (1, 8, 8, 1024, 512, 1), (1, 8, 8, 768, 256, 1),
    (1, 8, 16, 384, 128, 1), (1, 16, 16, 192, 64, 1), (2, 16, 32, 96, 32, 1),
    (2, 13, 21, 64, 64, 1), (2, 13, 21, 64, 96, 2), (1, 5, 7, 32, 32, 1), (3, 2, 2, 64, 64, 2),
    # stride 2 with Cin in {32, 64}: parity-stacked dgrad (even, odd and ragged sizes, BK = 32 and 64)
    (2, 16, 32, 32, 64, 2), (1, 13, 21, 32, 96, 2), (2, 18, 10, 64, 128, 2), (1, 7, 9, 32, 32, 2),
    # wide images: the narrow-output kernels (column taps stacked on N, 30-of-32 column tiles, ragged edges)
    (1, 12, 70, 32, 32, 1), (1, 9, 64, 96, 32, 1), (2, 8, 96, 64, 64, 1), (1, 6, 121, 64, 32, 1), (1, 7, 90, 32, 64, 1),
    # 32 -> 32 on dense tensors at least 128 wide: pixel pairs as operand rows (60-of-64 column tiles); odd width -> nconv
    (1, 12, 128, 32, 32, 1), (2, 9, 190, 32, 32, 1), (3, 5, 250, 32, 32, 1), (1, 6, 129, 32, 32, 1),
    # streamed-weight layers as CTA pairs (cta_group::2, an even number of 16 x 8 tiles per image): several pair tiles,
    # ragged edges, two M tiles per CTA (N = 128), N = 192 data gradient; an odd tile count falls back to single CTAs
    (2, 32, 32, 256, 256, 1), (1, 16, 24, 256, 256, 1), (1, 48, 16, 128, 128, 1), (1, 40, 20, 384, 128, 1),
    (1, 70, 12, 192, 64, 1),
]


@pytest.mark.parametrize("n,h,w,cin,cout,stride", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(n, h, w, cin, cout, stride):
    from unet_implementations_b200 import ops
    x, x_ref = rand_act(n, h, w, cin, seed=1)
    g = torch.Generator().manual_seed(2)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    wf, wd = ops.pack_conv_weights(wt.cuda())
    w_ref = wt.bfloat16().float().requires_grad_(True)
    xr = x_ref.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yr = F.conv2d(xr, w_ref, padding=1, stride=stride)
    y, stats = ops.conv_fprop(x, wf, stride)
    assert O.rel_l2(y.float(), yr.permute(0, 2, 3, 1)) <= BF16_TOL
    yb = y.float().cpu()
    s_ref = torch.stack([yb.sum(dim=(1, 2)), (yb * yb).sum(dim=(1, 2))], dim=-1)
    assert O.rel_l2(stats.sum(dim=1), s_ref) <= F32_TOL  # the epilogue's InstanceNorm partial sums
    oh, ow = y.shape[1], y.shape[2]
    dy, dy_ref = rand_act(n, oh, ow, cout, seed=3)
    yr.backward(dy_ref.permute(0, 3, 1, 2))
    dx = ops.conv_dgrad(dy, wd, (h, w), stride)
    assert O.rel_l2(dx.float(), xr.grad.permute(0, 2, 3, 1)) <= BF16_TOL
    ws2 = ops.pack_s2_dgrad_weights(wd) if stride == 2 else None
    if ws2 is not None:  # stride 2, Cin in {32, 64}: the parity-stacked single-launch form, written into a pitched buffer
        buf = torch.full((n, h, w, cin + 8), 3.0, dtype=torch.bfloat16, device="cuda")
        ops.conv_dgrad_s2(dy, ws2, (h, w), out=buf[..., :cin])
        assert O.rel_l2(buf[..., :cin].float(), xr.grad.permute(0, 2, 3, 1)) <= BF16_TOL
        assert float((buf[..., cin:].float() - 3).abs().max()) == 0
    dw = ops.conv_wgrad(x, dy, stride)
    assert dw.dtype == torch.float32 and tuple(dw.shape) == (cout, cin, 3, 3)
    assert O.rel_l2(dw, w_ref.grad) <= F32_TOL


def test_conv_32_to_32_wide_on_channel_slices():
    """The pair-row kernel needs dense tensors; a 32-channel slice of a wider buffer (in or out) must take the generic
    narrow kernel and still be right, leaving the other channels untouched."""
    from unet_implementations_b200 import ops
    xb, xb_ref = rand_act(1, 8, 136, 96, seed=14)
    x, x_ref = xb[..., 64:96], xb_ref[..., 64:96]
    g = torch.Generator().manual_seed(15)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.08
    wf, _ = ops.pack_conv_weights(wt.cuda(), need_dgrad=False)
    yr = F.conv2d(x_ref.permute(0, 3, 1, 2), wt.bfloat16().float(), padding=1).permute(0, 2, 3, 1)
    out = torch.full((1, 8, 136, 64), 7.0, dtype=torch.bfloat16, device="cuda")
    y, stats = ops.conv_fprop(x, wf, 1, out=out[..., 32:64])
    assert O.rel_l2(out[..., 32:64].float(), yr) <= BF16_TOL
    assert float((out[..., :32].float() - 7).abs().max()) == 0
    y2, _ = ops.conv_fprop(x, wf, 1)  # pitched in, dense out
    assert O.rel_l2(y2.float(), yr) <= BF16_TOL


def test_conv_reads_and_writes_channel_slices_of_a_concat_buffer():
    """Pitched views: the operand is channels [64,128) of a 192-wide buffer and the output lands in channels
    [32,96) of a 128-wide buffer whose other channels must stay untouched (how torch.cat disappears, unet.py:228)."""
    from unet_implementations_b200 import ops
    xb, xb_ref = rand_act(2, 16, 16, 192, seed=4)
    x, x_ref = xb[..., 64:128], xb_ref[..., 64:128]
    g = torch.Generator().manual_seed(5)
    wt = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    wf, _ = ops.pack_conv_weights(wt.cuda(), need_dgrad=False)
    out = torch.full((2, 16, 16, 128), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.conv_fprop(x, wf, 1, out=out[..., 32:96])
    ref = F.conv2d(x_ref.permute(0, 3, 1, 2), wt.bfloat16().float(), padding=1).permute(0, 2, 3, 1)
    assert O.rel_l2(out[..., 32:96].float(), ref) <= BF16_TOL
    assert float((out[..., :32].float() - 7).abs().max()) == 0 and float((out[..., 96:].float() - 7).abs().max()) == 0


def test_simt_conv_agrees_with_tensor_core_conv():
    from unet_implementations_b200 import ops
    x, _ = rand_act(2, 12, 20, 64, seed=6)
    g = torch.Generator().manual_seed(7)
    wt = (torch.randn(96, 64, 3, 3, generator=g) * 0.06).cuda()
    wf, wd = ops.pack_conv_weights(wt)
    for stride in (1, 2):
        y_tc, _ = ops.conv_fprop(x, wf, stride)
        y_si, _ = ops.conv_fprop(x, wf, stride, want_stats=False, simt=True)
        assert O.rel_l2(y_tc.float(), y_si.float()) <= 2e-3
        dy, _ = rand_act(2, y_tc.shape[1], y_tc.shape[2], 96, seed=8)
        assert O.rel_l2(ops.conv_dgrad(dy, wd, (12, 20), stride).float(),
                        ops.conv_dgrad(dy, wd, (12, 20), stride, simt=True).float()) <= 2e-3
        assert O.rel_l2(ops.conv_wgrad(x, dy, stride), ops.conv_wgrad(x, dy, stride, simt=True)) <= F32_TOL


def test_stem_conv_fprop_wgrad():
    from unet_implementations_b200 import ops
    n, h, w = 2, 48, 80
    g = torch.Generator().manual_seed(5)
    img = torch.randn(n, 3, h, w, generator=g)
    wt = (torch.randn(32, 3, 3, 3, generator=g) * 0.08).requires_grad_(True)
    y, stats = ops.stem_fprop(img.cuda(), wt.detach().cuda())
    yr = F.conv2d(img, wt, padding=1)
    assert O.rel_l2(y.float(), yr.permute(0, 2, 3, 1)) <= BF16_TOL
    yb = y.float().cpu()
    assert O.rel_l2(stats.sum(dim=1), torch.stack([yb.sum(dim=(1, 2)), (yb * yb).sum(dim=(1, 2))], -1)) <= F32_TOL
    dy, dy_ref = rand_act(n, h, w, 32, seed=6)
    yr.backward(dy_ref.permute(0, 3, 1, 2))
    assert O.rel_l2(ops.stem_wgrad(img.cuda(), dy), wt.grad) <= F32_TOL
    # tensor-core route: the image operand is rounded to bf16 (one rounding of X: bf16 tolerance)
    assert O.rel_l2(ops.stem_wgrad_tc(img.cuda(), dy), wt.grad) <= BF16_TOL


@pytest.mark.parametrize("n,h,w,c,p,pitch", [(2, 16, 24, 32, 0.0, None), (3, 8, 8, 128, 0.3, 192),
                                            (2, 32, 32, 96, 0.2, None), (2, 2, 2, 512, 0.3, None),
                                            (1, 64, 64, 64, 0.1, 96),
                                            # >= 4 MB per image: the fused L2-resident backward (several image
                                            # groups, ragged pixel ranges, with and without the second gradient)
                                            (5, 256, 264, 32, 0.2, None), (3, 128, 136, 128, 0.3, 160),
                                            (3, 192, 200, 64, 0.0, 96)])
def test_instance_norm_lrelu_dropout_fwd_bwd(n, h, w, c, p, pitch):
    """unet.py:118-127 fused: IN(eps=1e-5, affine, biased variance) -> LeakyReLU(0.01) -> channel dropout."""
    from unet_implementations_b200 import ops
    y, y_ref = rand_act(n, h, w, c, pitch, seed=7, scale=1.7)
    g = torch.Generator().manual_seed(8)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, generator=g) * 0.2).requires_grad_(True)
    drop = (torch.rand(n, c, generator=g) > p).float().div(1 - p) if p > 0 else None
    stats = torch.stack([y_ref.sum(dim=(1, 2)), (y_ref * y_ref).sum(dim=(1, 2))], -1).unsqueeze(1).contiguous().cuda()
    dropd = drop.cuda() if drop is not None else None
    mean, rstd, a, b = ops.in_finalize(stats, gamma.detach().cuda(), beta.detach().cuda(), dropd, 1e-5, h * w)
    z = ops.in_apply(y, a, b, 0.01)
    yr = y_ref.permute(0, 3, 1, 2).clone().requires_grad_(True)
    zr = F.leaky_relu(F.instance_norm(yr, weight=gamma, bias=beta, eps=1e-5), 0.01)
    if drop is not None:
        zr = zr * drop[:, :, None, None]
    assert O.rel_l2(z.float(), zr.permute(0, 2, 3, 1)) <= BF16_TOL
    if drop is not None:  # dropped channels are exact zeros
        zc = z.float().cpu()
        assert float(zc[(drop == 0)[:, None, None, :].expand_as(zc)].abs().max()) == 0.0
    dz, dz_ref = rand_act(n, h, w, c, seed=9)
    dz2, dz2_ref = rand_act(n, h, w, c, seed=10) if c == 128 else (None, 0)
    dy, dg, db = ops.in_backward(dz, dz2, y, a, b, mean, rstd, dropd, gamma.detach().cuda(), 0.01)
    zr.backward((dz_ref + dz2_ref).permute(0, 3, 1, 2))
    assert O.rel_l2(dy.float(), yr.grad.permute(0, 2, 3, 1)) <= 6e-3
    assert O.rel_l2(dg, gamma.grad) <= 1e-3 and O.rel_l2(db, beta.grad) <= 1e-3


def test_upsample_into_concat_fwd_bwd():
    """F.interpolate(bilinear, align_corners=False) exact 2x (unet.py:220-225) into channels [0,C) of a concat buffer."""
    from unet_implementations_b200 import ops
    n, h, w, c = 2, 6, 10, 64
    x, x_ref = rand_act(n, h, w, c, seed=11)
    cat = torch.zeros(n, 2 * h, 2 * w, c + 32, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x(x, cat[..., :c])
    xr = x_ref.permute(0, 3, 1, 2).requires_grad_(True)
    ur = F.interpolate(xr, size=(2 * h, 2 * w), mode="bilinear", align_corners=False)
    assert O.rel_l2(cat[..., :c].float(), ur.permute(0, 2, 3, 1)) <= BF16_TOL
    assert float(cat[..., c:].float().abs().max()) == 0.0
    dcat, dcat_ref = rand_act(n, 2 * h, 2 * w, c + 32, seed=12)
    dx = ops.upsample2x_backward(dcat[..., :c])
    ur.backward(dcat_ref[..., :c].permute(0, 3, 1, 2))
    assert O.rel_l2(dx.float(), xr.grad.permute(0, 2, 3, 1)) <= BF16_TOL
    # edge clamp, checked on a ramp: out[0] = in[0], out[1] = .75 in[0] + .25 in[1]
    ramp = torch.arange(8, dtype=torch.float32).view(1, 1, 8, 1).expand(1, 8, 8, 8).contiguous().bfloat16().cuda()
    out = torch.empty(1, 16, 16, 8, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x(ramp, out)
    assert out[0, 0, :4, 0].float().tolist() == [0.0, 0.25, 0.75, 1.25]


def test_head_fwd_bwd():
    """segmentation_output: Conv2d(32 -> 3, 1x1) + bias (unet.py:374-381) from bf16 NHWC to fp32 NCHW logits."""
    from unet_implementations_b200 import ops
    n, h, w = 2, 24, 40
    z, z_ref = rand_act(n, h, w, 32, seed=13)
    g = torch.Generator().manual_seed(14)
    wt = (torch.randn(3, 32, 1, 1, generator=g) * 0.3).requires_grad_(True)
    bias = torch.randn(3, generator=g).requires_grad_(True)
    logits = ops.head_forward(z, wt.detach().cuda(), bias.detach().cuda())
    zr = z_ref.permute(0, 3, 1, 2).clone().requires_grad_(True)
    lr = F.conv2d(zr, wt, bias)
    assert O.rel_l2(logits, lr) <= 1e-5
    dl = torch.randn(n, 3, h, w, generator=g) * 1e-3
    lr.backward(dl)
    dz, dw, db = ops.head_backward(dl.cuda(), z, wt.detach().cuda())
    assert O.rel_l2(dz.float(), zr.grad.permute(0, 2, 3, 1)) <= BF16_TOL
    assert O.rel_l2(dw, wt.grad) <= F32_TOL and O.rel_l2(db, bias.grad) <= F32_TOL


def test_full_size_layer_properties():
    """512x512 (the BASELINE size), where the oracle would take too long: size-independent properties.
    (a) the fused IN output has per-plane mean beta and variance gamma^2 (before the LeakyReLU: slope 1);
    (b) conv is linear: conv(x1 + x2) == conv(x1) + conv(x2) up to bf16 rounding;
    (c) the epilogue statistics equal a recount of the stored tensor."""
    from unet_implementations_b200 import ops
    n, s, c = 2, 512, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(n, s, s, c, generator=g, device="cuda").bfloat16()
    x2 = torch.randn(n, s, s, c, generator=g, device="cuda").bfloat16()
    wt = torch.randn(c, c, 3, 3, generator=g, device="cuda") * 0.08
    wf, _ = ops.pack_conv_weights(wt, need_dgrad=False)
    y1, st = ops.conv_fprop(x1, wf, 1)
    y2, _ = ops.conv_fprop(x2, wf, 1)
    y12, _ = ops.conv_fprop((x1.float() + x2.float()).bfloat16(), wf, 1)
    assert O.rel_l2(y12.float(), y1.float() + y2.float()) <= 8e-3
    yb = y1.float()
    recount = torch.stack([yb.sum(dim=(1, 2)), (yb * yb).sum(dim=(1, 2))], -1)
    assert O.rel_l2(st.sum(dim=1), recount) <= 1e-4
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda")
    mean, rstd, a, b = ops.in_finalize(st, gamma, beta, None, 1e-5, s * s)
    z = ops.in_apply(y1, a, b, 1.0).float()  # slope 1 = identity activation
    assert float((z.mean(dim=(1, 2)) - beta).abs().max()) <= 5e-3
    assert float((z.var(dim=(1, 2), unbiased=False).sqrt() - gamma).abs().max()) <= 5e-3


def test_consumers_with_the_apply_pass_fused_in():
    """upsample2x / head with (a, b, slope): the producer's InstanceNorm+LeakyReLU(+dropout) apply pass runs inside its
    single consumer.  Against the unfused sequence on fp32 CPU math of the same bf16 raw conv output."""
    from unet_implementations_b200 import ops
    n, h, w, c = 2, 12, 20, 32
    y, y_ref = rand_act(n, h, w, c, seed=21, scale=1.5)
    g = torch.Generator().manual_seed(22)
    a = (torch.rand(n, c, generator=g) + 0.3) * torch.where(torch.rand(n, c, generator=g) < 0.2, 0.0, 1.0)  # dropped channels: a = b = 0
    b = torch.randn(n, c, generator=g) * (a != 0)
    z_ref = F.leaky_relu(y_ref * a[:, None, None, :] + b[:, None, None, :], 0.01)
    ad, bd = a.cuda().contiguous(), b.cuda().contiguous()
    # upsample
    out = torch.zeros(n, 2 * h, 2 * w, c + 32, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x(y, out[..., :c], norm=(ad, bd, 0.01))
    up_ref = F.interpolate(z_ref.permute(0, 3, 1, 2), size=(2 * h, 2 * w), mode="bilinear", align_corners=False)
    assert O.rel_l2(out[..., :c].float(), up_ref.permute(0, 2, 3, 1)) <= BF16_TOL
    assert float(out[..., c:].float().abs().max()) == 0.0
    # head forward / backward
    wt = torch.randn(3, c, 1, 1, generator=g).requires_grad_(True)
    bias = torch.randn(3, generator=g).requires_grad_(True)
    zr = z_ref.permute(0, 3, 1, 2).clone().requires_grad_(True)
    lr = F.conv2d(zr, wt, bias)
    logits = ops.head_forward(y, wt.detach().cuda(), bias.detach().cuda(), norm=(ad, bd, 0.01))
    assert O.rel_l2(logits, lr) <= F32_TOL
    dl = torch.randn(n, 3, h, w, generator=g)
    lr.backward(dl)
    dz, dw, db = ops.head_backward(dl.cuda(), y, wt.detach().cuda(), norm=(ad, bd, 0.01))
    assert O.rel_l2(dz.float(), zr.grad.permute(0, 2, 3, 1)) <= BF16_TOL
    assert O.rel_l2(dw, wt.grad) <= F32_TOL and O.rel_l2(db, bias.grad) <= F32_TOL


@pytest.mark.parametrize("n,h,w", [(2, 12, 20), (3, 64, 96), (1, 7, 5)])
def test_head_backward_reduces_the_producing_units_norm_backward_sums(n, h, w):
    """Producer-side sums (b200unet_in_bwd_args.ext_part): the head backward, which reads the last unit's raw output
    anyway, also emits sum gm and sum gm * y per image; the norm backward fed with them (no reduction pass) must give
    what the self-reducing norm backward gives (unet.py:118-127 backward)."""
    from unet_implementations_b200 import ops
    c = 32
    y, y_ref = rand_act(n, h, w, c, seed=31, scale=1.5)
    g = torch.Generator().manual_seed(32)
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    beta = (torch.randn(c, generator=g) * 0.3).cuda()
    drop = torch.where(torch.rand(n, c, generator=g) < 0.25, 0.0, 1.25).cuda()
    stats = torch.stack([y.float().sum((1, 2)), (y.float() ** 2).sum((1, 2))], -1).reshape(n, 1, c, 2).contiguous()
    mean, rstd, a, b = ops.in_finalize(stats, gamma, beta, drop, 1e-5, h * w)
    wt = torch.randn(3, c, 1, 1, generator=g).cuda()
    dl = torch.randn(n, 3, h, w, generator=g).cuda()
    dz0, dw0, db0 = ops.head_backward(dl, y, wt, norm=(a, b, 0.01))
    dz, dw, db, part = ops.head_backward(dl, y, wt, norm=(a, b, 0.01), want_bwd_part=True)
    assert torch.equal(dz, dz0) and O.rel_l2(dw, dw0) <= 1e-6 and O.rel_l2(db, db0) <= 1e-6
    # the sums themselves, against fp64 torch math on the stored dz
    pre = y.double() * a[:, None, None, :].double() + b[:, None, None, :].double()
    gm = dz.double() * torch.where(pre > 0, 1.0, 0.01)
    t1, t2 = gm.sum((1, 2)), (gm * y.double()).sum((1, 2))
    got = part.double().sum(1)
    scale = gm.abs().sum((1, 2)).clamp_min(1e-30)
    assert float(((got[..., 0] - t1).abs() / scale).max()) <= 1e-5
    assert float(((got[..., 1] - t2).abs() / (gm * y.double()).abs().sum((1, 2)).clamp_min(1e-30)).max()) <= 1e-5
    dy0, dg0, dbt0 = ops.in_backward(dz, None, y, a, b, mean, rstd, drop, gamma, 0.01)
    dy1, dg1, dbt1 = ops.in_backward(dz, None, y, a, b, mean, rstd, drop, gamma, 0.01, ext_part=part)
    assert O.rel_l2(dy1.float(), dy0.float()) <= 2e-3   # bf16 outputs of the same value up to fp32 summation order
    assert O.rel_l2(dg1, dg0) <= 1e-4 and O.rel_l2(dbt1, dbt0) <= 1e-4


@pytest.mark.parametrize("n,h,w,c", [(2, 40, 190, 32), (1, 9, 128, 32), (2, 21, 70, 64), (1, 64, 256, 64), (1, 7, 131, 32)])
def test_data_gradient_epilogue_reduces_the_consuming_units_norm_backward_sums(n, h, w, c):
    """Producer-side sums in the narrow-output data-gradient kernels (pair rows for dense 32 -> 32 tensors at W >= 128,
    column-stacked otherwise): dx must be unchanged, the partial sums must equal fp64 sums over the STORED dx, and the
    norm backward fed with them must agree with the self-reducing one."""
    from unet_implementations_b200 import ops
    g = torch.Generator().manual_seed(41)
    dy, _ = rand_act(n, h, w, c, seed=42)
    wt = torch.randn(c, c, 3, 3, generator=g) * (2.0 / (9 * c)) ** 0.5
    _, wd = ops.pack_conv_weights(wt.cuda())
    y, _ = rand_act(n, h, w, c, seed=43, scale=1.5)
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    beta = (torch.randn(c, generator=g) * 0.3).cuda()
    drop = torch.where(torch.rand(n, c, generator=g) < 0.25, 0.0, 1.25).cuda()
    stats = torch.stack([y.float().sum((1, 2)), (y.float() ** 2).sum((1, 2))], -1).reshape(n, 1, c, 2).contiguous()
    mean, rstd, a, b = ops.in_finalize(stats, gamma, beta, drop, 1e-5, h * w)
    dx0 = ops.conv_dgrad(dy, wd, (h, w), 1)
    dx, part = ops.conv_dgrad(dy, wd, (h, w), 1, bwd_sums=(y, a, b, 0.01))
    assert part is not None, "this shape is expected on the narrow-output kernels"
    assert torch.equal(dx, dx0)
    pre = y.double() * a[:, None, None, :].double() + b[:, None, None, :].double()
    gm = dx.double() * torch.where(pre > 0, 1.0, 0.01)
    got = part.double().sum(1)
    t1, t2 = gm.sum((1, 2)), (gm * y.double()).sum((1, 2))
    assert float(((got[..., 0] - t1).abs() / gm.abs().sum((1, 2)).clamp_min(1e-30)).max()) <= 1e-5
    assert float(((got[..., 1] - t2).abs() / (gm * y.double()).abs().sum((1, 2)).clamp_min(1e-30)).max()) <= 1e-5
    dy0, dg0, db0 = ops.in_backward(dx, None, y, a, b, mean, rstd, drop, gamma, 0.01)
    dy1, dg1, db1 = ops.in_backward(dx, None, y, a, b, mean, rstd, drop, gamma, 0.01, ext_part=part)
    assert O.rel_l2(dy1.float(), dy0.float()) <= 2e-3
    assert O.rel_l2(dg1, dg0) <= 1e-4 and O.rel_l2(db1, db0) <= 1e-4


def test_data_gradient_written_as_two_dense_tensors():
    """The gradient of the level-0 concat buffer ([64 upsampled | 32 skip] channels, torch.cat backward, unet.py:228) as two
    dense tensors: bit-identical to the two channel slices of the single-tensor data gradient."""
    from unet_implementations_b200 import ops
    n, h, w, cin, cout = 2, 24, 40, 96, 32
    g = torch.Generator().manual_seed(51)
    dy, _ = rand_act(n, h, w, cout, seed=52)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    _, wd = ops.pack_conv_weights(wt.cuda())
    dx = ops.conv_dgrad(dy, wd, (h, w), 1)
    d1, d2 = ops.conv_dgrad_split(dy, wd, (h, w), 64)
    assert d1.shape == (n, h, w, 64) and d2.shape == (n, h, w, 32) and d1.is_contiguous() and d2.is_contiguous()
    assert torch.equal(d1, dx[..., :64]) and torch.equal(d2, dx[..., 64:])


def _guarded(shape, dtype=torch.bfloat16, guard=4096, fill=3.0):
    """A tensor of `shape` carved out of the middle of a larger allocation whose margins hold a sentinel value."""
    n = 1
    for d in shape:
        n *= d
    buf = torch.full((n + 2 * guard,), fill, dtype=dtype, device="cuda")
    return buf, buf[guard:guard + n].view(*shape)


def _guards_intact(buf, guard=4096, fill=3.0):
    return bool((buf[:guard].float() == fill).all()) and bool((buf[-guard:].float() == fill).all())


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 9, 190, 32, 32), (1, 7, 70, 32, 32), (1, 13, 21, 64, 64), (1, 5, 130, 96, 32),
                                            (1, 6, 9, 128, 256)])
def test_conv_kernels_stay_inside_ragged_outputs(n, h, w, cin, cout):
    """TMA stores clip at the tensor map's extents; the statistics and weight-gradient epilogues use plain stores.  Ragged
    sizes (tiles hanging over every edge), outputs placed between sentinel margins: nothing outside may change."""
    from unet_implementations_b200 import ops
    x, _ = rand_act(n, h, w, cin, seed=31)
    wt = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    wf, wd = ops.pack_conv_weights(wt)
    ybuf, y = _guarded((n, h, w, cout))
    ops.conv_fprop(x, wf, 1, out=y)
    dy, _ = rand_act(n, h, w, cout, seed=32)
    dbuf, dx = _guarded((n, h, w, cin))
    ops.conv_dgrad(dy, wd, (h, w), 1, out=dx)
    torch.cuda.synchronize()
    assert _guards_intact(ybuf) and _guards_intact(dbuf)
    assert torch.isfinite(y.float()).all() and torch.isfinite(dx.float()).all()
    assert float((y.float() - 3.0).abs().max()) > 0  # the output itself was written


@pytest.mark.parametrize("n,h,w,c", [(2, 13, 10, 64), (1, 70, 66, 32), (1, 64, 8, 128), (3, 1, 1, 8)])
def test_upsample_and_norm_kernels_stay_inside_ragged_outputs(n, h, w, c):
    from unet_implementations_b200 import ops
    x, _ = rand_act(n, h, w, c, seed=33)
    a = torch.rand(n, c, device="cuda") + 0.5
    b = torch.randn(n, c, device="cuda")
    ubuf, up = _guarded((n, 2 * h, 2 * w, c))
    ops.upsample2x(x, up, norm=(a, b, 0.01))
    dbuf, dx = _guarded((n, h, w, c))
    ops.upsample2x_backward(up, out=dx)
    zbuf, z = _guarded((n, h, w, c))
    ops.in_apply(x, a, b, 0.01, out=z)
    torch.cuda.synchronize()
    assert _guards_intact(ubuf) and _guards_intact(dbuf) and _guards_intact(zbuf)
    assert torch.isfinite(up.float()).all() and torch.isfinite(dx.float()).all() and torch.isfinite(z.float()).all()
