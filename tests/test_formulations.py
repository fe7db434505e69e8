"""CPU restatements (float64) of the algebra the CUDA kernels rely on, each against torch autograd of the reference op
it replaces -- the formulations are not obvious from the reference, so they are pinned where no GPU is needed:
  * dropout folded into the InstanceNorm affine and the closed-form norm backward (csrc/norm.cu; unet.py:13-35, :118-127);
  * the column taps stacked on N with a +-1 pixel recombination (csrc/conv_narrow.cu);
  * the stride-2 data gradient as four parity classes stacked on N over a 2x2 shift set, with the zero-block table of
    pack_s2_dgrad_weights_kernel (csrc/conv_fprop_dgrad.cu; unet.py:103 stride rule);
  * the weight gradient with the row tap on the input and the column tap on the output gradient (csrc/conv_wgrad_narrow.cu)."""
import torch
import torch.nn.functional as F

D = torch.float64


def test_folded_affine_and_closed_form_norm_backward():
    g = torch.Generator().manual_seed(0)
    N, C, H, W, slope, p, eps = 3, 6, 5, 7, 0.01, 0.3, 1e-5
    y = torch.randn(N, C, H, W, generator=g, dtype=D).requires_grad_(True)
    gamma = (torch.rand(C, generator=g, dtype=D) + 0.5).requires_grad_(True)
    beta = torch.randn(C, generator=g, dtype=D).requires_grad_(True)
    s = (torch.rand(N, C, generator=g) > p).to(D) / (1 - p)          # SpatialDropout2d scale per (n, c), >= 0
    z_ref = F.leaky_relu(F.instance_norm(y, weight=gamma, bias=beta, eps=eps), slope) * s[:, :, None, None]
    mean = y.detach().mean(dim=(2, 3))
    rstd = 1.0 / torch.sqrt(y.detach().var(dim=(2, 3), unbiased=False) + eps)
    a = s * gamma.detach() * rstd                                      # forward: z = lrelu(a * y + b)
    b = s * (beta.detach() - mean * gamma.detach() * rstd)
    z = F.leaky_relu(a[:, :, None, None] * y.detach() + b[:, :, None, None], slope)
    assert torch.allclose(z, z_ref.detach(), atol=1e-12)              # positively homogeneous activation, s >= 0
    dz = torch.randn(N, C, H, W, generator=g, dtype=D)
    z_ref.backward(dz)
    # backward as the kernels do it.  The reduce pass masks with the sign of a*y + b, which for a dropped channel
    # (a = b = 0) differs from the sign of the pre-activation -- harmless, because every use is multiplied by s = 0
    pre = a[:, :, None, None] * y.detach() + b[:, :, None, None]
    gm = torch.where(pre > 0, dz, dz * slope)
    yc = y.detach() - mean[:, :, None, None]
    T1, T2 = gm.sum(dim=(2, 3)), (gm * yc).sum(dim=(2, 3))
    S1, S2 = s * T1, s * rstd * T2                                     # sum g, sum g * xhat with g = s * gm
    gr, hw = gamma.detach() * rstd, H * W
    k1, k2, k3 = gr * s, gr * rstd * S2 / hw, gr * S1 / hw
    dy = k1[:, :, None, None] * gm - k2[:, :, None, None] * yc - k3[:, :, None, None]
    assert torch.allclose(dy, y.grad, atol=1e-10)
    assert torch.allclose(S2.sum(0), gamma.grad, atol=1e-10) and torch.allclose(S1.sum(0), beta.grad, atol=1e-10)


def test_column_taps_stacked_on_n():
    g = torch.Generator().manual_seed(1)
    H, W, C, Co = 6, 9, 4, 5
    x = torch.randn(H, W, C, generator=g, dtype=D)
    w = torch.randn(Co, C, 3, 3, generator=g, dtype=D)
    ref = F.conv2d(x.permute(2, 0, 1)[None], w, padding=1)[0].permute(1, 2, 0)
    xp = F.pad(x, (0, 0, 0, 0, 1, 1))
    E = sum(torch.einsum("hwc,okc->hwok", xp[kh:kh + H], w[:, :, kh, :].permute(0, 2, 1)) for kh in range(3))  # [H,W,Co,kw]
    out = E[..., 1] + F.pad(E[:, :-1, :, 0], (0, 0, 1, 0)) + F.pad(E[:, 1:, :, 2], (0, 0, 0, 1))
    assert torch.allclose(out, ref, atol=1e-12)


def test_stride2_data_gradient_as_stacked_parity_classes():
    g = torch.Generator().manual_seed(2)
    for H, W in ((8, 10), (7, 9)):
        Ci, Co = 3, 4
        x = torch.randn(1, Ci, H, W, generator=g, dtype=D).requires_grad_(True)
        w = torch.randn(Co, Ci, 3, 3, generator=g, dtype=D)
        y = F.conv2d(x, w, padding=1, stride=2)
        dy = torch.randn_like(y)
        y.backward(dy)
        OH, OW = y.shape[2:]
        dyp = F.pad(dy[0], (0, 1, 0, 1))                                # dy[a + dh, b + dw], zero past the edge
        dx = torch.zeros(Ci, H, W, dtype=D)
        for ph in range(2):
            for pw in range(2):
                acc = torch.zeros(Ci, OH, OW, dtype=D)
                for dh in range(2):
                    for dw in range(2):                                 # the table of pack_s2_dgrad_weights_kernel
                        kh = (1 if dh == 0 else -1) if ph == 0 else (2 if dh == 0 else 0)
                        kw = (1 if dw == 0 else -1) if pw == 0 else (2 if dw == 0 else 0)
                        if kh < 0 or kw < 0:
                            continue                                    # zero block: 7 of the 16 are
                        acc += torch.einsum("oab,oi->iab", dyp[:, dh:dh + OH, dw:dw + OW], w[:, :, kh, kw])
                hs, ws = (H - ph + 1) // 2, (W - pw + 1) // 2           # this class's sub-lattice of the input
                dx[:, ph::2, pw::2] = acc[:, :hs, :ws]
        assert torch.allclose(dx, x.grad[0], atol=1e-12)


def test_weight_gradient_with_row_tap_on_x_and_column_tap_on_dy():
    g = torch.Generator().manual_seed(3)
    H, W, C, Co = 5, 8, 3, 4
    x = torch.randn(1, C, H, W, generator=g, dtype=D)
    w = torch.randn(Co, C, 3, 3, generator=g, dtype=D, requires_grad=True)
    y = F.conv2d(x, w, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    xp = F.pad(x[0], (0, 0, 1, 1))                                      # rows -1 .. H
    dw = torch.zeros(Co, C, 3, 3, dtype=D)
    for kh in range(3):
        for kw in range(3):
            # dW[co, ci, kh, kw] = sum X[oh + kh - 1, w', ci] * dY[oh, w' - kw + 1, co]: the shift kw - 1 sits on dY
            dys = torch.roll(F.pad(dy[0], (1, 1)), shifts=kw - 1, dims=2)[:, :, 1:W + 1]
            dw[:, :, kh, kw] = torch.einsum("chw,ohw->oc", xp[:, kh:kh + H], dys)
    assert torch.allclose(dw, w.grad, atol=1e-12)


def test_bilinear_2x_taps_and_gather_backward():
    """csrc/resample.cu: out[2i] = .25 in[max(i-1,0)] + .75 in[i], out[2i+1] = .75 in[i] + .25 in[min(i+1,n-1)] per axis
    (F.interpolate(mode='bilinear', align_corners=False) at an exact factor 2, unet.py:220-225); backward in gather form
    din[i] = .75 (d[2i] + d[2i+1]) + .25 (d[2i-1] + d[2i+2]) where a tap that falls off the output did not exist --
    its weight went to the edge sample instead: din[0] gets .25 d[0] more, din[n-1] gets .25 d[2n-1] more."""
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 2, 5, 6, generator=g, dtype=D).requires_grad_(True)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)

    def up_axis(t, dim):
        n = t.shape[dim]
        idx = torch.arange(n)
        lo, hi = t.index_select(dim, (idx - 1).clamp(min=0)), t.index_select(dim, (idx + 1).clamp(max=n - 1))
        even, odd = 0.25 * lo + 0.75 * t, 0.75 * t + 0.25 * hi
        return torch.stack([even, odd], dim=dim + 1).flatten(dim, dim + 1)

    assert torch.allclose(up_axis(up_axis(x.detach(), 3), 2), ref.detach(), atol=1e-12)   # along w first, then h
    d = torch.randn_like(ref)
    ref.backward(d)

    def down_axis(t, dim):
        m = t.shape[dim]
        n = m // 2
        i = torch.arange(n)
        take = lambda j: t.index_select(dim, j.clamp(0, m - 1))  # noqa: E731  (clamped index = the edge sample)
        return 0.75 * (take(2 * i) + take(2 * i + 1)) + 0.25 * (take(2 * i - 1) + take(2 * i + 2))

    assert torch.allclose(down_axis(down_axis(d, 2), 3), x.grad, atol=1e-12)
