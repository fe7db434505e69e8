"""CPU restatement of the pixel-pair formulation of unet-implementations_b200/csrc/conv_pair.cu (the 32 -> 32 convs at
512^2): two horizontally adjacent pixels form one K = 64 operand row, four products per pair are stacked on N = 128,
and the outputs are recombined with one neighbour exchange.  Checked against F.conv2d and its autograd (nn.Conv2d of
Our_UNet/models/unet.py:106-115), so that the tap table of the kernel (`pc_tap`) and the flipped reading used for the
data gradient are pinned by a test that needs no GPU."""
import torch
import torch.nn.functional as F

# column tap of the (effective) kernel that product j takes from the pixel of parity pi; None = zero block
PC_TAP = {0: (1, 2), 1: (0, 1), 2: (None, 0), 3: (2, None)}


def pair_conv(x, w_eff):
    """x [H, W, C] (W even), w_eff [Cout, C, 3, 3] -> y [H, W, Cout] through the pair formulation (zero padding)."""
    H, W, C = x.shape
    Co = w_eff.shape[0]
    xp = F.pad(x, (0, 0, 0, 0, 1, 1))                      # rows -1 .. H
    pairs = xp.reshape(H + 2, W // 2, 2 * C)               # K = (parity, ci)
    acc = torch.zeros(H, W // 2, 4, Co, dtype=x.dtype)     # c_j[P]
    for kh in range(3):
        B = torch.zeros(4, Co, 2 * C, dtype=x.dtype)       # [(j, co), (parity, ci)] of this kh: 6 of 8 blocks non-zero
        for j, taps in PC_TAP.items():
            for pi, kw in enumerate(taps):
                if kw is not None:
                    B[j, :, pi * C:(pi + 1) * C] = w_eff[:, :, kh, kw]
        acc += torch.einsum("hpk,jok->hpjo", pairs[kh:kh + H], B)
    y0 = acc[:, :, 0] + F.pad(acc[:, :-1, 2], (0, 0, 1, 0))   # c0[P] + c2[P-1]
    y1 = acc[:, :, 1] + F.pad(acc[:, 1:, 3], (0, 0, 0, 1))    # c1[P] + c3[P+1]
    return torch.stack([y0, y1], dim=2).reshape(H, W, Co)


def test_pair_formulation_equals_conv2d_and_its_data_gradient():
    g = torch.Generator().manual_seed(0)
    H, W, C = 7, 12, 5
    x = torch.randn(H, W, C, generator=g, dtype=torch.float64)
    w = torch.randn(C, C, 3, 3, generator=g, dtype=torch.float64)
    xr = x.permute(2, 0, 1)[None].clone().requires_grad_(True)
    ref = F.conv2d(xr, w, padding=1)
    assert torch.allclose(pair_conv(x, w), ref[0].permute(1, 2, 0), atol=1e-12)
    # data gradient: the same computation on dy with the kernel flipped in both directions and its channels swapped
    # (kh -> 2 - kh, kw -> 2 - kw: what the REV instantiation reads from the [Cin][3][3][Cout] packing)
    dy = torch.randn(H, W, C, generator=g, dtype=torch.float64)
    ref.backward(dy.permute(2, 0, 1)[None])
    w_rev = w.flip(2, 3).transpose(0, 1).contiguous()
    assert torch.allclose(pair_conv(dy, w_rev), xr.grad[0].permute(1, 2, 0), atol=1e-12)


def test_pair_weight_tiles_are_three_quarters_dense():
    blocks = [kw is not None for taps in PC_TAP.values() for kw in taps]
    assert sum(blocks) == 6 and len(blocks) == 8  # the same 75 % of useful MACs as the N = 96 stacking
