"""Operator layer: torch tensors in, C-ABI kernel launches out.

Every function here enqueues work on torch's current CUDA stream through libb200unet.so and returns torch tensors
that own the memory.  Activations are NHWC bf16 tensors of logical shape [N, H, W, C]; a tensor may be a channel
slice of a wider buffer (its W-stride is the "pitch"), which is how the decoder concat buffer is filled in place.
No function falls back to torch math: if the library is missing or the device is not sm_100 the call raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import ConvDgradArgs, ConvFpropArgs, ConvWgradArgs, InBwdArgs

BF16 = torch.bfloat16
F32 = torch.float32


def _sfx(t: torch.Tensor) -> str:
    """Entry-point suffix for the activation storage type: bf16 (production) or fp32 (verification mode)."""
    if t.dtype == BF16:
        return ""
    if t.dtype == F32:
        return "_f32"
    raise TypeError(f"b200unet: activations must be bf16 or fp32, got {t.dtype}")


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pitch_of(t: torch.Tensor) -> int:
    """Pixel pitch (elements) of an NHWC bf16/fp32 tensor or channel-slice view; validates the layout."""
    assert t.dim() == 4 and t.dtype in (BF16, F32) and t.is_cuda, "expected a CUDA bf16/fp32 [N,H,W,C] tensor"
    n, h, w, c = t.shape
    pitch = t.stride(2)
    assert t.stride(3) == 1 and pitch >= c, f"channels must be contiguous (strides {t.stride()})"
    assert (h == 1 or t.stride(1) == w * pitch) and (n == 1 or t.stride(0) == h * w * pitch), (
        f"not a pitched NHWC view: shape {tuple(t.shape)} strides {t.stride()}"
    )
    assert pitch % 8 == 0 and t.data_ptr() % 16 == 0, "pitch must be a multiple of 8 elements, base 16-byte aligned"
    return pitch


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    return t


def require_device():
    if not torch.cuda.is_available():
        raise RuntimeError("b200unet: no CUDA device; this path has no CPU fallback")
    if _lib.call("b200unet_device_ok") != 1:
        raise RuntimeError("b200unet: the current device is not compute capability 10.x (B200, sm_100a)")


# ----------------------------------------------------------------------------------------------------- convolution
def pack_conv_weights(w_oihw: torch.Tensor, need_dgrad: bool = True, dtype=BF16):
    """fp32 [Cout,Cin,3,3] -> ([Cout,3,3,Cin], [Cin,3,3,Cout] or None) in `dtype` (bf16, or fp32 for the fp32 mode)."""
    w = _f32(w_oihw.detach())
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    wf = torch.empty((cout, 3, 3, cin), dtype=dtype, device=w.device)
    wd = torch.empty((cin, 3, 3, cout), dtype=dtype, device=w.device) if need_dgrad else None
    _lib.call("b200unet_pack_conv_weights" + _sfx(wf), _p(w), _p(wf), _p(wd), cout, cin, _stream())
    return wf, wd


def conv_out_hw(h: int, w: int, stride: int):
    return (h - 1) // stride + 1, (w - 1) // stride + 1


def conv_fprop(x, w_fprop, stride=1, out=None, want_stats=True, simt=False):
    """3x3/pad 1 conv.  Returns (y [N,OH,OW,Cout] bf16, stats fp32 [N,P,Cout,2] or None)."""
    n, h, w, cin = x.shape
    cout = w_fprop.shape[0]
    assert w_fprop.shape == (cout, 3, 3, cin) and w_fprop.dtype == x.dtype and w_fprop.is_contiguous()
    oh, ow = conv_out_hw(h, w, stride)
    y = out if out is not None else torch.empty((n, oh, ow, cout), dtype=x.dtype, device=x.device)
    assert tuple(y.shape) == (n, oh, ow, cout) and y.dtype == x.dtype
    f32 = x.dtype == F32
    simt = simt or f32  # the tensor-core path is bf16-only
    stats = None
    if want_stats:
        if simt:
            parts = _lib.call("b200unet_conv_fprop_simt_partials", oh, ow)
        else:
            parts = _lib.call("b200unet_conv_fprop_partials", n, oh, ow, cout)
            if parts < 0:
                raise RuntimeError(f"conv_fprop: Cout={cout} outside the tensor-core envelope")
        stats = torch.empty((n, parts, cout, 2), dtype=torch.float32, device=x.device)
    a = ConvFpropArgs(_p(x), pitch_of(x), _p(w_fprop), _p(y), pitch_of(y), _p(stats), n, h, w, cin, cout, stride)
    name = "b200unet_conv_fprop_f32" if f32 else ("b200unet_conv_fprop_simt" if simt else "b200unet_conv_fprop")
    _lib.call(name, ctypes.byref(a), _stream())
    return y, stats


def pack_s2_dgrad_weights(w_dgrad):
    """bf16 dgrad pack [Cin,3,3,Cout] of a stride-2 conv with Cin in {32, 64} -> the parity-stacked pack
    [4*Cin,4,Cout] of b200unet_conv_dgrad_s2, or None when that path does not apply."""
    cin, _, _, cout = w_dgrad.shape
    if w_dgrad.dtype != BF16 or _lib.call("b200unet_conv_dgrad_s2_supported", cin, cout) != 1:
        return None
    ws = torch.empty((4 * cin, 4, cout), dtype=BF16, device=w_dgrad.device)
    _lib.call("b200unet_pack_s2_dgrad_weights", _p(w_dgrad), _p(ws), cin, cout, _stream())
    return ws


def conv_dgrad_s2(dy, w_stacked, in_hw, out=None):
    """Stride-2 data gradient with the parity classes stacked on N (one launch).  w_stacked from pack_s2_dgrad_weights."""
    n, oh, ow, cout = dy.shape
    cin = w_stacked.shape[0] // 4
    h, w = in_hw
    assert conv_out_hw(h, w, 2) == (oh, ow) and w_stacked.shape == (4 * cin, 4, cout) and dy.dtype == BF16
    dx = out if out is not None else torch.empty((n, h, w, cin), dtype=BF16, device=dy.device)
    a = ConvDgradArgs(_p(dy), pitch_of(dy), _p(w_stacked), _p(dx), pitch_of(dx), n, h, w, cin, cout, 2)
    _lib.call("b200unet_conv_dgrad_s2", ctypes.byref(a), _stream())
    return dx


def conv_dgrad_split(dy, w_dgrad, in_hw, split):
    """Stride-1 data gradient written as TWO dense tensors: channels [0, split) and [split, Cin) -- the gradient of a
    decoder concat buffer without the concat (each half is consumed by a different kernel)."""
    n, oh, ow, cout = dy.shape
    cin = w_dgrad.shape[0]
    h, w = in_hw
    assert w_dgrad.shape == (cin, 3, 3, cout) and w_dgrad.dtype == dy.dtype == BF16 and (oh, ow) == (h, w) and 0 < split < cin
    d1 = torch.empty((n, h, w, split), dtype=BF16, device=dy.device)
    d2 = torch.empty((n, h, w, cin - split), dtype=BF16, device=dy.device)
    a = ConvDgradArgs(_p(dy), pitch_of(dy), _p(w_dgrad), _p(d1), pitch_of(d1), n, h, w, cin, cout, 1)
    a.dx2, a.dx2_pitch, a.dx_split = d2.data_ptr(), pitch_of(d2), split
    _lib.call("b200unet_conv_dgrad", ctypes.byref(a), _stream())
    return d1, d2


def conv_dgrad(dy, w_dgrad, in_hw, stride=1, out=None, simt=False, bwd_sums=None):
    """Gradient wrt the conv input.  dy [N,OH,OW,Cout]; w_dgrad [Cin,3,3,Cout]; returns dx [N,H,W,Cin] bf16.
    bwd_sums = (y, a, b, slope) of the unit whose output feeds this conv (dx is its dz): if the kernel that runs this
    shape supports it, the epilogue also reduces that unit's norm-backward sums and the call returns (dx, part) with
    part fp32 [N,P,Cin,2] for in_backward(ext_part=...); otherwise (dx, None)."""
    n, oh, ow, cout = dy.shape
    cin = w_dgrad.shape[0]
    h, w = in_hw
    assert w_dgrad.shape == (cin, 3, 3, cout) and w_dgrad.dtype == dy.dtype and w_dgrad.is_contiguous()
    assert conv_out_hw(h, w, stride) == (oh, ow)
    dx = out if out is not None else torch.empty((n, h, w, cin), dtype=dy.dtype, device=dy.device)
    a = ConvDgradArgs(_p(dy), pitch_of(dy), _p(w_dgrad), _p(dx), pitch_of(dx), n, h, w, cin, cout, stride)
    name = "b200unet_conv_dgrad_f32" if dy.dtype == F32 else ("b200unet_conv_dgrad_simt" if simt else "b200unet_conv_dgrad")
    part = None
    if bwd_sums is not None and name == "b200unet_conv_dgrad":
        y, sa, sb, slope = bwd_sums
        slots = _lib.call("b200unet_conv_dgrad_bwd_slots", n, h, w, cin, cout, stride)
        if slots > 0 and y.dtype == BF16 and tuple(y.shape) == (n, h, w, cin):
            part = torch.empty((n, slots, cin, 2), dtype=torch.float32, device=dy.device)
            a.bs_y, a.bs_y_pitch, a.bs_a, a.bs_b = y.data_ptr(), pitch_of(y), _f32(sa).data_ptr(), _f32(sb).data_ptr()
            a.bs_slope, a.bs_part = float(slope), part.data_ptr()
    _lib.call(name, ctypes.byref(a), _stream())
    return (dx, part) if bwd_sums is not None else dx


def conv_wgrad(x, dy, stride=1, simt=False, out=None):
    """Weight gradient in the nn.Conv2d layout: fp32 [Cout,Cin,3,3] (written into `out` when given)."""
    n, h, w, cin = x.shape
    _, oh, ow, cout = dy.shape
    assert conv_out_hw(h, w, stride) == (oh, ow)
    if out is not None:
        assert tuple(out.shape) == (cout, cin, 3, 3) and out.dtype == torch.float32 and out.is_contiguous()
        dw = out
    else:
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=x.device)
    assert x.dtype == dy.dtype
    if simt or x.dtype == F32:
        a = ConvWgradArgs(_p(x), pitch_of(x), _p(dy), pitch_of(dy), _p(dw), None, 0, n, h, w, cin, cout, stride)
        _lib.call("b200unet_conv_wgrad_f32" if x.dtype == F32 else "b200unet_conv_wgrad_simt", ctypes.byref(a), _stream())
        return dw
    nbytes = _lib.call("b200unet_conv_wgrad_workspace", n, h, w, cin, cout, stride)
    if nbytes < 0:
        raise RuntimeError(f"conv_wgrad: unsupported shape: {_lib.last_error()}")
    ws = torch.empty((max(nbytes, 4) // 4,), dtype=torch.float32, device=x.device)
    a = ConvWgradArgs(_p(x), pitch_of(x), _p(dy), pitch_of(dy), _p(dw), _p(ws), nbytes, n, h, w, cin, cout, stride)
    _lib.call("b200unet_conv_wgrad", ctypes.byref(a), _stream())
    return dw


def stem_fprop(img_nchw, w_oihw, out=None, want_stats=True):
    """Cin=3 -> 32 stem conv reading the fp32 NCHW image.  Returns (y [N,H,W,32] bf16, stats)."""
    img = _f32(img_nchw)
    n, c, h, w = img.shape
    assert c == 3 and tuple(w_oihw.shape) == (32, 3, 3, 3)
    y = out if out is not None else torch.empty((n, h, w, 32), dtype=BF16, device=img.device)
    stats = None
    if want_stats:
        parts = _lib.call("b200unet_stem_partials", h, w)
        stats = torch.empty((n, parts, 32, 2), dtype=torch.float32, device=img.device)
    _lib.call("b200unet_stem_fprop", _p(img), _p(_f32(w_oihw.detach())), _p(y), pitch_of(y), _p(stats), n, h, w,
              _stream())
    return y, stats


def image_to_nhwc32(img_nchw):
    """fp32 NCHW image (C <= 8) -> bf16 NHWC [N,H,W,32], channels C..31 zero: the stem's operand on the tensor-core
    path (one 64-byte row per pixel)."""
    img = _f32(img_nchw)
    n, c, h, w = img.shape
    assert c <= 8
    xp = torch.empty((n, h, w, 32), dtype=BF16, device=img.device)
    _lib.call("b200unet_image_to_nhwc32_bf16", _p(img), _p(xp), n, c, h * w, _stream())
    return xp


def preprocess_u8_nhwc32(images_u8, mean, std):
    """uint8 HWC batch [N,H,W,3] -> (x / 255 - mean) / std -> bf16 NHWC [N,H,W,32] (channels 3..31 zero) in one kernel:
    the dataset's tensor conversion (train.py:303-308) and the stem's layout change fused."""
    assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.size(3) == 3 and images_u8.is_cuda
    x = images_u8.contiguous()
    n, h, w, _ = x.shape
    xp = torch.empty((n, h, w, 32), dtype=BF16, device=x.device)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    _lib.call("b200unet_preprocess_u8_nhwc32", _p(x), _p(xp), m3, s3, n, h * w, _stream())
    return xp


def pack_stem_weights(w_oihw):
    """[Cout, C<=8, 3, 3] fp32 -> bf16 [Cout,3,3,32] with the input channels zero-padded to 32."""
    cout, c = w_oihw.shape[0], w_oihw.shape[1]
    w32 = torch.zeros((cout, 32, 3, 3), dtype=torch.float32, device=w_oihw.device)
    w32[:, :c] = w_oihw.detach()
    return pack_conv_weights(w32, need_dgrad=False)[0]


def stem_wgrad_tc(img_nchw, dy, xp=None, out=None, channels=None):
    """Stem weight gradient on the tensor-core path: the narrow-output wgrad kernel (Cin = 32 -> Cout = 32) runs on the
    zero-padded bf16 image and the 29 zero rows are dropped.  2.5x faster than the CUDA-core kernel at 512^2 x 32."""
    c = channels if channels is not None else img_nchw.shape[1]
    assert dy.dtype == BF16
    if xp is None:
        xp = image_to_nhwc32(img_nchw)
    full = conv_wgrad(xp, dy, 1)
    if out is not None:
        out.copy_(full[:, :c])
        return out
    return full[:, :c].contiguous()


def stem_wgrad(img_nchw, dy):
    img = _f32(img_nchw)
    n, c, h, w = img.shape
    nbytes = _lib.call("b200unet_stem_wgrad_workspace", n, h, w)
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=img.device)
    dw = torch.empty((32, 3, 3, 3), dtype=torch.float32, device=img.device)
    _lib.call("b200unet_stem_wgrad", _p(img), _p(dy), pitch_of(dy), _p(dw), _p(ws), nbytes, n, h, w, _stream())
    return dw


# ------------------------------------------------------------------------------------- InstanceNorm + LeakyReLU + drop
def in_finalize(stats, gamma, beta, drop_scale, eps, hw):
    """stats [N,P,C,2] -> (mean, rstd, a, b), each fp32 [N,C]."""
    n, parts, c, _ = stats.shape
    out = torch.empty((4, n, c), dtype=torch.float32, device=stats.device)
    mean, rstd, a, b = out[0], out[1], out[2], out[3]
    _lib.call("b200unet_in_finalize", _p(stats), parts, _p(_f32(gamma.detach())), _p(_f32(beta.detach())),
              _p(drop_scale), float(eps), _p(mean), _p(rstd), _p(a), _p(b), n, c, hw, _stream())
    return mean, rstd, a, b


def in_apply(y, a, b, slope, out=None):
    n, h, w, c = y.shape
    z = out if out is not None else torch.empty((n, h, w, c), dtype=y.dtype, device=y.device)
    assert tuple(z.shape) == (n, h, w, c) and z.dtype == y.dtype
    _lib.call("b200unet_in_apply" + _sfx(y), _p(y), pitch_of(y), _p(a), _p(b), float(slope), _p(z), pitch_of(z), n, h * w, c,
              _stream())
    return z


def in_backward(dz, dz2, y, a, b, mean, rstd, drop_scale, gamma, slope, out_dgamma=None, out_dbeta=None,
                ext_part=None, ext_part2=None, defer_params=False):
    """Backward of z = lrelu(IN(y))*drop.  dz2 (optional) is a second gradient contribution added to dz.
    Returns (dy bf16 [N,H,W,C], dgamma [C], dbeta [C]); the parameter gradients go to out_dgamma / out_dbeta if given.
    ext_part / ext_part2: fp32 [N,P,C,2] partial sums (sum gm, sum gm*y) already produced by the kernel that wrote dz /
    dz2 -- the reduction pass over (dz, y) is then skipped.
    defer_params: returns (dy, finish) instead; finish() enqueues the dgamma / dbeta reduction on the then-current
    stream and returns (dgamma, dbeta) -- it is off the critical path of backward."""
    n, h, w, c = y.shape
    hw = h * w
    nbytes = _lib.call("b200unet_in_backward_workspace", n, hw, c)
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=y.device)
    if out_dgamma is not None and out_dbeta is not None:
        assert out_dgamma.numel() == c and out_dbeta.numel() == c and out_dgamma.dtype == torch.float32
        dgb = (out_dgamma, out_dbeta)
    else:
        dgb = torch.empty((2, c), dtype=torch.float32, device=y.device)
    dy = torch.empty((n, h, w, c), dtype=y.dtype, device=y.device)
    assert dz.dtype == y.dtype and (dz2 is None or dz2.dtype == y.dtype)
    ep, ep2 = ext_part, ext_part2
    if ep is not None:
        assert ep.dtype == torch.float32 and ep.is_contiguous() and ep.shape[0] == n and tuple(ep.shape[2:]) == (c, 2)
    if ep2 is not None:
        assert ep2.dtype == torch.float32 and ep2.is_contiguous() and ep2.shape[0] == n and tuple(ep2.shape[2:]) == (c, 2)
    args = InBwdArgs(_p(dz), pitch_of(dz), _p(dz2), pitch_of(dz2) if dz2 is not None else 0, _p(y), pitch_of(y), _p(a),
                     _p(b), _p(mean), _p(rstd), _p(drop_scale), _p(_f32(gamma.detach())), float(slope), _p(dy),
                     pitch_of(dy), _p(dgb[0]), _p(dgb[1]), _p(ws), nbytes, n, hw, c,
                     _p(ep), ep.shape[1] if ep is not None else 0, _p(ep2), ep2.shape[1] if ep2 is not None else 0,
                     1 if defer_params else 0)
    _lib.call("b200unet_in_backward" + _sfx(y), ctypes.byref(args), _stream())
    if defer_params:
        # dgamma / dbeta are NOT computed yet: call the returned function (on any stream ordered after this one)
        def finish_params():
            _lib.call("b200unet_in_bwd_params", _p(ws), n, hw, c, _p(dgb[0]), _p(dgb[1]), _stream())
            return dgb[0], dgb[1]
        finish_params.keep = (ws, dgb)  # what the deferred kernel reads / writes: the caller keeps it alive until it has run
        return dy, finish_params
    return dy, dgb[0], dgb[1]


# ------------------------------------------------------------------------------------------------------- resampling
def upsample2x(x, out, norm=None):
    """Bilinear 2x of x [N,H,W,C] into out [N,2H,2W,C] (typically the leading channel slice of a concat buffer).
    norm = (a, b, slope): x is a RAW conv output and leaky_relu(a*x + b) is applied on the fly (the producer's apply
    pass fused into this, its only, consumer)."""
    n, h, w, c = x.shape
    assert tuple(out.shape) == (n, 2 * h, 2 * w, c) and out.dtype == x.dtype
    if norm is None:
        _lib.call("b200unet_upsample2x_fwd" + _sfx(x), _p(x), pitch_of(x), _p(out), pitch_of(out), n, h, w, c, _stream())
    else:
        a, b, slope = norm
        _lib.call("b200unet_upsample2x_norm_fwd" + _sfx(x), _p(x), pitch_of(x), _p(a), _p(b), float(slope), _p(out),
                  pitch_of(out), n, h, w, c, _stream())
    return out


def upsample2x_backward(dout, out=None):
    n, oh, ow, c = dout.shape
    assert oh % 2 == 0 and ow % 2 == 0
    dx = out if out is not None else torch.empty((n, oh // 2, ow // 2, c), dtype=dout.dtype, device=dout.device)
    _lib.call("b200unet_upsample2x_bwd" + _sfx(dout), _p(dout), pitch_of(dout), _p(dx), pitch_of(dx), n, oh // 2, ow // 2, c,
              _stream())
    return dx


def nchw_to_nhwc(x_nchw, out=None):
    x = _f32(x_nchw)
    n, c, h, w = x.shape
    y = out if out is not None else torch.empty((n, h, w, c), dtype=BF16, device=x.device)
    name = "b200unet_nchw_f32_to_nhwc_f32" if y.dtype == F32 else "b200unet_nchw_f32_to_nhwc_bf16"
    _lib.call(name, _p(x), _p(y), y.stride(2), n, c, h * w, _stream())
    return y


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _lib.call("b200unet_nhwc_bf16_to_nchw_f32", _p(x), x.stride(2), _p(y), n, c, h * w, _stream())
    return y


# ------------------------------------------------------------------------------------------------------ head + loss
def head_forward(z, weight, bias, norm=None):
    """1x1 conv C->K on NHWC z; returns fp32 NCHW logits.  norm = (a, b, slope): z is the RAW conv output of the last
    unit and its InstanceNorm/LeakyReLU/dropout apply is fused in."""
    n, h, w, c = z.shape
    k = weight.shape[0]
    logits = torch.empty((n, k, h, w), dtype=torch.float32, device=z.device)
    if norm is None:
        _lib.call("b200unet_head_fwd" + _sfx(z), _p(z), pitch_of(z), _p(_f32(weight.detach().reshape(k, c))),
                  _p(_f32(bias.detach())), _p(logits), n, h * w, c, k, _stream())
    else:
        a, b, slope = norm
        _lib.call("b200unet_head_norm_fwd" + _sfx(z), _p(z), pitch_of(z), _p(a), _p(b), float(slope),
                  _p(_f32(weight.detach().reshape(k, c))), _p(_f32(bias.detach())), _p(logits), n, h * w, c, k, _stream())
    return logits


def head_backward(dlogits, z, weight, norm=None, out_dw=None, out_db=None, want_bwd_part=False):
    """Returns (dz NHWC, dW [K,C,1,1], db [K]).  norm as in head_forward (z recomputed from the raw conv output).
    want_bwd_part (with norm): also returns the producing unit's norm-backward partial sums [N,P,C,2] as a 4th value."""
    n, h, w, c = z.shape
    k = weight.shape[0]
    dl = _f32(dlogits.contiguous())
    nbytes = _lib.call("b200unet_head_bwd_workspace", n, h * w, c, k)
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=z.device)
    dz = torch.empty((n, h, w, c), dtype=z.dtype, device=z.device)
    dw = out_dw if out_dw is not None else torch.empty((k, c, 1, 1), dtype=torch.float32, device=z.device)
    db = out_db if out_db is not None else torch.empty((k,), dtype=torch.float32, device=z.device)
    assert dw.numel() == k * c and db.numel() == k and dw.is_contiguous() and db.is_contiguous()
    if norm is None:
        _lib.call("b200unet_head_bwd" + _sfx(z), _p(dl), _p(z), pitch_of(z), _p(_f32(weight.detach().reshape(k, c))), _p(dz),
                  pitch_of(dz), _p(dw), _p(db), _p(ws), nbytes, n, h * w, c, k, _stream())
    elif want_bwd_part:
        a, b, slope = norm
        slots = _lib.call("b200unet_head_bwd_stat_slots", n, h * w)
        part = torch.empty((n, slots, c, 2), dtype=torch.float32, device=z.device)
        _lib.call("b200unet_head_norm_bwd_stats" + _sfx(z), _p(dl), _p(z), pitch_of(z), _p(a), _p(b), float(slope),
                  _p(_f32(weight.detach().reshape(k, c))), _p(dz), pitch_of(dz), _p(dw), _p(db), _p(ws), nbytes, _p(part),
                  n, h * w, c, k, _stream())
        return dz, dw, db, part
    else:
        a, b, slope = norm
        _lib.call("b200unet_head_norm_bwd" + _sfx(z), _p(dl), _p(z), pitch_of(z), _p(a), _p(b), float(slope),
                  _p(_f32(weight.detach().reshape(k, c))), _p(dz), pitch_of(dz), _p(dw), _p(db), _p(ws), nbytes, n, h * w, c,
                  k, _stream())
    return dz, dw, db


def loss_forward(logits, target, class_weights, dynamic, weight_ce, weight_dice, ignore_index, smooth):
    """Returns (loss_out fp32 [3] = total/CE/Dice, tables) -- see b200unet_loss_fwd."""
    lg = _f32(logits)
    n, k, h, w = lg.shape
    assert k == 3, "SimpleLoss kernels are built for 3 classes (losses.py:40)"
    assert target.dtype in (torch.int64, torch.uint8) and target.is_cuda and target.is_contiguous()
    assert tuple(target.shape) == (n, h, w)
    hw = h * w
    nbytes = _lib.call("b200unet_loss_workspace", n, hw)
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=lg.device)
    out = torch.empty((3,), dtype=torch.float32, device=lg.device)
    tables = torch.empty((3 + 6 * n,), dtype=torch.float32, device=lg.device)
    cw = _f32(class_weights) if class_weights is not None else None
    _lib.call("b200unet_loss_fwd" + ("_u8" if target.dtype == torch.uint8 else ""), _p(lg), _p(target), _p(cw),
              int(bool(dynamic)), float(weight_ce),
              float(weight_dice), int(ignore_index), float(smooth), _p(out), _p(tables), _p(ws), nbytes, n, hw,
              _stream())
    return out, tables


def loss_backward(logits, target, tables, grad_out, weight_ce, weight_dice, ignore_index):
    lg = _f32(logits)
    n, k, h, w = lg.shape
    dl = torch.empty_like(lg)
    go = _f32(grad_out.reshape(1).contiguous()) if grad_out is not None else None
    _lib.call("b200unet_loss_bwd" + ("_u8" if target.dtype == torch.uint8 else ""), _p(lg), _p(target), _p(tables), _p(go),
              float(weight_ce), float(weight_dice),
              int(ignore_index), _p(dl), n, h * w, _stream())
    return dl


# --------------------------------------------------------------------------------- autoencoder head + MSE (configs[3])
def recon_head_forward(y, bias):
    """Epilogue of the reconstruction conv: y [N,H,W,>=K] (raw 3x3 conv output, NHWC) + bias -> sigmoid -> fp32 NCHW."""
    n, h, w, _ = y.shape
    k = bias.numel()
    out = torch.empty((n, k, h, w), dtype=torch.float32, device=y.device)
    _lib.call("b200unet_recon_head_fwd" + _sfx(y), _p(y), pitch_of(y), _p(_f32(bias.detach())), _p(out), n, h * w, k,
              _stream())
    return out


def recon_head_backward(dout, out, cpad, dtype):
    """dpre = dout * out * (1 - out) as an NHWC [N,H,W,cpad] tensor of `dtype` (channels >= K zero) and db [K]."""
    do, o = _f32(dout.contiguous()), _f32(out)
    n, k, h, w = o.shape
    dpre = torch.empty((n, h, w, cpad), dtype=dtype, device=o.device)
    db = torch.empty((k,), dtype=torch.float32, device=o.device)
    nbytes = _lib.call("b200unet_recon_head_bwd_workspace", n, h * w)
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=o.device)
    _lib.call("b200unet_recon_head_bwd" + _sfx(dpre), _p(do), _p(o), _p(dpre), pitch_of(dpre), cpad, _p(db), _p(ws), nbytes,
              n, h * w, k, _stream())
    return dpre, db


def mse_forward(a, b):
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    n = a.numel()
    nbytes = _lib.call("b200unet_mse_workspace", n)
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=a.device)
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    _lib.call("b200unet_mse_fwd", _p(a), _p(b), _p(out), _p(ws), nbytes, n, _stream())
    return out


def mse_backward(a, b, grad_out):
    a, b = _f32(a), _f32(b)
    da = torch.empty_like(a)
    go = _f32(grad_out.reshape(1).contiguous()) if grad_out is not None else None
    _lib.call("b200unet_mse_bwd", _p(a), _p(b), _p(go), _p(da), a.numel(), _stream())
    return da

