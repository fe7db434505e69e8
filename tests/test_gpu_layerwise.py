"""GPU: teacher-forced per-layer gate of the BENCHMARKED bf16 path at full depth and full resolution.

north_star asks for "gradients within 1e-2 relative in bf16".  End to end that bound is unattainable for any bf16
pipeline of this network (at random init it amplifies rounding noise by ~1e5: the reference's own autocast(bf16) run
is 0.5 rel-L2 away from its fp32 run, profiles/r1_parity.md), so an end-to-end gate can only be loose -- loose enough
to hide a 30 % error in one layer.  This test closes that hole where it can be closed: ONE training step of the
default `UNet()` at 512 x 512 (batch 4), the production kernels and dispatch (tcgen05 convs, fused consumers, side-
stream weight gradients), with every intermediate tensor of the forward AND the backward recorded; then EVERY operator
instance -- 22 x (conv fprop, InstanceNorm statistics + folded affine, apply, norm backward, dgrad, wgrad), 5 x
(upsample forward / backward), head forward / backward -- is recomputed by the reference's own op
(`F.conv2d`, `F.instance_norm`, `F.leaky_relu`, `F.interpolate`, their autograd; Our_UNet/models/unet.py:106-127,
:203-231, :374-381) in fp32 on the same GPU (TF32 off) FROM THE SAME bf16 INPUTS the kernel saw.  Tolerances: one bf16
rounding of the output (rel-L2 <= 4e-3) for bf16 tensors, 1e-3 for fp32 parameter gradients, 1e-4 for fp32 statistics.
torch's GPU ops are test infrastructure here (an independent oracle at full size); the product path never calls them.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL_BF16 = 4e-3   # one round-to-nearest bf16 of a correctly computed value: 2^-9 / sqrt(3) = 1.1e-3 rel-L2
TOL_DW = 1e-3     # fp32 weight gradients (fp32 accumulation over up to 1M products, different summation orders)
TOL_STAT = 1e-4


def rel(got, ref):
    got, ref = got.double(), ref.double()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def act_of(y_nhwc, norm):
    """leaky_relu(a * y + b) in fp32 NCHW: the activation a fused consumer applies on the fly."""
    if norm is None:
        return nchw(y_nhwc)
    a, b, slope = norm
    return F.leaky_relu(nchw(y_nhwc) * a[:, :, None, None] + b[:, :, None, None], slope)


@pytest.fixture(scope="module")
def traced_step():
    from oracle import unet_oracle as O
    from unet_implementations_b200.models.losses import SimpleLoss
    from unet_implementations_b200.models.unet import UNet
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234)
    model = UNet().cuda().train()
    x, target = O.synthetic_batch(4, 512, seed=0)
    model._trace_fwd, model._trace_bwd = [], []
    torch.manual_seed(99)
    logits = model(x.cuda())
    loss = SimpleLoss()(logits, target.cuda())
    loss.backward()
    torch.cuda.synchronize()
    fwd, bwd = model._trace_fwd, model._trace_bwd
    model._trace_fwd = model._trace_bwd = None
    yield model, fwd, bwd, logits
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_every_forward_operator_against_the_reference_op(traced_step):
    model, fwd, _, logits = traced_step
    units = [r for r in fwd if r["kind"] == "unit"]
    assert len(units) == 22 and sum(r["kind"] == "up" for r in fwd) == 5 and fwd[-1]["kind"] == "head"
    worst = {}
    for r in fwd:
        if r["kind"] == "unit":
            conv, norm = r["conv"], r["norm"]
            xin = nchw(r["xin"])
            w = conv.weight.detach().bfloat16().float()
            if xin.shape[1] != w.shape[1]:  # the stem's operand is zero-padded to 32 channels
                assert float(xin[:, w.shape[1]:].abs().max()) == 0.0
                xin = xin[:, :w.shape[1]]
            y_ref = F.conv2d(xin, w, None, r["stride"], 1)
            e = rel(nchw(r["y"]), y_ref)
            worst["fprop"] = max(worst.get("fprop", 0), e)
            assert e <= TOL_BF16, f"unit {r['li']} fprop {tuple(w.shape)}: {e:.2e}"
            # statistics of the STORED bf16 output (what InstanceNorm then normalises), folded affine
            y = nchw(r["y"]).double()
            mean = y.mean((2, 3))
            var = y.var((2, 3), unbiased=False)
            rstd = (var + norm.eps).rsqrt()
            assert rel(r["mean"], mean) <= TOL_STAT or float((r["mean"].double() - mean).abs().max()) <= 1e-5 * float(y.abs().max())
            assert rel(r["rstd"], rstd) <= TOL_STAT, f"unit {r['li']} rstd"
            s = r["scale"].double() if r["scale"] is not None else torch.ones_like(mean)
            g, b = norm.weight.detach().double(), norm.bias.detach().double()
            assert rel(r["a"], s * g * rstd) <= TOL_STAT
            b_ref = s * (b - mean * g * rstd)
            assert float((r["b"].double() - b_ref).abs().max()) <= 1e-4 * float(b_ref.abs().max().clamp_min(1.0))
            if r["z"] is not None:
                # the reference's module sequence on the same y: IN -> LeakyReLU -> channel dropout (unet.py:118-127)
                z_ref = F.leaky_relu(F.instance_norm(y.float(), weight=norm.weight.detach(), bias=norm.bias.detach(), eps=norm.eps),
                                     r["slope"])
                if r["scale"] is not None:
                    z_ref = z_ref * r["scale"][:, :, None, None]
                e = rel(nchw(r["z"]), z_ref)
                worst["apply"] = max(worst.get("apply", 0), e)
                assert e <= TOL_BF16, f"unit {r['li']} norm+lrelu+dropout: {e:.2e}"
        elif r["kind"] == "up":
            src = act_of(r["src"], r["norm"])
            ref = F.interpolate(src, scale_factor=2, mode="bilinear", align_corners=False)
            e = rel(nchw(r["out"]), ref)
            worst["upsample"] = max(worst.get("upsample", 0), e)
            assert e <= TOL_BF16, f"upsample {tuple(src.shape)}: {e:.2e}"
        else:
            head = model.segmentation_output
            z = act_of(r["z"], r["norm"])
            ref = F.conv2d(z, head.weight.detach(), head.bias.detach())
            e = rel(r["logits"], ref)
            worst["head"] = max(worst.get("head", 0), e)
            assert e <= 1e-5, f"head: {e:.2e}"
    print("layerwise forward worst rel-L2:", {k: f"{v:.2e}" for k, v in worst.items()})


def test_every_backward_operator_against_the_reference_autograd(traced_step):
    model, fwd, bwd, _ = traced_step
    funits = {r["li"]: r for r in fwd if r["kind"] == "unit"}
    bunits = [r for r in bwd if r["kind"] == "unit"]
    assert len(bunits) == 22 and sum(r["kind"] == "up" for r in bwd) == 5 and bwd[0]["kind"] == "head"
    worst = {}

    def note(k, e):
        worst[k] = max(worst.get(k, 0), e)

    for r in bwd:
        if r["kind"] == "head":
            head = model.segmentation_output
            hf = fwd[-1]
            z = act_of(hf["z"], hf["norm"]).requires_grad_(True)
            w = head.weight.detach().clone().requires_grad_(True)
            b = head.bias.detach().clone().requires_grad_(True)
            F.conv2d(z, w, b).backward(r["dlogits"])
            note("head dz", rel(nchw(r["dz"]), z.grad))
            assert rel(nchw(r["dz"]), z.grad) <= TOL_BF16
            assert rel(r["dw"], w.grad) <= TOL_DW and rel(r["db"], b.grad) <= TOL_DW
        elif r["kind"] == "up":
            dout = nchw(r["dout"])
            x0 = torch.zeros(dout.shape[0], dout.shape[1], dout.shape[2] // 2, dout.shape[3] // 2, device=dout.device,
                             requires_grad=True)
            F.interpolate(x0, scale_factor=2, mode="bilinear", align_corners=False).backward(dout)
            e = rel(nchw(r["dx"]), x0.grad)
            note("upsample bwd", e)
            assert e <= TOL_BF16, f"upsample backward: {e:.2e}"
        else:
            f = funits[r["li"]]
            conv, norm = f["conv"], f["norm"]
            # ---- norm + LeakyReLU + dropout backward (native_batch_norm_backward + leaky_relu_backward + mul)
            y = nchw(r["y"]).requires_grad_(True)
            gam = norm.weight.detach().clone().requires_grad_(True)
            bet = norm.bias.detach().clone().requires_grad_(True)
            pre = F.instance_norm(y, weight=gam, bias=bet, eps=norm.eps)
            z = F.leaky_relu(pre, f["slope"])
            if f["scale"] is not None:
                z = z * f["scale"][:, :, None, None]
            dz = nchw(r["dz"])
            if r["dz2"] is not None:
                dz = dz + nchw(r["dz2"])
            z.backward(dz)
            e = rel(nchw(r["dy"]), y.grad)
            note("norm bwd dy", e)
            assert e <= TOL_BF16, f"unit {r['li']} norm backward dy: {e:.2e}"
            # LeakyReLU's derivative jumps at 0: a pixel whose pre-activation is 0 to rounding (the bf16 y sits on the
            # plane mean) may take either slope in any fp32 evaluation -- the reference's included.  Such knife-edge
            # pixels move dbeta by at most (1 - slope) * |dz| each (dgamma by that times x_hat ~ 0).  If the flat
            # tolerance fails, the (few) offending channels must be explained by that bound.
            e = rel(r["dgamma"], gam.grad)
            note("norm bwd dgamma", e)
            assert e <= TOL_DW, f"unit {r['li']} dgamma: {e:.2e}"
            e = rel(r["dbeta"], bet.grad)
            if e > TOL_DW:
                edge = (pre.detach().abs() < 1e-5).float() * dz.abs()
                if f["scale"] is not None:
                    edge = edge * f["scale"][:, :, None, None]
                budget = edge.sum((0, 2, 3)) * (1 - f["slope"])
                diff = (r["dbeta"] - bet.grad).abs()
                flat = TOL_DW * bet.grad.abs().max()
                assert float((diff - budget - flat).max()) <= 0, f"unit {r['li']} dbeta beyond the knife-edge budget"
                assert int((diff > flat).sum()) <= max(1, diff.numel() // 50), f"unit {r['li']} dbeta: too many channels off"
                e = rel(torch.where(diff > flat, bet.grad, r["dbeta"]), bet.grad)
            note("norm bwd dbeta", e)
            assert e <= TOL_DW, f"unit {r['li']} dbeta: {e:.2e}"
            # ---- conv backward from the SAME dy the kernels consumed (bf16)
            dy = nchw(r["dy"])
            w = conv.weight.detach().bfloat16().float()
            xin = f["xin"] if r.get("xin") is None else r["xin"]
            xin = nchw(xin)[:, :w.shape[1]]
            dw_ref = torch.nn.grad.conv2d_weight(xin, w.shape, dy, stride=f["stride"], padding=1)
            e = rel(r["dw"], dw_ref)
            note("wgrad", e)
            assert e <= TOL_DW, f"unit {r['li']} wgrad {tuple(w.shape)}: {e:.2e}"
            if r.get("dx") is not None:
                dx_ref = torch.nn.grad.conv2d_input(xin.shape, w, dy, stride=f["stride"], padding=1)
                e = rel(nchw(r["dx"]), dx_ref)
                note("dgrad", e)
                assert e <= TOL_BF16, f"unit {r['li']} dgrad {tuple(w.shape)}: {e:.2e}"
    print("layerwise backward worst rel-L2:", {k: f"{v:.2e}" for k, v in worst.items()})
    # the parameter gradients autograd received ARE the traced tensors (22 conv + 22 x 2 norm + head)
    for r in bunits:
        conv = funits[r["li"]]["conv"]
        assert torch.equal(conv.weight.grad, r["dw"].reshape(conv.weight.shape))
