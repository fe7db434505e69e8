"""Developer bring-up harness: runs one kernel family per process against torch's own GPU ops and prints a
one-line verdict per case.  Not part of the product or of the parity suite (tests/ uses the CPU oracle);
it exists so that a faulting kernel cannot take the other checks down with it.

    python tests/tools/bringup.py list
    python tests/tools/bringup.py <case> [<case> ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch
import torch.nn.functional as F

from unet_implementations_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"


def report(name, got, ref, tol, extra=None):
    got = got.double()
    ref = ref.double()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    rel_l2 = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    out = {"case": name, "max_abs": err.max().item(), "ref_max": denom, "rel_l2": rel_l2,
           "nan": bool(torch.isnan(got).any().item()), "ok": bool(rel_l2 < tol and not torch.isnan(got).any().item())}
    if extra:
        out.update(extra)
    if not out["ok"] and got.dim() == 4:
        # localise: error by channel block of 8 and by row block
        e = err
        out["err_by_c8"] = [round(v, 4) for v in e.reshape(-1, e.shape[-1]).amax(0).reshape(-1, 8).amax(1).tolist()[:32]]
        out["err_by_h"] = [round(v, 4) for v in e.amax(dim=(0, 2, 3)).tolist()[:32]]
        out["err_by_w"] = [round(v, 4) for v in e.amax(dim=(0, 1, 3)).tolist()[:32]]
    print(json.dumps(out), flush=True)
    return out["ok"]


def rand_act(n, h, w, c, pitch=None, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    pitch = pitch or c
    buf = torch.randn(n, h, w, pitch, generator=g).to(DEV).to(torch.bfloat16)
    return buf[..., :c] if pitch != c else buf


def conv_case(name, n, h, w, cin, cout, stride, simt=False, xpitch=None):
    x = rand_act(n, h, w, cin, xpitch, seed=1)
    g = torch.Generator().manual_seed(2)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).to(DEV)
    wf, wd = ops.pack_conv_weights(wt)
    w_r = wf.float().permute(0, 3, 1, 2).contiguous()  # bf16-rounded OIHW
    y, stats = ops.conv_fprop(x, wf, stride, simt=simt)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w_r, padding=1, stride=stride).permute(0, 2, 3, 1)
    ok = report(name + ".fprop", y.float(), ref, 5e-3)
    yb = y.float()
    s_ref = torch.stack([yb.sum(dim=(1, 2)), (yb * yb).sum(dim=(1, 2))], dim=-1)
    ok &= report(name + ".stats", stats.sum(dim=1), s_ref, 1e-4)
    # dgrad
    oh, ow = y.shape[1], y.shape[2]
    dy = rand_act(n, oh, ow, cout, seed=3)
    dx = ops.conv_dgrad(dy, wd, (h, w), stride, simt=simt)
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w_r.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1, stride=stride)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    ok &= report(name + ".dgrad", dx.float(), xr.grad.permute(0, 2, 3, 1), 5e-3)
    dw = ops.conv_wgrad(x, dy, stride, simt=simt)
    torch.cuda.synchronize()
    ok &= report(name + ".wgrad", dw, wr.grad, 5e-3)
    return ok


def case_stem():
    n, h, w = 2, 48, 80
    g = torch.Generator().manual_seed(5)
    img = torch.randn(n, 3, h, w, generator=g).to(DEV)
    wt = (torch.randn(32, 3, 3, 3, generator=g) * 0.08).to(DEV)
    y, stats = ops.stem_fprop(img, wt)
    torch.cuda.synchronize()
    ref = F.conv2d(img, wt, padding=1).permute(0, 2, 3, 1)
    ok = report("stem.fprop", y.float(), ref, 5e-3)
    yb = y.float()
    ok &= report("stem.stats", stats.sum(dim=1), torch.stack([yb.sum(dim=(1, 2)), (yb * yb).sum(dim=(1, 2))], -1), 1e-4)
    dy = rand_act(n, h, w, 32, seed=6)
    dw = ops.stem_wgrad(img, dy)
    torch.cuda.synchronize()
    wr = wt.clone().requires_grad_(True)
    F.conv2d(img, wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ok &= report("stem.wgrad", dw, wr.grad, 1e-4)
    return ok


def case_norm():
    ok = True
    for (n, h, w, c, p, pitch) in [(2, 16, 24, 32, 0.0, None), (3, 8, 8, 128, 0.3, 192), (2, 32, 32, 96, 0.2, None)]:
        y = rand_act(n, h, w, c, pitch, seed=7) * 1.7 + 0.3
        g = torch.Generator().manual_seed(8)
        gamma = (torch.rand(c, generator=g) + 0.5).to(DEV)
        beta = (torch.randn(c, generator=g) * 0.2).to(DEV)
        drop = None
        if p > 0:
            drop = (torch.rand(n, c, generator=g) > p).float().div(1 - p).to(DEV)
        yb = y.float()
        stats = torch.stack([yb.sum(dim=(1, 2)), (yb * yb).sum(dim=(1, 2))], -1).unsqueeze(1).contiguous()
        mean, rstd, a, b = ops.in_finalize(stats, gamma, beta, drop, 1e-5, h * w)
        z = ops.in_apply(y, a, b, 0.01)
        torch.cuda.synchronize()
        yr = yb.permute(0, 3, 1, 2).clone().requires_grad_(True)
        gr = gamma.clone().requires_grad_(True)
        br = beta.clone().requires_grad_(True)
        zr = F.leaky_relu(F.instance_norm(yr, weight=gr, bias=br, eps=1e-5), 0.01)
        if drop is not None:
            zr = zr * drop[:, :, None, None]
        ok &= report(f"norm.fwd.c{c}", z.float(), zr.permute(0, 2, 3, 1), 6e-3)
        dz = rand_act(n, h, w, c, seed=9)
        dz2 = rand_act(n, h, w, c, seed=10) if c == 128 else None
        dy, dg, db = ops.in_backward(dz, dz2, y, a, b, mean, rstd, drop, gamma, 0.01)
        torch.cuda.synchronize()
        gsum = dz.float() + (dz2.float() if dz2 is not None else 0)
        zr.backward(gsum.permute(0, 3, 1, 2))
        ok &= report(f"norm.bwd.dy.c{c}", dy.float(), yr.grad.permute(0, 2, 3, 1), 8e-3)
        ok &= report(f"norm.bwd.dgamma.c{c}", dg, gr.grad, 5e-3)
        ok &= report(f"norm.bwd.dbeta.c{c}", db, br.grad, 5e-3)
    return ok


def case_upsample():
    ok = True
    n, h, w, c = 2, 6, 10, 64
    x = rand_act(n, h, w, c, seed=11)
    cat = torch.zeros(n, 2 * h, 2 * w, c + 32, dtype=torch.bfloat16, device=DEV)
    ops.upsample2x(x, cat[..., :c])
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ur = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    ok &= report("upsample.fwd", cat[..., :c].float(), ur.permute(0, 2, 3, 1), 4e-3)
    ok &= report("upsample.untouched", cat[..., c:].float() + 1, torch.ones_like(cat[..., c:].float()), 1e-9)
    dcat = rand_act(n, 2 * h, 2 * w, c + 32, seed=12)
    dx = ops.upsample2x_backward(dcat[..., :c])
    torch.cuda.synchronize()
    ur.backward(dcat[..., :c].float().permute(0, 3, 1, 2))
    ok &= report("upsample.bwd", dx.float(), xr.grad.permute(0, 2, 3, 1), 4e-3)
    return ok


def case_head_loss():
    ok = True
    n, h, w = 3, 40, 56
    z = rand_act(n, h, w, 32, seed=13)
    g = torch.Generator().manual_seed(14)
    wt = (torch.randn(3, 32, 1, 1, generator=g) * 0.3).to(DEV)
    bias = torch.randn(3, generator=g).to(DEV)
    logits = ops.head_forward(z, wt, bias)
    torch.cuda.synchronize()
    zr = z.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    lr = F.conv2d(zr, wr, br)
    ok &= report("head.fwd", logits, lr, 1e-5)
    target = torch.randint(0, 3, (n, h, w), generator=g)
    target[torch.rand(n, h, w, generator=g) < 0.1] = 255
    target[1][target[1] == 2] = 0  # image 1 has no class 2
    target = target.to(DEV)
    out, tables = ops.loss_forward(logits, target, None, True, 1.0, 1.0, 255, 1e-5)
    torch.cuda.synchronize()
    # reference loss written out with torch ops (losses.py semantics)
    lg = lr
    valid = target != 255
    cnt = torch.stack([((target == c) & valid).sum() for c in range(3)]).float()
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    wts = valid.sum().float() / cnt
    wts = wts * (3 / wts.sum())
    ce = F.cross_entropy(lg, target, weight=wts, ignore_index=255)
    p = F.softmax(lg, dim=1)
    m = valid.float()
    dice = 0
    for c in range(3):
        tc = (target == c).float() * m
        ic = p[:, c] * m
        inter = (ic * tc).flatten(1).sum(1)
        union = ic.flatten(1).sum(1) + tc.flatten(1).sum(1)
        dice = dice + (1 - ((2 * inter + 1e-5) / (union + 1e-5)).mean())
    dice = dice / 3
    total = ce + dice
    ok &= report("loss.fwd", out, torch.stack([total, ce, dice]).detach(), 1e-5)
    gscale = torch.tensor(3.0, device=DEV)
    dl = ops.loss_backward(logits, target, tables, gscale, 1.0, 1.0, 255)
    torch.cuda.synchronize()
    (total * 3.0).backward(retain_graph=True)
    dl_ref = torch.autograd.grad(total * 3.0, lr, retain_graph=True)[0]
    ok &= report("loss.bwd", dl, dl_ref, 1e-4)
    dz, dw, db = ops.head_backward(dl, z, wt)
    torch.cuda.synchronize()
    zr.grad = None
    wr.grad = None
    br.grad = None
    lr.backward(dl)
    ok &= report("head.bwd.dz", dz.float(), zr.grad.permute(0, 2, 3, 1), 5e-3)
    ok &= report("head.bwd.dw", dw, wr.grad, 1e-4)
    ok &= report("head.bwd.db", db, br.grad, 1e-4)
    return ok


CASES = {
    # name: thunk
    "simt_s1": lambda: conv_case("simt_s1", 2, 12, 20, 16, 24, 1, simt=True),
    "simt_s2": lambda: conv_case("simt_s2", 2, 12, 20, 16, 24, 2, simt=True),
    "tc_64_64_s1": lambda: conv_case("tc_64_64_s1", 1, 16, 16, 64, 64, 1),
    "tc_64_128_s1": lambda: conv_case("tc_64_128_s1", 2, 16, 32, 64, 128, 1),
    "tc_128_256_s1": lambda: conv_case("tc_128_256_s1", 2, 16, 16, 128, 256, 1),
    "tc_32_32_s1": lambda: conv_case("tc_32_32_s1", 2, 32, 32, 32, 32, 1),
    "tc_96_32_s1": lambda: conv_case("tc_96_32_s1", 1, 32, 48, 96, 32, 1),
    "tc_192_64_s1": lambda: conv_case("tc_192_64_s1", 1, 16, 32, 192, 64, 1),
    "tc_32_64_s2": lambda: conv_case("tc_32_64_s2", 2, 32, 32, 32, 64, 2),
    "tc_64_128_s2": lambda: conv_case("tc_64_128_s2", 2, 32, 64, 64, 128, 2),
    "tc_256_512_s2": lambda: conv_case("tc_256_512_s2", 1, 32, 32, 256, 512, 2),
    "tc_odd_s1": lambda: conv_case("tc_odd_s1", 1, 24, 40, 64, 64, 1),
    "tc_odd_s2": lambda: conv_case("tc_odd_s2", 1, 24, 40, 64, 64, 2),
    "tc_pitch": lambda: conv_case("tc_pitch", 2, 16, 16, 64, 64, 1, xpitch=192),
    "tc_512_512_s1": lambda: conv_case("tc_512_512_s1", 2, 16, 16, 512, 512, 1),
    "stem": case_stem,
    "norm": case_norm,
    "upsample": case_upsample,
    "head_loss": case_head_loss,
}

if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "list":
        print(" ".join(CASES))
        sys.exit(0)
    ops.require_device()
    good = True
    for c in sys.argv[1:]:
        try:
            good &= bool(CASES[c]())
        except Exception as e:  # noqa: BLE001 - report and keep going
            print(json.dumps({"case": c, "ok": False, "exception": repr(e)[:400]}), flush=True)
            good = False
    sys.exit(0 if good else 1)
