"""Developer tool: per-unit dgamma / dbeta errors of the traced step (tests/test_gpu_layerwise.py) printed separately."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import unet_oracle as O
from unet_implementations_b200.models.losses import SimpleLoss
from unet_implementations_b200.models.unet import UNet

def nchw(t): return t.float().permute(0, 3, 1, 2).contiguous()
def rel(a, b): return ((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30)).item()
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(1234)
model = UNet().cuda().train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x, target = O.synthetic_batch(B, 512, seed=0)
model._trace_fwd, model._trace_bwd = [], []
torch.manual_seed(99)
loss = SimpleLoss()(model(x.cuda()), target.cuda()); loss.backward(); torch.cuda.synchronize()
fu = {r["li"]: r for r in model._trace_fwd if r["kind"] == "unit"}
for r in model._trace_bwd:
    if r["kind"] != "unit": continue
    f = fu[r["li"]]; norm = f["norm"]
    y = nchw(r["y"]).double().requires_grad_(True)
    gam = norm.weight.detach().double().requires_grad_(True); bet = norm.bias.detach().double().requires_grad_(True)
    z = F.leaky_relu(F.instance_norm(y, weight=gam, bias=bet, eps=norm.eps), f["slope"])
    if f["scale"] is not None: z = z * f["scale"].double()[:, :, None, None]
    dz = nchw(r["dz"]).double()
    if r["dz2"] is not None: dz = dz + nchw(r["dz2"]).double()
    z.backward(dz)
    dg, db = r["dgamma"].double(), r["dbeta"].double()
    eg, eb = rel(dg, gam.grad), rel(db, bet.grad)
    worst_c = (dg - gam.grad).abs().argmax().item()
    print(f"unit {r['li']:2d} C={y.shape[1]:4d} HW={y.shape[2]*y.shape[3]:7d} dz2={r['dz2'] is not None} drop={f['scale'] is not None} "
          f"dgamma {eg:.2e} dbeta {eb:.2e} |dgamma| {gam.grad.norm():.3e} |dbeta| {bet.grad.norm():.3e} worst c {worst_c}: ours {dg[worst_c]:.6e} ref {gam.grad[worst_c]:.6e}")
    if eb > 1e-4 or eg > 1e-4:
        d = (db - bet.grad).abs()
        top = d.topk(8)
        print("   dbeta worst channels", top.indices.tolist(), [f"{v:.3e}" for v in top.values.tolist()])
        print("   ours", [f"{db[c]:.5e}" for c in top.indices.tolist()], "ref", [f"{bet.grad[c]:.5e}" for c in top.indices.tolist()])
        # per-image contributions of the worst channel from first principles
        c = top.indices[0].item()
        xh = (y.detach() - y.detach().mean((2, 3), keepdim=True)) * (y.detach().var((2, 3), unbiased=False, keepdim=True) + norm.eps).rsqrt()
        pre = xh * gam.detach()[None, :, None, None] + bet.detach()[None, :, None, None]
        g = dz * torch.where(pre > 0, 1.0, f["slope"])
        if f["scale"] is not None: g = g * f["scale"].double()[:, :, None, None]
        print("   per-image sum g for channel", c, g[:, c].sum((1, 2)).tolist(), "scale", None if f["scale"] is None else f["scale"][:, c].tolist())
        print("   |pre| min for that channel", pre[:, c].abs().min().item(), "count |pre|<1e-6", (pre[:, c].abs() < 1e-6).sum().item())
        print("   gamma,beta", gam[c].item(), bet[c].item(), "a", f["a"][:, c].tolist(), "b", f["b"][:, c].tolist())
