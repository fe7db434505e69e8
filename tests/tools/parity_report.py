"""Developer tool: error of the CUDA path against the fp32 CPU oracle, next to the error of the oracle itself when it
is run under bf16 autocast on the CPU (the yardstick for what bf16 storage costs); and the fp32 verification mode of the
CUDA path, the fp32 oracle and the bf16 CUDA path against the fp64 oracle.  Not part of the test suite.
    python tests/tools/parity_report.py [size ...]  -> gpurun_out/parity_report.json, gpurun_out/parity_report.md"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import unet_oracle as O
from unet_implementations_b200.models.losses import SimpleLoss
from unet_implementations_b200.models.unet import UNet


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
    out = {}
    for size in sizes:
        torch.manual_seed(1234)
        model = UNet()
        cfg = O.config_of(model)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        model = model.cuda().train()
        B = 1 if size >= 512 else 2
        x, target = O.synthetic_batch(B, size, seed=0)
        torch.manual_seed(99)
        masks = O.draw_dropout_masks(cfg, B, x)
        model._mask_override = masks
        logits = model(x.cuda())
        loss = SimpleLoss()(logits, target.cuda())
        loss.backward()
        ref = O.training_step(sd, x, target, cfg, masks)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ref16 = O.training_step(sd, x, target, cfg, masks)
        refm = O.training_step(sd, x, target, cfg, masks, bf16_storage=True)
        rec = {"logits": O.rel_l2(logits, ref["logits"]), "logits_ref_bf16": O.rel_l2(ref16["logits"], ref["logits"]),
               "loss": abs(loss.item() - ref["loss"].item()) / ref["loss"].item(),
               "loss_ref_bf16": abs(ref16["loss"].item() - ref["loss"].item()) / ref["loss"].item(),
               "logits_vs_matched": O.rel_l2(logits, refm["logits"]),
               "matched_vs_fp32": O.rel_l2(refm["logits"], ref["logits"]), "grads": {}}
        for k, p in model.named_parameters():
            if p.dim() == 1 and k.endswith("bias") and "segmentation" not in k and ref["grads"][k].abs().max() < 1e-5:
                continue
            rec["grads"][k] = [O.rel_l2(p.grad, ref["grads"][k]), O.rel_l2(ref16["grads"][k], ref["grads"][k]),
                               O.rel_l2(p.grad, refm["grads"][k])]
        # ---- against the fp64 oracle: this path in fp32 mode, the reference's fp32 ops, this path in bf16
        ref64 = O.training_step(sd, x, target, cfg, masks, dtype=torch.float64)
        bf16_grads = {k: p.grad.detach().double().cpu() for k, p in model.named_parameters()}
        bf16_logits, bf16_loss = logits.detach().double().cpu(), loss.item()
        model.precision = "fp32"
        model.zero_grad(set_to_none=True)
        model._mask_override = masks
        l32 = model(x.cuda())
        loss32 = SimpleLoss()(l32, target.cuda())
        loss32.backward()
        keys = [k for k in rec["grads"]]
        e_fp32mode = [O.rel_l2(dict(model.named_parameters())[k].grad.double().cpu(), ref64["grads"][k]) for k in keys]
        e_ref32 = [O.rel_l2(ref["grads"][k].double(), ref64["grads"][k]) for k in keys]
        e_bf16 = [O.rel_l2(bf16_grads[k], ref64["grads"][k]) for k in keys]
        med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
        rec["vs_fp64"] = {
            "fp32_mode": {"logits": O.rel_l2(l32.double().cpu(), ref64["logits"]),
                          "loss": abs(loss32.item() - ref64["loss"].item()) / abs(ref64["loss"].item()),
                          "grad_worst": max(e_fp32mode), "grad_median": med(e_fp32mode),
                          "argmax_mismatch": int((l32.argmax(1).cpu() != ref64["logits"].argmax(1)).sum())},
            "reference_fp32": {"logits": O.rel_l2(ref["logits"].double(), ref64["logits"]),
                               "loss": abs(ref["loss"].item() - ref64["loss"].item()) / abs(ref64["loss"].item()),
                               "grad_worst": max(e_ref32), "grad_median": med(e_ref32),
                               "argmax_mismatch": int((ref["logits"].argmax(1) != ref64["logits"].argmax(1)).sum())},
            "bf16_mode": {"logits": O.rel_l2(bf16_logits, ref64["logits"]),
                          "loss": abs(bf16_loss - ref64["loss"].item()) / abs(ref64["loss"].item()),
                          "grad_worst": max(e_bf16), "grad_median": med(e_bf16),
                          "argmax_mismatch": int((bf16_logits.argmax(1) != ref64["logits"].argmax(1)).sum())},
            "reference_bf16_autocast": {"logits": O.rel_l2(ref16["logits"].double(), ref64["logits"]),
                                        "argmax_mismatch": int((ref16["logits"].float().argmax(1) != ref64["logits"].argmax(1)).sum())},
            "pixels": int(target.numel())}
        out[size] = rec
        print(size, {k: v for k, v in rec.items() if k != "grads"})
        for k, v in rec["grads"].items():
            print(f"   {k:55s} ours {v[0]:.4f}   ref-bf16 {v[1]:.4f}   ours-vs-matched {v[2]:.4f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)
    with open("gpurun_out/parity_report.md", "w") as f:
        f.write("| size | who | logits rel-L2 | loss rel | worst gradient tensor | median gradient tensor | argmax mismatches |\n|---|---|---|---|---|---|---|\n")
        for size, rec in out.items():
            v = rec["vs_fp64"]
            for who in ("fp32_mode", "reference_fp32", "bf16_mode"):
                w = v[who]
                f.write(f"| {size} | {who} | {w['logits']:.2e} | {w['loss']:.2e} | {w['grad_worst']:.2e} | {w['grad_median']:.2e} | "
                        f"{w['argmax_mismatch']} / {v['pixels']} |\n")
            w = v["reference_bf16_autocast"]
            f.write(f"| {size} | reference_bf16_autocast | {w['logits']:.2e} | | | | {w['argmax_mismatch']} / {v['pixels']} |\n")


if __name__ == "__main__":
    main()
