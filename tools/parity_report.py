"""Developer tool: error of the CUDA path against the fp32 CPU oracle, next to the error of the oracle itself when it
is run under bf16 autocast on the CPU (the yardstick for what bf16 storage costs).  Not part of the test suite.
    python tools/parity_report.py [size ...]  -> gpurun_out/parity_report.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
        out[size] = rec
        print(size, {k: v for k, v in rec.items() if k != "grads"})
        for k, v in rec["grads"].items():
            print(f"   {k:55s} ours {v[0]:.4f}   ref-bf16 {v[1]:.4f}   ours-vs-matched {v[2]:.4f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)


if __name__ == "__main__":
    main()
