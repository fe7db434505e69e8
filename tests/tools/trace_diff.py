"""Developer tool: layer-by-layer forward difference between the CUDA path and the matched-precision oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import unet_oracle as O
from unet_implementations_b200.models.unet import UNet

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(1234)
model = UNet()
cfg = O.config_of(model)
sd = {k: v.clone() for k, v in model.state_dict().items()}
model = model.cuda().train()
x, target = O.synthetic_batch(1, size, seed=0)
torch.manual_seed(99)
masks = O.draw_dropout_masks(cfg, 1, x)
model._mask_override = masks
model._trace = []
with torch.no_grad():
    logits = model(x.cuda())
for mode in (True, False):
    O.TRACE = []
    ref = O.unet_forward(sd, x, cfg, masks, True, bf16_storage=mode)
    print("matched" if mode else "fp32", "logits", O.rel_l2(logits, ref))
    for i, ((y, z), (yr, zr)) in enumerate(zip(model._trace, O.TRACE)):
        yr = yr - yr.mean(dim=(2, 3), keepdim=True) if False else yr
        print(f"  unit {i:2d} C={y.shape[-1]:4d} HW={y.shape[1]:4d}  y {O.rel_l2(y.float().permute(0,3,1,2), yr):.5f}  z {O.rel_l2(z.float().permute(0,3,1,2), zr):.5f}")
