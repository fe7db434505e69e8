"""Developer tool: row-major against column-major block order in the narrow weight-gradient kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from unet_implementations_b200 import ops
from wgrad_pairs_bench import timeit
g = torch.Generator(device="cuda").manual_seed(3)
for cin, cout, hw in ((32, 32, 512), (96, 32, 512), (64, 64, 256), (192, 64, 256)):
    x = torch.randn(32, hw, hw, cin, device="cuda", generator=g).bfloat16()
    dy = torch.randn(32, hw, hw, cout, device="cuda", generator=g).bfloat16()
    res = {}
    for rep in range(2):
        for cm in ("0", "1"):
            os.environ["B200UNET_WGRADN_COLMAJOR"] = cm
            res[cm] = ops.conv_wgrad(x, dy, 1)
            print(f"{cin}->{cout} @{hw} col_major={cm}: {timeit(lambda: ops.conv_wgrad(x, dy, 1)):.1f} us")
    print("   max rel diff", float((res["0"] - res["1"]).abs().max() / res["0"].abs().max()))
    del x, dy
