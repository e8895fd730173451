import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch, torch.nn.functional as F
import b200unet
from b200unet import ops
torch.backends.cudnn.allow_tf32 = False
def run(n, c, cout, h, w, k=3, pad=0):
    torch.manual_seed(0)
    x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, c, k, k, device="cuda") / (c * k * k) ** 0.5
    y = ops.conv_fwd([x], wt, None, pad, False)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), padding=pad).permute(0, 2, 3, 1)
    err = (y.float() - ref).abs().amax(dim=3)  # n, h, w
    bad = err > 0.05
    print(f"case n={n} c={c} cout={cout} h={h} w={w}: bad positions {int(bad.sum())} of {bad.numel()}")
    if bad.any():
        for i in range(n):
            rows = bad[i].any(dim=1).nonzero().flatten().tolist()
            cols = bad[i].any(dim=0).nonzero().flatten().tolist()
            print("  img", i, "bad rows", rows[:40], "bad cols", cols[:70])
run(1, 64, 64, 12, 40)
run(1, 64, 64, 12, 23)
run(1, 64, 64, 30, 21)
run(1, 64, 64, 40, 70)
run(3, 64, 64, 40, 70)
