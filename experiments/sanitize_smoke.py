"""Tiny invocation of every kernel family (for compute-sanitizer memcheck: one tool per run, small shapes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import b200unet
torch.manual_seed(0)
for args, ub, shape in [((1, 2, 3, 6, False, False, "upconv"), "paper", (2, 1, 60, 68)),
                        ((3, 6, 3, 2, True, True, "upsample", True), "deep", (2, 3, 24, 32))]:
    m = b200unet.UNet(*args, up_block=ub).cuda().train()
    x = torch.randn(*shape, device="cuda")
    out = m(x)
    y = torch.randint(0, args[1], (shape[0], out.shape[2], out.shape[3]), device="cuda")
    F.cross_entropy(out, y).backward()
    m.zero_grad(set_to_none=True)
    m.loss(x, y).backward()
    torch.cuda.synchronize()
    print("ok", args, float(out.abs().mean()))
