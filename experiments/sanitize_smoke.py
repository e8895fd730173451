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

# round 2: FusedAdam, uint8 input path, CVO kernels (fused forward / backward / w-v, materialising drop-ins)
from b200unet import cvo
m = b200unet.UNet(1, 2, 2, 5, True, False, "upconv").cuda().train()
opt = b200unet.FusedAdam(m.parameters(), lr=1e-3, model=m)
u8 = torch.randint(0, 256, (2, 36, 44, 1), dtype=torch.uint8, device="cuda")
y = torch.randint(0, 2, (2, 36, 44), device="cuda")
for _ in range(2):
    loss = m.loss(b200unet.pack_images(m, u8), y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
g = torch.Generator().manual_seed(1)
mk = lambda c, n, s: (torch.randn(1, c, n, generator=g) * s).cuda().requires_grad_(True)
f = [{"xyz": mk(3, 150, 0.3), "img": mk(5, 150, 0.5), "feature": mk(6, 150, 0.1)},
     {"xyz": mk(3, 131, 0.3), "img": mk(5, 131, 0.5), "feature": mk(6, 131, 0.1)}]
L = cvo.cvo_losses(f, ["xyz", "img", "feature"], {"xyz": 0.2, "img": 0.5, "feature": 0.1})
L["func_dist"].backward()
cvo.calc_w_v([f[0][k].detach() for k in ("xyz", "img", "feature")], [f[1][k].detach() for k in ("xyz", "img", "feature")],
             [0.2, 0.5, 0.1], 0)
k = cvo.kern_mat(f[0]["xyz"], f[1]["xyz"], 0.2)
k.sum().backward()
cvo.cross_prod(f[0]["xyz"].detach(), f[1]["xyz"].detach())
torch.cuda.synchronize()
print("ok round-2 kernels", float(L["func_dist"]))
