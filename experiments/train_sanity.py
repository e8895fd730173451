"""End-to-end sanity at the benchmark size: the paper U-Net (batch 16, 572^2) learns a synthetic but learnable task
(label = sign of the centre-cropped, slightly blurred input) for a few hundred steps with FusedAdam; the loss must fall
well below log(2).  Also a soak test of the CTA-pair kernels (thousands of cluster launches).
    python experiments/train_sanity.py [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch
import torch.nn.functional as F
import b200unet

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(0)
m = b200unet.UNet(1, 2, 5, 6, False, False, "upconv").cuda().train()
opt = b200unet.FusedAdam(m.parameters(), lr=1e-4, model=m)
B = 16
g = torch.Generator(device="cuda").manual_seed(1)
t0 = time.time()
for it in range(steps):
    x = torch.randn(B, 1, 572, 572, device="cuda", generator=g)
    sm = F.avg_pool2d(x, 5, stride=1, padding=2)
    y = (sm[:, 0, 92:92 + 388, 92:92 + 388] > 0).long()
    loss = m.loss(x, y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    if it % 25 == 0 or it == steps - 1:
        print(f"step {it:4d} loss {float(loss):.4f}", flush=True)
torch.cuda.synchronize()
print(f"{steps} steps in {time.time() - t0:.1f} s; final loss {float(loss):.4f}")
assert float(loss) < 0.45, "the network did not learn"
