import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch, b200unet
torch.manual_seed(0)
m = b200unet.UNet(1, 2, 5, 6, True, True, "upsample").cuda().train()
x = torch.randn(16, 1, 256, 256, device="cuda"); y = torch.randint(0, 2, (16, 256, 256), device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    loss = m.loss(x, y); m.zero_grad(set_to_none=True); loss.backward()
torch.cuda.synchronize(); print("ok", float(loss))
