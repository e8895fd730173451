import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch, b200unet
from b200unet import ops
torch.manual_seed(0)
m = b200unet.UNet(3, 6, 5, 2, True, True, "upsample", True, up_block="deep").cuda().train()
use_opt = len(sys.argv) > 1
opt = torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True, fused=True) if use_opt else None
x = torch.randn(12, 3, 192, 640, device="cuda"); y = torch.randint(0, 6, (12, 192, 640), device="cuda")
def step():
    l = m.loss(x, y)
    (opt.zero_grad(set_to_none=True) if opt else m.zero_grad(set_to_none=True))
    l.backward()
    if opt: opt.step()
    return l
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
rt = ctypes.CDLL("libcudart.so.12")
orig_check = ops.check
def chk(rc, what):
    status = ctypes.c_int(0)
    rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(status))
    if status.value == 2:
        print("capture INVALIDATED at/before", what, "rc", rc); sys.stdout.flush(); raise SystemExit(1)
    orig_check(rc, what)
ops.check = chk
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        l = step()
    g.replay(); torch.cuda.synchronize(); print("replay ok", float(l.detach()))
except BaseException as e:
    print("FAILED:", type(e).__name__, str(e)[:300])
