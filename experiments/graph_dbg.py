import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch, b200unet
from b200unet import ops, _lib
import ctypes as C
torch.manual_seed(0)
m = (b200unet.UNet(3, 6, 3, 2, True, True, "upsample", True, up_block="deep") if len(sys.argv) > 1 else b200unet.UNet(1, 2, 3, 6, False, False, "upconv")).cuda().train()
x = torch.randn(2, 3 if len(sys.argv) > 1 else 1, 92, 92, device="cuda"); y = torch.randint(0, 2, (2, 92, 92) if len(sys.argv) > 1 else (2, 52, 52), device="cuda")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        l = m.loss(x, y); m.zero_grad(set_to_none=True); l.backward()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
# wrap check to locate the first failing library call during capture
orig_check = ops.check
def chk(rc, what):
    st = torch.cuda.current_stream().cuda_stream
    import ctypes
    # query capture status through the runtime
    status = ctypes.c_int(0)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaStreamIsCapturing(ctypes.c_void_p(st), ctypes.byref(status))
    if status.value == 2:
        print("capture INVALIDATED after", what); sys.stdout.flush(); raise SystemExit(1)
    orig_check(rc, what)
ops.check = chk
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        l = m.loss(x, y)
        print("forward captured"); 
        m.zero_grad(set_to_none=True); l.backward()
        print("backward captured")
    g.replay(); torch.cuda.synchronize(); print("replay ok", float(l))
except BaseException as e:
    print("FAILED:", type(e).__name__, str(e)[:200])
