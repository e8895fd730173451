"""Whole training step (forward + fused loss + backward + Adam) captured in ONE CUDA graph: for the narrow / small
configurations the eager step is bound by Python launch overhead, not by the GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
sys.path.insert(0, ROOT)
import torch
import b200unet
from oracle import unet_oracle as O

CFG = {
    "config5": ((3, 6, 5, 2, True, True, "upsample", True), "deep", 12, 192, 640),
    "config2": ((1, 2, 5, 6, True, True, "upsample"), "paper", 16, 256, 256),
    "config3": ((1, 2, 5, 6, False, False, "upconv"), "paper", 32, 572, 572),
}
for name in sys.argv[1:] or ["config5"]:
    args, ub, b, h, w = CFG[name]
    torch.manual_seed(0)
    m = b200unet.UNet(*args, up_block=ub).cuda().train()
    opt = b200unet.FusedAdam(m.parameters(), lr=1e-4, model=m)
    spec = O.UNetSpec(*args[:7], non_neg=(args[7] if len(args) > 7 else False), up_block=ub)
    ho, wo = O.output_hw(spec, h, w)
    x = torch.randn(b, args[0], h, w, device="cuda")
    y = torch.randint(0, args[1], (b, ho, wo), device="cuda")

    def step():
        loss = m.loss(x, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    # warm-up and eager timing on a side stream: autograd nodes created on the legacy default stream and still alive
    # at capture time would make the capture wait on that stream and invalidate it
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t0) * 100
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        l_static = step()
    g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    graph_ms = (time.perf_counter() - t0) * 50
    print(f"{name}: eager {eager_ms:.2f} ms/step, CUDA graph {graph_ms:.2f} ms/step, loss {float(l_static):.5f}", flush=True)
