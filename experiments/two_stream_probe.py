"""Does running backward-weights on a side stream next to backward-data buy anything?  (Both are persistent 1-CTA/SM
tcgen05 kernels; only tails / launch gaps can overlap.)   python experiments/two_stream_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch
from b200unet import ops

dev = "cuda"
for (c, cout, h) in [(256, 256, 138), (128, 128, 282), (64, 64, 570), (512, 512, 66)]:
    n = 32
    x = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, c, 3, 3, device=dev) / (c * 9) ** 0.5
    y = ops.conv_fwd([x], w, None, 0, True)
    dz = torch.randn_like(y)
    dx = torch.empty_like(x)
    wp = ops.pack_conv_weight(w, [c], 1)
    side = torch.cuda.Stream()
    dw = torch.empty_like(w); db = torch.empty(cout, device=dev)

    def seq():
        ops.conv_wgrad(dz, [x], 3, 0, dw=dw, db=db)
        ops.conv_dgrad(dz, w, 0, [dx], [x], w_packed=wp)

    def par():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.conv_wgrad(dz, [x], 3, 0, dw=dw, db=db)
        ops.conv_dgrad(dz, w, 0, [dx], [x], w_packed=wp)
        torch.cuda.current_stream().wait_stream(side)

    for name, fn in (("sequential", seq), ("two streams", par)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{c}->{cout} @{h}: {name:12s} {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
