"""Single-layer timing harness for the convolution-family kernels (development tool, also the ncu target).

    python experiments/bench_conv.py --op fwd --n 32 --c 64 --cout 64 --h 570 --w 570 [--k 3] [--pad 0] [--iters 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch  # noqa: E402

from b200unet import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="fwd", choices=["fwd", "dgrad", "wgrad", "convt_fwd", "convt_dgrad", "convt_wgrad", "all"])
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--c", type=int, default=64)
ap.add_argument("--c2", type=int, default=0, help="channels of a second (concatenated) source")
ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--h", type=int, default=570)
ap.add_argument("--w", type=int, default=570)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--pad", type=int, default=0)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--impl", type=int, default=0)
a = ap.parse_args()

dev = "cuda"
torch.manual_seed(0)
cs = [a.c] + ([a.c2] if a.c2 else [])
cin = sum(cs)
srcs = [torch.randn(a.n, a.h, a.w, c, device=dev).to(torch.bfloat16) for c in cs]
wt = torch.randn(a.cout, cin, a.k, a.k, device=dev) / (cin * a.k * a.k) ** 0.5
b = torch.randn(a.cout, device=dev)
ho, wo = a.h + 2 * a.pad - a.k + 1, a.w + 2 * a.pad - a.k + 1
flops = 2.0 * a.n * ho * wo * a.cout * cin * a.k * a.k
wp0 = ops.pack_conv_weight(wt, cs, 0)
wp1 = ops.pack_conv_weight(wt, [cin], 1)
y = ops.conv_fwd(srcs, wt, b, a.pad, True, w_packed=wp0, impl=a.impl)
dz = torch.randn_like(y)
dsts = [torch.empty_like(s) for s in srcs]
masks = [torch.randn_like(s) for s in srcs]


def timeit(name, fn, fl):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"{name:12s} n={a.n} c={cs} cout={a.cout} {a.h}x{a.w} k={a.k}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


if a.op in ("fwd", "all"):
    timeit("conv_fwd", lambda: ops.conv_fwd(srcs, wt, b, a.pad, True, w_packed=wp0, impl=a.impl, out=y), flops)
if a.op in ("dgrad", "all"):
    timeit("conv_dgrad", lambda: ops.conv_dgrad(dz, wt, a.pad, dsts, masks, w_packed=wp1, impl=a.impl), flops)
if a.op in ("wgrad", "all"):
    dw = torch.empty_like(wt)
    db = torch.empty_like(b)
    timeit("conv_wgrad", lambda: ops.conv_wgrad(dz, srcs, a.k, a.pad, impl=a.impl, dw=dw, db=db), flops)
if a.op.startswith("convt"):
    x = srcs[0]
    wct = torch.randn(a.c, a.cout, 2, 2, device=dev) / (a.c * 4) ** 0.5
    fl = 2.0 * a.n * a.h * a.w * a.c * a.cout * 4
    yt = ops.convt_fwd(x, wct, b)
    dyt = torch.randn_like(yt)
    if a.op == "convt_fwd":
        timeit("convt_fwd", lambda: ops.convt_fwd(x, wct, b, out=yt), fl)
    elif a.op == "convt_dgrad":
        dx = torch.empty_like(x)
        timeit("convt_dgrad", lambda: ops.convt_dgrad(dyt, wct, dx, mask=x), fl)
    else:
        timeit("convt_wgrad", lambda: ops.convt_wgrad(x, dyt), fl)
