"""CPU emulation of candidate precision tiers for the BatchNorm graphs, to decide what the CUDA path
must carry before building it.  Runs the oracle graph with its torch.nn.functional calls swapped for emulating
autograd Functions:

  bf16      : today's product — operands and stored activations rounded to bf16 in forward and backward
  split     : forward operands (activations and weights) carried as hi + lo bf16 pairs (~16 mantissa bits), backward
              unchanged: dz bf16, dgrad weights bf16, wgrad/BN-backward read only the hi plane of the saved activation
  split_bwd : like split, but wgrad / BN backward read hi + lo

Usage: python experiments/precision_emu.py [cfg2|cfg5] [batch] [h] [w]
"""
import sys
import types

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import unet_oracle as O  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def split16(x):
    hi = bf(x)
    return hi + bf(x - hi)


MODE = "bf16"
BW_W = "bf"   # dgrad weight precision: bf | split
BW_DZ = "bf"  # gradient storage / operand precision: bf | split


def bw_w(x):
    return split16(x) if BW_W == "split" else bf(x)


def bw_dz(x):
    return split16(x) if BW_DZ == "split" else bf(x)


def fwd_round(x):
    return bf(x) if MODE == "bf16" else split16(x)


def saved_round(x):  # what backward reads of a saved forward activation
    return split16(x) if MODE == "split_bwd" else bf(x)


class Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding, transposed):
        ctx.save_for_backward(x, w)
        ctx.padding, ctx.transposed = padding, transposed
        xr, wr = fwd_round(x), fwd_round(w)
        if transposed:
            return F.conv_transpose2d(xr, wr, b, stride=2)
        return F.conv2d(xr, wr, b, padding=padding)

    @staticmethod
    def backward(ctx, dz):
        x, w = ctx.saved_tensors
        dzb, wb, xb = bw_dz(dz), bw_w(w), saved_round(x)
        with torch.enable_grad():
            xi = xb.detach().requires_grad_(True)
            wi = wb.detach().requires_grad_(True)
            if ctx.transposed:
                out = F.conv_transpose2d(xi, wi, None, stride=2)
            else:
                out = F.conv2d(xi, wi, None, padding=ctx.padding)
            dx, dw = torch.autograd.grad(out, [xi, wi], dzb)
        return dx, dw, dz.sum((0, 2, 3)), None, None


class BN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rm, rv, g, b, training):
        xr = fwd_round(x)
        ctx.save_for_backward(x, g)
        ctx.training = training
        return F.batch_norm(xr, rm, rv, g, b, training=training, momentum=0.1, eps=1e-5)

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        xs = saved_round(x)
        with torch.enable_grad():
            xi = xs.detach().requires_grad_(True)
            gi = g.detach().requires_grad_(True)
            bi = torch.zeros_like(g).requires_grad_(True)
            # statistics come from the forward (fp32 sums over the forward-precision values)
            y = F.batch_norm(xi, None, None, gi, bi, training=True, eps=1e-5)
            dx, dg, db = torch.autograd.grad(y, [xi, gi, bi], bw_dz(dy))
        return dx, None, None, dg, db, None


class Store(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return fwd_round(x)

    @staticmethod
    def backward(ctx, g):
        return bw_dz(g)


class Shim:
    def __getattr__(self, k):
        return getattr(F, k)

    @staticmethod
    def conv2d(x, w, b=None, padding=0):
        return Conv.apply(x, w, b, padding, False)

    @staticmethod
    def conv_transpose2d(x, w, b=None, stride=2):
        return Conv.apply(x, w, b, 0, True)

    @staticmethod
    def batch_norm(x, rm, rv, g, b, training=True, momentum=0.1, eps=1e-5):
        return BN.apply(x, rm, rv, g, b, training)


def run(spec, x, y, mode):
    global MODE
    MODE = mode
    sd = O.init_params(spec, seed=0)
    if mode == "fp32":
        return O.loss_and_grads(sd, x, y, spec)
    realF, realq = O.F, O._q
    O.F, O._q = Shim(), Store.apply
    try:
        return O.loss_and_grads(sd, x, y, spec)
    finally:
        O.F, O._q = realF, realq


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    torch.manual_seed(0)
    if which == "cfg2":
        b, h, w = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (4, 128, 128)
        spec = O.UNetSpec(in_channels=1, n_classes=2, depth=5, wf=6, padding=True, batch_norm=True, up_mode="upsample")
        x = torch.randn(b, 1, h, w)
    else:
        b, h, w = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (4, 192, 640)
        spec = O.UNetSpec(in_channels=3, n_classes=6, depth=5, wf=2, padding=True, batch_norm=True, up_mode="upsample",
                          non_neg=True, up_block="deep")
        x = torch.rand(b, 3, h, w)
    oh, ow = O.output_hw(spec, h, w)
    y = torch.randint(0, spec.n_classes, (b, oh, ow))
    ref_logits, ref_loss, ref_g, _ = run(spec, x, y, "fp32")
    names = list(ref_g.keys())
    flat = lambda g: torch.cat([g[k].flatten() for k in names])
    # clear-margin pixels for argmax agreement
    top2 = ref_logits.topk(2, dim=1).values
    global BW_W, BW_DZ
    for mode, BW_W, BW_DZ in (("split", "split", "bf"), ("split", "bf", "split"), ("split", "split", "split")):
        print("dgrad weights", BW_W, "dz", BW_DZ)
        logits, loss, g, _ = run(spec, x, y, mode)
        per = sorted(O.rel_l2(g[k], ref_g[k]) for k in names)
        agree = (logits.argmax(1) == ref_logits.argmax(1)).float().mean().item()
        print(f"{which} {mode:10s} logits {O.rel_l2(logits, ref_logits):.2e}  argmax {100 * agree:.3f}%  "
              f"grad-all {O.rel_l2(flat(g), flat(ref_g)):.2e}  per-tensor max {per[-1]:.2e} med {per[len(per) // 2]:.2e}  "
              f"loss {loss.item():.6f} vs {ref_loss.item():.6f}", flush=True)


if __name__ == "__main__":
    main()
