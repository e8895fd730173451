"""Randomised shape sweep of the convolution family: the tcgen05 kernels (IMPL_AUTO) against the library's own
CUDA-core implementation (IMPL_DIRECT) on identical inputs.  Shapes are drawn to hit the tile planner's edge cases:
odd extents, tiles straddling image borders, partial N tiles, one / two sources, CTA pairs with an odd tile count,
MB = 1/2/4, paired and grouped backward-weights.  (Both sides are OUR kernels; the direct one is itself pinned against
torch.nn.functional in test_ops_gpu.py.)"""
import random

import pytest
import torch

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def _cases():
    rng = random.Random(20240521)
    out = []
    chans = [8, 16, 24, 32, 40, 64, 72, 96, 128, 192, 256]
    for i in range(28):
        n = rng.choice([1, 1, 2, 3])
        c0 = rng.choice(chans)
        c1 = rng.choice([0, 0, 8, 32, 64, 128]) if i % 3 == 0 else 0
        cout = rng.choice(chans)
        k = 3 if i % 7 else 1
        pad = rng.choice([0, 1]) if k == 3 else 0
        h = rng.randint(3 if pad else 5, 70)
        w = rng.randint(3 if pad else 5, 90)
        out.append((n, c0, c1, cout, h, w, k, pad))
    out += [(1, 64, 0, 64, 131, 67, 3, 0), (2, 128, 128, 128, 29, 31, 3, 1), (1, 256, 0, 512, 9, 140, 3, 0),
            (4, 64, 64, 64, 40, 40, 3, 0), (1, 512, 0, 64, 17, 19, 3, 1)]
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: "x".join(map(str, c)))
def test_tcgen05_matches_direct(case):
    from b200unet import ops
    n, c0, c1, cout, h, w, k, pad = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(hash(case) % 100000)
    cs = [c0] + ([c1] if c1 else [])
    srcs = [torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16) for c in cs]
    cin = sum(cs)
    wt = (torch.randn(cout, cin, k, k, device=dev, generator=g) / (cin * k * k) ** 0.5)
    wq = wt.to(torch.bfloat16).float()  # both implementations then multiply the same weight values
    b = torch.randn(cout, device=dev, generator=g)
    ya = ops.conv_fwd(srcs, wq, b, pad, True, impl=ops.IMPL_AUTO)
    yd = ops.conv_fwd(srcs, wq, b, pad, True, impl=ops.IMPL_DIRECT)
    assert rel_l2(ya.float(), yd.float()) < 4e-3
    dz = (torch.randn(ya.shape, device=dev, generator=g) * (yd.float() > 0)).to(torch.bfloat16)
    masks = [torch.randn(s.shape, device=dev, generator=g).to(torch.bfloat16) for s in srcs]
    da = [torch.empty_like(s) for s in srcs]
    dd = [torch.empty_like(s) for s in srcs]
    ops.conv_dgrad(dz, wq, pad, da, masks, impl=ops.IMPL_AUTO)
    ops.conv_dgrad(dz, wq, pad, dd, masks, impl=ops.IMPL_DIRECT)
    for x, y in zip(da, dd):
        assert rel_l2(x.float(), y.float()) < 4e-3
    dwa, dba = ops.conv_wgrad(dz, srcs, k, pad, impl=ops.IMPL_AUTO)
    dwd, dbd = ops.conv_wgrad(dz, srcs, k, pad, impl=ops.IMPL_DIRECT)
    assert rel_l2(dwa, dwd) < 1e-3
    assert rel_l2(dba, dbd) < 1e-4 or float(dbd.abs().max()) < 1e-6
