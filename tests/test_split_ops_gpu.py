"""Split precision tier (include/b200unet.h: value = hi + lo bf16 planes) — every forward operator through the C ABI
against its torch.nn.functional fp32 counterpart evaluated on exactly the values the planes carry.

What limits the agreement: weights are carried as hi + lo too (2^-17 relative), the lo*lo operand term is dropped
(2^-18), accumulation is fp32 in a different order, and the result is re-split (2^-17).  Tolerances below are written
as relative L2 errors against fp32; the plain bf16 tier sits at ~2e-3 on the same comparisons.
"""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu

TOL = 2e-5


def split_of(x_nchw):
    """fp32 NCHW -> Split (NHWC planes) and the fp32 NCHW value the planes carry."""
    from b200unet import ops
    s = ops.Split.from_float(x_nchw.permute(0, 2, 3, 1).contiguous())
    return s, s.float().permute(0, 3, 1, 2).contiguous()


def w16(w):
    hi = w.to(torch.bfloat16).float()
    return hi + (w - hi).to(torch.bfloat16).float()


def to_nchw(s):
    return s.float().permute(0, 3, 1, 2).contiguous()


def test_input_conversion_carries_16_bits():
    from b200unet import ops
    torch.manual_seed(0)
    x = torch.randn(2, 3, 20, 24, device="cuda") * 3
    s = ops.to_nhwc(x, split=True)
    assert isinstance(s, ops.Split) and s.hi.shape == (2, 20, 24, 3)
    assert torch.equal(s.hi, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    err = (to_nchw(s) - x).abs() / x.abs().clamp_min(1e-30)
    assert float(err.max()) <= 2.0 ** -16


@pytest.mark.parametrize("cins,cout,k,pad,relu,hw", [
    ([64], 64, 3, 1, True, (20, 28)),
    ([64], 128, 3, 0, True, (18, 22)),
    ([128, 64], 64, 3, 1, True, (12, 20)),       # folded concat: 2 sources x 3 operand passes = 6 A sources
    ([16], 16, 3, 1, True, (24, 40)),            # narrow (channel-padded) layers of the repo's feature net
    ([24, 16], 64, 3, 1, True, (16, 24)),
    ([256], 128, 1, 0, False, (10, 14)),         # 1x1 after the bilinear upsample (unet.py:147)
])
def test_conv_fwd_split(cins, cout, k, pad, relu, hw):
    from b200unet import ops
    torch.manual_seed(1)
    h, w = hw
    srcs, refs = zip(*[split_of(torch.randn(2, c, h, w, device="cuda")) for c in cins])
    wt = torch.randn(cout, sum(cins), k, k, device="cuda") / (sum(cins) * k * k) ** 0.5
    b = torch.randn(cout, device="cuda")
    out = ops.conv_fwd(list(srcs), wt, b, pad, relu, impl=ops.IMPL_UMMA)
    assert isinstance(out, ops.Split)
    ref = F.conv2d(torch.cat(refs, 1), w16(wt), b, padding=pad)
    ref = F.relu(ref) if relu else ref
    e = rel_l2(to_nchw(out), ref)
    e_hi = rel_l2(out.hi.float().permute(0, 3, 1, 2), ref)
    print(f"conv split {cins}->{cout} k{k}: rel-L2 {e:.2e} (hi plane alone {e_hi:.2e})")
    assert e <= TOL
    assert e_hi > 20 * e  # the lo plane is doing the work


def test_first_layer_conv_split():
    from b200unet import ops
    torch.manual_seed(2)
    for cin in (1, 3):
        x = torch.randn(2, cin, 30, 34, device="cuda")
        s = ops.to_nhwc(x, split=True)
        wt = torch.randn(64, cin, 3, 3, device="cuda") / 3
        b = torch.randn(64, device="cuda")
        out = ops.conv_fwd([s], wt, b, 1, True)
        assert isinstance(out, ops.Split)
        ref = F.relu(F.conv2d(to_nchw(s), wt, b, padding=1))  # this kernel multiplies fp32 weights
        assert rel_l2(to_nchw(out), ref) <= TOL


def test_convt_fwd_split():
    from b200unet import ops
    torch.manual_seed(3)
    s, ref_x = split_of(torch.randn(2, 128, 9, 11, device="cuda"))
    wt = torch.randn(128, 64, 2, 2, device="cuda") / 128 ** 0.5
    b = torch.randn(64, device="cuda")
    out = ops.convt_fwd(s, wt, b, impl=ops.IMPL_UMMA)
    assert isinstance(out, ops.Split) and out.shape == (2, 18, 22, 64)
    ref = F.conv_transpose2d(ref_x, w16(wt), b, stride=2)
    assert rel_l2(to_nchw(out), ref) <= TOL


def test_batchnorm_fwd_split():
    from b200unet import ops
    torch.manual_seed(4)
    s, ref_x = split_of(F.relu(torch.randn(4, 32, 18, 22, device="cuda") + 0.3))
    g, bt = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda")
    rm, rv = torch.zeros(32, device="cuda"), torch.ones(32, device="cuda")
    out, mean, invstd = ops.bn_fwd_train(s, g, bt, rm, rv, 0.1, 1e-5)
    rm_ref, rv_ref = torch.zeros(32, device="cuda"), torch.ones(32, device="cuda")
    ref = F.batch_norm(ref_x, rm_ref, rv_ref, g, bt, training=True, momentum=0.1, eps=1e-5)
    assert isinstance(out, ops.Split)
    assert rel_l2(to_nchw(out), ref) <= TOL
    assert rel_l2(rm, rm_ref) <= 1e-5 and rel_l2(rv, rv_ref) <= 1e-5
    ev = ops.bn_fwd_eval(s, g, bt, rm, rv, 1e-5)
    ref_ev = F.batch_norm(ref_x, rm, rv, g, bt, training=False, eps=1e-5)
    assert rel_l2(to_nchw(ev), ref_ev) <= TOL


def test_maxpool_fwd_split_is_exact():
    from b200unet import ops
    torch.manual_seed(5)
    x = torch.randn(2, 16, 12, 20, device="cuda")
    x[0, :, 2:4, 2:4] = 1.0  # ties
    hi = x.to(torch.bfloat16).float()
    x[1, :, 4:6, 4:6] = hi[1, :, 4:6, 4:6]  # equal hi planes, ...
    x[1, :, 5, 5] += hi[1, :, 5, 5].abs() * 2.0 ** -12  # ... decided by the lo plane
    s, ref_x = split_of(x)
    y, idx8, idx64 = ops.maxpool_fwd(s, want_idx64=True)
    ref, ref_idx = F.max_pool2d(ref_x, 2, return_indices=True)
    assert torch.equal(to_nchw(y), ref)
    assert torch.equal(idx64.permute(0, 3, 1, 2), ref_idx)


def test_bilinear_fwd_split():
    from b200unet import ops
    torch.manual_seed(6)
    s, ref_x = split_of(torch.randn(2, 24, 7, 9, device="cuda"))
    y = ops.bilinear_fwd(s)
    ref = F.interpolate(ref_x, scale_factor=2, mode="bilinear", align_corners=False)
    assert isinstance(y, ops.Split) and rel_l2(to_nchw(y), ref) <= TOL


def test_head_split():
    from b200unet import ops
    torch.manual_seed(7)
    s, ref_x = split_of(torch.randn(2, 64, 10, 14, device="cuda"))
    w = torch.randn(6, 64, device="cuda") / 8
    b = torch.randn(6, device="cuda")
    logits = ops.head_fwd(s, w, b, True)
    ref = F.relu(F.conv2d(ref_x, w.view(6, 64, 1, 1), b))
    assert rel_l2(logits, ref) <= 1e-5
    y = torch.randint(0, 6, (2, 10, 14), device="cuda")
    loss, _, _ = ops.head_ce_fwd(s, w, b, True, y)
    assert abs(float(loss) - float(F.cross_entropy(ref, y))) <= 1e-5


def test_tier_mismatch_is_an_error():
    from b200unet import ops
    s, _ = split_of(torch.randn(1, 64, 8, 8, device="cuda"))
    wt = torch.randn(64, 64, 3, 3, device="cuda")
    plain_out = torch.empty(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="precision tier"):
        ops.conv_fwd([s], wt, None, 1, True, out=plain_out)
    with pytest.raises(RuntimeError, match="split precision tier"):
        ops.conv_fwd([s], wt, None, 1, True, impl=ops.IMPL_DIRECT)
