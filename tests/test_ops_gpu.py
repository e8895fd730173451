"""Kernel-level parity: every C-ABI operator against the torch.nn.functional call the reference makes
(unet.py:79, 92-100, 143-148, 65-71; README.md:58), on identical bf16-rounded inputs.  Integer results (pool
arg-max) are bit-exact; floating-point results are compared in fp32 with the tolerance written at each assert
(bf16 output rounding = 2^-9 relative per element)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import nchw, nhwc, rb, rel_l2

pytestmark = pytest.mark.gpu

BF16_OUT = 4e-3   # rel-L2 of a bf16-rounded tensor vs its fp32 value is ~2e-3
F32_OUT = 2e-3    # fp32 outputs accumulated from bf16 operands (order-of-summation noise only... plus bf16 weights)


@pytest.fixture(scope="module")
def ops():
    import b200unet
    assert torch.cuda.is_available()
    b200unet.load_library()
    return b200unet.ops


def dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------ layout
@pytest.mark.parametrize("shape", [(2, 1, 17, 23), (1, 3, 8, 9), (2, 64, 5, 7), (1, 72, 6, 6)])
def test_layout_roundtrip(ops, shape):
    x = torch.randn(shape, device=dev())
    y = ops.to_nhwc(x)
    assert torch.equal(y, nhwc(x))
    assert torch.equal(ops.to_nchw(y), nchw(y))


# ------------------------------------------------------------------ max pool (bit-exact incl. indices)
@pytest.mark.parametrize("n,c,h,w", [(2, 64, 12, 20), (1, 4, 9, 7), (2, 8, 6, 10), (1, 128, 5, 5), (1, 1, 4, 4)])
def test_maxpool_fwd_bit_exact(ops, n, c, h, w):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, c, h, w, device=dev(), generator=g)
    x = F.relu(x)  # whole windows of zeros -> exercises the tie rule
    x[0, 0, 0, 0] = float("nan")
    if h >= 4 and w >= 4:
        x[0, 0, 2, 2] = 1.0
        x[0, 0, 2, 3] = 1.0
        x[0, 0, 3, 2] = 1.0  # 3-way tie
    xb = nhwc(x)
    y, idx8, idx64 = ops.maxpool_fwd(xb, want_idx64=True)
    ref, ridx = F.max_pool2d(nchw(xb), 2, return_indices=True)
    got = nchw(y)
    assert torch.equal(torch.isnan(got), torch.isnan(ref))
    assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(ref))
    assert torch.equal(idx64.permute(0, 3, 1, 2), ridx)
    # idx8 code is consistent with idx64
    oh = torch.arange(h // 2, device=dev()).view(1, -1, 1, 1)
    ow = torch.arange(w // 2, device=dev()).view(1, 1, -1, 1)
    code = idx8.long()
    assert torch.equal((2 * oh + code // 2) * w + 2 * ow + code % 2, idx64)


@pytest.mark.parametrize("n,c,h,w,crop", [(2, 64, 12, 20, (2, 4, 6, 10)), (1, 4, 9, 7, None), (1, 8, 7, 9, (1, 1, 4, 6))])
@pytest.mark.parametrize("use_mask", [False, True])
def test_maxpool_bwd(ops, n, c, h, w, crop, use_mask):
    x = F.relu(torch.randn(n, c, h, w, device=dev()))
    xb = nhwc(x)
    y, idx8 = ops.maxpool_fwd(xb)
    dy = nhwc(torch.randn(n, c, h // 2, w // 2, device=dev()))
    dx = torch.full_like(xb, 7.0)  # garbage that must be overwritten
    add = None
    ref_add = torch.zeros(n, c, h, w, device=dev())
    ay = ax = 0
    if crop is not None:
        ay, ax, ah, aw = crop
        addv = nhwc(torch.randn(n, c, ah, aw, device=dev()))
        dx[:, ay:ay + ah, ax:ax + aw, :] = addv      # the decoder wrote its window into dx already
        add = dx[:, ay:ay + ah, ax:ax + aw, :]
        ref_add[:, :, ay:ay + ah, ax:ax + aw] = nchw(addv)
    mask = xb if use_mask else None
    ops.maxpool_bwd(dy, idx8, dx, add=add, add_y=ay, add_x=ax, mask=mask)
    xr = nchw(xb).requires_grad_(True)
    F.max_pool2d(xr, 2).backward(nchw(dy))
    ref = xr.grad + ref_add
    if use_mask:
        ref = ref * (nchw(xb) > 0)
    assert torch.equal(nchw(dx), rb(ref))


# ------------------------------------------------------------------ bilinear
@pytest.mark.parametrize("n,c,h,w", [(2, 64, 5, 7), (1, 4, 3, 3), (1, 8, 1, 6), (1, 16, 6, 1)])
def test_bilinear(ops, n, c, h, w):
    x = nhwc(torch.randn(n, c, h, w, device=dev()))
    y = ops.bilinear_fwd(x)
    xr = nchw(x).requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    assert rel_l2(nchw(y), ref.detach()) < BF16_OUT
    dy = nhwc(torch.randn(n, c, 2 * h, 2 * w, device=dev()))
    dx = torch.empty_like(x)
    ops.bilinear_bwd(dy, dx)
    ref.backward(nchw(dy))
    assert rel_l2(nchw(dx), xr.grad) < BF16_OUT
    m = nhwc(torch.randn(n, c, h, w, device=dev()))
    ops.bilinear_bwd(dy, dx, mask=m)
    assert rel_l2(nchw(dx), xr.grad * (nchw(m) > 0)) < BF16_OUT


# ------------------------------------------------------------------ batch norm
@pytest.mark.parametrize("n,c,h,w", [(4, 64, 9, 11), (2, 4, 16, 8), (3, 96, 5, 5), (2, 8, 3, 3)])
def test_batchnorm(ops, n, c, h, w):
    x = nhwc(F.relu(torch.randn(n, c, h, w, device=dev()) + 0.3))
    gamma = torch.rand(c, device=dev()) + 0.5
    beta = torch.randn(c, device=dev())
    rm, rv = torch.zeros(c, device=dev()), torch.ones(c, device=dev())
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y, mean, invstd = ops.bn_fwd_train(x, gamma, beta, rm, rv, 0.1, 1e-5)
    xr = nchw(x).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.batch_norm(xr, rm_ref, rv_ref, gr, br, training=True, momentum=0.1, eps=1e-5)
    assert rel_l2(nchw(y), ref.detach()) < BF16_OUT
    assert torch.allclose(rm, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)
    dy = nhwc(torch.randn(n, c, h, w, device=dev()))
    dx, dg, db = ops.bn_bwd(x, dy, gamma, mean, invstd, relu_mask=False)
    ref.backward(nchw(dy))
    assert rel_l2(nchw(dx), xr.grad) < BF16_OUT
    assert rel_l2(dg, gr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4
    dx2, _, _ = ops.bn_bwd(x, dy, gamma, mean, invstd, relu_mask=True)
    assert rel_l2(nchw(dx2), xr.grad * (nchw(x) > 0)) < BF16_OUT
    ye = ops.bn_fwd_eval(x, gamma, beta, rm, rv, 1e-5)
    refe = F.batch_norm(nchw(x), rm, rv, gamma, beta, training=False, eps=1e-5)
    assert rel_l2(nchw(ye), refe) < BF16_OUT


# ------------------------------------------------------------------ head (+ cross entropy)
@pytest.mark.parametrize("n,c,h,w,k,relu", [(2, 64, 9, 13, 2, False), (1, 64, 8, 8, 6, True), (2, 4, 5, 7, 3, False),
                                             (1, 128, 6, 6, 8, False), (1, 8, 4, 4, 1, True)])
def test_head_and_cross_entropy(ops, n, c, h, w, k, relu):
    x = nhwc(torch.randn(n, c, h, w, device=dev()))
    wt = torch.randn(k, c, device=dev()) * 0.2
    b = torch.randn(k, device=dev()) * 0.1
    y = torch.randint(0, k, (n, h, w), device=dev())
    y[0, 0, 0] = -100  # ignore_index
    logits = ops.head_fwd(x, wt, b, relu)
    xr = nchw(x).requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr.view(k, c, 1, 1), br)
    if relu:
        ref = F.relu(ref)
    assert rel_l2(logits, ref.detach()) < 1e-5
    # external-gradient backward
    dl = torch.randn_like(ref)
    dx, dw, db = ops.head_bwd(x, wt, b, relu, dl)
    ref.backward(dl, retain_graph=True)
    assert rel_l2(nchw(dx), xr.grad) < BF16_OUT
    assert rel_l2(dw.view(k, c), wr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4
    # fused cross entropy
    xr.grad = wr.grad = br.grad = None
    loss, state, lg = ops.head_ce_fwd(x, wt, b, relu, y, want_logits=True)
    ref_loss = F.cross_entropy(ref, y)
    assert abs(loss.item() - ref_loss.item()) < 1e-5 * max(1.0, abs(ref_loss.item()))
    assert rel_l2(lg, ref.detach()) < 1e-5
    gs = torch.tensor([0.5], device=dev())
    m = nhwc(torch.randn(n, c, h, w, device=dev()))
    dx, dw, db = ops.head_ce_bwd(x, wt, b, relu, y, state, grad_scale=gs, mask=m)
    (ref_loss * 0.5).backward()
    assert rel_l2(nchw(dx), xr.grad * (nchw(m) > 0)) < BF16_OUT
    assert rel_l2(dw.view(k, c), wr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4


# ------------------------------------------------------------------ convolution family
CONV_CASES = [
    # n, srcs(c...), cout, h, w, k, pad
    (2, (1,), 64, 20, 24, 3, 0),
    (1, (3,), 8, 11, 9, 3, 1),
    (2, (64,), 64, 14, 18, 3, 0),
    (1, (64, 64), 64, 12, 40, 3, 0),
    (1, (128,), 256, 9, 9, 3, 1),
    (2, (32, 16), 48, 10, 10, 3, 1),
    (1, (64, 4), 64, 8, 12, 3, 1),
    (1, (4,), 4, 7, 7, 3, 1),
    (2, (64,), 32, 9, 7, 1, 0),
    (1, (256,), 128, 6, 6, 1, 0),
    (3, (64,), 64, 40, 70, 3, 0),
    (1, (512,), 512, 10, 12, 3, 0),
    # enough pixel tiles and channels for CTA pairs (cta_group::2) in all three passes
    (2, (256,), 256, 30, 34, 3, 1),
    (1, (128, 128), 256, 26, 30, 3, 0),
    # first-layer kernels (1..4 input channels, cout % 16 == 0)
    (2, (3,), 32, 19, 23, 3, 1),
    (1, (2,), 16, 9, 30, 3, 0),
    (1, (4,), 128, 12, 13, 3, 1),
    (3, (1,), 64, 33, 41, 3, 1),
]


def _impls(ops):
    return [ops.IMPL_DIRECT, ops.IMPL_AUTO]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
@pytest.mark.parametrize("impl_name", ["direct", "auto"])
def test_conv_fwd_dgrad_wgrad(ops, case, impl_name):
    n, cs, cout, h, w, k, pad = case
    impl = ops.IMPL_DIRECT if impl_name == "direct" else ops.IMPL_AUTO
    cin = sum(cs)
    torch.manual_seed(hash(case) % 1000)
    # sources are windows of larger tensors (the cropped skip connection)
    bigs = [nhwc(torch.randn(n, c, h + 3, w + 5, device=dev())) for c in cs]
    srcs = [b_[:, 1:1 + h, 2:2 + w, :] for b_ in bigs]
    wt = torch.randn(cout, cin, k, k, device=dev()) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev()) * 0.1
    y = ops.conv_fwd(srcs, wt, b, pad, True, impl=impl)
    xs = [nchw(s).requires_grad_(True) for s in srcs]
    wq = rb(wt) if impl_name == "auto" else wt   # the tcgen05 path multiplies bf16 weights
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.relu(F.conv2d(torch.cat(xs, 1), wr, br, padding=pad))
    refq = F.relu(F.conv2d(torch.cat(xs, 1).detach(), wq, b, padding=pad))
    assert rel_l2(nchw(y), refq) < BF16_OUT
    # backward: dz = dy * (y > 0)
    dy = torch.randn_like(ref)
    dz = nhwc(dy * (ref.detach() > 0))
    dzr = nchw(dz)
    F.conv2d(torch.cat(xs, 1), wr, br, padding=pad).backward(dzr)
    dw, db = ops.conv_wgrad(dz, srcs, k, pad, impl=impl)
    assert rel_l2(dw, wr.grad) < F32_OUT
    assert rel_l2(db, br.grad) < 1e-4
    dsts = [torch.empty_like(b_) for b_ in bigs]
    dwin = [d[:, 1:1 + h, 2:2 + w, :] for d in dsts]
    masks = [nhwc(torch.randn(n, c, h + 3, w + 5, device=dev()))[:, 1:1 + h, 2:2 + w, :] for c in cs]
    # masks must be laid out like their destination: build them as windows of same-shaped tensors
    ops.conv_dgrad(dz, wt, pad, dwin, masks, impl=impl)
    refdx = torch.nn.grad.conv2d_input(torch.cat(xs, 1).shape, wq, dzr, padding=pad)
    o = 0
    for i, c in enumerate(cs):
        want = refdx[:, o:o + c] * (nchw(masks[i]) > 0)
        assert rel_l2(nchw(dwin[i]), want) < BF16_OUT, i
        o += c


CONVT_CASES = [(2, 64, 32, 7, 9), (1, 8, 4, 5, 5), (1, 128, 64, 12, 10), (1, 1024, 512, 4, 4), (2, 16, 16, 3, 6),
               (2, 256, 128, 24, 28), (1, 512, 256, 20, 22)]  # the last two: CTA pairs


@pytest.mark.parametrize("case", CONVT_CASES, ids=[str(c) for c in CONVT_CASES])
@pytest.mark.parametrize("impl_name", ["direct", "auto"])
def test_convt(ops, case, impl_name):
    n, cin, cout, h, w = case
    impl = ops.IMPL_DIRECT if impl_name == "direct" else ops.IMPL_AUTO
    x = nhwc(torch.randn(n, cin, h, w, device=dev()))
    wt = torch.randn(cin, cout, 2, 2, device=dev()) / (cin * 4) ** 0.5
    b = torch.randn(cout, device=dev()) * 0.1
    wq = rb(wt) if impl_name == "auto" else wt
    y = ops.convt_fwd(x, wt, b, impl=impl)
    xr = nchw(x).requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv_transpose2d(xr, wr, br, stride=2)
    assert rel_l2(nchw(y), F.conv_transpose2d(nchw(x), wq, b, stride=2)) < BF16_OUT
    dy = nhwc(torch.randn_like(ref))
    ref.backward(nchw(dy))
    dw, db = ops.convt_wgrad(x, dy, impl=impl)
    assert rel_l2(dw, wr.grad) < F32_OUT and rel_l2(db, br.grad) < 1e-4
    dx = torch.empty_like(x)
    m = nhwc(torch.randn(n, cin, h, w, device=dev()))
    ops.convt_dgrad(dy, wt, dx, mask=m, impl=impl)
    want = F.conv2d(nchw(dy), wq, stride=2) * (nchw(m) > 0)
    assert rel_l2(nchw(dx), want) < BF16_OUT


def test_relu_mask_and_channel_sum(ops):
    x = nhwc(torch.randn(2, 24, 5, 7, device=dev()))
    m = nhwc(torch.randn(2, 24, 5, 7, device=dev()))
    assert torch.equal(ops.relu_mask(x, m), torch.where(m.float() > 0, x, torch.zeros_like(x)))
    assert rel_l2(ops.channel_sum(x), x.float().sum((0, 1, 2))) < 1e-5


def test_errors_are_raised(ops):
    x = nhwc(torch.randn(1, 8, 4, 4, device=dev()))
    with pytest.raises(RuntimeError):
        ops.conv_fwd([x], torch.randn(8, 8, 3, 3, device=dev()), None, 0, True,
                     out=torch.empty(1, 3, 3, 8, dtype=torch.bfloat16, device=dev()))


@pytest.mark.parametrize("n,c,h,w,crop", [(2, 64, 12, 20, (2, 4, 6, 10)), (1, 8, 7, 9, (1, 1, 4, 6)), (1, 16, 10, 10, None)])
def test_maxpool_bwd_premasked_is_bit_identical(ops, n, c, h, w, crop):
    """b200unet_maxpool2x2_bwd_premasked (skip gradient masked by its producer, scattered term masked by [pooled > 0]) must
    give exactly the bytes of the masked form that reads the full-resolution ReLU mask."""
    torch.manual_seed(11)
    act = torch.relu(torch.randn(n, h, w, c, device="cuda")).to(torch.bfloat16)     # post-ReLU producer output (~half zeros)
    act[:, :2, :2] = 0                                                               # an all-zero window
    pooled, idx8 = ops.maxpool_fwd(act)
    g = torch.randn(n, h // 2, w // 2, c, device="cuda").to(torch.bfloat16)
    ref = torch.zeros(n, h, w, c, device="cuda", dtype=torch.bfloat16)
    new = torch.zeros_like(ref)
    if crop is not None:
        y0, x0, ch, cw = crop
        skip = torch.randn(n, ch, cw, c, device="cuda").to(torch.bfloat16)
        ref[:, y0:y0 + ch, x0:x0 + cw] = skip
        new[:, y0:y0 + ch, x0:x0 + cw] = torch.where(act[:, y0:y0 + ch, x0:x0 + cw] > 0, skip, torch.zeros_like(skip))
        ops.maxpool_bwd(g, idx8, ref, add=ref[:, y0:y0 + ch, x0:x0 + cw], add_y=y0, add_x=x0, mask=act)
        ops.maxpool_bwd(g, idx8, new, add=new[:, y0:y0 + ch, x0:x0 + cw], add_y=y0, add_x=x0, pooled=pooled)
    else:
        ops.maxpool_bwd(g, idx8, ref, mask=act)
        ops.maxpool_bwd(g, idx8, new, pooled=pooled)
    assert torch.equal(ref.view(torch.int16), new.view(torch.int16))
