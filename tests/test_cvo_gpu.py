"""CVO kernel-Gramian loss on the GPU through the C ABI (b200unet_cvo_*) against
 - the golden vectors produced by the UNMODIFIED reference source (oracle/make_golden_cvo.py), and
 - the CPU oracle (oracle/cvo_oracle.py, fp64) on seeded inputs, up to the reference's own maximum size
   (N = 96 * 128 = 12 288 points, options.py:109-110) through properties the oracle cannot reach in seconds.
fp32 arithmetic; tolerances: values 2e-5 relative to the largest entry, sums and gradients 1e-4 relative (rel-L2)."""
import os

import numpy as np
import pytest
import torch

from oracle import cvo_oracle as CO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cvo")


def _g(a, grad=False):
    return torch.from_numpy(np.asarray(a)).to("cuda", torch.float32).requires_grad_(grad)


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _support_agrees(k, want, x1, x2, coef):
    """the cut-off may differ only for pairs whose value is within fp32 rounding of 8.315e-3"""
    diff = (k.cpu() == 0) != (want.cpu() == 0)
    if not diff.any():
        return True
    exact = CO.kern_mat(x1.double().cpu(), x2.double().cpu(), coef)
    near = (torch.exp(-CO.sub_norm(x1.double().cpu(), x2.double().cpu()) / (2 * coef * coef)) - CO.THRE_T).abs() < 1e-6
    return bool((near | ~diff).all()) and exact is not None


@pytest.mark.parametrize("tag", ["xyz", "img", "feat"])
def test_kern_mat_against_reference_golden(tag):
    from b200unet import cvo
    z = np.load(os.path.join(GOLD, "cvo_kern_mat.npz"))
    x1, x2, coef = _g(z[f"{tag}_x1"], True), _g(z[f"{tag}_x2"], True), float(z[f"{tag}_coef"])
    k = cvo.kern_mat(x1, x2, dist_coef=coef)
    want = torch.from_numpy(z[f"{tag}_k"])
    assert _support_agrees(k.detach(), want, x1.detach(), x2.detach(), coef)
    same = (k.detach().cpu() == 0) == (want == 0)
    assert float((k.detach().cpu().double() - want)[same].abs().max()) < 2e-5
    (k * _g(z[f"{tag}_dy"])).sum().backward()
    if bool(same.all()):
        assert _rel(x1.grad, torch.from_numpy(z[f"{tag}_dx1"])) < 1e-4
        assert _rel(x2.grad, torch.from_numpy(z[f"{tag}_dx2"])) < 1e-4


def test_sub_norm_forward_backward_vs_oracle():
    from b200unet import cvo
    torch.manual_seed(3)
    for (b, c, n1, n2) in [(2, 3, 70, 41), (1, 5, 129, 257), (1, 16, 33, 64), (1, 1, 5, 3)]:
        x1 = torch.randn(b, c, n1, device="cuda", requires_grad=True)
        x2 = torch.randn(b, c, n2, device="cuda", requires_grad=True)
        dy = torch.randn(b, n1, n2, device="cuda")
        d = cvo.sub_norm(x1, x2)
        (d * dy).sum().backward()
        o1, o2 = x1.detach().double().cpu().requires_grad_(True), x2.detach().double().cpu().requires_grad_(True)
        do = CO.sub_norm(o1, o2)
        (do * dy.double().cpu()).sum().backward()
        assert _rel(d, do) < 1e-6
        assert _rel(x1.grad, o1.grad) < 1e-5 and _rel(x2.grad, o2.grad) < 1e-5


def test_cross_prod_and_subtract_golden():
    from b200unet import cvo
    z = np.load(os.path.join(GOLD, "cvo_cross.npz"))
    x1, x2 = _g(z["x1"]), _g(z["x2"])
    assert torch.allclose(cvo.cross_prod(x1, x2).cpu().double(), torch.from_numpy(z["cross_prod"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(cvo.cross_subtract(x1, x2).cpu().double(), torch.from_numpy(z["cross_subtract"]), rtol=1e-6, atol=2e-6)


@pytest.mark.parametrize("name", ["rbf", "dot_weighted"])
def test_fused_loss_chain_against_reference_golden(name):
    """calc_gramian -> calc_inner_prod -> calc_loss_from_inner_prod, values and the gradient of func_dist w.r.t. every input"""
    from b200unet import cvo
    z = np.load(os.path.join(GOLD, f"cvo_loss_{name}.npz"))
    kern, wmap, norm = bool(z["kernalize"]), bool(z["weight_map"]), bool(z["normalize"])
    items = ["xyz", "img", "feature"]
    coef = {"xyz": float(z["coef_xyz"]), "img": float(z["coef_img"]), "feature": float(z["coef_feature"]) if kern else None}
    f = [{k: _g(z[f"f{i}_{k}"], True) for k in items + ["feature_w"]} for i in range(2)]
    losses = cvo.cvo_losses(f, items, coef, with_self_terms=True, weight_key="feature_w" if wmap else None,
                            normalize_over_pts=norm)
    for k, v in losses.items():
        want = float(z[f"loss_{k}"])
        assert abs(float(v) - want) <= 1e-4 * max(abs(want), 1e-3), (k, float(v), want)
    losses["func_dist"].backward()
    for i in range(2):
        for k in items + (["feature_w"] if wmap else []):
            assert _rel(f[i][k].grad, torch.from_numpy(z[f"f{i}_{k}_grad_func_dist"])) < 2e-4, (i, k)
    w, v = cvo.calc_w_v([f[0][k].detach() for k in items], [f[1][k].detach() for k in items], [coef[k] for k in items], 0)
    assert _rel(w, torch.from_numpy(z["w"])) < 1e-4 and _rel(v, torch.from_numpy(z["v"])) < 1e-4


@pytest.mark.parametrize("shape", [(1, 300, 211), (2, 129, 640), (1, 7, 5), (3, 128, 128)])
def test_fused_inner_product_vs_oracle(shape):
    from b200unet import cvo
    b, n1, n2 = shape
    g = torch.Generator().manual_seed(n1 * 1000 + n2)
    mk = lambda c, n, s: (torch.randn(b, c, n, generator=g) * s).cuda().requires_grad_(True)
    xi = [mk(3, n1, 0.3), mk(5, n1, 0.5), mk(6, n1, 0.1)]
    xj = [mk(3, n2, 0.3), mk(5, n2, 0.5), mk(6, n2, 0.1)]
    wi = (torch.rand(b, 1, n1, generator=g) + 0.5).cuda().requires_grad_(True)
    wj = (torch.rand(b, 1, n2, generator=g) + 0.5).cuda().requires_grad_(True)
    coefs = [0.2, 0.5, 0.1]
    out = cvo.inner_product(xi, xj, coefs, wi, wj)
    out.backward()
    oi = [t.detach().double().cpu().requires_grad_(True) for t in xi + [wi]]
    oj = [t.detach().double().cpu().requires_grad_(True) for t in xj + [wj]]
    want = CO.cvo_inner_product(oi[:3], oj[:3], coefs, oi[3], oj[3])
    want.backward()
    assert abs(float(out) - float(want)) <= 1e-4 * abs(float(want)) + 1e-6
    for got, o in zip(xi + [wi] + xj + [wj], oi + oj):
        assert _rel(got.grad, o.grad) < 2e-4


def test_single_domain_and_unweighted_paths():
    from b200unet import cvo
    torch.manual_seed(5)
    x1 = (torch.randn(1, 3, 513, device="cuda") * 0.3).requires_grad_(True)
    x2 = (torch.randn(1, 3, 400, device="cuda") * 0.3).requires_grad_(True)
    out = cvo.inner_product([x1], [x2], [0.2])
    # the fused sum equals the sum of the materialised matrix (same library, different kernels)
    assert abs(float(out) - float(cvo.kern_mat(x1, x2, 0.2).double().sum())) <= 1e-4 * float(out)
    out.backward()
    o1, o2 = x1.detach().double().cpu().requires_grad_(True), x2.detach().double().cpu().requires_grad_(True)
    CO.kern_mat(o1, o2, 0.2).sum().backward()
    assert _rel(x1.grad, o1.grad) < 2e-4 and _rel(x2.grad, o2.grad) < 2e-4


def test_full_size_properties_12288_points():
    """N = 96 x 128 points (options.py:109-110), the size at which the reference stores 604 MB per Gramian: symmetry
    <f,g> = <g,f>, Cauchy-Schwarz <f,g>^2 <= <f,f><g,g>, <f,f> >= N (the diagonal is exactly 1), row-block additivity,
    and agreement with the materialised kern_mat product."""
    from b200unet import cvo
    n = 96 * 128
    g = torch.Generator().manual_seed(7)
    grid = torch.stack(torch.meshgrid(torch.linspace(-1, 1, 96), torch.linspace(-1.3, 1.3, 128), indexing="ij")).reshape(1, 2, n)
    depth = 1.5 + 0.3 * torch.rand(1, 1, n, generator=g)
    xyz1 = torch.cat([grid * depth, depth], 1).cuda()
    xyz2 = (xyz1.cpu() + 0.02 * torch.randn(1, 3, n, generator=g)).cuda()
    img1, img2 = torch.rand(1, 5, n, generator=g).cuda(), torch.rand(1, 5, n, generator=g).cuda()
    fe1, fe2 = (torch.randn(1, 6, n, generator=g) * 0.1).cuda(), (torch.randn(1, 6, n, generator=g) * 0.1).cuda()
    coefs = [0.2, 0.5, 0.1]
    f01 = float(cvo.inner_product([xyz1, img1, fe1], [xyz2, img2, fe2], coefs))
    f10 = float(cvo.inner_product([xyz2, img2, fe2], [xyz1, img1, fe1], coefs))
    f00 = float(cvo.inner_product([xyz1, img1, fe1], [xyz1, img1, fe1], coefs))
    f11 = float(cvo.inner_product([xyz2, img2, fe2], [xyz2, img2, fe2], coefs))
    assert f01 > 0 and abs(f01 - f10) <= 2e-5 * f01
    assert f00 >= n * (1 - 1e-6) and f11 >= n * (1 - 1e-6)
    assert f01 * f01 <= f00 * f11 * (1 + 1e-5)
    half = n // 2
    parts = sum(float(cvo.inner_product([t[:, :, s] for t in (xyz1, img1, fe1)], [xyz2, img2, fe2], coefs))
                for s in (slice(0, half), slice(half, n)))
    assert abs(parts - f01) <= 2e-5 * f01
    mat = cvo.kern_mat(xyz1, xyz2, 0.2) * cvo.kern_mat(img1, img2, 0.5) * cvo.kern_mat(fe1, fe2, 0.1)
    assert abs(float(mat.double().sum()) - f01) <= 1e-4 * f01
    w, v = cvo.calc_w_v([xyz1, img1, fe1], [xyz2, img2, fe2], coefs, 0)
    wv_ref = torch.cat([(mat.unsqueeze(-1) * cvo.cross_prod(xyz1, xyz2)).double().sum((1, 2)),
                        (mat.unsqueeze(-1) * cvo.cross_subtract(xyz1, xyz2)).double().sum((1, 2))], 1)
    wv_ref = wv_ref / wv_ref.norm(dim=1, keepdim=True)
    assert _rel(torch.cat([w, v], 1), wv_ref) < 1e-3


def test_gramian_drop_in_signature():
    from b200unet import cvo
    torch.manual_seed(9)
    f1, f2 = torch.rand(1, 4, 90, device="cuda") + 0.1, torch.rand(1, 4, 70, device="cuda") + 0.1
    for norm_mode, kern, nd in [(True, True, 1), (False, False, 2), (True, False, 2), (False, True, 0)]:
        g, s = cvo.gramian(f1, f2, norm_mode, kern, nd, dist_coef=0.5)
        go, so = CO.gramian(f1.double().cpu(), f2.double().cpu(), norm_mode, kern, nd, dist_coef=0.5)
        same = (g.cpu() == 0) == (go == 0)
        assert float(same.double().mean()) > 0.999
        assert float((g.cpu().double() - go)[same].abs().max()) < 1e-4 and abs(float(s) - float(so)) < 1e-5


def test_errors():
    from b200unet import cvo
    x = torch.randn(1, 3, 8, device="cuda")
    with pytest.raises(RuntimeError):
        cvo.kern_mat(x.cpu(), x.cpu(), 0.1)
    with pytest.raises(ValueError):
        cvo.inner_product([x], [torch.randn(1, 4, 8, device="cuda")], [0.1])
    with pytest.raises(ValueError):
        cvo.inner_product([x, x], [x, x], [None, None])
    with pytest.raises(ValueError):
        cvo.inner_product([torch.randn(1, 17, 8, device="cuda")], [torch.randn(1, 17, 8, device="cuda")], [0.1])
    with pytest.raises(ValueError):
        cvo.cross_prod(torch.randn(1, 4, 8, device="cuda"), torch.randn(1, 4, 8, device="cuda"))
