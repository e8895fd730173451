"""oracle/cvo_oracle.py against the golden vectors the UNMODIFIED reference source produced (oracle/make_golden_cvo.py):
geometry.py:47-136 kern_mat (+ its autograd through SubNormFunction), geometry.py:27-45, network_modules.py:1052-1189."""
import os

import numpy as np
import pytest
import torch

from oracle import cvo_oracle as CO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cvo")


def _t(a, grad=False):
    return torch.from_numpy(np.asarray(a)).clone().requires_grad_(grad)


def test_fixtures_present():
    for f in ("cvo_kern_mat.npz", "cvo_cross.npz", "cvo_loss_rbf.npz", "cvo_loss_dot_weighted.npz"):
        assert os.path.isfile(os.path.join(GOLD, f)), f


@pytest.mark.parametrize("tag", ["xyz", "img", "feat"])
def test_kern_mat_forward_backward(tag):
    z = np.load(os.path.join(GOLD, "cvo_kern_mat.npz"))
    x1, x2 = _t(z[f"{tag}_x1"], True), _t(z[f"{tag}_x2"], True)
    k = CO.kern_mat(x1, x2, dist_coef=float(z[f"{tag}_coef"]))
    want = _t(z[f"{tag}_k"])
    assert torch.equal(k == 0, want == 0)                       # the 8.315e-3 cut-off falls on the same pairs
    assert 0.02 < float((want > 0).double().mean()) < 0.9       # and the fixture exercises both sides of it
    assert torch.allclose(k, want, rtol=1e-12, atol=1e-14)
    (k * _t(z[f"{tag}_dy"])).sum().backward()
    assert torch.allclose(x1.grad, _t(z[f"{tag}_dx1"]), rtol=1e-10, atol=1e-12)
    assert torch.allclose(x2.grad, _t(z[f"{tag}_dx2"]), rtol=1e-10, atol=1e-12)


def test_cross_prod_and_subtract():
    z = np.load(os.path.join(GOLD, "cvo_cross.npz"))
    x1, x2 = _t(z["x1"]), _t(z["x2"])
    assert torch.allclose(CO.cross_prod(x1, x2), _t(z["cross_prod"]), rtol=1e-13, atol=1e-15)
    assert torch.allclose(CO.cross_subtract(x1, x2), _t(z["cross_subtract"]), rtol=0, atol=0)


@pytest.mark.parametrize("name", ["rbf", "dot_weighted"])
def test_loss_chain(name):
    z = np.load(os.path.join(GOLD, f"cvo_loss_{name}.npz"))
    kern, wmap, norm = bool(z["kernalize"]), bool(z["weight_map"]), bool(z["normalize"])
    items = ["xyz", "img", "feature"]
    coefs = [float(z["coef_xyz"]), float(z["coef_img"]), float(z["coef_feature"]) if kern else None]
    f = [{k: _t(z[f"f{i}_{k}"], True) for k in items + ["feature_w"]} for i in range(2)]
    ip = {}
    for (i, j) in [(0, 0), (1, 1), (0, 1)]:
        ip[(i, j)] = CO.cvo_inner_product([f[i][k] for k in items], [f[j][k] for k in items], coefs,
                                          f[i]["feature_w"] if wmap else None, f[j]["feature_w"] if wmap else None, norm)
    losses = CO.calc_loss_from_inner_prod(ip)
    for k, v in losses.items():
        assert torch.allclose(v, _t(z[f"loss_{k}"]), rtol=1e-11, atol=1e-13), k
    losses["func_dist"].backward()
    for i in range(2):
        for k in items + (["feature_w"] if wmap else []):
            assert torch.allclose(f[i][k].grad, _t(z[f"f{i}_{k}_grad_func_dist"]), rtol=1e-9, atol=1e-11), (i, k)
    # per-point product matrix and the se(3) direction of calc_w_v
    gl = [torch.matmul(f[0][k].transpose(1, 2), f[1][k]) if s is None else CO.kern_mat(f[0][k], f[1][k], s)
          for k, s in zip(items, coefs)]
    perp, _ = CO.calc_inner_prod(gl, f[0]["feature_w"] if wmap else None, f[1]["feature_w"] if wmap else None, False)
    assert torch.allclose(perp, _t(z["perp_01"]), rtol=1e-11, atol=1e-14)
    w, v = CO.calc_w_v(CO.inner_prod_from_gramians(gl).detach(), CO.cross_prod(f[0]["xyz"], f[1]["xyz"]).detach(),
                       CO.cross_subtract(f[0]["xyz"], f[1]["xyz"]).detach())
    assert torch.allclose(w, _t(z["w"]), rtol=1e-10, atol=1e-12) and torch.allclose(v, _t(z["v"]), rtol=1e-10, atol=1e-12)


def test_cutoff_is_the_squared_distance_threshold():
    # geometry.py:108-109: value cut-off 8.315e-3 <=> squared distance below -2 s^2 ln(8.315e-3)
    s = 0.2
    d = torch.tensor([CO.thre_d(s) * 0.999, CO.thre_d(s) * 1.001], dtype=torch.float64)
    x1 = torch.zeros(1, 1, 1, dtype=torch.float64)
    x2 = d.sqrt().reshape(1, 1, 2)
    k = CO.kern_mat(x1, x2, s)
    assert k[0, 0, 0] > 0 and k[0, 0, 1] == 0
