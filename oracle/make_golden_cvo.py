"""Generates tests/golden/cvo/cvo_*.npz by running the UNMODIFIED reference source of the CVO Gramian loss (this container
only; /root/reference is read, never copied).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_cvo.py

* `geometry.py` is imported as it is.  The three CUDA extensions it imports at line 4 have no source in the reference
  tree, so the interpreter is given pure-PyTorch stand-ins for them (defined below — the one unpinned assumption, see
  oracle/cvo_oracle.py's header), and a stub `dataloader` module for the one name geometry.py takes from it (the real
  module needs skimage / torchsnooper, and the name is not used on this path).
* The `innerProdLoss` methods of `network_modules.py` (which cannot be imported: open3d, a forked
  segmentation_models_pytorch, ...) are taken out of the file with `ast` and compiled unmodified.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("UNET_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "cvo")


def _install_stubs():
    def sub_norm_forward(x1, x2):
        d = x1.unsqueeze(-1) - x2.unsqueeze(-2)
        return (d * d).sum(dim=1)

    def sub_norm_backward(dy, x1, x2):
        d = x1.unsqueeze(-1) - x2.unsqueeze(-2)           # B*C*N1*N2
        g = 2 * d * dy.unsqueeze(1)
        return g.sum(dim=3), -g.sum(dim=2)

    def cross_prod_forward(x1, x2):
        a = x1.permute(0, 2, 1).unsqueeze(2)
        b = x2.permute(0, 2, 1).unsqueeze(1)
        a, b = torch.broadcast_tensors(a, b)
        return torch.cross(a, b, dim=-1)

    def cross_subtract_forward(x1, x2):
        return x1.permute(0, 2, 1).unsqueeze(2) - x2.permute(0, 2, 1).unsqueeze(1)

    m = types.ModuleType("sub_norm_cuda_half_paral")
    m.forward, m.backward = sub_norm_forward, sub_norm_backward
    sys.modules["sub_norm_cuda_half_paral"] = m
    m = types.ModuleType("cross_prod_cuda")
    m.forward = cross_prod_forward
    sys.modules["cross_prod_cuda"] = m
    m = types.ModuleType("cross_subtract_cuda")
    m.forward = cross_subtract_forward
    sys.modules["cross_subtract_cuda"] = m
    m = types.ModuleType("dataloader")
    m.pose_from_euler_t_Tensor = None
    sys.modules["dataloader"] = m


def load_reference_geometry():
    _install_stubs()
    sys.path.insert(0, REF)
    try:
        import geometry  # noqa: the reference's own file
    finally:
        sys.path.remove(REF)
    return geometry


def load_reference_loss_methods(names):
    """Unmodified source of the named innerProdLoss methods -> plain functions taking `self` first."""
    src = open(os.path.join(REF, "network_modules.py")).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "innerProdLoss":
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in names:
                    mod = ast.Module(body=[f], type_ignores=[])
                    ns = {"torch": torch, "np": np}
                    exec(compile(mod, os.path.join(REF, "network_modules.py"), "exec"), ns)
                    out[f.name] = ns[f.name]
    missing = set(names) - set(out)
    if missing:
        raise RuntimeError(f"not found in network_modules.py: {missing}")
    return out


def _cloud(g, b, c, n, spread):
    return (torch.randn(b, c, n, generator=g, dtype=torch.float64) * spread)


def main():
    os.makedirs(OUT, exist_ok=True)
    geometry = load_reference_geometry()
    fns = load_reference_loss_methods(["inner_prod_from_gramians", "calc_inner_prod", "calc_loss_from_inner_prod",
                                       "calc_w_v"])
    g = torch.Generator().manual_seed(20191023)

    # ---- case 1: kern_mat forward + backward (through SubNormFunction), three channel counts / scales
    rec = {}
    for tag, (b, c, n1, n2, spread, coef) in {"xyz": (2, 3, 37, 53, 0.4, 0.2), "img": (1, 5, 64, 48, 0.6, 0.5),
                                              "feat": (1, 6, 45, 45, 0.15, 0.1)}.items():
        x1 = _cloud(g, b, c, n1, spread).requires_grad_(True)
        x2 = _cloud(g, b, c, n2, spread).requires_grad_(True)
        k = geometry.kern_mat(x1, x2, dist_coef=coef)
        dy = torch.randn(k.shape, generator=g, dtype=torch.float64)
        (k * dy).sum().backward()
        rec.update({f"{tag}_x1": x1.detach().numpy(), f"{tag}_x2": x2.detach().numpy(), f"{tag}_coef": np.float64(coef),
                    f"{tag}_k": k.detach().numpy(), f"{tag}_dy": dy.numpy(), f"{tag}_dx1": x1.grad.numpy(),
                    f"{tag}_dx2": x2.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "cvo_kern_mat.npz"), **rec)

    # ---- case 2: cross_prod / cross_subtract
    x1, x2 = _cloud(g, 2, 3, 19, 1.0), _cloud(g, 2, 3, 23, 1.0)
    np.savez_compressed(os.path.join(OUT, "cvo_cross.npz"), x1=x1.numpy(), x2=x2.numpy(),
                        cross_prod=geometry.cross_prod(x1, x2).numpy(), cross_subtract=geometry.cross_subtract(x1, x2).numpy())

    # ---- case 3: the loss chain of network_modules.py (calc_gramian's kern_mat / matmul choice, calc_inner_prod,
    #      calc_loss_from_inner_prod, calc_w_v) on two frames, kernalized and plain-inner-product features,
    #      with and without weight maps / point normalisation
    for name, (kernalize, weight_map, normalize) in {"rbf": (True, False, False), "dot_weighted": (False, True, True)}.items():
        n = (57, 49)
        coefs = {"xyz": 0.2, "img": 0.5, "feature": 0.1}
        flat = []
        for i in range(2):
            feat = _cloud(g, 1, 4, n[i], 0.12)
            if not kernalize:
                feat = feat.abs()  # non_neg features (options.py:91: kernalize = not non_neg)
            flat.append({"xyz": _cloud(g, 1, 3, n[i], 0.35).requires_grad_(True),
                         "img": _cloud(g, 1, 5, n[i], 0.5).requires_grad_(True),
                         "feature": feat.requires_grad_(True),
                         "feature_w": (torch.rand(1, 1, n[i], generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)})
        gramians = {"xyz": {}, "img": {}, "feature": {}}
        list_of_ij = [(0, 0), (1, 1), (0, 1)]
        for item in gramians:
            for (i, j) in list_of_ij:
                if item == "feature" and not kernalize:     # network_modules.py:1012-1013
                    gramians[item][(i, j)] = torch.matmul(flat[i][item].transpose(1, 2), flat[j][item])
                else:                                        # network_modules.py:1015
                    gramians[item][(i, j)] = geometry.kern_mat(flat[i][item], flat[j][item], dist_coef=coefs[item])
        self = types.SimpleNamespace()
        self.opt = types.SimpleNamespace(opt_unet=types.SimpleNamespace(weight_map_mode=weight_map),
                                         normalize_inprod_over_pts=False, min_dist_mode=True, diff_mode=False,
                                         self_sparse_mode=False)
        self.inner_prod_from_gramians = lambda gr, ij, items=None: fns["inner_prod_from_gramians"](self, gr, ij, items)
        inner_prods = {}
        perp = fns["calc_inner_prod"](self, inner_prods, gramians, flat, ["xyz", "img", "feature"], list_of_ij)
        if normalize:
            # network_modules.py:1144-1147 reads `gramian_list`, a name that only exists in commented-out code, so the
            # reference raises NameError with normalize_inprod_over_pts=True; the stated intent (divide by N_i * N_j)
            # is applied here by hand and documented as such
            for (i, j) in list_of_ij:
                inner_prods[(i, j)] = inner_prods[(i, j)] / (n[i] * n[j])
        losses = fns["calc_loss_from_inner_prod"](self, inner_prods)
        losses["func_dist"].backward(retain_graph=True)
        rec = {"kernalize": np.bool_(kernalize), "weight_map": np.bool_(weight_map), "normalize": np.bool_(normalize),
               "coef_xyz": 0.2, "coef_img": 0.5, "coef_feature": 0.1}
        for i in range(2):
            for k, v in flat[i].items():
                rec[f"f{i}_{k}"] = v.detach().numpy()
                rec[f"f{i}_{k}_grad_func_dist"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        for k, v in losses.items():
            rec[f"loss_{k}"] = v.detach().numpy()
        rec["perp_01"] = perp[(0, 1)].detach().numpy()
        # calc_w_v on the (0,1) pair (network_modules.py:775-783)
        cp = geometry.cross_prod(flat[0]["xyz"].detach(), flat[1]["xyz"].detach())
        cs = geometry.cross_subtract(flat[0]["xyz"].detach(), flat[1]["xyz"].detach())
        gr_detached = {it: {(0, 1): gramians[it][(0, 1)].detach()} for it in gramians}
        w, v = fns["calc_w_v"](self, gr_detached, cp, cs, ["xyz", "img", "feature"], None)
        rec["w"], rec["v"] = w.numpy(), v.numpy()
        np.savez_compressed(os.path.join(OUT, f"cvo_loss_{name}.npz"), **rec)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("cvo_")))


if __name__ == "__main__":
    main()
