"""CPU restatement of the reference's CVO kernel-Gramian loss.  TEST INFRASTRUCTURE ONLY: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product never does.

What it restates (plain PyTorch, fp32 or fp64, any device):
  sub_norm            geometry.py:13-25   SubNormFunction over the ABSENT extension `sub_norm_cuda_half_paral`
  kern_mat            geometry.py:47-136  the live (else) branch: exp(-d / (2 s^2)) with the 8.315e-3 cut-off
  cross_prod / cross_subtract  geometry.py:27-45  over the ABSENT extensions `cross_prod_cuda`, `cross_subtract_cuda`
  gramian             geometry.py:138-181
  inner_prod_from_gramians / calc_inner_prod / calc_loss_from_inner_prod   network_modules.py:1096-1189
  calc_w_v            network_modules.py:1052-1094

PARITY PINNING.  geometry.py's Python and the network_modules.py methods are pinned: oracle/make_golden_cvo.py runs
the UNMODIFIED reference source for them and tests/test_cvo_oracle.py compares (tests/golden/cvo/cvo_*.npz).  The three
CUDA extensions those functions call have NO source in the reference tree (only stale cp36/cp37
binaries are named in .MISSING_LARGE_BLOBS), so their semantics are restated from the call sites and are
"parity unpinned":
  * sub_norm(x1, x2)[b, i, j] = sum_c (x1[b, c, i] - x2[b, c, j])^2 — the SQUARED distance: kern_mat feeds it to
    exp(-d / (2 * dist_coef^2)) and derives the matching cut-off `thre_d = -2 * dist_coef^2 * log(thre_t)`
    (geometry.py:108-118), which is the squared-exponential kernel and d2 threshold of the CVO paper the file cites;
  * cross_prod(x1, x2)[b, i, j, :] = x1[b, :, i] x x2[b, :, j] and cross_subtract(...) = x1[b, :, i] - x2[b, :, j]
    ("B*N1*N2*3", network_modules.py:775-777).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

THRE_T = 8.315e-3  # geometry.py:108


def sub_norm(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """B*C*N1, B*C*N2 -> B*N1*N2 squared distances (see the header: restated from geometry.py:108-118)."""
    d = x1.unsqueeze(-1) - x2.unsqueeze(-2)  # B*C*N1*N2
    return (d * d).sum(dim=1)


def kern_mat(pcl_1: torch.Tensor, pcl_2: torch.Tensor, dist_coef: float = 1e-1) -> torch.Tensor:
    """geometry.py:47-136 (else branch, lines 104-121)."""
    pcl_diff = sub_norm(pcl_1.contiguous(), pcl_2.contiguous())
    pcl_diff_exp = torch.exp(-pcl_diff / (2 * dist_coef * dist_coef))
    return torch.where(pcl_diff_exp >= THRE_T, pcl_diff_exp, torch.zeros_like(pcl_diff_exp))


def cross_prod(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """B*3*N1, B*3*N2 -> B*N1*N2*3 (geometry.py:27-31, 39-41)."""
    a = x1.permute(0, 2, 1).unsqueeze(2)  # B*N1*1*3
    b = x2.permute(0, 2, 1).unsqueeze(1)  # B*1*N2*3
    a, b = torch.broadcast_tensors(a, b)
    return torch.cross(a, b, dim=-1)


def cross_subtract(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """B*3*N1, B*3*N2 -> B*N1*N2*3 (geometry.py:33-37, 43-45)."""
    return x1.permute(0, 2, 1).unsqueeze(2) - x2.permute(0, 2, 1).unsqueeze(1)


def gramian(fea_flat_1, fea_flat_2, norm_mode, kernalize, norm_dim, dist_coef=1e0):
    """geometry.py:138-181 (the device of the two zero scalars follows the inputs instead of .get_device())."""
    fea_norm_sum_1 = torch.zeros((), dtype=fea_flat_1.dtype, device=fea_flat_1.device)
    fea_norm_sum_2 = torch.zeros((), dtype=fea_flat_2.dtype, device=fea_flat_2.device)
    if norm_dim == 1:
        fea_norm_1 = torch.norm(fea_flat_1, dim=1, keepdim=True)
        fea_norm_2 = torch.norm(fea_flat_2, dim=1, keepdim=True)
    elif norm_dim == 2:
        fea_norm_1 = torch.mean(torch.abs(fea_flat_1), dim=2, keepdim=True)
        fea_norm_2 = torch.mean(torch.abs(fea_flat_2), dim=2, keepdim=True)
    if norm_mode:
        fea_flat_1 = torch.div(fea_flat_1, fea_norm_1)
        fea_flat_2 = torch.div(fea_flat_2, fea_norm_2)
        if norm_dim == 2:
            fea_norm_sum_1 = -torch.mean(torch.norm(fea_flat_1, dim=2))
            fea_norm_sum_2 = -torch.mean(torch.norm(fea_flat_2, dim=2))
    elif norm_dim == 1 or norm_dim == 2:
        fea_norm_sum_1 = torch.mean(fea_norm_1)
        fea_norm_sum_2 = torch.mean(fea_norm_2)
    if not kernalize:
        g = torch.matmul(fea_flat_1.transpose(1, 2), fea_flat_2)
    else:
        g = kern_mat(fea_flat_1, fea_flat_2, dist_coef=dist_coef)
    return g, fea_norm_sum_1 + fea_norm_sum_2


def inner_prod_from_gramians(gramian_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """network_modules.py:1151-1165: the element-wise product of the Gramians of every domain."""
    inn_prod = gramian_list[0]
    for g in gramian_list[1:]:
        inn_prod = inn_prod * g
    return inn_prod


def calc_inner_prod(gramian_list: Sequence[torch.Tensor], w_i: Optional[torch.Tensor] = None,
                    w_j: Optional[torch.Tensor] = None, normalize_over_pts: bool = False):
    """network_modules.py:1096-1149 for one pair ij: (per-point product matrix, its sum).  w_i / w_j are the B*1*N
    `feature_w` maps of weight_map_mode."""
    inn_prod = inner_prod_from_gramians(gramian_list)
    if w_i is not None:
        inn_prod = inn_prod * w_i.transpose(1, 2) * w_j
    total = torch.sum(inn_prod)
    if normalize_over_pts:
        ni_nj = gramian_list[0].shape[1] * gramian_list[0].shape[2]
        total, inn_prod = total / ni_nj, inn_prod / ni_nj
    return inn_prod, total


def calc_loss_from_inner_prod(inner_prods: Dict[Tuple[int, int], torch.Tensor], with_self_terms: bool = True):
    """network_modules.py:1167-1189."""
    losses = {"inner_prod": inner_prods[(0, 1)]}
    if with_self_terms:
        losses["inner_prod_0_0"] = inner_prods[(0, 0)]
        losses["inner_prod_1_1"] = inner_prods[(1, 1)]
        losses["func_dist"] = inner_prods[(0, 0)] + inner_prods[(1, 1)] - 2 * inner_prods[(0, 1)]
        losses["cos_sim"] = 1 - inner_prods[(0, 1)] / torch.sqrt(inner_prods[(0, 0)] * inner_prods[(1, 1)])
    return losses


def calc_w_v(inn_prod: torch.Tensor, cross_prod_geo: torch.Tensor, cross_sub_geo: torch.Tensor):
    """network_modules.py:1052-1094: se(3) gradient direction from the per-point product matrix (B*N1*N2)."""
    w = torch.stack([torch.sum(inn_prod * cross_prod_geo[..., k], dim=(1, 2)) for k in range(3)], dim=1)
    v = torch.stack([torch.sum(inn_prod * cross_sub_geo[..., k], dim=(1, 2)) for k in range(3)], dim=1)
    wv = torch.cat((w, v), dim=1)
    wv_norm = wv.norm(dim=1)
    if wv_norm < 1e-6:
        wv = torch.zeros((wv_norm.shape[0], 6), dtype=inn_prod.dtype, device=inn_prod.device)
    else:
        wv = wv / wv_norm
    return wv[:, :3], wv[:, 3:]


def cvo_inner_product(items_i: List[torch.Tensor], items_j: List[torch.Tensor], dist_coefs: Sequence[Optional[float]],
                      w_i=None, w_j=None, normalize_over_pts=False) -> torch.Tensor:
    """The whole chain calc_gramian -> calc_inner_prod for one pair (network_modules.py:995-1015, 1096-1149):
    dist_coefs[k] is the RBF scale of domain k, or None for the plain inner-product Gramian (`not kernalize`)."""
    gl = []
    for a, b, s in zip(items_i, items_j, dist_coefs):
        gl.append(torch.matmul(a.transpose(1, 2), b) if s is None else kern_mat(a, b, dist_coef=s))
    return calc_inner_prod(gl, w_i, w_j, normalize_over_pts)[1]


def thre_d(dist_coef: float) -> float:
    """geometry.py:109: the squared-distance cut-off equivalent to the 8.315e-3 value cut-off."""
    return -2.0 * dist_coef * dist_coef * math.log(THRE_T)
