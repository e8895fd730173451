"""CVO kernel-Gramian loss on the GPU — the consumer of the U-Net's features in the
reference's own training loop.

Drop-in names (same arguments and meaning as the reference):
    sub_norm(x1, x2)                      geometry.py:13-25    SubNormFunction.apply
    kern_mat(pcl_1, pcl_2, dist_coef)     geometry.py:47-136
    cross_prod / cross_subtract           geometry.py:39-45
    gramian(f1, f2, norm_mode, kernalize, norm_dim, dist_coef)   geometry.py:138-181
These return the same B*N1*N2 tensors the reference materialises (autograd included), computed by one CUDA kernel each.

The product path is the FUSED form: `inner_product` evaluates calc_gramian + calc_inner_prod
(network_modules.py:995-1015, 1096-1149) for one frame pair — distance, cut-off, exponential and the product over the
domains — inside one kernel and keeps per-point sums only; its backward recomputes the pair weights.  `cvo_losses`
is calc_loss_from_inner_prod (network_modules.py:1167-1189) on top of it, `calc_w_v` the se(3) direction
(network_modules.py:1052-1094).  All compute goes through libb200unet.so (include/b200unet.h, b200unet_cvo_*); there
is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import stream_ptr as _stream

THRE_T = 8.315e-3  # geometry.py:108


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"b200unet.cvo.{what}: tensors must be on CUDA (there is no CPU path)")
    if t.dim() != 3:
        raise ValueError(f"b200unet.cvo.{what}: expected a B*C*N tensor, got shape {tuple(t.shape)}")
    return t.detach().to(torch.float32).contiguous()


def _workspace(b: int, n1: int, n2: int, c: int, device) -> torch.Tensor:
    nbytes = _lib.load().b200unet_cvo_workspace_bytes(b, n1, n2, c)
    return torch.empty(nbytes // 4, dtype=torch.float32, device=device)


def _pair_shapes(x1, x2, what):
    if x1.shape[0] != x2.shape[0] or x1.shape[1] != x2.shape[1]:
        raise ValueError(f"b200unet.cvo.{what}: batch / channel mismatch {tuple(x1.shape)} vs {tuple(x2.shape)}")
    return x1.shape[0], x1.shape[1], x1.shape[2], x2.shape[2]


class _MatFunction(torch.autograd.Function):
    """sub_norm (dist_coef None) or kern_mat as one materialising kernel, backward by recomputation."""

    @staticmethod
    def forward(ctx, x1, x2, dist_coef):
        what = "sub_norm" if dist_coef is None else "kern_mat"
        a, b_ = _f32c(x1, what), _f32c(x2, what)
        B, c, n1, n2 = _pair_shapes(a, b_, what)
        out = torch.empty((B, n1, n2), dtype=torch.float32, device=a.device)
        lib, st = _lib.load(), _stream()
        if dist_coef is None:
            _lib.check(lib.b200unet_cvo_sub_norm_fwd(a.data_ptr(), b_.data_ptr(), B, c, n1, n2, out.data_ptr(), st), what)
        else:
            _lib.check(lib.b200unet_cvo_kern_mat_fwd(a.data_ptr(), b_.data_ptr(), B, c, n1, n2, float(dist_coef),
                                                     out.data_ptr(), st), what)
        ctx.save_for_backward(a, b_)
        ctx.dist_coef, ctx.dtypes = dist_coef, (x1.dtype, x2.dtype)
        return out.to(x1.dtype)

    @staticmethod
    def backward(ctx, dy):
        a, b_ = ctx.saved_tensors
        B, c, n1, n2 = _pair_shapes(a, b_, "backward")
        if c > 16:
            raise RuntimeError("b200unet.cvo: the backward of sub_norm / kern_mat supports up to 16 channels")
        dy = dy.detach().to(torch.float32).contiguous()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx1 = torch.empty_like(a) if need1 else None
        dx2 = torch.empty_like(b_) if need2 else None
        ws = _workspace(B, n1, n2, c, a.device)
        lib, st = _lib.load(), _stream()
        p1 = dx1.data_ptr() if need1 else None
        p2 = dx2.data_ptr() if need2 else None
        if ctx.dist_coef is None:
            _lib.check(lib.b200unet_cvo_sub_norm_bwd(dy.data_ptr(), a.data_ptr(), b_.data_ptr(), B, c, n1, n2, ws.data_ptr(),
                                                     p1, p2, st), "sub_norm backward")
        else:
            _lib.check(lib.b200unet_cvo_kern_mat_bwd(dy.data_ptr(), a.data_ptr(), b_.data_ptr(), B, c, n1, n2,
                                                     float(ctx.dist_coef), ws.data_ptr(), p1, p2, st), "kern_mat backward")
        return (dx1.to(ctx.dtypes[0]) if need1 else None, dx2.to(ctx.dtypes[1]) if need2 else None, None)


def sub_norm(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """B*C*N1, B*C*N2 -> B*N1*N2 squared distances (SubNormFunction.apply, geometry.py:13-25)."""
    return _MatFunction.apply(x1, x2, None)


def kern_mat(pcl_1: torch.Tensor, pcl_2: torch.Tensor, dist_coef: float = 1e-1) -> torch.Tensor:
    """geometry.py:47-136: exp(-|x1_i - x2_j|^2 / (2 dist_coef^2)), zero where the value is below 8.315e-3."""
    return _MatFunction.apply(pcl_1, pcl_2, float(dist_coef))


def _cross(x1, x2, subtract, what):
    a, b_ = _f32c(x1, what), _f32c(x2, what)
    B, c, n1, n2 = _pair_shapes(a, b_, what)
    if c != 3:
        raise ValueError(f"b200unet.cvo.{what}: 3-channel point sets expected, got {c}")
    out = torch.empty((B, n1, n2, 3), dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().b200unet_cvo_cross_fwd(a.data_ptr(), b_.data_ptr(), B, n1, n2, int(subtract), out.data_ptr(),
                                                  _stream()), what)
    return out.to(x1.dtype)


def cross_prod(pcl_1: torch.Tensor, pcl_2: torch.Tensor) -> torch.Tensor:
    """geometry.py:39-41 (forward only, as CrossProdFunction): B*N1*N2*3 cross products."""
    return _cross(pcl_1, pcl_2, False, "cross_prod")


def cross_subtract(pcl_1: torch.Tensor, pcl_2: torch.Tensor) -> torch.Tensor:
    """geometry.py:43-45 (forward only): B*N1*N2*3 differences."""
    return _cross(pcl_1, pcl_2, True, "cross_subtract")


def _feature_norm(f: torch.Tensor, norm_dim: int):
    """The per-pixel L2 norm over channels (norm_dim 1) or the per-channel mean |.| over pixels (norm_dim 2)."""
    if norm_dim == 1:
        return f.norm(dim=1, keepdim=True)
    if norm_dim == 2:
        return f.abs().mean(dim=2, keepdim=True)
    return None


def gramian(fea_flat_1, fea_flat_2, norm_mode, kernalize, norm_dim, dist_coef=1e0):
    """geometry.py:138-181 with the same arguments and both return values (Gramian, norm penalty).  The O(N1*N2) part is
    `kern_mat` above (or a plain matmul when `kernalize` is false); the O(N) feature normalisation stays in torch."""
    feats = [fea_flat_1, fea_flat_2]
    norms = [_feature_norm(f, norm_dim) for f in feats]
    penalty = [torch.zeros((), dtype=f.dtype, device=f.device) for f in feats]
    if norm_mode:
        feats = [f / n for f, n in zip(feats, norms)]          # the reference divides without an epsilon
        if norm_dim == 2:                                       # sparsity measure of the normalised maps
            penalty = [-f.norm(dim=2).mean() for f in feats]
    elif norm_dim in (1, 2):
        penalty = [n.mean() for n in norms]
    if kernalize:
        g = kern_mat(feats[0], feats[1], dist_coef=dist_coef)
    else:
        g = torch.matmul(feats[0].transpose(1, 2), feats[1])
    return g, penalty[0] + penalty[1]


# ------------------------------------------------------------------------------------------------ fused inner product
def _items_struct(xs1, xs2, coefs):
    arr = (_lib.CvoItem * len(xs1))()
    for k, (a, b_, s) in enumerate(zip(xs1, xs2, coefs)):
        arr[k].x1, arr[k].x2, arr[k].c = a.data_ptr(), b_.data_ptr(), a.shape[1]
        arr[k].dist_coef = float(s) if s is not None else 0.0
    return arr


def _check_items(items_i, items_j, dist_coefs, what):
    if not (len(items_i) == len(items_j) == len(dist_coefs)) or not 1 <= len(items_i) <= 4:
        raise ValueError(f"b200unet.cvo.{what}: 1..4 domains, one dist_coef (or None) per domain")
    xs1 = [_f32c(t, what) for t in items_i]
    xs2 = [_f32c(t, what) for t in items_j]
    B, n1, n2 = xs1[0].shape[0], xs1[0].shape[2], xs2[0].shape[2]
    for a, b_ in zip(xs1, xs2):
        if a.shape[0] != B or b_.shape[0] != B or a.shape[2] != n1 or b_.shape[2] != n2 or a.shape[1] != b_.shape[1]:
            raise ValueError(f"b200unet.cvo.{what}: domains disagree on B / N / C: {tuple(a.shape)} vs {tuple(b_.shape)}")
    if sum(a.shape[1] for a in xs1) > 16:
        raise ValueError(f"b200unet.cvo.{what}: at most 16 channels over all domains")
    if sum(1 for s in dist_coefs if s is None) > 1:
        raise ValueError(f"b200unet.cvo.{what}: at most one plain inner-product domain (dist_coef None)")
    for s in dist_coefs:
        if s is not None and not s > 0:
            raise ValueError(f"b200unet.cvo.{what}: dist_coef must be positive (or None for the plain inner product)")
    return xs1, xs2, B, n1, n2


def _weights(w, B, n, what):
    if w is None:
        return None
    if w.numel() != B * n:
        raise ValueError(f"b200unet.cvo.{what}: weight map must hold B*N = {B * n} values, got {tuple(w.shape)}")
    return w.detach().to(torch.float32).reshape(B, n).contiguous()


class _InnerProdFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, k, dist_coefs, w_i, w_j, *items):
        items_i, items_j = items[:k], items[k:]
        xs1, xs2, B, n1, n2 = _check_items(items_i, items_j, dist_coefs, "inner_product")
        w1, w2 = _weights(w_i, B, n1, "inner_product"), _weights(w_j, B, n2, "inner_product")
        if (w1 is None) != (w2 is None):
            raise ValueError("b200unet.cvo.inner_product: give both weight maps or neither")
        ct = sum(a.shape[1] for a in xs1)
        ws = _workspace(B, n1, n2, ct, xs1[0].device)
        out = torch.empty(B, dtype=torch.float32, device=xs1[0].device)
        arr = _items_struct(xs1, xs2, dist_coefs)
        _lib.check(_lib.load().b200unet_cvo_inner_prod_fwd(arr, k, w1.data_ptr() if w1 is not None else None,
                                                           w2.data_ptr() if w2 is not None else None, B, n1, n2, -1,
                                                           ws.data_ptr(), out.data_ptr(), None, _stream()), "cvo_inner_prod_fwd")
        ctx.k, ctx.dist_coefs, ctx.geom = k, tuple(dist_coefs), (B, n1, n2, ct)
        ctx.has_w = w1 is not None
        ctx.w_shapes = (w_i.shape if w_i is not None else None, w_j.shape if w_j is not None else None)
        ctx.in_dtypes = [t.dtype for t in items] + [w_i.dtype if w_i is not None else None, w_j.dtype if w_j is not None else None]
        ctx.save_for_backward(*xs1, *xs2, *([w1, w2] if ctx.has_w else []))
        return out.to(items[0].dtype)

    @staticmethod
    def backward(ctx, grad_out):
        k = ctx.k
        saved = ctx.saved_tensors
        xs1, xs2 = saved[:k], saved[k:2 * k]
        w1, w2 = (saved[2 * k], saved[2 * k + 1]) if ctx.has_w else (None, None)
        B, n1, n2, ct = ctx.geom
        need = ctx.needs_input_grad  # (k, dist_coefs, w_i, w_j, *items)
        g = grad_out.detach().to(torch.float32).contiguous()
        dx1 = [torch.empty_like(a) if need[4 + i] else None for i, a in enumerate(xs1)]
        dx2 = [torch.empty_like(a) if need[4 + k + i] else None for i, a in enumerate(xs2)]
        dw1 = torch.empty_like(w1) if (ctx.has_w and need[2]) else None
        dw2 = torch.empty_like(w2) if (ctx.has_w and need[3]) else None
        P = C.c_void_p
        p1 = (P * k)(*[t.data_ptr() if t is not None else None for t in dx1])
        p2 = (P * k)(*[t.data_ptr() if t is not None else None for t in dx2])
        ws = _workspace(B, n1, n2, ct, xs1[0].device)
        arr = _items_struct(xs1, xs2, ctx.dist_coefs)
        _lib.check(_lib.load().b200unet_cvo_inner_prod_bwd(
            arr, k, w1.data_ptr() if w1 is not None else None, w2.data_ptr() if w2 is not None else None, B, n1, n2,
            g.data_ptr(), ws.data_ptr(), p1, p2, dw1.data_ptr() if dw1 is not None else None,
            dw2.data_ptr() if dw2 is not None else None, _stream()), "cvo_inner_prod_bwd")
        if dw1 is not None:
            # the kernel returns d out / d w1_i = g * S_i with S built from w2; the weights enter as a product w1_i * w2_j
            dw1 = dw1.reshape(ctx.w_shapes[0]).to(ctx.in_dtypes[2 * k])
        if dw2 is not None:
            dw2 = dw2.reshape(ctx.w_shapes[1]).to(ctx.in_dtypes[2 * k + 1])
        grads = [d.to(ctx.in_dtypes[i]) if d is not None else None for i, d in enumerate(dx1 + dx2)]
        return (None, None, dw1, dw2, *grads)


def inner_product(items_i: Sequence[torch.Tensor], items_j: Sequence[torch.Tensor], dist_coefs: Sequence[Optional[float]],
                  w_i: Optional[torch.Tensor] = None, w_j: Optional[torch.Tensor] = None,
                  normalize_over_pts: bool = False) -> torch.Tensor:
    """sum_b sum_ij [w_i w_j] prod_k K_k[b, i, j] for one frame pair — `inner_prods[ij]` of calc_inner_prod
    (network_modules.py:1096-1149) with the Gramians of calc_gramian (995-1015) never materialised.

    items_i[k] / items_j[k]: B*C_k*N_i / B*C_k*N_j tensors of domain k (xyz, hsv_graduv, feature_normalized ...);
    dist_coefs[k]: the RBF scale `self.dist_coef[item]`, or None for the plain inner-product Gramian the reference uses
    for non-kernalized features (network_modules.py:1012-1013).  w_i / w_j: the B*1*N `feature_w` maps of
    weight_map_mode.  normalize_over_pts divides by N_i * N_j (the stated intent of lines 1144-1147)."""
    k = len(items_i)
    out = _InnerProdFunction.apply(k, tuple(dist_coefs), w_i, w_j, *items_i, *items_j).sum()
    if normalize_over_pts:
        out = out / (items_i[0].shape[2] * items_j[0].shape[2])
    return out


def cvo_losses(flat_sel: Sequence[Dict[str, torch.Tensor]], items: Sequence[str], dist_coef: Dict[str, Optional[float]],
               with_self_terms: bool = True, weight_key: Optional[str] = None,
               normalize_over_pts: bool = False) -> Dict[str, torch.Tensor]:
    """calc_gramian + calc_inner_prod + calc_loss_from_inner_prod (network_modules.py:995-1189) for two frames.
    flat_sel[i][item]: B*C*N_i tensors; dist_coef[item]: scale or None (plain inner product).  Returns the reference's
    loss dictionary: inner_prod, and with the self terms inner_prod_0_0, inner_prod_1_1, func_dist, cos_sim."""
    coefs = [dist_coef[it] for it in items]
    ip: Dict[Tuple[int, int], torch.Tensor] = {}
    pairs = [(0, 0), (1, 1), (0, 1)] if with_self_terms else [(0, 1)]
    for (i, j) in pairs:
        ip[(i, j)] = inner_product([flat_sel[i][it] for it in items], [flat_sel[j][it] for it in items], coefs,
                                   flat_sel[i][weight_key] if weight_key else None,
                                   flat_sel[j][weight_key] if weight_key else None, normalize_over_pts)
    losses = {"inner_prod": ip[(0, 1)]}
    if with_self_terms:
        losses["inner_prod_0_0"] = ip[(0, 0)]
        losses["inner_prod_1_1"] = ip[(1, 1)]
        losses["func_dist"] = ip[(0, 0)] + ip[(1, 1)] - 2 * ip[(0, 1)]
        losses["cos_sim"] = 1 - ip[(0, 1)] / torch.sqrt(ip[(0, 0)] * ip[(1, 1)])
    return losses


@torch.no_grad()
def calc_w_v(items_i: Sequence[torch.Tensor], items_j: Sequence[torch.Tensor], dist_coefs: Sequence[Optional[float]],
             geo_item: int = 0, w_i: Optional[torch.Tensor] = None, w_j: Optional[torch.Tensor] = None):
    """network_modules.py:1052-1094 without the B*N1*N2*3 cross_prod / cross_subtract tensors: w = sum_ij P_ij (x_i x x_j),
    v = sum_ij P_ij (x_i - x_j) over the geometry domain `geo_item`, normalised to a unit 6-vector (zero if its norm
    is below 1e-6).  Forward only, like the reference (`cross_prod_geo.requires_grad = False`)."""
    xs1, xs2, B, n1, n2 = _check_items(items_i, items_j, dist_coefs, "calc_w_v")
    w1, w2 = _weights(w_i, B, n1, "calc_w_v"), _weights(w_j, B, n2, "calc_w_v")
    ct = sum(a.shape[1] for a in xs1)
    ws = _workspace(B, n1, n2, ct, xs1[0].device)
    out = torch.empty(B, dtype=torch.float32, device=xs1[0].device)
    wv = torch.empty((B, 6), dtype=torch.float32, device=xs1[0].device)
    arr = _items_struct(xs1, xs2, dist_coefs)
    _lib.check(_lib.load().b200unet_cvo_inner_prod_fwd(arr, len(xs1), w1.data_ptr() if w1 is not None else None,
                                                       w2.data_ptr() if w2 is not None else None, B, n1, n2, int(geo_item),
                                                       ws.data_ptr(), out.data_ptr(), wv.data_ptr(), _stream()), "cvo w/v")
    norm = wv.norm(dim=1, keepdim=True)
    wv = torch.where(norm < 1e-6, torch.zeros_like(wv), wv / norm)
    return wv[:, :3], wv[:, 3:]
