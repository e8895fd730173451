"""Operator-level Python wrappers over the C ABI (include/b200unet.h): one function per reference operator call
on the hot path (unet.py:79, 92-100, 143-148, 152-163, 65-71; README.md:58).  Tensors are NHWC bf16 CUDA
tensors unless noted; parameters are the reference's fp32 state_dict tensors.  Every function launches on the
current CUDA stream and raises RuntimeError if the library reports an error.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (IMPL_AUTO, IMPL_DIRECT, IMPL_UMMA, ConvDgradParams, ConvFwdParams, ConvTDgradParams,
                   ConvTFwdParams, ConvTWgradParams, ConvWgradParams, Split, check, hi_of, ptr, stream_ptr, view)

__all__ = ["IMPL_AUTO", "IMPL_DIRECT", "IMPL_UMMA", "Split", "hi_of"]

MAX_CLASSES = 8  # head / loss kernels: one accumulator per class in registers (csrc/head.cu)

_workspaces = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream)."""
    key = (torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def tensor_cores_available() -> bool:
    """True on an sm_100 device (the tcgen05 kernels can run)."""
    return bool(_lib.load().b200unet_device_ok())


def _nhwc_empty(n, h, w, c, device, split: bool = False):
    """Output activation: a bf16 tensor, or a hi/lo pair in the split precision tier."""
    hi = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=device)
    return Split(hi, torch.empty_like(hi)) if split else hi


# ------------------------------------------------------------------ layout
def to_nhwc(x: torch.Tensor, split: bool = False, c_pad: Optional[int] = None):
    """fp32 NCHW (module input, unet.py:73) -> bf16 NHWC (hi/lo pair in the split tier).  c_pad > C: the result has
    c_pad channels, the extra ones zero."""
    assert x.dim() == 4 and x.dtype == torch.float32 and x.is_cuda
    x = x.contiguous()
    n, c, h, w = x.shape
    if c_pad is not None and c_pad > c:
        hi = torch.zeros((n, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
        y = Split(hi, torch.zeros_like(hi)) if split else hi
        v = view(y[..., :c])
    else:
        y = _nhwc_empty(n, h, w, c, x.device, split)
        v = view(y)
    check(_lib.load().b200unet_nchw_f32_to_nhwc_bf16(x.data_ptr(), C.byref(v), stream_ptr()), "nchw_f32_to_nhwc_bf16")
    return y


def u8_to_nhwc(img: torch.Tensor, split: bool = False, c_pad: Optional[int] = None, scale: Optional[torch.Tensor] = None,
               shift: Optional[torch.Tensor] = None, divide_255: bool = True):
    """uint8 [N,H,W,C] device batch -> the NHWC bf16 operand of the first convolution: img / 255 (dataloader.py:258-264)
    [* scale + shift per channel], channels zero-padded to c_pad, hi/lo planes in the split tier."""
    assert img.dim() == 4 and img.dtype == torch.uint8 and img.is_cuda and img.is_contiguous()
    n, h, w, c = img.shape
    cp = max(c, c_pad or c)
    y = _nhwc_empty(n, h, w, cp, img.device, split)
    v = view(y)
    check(_lib.load().b200unet_u8_nhwc_to_bf16(img.data_ptr(), c, C.byref(v), ptr(scale), ptr(shift), int(divide_255),
                                               stream_ptr()), "u8_nhwc_to_bf16")
    return y


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    v = view(x)
    check(_lib.load().b200unet_nhwc_bf16_to_nchw_f32(C.byref(v), y.data_ptr(), stream_ptr()), "nhwc_bf16_to_nchw_f32")
    return y


def im2col3x3(x: torch.Tensor, pad: int) -> torch.Tensor:
    """First-layer patches: [N,H,W,Cin] (Cin <= 7) -> [N,Ho,Wo,Kp], channel c*9 + r*3 + s, Kp = roundup(9*Cin, 16)."""
    n, h, w, c = x.shape
    kp = (9 * c + 15) // 16 * 16
    out = _nhwc_empty(n, h + 2 * pad - 2, w + 2 * pad - 2, kp, x.device)
    vx, vo = view(x), view(out)
    check(_lib.load().b200unet_im2col3x3(C.byref(vx), C.byref(vo), pad, stream_ptr()), "im2col3x3")
    return out


def _pack_out(nbytes: int, device, out: Optional[torch.Tensor]) -> torch.Tensor:
    if out is not None and out.numel() == nbytes // 2 and out.dtype == torch.bfloat16:
        return out
    return torch.empty(nbytes // 2, dtype=torch.bfloat16, device=device)


def pack_conv_weight(w: torch.Tensor, src_c: Sequence[int], mode: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [cout][cin][k][k] -> bf16 GEMM operand (mode 0: fprop, 1: dgrad, 2: fprop of the split tier)."""
    lib = _lib.load()
    cout, cin, k, _ = w.shape
    assert sum(src_c) == cin and w.dtype == torch.float32 and w.is_contiguous()
    arr = (C.c_int * len(src_c))(*src_c)
    nbytes = lib.b200unet_pack_conv_weight_bytes(cout, len(src_c), arr, k * k, mode)
    out = _pack_out(nbytes, w.device, out)
    check(lib.b200unet_pack_conv_weight(w.data_ptr(), cout, len(src_c), arr, k * k, mode, out.data_ptr(),
                                        stream_ptr()), "pack_conv_weight")
    return out


def pack_convt_weight(w: torch.Tensor, mode: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    cin, cout = w.shape[0], w.shape[1]
    assert w.dtype == torch.float32 and w.is_contiguous()
    nbytes = lib.b200unet_pack_convt_weight_bytes(cin, cout, mode)
    out = _pack_out(nbytes, w.device, out)
    check(lib.b200unet_pack_convt_weight(w.data_ptr(), cin, cout, mode, out.data_ptr(), stream_ptr()),
          "pack_convt_weight")
    return out


# ------------------------------------------------------------------ convolution family
def conv_fwd(srcs: Sequence[torch.Tensor], w: torch.Tensor, bias: Optional[torch.Tensor], pad: int, relu: bool,
             w_packed: Optional[torch.Tensor] = None, impl: int = IMPL_AUTO,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conv2d(k=3|1, padding=pad)(cat(srcs, channel)) [+bias] [+ReLU]; srcs are already-cropped windows."""
    lib = _lib.load()
    cout, cin, k, _ = w.shape
    n, h, wd, _ = srcs[0].shape
    ho, wo = h + 2 * pad - (k - 1), wd + 2 * pad - (k - 1)
    split = isinstance(srcs[0], Split)
    if out is None:
        out = _nhwc_empty(n, ho, wo, cout, srcs[0].device, split)
    p = ConvFwdParams()
    for i, s in enumerate(srcs):
        p.src[i] = view(s)
    p.num_src, p.taps, p.pad = len(srcs), k * k, pad
    p.w_f32, p.bias, p.relu = w.data_ptr(), ptr(bias), int(relu)
    p.dst, p.impl = view(out), impl
    if impl != IMPL_DIRECT and lib.b200unet_conv_fwd_impl(C.byref(p)) == IMPL_UMMA:
        if callable(w_packed):
            w_packed = w_packed()
        if w_packed is None:
            w_packed = pack_conv_weight(w, [s.shape[3] for s in srcs], 2 if split else 0)
        p.w_packed = w_packed.data_ptr()
    check(lib.b200unet_conv_fwd(C.byref(p), stream_ptr()), "conv_fwd")
    return out


def conv_dgrad(dz: torch.Tensor, w: torch.Tensor, pad: int, dsts: Sequence[torch.Tensor],
               masks: Sequence[Optional[torch.Tensor]] = (None, None), w_packed: Optional[torch.Tensor] = None,
               impl: int = IMPL_AUTO) -> None:
    """Backward-data of conv_fwd into the (pre-allocated) destinations, split by input channel."""
    lib = _lib.load()
    k = w.shape[2]
    p = ConvDgradParams()
    p.dz, p.taps, p.pad = view(dz), k * k, pad
    p.w_f32 = w.data_ptr()
    for i, d in enumerate(dsts):
        p.dst[i] = view(d)
        m = masks[i] if i < len(masks) else None
        if m is not None:
            assert m.stride() == d.stride() and m.shape == d.shape
        p.mask[i] = ptr(m)
    p.num_dst, p.impl = len(dsts), impl
    if impl != IMPL_DIRECT and lib.b200unet_conv_dgrad_impl(C.byref(p)) == IMPL_UMMA:
        if callable(w_packed):
            w_packed = w_packed()
        if w_packed is None:
            w_packed = pack_conv_weight(w, [w.shape[1]], 1)
        p.w_packed = w_packed.data_ptr()
    check(lib.b200unet_conv_dgrad(C.byref(p), stream_ptr()), "conv_dgrad")


def conv_wgrad(dz: torch.Tensor, srcs: Sequence[torch.Tensor], k: int, pad: int, want_db: bool = True,
               impl: int = IMPL_AUTO, dw: Optional[torch.Tensor] = None,
               db: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    lib = _lib.load()
    cout = dz.shape[3]
    cin = sum(s.shape[3] for s in srcs)
    if dw is None:
        dw = torch.empty((cout, cin, k, k), dtype=torch.float32, device=dz.device)
    if db is None and want_db:
        db = torch.empty((cout,), dtype=torch.float32, device=dz.device)
    p = ConvWgradParams()
    p.dz = view(dz)
    for i, s in enumerate(srcs):
        p.src[i] = view(s)
    p.num_src, p.taps, p.pad = len(srcs), k * k, pad
    p.dw_f32, p.db_f32, p.impl = dw.data_ptr(), ptr(db), impl
    nbytes = lib.b200unet_conv_wgrad_workspace_bytes(C.byref(p))
    ws = workspace(nbytes, dz.device)
    check(lib.b200unet_conv_wgrad(C.byref(p), ws.data_ptr(), ws.numel(), stream_ptr()), "conv_wgrad")
    return dw, db


def convt_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], w_packed=None, impl: int = IMPL_AUTO,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ConvTranspose2d(k=2, s=2): w fp32 [cin][cout][2][2]."""
    lib = _lib.load()
    n, h, wd, cin = x.shape
    cout = w.shape[1]
    split = isinstance(x, Split)
    if out is None:
        out = _nhwc_empty(n, 2 * h, 2 * wd, cout, x.device, split)
    p = ConvTFwdParams()
    p.x, p.y = view(x), view(out)
    p.w_f32, p.bias, p.impl = w.data_ptr(), ptr(bias), impl
    if impl != IMPL_DIRECT and lib.b200unet_convt_fwd_impl(C.byref(p)) == IMPL_UMMA:
        if callable(w_packed):
            w_packed = w_packed()
        if w_packed is None:
            w_packed = pack_convt_weight(w, 2 if split else 0)
        p.w_packed = w_packed.data_ptr()
    check(lib.b200unet_convt_fwd(C.byref(p), stream_ptr()), "convt_fwd")
    return out


def convt_dgrad(dy: torch.Tensor, w: torch.Tensor, dx: torch.Tensor, mask: Optional[torch.Tensor] = None,
                w_packed=None, impl: int = IMPL_AUTO) -> None:
    lib = _lib.load()
    p = ConvTDgradParams()
    p.dy, p.dx = view(dy), view(dx)
    p.w_f32, p.mask, p.impl = w.data_ptr(), ptr(mask), impl
    if impl != IMPL_DIRECT and lib.b200unet_convt_dgrad_impl(C.byref(p)) == IMPL_UMMA:
        if callable(w_packed):
            w_packed = w_packed()
        if w_packed is None:
            w_packed = pack_convt_weight(w, 1)
        p.w_packed = w_packed.data_ptr()
    check(lib.b200unet_convt_dgrad(C.byref(p), stream_ptr()), "convt_dgrad")


def convt_wgrad(x: torch.Tensor, dy: torch.Tensor, want_db: bool = True, impl: int = IMPL_AUTO,
                dw: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None):
    lib = _lib.load()
    cin, cout = x.shape[3], dy.shape[3]
    if dw is None:
        dw = torch.empty((cin, cout, 2, 2), dtype=torch.float32, device=x.device)
    if db is None and want_db:
        db = torch.empty((cout,), dtype=torch.float32, device=x.device)
    p = ConvTWgradParams()
    p.x, p.dy = view(x), view(dy)
    p.dw_f32, p.db_f32, p.impl = dw.data_ptr(), ptr(db), impl
    nbytes = lib.b200unet_convt_wgrad_workspace_bytes(C.byref(p))
    ws = workspace(nbytes, x.device)
    check(lib.b200unet_convt_wgrad(C.byref(p), ws.data_ptr(), ws.numel(), stream_ptr()), "convt_wgrad")
    return dw, db


# ------------------------------------------------------------------ pool / upsample
def maxpool_fwd(x: torch.Tensor, want_idx64: bool = False):
    """F.max_pool2d(x, 2): returns (y, idx8[, idx64]); idx8 = a*2+b code of the arg-max inside the window."""
    n, h, w, c = x.shape
    y = _nhwc_empty(n, h // 2, w // 2, c, x.device, isinstance(x, Split))
    idx8 = torch.empty((n, h // 2, w // 2, c), dtype=torch.uint8, device=x.device)
    idx64 = torch.empty((n, h // 2, w // 2, c), dtype=torch.int64, device=x.device) if want_idx64 else None
    vx, vy = view(x), view(y)
    check(_lib.load().b200unet_maxpool2x2_fwd(C.byref(vx), C.byref(vy), idx8.data_ptr(), ptr(idx64), stream_ptr()),
          "maxpool2x2_fwd")
    return (y, idx8, idx64) if want_idx64 else (y, idx8)


def maxpool_bwd(dy: torch.Tensor, idx8: torch.Tensor, dx: torch.Tensor, add: Optional[torch.Tensor] = None,
                add_y: int = 0, add_x: int = 0, mask: Optional[torch.Tensor] = None,
                pooled: Optional[torch.Tensor] = None) -> None:
    """pooled (the pool's forward output) selects the pre-masked form: `add` is already ReLU-masked by its producer and
    the scattered term is masked by [pooled > 0]; `mask` must then be None."""
    vdy, vdx = view(dy), view(dx)
    vadd = view(add) if add is not None else None
    if pooled is not None:
        assert mask is None
        vy = view(hi_of(pooled))
        check(_lib.load().b200unet_maxpool2x2_bwd_premasked(C.byref(vdy), idx8.data_ptr(), C.byref(vy), C.byref(vdx),
                                                            C.byref(vadd) if vadd is not None else None, add_y, add_x,
                                                            stream_ptr()), "maxpool2x2_bwd_premasked")
        return
    if mask is not None:
        assert mask.stride() == dx.stride() and mask.shape == dx.shape
    check(_lib.load().b200unet_maxpool2x2_bwd(C.byref(vdy), idx8.data_ptr(), C.byref(vdx),
                                              C.byref(vadd) if vadd is not None else None, add_y, add_x, ptr(mask),
                                              stream_ptr()), "maxpool2x2_bwd")


def bilinear_fwd(x):
    n, h, w, c = x.shape
    y = _nhwc_empty(n, 2 * h, 2 * w, c, x.device, isinstance(x, Split))
    vx, vy = view(x), view(y)
    check(_lib.load().b200unet_bilinear_up2x_fwd(C.byref(vx), C.byref(vy), stream_ptr()), "bilinear_up2x_fwd")
    return y


def bilinear_bwd(dy: torch.Tensor, dx: torch.Tensor, mask: Optional[torch.Tensor] = None) -> None:
    vdy, vdx = view(dy), view(dx)
    check(_lib.load().b200unet_bilinear_up2x_bwd(C.byref(vdy), C.byref(vdx), ptr(mask), stream_ptr()),
          "bilinear_up2x_bwd")


# ------------------------------------------------------------------ batch norm
def bn_fwd_train(x, gamma, beta, running_mean, running_var, momentum: float, eps: float):
    lib = _lib.load()
    n, h, w, c = x.shape
    y = _nhwc_empty(n, h, w, c, x.device, isinstance(x, Split))
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    invstd = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = workspace(lib.b200unet_bn_workspace_bytes(c), x.device)
    vx, vy = view(x), view(y)
    check(lib.b200unet_bn_fwd_train(C.byref(vx), C.byref(vy), gamma.data_ptr(), beta.data_ptr(), ptr(running_mean),
                                    ptr(running_var), momentum, eps, mean.data_ptr(), invstd.data_ptr(),
                                    ws.data_ptr(), ws.numel(), stream_ptr()), "bn_fwd_train")
    return y, mean, invstd


def bn_fwd_eval(x, gamma, beta, running_mean, running_var, eps: float):
    n, h, w, c = x.shape
    y = _nhwc_empty(n, h, w, c, x.device, isinstance(x, Split))
    vx, vy = view(x), view(y)
    check(_lib.load().b200unet_bn_fwd_eval(C.byref(vx), C.byref(vy), gamma.data_ptr(), beta.data_ptr(),
                                           running_mean.data_ptr(), running_var.data_ptr(), eps, stream_ptr()),
          "bn_fwd_eval")
    return y


def bn_bwd(x, dy, gamma, mean, invstd, relu_mask: bool, dx=None, dgamma=None, dbeta=None, frozen_stats: bool = False):
    """BatchNorm2d backward (+ ReLU mask of the conv in front).  frozen_stats: `mean` / `invstd` are the running
    statistics of an eval-mode forward (torch's batch_norm backward with training=False)."""
    lib = _lib.load()
    c = x.shape[3]
    if dx is None:
        dx = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
    if dbeta is None:
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = workspace(lib.b200unet_bn_workspace_bytes(c), x.device)
    vx, vdy, vdx = view(x), view(dy), view(dx)
    check(lib.b200unet_bn_bwd(C.byref(vx), C.byref(vdy), C.byref(vdx), gamma.data_ptr(), mean.data_ptr(),
                              invstd.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), int(relu_mask) | (2 if frozen_stats else 0),
                              ws.data_ptr(),
                              ws.numel(), stream_ptr()), "bn_bwd")
    return dx, dgamma, dbeta


# ------------------------------------------------------------------ head
def head_fwd(x, w, b, relu: bool) -> torch.Tensor:
    """1x1 classifier: returns fp32 NCHW logits (unet.py:84)."""
    n, h, wd, c = x.shape
    k = w.shape[0]
    logits = torch.empty((n, k, h, wd), dtype=torch.float32, device=x.device)
    vx = view(x)
    check(_lib.load().b200unet_head_fwd(C.byref(vx), w.data_ptr(), ptr(b), k, int(relu), logits.data_ptr(),
                                        stream_ptr()), "head_fwd")
    return logits


def head_bwd(x, w, b, relu: bool, dlogits, dx=None, mask=None, want_dx: bool = True, dw=None, db=None):
    lib = _lib.load()
    k, c = w.shape[0], x.shape[3]
    if dx is None and want_dx:
        dx = torch.empty_like(x)
    if dw is None:
        dw = torch.empty((k, c, 1, 1), dtype=torch.float32, device=x.device)
    if db is None:
        db = torch.empty((k,), dtype=torch.float32, device=x.device)
    ws = workspace(lib.b200unet_head_workspace_bytes(c, k), x.device)
    vx = view(x)
    vdx = view(dx) if dx is not None else None
    check(lib.b200unet_head_bwd(C.byref(vx), w.data_ptr(), ptr(b), k, int(relu), dlogits.contiguous().data_ptr(),
                                C.byref(vdx) if vdx is not None else None, ptr(mask), dw.data_ptr(), db.data_ptr(),
                                ws.data_ptr(), ws.numel(), stream_ptr()), "head_bwd")
    return dx, dw, db


def head_ce_fwd(x, w, b, relu: bool, labels, want_logits: bool = False):
    """Head fused with F.cross_entropy(mean): returns (loss scalar tensor, ce_state, logits or None)."""
    lib = _lib.load()
    n, h, wd, c = x.shape
    k = w.shape[0]
    assert labels.dtype == torch.int64 and tuple(labels.shape) == (n, h, wd) and labels.is_contiguous()
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    state = torch.empty(2, dtype=torch.float32, device=x.device)
    logits = torch.empty((n, k, h, wd), dtype=torch.float32, device=x.device) if want_logits else None
    ws = workspace(lib.b200unet_head_workspace_bytes(c, k), x.device)
    vx = view(x)
    check(lib.b200unet_head_ce_fwd(C.byref(vx), w.data_ptr(), ptr(b), k, int(relu), labels.data_ptr(),
                                   loss.data_ptr(), ptr(logits), state.data_ptr(), ws.data_ptr(), ws.numel(),
                                   stream_ptr()), "head_ce_fwd")
    return loss, state, logits


def head_ce_bwd(x, w, b, relu: bool, labels, state, grad_scale=None, dx=None, mask=None, want_dx: bool = True,
                dw=None, db=None):
    lib = _lib.load()
    k, c = w.shape[0], x.shape[3]
    if dx is None and want_dx:
        dx = torch.empty_like(x)
    if dw is None:
        dw = torch.empty((k, c, 1, 1), dtype=torch.float32, device=x.device)
    if db is None:
        db = torch.empty((k,), dtype=torch.float32, device=x.device)
    ws = workspace(lib.b200unet_head_workspace_bytes(c, k), x.device)
    vx = view(x)
    vdx = view(dx) if dx is not None else None
    check(lib.b200unet_head_ce_bwd(C.byref(vx), w.data_ptr(), ptr(b), k, int(relu), labels.data_ptr(),
                                   ptr(grad_scale), state.data_ptr(), C.byref(vdx) if vdx is not None else None,
                                   ptr(mask), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "head_ce_bwd")
    return dx, dw, db


def relu_mask(x, mask, out=None):
    if out is None:
        out = torch.empty_like(x)
    vx, vo = view(x), view(out)
    check(_lib.load().b200unet_relu_mask(C.byref(vx), mask.data_ptr(), C.byref(vo), stream_ptr()), "relu_mask")
    return out


def channel_sum(x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    c = x.shape[3]
    if out is None:
        out = torch.empty(c, dtype=torch.float32, device=x.device)
    ws = workspace(lib.b200unet_bn_workspace_bytes(c), x.device)
    vx = view(x)
    check(lib.b200unet_channel_sum(C.byref(vx), out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "channel_sum")
    return out
