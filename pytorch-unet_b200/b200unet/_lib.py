"""ctypes binding of libb200unet.so — the C ABI declared in include/b200unet.h.

The library is the product's only compute path: if it is missing, or a call fails, this module raises; there is
no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200UNET_LIB") or os.path.join(_HERE, "libb200unet.so")  # override: A/B builds of the same ABI

IMPL_AUTO, IMPL_DIRECT, IMPL_UMMA = 0, 1, 2


class View(C.Structure):
    """b200_view: strided NHWC bf16 window (+ optional low-order plane of the split precision tier)."""

    _fields_ = [("ptr", C.c_void_p), ("lo", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("stride_n", C.c_int64), ("stride_h", C.c_int64), ("stride_w", C.c_int64)]


class ConvFwdParams(C.Structure):
    _fields_ = [("src", View * 2), ("num_src", C.c_int32), ("taps", C.c_int32), ("pad", C.c_int32),
                ("w_packed", C.c_void_p), ("w_f32", C.c_void_p), ("bias", C.c_void_p), ("relu", C.c_int32),
                ("dst", View), ("impl", C.c_int32)]


class ConvDgradParams(C.Structure):
    _fields_ = [("dz", View), ("taps", C.c_int32), ("pad", C.c_int32), ("w_packed", C.c_void_p),
                ("w_f32", C.c_void_p), ("dst", View * 2), ("num_dst", C.c_int32), ("mask", C.c_void_p * 2),
                ("impl", C.c_int32)]


class ConvWgradParams(C.Structure):
    _fields_ = [("dz", View), ("src", View * 2), ("num_src", C.c_int32), ("taps", C.c_int32), ("pad", C.c_int32),
                ("dw_f32", C.c_void_p), ("db_f32", C.c_void_p), ("impl", C.c_int32)]


class ConvTFwdParams(C.Structure):
    _fields_ = [("x", View), ("y", View), ("w_packed", C.c_void_p), ("w_f32", C.c_void_p), ("bias", C.c_void_p),
                ("impl", C.c_int32)]


class ConvTDgradParams(C.Structure):
    _fields_ = [("dy", View), ("dx", View), ("w_packed", C.c_void_p), ("w_f32", C.c_void_p), ("mask", C.c_void_p),
                ("impl", C.c_int32)]


class ConvTWgradParams(C.Structure):
    _fields_ = [("x", View), ("dy", View), ("dw_f32", C.c_void_p), ("db_f32", C.c_void_p), ("impl", C.c_int32)]


class CvoItem(C.Structure):
    """b200_cvo_item"""

    _fields_ = [("x1", C.c_void_p), ("x2", C.c_void_p), ("c", C.c_int32), ("dist_coef", C.c_float)]


class AdamJob(C.Structure):
    """b200_adam_job"""

    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("pack_fwd", C.c_void_p), ("pack_dgrad", C.c_void_p), ("numel", C.c_int64), ("kind", C.c_int32),
                ("dim0", C.c_int32), ("dim1", C.c_int32), ("taps", C.c_int32), ("src0_c", C.c_int32),
                ("split", C.c_int32), ("block0", C.c_int32), ("nblocks", C.c_int32)]


_VP = C.POINTER(View)
_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol include/b200unet.h declares
SIGNATURES = {
    "b200unet_abi_version": (_I, []),
    "b200unet_last_error": (C.c_char_p, []),
    "b200unet_device_ok": (_I, []),
    "b200unet_num_sms": (_I, []),
    "b200unet_launch_count": (C.c_ulonglong, []),
    "b200unet_fallback_count": (C.c_ulonglong, []),
    "b200unet_conv_fwd": (_I, [C.POINTER(ConvFwdParams), _P]),
    "b200unet_conv_dgrad": (_I, [C.POINTER(ConvDgradParams), _P]),
    "b200unet_conv_wgrad_workspace_bytes": (_SZ, [C.POINTER(ConvWgradParams)]),
    "b200unet_conv_wgrad": (_I, [C.POINTER(ConvWgradParams), _P, _SZ, _P]),
    "b200unet_convt_fwd": (_I, [C.POINTER(ConvTFwdParams), _P]),
    "b200unet_convt_dgrad": (_I, [C.POINTER(ConvTDgradParams), _P]),
    "b200unet_convt_wgrad_workspace_bytes": (_SZ, [C.POINTER(ConvTWgradParams)]),
    "b200unet_convt_wgrad": (_I, [C.POINTER(ConvTWgradParams), _P, _SZ, _P]),
    "b200unet_conv_fwd_impl": (_I, [C.POINTER(ConvFwdParams)]),
    "b200unet_conv_dgrad_impl": (_I, [C.POINTER(ConvDgradParams)]),
    "b200unet_conv_wgrad_impl": (_I, [C.POINTER(ConvWgradParams)]),
    "b200unet_convt_fwd_impl": (_I, [C.POINTER(ConvTFwdParams)]),
    "b200unet_convt_dgrad_impl": (_I, [C.POINTER(ConvTDgradParams)]),
    "b200unet_convt_wgrad_impl": (_I, [C.POINTER(ConvTWgradParams)]),
    "b200unet_pack_conv_weight_bytes": (_SZ, [_I, _I, C.POINTER(C.c_int), _I, _I]),
    "b200unet_pack_conv_weight": (_I, [_P, _I, _I, C.POINTER(C.c_int), _I, _I, _P, _P]),
    "b200unet_pack_convt_weight_bytes": (_SZ, [_I, _I, _I]),
    "b200unet_pack_convt_weight": (_I, [_P, _I, _I, _I, _P, _P]),
    "b200unet_maxpool2x2_fwd": (_I, [_VP, _VP, _P, _P, _P]),
    "b200unet_maxpool2x2_bwd": (_I, [_VP, _P, _VP, _VP, _I, _I, _P, _P]),
    "b200unet_maxpool2x2_bwd_premasked": (_I, [_VP, _P, _VP, _VP, _VP, _I, _I, _P]),
    "b200unet_bilinear_up2x_fwd": (_I, [_VP, _VP, _P]),
    "b200unet_bilinear_up2x_bwd": (_I, [_VP, _VP, _P, _P]),
    "b200unet_bn_workspace_bytes": (_SZ, [_I]),
    "b200unet_bn_fwd_train": (_I, [_VP, _VP, _P, _P, _P, _P, _F, _F, _P, _P, _P, _SZ, _P]),
    "b200unet_bn_fwd_eval": (_I, [_VP, _VP, _P, _P, _P, _P, _F, _P]),
    "b200unet_bn_bwd": (_I, [_VP, _VP, _VP, _P, _P, _P, _P, _P, _I, _P, _SZ, _P]),
    "b200unet_head_workspace_bytes": (_SZ, [_I, _I]),
    "b200unet_head_fwd": (_I, [_VP, _P, _P, _I, _I, _P, _P]),
    "b200unet_head_bwd": (_I, [_VP, _P, _P, _I, _I, _P, _VP, _P, _P, _P, _P, _SZ, _P]),
    "b200unet_head_ce_fwd": (_I, [_VP, _P, _P, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "b200unet_head_ce_bwd": (_I, [_VP, _P, _P, _I, _I, _P, _P, _P, _VP, _P, _P, _P, _P, _SZ, _P]),
    "b200unet_nchw_f32_to_nhwc_bf16": (_I, [_P, _VP, _P]),
    "b200unet_nhwc_bf16_to_nchw_f32": (_I, [_VP, _P, _P]),
    "b200unet_u8_nhwc_to_bf16": (_I, [_P, _I, _VP, _P, _P, _I, _P]),
    "b200unet_im2col3x3": (_I, [_VP, _VP, _I, _P]),
    "b200unet_channel_sum": (_I, [_VP, _P, _P, _SZ, _P]),
    "b200unet_relu_mask": (_I, [_VP, _P, _VP, _P]),
    "b200unet_adam_plan": (_I, [C.POINTER(AdamJob), _I]),
    "b200unet_adam_upload": (_I, [_P, C.POINTER(AdamJob), _I, _P]),
    "b200unet_adam_step": (_I, [_P, _I, _I, _F, _F, _F, _F, _F, _P, _P]),
    "b200unet_cvo_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "b200unet_cvo_sub_norm_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "b200unet_cvo_sub_norm_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "b200unet_cvo_kern_mat_fwd": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P]),
    "b200unet_cvo_kern_mat_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "b200unet_cvo_cross_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "b200unet_cvo_inner_prod_fwd": (_I, [C.POINTER(CvoItem), _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "b200unet_cvo_inner_prod_bwd": (_I, [C.POINTER(CvoItem), _I, _P, _P, _I, _I, _I, _P, _P, C.POINTER(_P), C.POINTER(_P),
                                         _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the shared object (once) and types every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python pytorch-unet_b200/build.py` "
            "(there is no CPU / PyTorch fallback for the U-Net hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b200unet_abi_version() != 4:
        raise RuntimeError("libb200unet.so ABI version mismatch: rebuild with pytorch-unet_b200/build.py")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200unet_last_error().decode(errors="replace")
        raise RuntimeError(f"b200unet {what} failed (rc={rc}): {msg}")


class Split:
    """Activation of the split precision tier: value = hi + lo, two NHWC bf16 tensors of identical geometry
    (include/b200unet.h).  Forward operators accept and return it in place of a plain tensor; backward operators take
    `.hi` only."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor):
        assert hi.shape == lo.shape and hi.stride() == lo.stride() and hi.dtype == lo.dtype == torch.bfloat16
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    def __getitem__(self, idx) -> "Split":  # windows (center crop) apply to both planes
        return Split(self.hi[idx], self.lo[idx])

    def float(self) -> torch.Tensor:
        return self.hi.float() + self.lo.float()

    @staticmethod
    def from_float(x: torch.Tensor) -> "Split":
        hi = x.to(torch.bfloat16)
        return Split(hi, (x - hi.float()).to(torch.bfloat16))


def hi_of(t):
    """The bf16 plane backward operators read."""
    return t.hi if isinstance(t, Split) else t


def view(t) -> View:
    """b200_view of an NHWC bf16 tensor (any strides with unit channel stride) or of a Split pair."""
    lo = None
    if isinstance(t, Split):
        t, lo = t.hi, t.lo
    assert t.dim() == 4 and t.dtype == torch.bfloat16, (t.shape, t.dtype)
    assert t.stride(3) == 1 or t.shape[3] == 1
    return View(t.data_ptr(), None if lo is None else lo.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3],
                t.stride(0), t.stride(1), t.stride(2))


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
