import torch


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(x_nchw):
    """fp32 NCHW -> bf16 NHWC (values rounded to bf16)."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def rb(x):
    """round to bf16, keep fp32"""
    return x.to(torch.bfloat16).float()
