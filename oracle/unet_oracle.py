"""CPU fp32 restatement of the reference U-Net hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; the product package (pytorch-unet_b200/) never does.

What is restated (reference file:line, relative to minghanz/pytorch-unet):
  * UNet.forward                unet.py:73-84   / unet_original.py:64-75
  * UNetConvBlock               unet.py:87-106  / unet_original.py:78-97   (Conv3x3 -> ReLU -> [BatchNorm])
  * UNetUpBlock                 unet.py:139-166 / unet_original.py:100-127 (paper decoder block)
  * UNetUpBlockDeep             unet.py:169-199                            (decoder block `unet.py` really uses)
  * head `last` (+ReLU non_neg) unet.py:65-71, 84
  * loss F.cross_entropy        README.md:57-62 (mean over N*H'*W')

The arithmetic itself lives in a third-party dependency of the reference that is not vendored in its tree
(PyTorch; the reference pins no version — circa torch 1.2/1.3 by its era).  This restatement binds to the
published semantics of torch.nn.functional as installed (torch 2.11): conv2d, conv_transpose2d(k=2,s=2),
max_pool2d(2), interpolate(bilinear, scale 2, align_corners=False), batch_norm(eps 1e-5, momentum 0.1),
cross_entropy(mean).

Pinning: the reference ships no tests, golden vectors or fixtures for this path.  The
pin is therefore made from outputs of the reference itself: oracle/make_golden.py imports the unmodified
reference modules from /root/reference, runs them on seeded inputs and stores inputs, weights, logits, loss
and gradients under tests/golden/*.npz; tests/test_oracle.py checks this restatement against those files.

Everything works on a plain ``dict[str, Tensor]`` in the reference's state_dict schema, so the same weights
can be loaded into the reference module, this oracle and the CUDA module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetSpec:
    """Constructor arguments of the reference UNet (unet.py:9-19) plus which decoder block it builds."""

    in_channels: int = 1
    n_classes: int = 2
    depth: int = 5
    wf: int = 6
    padding: bool = False
    batch_norm: bool = False
    up_mode: str = "upconv"
    non_neg: bool = False
    up_block: str = "paper"  # 'paper' = UNetUpBlock (unet_original.py), 'deep' = UNetUpBlockDeep (unet.py)

    def down_widths(self) -> List[int]:
        return [2 ** (self.wf + i) for i in range(self.depth)]


# --------------------------------------------------------------------------- parameter schema / init
def param_shapes(spec: UNetSpec) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape for learnable tensors, in the reference's registration order
    (unet.py:46-71; key names verified against the real modules in tests/test_oracle.py)."""
    assert spec.up_mode in ("upconv", "upsample")  # unet.py:45
    shapes: Dict[str, Tuple[int, ...]] = {}

    def conv_block(prefix: str, cin: int, cout: int) -> None:
        i1, i2 = (0, 3) if spec.batch_norm else (0, 2)
        shapes[f"{prefix}.block.{i1}.weight"] = (cout, cin, 3, 3)
        shapes[f"{prefix}.block.{i1}.bias"] = (cout,)
        if spec.batch_norm:
            shapes[f"{prefix}.block.2.weight"] = (cout,)
            shapes[f"{prefix}.block.2.bias"] = (cout,)
        shapes[f"{prefix}.block.{i2}.weight"] = (cout, cout, 3, 3)
        shapes[f"{prefix}.block.{i2}.bias"] = (cout,)
        if spec.batch_norm:
            shapes[f"{prefix}.block.5.weight"] = (cout,)
            shapes[f"{prefix}.block.5.bias"] = (cout,)

    prev = spec.in_channels
    for i, w in enumerate(spec.down_widths()):
        conv_block(f"down_path.{i}", prev, w)
        prev = w
    for j, i in enumerate(reversed(range(spec.depth - 1))):
        skip = 2 ** (spec.wf + i)
        if spec.up_block == "paper":
            up_out, blk_in, blk_out = skip, prev, skip  # unet_original.py:57-60
        else:
            up_out, blk_in, blk_out = prev, prev + skip, prev  # unet.py:60, 173-179 (prev never shrinks)
        if spec.up_mode == "upconv":
            shapes[f"up_path.{j}.up.weight"] = (prev, up_out, 2, 2)
            shapes[f"up_path.{j}.up.bias"] = (up_out,)
        else:
            shapes[f"up_path.{j}.up.1.weight"] = (up_out, prev, 1, 1)
            shapes[f"up_path.{j}.up.1.bias"] = (up_out,)
        conv_block(f"up_path.{j}.conv_block", blk_in, blk_out)
        prev = blk_out
    head = "last.0" if spec.non_neg else "last"
    shapes[f"{head}.weight"] = (spec.n_classes, prev, 1, 1)
    shapes[f"{head}.bias"] = (spec.n_classes,)
    return shapes


def bn_buffer_names(spec: UNetSpec) -> List[str]:
    if not spec.batch_norm:
        return []
    return [k[: -len(".weight")] for k, s in param_shapes(spec).items() if k.endswith(".weight") and len(s) == 1]


def init_params(spec: UNetSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """PyTorch-default init (kaiming_uniform(a=sqrt 5) => U(+-1/sqrt(fan_in)); BN gamma 1 / beta 0)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    shapes = param_shapes(spec)
    for k, s in shapes.items():
        if len(s) == 4:
            fan_in = s[1] * s[2] * s[3]
            bound = 1.0 / math.sqrt(fan_in)
            out[k] = (torch.rand(s, generator=g) * 2 - 1) * bound
            out[k[: -len("weight")] + "bias"] = (torch.rand(shapes[k[: -len("weight")] + "bias"], generator=g) * 2 - 1) * bound
        elif k.endswith(".weight"):  # BN gamma
            out[k] = torch.ones(s)
            out[k[: -len("weight")] + "bias"] = torch.zeros(s)
    for b in bn_buffer_names(spec):
        c = shapes[b + ".weight"][0]
        out[b + ".running_mean"] = torch.zeros(c)
        out[b + ".running_var"] = torch.ones(c)
        out[b + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return {k: out[k] for k in list(shapes.keys()) + [k for k in out if k not in shapes]}


# --------------------------------------------------------------------------- forward
class _RoundBF16(torch.autograd.Function):
    """Storage-precision emulation for the `act_bf16` mode: value rounded to bf16 in forward, gradient in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


_ACT_BF16 = False


def _q(x):
    """Identity in the fp32 oracle; bf16 storage rounding when the product's activation precision is emulated
    (used only to separate "precision" from "logic" differences in the parity tests)."""
    return _RoundBF16.apply(x) if _ACT_BF16 else x


def _conv_block(sd, prefix: str, x, spec: UNetSpec, training: bool, new_stats: Optional[dict]):
    # unet.py:92-100: Conv2d(3x3, padding=int(padding)) -> ReLU -> [BatchNorm2d], twice.
    convs, bns = ((0, 3), (2, 5)) if spec.batch_norm else ((0, 2), (None, None))
    for ci, bi in zip(convs, bns):
        x = F.conv2d(x, sd[f"{prefix}.block.{ci}.weight"], sd[f"{prefix}.block.{ci}.bias"], padding=int(spec.padding))
        x = _q(F.relu(x))
        if spec.batch_norm:
            name = f"{prefix}.block.{bi}"
            rm = sd[name + ".running_mean"].clone()
            rv = sd[name + ".running_var"].clone()
            x = F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], training=training, momentum=0.1, eps=1e-5)
            x = _q(x)
            if new_stats is not None and training:
                new_stats[name + ".running_mean"] = rm
                new_stats[name + ".running_var"] = rv
    return x


def center_crop_offsets(bridge_hw: Tuple[int, int], target_hw: Tuple[int, int]) -> Tuple[int, int]:
    # unet.py:152-158
    return (bridge_hw[0] - target_hw[0]) // 2, (bridge_hw[1] - target_hw[1]) // 2


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, spec: UNetSpec, training: bool = True,
            new_stats: Optional[dict] = None, taps: Optional[dict] = None, act_bf16: bool = False) -> torch.Tensor:
    """Logits N x n_classes x H' x W' (unet.py:73-84).  `taps`, if given, collects intermediate tensors.
    act_bf16=True rounds every stored activation (and its gradient) to bf16 like the CUDA path does."""
    global _ACT_BF16
    prev, _ACT_BF16 = _ACT_BF16, act_bf16
    try:
        return _forward(sd, x, spec, training, new_stats, taps)
    finally:
        _ACT_BF16 = prev


def _forward(sd, x, spec, training, new_stats, taps):
    x = _q(x)
    bridges = []
    for i in range(spec.depth):
        x = _conv_block(sd, f"down_path.{i}", x, spec, training, new_stats)
        if taps is not None:
            taps[f"down.{i}"] = x
        if i != spec.depth - 1:
            bridges.append(x)
            x = F.max_pool2d(x, 2)  # unet.py:79
    for j in range(spec.depth - 1):
        bridge = bridges[-j - 1]
        if spec.up_mode == "upconv":  # unet.py:143 / 173
            up = _q(F.conv_transpose2d(x, sd[f"up_path.{j}.up.weight"], sd[f"up_path.{j}.up.bias"], stride=2))
        else:  # unet.py:145-148 / 175-178
            up = _q(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
            up = _q(F.conv2d(up, sd[f"up_path.{j}.up.1.weight"], sd[f"up_path.{j}.up.1.bias"]))
        dy, dx = center_crop_offsets(bridge.shape[2:], up.shape[2:])
        crop = bridge[:, :, dy:dy + up.shape[2], dx:dx + up.shape[3]]
        x = torch.cat([up, crop], 1)  # unet.py:163 — up first, bridge second
        x = _conv_block(sd, f"up_path.{j}.conv_block", x, spec, training, new_stats)
        if taps is not None:
            taps[f"up.{j}"] = x
    head = "last.0" if spec.non_neg else "last"
    x = F.conv2d(x, sd[head + ".weight"], sd[head + ".bias"])
    if spec.non_neg:  # unet.py:65-69
        x = F.relu(x)
    return x


def loss_and_grads(sd: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor, spec: UNetSpec,
                   training: bool = True, act_bf16: bool = False):
    """README.md:57-62 training step up to backward(): returns (logits, loss, {param: grad}, new BN stats)."""
    shapes = param_shapes(spec)
    leaves = {k: (v.detach().clone().requires_grad_(True) if k in shapes else v) for k, v in sd.items()}
    new_stats: dict = {}
    global _ACT_BF16
    prev, _ACT_BF16 = _ACT_BF16, act_bf16  # must also cover the backward pass
    try:
        logits = _forward(leaves, x, spec, training, new_stats, None)
        loss = F.cross_entropy(logits, y)
        names = list(shapes.keys())
        grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    finally:
        _ACT_BF16 = prev
    return logits.detach(), loss.detach(), dict(zip(names, grads)), new_stats


def output_hw(spec: UNetSpec, h: int, w: int) -> Tuple[int, int]:
    shrink = 0 if spec.padding else 4
    sizes = []
    for i in range(spec.depth):
        h, w = h - shrink, w - shrink
        if i != spec.depth - 1:
            sizes.append((h, w))
            h, w = h // 2, w // 2
    for _ in range(spec.depth - 1):
        h, w = 2 * h - shrink, 2 * w - shrink
    return h, w


# --------------------------------------------------------------------------- per-op oracles (kernel level)
def pool2x2_with_indices(x: torch.Tensor):
    """F.max_pool2d(x, 2) plus the int64 argmax (flat h*W+w in the input plane), unet.py:79."""
    return F.max_pool2d(x, 2, return_indices=True)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
