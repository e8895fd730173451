"""Input pipeline at the module boundary.

The reference converts every sample on the host — uint8 image / 255 in float64 (dataloader.py:258-264), HWC -> CHW
transpose and a per-sample `.to(device, dtype=torch.float)` inside `ToTensor.__call__` (dataloader.py:553-579) — so a
batch crosses PCIe as fp32 (4 bytes per value) in many small synchronous copies, and the model then re-lays it out.

Here a batch stays uint8 NHWC (what image decoders produce) in PINNED host memory, crosses the bus once as 1 byte per
value on a copy stream while the previous step computes, and one CUDA kernel (`b200unet_u8_nhwc_to_bf16`) writes the
normalised NHWC bf16 operand `down_path.0` reads (channel padding and the split tier's hi/lo planes included).  The
result is a `PackedImages`, which `UNet.forward` / `UNet.loss` accept in place of the fp32 NCHW tensor.

    stager = ImageStager(model, batch=32, height=572, width=572, channels=1)
    stager.put(u8_batch_0)                      # numpy / torch uint8 [N,H,W,C] (or [N,H,W]) -> async H2D
    for step in range(...):
        x = stager.get()                        # PackedImages of the batch put() last
        stager.put(next_u8_batch)               # travels while this step computes
        loss = model.loss(x, y)
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


class PackedImages:
    """A batch already in the layout the first convolution consumes: NHWC bf16 (a `Split` hi/lo pair in the split
    precision tier), channels zero-padded as the module pads them.  `shape` is the logical NCHW shape."""

    __slots__ = ("data", "shape")

    def __init__(self, data, shape):
        self.data, self.shape = data, tuple(shape)

    @property
    def is_cuda(self) -> bool:
        return True

    @property
    def device(self):
        return self.data.device


def pack_images(model, img_u8: torch.Tensor, mean: Optional[Sequence[float]] = None,
                std: Optional[Sequence[float]] = None) -> PackedImages:
    """uint8 [N,H,W,C] CUDA batch -> PackedImages for `model` (a b200unet.UNet or its DataParallel wrapper):
    value = u8 / 255 (dataloader.py:258-264), then optionally (value - mean[c]) / std[c]."""
    m = getattr(model, "module", model)
    if img_u8.dim() == 3:
        img_u8 = img_u8.unsqueeze(-1)
    n, h, w, c = img_u8.shape
    if c != m.in_channels:
        raise ValueError(f"pack_images: batch has {c} channels, the model takes {m.in_channels}")
    scale = shift = None
    if mean is not None or std is not None:
        mean_t = torch.as_tensor(mean if mean is not None else [0.0] * c, dtype=torch.float32, device=img_u8.device)
        std_t = torch.as_tensor(std if std is not None else [1.0] * c, dtype=torch.float32, device=img_u8.device)
        scale, shift = (1.0 / std_t).contiguous(), (-mean_t / std_t).contiguous()
    data = ops.u8_to_nhwc(img_u8.contiguous(), split=m.precision == "split", c_pad=m._cpad_image(c), scale=scale,
                          shift=shift)
    return PackedImages(data, (n, c, h, w))


class ImageStager:
    """Double-buffered pinned uint8 staging: `put()` copies a host batch into pinned memory and enqueues ONE H2D copy
    on a side stream; `get()` makes the compute stream wait for it and runs the conversion kernel."""

    def __init__(self, model, batch: int, height: int, width: int, channels: int, mean=None, std=None, device=None):
        m = getattr(model, "module", model)
        self.model, self.mean, self.std = model, mean, std
        self.device = device if device is not None else next(m.parameters()).device
        shape = (batch, height, width, channels)
        self._host = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._dev = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._free = [torch.cuda.Event(), torch.cuda.Event()]
        self._stream = torch.cuda.Stream(device=self.device)
        self._put, self._got = 0, 0
        self.h2d_bytes_per_batch = int(np.prod(shape))

    def put(self, batch_u8) -> None:
        if self._put - self._got >= 2:
            raise RuntimeError("ImageStager: two batches are already staged; call get() first")
        slot = self._put % 2
        src = torch.as_tensor(batch_u8)
        if src.dim() == 3:
            src = src.unsqueeze(-1)
        if src.dtype != torch.uint8 or tuple(src.shape) != tuple(self._host[slot].shape):
            raise ValueError(f"ImageStager.put: expected uint8 {tuple(self._host[slot].shape)}, got {src.dtype} {tuple(src.shape)}")
        if self._put >= 2:
            self._ready[slot].synchronize()  # the H2D copy that last read this pinned buffer (two puts ago) is done
        self._host[slot].copy_(src)          # host memcpy into pinned memory
        with torch.cuda.stream(self._stream):
            if self._put >= 2:
                self._stream.wait_event(self._free[slot])  # the kernel that read the device buffer has run
            self._dev[slot].copy_(self._host[slot], non_blocking=True)
            self._ready[slot].record(self._stream)
        self._put += 1

    def get(self) -> PackedImages:
        if self._got >= self._put:
            raise RuntimeError("ImageStager.get: nothing staged; call put() first")
        slot = self._got % 2
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[slot])
        x = pack_images(self.model, self._dev[slot], self.mean, self.std)
        self._free[slot].record(cur)        # the device uint8 buffer may be overwritten after the conversion kernel
        self._got += 1
        return x
