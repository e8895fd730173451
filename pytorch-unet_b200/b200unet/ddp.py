"""Batch-sharded data parallelism for `b200unet.UNet` (one process per GPU, torch.distributed over NCCL / NVLink).

The reference has no distributed code; the semantics implemented are those of stock DDP: every
rank holds a full replica, runs the step on its own images, and weight gradients are averaged over ranks.
BatchNorm statistics stay per rank.

Mechanics (this is why the U-Net backward is one hand-scheduled autograd node): the backward pass produces weight
gradients in reverse-forward order straight into ONE flat fp32 arena whose layout follows that order, so a bucket
is a contiguous slice that needs no flatten/copy.  As soon as the last gradient of a bucket has been written, an
asynchronous all-reduce of that slice is enqueued (NCCL runs it on its own stream, ordered after the kernels that
wrote it) and overlaps the rest of the backward; the end of backward waits for the outstanding handles.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def backward_order(model) -> List[str]:
    """Parameter names in the order the hand-scheduled backward finishes them (head, decoder deep->shallow... in
    reverse forward order; inside a layer weight before bias)."""
    names = list(model._param_names)
    # forward order groups are (weight, bias) pairs; reversing pairs keeps weight before bias
    pairs = [names[i:i + 2] for i in range(0, len(names), 2)]
    out: List[str] = []
    for p in reversed(pairs):
        out.extend(p)
    return out


class GradBucketer:
    """Flat gradient arena + bucketed asynchronous all-reduce.  Independent of the model so that the bucket logic
    can be exercised on CPU with the gloo backend."""

    def __init__(self, names: Sequence[str], shapes: Dict[str, Tuple[int, ...]], bucket_bytes: int = 32 << 20,
                 process_group=None, average: bool = True, grad_dtype: str = "fp32"):
        assert grad_dtype in ("fp32", "bf16")
        self.grad_dtype = grad_dtype  # wire precision of the all-reduce (the arena and the optimizer stay fp32)
        self.profile = False          # bench.py: measure how long the end of backward waits for the collectives
        # deferred mode (b200unet.GraphedTrainStep): backward only fills the arena; the caller reduces it in ONE collective
        # between two CUDA graphs (a NCCL call issued from inside a stream capture deadlocked with the async bucket logic)
        self.defer = False
        # streams other than the current one that write gradients (UNet's side stream for overlapped backward-weights): a
        # bucket's reduction is ordered after everything enqueued on them so far
        self.sync_streams: list = []
        self.deferred_arena: Optional[torch.Tensor] = None
        self._prof: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self.names = list(names)
        self.shapes = dict(shapes)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # ReduceOp.AVG exists in NCCL only; other backends (gloo: CPU tests, single-GPU multi-process tests) sum and divide
        self._nccl = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        self.average = average
        self.offsets: Dict[str, Tuple[int, int]] = {}
        self.bucket_of: Dict[str, int] = {}
        self.buckets: List[Tuple[int, int, int]] = []  # (start, end, n_params)
        off = 0
        b_start, b_count = 0, 0
        for n in self.names:
            numel = 1
            for s in self.shapes[n]:
                numel *= s
            # keep every slice 16-byte aligned
            self.offsets[n] = (off, numel)
            self.bucket_of[n] = len(self.buckets)
            off += (numel + 3) // 4 * 4
            b_count += 1
            if (off - b_start) * 4 >= bucket_bytes:
                self.buckets.append((b_start, off, b_count))
                b_start, b_count = off, 0
        if b_count:
            self.buckets.append((b_start, off, b_count))
        self.total = off
        self.arena: Optional[torch.Tensor] = None
        self._pending: List[int] = []
        self._handles: list = []

    # ---- called by the model's backward
    def begin(self, device) -> None:
        self.arena = torch.empty(self.total, dtype=torch.float32, device=device)
        self._pending = [c for (_, _, c) in self.buckets]
        self._handles = []

    def alloc(self, name: str, shape: Tuple[int, ...], device) -> torch.Tensor:
        if self.arena is None:
            self.begin(device)
        off, numel = self.offsets[name]
        return self.arena[off:off + numel].view(shape)

    def ready(self, name: str) -> None:
        b = self.bucket_of[name]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._reduce(b)

    def _reduce(self, b: int) -> None:
        if self.world <= 1 or self.defer:
            return
        s, e, _ = self.buckets[b]
        buf = self.arena[s:e]
        if buf.is_cuda:
            for st in self.sync_streams:
                torch.cuda.current_stream(buf.device).wait_stream(st)
        if self.grad_dtype == "bf16" and buf.is_cuda and self._nccl:
            # half the bytes on the wire: the bucket travels as bf16 and is widened back after the reduction
            wire = buf.to(torch.bfloat16)
            h = dist.all_reduce(wire, op=dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM, group=self.pg,
                                async_op=True)
            self._handles.append((h, None, (buf, wire)))
        elif self.average and buf.is_cuda and self._nccl:
            self._handles.append((dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg, async_op=True), None, None))
        else:
            h = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._handles.append((h, buf if self.average else None, None))

    def reduce_all(self, arena: torch.Tensor) -> None:
        """One blocking-on-stream all-reduce (average) of a whole arena: the collective of the graphed step."""
        if self.world <= 1:
            return
        if self._nccl and self.average:
            dist.all_reduce(arena, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=self.pg)
            if self.average:
                arena.div_(self.world)

    def finish(self) -> None:
        if self.defer:
            self.deferred_arena, self.arena, self._handles = self.arena, None, []
            return
        prof = self.profile and self._handles and self.arena is not None and self.arena.is_cuda \
            and not torch.cuda.is_current_stream_capturing()
        if prof:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for h, scale_buf, widen in self._handles:
            h.wait()
            if scale_buf is not None:
                scale_buf.div_(self.world)
            if widen is not None:
                widen[0].copy_(widen[1])
        if prof:
            e1.record()
            self._prof.append((e0, e1))
        self._handles = []
        self.arena = None  # the views handed out keep the storage alive

    def profile_summary(self) -> dict:
        """Mean time per step the compute stream spent waiting for outstanding all-reduces at the end of backward
        (the part of the collective that did NOT overlap), from CUDA events on the compute stream."""
        if not self._prof:
            return {"buckets": len(self.buckets), "bucket_mb": [round((e - s) * 4 / 2 ** 20, 2) for s, e, _ in self.buckets]}
        torch.cuda.synchronize()
        waits = [a.elapsed_time(b) for a, b in self._prof]
        self._prof = []
        return {"exposed_allreduce_wait_ms": sum(waits) / len(waits), "max_ms": max(waits), "steps": len(waits),
                "buckets": len(self.buckets), "bucket_mb": [round((e - s) * 4 / 2 ** 20, 2) for s, e, _ in self.buckets],
                "grad_dtype": self.grad_dtype}


class DataParallel(torch.nn.Module):
    """`DataParallel(UNet(...).cuda())`: same call surface as the wrapped module (`forward`, `loss`)."""

    def __init__(self, module, bucket_bytes: int = 32 << 20, process_group=None, broadcast: bool = True,
                 grad_dtype: str = "fp32"):
        super().__init__()
        self.module = module
        if dist.is_initialized() and broadcast:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        shapes = {n: tuple(p.shape) for n, p in module.named_parameters()}
        self.bucketer = GradBucketer(backward_order(module), shapes, bucket_bytes, process_group, grad_dtype=grad_dtype)
        module._grad_alloc = self.bucketer.alloc
        module._grad_ready = self.bucketer.ready
        module._grads_done = self.bucketer.finish
        self.bucketer.sync_streams = module._side_streams  # the same list object: filled lazily by the module

    def forward(self, x):
        return self.module(x)

    def loss(self, x, target):
        return self.module.loss(x, target)
