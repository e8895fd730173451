"""Drop-in `UNet` for minghanz/pytorch-unet on B200: same constructor, `forward` contract and `state_dict` schema
as the reference (unet.py:8-84 / unet_original.py:8-75), but forward and backward run on the hand-written sm_100a
kernels of libb200unet.so through `b200unet.ops`.  Nothing here falls back to torch.nn compute: the nn.Conv2d /
nn.BatchNorm2d / nn.ConvTranspose2d children exist only as parameter containers, so that parameter names, shapes,
initialisation order and checkpoints are those of the reference.

Internals: activations are NHWC bf16; the whole network is ONE autograd node (`_UNetFunction`) whose backward is a
hand-scheduled reverse pass, so that
  * the center-crop + concat of the skip connection (unet.py:152-163) is never materialised: the consumer
    convolution reads two windows, its backward-data writes two destinations;
  * every ReLU backward is fused into the kernel that produces the gradient (mask operand), every bias/ReLU
    forward into the convolution epilogue;
  * the gradient of a skip tensor = pool backward + the window written by the decoder, summed in one kernel;
  * weight gradients are produced in reverse-forward order into one flat fp32 arena, bucket by bucket, which is
    what the data-parallel wrapper all-reduces while the rest of the backward is still running.
"""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .input import PackedImages


class _ImagesArg:
    """Carries a PackedImages through torch.autograd.Function.apply (a non-tensor argument: no gradient)."""

    def __init__(self, images: PackedImages):
        self.images = images


# ---------------------------------------------------------------------------------------------------------------
# parameter containers with the reference's module tree (names are part of the checkpoint format)
# ---------------------------------------------------------------------------------------------------------------
class UNetConvBlock(nn.Module):
    """Conv3x3 -> ReLU -> [BatchNorm] -> Conv3x3 -> ReLU -> [BatchNorm]   (unet.py:87-106)."""

    def __init__(self, in_size: int, out_size: int, padding: bool, batch_norm: bool):
        super().__init__()
        block: List[nn.Module] = [nn.Conv2d(in_size, out_size, kernel_size=3, padding=int(padding)), nn.ReLU()]
        if batch_norm:
            block.append(nn.BatchNorm2d(out_size))
        block += [nn.Conv2d(out_size, out_size, kernel_size=3, padding=int(padding)), nn.ReLU()]
        if batch_norm:
            block.append(nn.BatchNorm2d(out_size))
        self.block = nn.Sequential(*block)
        self.batch_norm = batch_norm

    def convs(self) -> Tuple[nn.Conv2d, nn.Conv2d]:
        return (self.block[0], self.block[3 if self.batch_norm else 2])

    def bns(self) -> Tuple[Optional[nn.BatchNorm2d], Optional[nn.BatchNorm2d]]:
        return (self.block[2], self.block[5]) if self.batch_norm else (None, None)


class _UpBase(nn.Module):
    def __init__(self, up_in: int, up_out: int, block_in: int, block_out: int, up_mode: str, padding: bool,
                 batch_norm: bool):
        super().__init__()
        if up_mode == "upconv":
            self.up = nn.ConvTranspose2d(up_in, up_out, kernel_size=2, stride=2)
        else:
            self.up = nn.Sequential(nn.Upsample(mode="bilinear", scale_factor=2), nn.Conv2d(up_in, up_out, kernel_size=1))
        self.conv_block = UNetConvBlock(block_in, block_out, padding, batch_norm)
        self.up_mode = up_mode


class UNetUpBlock(_UpBase):
    """Paper decoder block (unet.py:139-166, unet_original.py:100-127): up halves the channels."""

    def __init__(self, in_size: int, out_size: int, up_mode: str, padding: bool, batch_norm: bool):
        super().__init__(in_size, out_size, in_size, out_size, up_mode, padding, batch_norm)


class UNetUpBlockDeep(_UpBase):
    """Decoder block `unet.py` actually instantiates (unet.py:169-199): up keeps the channels."""

    def __init__(self, in_size: int, skip_in_size: int, out_size: int, up_mode: str, padding: bool, batch_norm: bool):
        super().__init__(in_size, in_size, in_size + skip_in_size, out_size, up_mode, padding, batch_norm)


# ---------------------------------------------------------------------------------------------------------------
# the autograd node
# ---------------------------------------------------------------------------------------------------------------
class _Tape:
    """What one forward call keeps for its backward."""

    def __init__(self):
        self.blocks: Dict[str, dict] = {}
        self.pools: List[dict] = []
        self.ups: List[dict] = []
        self.head: dict = {}
        self.P: Dict[str, torch.Tensor] = {}


def _crop_window(bridge: torch.Tensor, h: int, w: int) -> Tuple[torch.Tensor, int, int]:
    """center_crop (unet.py:152-158) as a strided window, no copy."""
    dy, dx = (bridge.shape[1] - h) // 2, (bridge.shape[2] - w) // 2
    return bridge[:, dy:dy + h, dx:dx + w, :], dy, dx


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: "UNet", x: torch.Tensor, labels: Optional[torch.Tensor], *params: torch.Tensor):
        names = model._param_names
        P = dict(zip(names, params))
        keep = ctx.needs_input_grad[1] or any(ctx.needs_input_grad[3:])  # False under no_grad / nothing requires grad
        tape = _Tape() if keep else None
        out = model._run_forward(x, labels, P, tape)
        ctx.model, ctx.tape, ctx.P, ctx.labels = model, tape, P, labels
        ctx.want_dx = bool(ctx.needs_input_grad[1])
        ctx.x_dtype = getattr(x, "dtype", torch.float32)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        model, tape = ctx.model, ctx.tape
        if tape is None:
            raise RuntimeError("b200unet: backward called on a forward that did not record (no_grad / frozen)")
        ctx.tape = None
        if grad_out is None:
            return (None, None, None) + tuple(None for _ in model._param_names)
        grads = model._run_backward(tape, ctx.P, ctx.labels, grad_out, want_dx=ctx.want_dx)
        dx = grads.pop("__input__", None)
        if dx is not None and dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        return (None, dx, None) + tuple(grads.get(n) for n in model._param_names)


# ---------------------------------------------------------------------------------------------------------------
# the module
# ---------------------------------------------------------------------------------------------------------------
class UNet(nn.Module):
    """U-Net (Ronneberger et al. 2015) with the reference's options.

    Positional arguments are the reference's (unet.py:9-19): in_channels, n_classes, depth, wf, padding, batch_norm,
    up_mode, non_neg.  `up_block` selects the decoder: 'paper' = UNetUpBlock (the graph of unet_original.py, the
    7-argument signature in README.md:14-15) or 'deep' = UNetUpBlockDeep (what unet.py:60 builds); by default it is
    'deep' exactly when the 8th argument `non_neg` is given (the call network_modules.py:73 makes), else 'paper'.
    Output: fp32 NCHW logits, differentiable w.r.t. the parameters and the input.

    `precision` selects how forward activations and conv operands are carried: 'bf16' (one bf16 plane, one tensor-core
    pass) or 'split' (hi + lo bf16 planes, three passes, ~16 mantissa bits; the backward pass is the same bf16 one in
    both).  'auto' = 'split' when batch_norm else 'bf16': measured against the fp32 reference, the BatchNorm graphs miss
    the parity tolerance in plain bf16 (logits 2e-2..1e-1) and meet it in the split tier (3e-5..2e-4), while the
    paper graph without BatchNorm meets it in bf16.
    """

    def __init__(self, in_channels: int = 1, n_classes: int = 2, depth: int = 5, wf: int = 6, padding: bool = False,
                 batch_norm: bool = False, up_mode: str = "upconv", non_neg: Optional[bool] = None,
                 up_block: Optional[str] = None, conv_impl: int = ops.IMPL_AUTO, precision: str = "auto"):
        super().__init__()
        assert up_mode in ("upconv", "upsample")  # unet.py:45
        if up_block is None:
            # Only unet.py's UNet has the 8th argument, and that class always builds UNetUpBlockDeep (unet.py:60): a
            # call that passes `non_neg` (network_modules.py:73 does, positionally) gets that graph, so its checkpoints
            # load.  The 7-argument signature of README.md:14-15 / unet_original.py builds the paper decoder.
            up_block = "deep" if non_neg is not None else "paper"
        non_neg = bool(non_neg)
        assert up_block in ("paper", "deep")
        assert precision in ("auto", "bf16", "split")
        if not 1 <= n_classes <= ops.MAX_CLASSES:
            raise ValueError(f"b200unet.UNet: n_classes must be in 1..{ops.MAX_CLASSES} (the fused classifier / loss "
                             f"kernels keep one accumulator per class in registers), got {n_classes}")
        self.precision = ("split" if batch_norm else "bf16") if precision == "auto" else precision
        self.padding = padding
        self.depth = depth
        self.batch_norm = batch_norm
        self.up_mode = up_mode
        self.non_neg = non_neg
        self.up_block = up_block
        self.n_classes = n_classes
        self.in_channels = in_channels
        self.conv_impl = conv_impl
        prev = in_channels
        self.down_path = nn.ModuleList()
        for i in range(depth):
            self.down_path.append(UNetConvBlock(prev, 2 ** (wf + i), padding, batch_norm))
            prev = 2 ** (wf + i)
        self.up_path = nn.ModuleList()
        for i in reversed(range(depth - 1)):
            if up_block == "paper":
                self.up_path.append(UNetUpBlock(prev, 2 ** (wf + i), up_mode, padding, batch_norm))
                prev = 2 ** (wf + i)
            else:  # unet.py:60 — prev_channels is never updated (unet.py:63 is commented out)
                self.up_path.append(UNetUpBlockDeep(prev, 2 ** (wf + i), prev, up_mode, padding, batch_norm))
        if non_neg:  # unet.py:65-69
            self.last = nn.Sequential(nn.Conv2d(prev, n_classes, kernel_size=1), nn.ReLU())
        else:
            self.last = nn.Conv2d(prev, n_classes, kernel_size=1)
        self._param_names = [n for n, _ in self.named_parameters()]
        self._padspec = self._build_padspec(in_channels, n_classes, depth, wf)
        self._real_shapes = {n: tuple(p.shape) for n, p in self.named_parameters()}
        self._cur_grads: Dict[str, torch.Tensor] = {}
        # hooks for the data-parallel wrapper: grad arena allocator + "these gradients are final" callback
        self._grad_alloc: Optional[Callable[[str, Tuple[int, ...], torch.device], torch.Tensor]] = None
        self._grad_ready: Optional[Callable[[str], None]] = None
        self._grads_done: Optional[Callable[[], None]] = None
        self._pack_cache: Dict[Tuple[str, int], dict] = {}
        self._nbt_pending: List[torch.Tensor] = []
        # backward: the encoder's first-convolution backward-weights launches run on a side stream, next to the HBM-bound pool
        # backward that follows them on the critical path (B200UNET_NO_SIDE_STREAM=1 disables)
        self.side_stream_wgrad = os.environ.get("B200UNET_NO_SIDE_STREAM") is None
        self._side_streams: List[torch.cuda.Stream] = []   # shared with the data-parallel bucketer (it orders reductions after them)
        self._side_keep: List[object] = []
        self._overlap_first = os.environ.get("B200UNET_NO_OVERLAP_FIRST") is None
        self._train_forwards = 0  # training-mode forwards so far: eval-mode packs are reused only within one value

    def __getstate__(self):
        # CUDA streams cannot be pickled / deep-copied (copy.deepcopy(model), torch.save(model)); they are created lazily
        state = dict(self.__dict__)
        state["_side_streams"], state["_side_keep"] = [], []
        return state

    # ---------------------------------------------------------------- public API
    def invalidate_packs(self) -> None:
        """Drop the bf16 operand copies `FusedAdam` maintains.  They are trusted while the parameter keeps its storage
        and autograd version; call this after modifying weights in a way that bumps neither (writes through `.data`,
        custom kernels) while training with FusedAdam."""
        self._pack_cache.clear()

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): parameters move, copies are stale
        self._pack_cache.clear()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """nn.Module.load_state_dict, plus a clear message when the checkpoint belongs to the other decoder variant: the
        two are told apart by the shape of the first up-sampling weight ([C, C/2, ..] paper vs [C, C, ..] Deep)."""
        for key in ("up_path.0.up.weight", "up_path.0.up.1.weight"):
            mine = dict(self.named_parameters()).get(key)
            theirs = state_dict.get(key) if hasattr(state_dict, "get") else None
            if mine is not None and theirs is not None and tuple(mine.shape) != tuple(theirs.shape) \
                    and mine.shape[0] != mine.shape[1] and theirs.shape[0] == theirs.shape[1] and self.up_block == "paper":
                raise RuntimeError("b200unet.UNet: this checkpoint is from the UNetUpBlockDeep graph that unet.py builds "
                                   "(unet.py:60); construct the module with up_block='deep' (or with the 8-argument call)")
            if mine is not None and theirs is not None and tuple(mine.shape) != tuple(theirs.shape) \
                    and mine.shape[0] == mine.shape[1] and theirs.shape[0] != theirs.shape[1] and self.up_block == "deep":
                raise RuntimeError("b200unet.UNet: this checkpoint is from the paper decoder (unet_original.py / "
                                   "UNetUpBlock); construct the module with up_block='paper' (or the 7-argument call)")
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._apply_fn(x, None)

    def loss(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """F.cross_entropy(self(x), target) (README.md:57-58) with the classifier fused into the loss kernel: the
        logits never touch HBM.  Returns the scalar mean loss (ignore_index = -100)."""
        return self._apply_fn(x, target)

    def _apply_fn(self, x, labels):
        if not x.is_cuda:
            raise RuntimeError("b200unet.UNet runs on CUDA (sm_100a) only: there is no CPU path")
        if isinstance(x, PackedImages):  # b200unet.input: the batch is already the first convolution's operand
            params = [p for _, p in self.named_parameters()]
            return _UNetFunction.apply(self, _ImagesArg(x), labels, *params)
        params = [p for _, p in self.named_parameters()]
        return _UNetFunction.apply(self, x, labels, *params)

    # ---------------------------------------------------------------- internal channel padding
    # The tensor-core kernels need NHWC channel counts that are multiples of 8 (TMA strides) and the first-layer
    # kernels multiples of 16.  Narrow variants (the repo's feature net: wf=2 -> 4 / 8 channels, options.py:19-25)
    # therefore run with their 4- and 8-channel tensors padded to 16 channels INSIDE the module: parameters are
    # zero-padded copies made per call, pad channels carry exact zeros through conv/ReLU/BN/pool/upsample, gradients
    # are un-padded before they are returned.  state_dict, parameter shapes and results are unchanged; for the paper
    # widths (64..1024) the spec is empty and none of this code runs.
    @staticmethod
    def _cpad(c: int) -> int:
        return c if (c % 8 == 0 and c >= 16) else max(16, (c + 7) // 8 * 8)

    @staticmethod
    def _cpad_image(c: int) -> int:
        """Channel count the image is carried with: 1..4 channels go to the first-layer CUDA-core kernels as they are;
        5, 6, 7, 9, ... are zero-padded to a multiple of 8 so that the tensor-core kernels take the first layer (the
        reference accepts any in_channels, unet.py:49-52)."""
        return c if (c <= 4 or c % 8 == 0) else (c + 7) // 8 * 8

    def _build_padspec(self, in_channels, n_classes, depth, wf):
        spec: Dict[str, tuple] = {}
        cp = self._cpad

        def conv(name, srcs_real, cout, pad_src=True, srcs_pad=None):
            if srcs_pad is None:
                srcs_pad = [cp(c) if pad_src else self._cpad_image(c) for c in srcs_real]
            segs, ro, po = [], 0, 0
            for r, pd in zip(srcs_real, srcs_pad):
                segs.append((ro, r, po))
                ro, po = ro + r, po + pd
            if cp(cout) != cout or po != ro:
                spec[name + ".weight"] = ([(0, cout, 0)], cp(cout), segs, po)
            if cp(cout) != cout:
                spec[name + ".bias"] = ([(0, cout, 0)], cp(cout), None, None)

        def block(prefix, srcs_real, cout, first=False, srcs_pad=None):
            i2, b1, b2 = (3, 2, 5) if self.batch_norm else (2, None, None)
            conv(f"{prefix}.block.0", srcs_real, cout, pad_src=not first, srcs_pad=srcs_pad)
            conv(f"{prefix}.block.{i2}", [cout], cout)
            if self.batch_norm and cp(cout) != cout:
                for b in (b1, b2):
                    spec[f"{prefix}.block.{b}.weight"] = ([(0, cout, 0)], cp(cout), None, None)
                    spec[f"{prefix}.block.{b}.bias"] = ([(0, cout, 0)], cp(cout), None, None)

        prev = in_channels
        for i in range(depth):
            block(f"down_path.{i}", [prev], 2 ** (wf + i), first=(i == 0))
            prev = 2 ** (wf + i)
        for j, i in enumerate(reversed(range(depth - 1))):
            skip = 2 ** (wf + i)
            up_out, blk_out = (skip, skip) if self.up_block == "paper" else (prev, prev)
            up_pad = cp(up_out)
            if self.up_mode == "upconv":  # weight [cin, cout, 2, 2]
                if self.precision == "split" and up_pad % 32:
                    # the transposed convolution of the split tier exists on the tensor-core kernel only, whose four
                    # output quadrants must each be a whole number of 32-column epilogue passes
                    up_pad = (up_pad + 31) // 32 * 32
                if cp(prev) != prev or up_pad != up_out:
                    spec[f"up_path.{j}.up.weight"] = ([(0, prev, 0)], cp(prev), [(0, up_out, 0)], up_pad)
                if up_pad != up_out:
                    spec[f"up_path.{j}.up.bias"] = ([(0, up_out, 0)], up_pad, None, None)
            else:
                conv(f"up_path.{j}.up.1", [prev], up_out)
            block(f"up_path.{j}.conv_block", [up_out, skip], blk_out, srcs_pad=[up_pad, cp(skip)])
            prev = blk_out
        if cp(prev) != prev:
            spec[("last.0" if self.non_neg else "last") + ".weight"] = ([(0, n_classes, 0)], n_classes, [(0, prev, 0)], cp(prev))
        return spec

    @staticmethod
    def _pad_tensor(t: torch.Tensor, sp: tuple) -> torch.Tensor:
        segs0, size0, segs1, size1 = sp
        shape = list(t.shape)
        shape[0] = size0
        if segs1 is not None:
            shape[1] = size1
        out = torch.zeros(shape, dtype=t.dtype, device=t.device)
        for r0, l0, p0 in segs0:
            if segs1 is None:
                out[p0:p0 + l0] = t[r0:r0 + l0]
            else:
                for r1, l1, p1 in segs1:
                    out[p0:p0 + l0, p1:p1 + l1] = t[r0:r0 + l0, r1:r1 + l1]
        return out

    @staticmethod
    def _unpad_into(dst: torch.Tensor, padded: torch.Tensor, sp: tuple) -> None:
        segs0, _, segs1, _ = sp
        for r0, l0, p0 in segs0:
            if segs1 is None:
                dst[r0:r0 + l0] = padded[p0:p0 + l0]
            else:
                for r1, l1, p1 in segs1:
                    dst[r0:r0 + l0, r1:r1 + l1] = padded[p0:p0 + l0, p1:p1 + l1]

    def _padded_params(self, P: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if not self._padspec:
            return P
        return {k: (self._pad_tensor(v.detach(), self._padspec[k]) if k in self._padspec else v) for k, v in P.items()}

    # ---------------------------------------------------------------- helpers
    def _packed(self, name: str, w: torch.Tensor, mode: int, src_c=None, transposed_conv: bool = False) -> torch.Tensor:
        """Packed bf16 operand copy of weight `name` (mode: 0 forward, 1 backward-data, 2 forward of the split tier).
        The buffer of a (name, mode) entry is persistent — a repack writes into it — so that `FusedAdam` can rewrite
        it inside its own kernel.  An entry is REUSED without repacking only if the parameter still has that storage
        and autograd version AND either FusedAdam stamped it after its last update, or the module is in eval mode and
        the entry was packed in eval mode with no training-mode forward since (inference: weights are static).  In
        training with any other optimizer every call repacks — torch's own fused Adam, for one, updates parameters
        without bumping their version, so a cache keyed on versions alone would serve stale weights.  Internally padded
        parameters are fresh copies per call."""
        if name in self._padspec:
            return (ops.pack_convt_weight(w.detach(), mode) if transposed_conv
                    else ops.pack_conv_weight(w.detach(), src_c if src_c is not None else [w.shape[1]], mode))
        key = (name, mode)
        e = self._pack_cache.get(key)
        if e is not None and e["version"] == w._version and e["ptr"] == w.data_ptr() and (
                e["stamped"] or (not self.training and e["eval_epoch"] == self._train_forwards)):
            return e["tensor"]
        out = e["tensor"] if e is not None and e["tensor"].device == w.device else None
        if transposed_conv:
            packed = ops.pack_convt_weight(w.detach(), mode, out=out)
        else:
            src_c = list(src_c) if src_c is not None else [w.shape[1]]
            packed = ops.pack_conv_weight(w.detach(), src_c, mode, out=out)
        new = {"version": w._version, "ptr": w.data_ptr(), "tensor": packed, "mode": mode,
               "src_c": src_c, "transposed": transposed_conv, "stamped": False,
               "eval_epoch": self._train_forwards if not self.training else -1}
        if e is not None:
            e.update(new)  # in place: FusedAdam keeps references to the entries it stamps
        else:
            self._pack_cache[key] = new
        return packed

    def _side_stream(self, device) -> torch.cuda.Stream:
        if not self._side_streams:
            self._side_streams.append(torch.cuda.Stream(device=device))
        return self._side_streams[0]

    def _new_grad(self, name: str, like: torch.Tensor) -> torch.Tensor:
        """Gradient buffer for parameter `name` (`like` has the shape the kernels produce, i.e. the padded one)."""
        if name in self._padspec:  # kernels write the padded gradient into a scratch tensor, _done() un-pads it
            return torch.empty(like.shape, dtype=torch.float32, device=like.device)
        if self._grad_alloc is not None:
            return self._grad_alloc(name, tuple(like.shape), like.device)
        return torch.empty(like.shape, dtype=torch.float32, device=like.device)

    def _done(self, *names: str) -> None:
        for n in names:
            if n in self._padspec:
                real_shape = self._real_shapes[n]
                padded = self._cur_grads[n]
                if self._grad_alloc is not None:
                    real = self._grad_alloc(n, real_shape, padded.device)
                else:
                    real = torch.empty(real_shape, dtype=torch.float32, device=padded.device)
                self._unpad_into(real, padded, self._padspec[n])
                self._cur_grads[n] = real
            if self._grad_ready is not None:
                self._grad_ready(n)

    # ---------------------------------------------------------------- forward
    def _conv(self, name, srcs, P, pad, relu=True):
        w, b = P[name + ".weight"], P[name + ".bias"]
        mode = 2 if isinstance(srcs[0], ops.Split) else 0
        return ops.conv_fwd(srcs, w.detach(), b.detach(), pad, relu, impl=self.conv_impl,
                            w_packed=lambda: self._packed(name + ".weight", w, mode, [s.shape[3] for s in srcs]))

    def _first_layer_patches(self, prefix: str, srcs, w) -> bool:
        return (prefix == "down_path.0" and self.conv_impl == ops.IMPL_AUTO and len(srcs) == 1
                and srcs[0].shape[3] * 9 <= 64 and w.shape[0] % 16 == 0 and ops.tensor_cores_available())

    def _block_forward(self, prefix: str, blk: UNetConvBlock, srcs, P, tape):
        """UNetConvBlock.forward (unet.py:104-106).  Returns (output, post-ReLU activation of the 2nd conv)."""
        pad = int(self.padding)
        rec = {"srcs": [ops.hi_of(s) for s in srcs]}  # the backward pass reads the hi planes only (b200unet.h)
        names = [f"{prefix}.block.0", f"{prefix}.block.{3 if blk.batch_norm else 2}"]
        bn_names = [f"{prefix}.block.2", f"{prefix}.block.5"]
        cur = srcs
        for i in range(2):
            a = self._conv(names[i], cur, P, pad)
            rec[f"a{i}"] = ops.hi_of(a)
            if blk.batch_norm:
                bn = blk.bns()[i]
                g, bt = P[bn_names[i] + ".weight"].detach(), P[bn_names[i] + ".bias"].detach()
                rm, rv = bn.running_mean, bn.running_var
                padded_stats = rm is not None and rm.numel() != a.shape[3]
                if padded_stats:  # internal channel padding: the kernels see padded copies of the running statistics
                    extra = a.shape[3] - rm.numel()
                    rm = torch.cat([rm, rm.new_zeros(extra)])
                    rv = torch.cat([rv, rv.new_ones(extra)])
                if self.training or bn.running_mean is None:
                    o, mean, invstd = ops.bn_fwd_train(a, g, bt, rm if self.training else None,
                                                       rv if self.training else None,
                                                       bn.momentum if bn.momentum is not None else 0.1, bn.eps)
                    if self.training and padded_stats:
                        bn.running_mean.copy_(rm[:bn.running_mean.numel()])
                        bn.running_var.copy_(rv[:bn.running_var.numel()])
                    if self.training and bn.num_batches_tracked is not None:
                        self._nbt_pending.append(bn.num_batches_tracked)  # one fused increment per forward
                    rec[f"bn{i}"] = (mean, invstd)
                else:
                    o = ops.bn_fwd_eval(a, g, bt, rm, rv, bn.eps)
                    # backward through frozen statistics (fine-tuning in eval mode): torch's BatchNorm backward with
                    # training=False, i.e. mean / invstd are constants
                    rec[f"bn{i}"] = (rm, torch.rsqrt(rv + bn.eps)) if tape is not None else None
                    rec[f"bn{i}_frozen"] = True
            else:
                o = a
            rec[f"o{i}"] = ops.hi_of(o)
            cur = [o]
        if tape is not None:
            tape.blocks[prefix] = rec
        return cur[0], rec["a1"]

    def _run_forward(self, x, labels, P, tape):
        self._nbt_pending = []
        if self.training:
            self._train_forwards += 1
        P = self._padded_params(P)
        if tape is not None:
            tape.P = P
        if isinstance(x, _ImagesArg):
            cur = x.images.data
        else:
            if x.dtype != torch.float32:
                x = x.float()
            cur = ops.to_nhwc(x, split=self.precision == "split", c_pad=self._cpad_image(x.shape[1]))
        bridges = []
        last_act = None
        for i, down in enumerate(self.down_path):
            cur, last_act = self._block_forward(f"down_path.{i}", down, [cur], P, tape)
            if i != self.depth - 1:
                bridges.append((cur, last_act))
                pooled, idx8 = ops.maxpool_fwd(cur)  # unet.py:79
                if tape is not None:
                    tape.pools.append({"idx8": idx8, "src_shape": cur.shape, "act": last_act, "pooled": pooled})
                cur = pooled
        for j, up in enumerate(self.up_path):
            bridge, _ = bridges[-j - 1]
            rec = {"x": ops.hi_of(cur), "x_act": last_act}
            if self.up_mode == "upconv":
                w, b = P[f"up_path.{j}.up.weight"].detach(), P[f"up_path.{j}.up.bias"].detach()
                wname, wparam, mode = f"up_path.{j}.up.weight", P[f"up_path.{j}.up.weight"], 2 if isinstance(cur, ops.Split) else 0
                upv = ops.convt_fwd(cur, w, b, impl=self.conv_impl,
                                    w_packed=lambda: self._packed(wname, wparam, mode, transposed_conv=True))
            else:
                # Sequential(Upsample(bilinear, x2), Conv2d 1x1) (unet.py:145-148, 175-178) evaluated as conv THEN
                # upsample: a 1x1 convolution is per-pixel linear and the interpolation weights sum to one, so the two
                # commute exactly (bias included; SURVEY.md 8c measured 2.4e-7 in fp32) — the GEMM runs on a quarter
                # of the pixels and the full-resolution tensor is written once instead of written, read and written.
                v = self._conv(f"up_path.{j}.up.1", [cur], P, 0, relu=False)
                upv = ops.bilinear_fwd(v)
            win, dy, dx = _crop_window(bridge, upv.shape[1], upv.shape[2])
            rec.update({"crop": (dy, dx), "bridge_shape": bridge.shape, "up_shape": upv.shape})
            if tape is not None:
                tape.ups.append(rec)
            cur, last_act = self._block_forward(f"up_path.{j}.conv_block", up.conv_block, [upv, win], P, tape)
        if self._nbt_pending:  # BatchNorm2d.num_batches_tracked += 1 for every layer, as one multi-tensor launch
            torch._foreach_add_(self._nbt_pending, 1)
            self._nbt_pending = []
        hname = "last.0" if self.non_neg else "last"
        hw, hb = P[hname + ".weight"].detach(), P[hname + ".bias"].detach()
        hw2 = hw.view(hw.shape[0], hw.shape[1])
        if tape is not None:
            tape.head = {"x": ops.hi_of(cur), "act": last_act}
        if labels is None:
            return ops.head_fwd(cur, hw2, hb, self.non_neg)
        loss, state, _ = ops.head_ce_fwd(cur, hw2, hb, self.non_neg, labels.contiguous())
        if tape is not None:
            tape.head["state"] = state
        return loss

    # ---------------------------------------------------------------- backward
    def _block_backward(self, prefix: str, blk: UNetConvBlock, tape, P, g, grads, src_dsts, src_masks,
                        overlap_wgrad: bool = False):
        """Reverse of _block_forward.  `g` = gradient w.r.t. the block output (already ReLU-masked when there is no
        BatchNorm).  src_dsts: destination tensors for the gradient of each source (None = not needed)."""
        rec = tape.blocks.pop(prefix)
        pad = int(self.padding)
        names = [f"{prefix}.block.0", f"{prefix}.block.{3 if blk.batch_norm else 2}"]
        bn_names = [f"{prefix}.block.2", f"{prefix}.block.5"]
        for i in (1, 0):
            a = rec[f"a{i}"]
            if blk.batch_norm:
                mean, invstd = rec[f"bn{i}"]
                gam = P[bn_names[i] + ".weight"]
                dgam = self._new_grad(bn_names[i] + ".weight", gam)
                dbet = self._new_grad(bn_names[i] + ".bias", gam)
                dz, _, _ = ops.bn_bwd(a, g, gam.detach(), mean, invstd, True, dx=g, dgamma=dgam, dbeta=dbet,
                                      frozen_stats=rec.get(f"bn{i}_frozen", False))
                grads[bn_names[i] + ".weight"], grads[bn_names[i] + ".bias"] = dgam, dbet
            else:
                dz = g
            w = P[names[i] + ".weight"]
            srcs = rec["srcs"] if i == 0 else [rec["o0"]]
            dw = self._new_grad(names[i] + ".weight", w)
            db = self._new_grad(names[i] + ".bias", P[names[i] + ".bias"])
            side = None
            if (i == 1 and prefix == "down_path.0" and self.side_stream_wgrad and self._overlap_first and dz.is_cuda
                    and not blk.batch_norm and (names[i] + ".weight") not in self._padspec
                    and self._first_layer_patches(prefix, rec["srcs"], P[names[0] + ".weight"])):
                # last block of the backward pass: its second convolution's backward-weights (tensor-bound) goes to the side
                # stream, next to the first layer's backward-weights that follows on this one (im2col + a K = 16 GEMM over the
                # full-resolution gradient: HBM-bound)
                side = self._side_stream(dz.device)
            if (i == 0 and overlap_wgrad and src_dsts is not None and self.side_stream_wgrad and dz.is_cuda
                    and (names[i] + ".weight") not in self._padspec):  # padded gradients are un-padded on this stream at once
                # critical path first: backward-data of this convolution feeds the pool backward of the level above, which
                # is HBM-bound; this convolution's backward-weights (tensor-bound, nobody waits for it) then runs next to
                # it on a side stream instead of in front of it
                side = self._side_stream(dz.device)
            if i == 0 and self._first_layer_patches(prefix, srcs, w):
                # 1..7 input channels: the 3x3 patch (9*Cin values) fits one 64-wide K chunk, so backward-weights of
                # the first layer runs as a 1x1 problem on the tensor-core kernel over the im2col'ed image (built here,
                # 2*Kp bytes per pixel, dropped right after).  Measured 0.77 ms vs 1.24 ms for the CUDA-core kernel at
                # batch 32; the forward stays on the first-layer kernel.
                patches = ops.im2col3x3(srcs[0], pad)
                dw1, _ = ops.conv_wgrad(dz, [patches], 1, 0, impl=self.conv_impl, db=db)
                dw.copy_(dw1[:, :w.shape[1] * 9, 0, 0].reshape(w.shape))
                del patches
            elif side is None:
                ops.conv_wgrad(dz, srcs, 3, pad, impl=self.conv_impl, dw=dw, db=db)
            grads[names[i] + ".weight"], grads[names[i] + ".bias"] = dw, db
            if i == 1:
                g = torch.empty_like(rec["o0"])
                ops.conv_dgrad(dz, w.detach(), pad, [g], [None if blk.batch_norm else rec["a0"]], impl=self.conv_impl,
                               w_packed=lambda: self._packed(names[i] + ".weight", w, 1))
            elif src_dsts is not None:
                cin_dst = sum(d.shape[3] for d in src_dsts)
                if cin_dst != w.shape[1]:
                    # gradient w.r.t. the image (unet.py:73-84 is differentiable in x): the 1..4-channel image gradient is
                    # produced with its channels zero-padded to 8, which is what the tensor-core kernel can store
                    wp = torch.zeros((w.shape[0], cin_dst, w.shape[2], w.shape[3]), dtype=w.dtype, device=w.device)
                    wp[:, :w.shape[1]] = w.detach()
                    ops.conv_dgrad(dz, wp, pad, src_dsts, src_masks, impl=self.conv_impl)
                else:
                    ops.conv_dgrad(dz, w.detach(), pad, src_dsts, src_masks, impl=self.conv_impl,
                                   w_packed=lambda: self._packed(names[i] + ".weight", w, 1))
            if side is not None:
                main = torch.cuda.current_stream(dz.device)
                side.wait_stream(main)       # dz (and everything before it) is complete for the side stream
                with torch.cuda.stream(side):
                    ops.conv_wgrad(dz, srcs, 3, pad, impl=self.conv_impl, dw=dw, db=db)
                # the caching allocator must not hand these buffers to main-stream work before the side kernel has read them
                self._side_keep.extend([dz, dw, db] + [ops.hi_of(t) for t in srcs])
            self._done(*( [bn_names[i] + ".weight", bn_names[i] + ".bias"] if blk.batch_norm else [] ),
                       names[i] + ".weight", names[i] + ".bias")

    def _run_backward(self, tape, P, labels, grad_out, want_dx: bool = False) -> Dict[str, torch.Tensor]:
        P = tape.P
        grads: Dict[str, torch.Tensor] = {}
        self._cur_grads = grads
        bn = self.batch_norm
        hname = "last.0" if self.non_neg else "last"
        hw, hb = P[hname + ".weight"], P[hname + ".bias"]
        hw2 = hw.detach().view(hw.shape[0], hw.shape[1])
        hx, hact = tape.head["x"], tape.head["act"]
        g = torch.empty_like(hx)
        mask = None if bn else hact
        dwh = self._new_grad(hname + ".weight", hw)
        dbh = self._new_grad(hname + ".bias", hb)
        if labels is None:
            ops.head_bwd(hx, hw2, hb.detach(), self.non_neg, grad_out.float(), dx=g, mask=mask, dw=dwh, db=dbh)
        else:
            gs = grad_out.detach().reshape(1).float()
            ops.head_ce_bwd(hx, hw2, hb.detach(), self.non_neg, labels.contiguous(), tape.head["state"],
                            grad_scale=gs, dx=g, mask=mask, dw=dwh, db=dbh)
        grads[hname + ".weight"], grads[hname + ".bias"] = dwh, dbh
        self._done(hname + ".weight", hname + ".bias")

        bridge_grads: Dict[int, Tuple[torch.Tensor, Tuple[int, int, int, int]]] = {}
        # decoder, deepest-first order reversed
        for j in reversed(range(len(self.up_path))):
            up = self.up_path[j]
            rec = tape.ups[j]
            level = self.depth - 2 - j  # index of the bridge in down_path
            d_up = torch.empty(rec["up_shape"], dtype=torch.bfloat16, device=g.device)
            gb = torch.empty(rec["bridge_shape"], dtype=torch.bfloat16, device=g.device)
            dy, dx = rec["crop"]
            win = gb[:, dy:dy + d_up.shape[1], dx:dx + d_up.shape[2], :]
            bridge_grads[level] = (gb, (dy, dx, d_up.shape[1], d_up.shape[2]))
            # without BatchNorm the bridge's ReLU mask is applied right here, by the backward-data kernel that writes the
            # skip gradient (its mask operand is a window of the bridge activation laid out like `win`): the pool backward
            # below then never reads the full-resolution mask (it masks the scattered term with [pooled > 0])
            # (only where that backward-data launch is compute-bound, i.e. >= 128 channels: on the 64-channel level the extra
            # mask read made it 0.49 -> 0.66 ms, more than the pool backward saves)
            premask = (not bn) and gb.shape[3] >= 128
            tape.pools[level]["premasked"] = premask
            bmask = tape.pools[level]["act"][:, dy:dy + d_up.shape[1], dx:dx + d_up.shape[2], :] if premask else None
            self._block_backward(f"up_path.{j}.conv_block", up.conv_block, tape, P, g, grads, [d_up, win], [None, bmask])
            xin, xact = rec["x"], rec["x_act"]
            g = torch.empty_like(xin)
            xmask = None if bn else xact
            if self.up_mode == "upconv":
                wn = f"up_path.{j}.up"
                w = P[wn + ".weight"]
                dw = self._new_grad(wn + ".weight", w)
                db = self._new_grad(wn + ".bias", P[wn + ".bias"])
                if self.side_stream_wgrad and d_up.is_cuda and (wn + ".bias") not in self._padspec:
                    # the bias gradient (a column sum over d_up: HBM-bound) runs on the side stream next to the L2-fed
                    # tensor-core kernels below instead of after them
                    side = self._side_stream(d_up.device)
                    side.wait_stream(torch.cuda.current_stream(d_up.device))
                    with torch.cuda.stream(side):
                        ops.channel_sum(d_up, out=db)
                    self._side_keep.extend([d_up, db])
                    ops.convt_wgrad(xin, d_up, want_db=False, impl=self.conv_impl, dw=dw)
                else:
                    ops.convt_wgrad(xin, d_up, impl=self.conv_impl, dw=dw, db=db)
                ops.convt_dgrad(d_up, w.detach(), g, mask=xmask, impl=self.conv_impl,
                                w_packed=lambda: self._packed(wn + ".weight", w, 1, transposed_conv=True))
            else:
                wn = f"up_path.{j}.up.1"
                w = P[wn + ".weight"]
                dw = self._new_grad(wn + ".weight", w)
                db = self._new_grad(wn + ".bias", P[wn + ".bias"])
                # reverse of conv-then-upsample: gradient back to low resolution first, 1x1 backward there
                gv = torch.empty((xin.shape[0], xin.shape[1], xin.shape[2], d_up.shape[3]), dtype=torch.bfloat16,
                                 device=g.device)
                ops.bilinear_bwd(d_up, gv)
                ops.conv_wgrad(gv, [xin], 1, 0, impl=self.conv_impl, dw=dw, db=db)
                ops.conv_dgrad(gv, w.detach(), 0, [g], [xmask], impl=self.conv_impl,
                               w_packed=lambda: self._packed(wn + ".weight", w, 1))
            grads[wn + ".weight"], grads[wn + ".bias"] = dw, db
            self._done(wn + ".weight", wn + ".bias")
            rec.clear()
        # encoder
        for i in reversed(range(self.depth)):
            down = self.down_path[i]
            if i != self.depth - 1:
                pool = tape.pools[i]
                gb, (dy, dx, wh, ww) = bridge_grads.pop(i)
                add = gb[:, dy:dy + wh, dx:dx + ww, :]
                if pool.get("premasked"):
                    ops.maxpool_bwd(g, pool["idx8"], gb, add=add, add_y=dy, add_x=dx, pooled=pool["pooled"])
                else:
                    ops.maxpool_bwd(g, pool["idx8"], gb, add=add, add_y=dy, add_x=dx, mask=None if bn else pool["act"])
                g = gb
            if i == 0 and want_dx:
                ximg = tape.blocks["down_path.0"]["srcs"][0]
                gx = torch.empty(tuple(ximg.shape[:3]) + (max(8, (ximg.shape[3] + 7) // 8 * 8),), dtype=torch.bfloat16,
                                 device=g.device)
                self._block_backward(f"down_path.{i}", down, tape, P, g, grads, [gx], [None])
                grads["__input__"] = ops.to_nchw(gx)[:, :self.in_channels]
            elif i == 0:
                self._block_backward(f"down_path.{i}", down, tape, P, g, grads, None, None)
            else:
                prev_pooled_shape = tape.blocks[f"down_path.{i}"]["srcs"][0].shape
                gp = torch.empty(prev_pooled_shape, dtype=torch.bfloat16, device=g.device)
                self._block_backward(f"down_path.{i}", down, tape, P, g, grads, [gp], [None], overlap_wgrad=True)
                g = gp
        if self._side_streams and self._side_keep:
            torch.cuda.current_stream(g.device).wait_stream(self._side_streams[0])  # join: every gradient is complete
            self._side_keep = []
        if self._grads_done is not None:
            self._grads_done()
        self._cur_grads = {}
        return grads
