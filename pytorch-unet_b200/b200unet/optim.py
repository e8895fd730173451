"""Fused Adam for `b200unet.UNet` (reference: `torch.optim.Adam(model.parameters())`,
README.md:49, run.py:71, LR decay run.py:358-363).

One kernel launch per `step()` updates every parameter tensor (torch's Adam arithmetic, amsgrad=False) AND rewrites the
packed bf16 operand copies the tensor-core kernels read, so the next forward/backward launches no pack kernels
(~90 tiny launches per step otherwise).  `state_dict()` has torch.optim.Adam's schema (`step`, `exp_avg`,
`exp_avg_sq` per parameter), so checkpoints (`run.py:105, 116, 428`) round-trip in both directions.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import AdamJob, check, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    """`FusedAdam(model.parameters(), lr=..., model=model)`.  `model` (a b200unet.UNet, optional) enables the fused
    refresh of its packed weights; without it this is a plain single-launch Adam."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 model=None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=True, differentiable=False, fused=True)
        super().__init__(params, defaults)
        self._model = getattr(model, "module", model)  # accept the DataParallel wrapper
        self._names: Dict[int, str] = {}
        if self._model is not None:
            self._names = {id(p): n for n, p in self._model.named_parameters()}
        self._groups_rt: List[dict] = [dict() for _ in self.param_groups]

    # ------------------------------------------------------------------ state
    def _init_state(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not torch.is_tensor(st["step"]) or st["step"].device != p.device:  # state loaded from a CPU-step checkpoint
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=p.device)
        return st

    def state_dict(self):
        """torch.optim.Adam's schema.  Internally all parameters of a group share ONE device-side step counter; the
        checkpoint gets an independent copy per parameter, as torch.optim.Adam expects when it loads it."""
        sd = super().state_dict()
        sd["state"] = {k: {kk: (vv.clone() if kk == "step" and torch.is_tensor(vv) else vv) for kk, vv in st.items()}
                       for k, st in sd["state"].items()}
        return sd

    # ------------------------------------------------------------------ job table
    def _pack_entries(self, p: torch.Tensor):
        """(fwd entry, dgrad entry) of the model's pack cache for parameter p, or (None, None)."""
        m = self._model
        name = self._names.get(id(p))
        if m is None or name is None or name in m._padspec or p.dim() != 4:
            return None, None
        fwd = None
        for mode in (0, 2):
            e = m._pack_cache.get((name, mode))
            if e is not None and e["ptr"] == p.data_ptr():
                fwd = e
        d = m._pack_cache.get((name, 1))
        if d is not None and d["ptr"] != p.data_ptr():
            d = None
        return fwd, d

    def _build(self, gi: int, plist: List[torch.Tensor]) -> None:
        lib = _lib.load()
        jobs = (AdamJob * len(plist))()
        entries = []
        for j, p in enumerate(plist):
            st = self._init_state(p)
            J = jobs[j]
            J.param, J.grad = p.data_ptr(), p.grad.data_ptr()
            J.exp_avg, J.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            J.numel = p.numel()
            fwd, dg = self._pack_entries(p)
            if fwd is None and dg is None:
                J.kind = 0
            else:
                any_e = fwd if fwd is not None else dg
                J.kind = 2 if any_e["transposed"] else 1
                J.dim0, J.dim1, J.taps = p.shape[0], p.shape[1], p.shape[2] * p.shape[3]
                J.src0_c = any_e["src_c"][0] if not any_e["transposed"] else p.shape[1]
                J.split = int(fwd is not None and fwd["mode"] == 2)
                J.pack_fwd = fwd["tensor"].data_ptr() if fwd is not None else None
                J.pack_dgrad = dg["tensor"].data_ptr() if dg is not None else None
            entries.append((fwd, dg))
        blocks = lib.b200unet_adam_plan(jobs, len(plist))
        if blocks < 0:
            check(blocks, "adam_plan")
        dev = plist[0].device
        rt = self._groups_rt[gi]
        nbytes = C.sizeof(AdamJob) * len(plist)
        if rt.get("table") is None or rt["table"].numel() != nbytes:
            rt["table"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rt["step"] = torch.zeros(1, dtype=torch.float32, device=dev)  # device-side update counter
            rt["step_view"] = None
        st0 = self.state[plist[0]]["step"]
        if st0 is not rt["step_view"]:  # fresh state, or one that load_state_dict() just replaced
            rt["step"].copy_(st0.reshape(1))
            rt["step_view"] = rt["step"].view(())  # the `step` entry of every parameter (torch.optim.Adam schema)
        check(lib.b200unet_adam_upload(rt["table"].data_ptr(), jobs, len(plist), stream_ptr()), "adam_upload")
        for p in plist:
            self.state[p]["step"] = rt["step_view"]
        rt.update({"sig": self._signature(plist), "blocks": blocks, "n": len(plist), "entries": entries})

    def _signature(self, plist):
        sig = []
        for p in plist:
            fwd, dg = self._pack_entries(p)
            st = self.state.get(p)
            sig.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr() if st else 0,
                        id(st["step"]) if st else 0, None if fwd is None else fwd["tensor"].data_ptr(),
                        None if dg is None else dg["tensor"].data_ptr()))
        return sig

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or p.grad.dtype != torch.float32 \
                        or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters and gradients must be contiguous fp32 CUDA tensors")
            while len(self._groups_rt) <= gi:
                self._groups_rt.append(dict())
            rt = self._groups_rt[gi]
            if rt.get("sig") != self._signature(plist):
                self._build(gi, plist)
            rt["step"] += 1
            # hyper-parameters are kernel arguments: lr schedules (run.py:358-363) take effect at the next step()
            check(lib.b200unet_adam_step(rt["table"].data_ptr(), rt["n"], rt["blocks"], float(group["lr"]),
                                         float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                         float(group["weight_decay"]), rt["step"].data_ptr(), stream_ptr()),
                  "adam_step")
            for p, (fwd, dg) in zip(plist, rt["entries"]):
                torch.autograd.graph.increment_version(p)  # the kernel wrote through the raw pointer
                for e in (fwd, dg):
                    if e is not None:
                        e["version"], e["stamped"] = p._version, True
        return loss
