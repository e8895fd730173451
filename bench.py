#!/usr/bin/env python
"""Headline benchmark: training images/s of the reference U-Net on synthetic images, B200-native path vs the reference's
CPU path, with the roofline evidence in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--batch B]

Default workload = BASELINE.json configs[2] ("config 3"): paper-default U-Net (in=1, n_classes=2, depth 5, wf 6, valid
padding, upconv), 1x572x572, batch 32 per GPU — it fits one B200, so it is the N=1 workload too — data-parallel over N
GPUs of one node.  --config 2 / 4 / 5 select the other BASELINE configurations (same JSON contract).

A step = README.md:55-62 of the reference: forward, F.cross_entropy, zero_grad, backward, Adam step.
  value        whole-job images/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e          same step through the public module API from PINNED HOST buffers: H2D copy of X and y and a D2H read of
               the loss inside the timed region
  roofline     the 3x3-convolution tcgen05 kernels (fprop + dgrad + wgrad, ~96 % of the FLOPs): algorithmic FLOPs of those
               launches / their summed CUDA-event durations over a SECOND pass of the same K steps (per-operator brackets on
               the launching stream, single-stream order: with the module's side stream active a bracket would also time
               the kernels it overlaps with; `instrumented_ms_per_step` is that pass's step time), against the measured dense bf16
               peak (MEASURED_PEAKS.json, sustained figure: the kernels run inside a long step); `traffic` = DRAM bytes
               per step of those launches from the committed ncu launch list, next to the algorithmic bytes
  layers       every convolution-family launch of a step: shapes, ms, TFLOP/s
  hbm_kernels  the HBM-bound operators (pool, bilinear, BatchNorm, head + loss, first layer, layout): algorithmic bytes /
               CUDA-event time against the measured copy bandwidth and against 8 TB/s
  cudnn_baseline  the same module tree run through stock PyTorch (cuDNN) on the same GPU: bf16 autocast, channels_last
               weights and activations, cudnn.benchmark — the "existing Blackwell kernels" bar
  cpu_baseline / --impl reference   the reference's CPU path on the host cores on a bounded sample (batch 1 per step): the
               UNMODIFIED reference module when $UNET_REFERENCE_DIR / baseline/_ref / /root/reference holds it
               (kind "reference"), else the oracle port oracle/unet_oracle.py (kind "port").
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-unet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

# BASELINE.json configs (1-based numbering as in BASELINE.md); args = the reference constructor's positional arguments
CONFIGS = {
    2: dict(args=(1, 2, 5, 6, True, True, "upsample"), batch=16, h=256, w=256, dist="randn",
            metric="train images/s UNet d5 wf6 256^2 (same padding, BatchNorm, bilinear upsample)",
            workload="unet_paper_d5_wf6_in1_c2_same_bn_upsample_1x256x256_batch16_per_gpu_fwd+ce+bwd+adam"),
    3: dict(args=(1, 2, 5, 6, False, False, "upconv"), batch=32, h=572, w=572, dist="randn",
            metric="train images/s UNet d5 wf6 572^2 (paper default, valid conv, upconv)",
            workload="unet_paper_d5_wf6_in1_c2_valid_upconv_1x572x572_batch32_per_gpu_fwd+ce+bwd+adam"),
    4: dict(args=(3, 2, 4, 5, True, False, "upconv"), batch=8, h=1024, w=1024, dist="randn",
            metric="train images/s UNet d4 wf5 in3 1024^2 (same padding, upconv)",
            workload="unet_paper_d4_wf5_in3_c2_same_upconv_3x1024x1024_batch8_per_gpu_fwd+ce+bwd+adam"),
    5: dict(args=(3, 6, 5, 2, True, True, "upsample", True), batch=12, h=192, w=640, dist="rand",
            metric="train images/s feature UNet (unet.py Deep decoder, d5 wf2, BatchNorm, upsample, non_neg) 3x192x640",
            workload="unet_py_deep_d5_wf2_in3_c6_same_bn_upsample_nonneg_3x192x640_batch12_per_gpu_fwd+ce+bwd+adam"),
}
HBM_SPEC_GBS = 8000.0  # the figure north_star names


def out_hw(cfg):
    """Output extent of the reference forward (unet.py:73-84): valid convs shrink by 4 per block, pools floor."""
    args = cfg["args"]
    depth, padding = args[2], args[4]
    h, w = cfg["h"], cfg["w"]
    shrink = 0 if padding else 4
    for i in range(depth):
        h, w = h - shrink, w - shrink
        if i != depth - 1:
            h, w = h // 2, w // 2
    for _ in range(depth - 1):
        h, w = 2 * h - shrink, 2 * w - shrink
    return h, w


def make_batch(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    args = cfg["args"]
    fn = torch.randn if cfg["dist"] == "randn" else torch.rand
    x = fn(batch, args[0], cfg["h"], cfg["w"], generator=g)
    ho, wo = out_hw(cfg)
    y = torch.randint(0, args[1], (batch, ho, wo), generator=g)
    return x, y


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d.get("bf16_tflops_sustained", 1400.0)), float(d.get("hbm_gbs", 6650.0)), "measured"
    return 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML (the nvidia-smi data source) every 50 ms in a thread."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop, self._thr, self._err = threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self._err = repr(e)
            return
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}

        def run():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for k, b in bits.items():
                        if r & b:
                            self.reasons.add(k)
                except Exception as e:  # noqa: BLE001
                    self._err = repr(e)
                    return
                time.sleep(0.05)

        self._thr = threading.Thread(target=run, daemon=True)
        self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        sm = sorted(self.samples)
        med = sm[len(sm) // 2] if sm else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm)}
        if self._err:
            out["error"] = self._err
        return out


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def find_reference_dir():
    """Directory holding the UNMODIFIED reference sources (unet.py + unet_original.py), or None.  /root/reference does
    not exist on the GPU box; an operator may mount it and point UNET_REFERENCE_DIR at it."""
    for d in (os.environ.get("UNET_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, "unet.py")) and os.path.isfile(os.path.join(d, "unet_original.py")):
            return d
    return None


def reference_module(cfg, ref_dir):
    """The reference's own nn.Module for this configuration: unet_original.UNet for the 7-argument paper graphs,
    unet.UNet for the 8-argument feature net (unet.py:6 imports torchsnooper without using it: an empty stub)."""
    args = cfg["args"]
    fname, modname = ("unet.py", "_ref_unet") if len(args) == 8 else ("unet_original.py", "_ref_unet_original")
    sys.modules.setdefault("torchsnooper", types.ModuleType("torchsnooper"))
    spec = importlib.util.spec_from_file_location(modname, os.path.join(ref_dir, fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.UNet(*args)


def cpu_reference_step_fn(cfg, batch: int):
    """(step function, kind): one README.md:55-62 training step of the reference's CPU path on `batch` images."""
    x, y = make_batch(cfg, batch, 1234)
    ref_dir = find_reference_dir()
    if ref_dir is not None:
        torch.manual_seed(0)
        model = reference_module(cfg, ref_dir).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)

        def step():
            loss = F.cross_entropy(model(x), y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.detach())

        return step, "reference"
    from oracle import unet_oracle as O
    args = cfg["args"]
    spec = O.UNetSpec(*args[:7], non_neg=(args[7] if len(args) > 7 else False), up_block="deep" if len(args) > 7 else "paper")
    sd = O.init_params(spec, seed=0)
    shapes = O.param_shapes(spec)
    params = {k: (torch.nn.Parameter(v.clone()) if k in shapes else v.clone()) for k, v in sd.items()}
    opt = torch.optim.Adam([v for k, v in params.items() if k in shapes], lr=1e-4)

    def step():
        logits = O.forward(params, x, spec, training=True)   # reference unet_original.py:64-75 / unet.py:73-84
        loss = F.cross_entropy(logits, y)                    # README.md:58
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, "port"


def workload_config(cfg, batch, world, extra=None):
    c = {"workload": cfg["workload"], "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world}",
         "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)"}
    if extra:
        c.update(extra)
    return c


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = cpu_reference_step_fn(cfg, 1)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    ips = 1.0 / dt
    what = "the unmodified reference module" if kind == "reference" else "oracle port of the reference graph"
    sample = (f"1 image of the workload ({cfg['args'][0]}x{cfg['h']}x{cfg['w']}) per step, {what}, torch CPU fp32, "
              f"{cores} threads, {dt * 1e3:.0f} ms/step")
    line = {"impl": "reference", "metric": cfg["metric"], "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(cfg, args.batch or cfg["batch"], max(args.gpus, 1)),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------------ our arm (B200)
def _planes(t):
    """bytes of an activation operand: NHWC bf16 window, two planes in the split precision tier"""
    if t is None:
        return 0
    if hasattr(t, "hi"):
        return 2 * t.hi.numel() * 2
    return t.numel() * t.element_size()


class OpTimer:
    """Brackets operator calls of b200unet.ops with CUDA events on the launching stream and keeps, per call, the
    algorithmic FLOPs (2 * MACs of the mathematical operator) and algorithmic HBM bytes."""

    CONV = ("conv_fwd", "conv_dgrad", "conv_wgrad", "convt_fwd", "convt_dgrad", "convt_wgrad")
    HBM = ("maxpool_fwd", "maxpool_bwd", "bilinear_fwd", "bilinear_bwd", "bn_fwd_train", "bn_fwd_eval", "bn_bwd",
           "head_fwd", "head_bwd", "head_ce_fwd", "head_ce_bwd", "to_nhwc", "im2col3x3")

    def __init__(self, ops):
        self.ops, self.records, self.enabled = ops, [], False
        self._orig = {}
        for name in self.CONV + self.HBM:
            self._orig[name] = getattr(ops, name)
            setattr(ops, name, self._wrap(name))

    # ---- algorithmic work of one call: (flops, bytes, k, shape0, shape1)
    def _work(self, name, a, kw, ret):
        first = lambda v: v[0] if isinstance(v, (list, tuple)) else v  # noqa: E731
        if name == "conv_fwd":
            srcs, w, pad = a[0], a[1], a[3]
            cout, cin, k, _ = w.shape
            n, h, wd, _ = srcs[0].shape
            ho, wo = h + 2 * pad - k + 1, wd + 2 * pad - k + 1
            by = sum(_planes(s) for s in srcs) + _planes(ret)
            return 2.0 * n * ho * wo * cout * cin * k * k, by, k, tuple(srcs[0].shape), tuple(w.shape)
        if name == "conv_dgrad":
            dz, w, dsts = a[0], a[1], a[3]
            masks = a[4] if len(a) > 4 else kw.get("masks", ())
            cout, cin, k, _ = w.shape
            n, h, wd, _ = dz.shape
            by = _planes(dz) + sum(_planes(d) for d in dsts) + sum(_planes(m) for m in masks if m is not None)
            return 2.0 * n * h * wd * cout * cin * k * k, by, k, tuple(dz.shape), tuple(w.shape)
        if name == "conv_wgrad":
            dz, srcs, k = a[0], a[1], a[2]
            n, h, wd, cout = dz.shape
            cin = sum(s.shape[3] for s in srcs)
            by = _planes(dz) + sum(_planes(s) for s in srcs) + cout * cin * k * k * 4
            return 2.0 * n * h * wd * cout * cin * k * k, by, k, tuple(dz.shape), tuple(srcs[0].shape)
        if name == "convt_fwd":
            x, w = a[0], a[1]
            fl = 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * w.shape[0] * w.shape[1] * 4
            return fl, _planes(x) + _planes(ret), 2, tuple(x.shape), tuple(w.shape)
        if name == "convt_dgrad":
            dy, w, dx = a[0], a[1], a[2]
            fl = 2.0 * dy.shape[0] * (dy.shape[1] // 2) * (dy.shape[2] // 2) * w.shape[0] * w.shape[1] * 4
            return fl, _planes(dy) + _planes(dx) + _planes(kw.get("mask")), 2, tuple(dy.shape), tuple(w.shape)
        if name == "convt_wgrad":
            x, dy = a[0], a[1]
            fl = 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] * dy.shape[3] * 4
            return fl, _planes(x) + _planes(dy) + x.shape[3] * dy.shape[3] * 16, 2, tuple(x.shape), tuple(dy.shape)
        # ---- HBM-bound operators: bytes every element is read / written once
        x = first(a[0])
        shp = tuple(x.shape)
        if name == "maxpool_fwd":
            y, idx = ret[0], ret[1]
            return 0.0, _planes(x) + _planes(y) + idx.numel(), 0, shp, ()
        if name == "maxpool_bwd":
            dy, idx, dx = a[0], a[1], a[2]
            add = a[3] if len(a) > 3 else kw.get("add")
            return (0.0, _planes(dy) + idx.numel() + _planes(dx) + _planes(add) + _planes(kw.get("mask")) +
                    _planes(kw.get("pooled")), 0, shp, ())
        if name == "bilinear_fwd":
            return 0.0, _planes(x) + _planes(ret), 0, shp, ()
        if name == "bilinear_bwd":
            return 0.0, _planes(a[0]) + _planes(a[1]) + _planes(kw.get("mask") if len(a) < 3 else a[2]), 0, shp, ()
        if name == "bn_fwd_train":   # statistics pass + apply pass: x twice, y once
            return 0.0, 2 * _planes(x) + _planes(ret[0]), 0, shp, ()
        if name == "bn_fwd_eval":
            return 0.0, _planes(x) + _planes(ret), 0, shp, ()
        if name == "bn_bwd":         # reduction pass (x, dy) + apply pass (x, dy -> dx)
            return 0.0, 2 * (_planes(a[0]) + _planes(a[1])) + _planes(a[1]), 0, shp, ()
        if name in ("head_fwd", "head_ce_fwd"):
            lab = a[4].numel() * 8 if name == "head_ce_fwd" else x.shape[0] * x.shape[1] * x.shape[2] * a[1].shape[0] * 4
            return 0.0, _planes(x) + lab, 0, shp, ()
        if name in ("head_bwd", "head_ce_bwd"):
            lab = x.shape[0] * x.shape[1] * x.shape[2] * 8
            return 0.0, 2 * _planes(x) + lab, 0, shp, ()   # read x (+labels), write dx
        if name == "to_nhwc":
            return 0.0, x.numel() * 4 + _planes(ret), 0, shp, ()
        if name == "im2col3x3":
            return 0.0, _planes(x) + _planes(ret), 0, shp, ()
        return 0.0, 0, 0, shp, ()

    def _wrap(self, name):
        orig = self._orig[name]

        def fn(*a, **kw):
            if not self.enabled:
                return orig(*a, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **kw)
            e1.record()
            fl, by, k, s0, s1 = self._work(name, a, kw, r)
            self.records.append((name, k, fl, by, s0, s1, e0, e1))
            return r

        return fn

    def resolve(self):
        """event pairs -> milliseconds (after a synchronize)"""
        return [(n, k, fl, by, s0, s1, e0.elapsed_time(e1)) for (n, k, fl, by, s0, s1, e0, e1) in self.records]


def stock_forward(model, x):
    """The SAME module tree run by torch's own operators (cuDNN): UNet.forward of the reference, unet.py:73-84 with
    UNetConvBlock.forward :104-106 and UNetUpBlock.forward :160-166 / UNetUpBlockDeep :193-199 — the children of
    b200unet.UNet are real nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d containers, so their own forward is stock
    PyTorch.  Used only for the cudnn_baseline leg."""
    blocks = []
    for i, down in enumerate(model.down_path):
        x = down.block(x)
        if i != len(model.down_path) - 1:
            blocks.append(x)
            x = F.max_pool2d(x, 2)
    for i, up in enumerate(model.up_path):
        u = up.up(x)
        b = blocks[-i - 1]
        dy, dx = (b.shape[2] - u.shape[2]) // 2, (b.shape[3] - u.shape[3]) // 2
        x = up.conv_block.block(torch.cat([u, b[:, :, dy:dy + u.shape[2], dx:dx + u.shape[3]]], 1))
    return model.last(x)


def time_cudnn_baseline(cfg, B, dev, steps=20, warmup=5):
    """Existing-Blackwell-kernel bar: stock PyTorch / cuDNN, bf16 autocast, channels_last weights AND activations,
    cudnn.benchmark, fused Adam; >= 20 timed steps with CUDA events."""
    import b200unet
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = b200unet.UNet(*cfg["args"]).to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x, y = make_batch(cfg, B, 1234)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    y = y.to(dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = stock_forward(model, x)
        loss = F.cross_entropy(logits.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": B / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
           "how": "same module tree through torch's own operators (cuDNN): torch.autocast(bf16), weights and activations "
                  "channels_last, cudnn.benchmark=True, torch.optim.Adam(fused=True)"}
    del model, opt, x, y
    torch.cuda.empty_cache()
    return out


def run_ours(args, cfg):
    import torch.distributed as dist
    import b200unet
    from b200unet import ops
    from b200unet.ddp import DataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = b200unet.load_library()
    if not lib.b200unet_device_ok():
        raise RuntimeError("bench.py needs a B200 (sm_100) device: there is no fallback path")

    B = args.batch or cfg["batch"]
    torch.manual_seed(0)
    model = b200unet.UNet(*cfg["args"], **({"precision": args.precision} if args.precision else {})).to(dev)
    model.train()
    net = DataParallel(model, bucket_bytes=args.bucket_mb << 20, grad_dtype=args.grad_dtype) if world > 1 else model
    if args.optimizer == "fused":  # torch.optim.Adam's arithmetic + packed-weight refresh in one launch (row N1)
        opt = b200unet.FusedAdam(model.parameters(), lr=1e-4, model=model)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=args.graph)

    xh, yh = make_batch(cfg, B, 1234 + rank)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    xd, yd = xh.to(dev), yh.to(dev)

    graphed = None
    if args.graph:
        # the whole step as ONE CUDA graph (b200unet.GraphedTrainStep): for the small / narrow configurations the eager
        # step is bound by Python launch overhead.  With N > 1 the bucketed all-reduces are captured too.
        graphed = b200unet.GraphedTrainStep(net, opt, xd, yd)

    def step_device():
        if graphed is not None:
            return graphed(xd, yd)
        loss = net.loss(xd, yd)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    # end-to-end: every step copies ITS inputs from pinned host memory and reads its loss back.  The copy of step i+1
    # is issued on a side stream while step i computes (double-buffered device inputs), as a training loop would do.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(xd), torch.empty_like(yd)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            bufs[slot][0].copy_(xh, non_blocking=True)
            bufs[slot][1].copy_(yh, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        i = state["i"]
        slot = i % 2
        if not state["primed"]:
            consumed[0].record()
            consumed[1].record()
            prefetch(slot)
            state["primed"] = True
        prefetch(1 - slot)  # next step's inputs travel while this step computes
        torch.cuda.current_stream().wait_event(ready[slot])
        if graphed is not None:
            loss = graphed(bufs[slot][0], bufs[slot][1])
        else:
            loss = net.loss(bufs[slot][0], bufs[slot][1])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        consumed[slot].record()
        state["i"] = i + 1
        return loss.item()  # D2H read of the loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        own = e0.elapsed_time(e1) / steps
        ms, per_rank = own, [own]
        if world > 1:
            t = torch.tensor([own], device=dev)
            allr = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            per_rank = [float(v.item()) for v in allr]
            ms = max(per_rank)
        return ms, per_rank

    timer = OpTimer(ops)
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.b200unet_launch_count()
    if world > 1:
        net.bucketer.profile = True
    # ---- the timed region: K steps exactly as a user runs them (no per-operator events; the module's side stream overlaps
    # HBM-bound backward kernels with tensor-bound ones)
    ms, per_rank = timed(step_device, args.steps)
    launches = int(lib.b200unet_launch_count() - n0)
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    ddp_stats = None
    if world > 1:
        ddp_stats = net.bucketer.profile_summary()
        net.bucketer.profile = False
    # ---- the same K steps once more with every operator bracketed by CUDA events on its launching stream, in
    # SINGLE-STREAM order: with the side stream active a bracket would also measure the kernels it overlaps with (the
    # per-kernel sums then exceed the step), so per-kernel times / roofline / hbm_kernels come from this pass and its own
    # step time is reported next to them (`instrumented_ms_per_step`).  A graph replay runs no Python: nothing to bracket.
    ms_instr = None
    if graphed is None:
        side_was = model.side_stream_wgrad
        model.side_stream_wgrad = False
        step_device()
        timer.enabled = True
        ms_instr, _ = timed(step_device, args.steps)
        timer.enabled = False
        model.side_stream_wgrad = side_was
        step_device()
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    loss_val = step_e2e()

    dp_check = None
    if world > 1:
        # data-parallel equivalence, driver-visible: after the same number of averaged-gradient updates every replica
        # must hold bit-identical weights (they started identical and applied identical all-reduced gradients)
        cs = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        ab = torch.stack([p.detach().double().abs().sum() for p in model.parameters()]).sum().reshape(1)
        both = torch.cat([cs, ab])
        allc = [torch.empty_like(both) for _ in range(world)]
        dist.all_gather(allc, both)
        same = all(torch.equal(allc[0], c) for c in allc)
        dp_check = {"weights_bit_identical_across_ranks": bool(same), "weight_checksum": float(cs.item())}
        if not same:
            raise RuntimeError(f"data-parallel replicas diverged: {[c.tolist() for c in allc]}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B / (ms * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    peak_tf, peak_gbs, which = measured_peaks()
    recs = timer.resolve()
    steps = args.steps

    # ---- roofline of the dominant kernels: every 3x3 convolution launch (fprop + dgrad + wgrad)
    k3 = [r for r in recs if r[0] in OpTimer.CONV and r[1] == 3 and not (r[0] == "conv_fwd" and r[5][1] < 8)]
    fl3 = sum(r[2] for r in k3) / steps
    ms3 = sum(r[6] for r in k3) / steps
    by3 = sum(r[3] for r in k3) / steps
    achieved = fl3 / (ms3 * 1e-3) / 1e12 if ms3 > 0 else 0.0
    # DRAM traffic, like for like: every tcgen05 launch of a step (3x3 / 1x1 conv + ConvTranspose fprop, dgrad, wgrad and the
    # first layer's wgrad = everything in OpTimer.CONV except the first layer's CUDA-core / mma.sync forward) — measured
    # bytes from the committed ncu launch list against the algorithmic bytes of exactly those launches
    tc = [r for r in recs if r[0] in OpTimer.CONV and not (r[0] == "conv_fwd" and r[5][1] < 8)]
    by_tc = sum(r[3] for r in tc) / steps
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", f"r02_traffic_cfg{args.config}.json")
    if os.path.isfile(tpath) and B == cfg["batch"]:
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic, traffic_src = tj.get("dram_bytes_per_step_tcgen05"), tj.get("source")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "frac_of_burst_1626.6": achieved / 1626.6, "frac_of_nominal_2250": achieved / 2250.0,
                "traffic": traffic,
                "traffic_unit": "DRAM bytes per STEP over all tcgen05 launches (ncu dram__bytes_read + write); compare with "
                                "algorithmic_bytes_per_step_tcgen05",
                "traffic_source": traffic_src, "algorithmic_bytes_per_step_tcgen05": by_tc,
                "traffic_over_algorithmic": (traffic / by_tc) if (traffic and by_tc) else None,
                "algorithmic_bytes_per_step": by3,
                "kernel": "umma_conv_kernel + wgrad_umma_kernel (all 3x3 conv fprop/dgrad/wgrad launches of a step)",
                "peak_source": f"{which} bf16_tflops_sustained", "flops_per_step": fl3, "ms_per_step_in_kernel": ms3,
                "launches_per_step": len(k3) / steps, "share_of_step": ms3 / ms if ms > 0 else None,
                "measured_in": "second pass of the same K steps, operators bracketed by CUDA events, single-stream order",
                "instrumented_ms_per_step": ms_instr}

    # ---- per-layer table of the convolution family, and the per-op-kind breakdown
    agg = {}
    for n, k, fl, by, s0, s1, t in recs:
        if n in OpTimer.CONV:
            d = agg.setdefault((n, s0, s1), [0.0, 0.0, 0.0, 0])
            d[0] += t
            d[1] += fl
            d[2] += by
            d[3] += 1
    layers = [{"op": n, "a": list(s0), "b": list(s1), "ms": round(t / c, 4), "tflops": round(fl / t / 1e9, 1),
               "gb": round(by / c / 1e9, 4)} for (n, s0, s1), (t, fl, by, c) in agg.items()]
    breakdown = {}
    for n, k, fl, by, s0, s1, t in recs:
        if n in OpTimer.CONV:
            d = breakdown.setdefault(f"{n}_k{k}", {"ms": 0.0, "flops": 0.0, "calls": 0})
            d["ms"] += t
            d["flops"] += fl
            d["calls"] += 1
    breakdown = {k: {"ms_per_step": v["ms"] / steps, "tflops": v["flops"] / max(v["ms"], 1e-9) / 1e9,
                     "calls_per_step": v["calls"] / steps} for k, v in sorted(breakdown.items())}

    # ---- HBM-bound operators: achieved GB/s against the measured copy bandwidth and against 8 TB/s
    hbm = {}
    for n, k, fl, by, s0, s1, t in recs:
        key = n
        if n == "conv_fwd" and s1 and s1[1] < 8:
            key = "first_layer_fwd"   # Cin <= 4: CUDA-core kernel, bound by writing the 64-channel tensor
        elif n in OpTimer.CONV:
            continue
        d = hbm.setdefault(key, {"ms": 0.0, "bytes": 0.0, "calls": 0})
        d["ms"] += t
        d["bytes"] += by
        d["calls"] += 1
    hbm_kernels = {k: {"ms_per_step": v["ms"] / steps, "gb_per_step": v["bytes"] / steps / 1e9,
                       "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6,
                       "frac_of_measured": v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak_gbs,
                       "frac_of_8tbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6 / HBM_SPEC_GBS,
                       "calls_per_step": v["calls"] / steps} for k, v in sorted(hbm.items())}
    conv_ms = sum(r[6] for r in recs if r[0] in OpTimer.CONV) / steps
    step_split = {"conv3x3_ms": ms3, "conv_family_ms": conv_ms,
                  "hbm_ops_ms": sum(v["ms_per_step"] for v in hbm_kernels.values()),
                  "other_ms": (ms_instr if ms_instr is not None else ms) - conv_ms -
                              sum(v["ms_per_step"] for k, v in hbm_kernels.items() if k != "first_layer_fwd"),
                  "note": "split of the instrumented pass (roofline.instrumented_ms_per_step); other = optimizer, "
                          "weight-gradient reductions outside the bracketed calls, launch gaps"}

    cpu = None
    if world == 1 and not args.no_cpu:
        # bounded sample: ~10-15 s of CPU work (1 warm-up step sizes the number of timed steps)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cstep, kind = cpu_reference_step_fn(cfg, 1)
        t0 = time.perf_counter()
        cstep()
        t_warm = time.perf_counter() - t0
        n_timed = max(2, min(20, int(12.0 / max(t_warm, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(n_timed):
            cstep()
        cms = (time.perf_counter() - t0) / n_timed * 1e3
        what = "the unmodified reference module" if kind == "reference" else "oracle port of the reference graph"
        cpu = {"value": 1e3 / cms, "unit": "images/s", "cores": cores, "kind": kind,
               "sample": f"1 image of the workload ({cfg['args'][0]}x{cfg['h']}x{cfg['w']}) per step, 1 warm-up + {n_timed} "
                         f"timed steps, {what} (torch CPU fp32, {cores} threads), {cms:.0f} ms/step"}
    cudnn = None
    if world == 1 and not args.no_cudnn:
        del graphed
        try:
            cudnn = time_cudnn_baseline(cfg, B, dev)
            cudnn["ours_over_cudnn"] = value / cudnn["value"]
        except Exception as e:  # noqa: BLE001  (e.g. out of memory at a user-chosen batch): report, do not fail the bench
            cudnn = {"error": repr(e)[:300]}

    line = {"metric": cfg["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(cfg, B, world),
            "run": {"optimizer": ("b200unet.FusedAdam (Adam + packed-weight refresh, one launch)"
                                  if args.optimizer == "fused" else "torch.optim.Adam(fused=True)"),
                    "precision_tier": model.precision, "cuda_graph": bool(args.graph), "loss": float(loss_val)},
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": xh.numel() * 4 + yh.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "cudnn_baseline": cudnn, "conv_breakdown": breakdown, "hbm_kernels": hbm_kernels, "step_split": step_split,
            "per_rank_ms": per_rank, "ddp": ddp_stats, "dp_check": dp_check, "layers": layers}
    _emit(line)
    if args.detail:
        with open(args.detail, "w") as fh:
            for L in layers:
                fh.write(f"{L['op']:12s} {str(tuple(L['a'])):26s} {str(tuple(L['b'])):26s} ms/call {L['ms']:8.3f} "
                         f"TFLOP/s {L['tflops']:8.1f} GB {L['gb']:7.3f}\n")
            for k, v in hbm_kernels.items():
                fh.write(f"{k:16s} ms/step {v['ms_per_step']:7.3f} GB/step {v['gb_per_step']:7.3f} GB/s {v['gbs']:8.1f} "
                         f"({v['frac_of_measured']:.2f} of measured {peak_gbs:.0f})\n")
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the first
    communicator), so everything else that goes to fd 1 is re-routed to stderr and the JSON line is written to the
    original stdout at the end."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based)")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the configuration's)")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    ap.add_argument("--precision", default=None, choices=["bf16", "split"], help="override the module's precision tier")
    ap.add_argument("--graph", action="store_true", help="capture the whole training step in one CUDA graph")
    ap.add_argument("--bucket-mb", type=int, default=32, help="gradient all-reduce bucket size (N > 1)")
    ap.add_argument("--grad-dtype", default="fp32", choices=["fp32", "bf16"], help="wire precision of the gradient all-reduce")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cudnn", action="store_true", help="skip the cudnn_baseline leg")
    ap.add_argument("--detail", default=None, help="also write the per-layer table to this file")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
