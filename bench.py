#!/usr/bin/env python
"""Headline benchmark: training images/s of the paper-default U-Net (in=1, n_classes=2, depth 5, wf 6, valid padding,
upconv) on 1x572x572 synthetic images, batch 32 per GPU (BASELINE.json configs[2]; it fits one B200, so it is the
N=1 workload too), data-parallel over N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A step = README.md:55-62 of the reference: forward, F.cross_entropy, zero_grad, backward, Adam step.
  value  : whole-job images/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e    : same step through the public module API from PINNED HOST buffers: H2D copy of X and y and a D2H read of
           the loss inside the timed region
  roofline: the 3x3-convolution tcgen05 kernels (fprop + dgrad + wgrad, 96 % of the FLOPs): algorithmic FLOPs of
           those launches / their summed CUDA-event durations inside the timed region, against the measured dense
           bf16 peak (MEASURED_PEAKS.json, sustained figure because the kernels run inside a long step)
  cpu_baseline / --impl reference: the reference's CPU path (oracle/unet_oracle.py, a restatement of unet_original.py
           on torch.nn.functional — the reference itself is pure PyTorch and does not exist on the GPU box) timed on
           the host cores on a bounded sample (batch 1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-unet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

H_IN = 572
METRIC = "train images/s UNet d5 wf6 572^2 (paper default, valid conv, upconv)"
WORKLOAD = "unet_paper_d5_wf6_in1_c2_valid_upconv_1x572x572_batch32_per_gpu_fwd+ce+bwd+adam"


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
def cpu_reference_step_fn(batch: int):
    from oracle import unet_oracle as O
    spec = O.UNetSpec(1, 2, 5, 6, False, False, "upconv")
    sd = O.init_params(spec, seed=0)
    shapes = O.param_shapes(spec)
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items() if k in shapes}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 1, H_IN, H_IN, generator=g)
    ho, wo = O.output_hw(spec, H_IN, H_IN)
    y = torch.randint(0, 2, (batch, ho, wo), generator=g)

    def step():
        logits = O.forward(params, x, spec, training=True)   # reference unet_original.py:64-75
        loss = F.cross_entropy(logits, y)                    # README.md:58
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_cpu(steps: int, warmup: int, batch: int = 1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_reference_step_fn(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, ms, cores = time_cpu(args.steps, args.warmup, batch=1)
    sample = f"batch 1 of the 1x{H_IN}x{H_IN} workload per step, oracle port of unet_original.py on torch CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "batch 1 per step (CPU)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------------ our arm (B200)
class ConvTimer:
    """Brackets every convolution-family operator call with CUDA events on the launching stream and keeps the
    algorithmic FLOPs (2 * MACs of the mathematical operator) of each call."""

    def __init__(self, ops):
        self.ops, self.records, self.enabled, self.detail = ops, [], False, None
        self._orig = {}
        for name in ("conv_fwd", "conv_dgrad", "conv_wgrad", "convt_fwd", "convt_dgrad", "convt_wgrad"):
            self._orig[name] = getattr(ops, name)
            setattr(ops, name, self._wrap(name))

    def _flops(self, name, a, kw):
        if name == "conv_fwd":
            srcs, w, pad = a[0], a[1], a[3]
            cout, cin, k, _ = w.shape
            n, h, wd, _ = srcs[0].shape
            return 2.0 * n * (h + 2 * pad - k + 1) * (wd + 2 * pad - k + 1) * cout * cin * k * k, k
        if name == "conv_dgrad":
            dz, w = a[0], a[1]
            cout, cin, k, _ = w.shape
            n, h, wd, _ = dz.shape
            return 2.0 * n * h * wd * cout * cin * k * k, k
        if name == "conv_wgrad":
            dz, srcs, k = a[0], a[1], a[2]
            n, h, wd, cout = dz.shape
            cin = sum(s.shape[3] for s in srcs)
            return 2.0 * n * h * wd * cout * cin * k * k, k
        x = a[0]
        if name == "convt_fwd":
            w = a[1]
            return 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * w.shape[0] * w.shape[1] * 4, 2
        if name == "convt_dgrad":
            dy, w = a[0], a[1]
            return 2.0 * dy.shape[0] * (dy.shape[1] // 2) * (dy.shape[2] // 2) * w.shape[0] * w.shape[1] * 4, 2
        dy = a[1]
        return 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] * dy.shape[3] * 4, 2

    def _wrap(self, name):
        orig = self._orig[name]

        def fn(*a, **kw):
            if not self.enabled:
                return orig(*a, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **kw)
            e1.record()
            fl, k = self._flops(name, a, kw)
            self.records.append((name, k, fl, e0, e1))
            if self.detail is not None:
                t0 = a[0][0] if isinstance(a[0], (list, tuple)) else a[0]
                t1 = a[1][0] if isinstance(a[1], (list, tuple)) else a[1]
                self.detail.append((name, tuple(t0.shape), tuple(t1.shape), fl, e0, e1))
            return r

        return fn

    def detail_table(self):
        agg = {}
        for name, s0, s1, fl, e0, e1 in self.detail or []:
            d = agg.setdefault((name, s0, s1), [0.0, 0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += fl
            d[2] += 1
        rows = [f"{n:12s} {str(s0):24s} {str(s1):24s} calls {c:3d} ms/call {ms / c:8.3f} TFLOP/s {fl / ms / 1e9:8.1f}"
                for (n, s0, s1), (ms, fl, c) in agg.items()]
        return "\n".join(rows)

    def summary(self):
        out = {}
        for name, k, fl, e0, e1 in self.records:
            key = f"{name}_k{k}"
            d = out.setdefault(key, {"ms": 0.0, "flops": 0.0, "calls": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["calls"] += 1
        return out


def run_ours(args):
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

    B = args.batch
    torch.manual_seed(0)
    model = b200unet.UNet(1, 2, 5, 6, False, False, "upconv").to(dev)
    model.train()
    net = DataParallel(model) if world > 1 else model
    if args.optimizer == "fused":  # torch.optim.Adam's arithmetic + packed-weight refresh in one launch (row N1)
        opt = b200unet.FusedAdam(model.parameters(), lr=1e-4, model=model)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

    g = torch.Generator().manual_seed(1234 + rank)
    xh = torch.randn(B, 1, H_IN, H_IN, generator=g).pin_memory()
    yh = torch.randint(0, 2, (B, 388, 388), generator=g).pin_memory()
    xd, yd = xh.to(dev), yh.to(dev)

    def step_device():
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
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    timer = ConvTimer(ops)
    if args.detail and rank == 0:
        timer.detail = []
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.b200unet_launch_count()
    timer.enabled = True
    ms = timed(step_device, args.steps)
    timer.enabled = False
    launches = int(lib.b200unet_launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    loss_val = step_e2e()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B / (ms * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    peak_tf, peak_gbs, which = measured_peaks()
    summ = timer.summary()
    k3 = {k: v for k, v in summ.items() if k.endswith("_k3")}
    fl3 = sum(v["flops"] for v in k3.values()) / args.steps
    ms3 = sum(v["ms"] for v in k3.values()) / args.steps
    achieved = fl3 / (ms3 * 1e-3) / 1e12 if ms3 > 0 else 0.0
    calls3 = sum(v["calls"] for v in k3.values()) / args.steps
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(tpath):  # ncu dram__bytes_read+write per tcgen05 launch of the same bench command (batch 32)
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_note": "mean DRAM bytes per tcgen05 launch (ncu, profiles/r01_traffic.json)",
                "kernel": "umma_conv_kernel + wgrad_umma_kernel (all 3x3 conv fprop/dgrad/wgrad launches of a step)",
                "peak_source": f"{which} bf16_tflops_sustained", "flops_per_step": fl3, "ms_per_step_in_kernel": ms3,
                "launches_per_step": calls3,
                "share_of_step": ms3 / ms if ms > 0 else None}
    breakdown = {k: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / max(v["ms"], 1e-9) / 1e9,
                     "calls_per_step": v["calls"] / args.steps} for k, v in sorted(summ.items())}
    cpu = None
    if world == 1 and not args.no_cpu:
        # bounded sample: ~10-15 s of CPU work (1 warm-up step sizes the number of timed steps)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cstep = cpu_reference_step_fn(1)
        t0 = time.perf_counter()
        cstep()
        t_warm = time.perf_counter() - t0
        n_timed = max(2, min(20, int(12.0 / max(t_warm, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(n_timed):
            cstep()
        cms = (time.perf_counter() - t0) / n_timed * 1e3
        cpu = {"value": 1e3 / cms, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"batch 1 of the workload (1x{H_IN}x{H_IN}) per step, 1 warm-up + {n_timed} timed steps, oracle port "
                         f"of unet_original.py (torch CPU fp32, {cores} threads), {cms:.0f} ms/step"}
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)",
                       "optimizer": ("b200unet.FusedAdam (Adam + packed-weight refresh, one launch)"
                                     if args.optimizer == "fused" else "torch.optim.Adam(fused=True)"),
                       "loss": float(loss_val)},
            "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": xh.numel() * 4 + yh.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "conv_breakdown": breakdown}
    _emit(line)
    if args.detail:
        with open(args.detail, "w") as fh:
            fh.write(timer.detail_table() + "\n")
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
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--detail", default=None, help="write a per-layer table of the convolution launches to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
