"""Training-step throughput of the other BASELINE.json configurations (2, 4, 5) on one B200 — information for DESIGN.md,
not the headline bench.   python experiments/bench_configs.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
sys.path.insert(0, ROOT)
import torch
import b200unet
from oracle import unet_oracle as O

CFG = {
    "config2 same+BN+upsample 1x256x256 b16": ((1, 2, 5, 6, True, True, "upsample"), "paper", 16, 256, 256),
    "config3 paper valid 1x572x572 b32": ((1, 2, 5, 6, False, False, "upconv"), "paper", 32, 572, 572),
    "config4 d4 wf5 in3 same 3x1024x1024 b8": ((3, 2, 4, 5, True, False, "upconv"), "paper", 8, 1024, 1024),
    "config5 deep feature variant 3x192x640 b12": ((3, 6, 5, 2, True, True, "upsample", True), "deep", 12, 192, 640),
}
RUNS = []
for name, cfg in CFG.items():
    RUNS.append((name, cfg, "auto"))
    if cfg[0][5]:  # BatchNorm graphs: auto = split tier; also time the plain bf16 tier
        RUNS.append((name + " [precision=bf16]", cfg, "bf16"))
for name, (args, ub, b, h, w), tier in RUNS:
    torch.manual_seed(0)
    m = b200unet.UNet(*args, up_block=ub, precision=tier).cuda().train()
    opt = b200unet.FusedAdam(m.parameters(), lr=1e-4, model=m)
    spec = O.UNetSpec(*args[:7], non_neg=(args[7] if len(args) > 7 else False), up_block=ub)
    ho, wo = O.output_hw(spec, h, w)
    x = torch.randn(b, args[0], h, w, device="cuda")
    y = torch.randint(0, args[1], (b, ho, wo), device="cuda")
    def step():
        loss = m.loss(x, y); opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:64s} {m.precision:5s} {ms:8.2f} ms/step  {b / ms * 1e3:9.1f} img/s  (host wall {1e2 * (time.perf_counter() - t0):.2f} ms/step)", flush=True)
