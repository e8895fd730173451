"""Forward-only (eval, torch.no_grad) throughput — the inference half of next-row N2 (reference: eval loop run.py:248-291,
full-size eval shape 480x640, options.py:105-107).   python experiments/bench_infer.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch
import b200unet

CFG = {
    "paper d5 wf6 valid upconv, 1x572x572, batch 32": ((1, 2, 5, 6, False, False, "upconv"), "paper", 32, 572, 572),
    "repo feature net (deep, BN, wf=2), 3x480x640, batch 8": ((3, 6, 5, 2, True, True, "upsample", True), "deep", 8, 480, 640),
    "repo feature net (deep, BN, wf=2), 3x192x640, batch 12": ((3, 6, 5, 2, True, True, "upsample", True), "deep", 12, 192, 640),
}
for name, (args, ub, b, h, w) in CFG.items():
    for tier in (("auto", "bf16") if args[5] else ("auto",)):
        torch.manual_seed(0)
        m = b200unet.UNet(*args, up_block=ub, precision=tier).cuda().eval()
        x = torch.randn(b, args[0], h, w, device="cuda")
        with torch.no_grad():
            for _ in range(3):
                y = m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y = m(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:58s} {m.precision:5s} {ms:8.2f} ms/forward {b / ms * 1e3:9.1f} img/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        del m, x, y
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
