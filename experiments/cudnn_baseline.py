"""The 'existing Blackwell kernel' bar: the reference graph through stock PyTorch/cuDNN on the same B200
(oracle restatement of unet_original.py moved to CUDA), forward + F.cross_entropy + backward + fused Adam, batch 32, 572^2.
Three modes: strict fp32, TF32 (PyTorch default for convs), bf16 autocast + channels_last."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from oracle import unet_oracle as O

spec = O.UNetSpec(1, 2, 5, 6, False, False, "upconv")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.randn(B, 1, 572, 572, device="cuda")
y = torch.randint(0, 2, (B, 388, 388), device="cuda")
torch.backends.cudnn.benchmark = True
for mode in ("fp32", "tf32", "bf16_autocast_channels_last"):
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    sd = {k: torch.nn.Parameter(v.cuda()) for k, v in O.init_params(spec, seed=0).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4, fused=True)
    xin = x.contiguous(memory_format=torch.channels_last) if "channels_last" in mode else x

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled="bf16" in mode):
            logits = O.forward(sd, xin, spec)
            loss = F.cross_entropy(logits.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"stock PyTorch/cuDNN {mode:30s} batch {B}: {ms:8.2f} ms/step  {B / ms * 1e3:8.1f} img/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{mode}: failed: {type(e).__name__}: {str(e)[:120]}", flush=True)
    del sd, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
