"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (this container only).

    python oracle/make_golden.py

Each fixture holds: the constructor spec, the full state_dict, a seeded input, labels, and the reference's
fp32 CPU outputs of one README training step (README.md:57-62): logits, loss, every parameter gradient and the
updated BatchNorm running statistics.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader  # noqa: E402
from oracle.unet_oracle import UNetSpec  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (spec, input N,H,W, label kind)
CASES = {
    # the paper graph (BASELINE configs 1/3) at the smallest valid-conv size for depth 5, narrow
    "paper_valid_d5_wf2": (UNetSpec(1, 2, 5, 2, False, False, "upconv"), (1, 188, 188)),
    # 64-channel-multiple widths so that the tcgen05 kernels (not the small-channel ones) are exercised
    "paper_valid_d2_wf6": (UNetSpec(1, 2, 2, 6, False, False, "upconv"), (2, 44, 52)),
    # BASELINE config 2 graph: same padding + BN + bilinear upsample
    "paper_same_bn_upsample_d3_wf3": (UNetSpec(1, 2, 3, 3, True, True, "upsample"), (2, 32, 40)),
    # BASELINE config 4 graph: in=3, depth 4, same padding, upconv, odd sizes (crop is not a no-op)
    "paper_same_d4_wf3_in3_odd": (UNetSpec(3, 2, 4, 3, True, False, "upconv"), (1, 50, 44)),
    # BASELINE config 5 graph: unet.py's Deep decoder, options.py:19-25 values
    "deep_cfg5_d5_wf2": (UNetSpec(3, 6, 5, 2, True, True, "upsample", True, "deep"), (2, 32, 48)),
    # Deep decoder with transposed convolutions and valid padding
    "deep_valid_upconv_d3_wf3": (UNetSpec(1, 3, 3, 3, False, False, "upconv", False, "deep"), (1, 60, 68)),
}


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for name, (spec, (n, h, w)) in CASES.items():
        torch.manual_seed(1234)
        model = reference_loader.build_reference_module(spec)
        model.train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        x = torch.randn(n, spec.in_channels, h, w)
        logits = model(x)
        # structured labels: quantile buckets of the centre-cropped first input channel
        ho, wo = logits.shape[2:]
        dy, dx = (h - ho) // 2, (w - wo) // 2
        c = x[:, 0, dy:dy + ho, dx:dx + wo]
        qs = torch.quantile(c.flatten(), torch.linspace(0, 1, spec.n_classes + 1)[1:-1])
        y = torch.bucketize(c, qs)
        loss = F.cross_entropy(logits, y)
        model.zero_grad()
        loss.backward()
        arrays = {"x": x.numpy(), "y": y.numpy().astype(np.int64), "logits": logits.detach().numpy(),
                  "loss": np.float32(loss.item())}
        for k, v in sd0.items():
            arrays["sd/" + k] = v.numpy()
        for k, p in model.named_parameters():
            arrays["grad/" + k] = p.grad.numpy()
        for k, v in model.state_dict().items():
            if "running_" in k or "num_batches" in k:
                arrays["sd_after/" + k] = v.numpy()
        arrays["spec"] = np.frombuffer(json.dumps(spec.__dict__).encode(), dtype=np.uint8)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: logits {tuple(logits.shape)} loss {loss.item():.6f} -> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
