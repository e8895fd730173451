"""precision_emu.py on the small BatchNorm cases tests/test_unet_gpu.py uses (CPU, seconds)."""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "experiments")
import precision_emu as E  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402

CASES = {
    "paper_same_bn_upsample_wf4": (O.UNetSpec(1, 2, 3, 4, True, True, "upsample"), (4, 32, 32)),
    "deep_cfg5": (O.UNetSpec(3, 6, 4, 2, True, True, "upsample", True, "deep"), (3, 32, 40)),
}
for name, (spec, (n, h, w)) in CASES.items():
    torch.manual_seed(7)
    sd = O.init_params(spec, seed=3)
    x = torch.randn(n, spec.in_channels, h, w)
    ho, wo = O.output_hw(spec, h, w)
    c = x[:, 0, :ho, :wo]
    qs = torch.quantile(c.flatten(), torch.linspace(0, 1, spec.n_classes + 1)[1:-1])
    y = torch.bucketize(c, qs)
    ref_logits, ref_loss, ref_g, _ = O.loss_and_grads(sd, x, y, spec)
    names = list(ref_g)
    flat = lambda g: torch.cat([g[k].flatten() for k in names])
    for mode in ("bf16", "split"):
        E.MODE, E.BW_W, E.BW_DZ = mode, "bf", "bf"
        realF, realq = O.F, O._q
        O.F, O._q = E.Shim(), E.Store.apply
        try:
            logits, loss, g, _ = O.loss_and_grads(sd, x, y, spec)
        finally:
            O.F, O._q = realF, realq
        agree = (logits.argmax(1) == ref_logits.argmax(1)).float().mean().item()
        print(f"{name} {mode:6s} logits {O.rel_l2(logits, ref_logits):.2e} argmax {100 * agree:.3f}% "
              f"grad-all {O.rel_l2(flat(g), flat(ref_g)):.2e}")
