"""bench.py's algorithmic-work accounting (the numerator of `roofline.achieved`) on CPU, with meta tensors: the 3x3-conv
FLOPs of one training step of BASELINE config 3 must equal SURVEY.md 8(d) / BASELINE.md section 2 — 868.3 GFLOP per image
(fprop + dgrad + wgrad, no dgrad for the first convolution) x 32 images — and the output extent helper must reproduce the
reference's 572 -> 388."""
import importlib.util
import os
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_output_extent_and_configs_match_baseline_json():
    import json
    b = _bench()
    assert b.out_hw(b.CONFIGS[3]) == (388, 388)          # unet_original.py: valid convolutions, README 572 -> 388
    assert b.out_hw(b.CONFIGS[2]) == (256, 256) and b.out_hw(b.CONFIGS[4]) == (1024, 1024)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5 and set(b.CONFIGS) == {2, 3, 4, 5}


def test_conv3x3_flops_of_config3_match_the_survey():
    b = _bench()
    names = b.OpTimer.CONV + b.OpTimer.HBM
    timer = b.OpTimer(types.SimpleNamespace(**{n: (lambda *a, **k: None) for n in names}))
    B, meta = 32, "meta"
    t = lambda *s: torch.empty(s, device=meta, dtype=torch.bfloat16)   # noqa: E731
    w = lambda co, ci: torch.empty((co, ci, 3, 3), device=meta)         # noqa: E731
    total = 0.0
    # encoder: (cin, cout, input extent) of the two valid 3x3 convolutions per level, then the decoder's
    h, cin = 572, 1
    enc = []
    for lvl in range(5):
        cout = 64 << lvl
        enc.append((cin, cout, h))
        enc.append((cout, cout, h - 2))
        h, cin = (h - 4) // 2 if lvl < 4 else h - 4, cout
    dec = []
    for lvl in reversed(range(4)):
        cout = 64 << lvl
        h = 2 * h
        dec.append((2 * cout, cout, h))
        dec.append((cout, cout, h - 2))
        h -= 4
    assert h == 388
    for i, (ci, co, hin) in enumerate(enc + dec):
        fl, _, k, _, _ = timer._work("conv_fwd", ([t(B, hin, hin, ci)], w(co, ci), None, 0, True), {}, t(B, hin - 2, hin - 2, co))
        assert k == 3
        flw = timer._work("conv_wgrad", (t(B, hin - 2, hin - 2, co), [t(B, hin, hin, ci)], 3, 0), {}, None)[0]
        fld = 0.0 if i == 0 else timer._work("conv_dgrad", (t(B, hin - 2, hin - 2, co), w(co, ci), 0, [t(B, hin, hin, ci)]), {}, None)[0]
        assert fl == flw and (i == 0 or fl == fld)
        total += fl + flw + fld
    assert abs(total / B / 1e9 - 868.3) < 0.5, total / B / 1e9
