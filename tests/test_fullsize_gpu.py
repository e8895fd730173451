"""BASELINE.json configurations at their FULL sizes on the B200 (through the drop-in module -> C ABI).

config 1 (paper default, batch 1, 572x572) is small enough for the CPU oracle: logits / loss / gradients are compared
against it with north_star's tolerances; so is config 4 at its 1024x1024 resolution with one image, and so are the two BatchNorm configurations (2 and 5, full batch: the batch
statistics are part of the result), which run in the module's split precision tier.  For the larger configurations the
oracle would take minutes, so they are checked through size-independent properties of the training step:
  * run-to-run bit-exactness (every kernel reduces in a fixed order),
  * linearity of the backward pass in the upstream gradient,
  * batch additivity: gradients on [a; b] equal the pixel-weighted mean of the gradients on a and on b (no BatchNorm) —
    this is exactly the identity data parallelism relies on, and it crosses different tile plans (n_img changes),
  * fused loss == F.cross_entropy(logits) and a finite, sane loss at random init.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import rel_l2
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def make(spec_args, up_block="paper"):
    import b200unet
    torch.manual_seed(0)
    return b200unet.UNet(*spec_args, up_block=up_block).cuda().train()


def grads_of(model, x, y, scale=1.0):
    model.zero_grad(set_to_none=True)
    loss = model.loss(x, y)
    (loss * scale).backward()
    return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def cat(g):
    return torch.cat([v.flatten().float() for v in g.values()])


def test_config1_paper_default_batch1_572_vs_oracle():
    spec = O.UNetSpec(1, 2, 5, 6, False, False, "upconv")
    sd = O.init_params(spec, seed=0)
    torch.manual_seed(1)
    x = torch.randn(1, 1, 572, 572)
    c = x[:, 0, 92:92 + 388, 92:92 + 388]
    y = (c > 0).long()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd, x, y, spec)
    import b200unet
    model = b200unet.UNet(1, 2, 5, 6, False, False, "upconv").cuda().train()
    model.load_state_dict(sd)
    logits = model(x.cuda())
    loss = F.cross_entropy(logits, y.cuda())
    loss.backward()
    e = rel_l2(logits.detach().cpu(), ref_logits)
    agree = float((logits.argmax(1).cpu() == ref_logits.argmax(1)).float().mean())
    keys = list(ref_grads)
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_grads[k].flatten() for k in keys]))
    print(f"[config1 572^2] logits rel-L2 {e:.3e} argmax agreement {agree:.5f} grad rel-L2 {eg:.3e} "
          f"loss {float(loss):.6f} vs {float(ref_loss):.6f}")
    assert e <= 1e-2 and eg <= 2e-2 and agree >= 0.999   # BASELINE.json north_star tolerances
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * max(1.0, abs(float(ref_loss)))


def test_config4_full_resolution_batch1_vs_oracle():
    """BASELINE config 4 (in=3, depth 4, wf 5, same padding) at its full 1024x1024 resolution, one image (the oracle
    needs ~20 s of CPU for it): 32-channel layers, zero padding at every level, the widest tensors of all configs."""
    import b200unet
    spec = O.UNetSpec(3, 2, 4, 5, True, False, "upconv")
    sd = O.init_params(spec, seed=0)
    torch.manual_seed(3)
    x = torch.randn(1, 3, 1024, 1024)
    y = (x[:, 0] > 0).long()
    ref_logits, ref_loss, ref_grads, _ = O.loss_and_grads(sd, x, y, spec)
    model = b200unet.UNet(3, 2, 4, 5, True, False, "upconv").cuda().train()
    model.load_state_dict(sd)
    logits = model(x.cuda())
    loss = F.cross_entropy(logits, y.cuda())
    loss.backward()
    e = rel_l2(logits.detach().cpu(), ref_logits)
    agree = float((logits.argmax(1).cpu() == ref_logits.argmax(1)).float().mean())
    keys = list(ref_grads)
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_grads[k].flatten() for k in keys]))
    print(f"[config4 1024^2 b1] logits rel-L2 {e:.3e} argmax agreement {agree:.5f} grad rel-L2 {eg:.3e} "
          f"loss {float(loss):.6f} vs {float(ref_loss):.6f}")
    assert e <= 1e-2 and eg <= 2e-2 and agree >= 0.999
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * max(1.0, abs(float(ref_loss)))


def _oracle_chunked(sd, x, y, spec, chunk):
    """The fp32 CPU oracle on a batch too large to hold at once: without BatchNorm every image is independent, the loss
    is the mean over all N*H'*W' pixels (README.md:58), so logits are the per-chunk logits and the gradient is the
    pixel-weighted mean of the per-chunk gradients.  Same arithmetic as one call, bounded host memory."""
    assert not spec.batch_norm
    n = x.shape[0]
    logits, loss, grads = [], 0.0, None
    for i in range(0, n, chunk):
        lg, ls, g, _ = O.loss_and_grads(sd, x[i:i + chunk], y[i:i + chunk], spec)
        wgt = x[i:i + chunk].shape[0] / n
        logits.append(lg)
        loss += float(ls) * wgt
        grads = {k: v * wgt for k, v in g.items()} if grads is None else {k: grads[k] + g[k] * wgt for k in g}
    return torch.cat(logits), loss, grads


def _check_vs_oracle(tag, args, b, h, w, chunk, seed):
    """Full BASELINE batch through the module vs the CPU oracle; north_star tolerances verbatim."""
    import b200unet
    spec = O.UNetSpec(*args[:7])
    sd = O.init_params(spec, seed=0)
    torch.manual_seed(seed)
    x = torch.randn(b, args[0], h, w)
    ho, wo = O.output_hw(spec, h, w)
    dy, dx = (h - ho) // 2, (w - wo) // 2
    y = (x[:, 0, dy:dy + ho, dx:dx + wo] > 0).long()
    ref_logits, ref_loss, ref_grads = _oracle_chunked(sd, x, y, spec, chunk)
    model = b200unet.UNet(*args).cuda().train()
    model.load_state_dict(sd)
    logits = model(x.cuda())
    loss = F.cross_entropy(logits, y.cuda())
    loss.backward()
    e = rel_l2(logits.detach().cpu(), ref_logits)
    agree = float((logits.argmax(1).cpu() == ref_logits.argmax(1)).float().mean())
    keys = list(ref_grads)
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_grads[k].flatten() for k in keys]))
    print(f"[{tag}] logits rel-L2 {e:.3e} argmax agreement {agree:.5f} grad rel-L2 {eg:.3e} "
          f"loss {float(loss):.6f} vs {ref_loss:.6f}")
    assert e <= 1e-2 and eg <= 2e-2 and agree >= 0.999   # BASELINE.json north_star tolerances
    assert abs(float(loss) - ref_loss) <= 1e-3 * max(1.0, abs(ref_loss))


def test_config3_batch32_572_vs_oracle():
    """The headline configuration (BASELINE configs[2]: paper default, 1x572x572, batch 32 per GPU) at its full batch."""
    _check_vs_oracle("config3 572^2 b32", (1, 2, 5, 6, False, False, "upconv"), 32, 572, 572, chunk=4, seed=5)


def test_config4_batch8_1024_vs_oracle():
    """BASELINE configs[3] (in=3, depth 4, wf 5, same padding, 1024x1024) at its full batch of 8."""
    _check_vs_oracle("config4 1024^2 b8", (3, 2, 4, 5, True, False, "upconv"), 8, 1024, 1024, chunk=1, seed=6)


BN_FULL = {
    # name: (ctor args, up_block, batch, H, W, gradient tolerance)
    "config2_same_bn_upsample_b16_256": ((1, 2, 5, 6, True, True, "upsample"), "paper", 16, 256, 256, 2e-2),
    "config5_deep_feature_b12_192x640": ((3, 6, 5, 2, True, True, "upsample", True), "deep", 12, 192, 640, 2e-2),
}


@pytest.mark.parametrize("name", list(BN_FULL))
def test_batchnorm_configs_full_size_vs_oracle(name):
    import b200unet
    args, up_block, b, h, w, tol_grad = BN_FULL[name]
    spec = O.UNetSpec(*args[:7], non_neg=(args[7] if len(args) > 7 else False), up_block=up_block)
    sd = O.init_params(spec, seed=0)
    torch.manual_seed(2)
    x = (torch.rand if name.startswith("config5") else torch.randn)(b, args[0], h, w)
    c = x[:, 0].contiguous()
    qs = torch.quantile(c.flatten()[:1 << 20], torch.linspace(0, 1, spec.n_classes + 1)[1:-1])
    y = torch.bucketize(c, qs)
    ref_logits, ref_loss, ref_grads, ref_stats = O.loss_and_grads(sd, x, y, spec)
    model = b200unet.UNet(*args, up_block=up_block).cuda().train()
    assert model.precision == "split"
    model.load_state_dict(sd)
    logits = model(x.cuda())
    loss = F.cross_entropy(logits, y.cuda())
    loss.backward()
    e = rel_l2(logits.detach().cpu(), ref_logits)
    agree = float((logits.argmax(1).cpu() == ref_logits.argmax(1)).float().mean())
    keys = list(ref_grads)
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_grads[k].flatten() for k in keys]))
    print(f"[{name}] logits rel-L2 {e:.3e} argmax agreement {agree:.5f} grad rel-L2 {eg:.3e} "
          f"loss {float(loss):.6f} vs {float(ref_loss):.6f}")
    assert e <= 1e-2 and agree >= 0.999 and eg <= tol_grad
    assert abs(float(loss) - float(ref_loss)) <= 1e-4 * max(1.0, abs(float(ref_loss)))
    after = model.state_dict()
    for k, v in ref_stats.items():
        assert torch.allclose(after[k].cpu(), v, rtol=1e-3, atol=1e-5), k


FULL = {
    # name: (ctor args, up_block, batch, H, W, has_bn)
    "config3_paper_valid_b32_572": ((1, 2, 5, 6, False, False, "upconv"), "paper", 32, 572, 572, False),
    "config2_same_bn_upsample_b16_256": ((1, 2, 5, 6, True, True, "upsample"), "paper", 16, 256, 256, True),
    "config4_d4_wf5_in3_same_b8_1024": ((3, 2, 4, 5, True, False, "upconv"), "paper", 8, 1024, 1024, False),
    "config5_deep_feature_b12_192x640": ((3, 6, 5, 2, True, True, "upsample", True), "deep", 12, 192, 640, True),
}


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_properties(name):
    args, up_block, b, h, w, has_bn = FULL[name]
    model = make(args, up_block)
    spec = O.UNetSpec(*args[:7], non_neg=(args[7] if len(args) > 7 else False), up_block=up_block)
    ho, wo = O.output_hw(spec, h, w)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.rand if name.startswith("config5") else torch.randn)(b, args[0], h, w, device="cuda", generator=g)
    y = torch.randint(0, args[1], (b, ho, wo), device="cuda", generator=g)
    bn_state = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}

    def reset_bn():
        if bn_state:
            model.load_state_dict(bn_state, strict=False)

    loss1, g1 = grads_of(model, x, y)
    assert math.isfinite(loss1) and 0.2 * math.log(args[1]) < loss1 < 5 * math.log(args[1]) + 1
    assert all(torch.isfinite(v).all() for v in g1.values())
    # bit-exact reproducibility
    reset_bn()
    loss2, g2 = grads_of(model, x, y)
    assert loss1 == loss2
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k
    # linearity in the upstream gradient (bf16 storage of activation gradients: not exact, but close)
    reset_bn()
    _, g3 = grads_of(model, x, y, scale=4.0)   # power of two: bf16 rounding is scale invariant -> exact
    assert rel_l2(cat(g3), 4.0 * cat(g1)) < 1e-6
    # fused loss == F.cross_entropy(logits)
    reset_bn()
    with torch.no_grad():
        lg = model(x)
    assert lg.shape == (b, args[1], ho, wo)
    assert abs(float(F.cross_entropy(lg, y)) - loss1) < 2e-4 * max(1.0, loss1)
    # batch additivity (no BatchNorm): different batch sizes use different tile plans / split counts
    if not has_bn:
        k = b // 4
        la, ga = grads_of(model, x[:k], y[:k])
        lb, gb = grads_of(model, x[k:], y[k:])
        mix = (cat(ga) * k + cat(gb) * (b - k)) / b
        e = rel_l2(mix, cat(g1))
        print(f"[{name}] batch-additivity rel-L2 {e:.3e}")
        assert e < 2e-3
        assert abs((la * k + lb * (b - k)) / b - loss1) < 1e-5 * max(1.0, loss1)
    print(f"[{name}] loss {loss1:.5f} ok")
