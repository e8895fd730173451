"""Drop-in behaviour of b200unet.UNet beyond the plain training step, through the C ABI:

  * gradient w.r.t. the input image (the reference forward is differentiable in x, unet.py:73-84);
  * backward through an eval-mode (frozen statistics) BatchNorm model — fine-tuning with `model.eval()`;
  * any in_channels (unet.py:49-52 accepts every count), with and without BatchNorm;
  * the full-size evaluation shape of the repo's feature net, 3x480x640 (options.py:105-107), eval mode + no_grad;
  * checkpoints shaped like run.py:418-431 (`model_state_dict`, `optimizer_state_dict`): reference -> here -> back.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import rel_l2
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)


def _build(spec: O.UNetSpec, **kw):
    import b200unet
    return b200unet.UNet(spec.in_channels, spec.n_classes, spec.depth, spec.wf, spec.padding, spec.batch_norm,
                         spec.up_mode, spec.non_neg, up_block=spec.up_block, **kw)


def _oracle_with_input_grad(sd, x, y, spec, training=True):
    shapes = O.param_shapes(spec)
    leaves = {k: (v.detach().clone().requires_grad_(True) if k in shapes else v.clone()) for k, v in sd.items()}
    xi = x.clone().requires_grad_(True)
    logits = O.forward(leaves, xi, spec, training=training)
    loss = F.cross_entropy(logits, y)
    names = list(shapes)
    grads = torch.autograd.grad(loss, [xi] + [leaves[k] for k in names])
    return logits.detach(), float(loss), grads[0], dict(zip(names, grads[1:]))


@pytest.mark.parametrize("case", ["paper_valid_in1", "paper_same_in3", "deep_bn_in3"])
def test_input_gradient_matches_oracle(case):
    spec, (n, h, w) = {
        "paper_valid_in1": (O.UNetSpec(1, 2, 3, 6, False, False, "upconv"), (2, 92, 108)),
        "paper_same_in3": (O.UNetSpec(3, 2, 3, 5, True, False, "upconv"), (1, 70, 54)),
        "deep_bn_in3": (O.UNetSpec(3, 4, 3, 3, True, True, "upsample", True, "deep"), (2, 40, 48)),
    }[case]
    sd = O.init_params(spec, seed=3)
    torch.manual_seed(11)
    x = torch.randn(n, spec.in_channels, h, w)
    ho, wo = O.output_hw(spec, h, w)
    dy, dx = (h - ho) // 2, (w - wo) // 2
    c = x[:, 0, dy:dy + ho, dx:dx + wo].contiguous()
    y = torch.bucketize(c, torch.quantile(c.flatten(), torch.linspace(0, 1, spec.n_classes + 1)[1:-1]))
    _, ref_loss, ref_dx, ref_g = _oracle_with_input_grad(sd, x, y, spec)
    model = _build(spec).cuda().train()
    model.load_state_dict(sd)
    xg = x.cuda().requires_grad_(True)
    loss = F.cross_entropy(model(xg), y.cuda())
    loss.backward()
    assert xg.grad is not None and xg.grad.shape == x.shape and xg.grad.dtype == torch.float32
    e = rel_l2(xg.grad.cpu(), ref_dx)
    keys = list(ref_g)
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_g[k].flatten() for k in keys]))
    cos = float(F.cosine_similarity(xg.grad.cpu().flatten().double(), ref_dx.flatten().double(), dim=0))
    print(f"[{case}] input-gradient rel-L2 {e:.3e} (cosine {cos:.5f}), weight-gradient rel-L2 {eg:.3e}")
    # north_star sets no tolerance for dL/dx.  Unlike a weight gradient (a sum over ~10^4..10^7 pixels, in which the
    # bf16 rounding of the stored activation gradients averages out) it is a per-pixel quantity, so it carries the bf16
    # storage noise un-averaged: the CPU oracle with bf16 storage EMULATED (act_bf16=True) differs from the fp32 oracle
    # by 8.2e-2 / 8.4e-2 on the first two cases — the kernels measure 8.2e-2 on the first.  Bound: 1.2e-1 and a
    # direction within 0.5 %; the weight gradients of the same step keep the 2e-2 bar.
    assert e <= 1.2e-1 and cos >= 0.995
    assert eg <= (6e-2 if spec.batch_norm else 2e-2)
    # frozen parameters, only the image requires grad (feature visualisation / adversarial use)
    for p in model.parameters():
        p.requires_grad_(False)
    model.zero_grad(set_to_none=True)
    xg2 = x.cuda().requires_grad_(True)
    bn_state = {k: v.clone() for k, v in sd.items() if "running" in k or "num_batches" in k}
    model.load_state_dict(bn_state, strict=False)
    F.cross_entropy(model(xg2), y.cuda()).backward()
    assert rel_l2(xg2.grad, xg.grad) < 1e-6
    assert all(p.grad is None for p in model.parameters())


def test_backward_through_eval_mode_batchnorm():
    """model.eval() then loss.backward(): BatchNorm uses and back-propagates through its running statistics as
    constants (torch's batch_norm backward with training=False), which the reference supports."""
    spec = O.UNetSpec(1, 3, 3, 3, True, True, "upsample")
    sd = O.init_params(spec, seed=5)
    g = torch.Generator().manual_seed(1)
    for k in list(sd):
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        if k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    x = torch.randn(2, 1, 40, 32, generator=g)
    y = torch.randint(0, 3, (2, 40, 32), generator=g)
    ref_logits, ref_loss, ref_dx, ref_g = _oracle_with_input_grad(sd, x, y, spec, training=False)
    model = _build(spec).cuda()
    model.load_state_dict(sd)
    model.eval()
    logits = model(x.cuda())
    loss = F.cross_entropy(logits, y.cuda())
    loss.backward()
    assert rel_l2(logits.detach().cpu(), ref_logits) <= 1e-2
    keys = list(ref_g)
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_g[k].flatten() for k in keys]))
    print(f"[eval-mode backward] logits {rel_l2(logits.detach().cpu(), ref_logits):.3e} grads {eg:.3e}")
    assert eg <= 2e-2
    after = model.state_dict()
    for k in sd:
        if "running" in k or "num_batches" in k:   # eval mode: statistics untouched
            assert torch.equal(after[k].cpu(), sd[k]), k


@pytest.mark.parametrize("batch_norm", [False, True], ids=["no-bn", "bn"])
@pytest.mark.parametrize("cin", [1, 2, 3, 4, 5, 6, 7, 8, 9, 12])
def test_any_in_channels(cin, batch_norm):
    """unet.py:49-52 takes any in_channels; 5, 6, 7, 9.. are carried zero-padded to a multiple of 8 so that the
    tensor-core kernels run the first layer (they used to raise in the split tier / fall to the CUDA-core kernels)."""
    spec = O.UNetSpec(cin, 2, 2, 4, True, batch_norm, "upconv")
    sd = O.init_params(spec, seed=cin)
    torch.manual_seed(cin)
    x = torch.randn(2, cin, 72, 60)   # enough pixels for the bf16 rounding noise of a weight gradient to average out
    y = (x[:, 0] > 0).long()
    ref_logits, ref_loss, ref_g, _ = O.loss_and_grads(sd, x, y, spec)
    model = _build(spec).cuda().train()
    model.load_state_dict(sd)
    logits = model(x.cuda())
    F.cross_entropy(logits, y.cuda()).backward()
    assert rel_l2(logits.detach().cpu(), ref_logits) <= 1e-2
    keys = list(ref_g)
    for k in keys:
        assert dict(model.named_parameters())[k].grad.shape == ref_g[k].shape, k
    eg = rel_l2(torch.cat([dict(model.named_parameters())[k].grad.cpu().flatten() for k in keys]),
                torch.cat([ref_g[k].flatten() for k in keys]))
    assert eg <= 2e-2, (cin, batch_norm, eg)


def test_eval_shape_480x640_feature_net_vs_oracle():
    """N2: the repo's evaluation runs the feature net on full 3x480x640 frames (options.py:105-107) under
    model.eval() + torch.no_grad() (run.py:106-108, 264): eval-mode BatchNorm (running statistics), no tape."""
    spec = O.UNetSpec(3, 6, 5, 2, True, True, "upsample", True, "deep")   # options.py:4-25 after setto()
    sd = O.init_params(spec, seed=9)
    g = torch.Generator().manual_seed(2)
    for k in list(sd):   # a trained checkpoint has non-trivial statistics
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05 + 0.1
        if k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.25
    x = torch.rand(2, 3, 480, 640, generator=g)   # the repo feeds [0, 1] RGB (dataloader.py:258-264)
    ref = O.forward(sd, x, spec, training=False)
    import b200unet
    model = b200unet.UNet(3, 6, 5, 2, True, True, "upsample", True).cuda()   # network_modules.py:73
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        out = model(x.cuda())
    assert out.shape == (2, 6, 480, 640) and out.dtype == torch.float32
    e = rel_l2(out.cpu(), ref)
    agree = float((out.argmax(1).cpu() == ref.argmax(1)).float().mean())
    print(f"[eval 3x480x640] features rel-L2 {e:.3e} argmax agreement {agree:.5f}")
    assert e <= 1e-2 and agree >= 0.999
    assert float(out.min()) >= 0.0   # non_neg head (unet.py:65-69)


def test_run_py_shaped_checkpoint_round_trip(tmp_path):
    """run.py:418-431 saves {'epoch', 'model_state_dict', 'loss_model_state_dict', 'optimizer_state_dict', 'loss'} and
    run.py:101-116 loads it back.  A checkpoint holding the REFERENCE's own state_dict (golden fixture written by the
    unmodified unet.py) and a torch.optim.Adam state must load into b200unet.UNet + FusedAdam, training must continue
    from it exactly like torch.optim.Adam would, and what we save must load back into torch.optim.Adam."""
    import b200unet
    z = np.load(os.path.join(HERE, "golden", "deep_cfg5_d5_wf2.npz"))
    spec = json.loads(bytes(z["spec"]).decode())
    ref_sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    x, y = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["y"]).cuda()
    args = (spec["in_channels"], spec["n_classes"], spec["depth"], spec["wf"], spec["padding"], spec["batch_norm"],
            spec["up_mode"], spec["non_neg"])

    def one_step(model, opt):
        loss = model.loss(x, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss)

    # "the reference trained for a while": two steps with torch.optim.Adam, then a run.py-shaped checkpoint file
    m0 = b200unet.UNet(*args).cuda().train()
    m0.load_state_dict(ref_sd)
    o0 = torch.optim.Adam(m0.parameters(), lr=1e-3)
    for _ in range(2):
        one_step(m0, o0)
    path = os.path.join(tmp_path, "epoch00_2000.pth")
    torch.save({"epoch": 0, "model_state_dict": m0.state_dict(), "loss_model_state_dict": {},
                "optimizer_state_dict": o0.state_dict(), "loss": 0.5}, path)
    ck = torch.load(path)
    assert list(ck["model_state_dict"].keys()) == list(ref_sd.keys())   # the reference's schema, key for key
    # continue with torch.optim.Adam (what run.py does) ...
    l_ref = one_step(m0, o0)
    # ... and from the checkpoint with b200unet.FusedAdam
    m1 = b200unet.UNet(*args).cuda().train()                            # run.py:103, 114
    m1.load_state_dict(ck["model_state_dict"])
    o1 = b200unet.FusedAdam(m1.parameters(), lr=1e-3, model=m1)
    o1.load_state_dict(ck["optimizer_state_dict"])                      # run.py:105, 116
    l1 = one_step(m1, o1)
    assert abs(l1 - l_ref) <= 1e-5 * max(1.0, abs(l_ref))
    for (k, a), (_, b) in zip(m0.state_dict().items(), m1.state_dict().items()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-4, atol=1e-6), k
    # and back: our checkpoint into torch.optim.Adam
    path2 = os.path.join(tmp_path, "epoch00_2001.pth")
    torch.save({"epoch": 0, "model_state_dict": m1.state_dict(), "loss_model_state_dict": {},
                "optimizer_state_dict": o1.state_dict(), "loss": l1}, path2)
    ck2 = torch.load(path2)
    m2 = b200unet.UNet(*args).cuda().train()
    m2.load_state_dict(ck2["model_state_dict"])
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    o2.load_state_dict(ck2["optimizer_state_dict"])
    l2, l0 = one_step(m2, o2), one_step(m0, o0)
    assert abs(l2 - l0) <= 1e-5 * max(1.0, abs(l0))


def test_out_of_range_labels_poison_the_loss():
    """F.cross_entropy raises a device-side assert for labels outside [0, n_classes) (other than ignore_index); the fused
    loss must not silently treat them as class 0: the loss becomes NaN."""
    spec = O.UNetSpec(1, 3, 2, 3, True, False, "upconv")
    model = _build(spec).cuda().train()
    x = torch.randn(1, 1, 16, 16, device="cuda")
    y = torch.randint(0, 3, (1, 16, 16), device="cuda")
    assert torch.isfinite(model.loss(x, y))
    y[0, 3, 4] = 7
    assert torch.isnan(model.loss(x, y))
    y[0, 3, 4] = -100   # ignore_index stays legal
    assert torch.isfinite(model.loss(x, y))


def test_cuda_core_fallback_is_counted_and_still_correct(capfd):
    """B200_IMPL_AUTO must not drop to the CUDA-core kernels silently (10-50x slower): a 12-channel convolution (not a
    multiple of 8, so no TMA / tcgen05) is served by the fallback, counted by b200unet_fallback_count() and reported on
    stderr once; the 16-channel one next to it stays on the tensor cores and is not counted."""
    import torch.nn.functional as F
    from b200unet import ops, load_library
    lib = load_library()
    torch.manual_seed(0)
    for cin, expect_fallback in [(12, True), (16, False)]:
        x = torch.randn(1, 20, 24, cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(16, cin, 3, 3, device="cuda") * 0.1)
        n0 = lib.b200unet_fallback_count()
        y = ops.conv_fwd([x], w, None, 1, True)
        assert (lib.b200unet_fallback_count() > n0) == expect_fallback
        want = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float() if not expect_fallback else w, padding=1))
        got = y.float().permute(0, 3, 1, 2)
        assert float((got - want).norm() / want.norm()) < 1e-2
    capfd.readouterr()  # the stderr note is printed once per process and entry point: not asserted here
