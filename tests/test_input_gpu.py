"""Input pipeline at the boundary (reference dataloader.py:258-264, 553-579): a uint8 NHWC batch
staged through pinned memory and converted by one kernel must give the model exactly what the reference pipeline
(image / 255 -> float32 CHW -> module) gives it."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _reference_pipeline(u8_nhwc: np.ndarray) -> torch.Tensor:
    """dataloader.py:258-264 (img.astype(float) / 255), ToTensor transpose (2,0,1) and .to(device, dtype=torch.float)"""
    img = u8_nhwc.astype(float) / 255
    return torch.from_numpy(img.transpose(0, 3, 1, 2)).to("cuda", dtype=torch.float)


@pytest.mark.parametrize("spec", [O.UNetSpec(1, 2, 3, 6, False, False, "upconv"),
                                  O.UNetSpec(3, 6, 3, 2, True, True, "upsample", True, "deep"),
                                  O.UNetSpec(5, 2, 2, 4, True, False, "upconv")],
                         ids=["gray-paper-bf16", "rgb-feature-net-split", "5ch-padded"])
def test_u8_batch_equals_reference_pipeline(spec):
    import b200unet
    rng = np.random.default_rng(0)
    n, h, w = 2, 60, 76
    u8 = rng.integers(0, 256, (n, h, w, spec.in_channels), dtype=np.uint8)
    torch.manual_seed(0)
    model = b200unet.UNet(spec.in_channels, spec.n_classes, spec.depth, spec.wf, spec.padding, spec.batch_norm,
                          spec.up_mode, spec.non_neg, up_block=spec.up_block).cuda().train()
    ho, wo = O.output_hw(spec, h, w)
    y = torch.randint(0, spec.n_classes, (n, ho, wo), device="cuda")
    bn = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
    # reference-shaped input
    x32 = _reference_pipeline(u8)
    ref_logits = model(x32)
    F.cross_entropy(ref_logits, y).backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    model.load_state_dict(bn, strict=False)
    # uint8 pipeline
    packed = b200unet.pack_images(model, torch.from_numpy(u8).cuda())
    assert packed.shape == (n, spec.in_channels, h, w)
    logits = model(packed)
    F.cross_entropy(logits, y).backward()
    assert torch.equal(logits, ref_logits)                      # bit-identical operand -> bit-identical network
    for k, p in model.named_parameters():
        assert torch.equal(p.grad, ref_grads[k]), k


def test_mean_std_normalisation_and_grayscale_rank3():
    import b200unet
    from b200unet import ops
    u8 = torch.randint(0, 256, (2, 20, 24, 3), dtype=torch.uint8, device="cuda")
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    model = b200unet.UNet(3, 2, 2, 4, True, False, "upconv").cuda()
    got = b200unet.pack_images(model, u8, mean, std).data.float()
    want = ((u8.float() / 255) - torch.tensor(mean, device="cuda")) / torch.tensor(std, device="cuda")
    assert torch.allclose(got, want.to(torch.bfloat16).float(), atol=2e-2, rtol=1e-2)
    g = b200unet.UNet(1, 2, 2, 4, True, False, "upconv").cuda()
    gray = torch.randint(0, 256, (2, 20, 24), dtype=torch.uint8, device="cuda")
    assert b200unet.pack_images(g, gray).shape == (2, 1, 20, 24)
    with pytest.raises(ValueError):
        b200unet.pack_images(g, u8)


def test_image_stager_double_buffering():
    import b200unet
    model = b200unet.UNet(1, 2, 2, 4, True, False, "upconv").cuda().eval()
    st = b200unet.ImageStager(model, batch=2, height=32, width=40, channels=1)
    rng = np.random.default_rng(1)
    batches = [rng.integers(0, 256, (2, 32, 40, 1), dtype=np.uint8) for _ in range(5)]
    outs = []
    st.put(batches[0])
    with torch.no_grad():
        for i in range(5):
            x = st.get()
            if i + 1 < 5:
                st.put(batches[i + 1])          # travels while this batch is processed
            outs.append(model(x))
        for i in range(5):
            assert torch.equal(outs[i], model(_reference_pipeline(batches[i])))
    assert st.h2d_bytes_per_batch == 2 * 32 * 40
    with pytest.raises(RuntimeError):
        st.get()
