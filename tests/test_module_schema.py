"""Drop-in boundary: constructor signatures, module tree, state_dict schema and initialisation of b200unet.UNet are
those of the reference (golden fixtures carry the reference's own state_dict)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import b200unet

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def build_from_spec(spec: dict):
    return b200unet.UNet(spec["in_channels"], spec["n_classes"], spec["depth"], spec["wf"], spec["padding"],
                         spec["batch_norm"], spec["up_mode"], spec["non_neg"], up_block=spec["up_block"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_state_dict_schema_matches_reference(path):
    z = np.load(path)
    spec = json.loads(bytes(z["spec"]).decode())
    ref_sd = {k[3:]: z[k] for k in z.files if k.startswith("sd/")}
    torch.manual_seed(1234)  # make_golden.py seeds 1234 before constructing the reference module
    m = build_from_spec(spec)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref_sd[k].shape), k
        assert str(v.dtype).replace("torch.", "") == str(ref_sd[k].dtype), k
        # same construction order and init calls => same random-init weights as the reference under the same seed
        assert np.array_equal(v.numpy(), ref_sd[k]), k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ref_sd.items()})


def test_reference_call_signatures():
    b200unet.UNet()  # README.md:14-15 defaults
    b200unet.UNet(3, 6, 5, 2, True, True, "upsample", True)  # network_modules.py:73 8-positional call
    m = b200unet.UNet(in_channels=1, n_classes=2, depth=5, wf=6, padding=False, batch_norm=False, up_mode="upconv")
    assert m.padding is False and m.depth == 5 and len(m.down_path) == 5 and len(m.up_path) == 4
    with pytest.raises(AssertionError):
        b200unet.UNet(up_mode="nearest")  # unet.py:45
    assert sum(p.numel() for p in m.parameters()) == 31030658


def test_eight_argument_call_builds_the_graph_of_unet_py():
    """unet.py (the file run.py / network_modules.py import) is the only variant with `non_neg`, and it always builds
    UNetUpBlockDeep (unet.py:60): the 8-argument call must produce that graph so that its checkpoints load; the
    7-argument README call builds the paper decoder of unet_original.py."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "deep_cfg5_d5_wf2.npz"))
    ref_sd = {k[3:]: z[k] for k in z.files if k.startswith("sd/")}   # state_dict of the unmodified unet.UNet(3,6,5,2,True,True,'upsample',True)
    torch.manual_seed(1234)
    m = b200unet.UNet(3, 6, 5, 2, True, True, "upsample", True)
    assert m.up_block == "deep"
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref_sd[k].shape), k
        assert np.array_equal(v.numpy(), ref_sd[k]), k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ref_sd.items()})
    assert b200unet.UNet(3, 6, 5, 2, True, True, "upsample", False).up_block == "deep"      # 8 arguments, non_neg False
    assert b200unet.UNet(3, 6, 5, 2, True, True, "upsample").up_block == "paper"            # 7 arguments
    assert b200unet.UNet(3, 6, 5, 2, True, True, "upsample", True, up_block="paper").up_block == "paper"
    # the last layer of the Deep graph takes 2**(wf+depth-1) channels (unet.py:60-71)
    assert m.last[0].weight.shape == (6, 2 ** (2 + 5 - 1), 1, 1)


def test_loading_the_other_decoder_variant_says_so():
    deep = b200unet.UNet(1, 2, 3, 3, False, False, "upconv", False)
    paper = b200unet.UNet(1, 2, 3, 3, False, False, "upconv")
    with pytest.raises(RuntimeError, match="up_block"):
        paper.load_state_dict(deep.state_dict())
    with pytest.raises(RuntimeError, match="up_block"):
        deep.load_state_dict(paper.state_dict())


def test_n_classes_limit_is_checked_at_construction():
    b200unet.UNet(n_classes=8)
    with pytest.raises(ValueError, match="n_classes"):
        b200unet.UNet(n_classes=9)
