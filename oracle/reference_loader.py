"""Imports the UNMODIFIED reference modules from a read-only checkout.  TEST INFRASTRUCTURE ONLY.

`unet.py:6` imports `torchsnooper`, which is neither installed nor used by that file; an empty module is
injected so the import succeeds.  Nothing is copied out of the reference tree.  The
reference exists only in the authoring container (/root/reference): GPU-box tests never call this.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_CANDIDATES = [os.environ.get("UNET_REFERENCE_DIR", ""), "/root/reference"]


def reference_dir() -> str | None:
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "unet_original.py")):
            return c
    return None


def load():
    """Returns (unet_original module, unet module) of the reference; raises if no checkout is visible."""
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("reference checkout not found (set UNET_REFERENCE_DIR)")
    if "torchsnooper" not in sys.modules:
        sys.modules["torchsnooper"] = types.ModuleType("torchsnooper")
    if d not in sys.path:
        sys.path.append(d)
    return importlib.import_module("unet_original"), importlib.import_module("unet")


def build_reference_module(spec):
    """Reference nn.Module for a UNetSpec (paper block -> unet_original.UNet, deep block -> unet.UNet)."""
    orig, deep = load()
    if spec.up_block == "paper":
        assert not spec.non_neg
        return orig.UNet(spec.in_channels, spec.n_classes, spec.depth, spec.wf, spec.padding, spec.batch_norm,
                         spec.up_mode)
    return deep.UNet(spec.in_channels, spec.n_classes, spec.depth, spec.wf, spec.padding, spec.batch_norm,
                     spec.up_mode, spec.non_neg)
