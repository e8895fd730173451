"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/b200unet.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200unet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200unet_\w+)\s*\(", text)))


def test_library_exports_header_symbols(lib_built):
    lib = ctypes.CDLL(lib_built)
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), s


def test_python_binding_covers_header(lib_built):
    from b200unet import _lib
    assert sorted(_lib.SIGNATURES.keys()) == header_symbols()
    lib = _lib.load()
    assert lib.b200unet_abi_version() == 4
    assert lib.b200unet_bn_workspace_bytes(64) > 0
    assert lib.b200unet_head_workspace_bytes(64, 2) > 0


def test_sass_is_sm100a(lib_built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sass_contains_tcgen05_tmem_and_tma(lib_built):
    """The hot path is tensor-memory code, not a CUDA-core fallback: tcgen05.mma (UTCHMMA, single-CTA and .2CTA), TMA
    loads (UTMALDG), TMEM loads / stores (LDTM / STTM) and tcgen05.commit (UTCBAR) are present in the built library."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    for mnemonic, at_least in [("UTCHMMA", 100), ("UTCHMMA.2CTA", 50), ("UTMALDG", 40), ("LDTM", 10), ("STTM", 10),
                               ("UTCBAR", 40)]:
        assert sass.count(mnemonic) >= at_least, (mnemonic, sass.count(mnemonic))


def test_bad_arguments_are_reported_not_crashed(lib_built):
    from b200unet import _lib
    lib = _lib.load()
    rc = lib.b200unet_maxpool2x2_fwd(None, None, None, None, None)
    assert rc < 0 and b"maxpool_fwd" in lib.b200unet_last_error()
    rc = lib.b200unet_conv_fwd(None, None)
    assert rc < 0


def test_no_cpu_path():
    import pytest
    import torch
    import b200unet
    m = b200unet.UNet(depth=2, wf=2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 20, 20))
