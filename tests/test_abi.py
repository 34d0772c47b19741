"""CPU: the C-ABI shared library loads and exports every symbol include/mmx.h declares, and the ctypes table
(_lib.SIGNATURES) lists exactly those symbols.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from motionmixerconv_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "mmx.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mmx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    assert declared_symbols() == sorted(L.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = L.bind(C.CDLL(L.LIB_PATH))          # bind() raises AttributeError on a missing symbol
    for name in declared_symbols():
        assert getattr(lib, name) is not None
    assert lib.mmx_version() >= 100
    assert lib.mmx_last_error() is not None


def test_argument_validation_needs_no_gpu():
    """Invalid descriptors are rejected before any CUDA call: error code + message, no crash."""
    lib = L.load()
    d = L.MmxMlpHeadDesc(0, 10, 10, 50, 66)
    rc = lib.mmx_mlp_head_fwd(C.byref(d), C.byref(L.MmxMlpHeadParams()), None, None, None)
    assert rc == -1 and b"null tensor" in lib.mmx_last_error()
    rc = lib.mmx_adam_step(None, None, None, None, 10, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        L.check(lib, rc, "mmx_adam_step")


def test_product_package_has_no_cpu_path():
    import torch
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    m = MlpMixer(66, 1, 50, 20, 50, 10, 10, activation="gelu", input_size=66)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 10, 66))
    import motionmixerconv_b200
    pkg = os.path.dirname(motionmixerconv_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            with open(os.path.join(pkg, fn)) as f:
                assert "oracle" not in f.read().replace("the oracle", ""), fn   # the product never imports the checker
