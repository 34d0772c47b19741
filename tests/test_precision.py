"""CPU: the precision switch of the MixerBlock kernels (include/mmx.h MMX_PREC_*) — host-side plumbing only.
The reference's constructor signatures stay untouched: precision travels as a module attribute / process default."""
import ctypes as C
import inspect
import re
import os

import pytest

from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import functional as F_
from motionmixerconv_b200.mlp_mixer import MixerBlock, MlpMixer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_constants_match_binding():
    hdr = open(os.path.join(ROOT, "include", "mmx.h")).read()
    vals = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define MMX_PREC_(\w+) (\d+)", hdr)}
    assert vals == L.MMX_PREC
    # `precision` is the last field of MmxMlpBlockDesc in the header and in the ctypes mirror
    body = hdr[hdr.index("typedef struct {\n    int B, T, H, tok, ch;"):hdr.index("} MmxMlpBlockDesc;")]
    assert body.strip().splitlines()[-1].strip().startswith("int precision;")
    assert L.MmxMlpBlockDesc._fields_[-1] == ("precision", C.c_int)


def test_desc_carries_precision_and_default():
    old = F_.get_precision()
    try:
        d = F_.mlp_block_desc(4, 10, 50, 20, 50, 1, "mish", True, False, True, 0, 0.1, 1, 2)
        assert d.precision == L.MMX_PREC[old]
        F_.set_precision("tf32")
        assert F_.mlp_block_desc(4, 10, 50, 20, 50, 1, "mish", True, False, True, 0, 0.1, 1, 2).precision == 1
        assert F_.mlp_block_desc(4, 10, 50, 20, 50, 1, "mish", True, False, True, 0, 0.1, 1, 2, "fp32").precision == 0
        with pytest.raises(ValueError):
            F_.set_precision("bf16")
    finally:
        F_.set_precision(old)


def test_model_attribute_not_constructor_argument():
    cfg = dict(num_classes=66, num_blocks=2, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10, pred_len=10,
               activation="mish", regularization=0.1, input_size=66, r_se=8, use_se=True)
    assert "precision" not in inspect.signature(MlpMixer.__init__).parameters          # reference signature (mlp_mixer.py:254-258)
    assert "precision" not in inspect.signature(MixerBlock.__init__).parameters
    m = MlpMixer(**cfg)
    assert all(mb.precision is None and mb.meta()[-1] is None for mb in m.Mixer_Block)
    assert m.set_precision("tf32") is m
    assert all(mb.meta()[-1] == "tf32" for mb in m.Mixer_Block)
    assert "precision" not in "".join(m.state_dict().keys())                            # state_dict layout untouched
    with pytest.raises(ValueError):
        m.set_precision("fp8")
    m.set_precision(None)
    assert all(mb.precision is None for mb in m.Mixer_Block)
