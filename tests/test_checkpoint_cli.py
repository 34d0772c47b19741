"""CPU: the checkpoint validator (motionmixerconv_b200/checkpoint.py) on state_dicts written by the REFERENCE modules (oracle/_ref)
and on the golden fixtures' parameters; detects missing / reordered keys, wrong shapes, broken se2 aliases."""
import json

import numpy as np
import pytest
import torch

from motionmixerconv_b200 import checkpoint as CK
from oracle import make_ref
from tests.golden_util import Golden, golden_cases


@pytest.mark.parametrize("case", golden_cases())
def test_fixture_parameters_validate(case):
    g = Golden(case)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in g.params.items()}
    assert CK.validate_state_dict(sd, g.family, g.cfg) == []


@pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not staged")
def test_reference_checkpoint_file_through_the_cli(tmp_path, capsys):
    _, RefConv, _ = make_ref.import_reference()
    cfg = dict(num_blocks=2, dimPosIn=33, dimPosEmb=48, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=4, conv1_kernel_shape=[5, 5],
               mode_conv="twice", activation="mish", regularization=-1.0, use_se=True, r_se=8, encoder_n_harmonic_functions=8, encoder_omega0=0.1)
    torch.manual_seed(0)
    ref = RefConv(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items()})
    path = tmp_path / "model.pt"
    torch.save(ref.state_dict(), path)
    assert CK.main([str(path), "--family", "conv", "--cfg", json.dumps(cfg)]) == 0
    assert "OK" in capsys.readouterr().out
    sd = ref.state_dict()
    bad = dict(sd)
    bad["Mixer_Block.0.se2.excitationBlock.0.weight"] = sd["Mixer_Block.0.se2.excitationBlock.0.weight"] + 1.0      # alias broken
    bad["LN.weight"] = torch.zeros(7)                                                                                # wrong shape
    del bad["fc_out.bias"]                                                                                           # missing
    problems = CK.validate_state_dict(bad, "conv", cfg)
    text = "\n".join(problems)
    assert "missing keys" in text and "LN.weight: shape" in text and "se2" in text
    reordered = dict(reversed(list(sd.items())))
    assert any("key order" in p for p in CK.validate_state_dict(reordered, "conv", cfg))
