"""GPU (-m gpu): the drop-in claim end to end.  The reference's own ``train()`` (h36m/train_mixer_h36m.py:47-279, unmodified copy
under oracle/_ref) is run twice through tests/ref_train.py -- once with the reference's module on the CPU, once with this
package's module on the GPU (same constructor call, same initial state_dict, same synthetic dataset and DataLoader order) -- and
the per-epoch training / validation / test losses, the AUC-PCK metric and the saved ``model.pt`` must agree.

Losses are compared (functions of the parameters), not raw parameters: Adam turns the rounding noise of exactly-cancelling
gradients (a bias in front of a LayerNorm) into +-lr steps of parameters that do not influence the output.
"""
import numpy as np
import pytest
import torch

from oracle import make_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not make_ref.train_script_available(), reason="oracle/_ref training script not staged")]

MLP_CFG = dict(num_classes=66, num_blocks=2, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10, pred_len=10,
               activation="mish", regularization=0, input_size=66, r_se=8, use_se=True)
MLP_BN_CFG = dict(MLP_CFG, regularization=-1.0, activation="gelu", hidden_dim=32, channels_mlp_dim=32)
CONV_CFG = dict(num_blocks=2, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=10, conv_nChan=2, conv1_kernel_shape=(1, 3),
                conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation="mish", regularization=0, use_se=True, r_se=8)


def _both(family, cfg, tmp_path, precision=None, rtol=2e-4):
    from tests import ref_train as RT
    tr = RT.load_reference_train()
    RefMlp, RefConv, _ = make_ref.import_reference()
    if family == "mlp":
        from motionmixerconv_b200.mlp_mixer import MlpMixer as Ours
        Ref = RefMlp
    else:
        from motionmixerconv_b200.conv_mixer_model import ConvMixer as Ours
        Ref = RefConv
    torch.manual_seed(0)
    ref = Ref(**cfg)
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    ours = Ours(**cfg)
    ours.load_state_dict(init, strict=True)            # the reference's state_dict loads strictly into the drop-in
    ours = ours.to("cuda")
    if precision is not None:
        ours.set_precision(precision)
    a = RT.run_train(tr, ref, "reference_cpu", RT.train_args(str(tmp_path), "cpu"))
    b = RT.run_train(tr, ours, "dropin_gpu", RT.train_args(str(tmp_path), "cuda"))
    for key in ("train", "val", "test", "mpjpe"):
        np.testing.assert_allclose(b[key], a[key], rtol=rtol, err_msg=key)
    np.testing.assert_allclose(b["auc_pck"], a["auc_pck"], atol=2e-3)
    assert b["train"][-1] < b["train"][0]
    # the checkpoint the script wrote from the drop-in loads strictly into the reference module and reproduces its predictions
    sd = torch.load(b["state_path"], map_location="cpu")
    fresh = Ref(**cfg)
    fresh.load_state_dict(sd, strict=True)
    fresh.eval()
    ours.eval()
    x = RT.SyntheticH36M("", 10, 10, 1, split=2).data[:8, :10, :66].contiguous() / 1000
    with torch.no_grad():
        want = fresh(x)
        got = ours(x.cuda()).cpu()
    assert (got - want).abs().max().item() <= max(rtol, 1e-4) * want.abs().max().item()


def test_mlp_mixer_in_the_reference_train_loop(tmp_path):
    _both("mlp", MLP_CFG, tmp_path)


def test_mlp_mixer_tensor_core_mode_in_the_reference_train_loop(tmp_path):
    _both("mlp", MLP_CFG, tmp_path, precision="tf32", rtol=2e-3)


def test_mlp_mixer_batchnorm_in_the_reference_train_loop(tmp_path):
    _both("mlp", MLP_BN_CFG, tmp_path, rtol=5e-4)


def test_conv_mixer_in_the_reference_train_loop(tmp_path):
    _both("conv", CONV_CFG, tmp_path)
