"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs ``/root/reference``; the GPU box does not have it):

    python tests/golden/make_golden.py

For every case it instantiates the unmodified reference module
(``h36m.mlp_mixer.MlpMixer`` / ``h36m.conv_mixer_model.ConvMixer``), seeds it, runs
forward -> ``mpjpe_error`` -> backward on CPU fp32, then 3 ``torch.optim.Adam(lr=1e-3,
weight_decay=1e-5)`` steps exactly as ``h36m/train_mixer_h36m.py:63,126,180-193`` does, and
stores inputs, parameters, outputs, gradients and post-Adam parameters in ``<case>.npz``.
The fixtures pin both the numpy oracle (``tests/test_oracle_golden.py``, CPU) and the
CUDA path (``tests/test_gpu_golden.py``, ``-m gpu``).
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("MMX_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from h36m.conv_mixer_model import ConvMixer  # noqa: E402
from h36m.mlp_mixer import MlpMixer  # noqa: E402
from h36m.utils.utils_mixer import mpjpe_error  # noqa: E402

from tests.synthetic import synthetic_pose_windows  # noqa: E402

CASES = {
    # ---- MlpMixer -------------------------------------------------------------------
    "mlp_k2": dict(family="mlp", B=6, scale="h36m", cfg=dict(
        num_classes=66, num_blocks=4, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10,
        pred_len=10, activation="mish", regularization=0, input_size=66, r_se=8, use_se=True)),
    "mlp_k4": dict(family="mlp", B=4, scale="amass", cfg=dict(
        num_classes=54, num_blocks=2, hidden_dim=128, tokens_mlp_dim=20, channels_mlp_dim=128, seq_len=10,
        pred_len=25, activation="gelu", regularization=0, input_size=54, r_se=8, use_se=True)),
    "mlp_odd_nose": dict(family="mlp", B=5, scale="amass", cfg=dict(
        num_classes=9, num_blocks=2, hidden_dim=24, tokens_mlp_dim=12, channels_mlp_dim=40, seq_len=8,
        pred_len=5, activation="gelu", regularization=0, input_size=9, use_se=False)),
    "mlp_maxpool": dict(family="mlp", B=5, scale="amass", cfg=dict(
        num_classes=33, num_blocks=2, hidden_dim=36, tokens_mlp_dim=16, channels_mlp_dim=36, seq_len=10,
        pred_len=10, activation="mish", regularization=0, input_size=33, r_se=4, use_max_pooling=True,
        use_se=True)),
    "mlp_bn": dict(family="mlp", B=8, scale="amass", cfg=dict(
        num_classes=33, num_blocks=2, hidden_dim=32, tokens_mlp_dim=20, channels_mlp_dim=32, seq_len=10,
        pred_len=10, activation="gelu", regularization=-1.0, input_size=33, r_se=8, use_se=True)),
    # BatchNorm1d + the max squeeze, mish (mlp_mixer.py:21 with :72-73)
    "mlp_bn_maxpool": dict(family="mlp", B=7, scale="amass", cfg=dict(
        num_classes=33, num_blocks=2, hidden_dim=40, tokens_mlp_dim=12, channels_mlp_dim=24, seq_len=10,
        pred_len=10, activation="mish", regularization=-1.0, input_size=33, r_se=4, use_max_pooling=True,
        use_se=True)),
    # ---- ConvMixer ------------------------------------------------------------------
    "conv_k1": dict(family="conv", B=6, scale="h36m", cfg=dict(
        num_blocks=4, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1,
        conv1_kernel_shape=(1, 3), conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice",
        activation="mish", regularization=0, use_se=True, r_se=8, encoder_n_harmonic_functions=8,
        encoder_omega0=0.1)),
    "conv_harm64": dict(family="conv", B=4, scale="h36m", cfg=dict(
        num_blocks=1, dimPosIn=6, dimPosEmb=20, dimPosOut=6, in_nTP=10, out_nTP=15, conv_nChan=2,
        conv1_kernel_shape=(1, 3), conv1_padding=(0, 1), mode_conv="twice", activation="gelu",
        regularization=0, use_se=True, r_se=4, encoder_n_harmonic_functions=64, encoder_omega0=0.1)),
    "conv_k3": dict(family="conv", B=3, scale="ais", cfg=dict(
        num_blocks=2, dimPosIn=33, dimPosEmb=192, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=4,
        conv1_kernel_shape=(5, 9), mode_conv="twice", activation="mish", regularization=0, use_se=True,
        r_se=8, encoder_n_harmonic_functions=0, encoder_omega0=0)),
    "conv_k3_bn": dict(family="conv", B=4, scale="ais", cfg=dict(
        num_blocks=2, dimPosIn=33, dimPosEmb=64, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=4,
        conv1_kernel_shape=(5, 9), mode_conv="twice", activation="mish", regularization=-1.0, use_se=True,
        r_se=8, encoder_n_harmonic_functions=0, encoder_omega0=0)),
    "conv_once_se": dict(family="conv", B=4, scale="ais", cfg=dict(
        num_blocks=2, dimPosIn=33, dimPosEmb=48, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=8,
        conv1_kernel_shape=(5, 5), mode_conv="once", activation="mish", regularization=0, use_se=True,
        r_se=8, encoder_n_harmonic_functions=0, encoder_omega0=0)),
    "conv_once_nose": dict(family="conv", B=4, scale="ais", cfg=dict(
        num_blocks=2, dimPosIn=12, dimPosEmb=20, dimPosOut=12, in_nTP=6, out_nTP=4, conv_nChan=2,
        conv1_kernel_shape=(3, 3), mode_conv="once", activation="gelu", regularization=0, use_se=False,
        encoder_n_harmonic_functions=0, encoder_omega0=0)),
    # the reference's Optuna grid (optuna_search/conv_optuna_main.py:339-342): C = 8, E = 192, a kernel beyond the fused kernels' tile
    "conv_c8_k5x9": dict(family="conv", B=2, scale="ais", cfg=dict(
        num_blocks=1, dimPosIn=33, dimPosEmb=192, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=8,
        conv1_kernel_shape=(5, 9), mode_conv="twice", activation="mish", regularization=0, use_se=True,
        r_se=8, encoder_n_harmonic_functions=0, encoder_omega0=0)),
    # BatchNorm2d + the max squeeze (conv_mixer_model.py:60-62 with :115-116)
    "conv_maxpool_bn": dict(family="conv", B=5, scale="ais", cfg=dict(
        num_blocks=2, dimPosIn=33, dimPosEmb=64, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=4,
        conv1_kernel_shape=(5, 5), mode_conv="twice", activation="gelu", regularization=-1.0, use_se=True,
        r_se=4, use_max_pooling=True, encoder_n_harmonic_functions=0, encoder_omega0=0)),
    "conv_evenk": dict(family="conv", B=3, scale="ais", cfg=dict(
        num_blocks=1, dimPosIn=33, dimPosEmb=40, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=3,
        conv1_kernel_shape=(1, 29), mode_conv="twice", activation="gelu", regularization=0, use_se=True,
        r_se=4, use_max_pooling=True, encoder_n_harmonic_functions=0, encoder_omega0=0)),
}


def _np(t):
    return t.detach().cpu().numpy().copy()


def build(case):
    spec = CASES[case]
    cfg = spec["cfg"]
    torch.manual_seed(0)
    if spec["family"] == "mlp":
        model = MlpMixer(**cfg)
        T, To, D, Dout = cfg["seq_len"], cfg["pred_len"], cfg["input_size"], cfg["num_classes"]
    else:
        model = ConvMixer(**cfg)
        T, To, D, Dout = cfg["in_nTP"], cfg["out_nTP"], cfg["dimPosIn"], cfg["dimPosOut"]
    assert D == Dout
    x, gt = synthetic_pose_windows(spec["B"], T, To, D, scale=spec["scale"], seed=1234)
    x = torch.from_numpy(x)
    gt = torch.from_numpy(gt)
    out = {"cfg": json.dumps(cfg), "family": spec["family"], "x": _np(x), "gt": _np(gt)}
    model.train()
    for k, v in model.state_dict().items():
        out["p/" + k] = _np(v)
    # eval-mode forward first (does not touch BN running stats)
    model.eval()
    with torch.no_grad():
        out["pred_eval"] = _np(model(x))
    model.train()
    xg = x.clone().requires_grad_(True)
    pred = model(xg)
    loss = mpjpe_error(pred, gt)
    loss.backward()
    out["pred"] = _np(pred)
    out["loss"] = np.float32(loss.item())
    out["dx"] = _np(xg.grad)
    for k, p in model.named_parameters():
        out["g/" + k] = _np(p.grad)
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["p1/" + k] = _np(v)  # BN buffers after ONE training forward
    # 3 Adam steps from the initial weights (the forward above already consumed one BN update,
    # so reload the initial state first)
    model.load_state_dict({k: torch.from_numpy(out["p/" + k]) for k in model.state_dict()})
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-05)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        l = mpjpe_error(model(x), gt)
        l.backward()
        opt.step()
        losses.append(l.item())
    out["losses"] = np.asarray(losses, dtype=np.float64)
    for k, v in model.state_dict().items():
        out["p3/" + k] = _np(v)
    return out


if __name__ == "__main__":
    torch.set_num_threads(1)
    only = sys.argv[1:]
    for case in CASES:
        if only and case not in only:
            continue
        data = build(case)
        path = os.path.join(HERE, case + ".npz")
        np.savez_compressed(path, **data)
        print("%-16s %7.1f KB  loss=%.6f" % (case, os.path.getsize(path) / 1024, float(data["loss"])))
