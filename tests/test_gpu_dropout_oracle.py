"""GPU (-m gpu): training-mode dropout of the FP32 kernel families against the oracle WITH THE SAME MASKS.

The kernels draw their masks from counter-based generators (a pure function of seed, step, site, element);
tests/masks_np.py holds bit-exact numpy twins, so the oracle (oracle/mixer_np.py, dropout-with-given-mask path) can be run on
exactly the masks a kernel used and forward / loss / dx / every parameter gradient compared at the FP32 bar (1e-5, noise-aware).
Families: warp-per-sequence-pair MixerBlock kernels (the K2 shape), generic CTA-per-tile MixerBlock kernels, ConvMixer halves.
(The tcgen05 family's twin is pinned and used in tests/test_gpu_mlp_tc5.py.)
"""
import numpy as np
import pytest
import torch

from oracle import mixer_np as O
from tests import masks_np as MK
from tests.golden_util import Golden, check_close, grad_scale
from tests.synthetic import synthetic_pose_windows

pytestmark = pytest.mark.gpu
TOL = 1e-5
SEED = 97531


def _run(model, x, gt):
    from motionmixerconv_b200.functional import mpjpe_error
    model.zero_grad()
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    pred = model(xg)
    loss = mpjpe_error(pred, torch.from_numpy(gt).cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    return pred.detach().cpu().numpy(), float(loss.detach()), grads, xg.grad.cpu().numpy()


def _oracles(cls, cfg, params, x, gt, masks):
    res = []
    for dt in (np.float32, np.float64):
        o = cls(cfg, params, dtype=dt)
        p = o.forward(x, training=True, masks=masks)
        l, dp = O.mpjpe(p, gt.astype(dt))
        g, dx = o.backward(dp)
        res.append((p, float(l), g, dx))
    return res


def _compare(got, o32, o64, skip=()):
    pred, loss, grads, dx = got
    check_close("pred", pred, o32[0], o64[0], rtol=TOL)
    assert abs(loss - o64[1]) <= TOL * abs(o64[1])
    floor = 5e-6 * grad_scale(o32[2])
    for k, w in o32[2].items():
        if k in grads and k not in skip:
            check_close("grad " + k, grads[k], w, o64[2][k], rtol=TOL, atol=floor)
    check_close("dx", dx, o32[3], o64[3], rtol=TOL, atol=1e-6 * float(np.abs(o32[3]).max()))


def _mlp(cfg, params):
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    m = MlpMixer(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    return m.cuda().set_precision("fp32").train()


@pytest.mark.parametrize("B", [6, 333])
def test_mlp_warp_kernels_dropout_vs_oracle(B):
    """K2 shape (T=10, tok=20, H=ch=50): the warp-per-sequence-pair kernels, the round-1 bench headline."""
    g = Golden("mlp_k2")
    c = dict(g.cfg, regularization=0.1)
    assert c["hidden_dim"] <= 64
    x, gt = (g.x, g.gt) if B == 6 else synthetic_pose_windows(B, 10, 10, 66, scale="h36m", seed=21)
    torch.manual_seed(SEED)
    model = _mlp(c, g.params)
    for step in (0, 1):
        masks = MK.mlp_warp_masks(c, len(x), SEED, step=step)
        o32, o64 = _oracles(O.MlpMixerOracle, c, g.params, x, gt, masks)
        _compare(_run(model, x, gt), o32, o64)


def test_mlp_generic_kernels_dropout_vs_oracle():
    """A shape the warp variant does not serve (H = 72 > 64): the generic CTA-per-tile kernels."""
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    c = dict(Golden("mlp_k2").cfg, num_blocks=2, hidden_dim=72, channels_mlp_dim=40, regularization=0.2)
    torch.manual_seed(3)
    params = {k: v.detach().cpu().numpy() for k, v in MlpMixer(**c).state_dict().items()}
    x, gt = synthetic_pose_windows(130, 10, 10, 66, scale="amass", seed=9)
    torch.manual_seed(SEED)
    model = _mlp(c, params)
    masks = MK.mlp_generic_masks(c, len(x), SEED, step=0)
    o32, o64 = _oracles(O.MlpMixerOracle, c, params, x, gt, masks)
    _compare(_run(model, x, gt), o32, o64)


@pytest.mark.parametrize("case", ["conv_k1", "conv_k3"])
def test_conv_kernels_dropout_vs_oracle(case):
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    g = Golden(case)
    c = dict(g.cfg, regularization=0.1)
    torch.manual_seed(SEED)
    m = ConvMixer(**c)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.params.items()}, strict=True)
    model = m.cuda().train()
    masks = MK.conv_masks(c, len(g.x), SEED, step=0)
    o32, o64 = _oracles(O.ConvMixerOracle, c, g.params, g.x, g.gt, masks)
    skip = [k for k in o32[2] if ".se2." in k]
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    check_close("pred", pred, o32[0], o64[0], rtol=TOL)
    assert abs(loss - o64[1]) <= TOL * abs(o64[1])
    floor = 5e-6 * grad_scale(o32[2])
    for k in O.trainable_keys(g.params):
        if k not in skip:
            check_close("grad " + k, grads[k], o32[2][k], o64[2][k], rtol=TOL, atol=floor * (50 if k == "encoder.channelUpscaling.bias" else 1))
    check_close("dx", dx, o32[3], o64[3], rtol=TOL, atol=1e-6 * float(np.abs(o32[3]).max()))
