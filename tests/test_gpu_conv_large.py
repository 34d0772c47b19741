"""GPU (-m gpu): ConvMixerBlock halves on the stage-kernel chain (csrc/mmx_api_conv_large.cu, functional.ConvHalfLarge) -- the
shapes the fused kernels do not hold in shared memory (the reference's Optuna grid at C = 8, E = 192: kernels 5x9 ... 9x29,
optuna_search/conv_optuna_main.py:339-342) and BatchNorm with the max squeeze -- against the numpy oracle (fp32 + fp64) and,
forced onto small shapes, against the golden fixtures generated from the reference.  Tolerance 1e-5 (fp32 mode)."""
import numpy as np
import pytest
import torch

from oracle import mixer_np as O
from tests import masks_np as MK
from tests.golden_util import Golden, check_close, grad_scale
from tests.synthetic import synthetic_pose_windows

pytestmark = pytest.mark.gpu
TOL = 1e-5
SEED = 4321


def _build(cfg, params=None, seed=3, jitter_bn=False):
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    torch.manual_seed(seed)
    m = ConvMixer(**cfg)
    if params is not None:
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    if jitter_bn:
        with torch.no_grad():
            for k, p in m.named_parameters():
                if ".reg." in k:
                    p.add_(0.3 * torch.randn_like(p))
    sd = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    return m.cuda(), sd


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


def _oracles(cfg, params, x, gt, masks=None):
    out = []
    for dt in (np.float32, np.float64):
        o = O.ConvMixerOracle(cfg, params, dtype=dt)
        p = o.forward(x, training=True, masks=masks)
        l, dp = O.mpjpe(p, gt.astype(dt))
        g, dx = o.backward(dp)
        out.append((p, float(l), g, dx, o))
    return out


def _compare(got, o32, o64, params):
    pred, loss, grads, dx = got
    check_close("pred", pred, o32[0], o64[0], rtol=TOL)
    assert abs(loss - o64[1]) <= TOL * abs(o64[1])
    floor = 5e-6 * grad_scale(o32[2])
    for k in O.trainable_keys(params):
        if ".se2." in k:
            continue
        check_close("grad " + k, grads[k], o32[2][k], o64[2][k], rtol=TOL, atol=floor * (50 if k == "encoder.channelUpscaling.bias" else 1))
    check_close("dx", dx, o32[3], o64[3], rtol=TOL, atol=1e-6 * float(np.abs(o32[3]).max()))


BASE = dict(num_blocks=2, dimPosIn=33, dimPosEmb=192, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=8, activation="mish",
            use_se=True, r_se=8, encoder_n_harmonic_functions=0, encoder_omega0=0)
LARGE = {
    # (config, batch): the Optuna grid's shapes that return MMX_E_UNSUPPORTED from the fused kernels
    "c8_e192_k5x9_twice": (dict(BASE, conv1_kernel_shape=(5, 9), mode_conv="twice", regularization=0), 5),
    "c8_e192_k9x29_once_bn": (dict(BASE, conv1_kernel_shape=(9, 29), mode_conv="once", regularization=-1.0), 6),
    "c8_e192_k1x29_twice_gelu": (dict(BASE, conv1_kernel_shape=(1, 29), mode_conv="twice", regularization=0, activation="gelu"), 3),
    "c4_e64_k5x5_bn_maxpool": (dict(BASE, dimPosEmb=64, conv_nChan=4, conv1_kernel_shape=(5, 5), mode_conv="twice", regularization=-1.0,
                                   use_max_pooling=True, r_se=4), 7),
    "c3_e50_k3x4_even_kernel_maxpool": (dict(BASE, dimPosEmb=50, conv_nChan=3, conv1_kernel_shape=(3, 4), mode_conv="once",
                                            regularization=0, use_max_pooling=True), 9),
}


@pytest.mark.parametrize("name", sorted(LARGE))
def test_large_halves_vs_oracle(name, monkeypatch):
    cfg, B = LARGE[name]
    if cfg["dimPosEmb"] < 192:
        monkeypatch.setenv("MMX_CONV_FORCE_LARGE", "1")
    model, params = _build(cfg, jitter_bn=cfg["regularization"] == -1.0)
    model.train()
    halves = (0, 1) if cfg["mode_conv"] == "twice" else (0,)
    assert all(any(mb.uses_large_path(h, B) for h in halves) for mb in model.Mixer_Block)
    x, gt = synthetic_pose_windows(B, 10, 10, 33, scale="ais", seed=11)
    got = _run(model, x, gt)
    o32, o64 = _oracles(cfg, params, x, gt)
    _compare(got, o32, o64, params)
    if cfg["regularization"] == -1.0:
        sd = model.state_dict()
        for k in sd:
            if "running_" in k:
                np.testing.assert_allclose(sd[k].cpu().numpy(), o64[4].p[k], rtol=1e-5, atol=1e-7, err_msg=k)
        ref = O.ConvMixerOracle(cfg, {k: v.cpu().numpy() for k, v in sd.items()}, dtype=np.float64)
        model.eval()
        with torch.no_grad():
            pe = model(torch.from_numpy(x).cuda()).cpu().numpy()
        pe64 = ref.forward(x, training=False)
        check_close("pred_eval", pe, pe64.astype(np.float32), pe64, rtol=TOL)


def test_large_half_dropout_vs_oracle_with_the_same_masks():
    cfg = dict(BASE, conv1_kernel_shape=(5, 9), mode_conv="twice", regularization=0.1, num_blocks=1)
    torch.manual_seed(SEED)
    model, params = _build(cfg, seed=SEED)
    model.train()
    x, gt = synthetic_pose_windows(4, 10, 10, 33, scale="ais", seed=5)
    masks = MK.conv_masks(cfg, len(x), SEED, step=0)
    got = _run(model, x, gt)
    o32, o64 = _oracles(cfg, params, x, gt, masks)
    _compare(got, o32, o64, params)


@pytest.mark.parametrize("case", ["conv_k1", "conv_once_se", "conv_evenk", "conv_k3_bn"])
def test_golden_fixtures_through_the_stage_kernel_chain(case, monkeypatch):
    """The fixtures generated from the reference, with every half forced onto the chain (C = 1 ... 4, harmonic encoder,
    'once' quirks, even kernels with asymmetric padding, BatchNorm)."""
    monkeypatch.setenv("MMX_CONV_FORCE_LARGE", "1")
    g = Golden(case)
    model, _ = _build(g.cfg, g.params)
    model.train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    o64 = O.ConvMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(g.x)
    _, dp64 = O.mpjpe(p64, g.gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    check_close("pred", pred, g.pred, p64, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 5e-6 * grad_scale(g.grads)
    for k, want in g.grads.items():
        if ".se2." in k:
            continue
        check_close("grad " + k, grads[k], want, g64[k], rtol=TOL, atol=floor * (50 if k == "encoder.channelUpscaling.bias" else 1))
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainstep_on_a_large_batchnorm_config(use_graph):
    """TrainStep (flat buffers, CUDA graph) with halves on the chain: three Adam steps against the numpy oracle's."""
    from motionmixerconv_b200.train import TrainStep
    cfg, B = LARGE["c8_e192_k9x29_once_bn"]
    model, params = _build(cfg)
    model.train()
    x, gt = synthetic_pose_windows(B, 10, 10, 33, scale="ais", seed=11)
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, use_cuda_graph=use_graph)
    xs, gts = torch.from_numpy(x).cuda(), torch.from_numpy(gt).cuda()
    losses = [float(ts.step(xs, gts)) for _ in range(3)]
    want = O.train_steps(O.ConvMixerOracle(cfg, params), x, gt, 3)
    np.testing.assert_allclose(losses, want, rtol=5e-5)


def test_batchnorm_with_max_squeeze_golden():
    """BatchNorm2d + use_max_pooling (conv_mixer_model.py:60-62 with :115-116) against the fixture generated from the reference:
    training forward / backward, running statistics after one update, eval forward."""
    g = Golden("conv_maxpool_bn")
    model, _ = _build(g.cfg, g.params)
    model.train()
    assert all(mb.uses_large_path(0, len(g.x)) for mb in model.Mixer_Block)
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    o64 = O.ConvMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(g.x)
    _, dp64 = O.mpjpe(p64, g.gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    check_close("pred", pred, g.pred, p64, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 5e-6 * grad_scale(g.grads)
    for k, want in g.grads.items():
        if ".se2." in k:
            continue
        check_close("grad " + k, grads[k], want, g64[k], rtol=TOL, atol=floor * (50 if k == "encoder.channelUpscaling.bias" else 1))
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))
    sd = model.state_dict()
    for k in sd:
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), g.params1[k], rtol=1e-5, atol=1e-7, err_msg=k)
    fresh, _ = _build(g.cfg, g.params)
    fresh.eval()
    with torch.no_grad():
        pe = fresh(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval", pe, g.pred_eval, rtol=TOL)
