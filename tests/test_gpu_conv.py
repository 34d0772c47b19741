"""GPU (-m gpu): ConvMixer CUDA path through the product modules / C ABI vs the golden fixtures
(generated from the reference) and vs the numpy oracle on larger seeded batches.

Tolerance: fp32, 1e-5 relative per tensor (north-star), noise-aware (tests.golden_util.check_close).
"""
import numpy as np
import pytest
import torch

from oracle import mixer_np as O
from tests.golden_util import Golden, check_close, golden_cases, grad_scale
from tests.synthetic import synthetic_pose_windows

pytestmark = pytest.mark.gpu
TOL = 1e-5
# mathematically ZERO gradients (a per-channel constant in front of LayerNorms only): what any implementation returns is the
# rounding noise of a long cancelling sum, which scales with the summands, not with the other gradients
ZERO_GRAD_SLACK = {"encoder.channelUpscaling.bias": 50}
CONV_CASES = [c for c in golden_cases("conv") if not c.endswith("_bn")]


def _model(cfg, params):
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    m = ConvMixer(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    return m.cuda()


def _run(model, x, gt):
    from motionmixerconv_b200.functional import mpjpe_error
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    pred = model(xg)
    loss = mpjpe_error(pred, torch.from_numpy(gt).cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    return pred.detach().cpu().numpy(), float(loss.detach()), grads, xg.grad.cpu().numpy()


@pytest.mark.parametrize("case", CONV_CASES)
def test_golden(case):
    g = Golden(case)
    model = _model(g.cfg, g.params).train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    o64 = O.ConvMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(g.x)
    _, dp64 = O.mpjpe(p64, g.gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    check_close("pred", pred, g.pred, p64, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 5e-6 * grad_scale(g.grads)   # exactly-cancelling gradients (a bias in front of a LayerNorm) are pure rounding noise
    for k, want in g.grads.items():
        if ".se2." in k:
            continue
        check_close("grad " + k, grads[k], want, g64[k], rtol=TOL, atol=floor * ZERO_GRAD_SLACK.get(k, 1))
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))
    model.eval()
    with torch.no_grad():
        pe = model(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval", pe, g.pred_eval, rtol=TOL)


@pytest.mark.parametrize("case,B", [("conv_k1", 333), ("conv_k3", 37), ("conv_evenk", 65), ("conv_once_se", 70), ("conv_harm64", 130),
                                    ("conv_k1", 4096)])      # the last: several tiles per CTA of every kernel at a bench-sized batch
def test_vs_oracle_ragged_batch(case, B):
    """Batch sizes that are not multiples of the CTA tile: several tiles per CTA + a partial last tile."""
    g = Golden(case)
    c = g.cfg
    x, gt = synthetic_pose_windows(B, c["in_nTP"], c["out_nTP"], c["dimPosIn"], scale="amass", seed=7)
    model = _model(c, g.params).train()
    pred, loss, grads, dx = _run(model, x, gt)
    res = {}
    for dt in (np.float32, np.float64):
        o = O.ConvMixerOracle(c, g.params, dtype=dt)
        p = o.forward(x)
        l, dp = O.mpjpe(p, gt.astype(dt))
        gr, dxx = o.backward(dp)
        res[dt] = (p, l, gr, dxx)
    p32, l32, g32, dx32 = res[np.float32]
    p64, l64, g64, dx64 = res[np.float64]
    check_close("pred", pred, p32, p64, rtol=TOL)
    assert abs(loss - float(l64)) <= TOL * abs(float(l64))
    floor = 5e-6 * grad_scale(g32)      # exactly-cancelling gradients (e.g. a bias in front of a LayerNorm) are pure rounding noise
    for k in O.trainable_keys(g.params):
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor * ZERO_GRAD_SLACK.get(k, 1))
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))


def test_state_dict_roundtrip_and_errors():
    g = Golden("conv_k1")
    model = _model(g.cfg, g.params)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g.params.keys())
    # se2.* are aliases of se.* (same storage), as in the reference (conv_mixer_model.py:257)
    assert sd["Mixer_Block.0.se2.excitationBlock.0.weight"].data_ptr() == sd["Mixer_Block.0.se.excitationBlock.0.weight"].data_ptr()
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 10, 66))            # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 9, 66).cuda())      # wrong in_nTP
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    with pytest.raises(ValueError):
        ConvMixer(1, 6, 8, 6, 10, 5, activation="relu")
    with pytest.raises(ValueError):
        ConvMixer(1, 6, 8, 6, 10, 5, mode_conv="thrice")


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainstep_matches_golden_three_adam_steps(use_graph):
    from motionmixerconv_b200.train import TrainStep
    g = Golden("conv_k1")
    model = _model(g.cfg, g.params).train()
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, use_cuda_graph=use_graph)
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    losses = [float(ts.step(x, gt)) for _ in range(3)]
    np.testing.assert_allclose(losses, g.losses, rtol=2e-5)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g.params.keys())
    bad = tot = 0
    for k in O.trainable_keys(g.params):
        upd = sd[k].cpu().numpy() - g.params[k]
        want = g.params3[k] - g.params[k]
        bad += int((np.abs(upd - want) > 3e-3 * 1e-2).sum())
        tot += upd.size
    assert bad / tot <= 0.005, (bad, tot)


def test_dropout_training_mode():
    from motionmixerconv_b200.functional import mpjpe_error
    g = Golden("conv_k1")
    cfg = dict(g.cfg, regularization=0.1)
    model = _model(cfg, g.params).train()
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    a, b = model(x), model(x)
    assert (a - b).abs().max().item() > 0            # fresh masks per call
    mpjpe_error(a, gt).backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    with torch.no_grad():
        e = model(x).cpu().numpy()
    np.testing.assert_allclose(e, g.pred_eval, rtol=0, atol=1e-5 * np.abs(g.pred_eval).max())


# ---- BatchNorm2d halves (regularization == -1, the Optuna "production" setting) ----------------
def test_batchnorm_golden_train_and_eval():
    g = Golden("conv_k3_bn")
    model = _model(g.cfg, g.params).train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    o64 = O.ConvMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(g.x)
    _, dp64 = O.mpjpe(p64, g.gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    check_close("pred", pred, g.pred, p64, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 5e-6 * grad_scale(g.grads)   # exactly-cancelling gradients (a bias in front of a LayerNorm) are pure rounding noise
    for k, want in g.grads.items():
        if ".se2." in k:
            continue
        check_close("grad " + k, grads[k], want, g64[k], rtol=TOL, atol=floor * ZERO_GRAD_SLACK.get(k, 1))
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))
    sd = model.state_dict()
    for k in sd:                                    # running statistics after ONE training forward == the reference's
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), o64.p[k], rtol=1e-5, atol=1e-7, err_msg=k)
        if "num_batches_tracked" in k:
            assert int(sd[k]) == int(g.params[k]) + 1
    # eval mode uses the running statistics (folded into a per-channel affine inside the fused kernel)
    ref = O.ConvMixerOracle(g.cfg, {k: v.cpu().numpy() for k, v in sd.items()}, dtype=np.float64)
    pe64 = ref.forward(g.x, training=False)
    model.eval()
    with torch.no_grad():
        pe = model(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval", pe, pe64.astype(np.float32), pe64, rtol=TOL)


@pytest.mark.parametrize("use_graph", [False, True])
def test_batchnorm_trainstep_three_adam_steps(use_graph):
    from motionmixerconv_b200.train import TrainStep
    g = Golden("conv_k3_bn")
    model = _model(g.cfg, g.params).train()
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, use_cuda_graph=use_graph)
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    losses = [float(ts.step(x, gt)) for _ in range(3)]
    np.testing.assert_allclose(losses, g.losses, rtol=5e-5)
    sd = model.state_dict()
    assert int(sd["Mixer_Block.0.conv1.reg.num_batches_tracked"]) == 3
    for k in g.params3:
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), g.params3[k], rtol=2e-4, atol=1e-6, err_msg=k)
