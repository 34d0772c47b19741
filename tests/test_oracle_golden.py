"""CPU: the numpy oracle (oracle/mixer_np.py) against the fixtures generated from the reference.

Tolerance: fp32, 1e-5 relative (max-abs error over max-abs reference, per tensor) — the north-star's
fp32 criterion — in the noise-aware form of ``tests.golden_util.check_close``.
"""
import numpy as np
import pytest

from oracle import mixer_np as O
from tests.golden_util import Golden, check_close, golden_cases, grad_scale, rel_err

TOL = 1e-5


def make_oracle(g, dtype=np.float32):
    cls = O.MlpMixerOracle if g.family == "mlp" else O.ConvMixerOracle
    return cls(g.cfg, g.params, dtype=dtype)


def run(g, dtype):
    o = make_oracle(g, dtype)
    pred = o.forward(g.x, training=True)
    loss, dpred = O.mpjpe(pred, g.gt.astype(dtype))
    grads, dx = o.backward(dpred)
    return o, pred, loss, grads, dx


@pytest.mark.parametrize("case", golden_cases())
def test_forward_backward(case):
    g = Golden(case)
    o, pred, loss, grads, dx = run(g, np.float32)
    _, pred64, loss64, grads64, dx64 = run(g, np.float64)
    check_close("pred", pred, g.pred, pred64, rtol=TOL)
    assert abs(float(loss) - g.loss) <= TOL * abs(g.loss)
    assert set(grads) == set(g.grads)
    floor = 1e-6 * grad_scale(g.grads)
    for k, want in g.grads.items():
        check_close("grad " + k, grads[k], want, grads64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))
    # BN running statistics after one training forward
    for k, want in g.params1.items():
        if "num_batches" in k:
            assert int(o.p[k]) == int(want)
        else:
            check_close("buffer " + k, o.p[k], want, rtol=TOL)


@pytest.mark.parametrize("case", golden_cases())
def test_eval_forward(case):
    g = Golden(case)
    o = make_oracle(g)
    check_close("pred_eval", o.forward(g.x, training=False), g.pred_eval, rtol=TOL)


@pytest.mark.parametrize("case", ["mlp_k2", "mlp_bn", "conv_k1", "conv_k3_bn", "conv_once_se"])
def test_three_adam_steps(case):
    g = Golden(case)
    o = make_oracle(g)
    losses = O.train_steps(o, g.x, g.gt, 3)
    np.testing.assert_allclose(losses, g.losses, rtol=2e-5)
    # Adam's first steps move every weight by ~lr whatever the gradient size, which turns the
    # rounding noise of gradients whose true value is 0 (e.g. encoder.channelUpscaling.bias: the
    # LayerNorms make the output invariant to it) into +-lr.  So compare the UPDATE p3-p0 against
    # 3*lr element-wise and bound the fraction of disagreeing elements over the whole model.
    bad = tot = 0
    for k in O.trainable_keys(g.params):
        upd = o.p[k] - g.params[k]
        want = g.params3[k] - g.params[k]
        bad += int((np.abs(upd - want) > 3e-3 * 1e-2).sum())
        tot += upd.size
    assert bad / tot <= 0.005, (bad, tot)


def test_fp64_noise_floor():
    """fp32 oracle vs fp64 oracle: documents the head-room under the 1e-5 criterion."""
    g = Golden("mlp_k2")
    p32 = make_oracle(g, np.float32).forward(g.x)
    p64 = make_oracle(g, np.float64).forward(g.x)
    assert rel_err(p32, p64) < 5e-6


# ---------------------------------------------------------------------------------------------
# the torch-CPU functional port (oracle/mixer_torch.py) that bench.py times as the CPU baseline
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", golden_cases())
def test_torch_port_matches_reference(case):
    import torch
    from oracle import mixer_torch as MT
    g = Golden(case)
    p = {k: torch.from_numpy(np.array(v)) for k, v in g.params.items()}
    for k in p:
        if MT.is_trainable(k) and p[k].dtype.is_floating_point:
            p[k].requires_grad_(True)
    fwd = MT.mlpmixer_forward if g.family == "mlp" else MT.convmixer_forward
    x = torch.from_numpy(g.x).requires_grad_(True)
    pred = fwd(p, g.cfg, x, training=True)
    loss = MT.mpjpe_error(pred, torch.from_numpy(g.gt))
    loss.backward()
    check_close("pred", pred.detach().numpy(), g.pred, rtol=1e-6)
    assert abs(float(loss) - g.loss) <= 1e-6 * abs(g.loss)
    floor = 1e-6 * grad_scale(g.grads)
    _, _, _, grads64, _ = run(g, np.float64)     # summation order differs with the thread count: noise-aware
    for k, want in g.grads.items():
        check_close("grad " + k, p[k].grad.numpy(), want, grads64[k], rtol=1e-5, atol=floor)
    # the layout produced without the reference (bench.py on the GPU box) has the same keys/shapes
    rp = MT.random_params(g.family, g.cfg)
    assert list(rp.keys()) == list(g.params.keys())
    for k in rp:
        assert tuple(rp[k].shape) == tuple(g.params[k].shape), k
