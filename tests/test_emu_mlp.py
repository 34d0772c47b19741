"""CPU: the MotionMixer kernel SOURCE (csrc/mmx_mlp.cuh, mmx_loss_adam.cuh) run phase by phase in the host
emulator (tests/emu/harness.py — test infrastructure) against the golden fixtures generated from the reference."""
import numpy as np
import pytest

from oracle import mixer_np as O
from tests.emu import harness as H
from tests.golden_util import Golden, check_close, golden_cases, grad_scale

TOL = 1e-5
MLP_CASES = [c for c in golden_cases("mlp") if c != "mlp_bn"]


@pytest.mark.parametrize("case", MLP_CASES)
def test_emulated_kernels_match_golden(case):
    g = Golden(case)
    n = min(g.x.shape[0], 12)
    x, gt = g.x[:n], g.gt[:n]
    o32 = O.MlpMixerOracle(g.cfg, g.params, dtype=np.float32)
    o64 = O.MlpMixerOracle(g.cfg, g.params, dtype=np.float64)
    p32, p64 = o32.forward(x), o64.forward(x)
    l32, dp32 = O.mpjpe(p32, gt)
    l64, dp64 = O.mpjpe(p64, gt.astype(np.float64))
    (g32, dx32), (g64, dx64) = o32.backward(dp32), o64.backward(dp64)
    m = H.EmuMlpMixer(g.cfg, g.params, training=True)
    pred = m.forward(x)
    check_close("pred", pred, p32, p64, rtol=TOL)
    loss, dpred = H.mpjpe(pred, gt)
    assert abs(loss - float(l64)) <= TOL * abs(float(l64))
    grads, dx = m.backward(dpred)
    floor = 1e-6 * grad_scale(g32)
    for k in O.trainable_keys(g.params):
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))


def test_emulated_adam_matches_oracle():
    rng = np.random.default_rng(0)
    n = 1003
    p = rng.standard_normal(n).astype(np.float32)
    g = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    p2, m2, v2 = p.copy(), m.copy(), v.copy()
    for step in (1, 2, 3):
        H.adam_step(p, g, m, v, step)
        O.adam_step(p2, g, m2, v2, step, 1e-3)
    np.testing.assert_allclose(p, p2, rtol=1e-6, atol=1e-6)   # a few ulp: different association of the same formula
    np.testing.assert_allclose(v, v2, rtol=1e-6, atol=1e-12)
