"""CPU: the MotionMixer kernel SOURCE (csrc/mmx_mlp.cuh, mmx_loss_adam.cuh) run phase by phase in the host
emulator (tests/emu/harness.py — test infrastructure) against the golden fixtures generated from the reference."""
import numpy as np
import pytest

from oracle import mixer_np as O
from tests.emu import harness as H
from tests.golden_util import Golden, check_close, golden_cases, grad_scale

TOL = 1e-5
MLP_CASES = [c for c in golden_cases("mlp") if not c.startswith("mlp_bn")]


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


WARP_CASES = [  # (hidden, channels, activation, use_se, r_se, B): all route to the warp-per-sequence-pair kernels (T=10, tok=20)
    (50, 50, "mish", True, 8, 7),
    (36, 64, "gelu", True, 4, 150),      # H != ch, SE hidden width 2, several groups per warp
    (64, 24, "gelu", False, 4, 33),      # no SE, H % 4 == 0 (vector loads), odd batch (dead half-warp)
    (10, 6, "mish", True, 8, 4),         # tiny widths: most lanes idle
]


@pytest.mark.parametrize("H_,ch,act,use_se,r_se,B", WARP_CASES)
def test_emulated_warp_variant_vs_oracle(H_, ch, act, use_se, r_se, B, monkeypatch):
    from oracle import mixer_torch as MT
    from tests.synthetic import synthetic_pose_windows
    cfg = dict(num_classes=12, num_blocks=2, hidden_dim=H_, tokens_mlp_dim=20, channels_mlp_dim=ch, seq_len=10, pred_len=7,
               activation=act, regularization=0, input_size=12, r_se=r_se, use_se=use_se)
    params = {k: v.numpy() for k, v in MT.random_params("mlp", cfg, 5).items()}
    x, gt = synthetic_pose_windows(B, 10, 7, 12, scale="amass", seed=9)
    o32, o64 = O.MlpMixerOracle(cfg, params), O.MlpMixerOracle(cfg, params, dtype=np.float64)
    p32, p64 = o32.forward(x), o64.forward(x)
    _, dp32 = O.mpjpe(p32, gt)
    _, dp64 = O.mpjpe(p64, gt.astype(np.float64))
    (g32, dx32), (g64, dx64) = o32.backward(dp32), o64.backward(dp64)
    res = {}
    for v1 in ("0", "1"):                      # warp variant, and the generic kernels on the same problem
        monkeypatch.setenv("MMX_MLP_V1", v1)
        m = H.EmuMlpMixer(cfg, params, training=True)
        pred = m.forward(x)
        check_close("pred", pred, p32, p64, rtol=TOL)
        _, dpred = H.mpjpe(pred, gt)
        grads, dx = m.backward(dpred)
        floor = 1e-6 * grad_scale(g32)
        for k in O.trainable_keys(params):
            check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor)
        check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))
        res[v1] = pred
    assert np.abs(res["0"] - res["1"]).max() <= 1e-5 * np.abs(p32).max()


@pytest.mark.parametrize("v1", ["0", "1"])
def test_emulated_dropout_masks_consistent_between_forward_and_backward(v1, monkeypatch):
    """regularization > 0: the backward regenerates the forward's Philox masks.  With a fixed (seed, step) the block
    is a deterministic function, so a finite difference of sum(y*r) along a random direction must match <dx, v> and
    <dW, dV> from the backward — a mask mismatch between the two kernels shows up as an O(1) error."""
    import ctypes as C
    from motionmixerconv_b200 import _lib as L
    monkeypatch.setenv("MMX_MLP_V1", v1)
    rng = np.random.default_rng(3)
    B, T, Hd, tok, ch = 5, 10, 50, 20, 50
    shapes = [(Hd,), (Hd,), (tok, T), (tok,), (T, tok), (T,), (Hd,), (Hd,), (ch, Hd), (ch,), (Hd, ch), (Hd,), (1, T), (T, 1)]
    params = [H.f32(rng.standard_normal(s) * 0.3) for s in shapes]
    params[0] += 1.0
    params[6] += 1.0
    x = H.f32(rng.standard_normal((B, T, Hd)))
    r = H.f32(rng.standard_normal((B, T, Hd)))
    desc = L.MmxMlpBlockDesc(B, T, Hd, tok, ch, 1, L.MMX_ACT["mish"], 1, 0, 1, 2, L.MmxDropout(0.3, 1234, 7, None))
    fields = ("ln1_w", "ln1_b", "tok_w1", "tok_b1", "tok_w2", "tok_b2", "ln2_w", "ln2_b", "ch_w1", "ch_b1", "ch_w2", "ch_b2", "se_w1", "se_w2")

    def table(arrs):
        t = L.MmxMlpBlockParams()
        for f, a in zip(fields, arrs):
            setattr(t, f, H.ptr(a))
        return t

    def fwd(xx, pp):
        y = np.empty_like(xx)
        H.call("mmx_mlp_block_fwd", C.byref(desc), C.byref(table(pp)), H.ptr(xx), H.ptr(y), None)
        return y

    y0 = fwd(x, params)
    assert np.array_equal(y0, fwd(x, params))                       # same masks on every call with the same (seed, step)
    nodrop = L.MmxMlpBlockDesc(B, T, Hd, tok, ch, 1, L.MMX_ACT["mish"], 1, 0, 0, 2, L.MmxDropout(0.3, 1234, 7, None))
    y_eval = np.empty_like(x)
    H.call("mmx_mlp_block_fwd", C.byref(nodrop), C.byref(table(params)), H.ptr(x), H.ptr(y_eval), None)
    assert np.abs(y0 - y_eval).max() > 1e-3                         # dropout did something
    grads = [np.zeros_like(p) for p in params]
    dx = np.empty_like(x)
    H.call("mmx_mlp_block_bwd", C.byref(desc), C.byref(table(params)), C.byref(table(grads)), H.ptr(x), H.ptr(r), H.ptr(dx), None)
    v = H.f32(rng.standard_normal(x.shape))
    eps = 1e-2
    fd = (np.sum(fwd(x + eps * v, params).astype(np.float64) * r) - np.sum(fwd(x - eps * v, params).astype(np.float64) * r)) / (2 * eps)
    an = float(np.sum(dx.astype(np.float64) * v))
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1.0), (fd, an)
    for idx in (2, 4, 8, 10, 3):                                    # tok_w1, tok_w2, ch_w1, ch_w2, tok_b1
        dv = H.f32(rng.standard_normal(params[idx].shape))
        pp, pm = list(params), list(params)
        pp[idx] = H.f32(params[idx] + eps * dv)
        pm[idx] = H.f32(params[idx] - eps * dv)
        fd = (np.sum(fwd(x, pp).astype(np.float64) * r) - np.sum(fwd(x, pm).astype(np.float64) * r)) / (2 * eps)
        an = float(np.sum(grads[idx].astype(np.float64) * dv))
        assert abs(fd - an) <= 3e-2 * max(abs(an), 1.0), (fields[idx], fd, an)
