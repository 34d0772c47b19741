"""CPU: the ConvMixer kernel SOURCE (csrc/mmx_conv*.cuh) run phase by phase in the host emulator
(tests/emu/harness.py — test infrastructure) against the golden fixtures generated from the reference.

This is what keeps indexing / math of the CUDA kernels checkable in a container without a GPU; the GPU
parity tests proper are tests/test_gpu_conv.py (-m gpu).
"""
import numpy as np
import pytest

from oracle import mixer_np as O
from tests.emu import harness as H
from tests.golden_util import Golden, check_close, golden_cases, grad_scale

TOL = 1e-5
# mathematically ZERO gradients (a per-channel constant in front of LayerNorms only): what any implementation returns is the
# rounding noise of a long cancelling sum, which scales with the summands, not with the other gradients
ZERO_GRAD_SLACK = {"encoder.channelUpscaling.bias": 50}
# includes conv_k3_bn: BatchNorm2d halves (two-pass kernels); the fixtures served by the stage-kernel chain (not in the emulator: plain
# CUDA kernels) are excluded -- a tile beyond the fused kernels, BatchNorm with the max squeeze
CONV_CASES = [c for c in golden_cases("conv") if c not in ("conv_c8_k5x9", "conv_maxpool_bn")]


@pytest.mark.parametrize("case", CONV_CASES)
def test_emulated_kernels_match_golden(case):
    g = Golden(case)
    n = min(g.x.shape[0], 6)                       # the emulator is slow: a few sequences are enough for indexing
    x, gt = g.x[:n], g.gt[:n]
    o64 = O.ConvMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(x)
    l64, dp64 = O.mpjpe(p64, gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    o32 = O.ConvMixerOracle(g.cfg, g.params, dtype=np.float32)
    p32 = o32.forward(x)
    l32, dp32 = O.mpjpe(p32, gt)
    g32, dx32 = o32.backward(dp32)

    m = H.EmuConvMixer(g.cfg, g.params, training=True)
    pred = m.forward(x)
    check_close("pred", pred, p32, p64, rtol=TOL)
    loss, dpred = H.mpjpe(pred, gt)
    assert abs(loss - float(l64)) <= TOL * abs(float(l64))
    grads, dx = m.backward(dpred)
    if m.bn:      # running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance); eval mode uses them
        for k, v in m.running.items():
            if "num_batches" not in k:
                np.testing.assert_allclose(v, o64.p[k], rtol=1e-5, atol=1e-7)
        me = H.EmuConvMixer(g.cfg, {**g.params, **{k: o64.p[k] for k in m.running}}, training=False)
        pe = O.ConvMixerOracle(g.cfg, {**g.params, **{k: o64.p[k] for k in m.running}}, dtype=np.float64).forward(x, training=False)
        check_close("pred_eval", me.forward(x), pe.astype(np.float32), pe, rtol=TOL)
    floor = 5e-6 * grad_scale(g32)      # exactly-cancelling gradients (e.g. a bias in front of a LayerNorm) are pure rounding noise
    for k in O.trainable_keys(g.params):
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor * ZERO_GRAD_SLACK.get(k, 1))
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))


@pytest.mark.parametrize("case,B,S,xg", [("conv_k1", 7, 2, 0), ("conv_k3", 3, 1, 1), ("conv_once_se", 5, 2, 1), ("conv_evenk", 5, 3, 0)])
def test_emulated_multi_tile_and_streamed_inputs(case, B, S, xg, monkeypatch):
    """Forced small tiles: several tiles per CTA with a ragged last one (the persistent loops, the accumulators
    that live across tiles) — once with the backward's X / dY tiles in shared memory, once streamed from global."""
    from tests.synthetic import synthetic_pose_windows
    g = Golden(case)
    c = g.cfg
    x, gt = synthetic_pose_windows(B, c["in_nTP"], c["out_nTP"], c["dimPosIn"], scale="amass", seed=3)
    o32 = O.ConvMixerOracle(c, g.params, dtype=np.float32)
    o64 = O.ConvMixerOracle(c, g.params, dtype=np.float64)
    p32, p64 = o32.forward(x), o64.forward(x)
    _, dp32 = O.mpjpe(p32, gt)
    _, dp64 = O.mpjpe(p64, gt.astype(np.float64))
    (g32, dx32), (g64, dx64) = o32.backward(dp32), o64.backward(dp64)
    for name in ("MMX_CONV_S_FWD", "MMX_CONV_S_BWD", "MMX_CHEAD_S_FWD", "MMX_CHEAD_S_BWD"):
        monkeypatch.setenv(name, str(S))
    monkeypatch.setenv("MMX_CONV_X_GLOBAL", str(xg))
    m = H.EmuConvMixer(c, g.params, training=True)
    pred = m.forward(x)
    check_close("pred", pred, p32, p64, rtol=TOL)
    _, dpred = H.mpjpe(pred, gt)
    grads, dx = m.backward(dpred)
    floor = 5e-6 * grad_scale(g32)      # exactly-cancelling gradients (e.g. a bias in front of a LayerNorm) are pure rounding noise
    for k in O.trainable_keys(g.params):
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor * ZERO_GRAD_SLACK.get(k, 1))
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))


def test_emulated_harmonic_embedding_exact_angle_doubling():
    """The harmonic encoder evaluates sin/cos(fl32(x*f_h)), f_h = omega0*2^h up to 9e17, from ONE 128-bit phase per input
    value and bit shifts (mmx_conv_io.cuh).  With an identity embed_mlp the kernel returns the embedding itself: it must
    match sin/cos of the exact fp32 argument (positional_encoder.py:86-89) for every harmonic, and the generic path
    (frequency table that is not a power-of-two ladder) must agree."""
    import ctypes as C
    from motionmixerconv_b200 import _lib as L
    rng = np.random.default_rng(0)
    Hn, D, T = 64, 1, 8
    xs = np.concatenate([rng.standard_normal(40) * 0.4, [0.0, 1e-30, -3.7e-12, 7.25, -1234.5, 2.0 ** -20, 0.5, -0.5]]).astype(np.float32)
    B = len(xs) // T
    x = H.f32(xs[:B * T].reshape(B, T, D))
    K = 2 * Hn * D
    for ladder in (True, False):
        freq = (np.float32(0.1) * (2.0 ** np.arange(Hn))).astype(np.float32)
        if not ladder:
            freq = (freq * np.float32(1.0 + 2.0 ** -20)).astype(np.float32)
            freq[5] = np.float32(3.3)                              # breaks the ladder: generic sinf/cosf path
        w, b = H.f32(np.eye(K)), H.f32(np.zeros(K))
        wc, bc = H.f32(np.ones((1, 1))), H.f32(np.zeros(1))
        tab = L.MmxEncoderParams()
        tab.freq, tab.w, tab.b, tab.wc, tab.bc = H.ptr(freq), H.ptr(w), H.ptr(b), H.ptr(wc), H.ptr(bc)
        m = np.empty((B * T, K), np.float32)
        y = np.empty((B, 1, T, K), np.float32)
        H.call("mmx_pose_encoder_fwd", C.byref(L.MmxEncoderDesc(B, T, D, K, 1, Hn)), C.byref(tab), H.ptr(x), H.ptr(m), H.ptr(y), None)
        arg = (x.reshape(-1, 1) * freq[None, :]).astype(np.float32).astype(np.float64)   # ONE fp32 multiply, then exact
        want = np.concatenate([np.sin(arg), np.cos(arg)], axis=1)
        assert np.abs(m - want).max() <= 4e-7, (ladder, np.abs(m - want).max())


def test_emulated_conv_dropout_masks_consistent_between_forward_and_backward():
    """regularization > 0 in a ConvMixerBlock half: the backward regenerates the forward's Philox masks (finite-difference
    check of sum(y*r) along random directions, fixed (seed, step))."""
    import ctypes as C
    from motionmixerconv_b200 import _lib as L
    rng = np.random.default_rng(5)
    B, Cn, T, E = 3, 2, 10, 20
    kt, kp = 3, 5
    params = [H.f32(1.0 + 0.2 * rng.standard_normal(E)), H.f32(0.2 * rng.standard_normal(E)),
              H.f32(0.3 * rng.standard_normal((Cn, Cn, kt, kp))), H.f32(0.1 * rng.standard_normal(Cn)),
              H.f32(0.5 * rng.standard_normal((2, T))), H.f32(0.5 * rng.standard_normal((T, 2)))]
    x = H.f32(rng.standard_normal((B, Cn, T, E)))
    r = H.f32(rng.standard_normal((B, Cn, T, E)))
    desc = L.MmxConvHalfDesc(B, Cn, T, E, kt, kp, 1, 2, 2, L.MMX_ACT["mish"], 1, 0, 1, 3, L.MmxDropout(0.3, 99, 4, None))

    def table(arrs):
        t = L.MmxConvHalfParams()
        for f, a in zip(("ln_w", "ln_b", "conv_w", "conv_b", "se_w1", "se_w2"), arrs):
            setattr(t, f, H.ptr(a))
        return t

    def fwd(xx, pp):
        y = np.empty_like(xx)
        H.call("mmx_conv_half_fwd", C.byref(desc), C.byref(table(pp)), H.ptr(xx), H.ptr(y), None)
        return y

    y0 = fwd(x, params)
    assert np.array_equal(y0, fwd(x, params))
    grads = [np.zeros_like(p) for p in params]
    dx = np.empty_like(x)
    H.call("mmx_conv_half_bwd", C.byref(desc), C.byref(table(params)), C.byref(table(grads)), H.ptr(x), H.ptr(r), H.ptr(dx), None)
    eps = 1e-2
    v = H.f32(rng.standard_normal(x.shape))
    fd = (np.sum(fwd(x + eps * v, params).astype(np.float64) * r) - np.sum(fwd(x - eps * v, params).astype(np.float64) * r)) / (2 * eps)
    an = float(np.sum(dx.astype(np.float64) * v))
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1.0), (fd, an)
    for idx in (2, 3, 0):                                  # conv weight, conv bias, LN weight
        dv = H.f32(rng.standard_normal(params[idx].shape))
        pp, pm = list(params), list(params)
        pp[idx] = H.f32(params[idx] + eps * dv)
        pm[idx] = H.f32(params[idx] - eps * dv)
        fd = (np.sum(fwd(x, pp).astype(np.float64) * r) - np.sum(fwd(x, pm).astype(np.float64) * r)) / (2 * eps)
        an = float(np.sum(grads[idx].astype(np.float64) * dv))
        assert abs(fd - an) <= 3e-2 * max(abs(an), 1.0), (idx, fd, an)


def test_conv_half_plan_capability_query():
    """mmx_conv_half_plan on the CPU build of the launch layer (same planner, same 227 KB budget as the B200 build): the
    reference's default shapes are served by the fused kernels, the wide Optuna-grid shapes are not (the module then routes
    them to the stage-kernel chain, functional.ConvHalfLarge)."""
    import ctypes as C
    from motionmixerconv_b200 import _lib as L
    from tests.emu.harness import emu
    lib = emu()

    def plan(Cn, E, kt, kp, bwd):
        d = L.MmxConvHalfDesc(256, Cn, 10, E, kt, kp, (kt - 1) // 2, (kp - 1) // 2, 1, 1, 1, 0, 1, 0, L.MmxDropout(0.0, 0, 0, None))
        S, smem = C.c_int(0), C.c_int(0)
        rc = lib.mmx_conv_half_plan(C.byref(d), int(bwd), C.byref(S), C.byref(smem))
        return rc, S.value, smem.value

    for bwd in (0, 1):
        rc, S, smem = plan(1, 50, 1, 3, bwd)            # train_mixer_h36m.py __main__ config
        assert rc == 0 and S >= 4 and 0 < smem <= 227 * 1024
        rc, S, smem = plan(4, 192, 5, 9, bwd)           # the AIS autoregressive config (K3)
        assert rc == 0 and S == 1
        rc, _, _ = plan(8, 192, 9, 29, bwd)             # Optuna grid, optuna_search/conv_optuna_main.py:339-342
        assert rc == -2 and b"shared memory" in lib.mmx_last_error() or b"weight-gradient" in lib.mmx_last_error()
    assert plan(9, 50, 1, 3, 0)[0] == -2                # conv_nChan > 8
