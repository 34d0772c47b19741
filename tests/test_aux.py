"""The callers either side of model(x): the step-input gather/split kernel and the PCK-AUC metric.
CPU: kernel source in the emulator vs numpy restatements of train_mixer_h36m.py:117-120,179 and utils_mixer.py:20-45.
GPU (-m gpu): the product functions vs the same semantics in torch ops."""
import ctypes as C

import numpy as np
import pytest


def _ref_auc_pck(pred, gt):
    """utils_mixer.py:20-45 restated with numpy (thresholds compared in fp32, as torch does for a fp32 tensor vs a Python float)."""
    dist = np.sqrt(((pred - gt) ** 2).sum(-1, dtype=np.float32)).astype(np.float32)
    th = np.arange(0.001, 0.3, 0.001).astype(np.float32)
    pck = np.array([(dist <= t).mean(dtype=np.float64) for t in th], np.float32)
    return float(np.trapz(pck.astype(np.float64), dx=0.001) / 0.299), pck


def test_emulated_window_split_and_pck_hist():
    from tests.emu import harness as H
    rng = np.random.default_rng(0)
    B, Ttot, Dfull = 7, 35, 96
    batch = H.f32(rng.standard_normal((B, Ttot, Dfull)) * 300)
    dim = np.sort(rng.choice(Dfull, 66, replace=False)).astype(np.int32)
    x, gt = np.empty((B, 10, 66), np.float32), np.empty((B, 25, 66), np.float32)
    H.call("mmx_window_split", H.ptr(batch), B, Ttot, Dfull, dim.ctypes.data, 66, 10, 25, C.c_float(1 / 1000), C.c_float(1.0),
           H.ptr(x), H.ptr(gt), None)
    assert np.array_equal(gt, batch[:, 10:35][:, :, dim])
    np.testing.assert_allclose(x, batch[:, :10][:, :, dim] / np.float32(1000), rtol=2e-7)
    pred = H.f32(rng.standard_normal((64, 25, 22, 3)) * 0.08)
    tgt = H.f32(rng.standard_normal((64, 25, 22, 3)) * 0.08)
    th = np.arange(0.001, 0.3, 0.001).astype(np.float32)
    hist = np.zeros(len(th) + 1, np.int32)
    H.call("mmx_pck_hist", H.ptr(pred), H.ptr(tgt), pred.size // 3, th.ctypes.data, len(th), hist.ctypes.data, None)
    _, pck = _ref_auc_pck(pred, tgt)
    assert hist.sum() == pred.size // 3
    assert np.array_equal((np.cumsum(hist[:-1]) / (pred.size // 3)).astype(np.float32), pck)      # bit-exact counts


@pytest.mark.gpu
def test_window_split_and_auc_pck_on_gpu():
    import torch
    from motionmixerconv_b200.functional import auc_pck_metric, window_split
    g = torch.Generator().manual_seed(0)
    batch = (torch.randn(300, 35, 96, generator=g) * 300).cuda()
    dim_used = np.sort(np.random.default_rng(1).choice(96, 66, replace=False))
    x, gt = window_split(batch, dim_used, 10, 25, x_scale=1 / 1000)
    idx = torch.from_numpy(dim_used).cuda()
    assert torch.equal(gt, batch[:, 10:35, idx])
    torch.testing.assert_close(x, batch[:, :10, idx] / 1000, rtol=2e-7, atol=0)
    pred = (torch.randn(256, 25, 32, 3, generator=g) * 0.08).cuda()
    tgt = (torch.randn(256, 25, 32, 3, generator=g) * 0.08).cuda()
    got = float(auc_pck_metric(pred, tgt))
    want, _ = _ref_auc_pck(pred.cpu().numpy(), tgt.cpu().numpy())
    assert abs(got - want) <= 2e-6, (got, want)


@pytest.mark.gpu
def test_trainstep_from_raw_window_equals_step_on_split_tensors():
    import torch
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    from motionmixerconv_b200.train import TrainStep
    from tests.golden_util import Golden
    g = Golden("mlp_k2")
    dim_used = np.arange(15, 81)
    raw = (torch.randn(32, 35, 96, generator=torch.Generator().manual_seed(3)) * 200).cuda()
    losses = []
    for mode in ("raw", "split"):
        m = MlpMixer(**g.cfg)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.params.items()})
        ts = TrainStep(m.cuda().train(), lr=1e-3, weight_decay=1e-5)
        if mode == "raw":
            losses.append([float(ts.step_raw(raw, dim_used, 10, 10, x_scale=1 / 1000)) for _ in range(2)])
        else:
            idx = torch.from_numpy(dim_used).cuda()
            losses.append([float(ts.step(raw[:, :10, idx] / 1000, raw[:, 10:20, idx])) for _ in range(2)])
    np.testing.assert_allclose(losses[0], losses[1], rtol=2e-6)
