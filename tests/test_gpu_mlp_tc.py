"""GPU (-m gpu): the TF32 tensor-core MixerBlock kernels (precision="tf32", csrc/mmx_mlp_tc.cuh) through the product
modules / C ABI against the golden fixture (generated from the reference) and the numpy oracle.

Tolerances (what plain TF32 operands measurably deliver through 4 blocks x (forward + recompute + backward), see
tools/tc_model_check.py; the FP32 mode -- tests/test_gpu_mlp.py -- is the 1e-5 path and the bench headline):
  * predictions, loss: 2e-3 relative per tensor (the north star's bf16/TF32 bar; measured <= 1.1e-3 / 1e-6);
  * parameter gradients at batch sizes that average the rounding noise (>= 256 sequences): 2e-3, with the FP32 tests' floor
    convention (0.1 * rtol * the model's largest gradient: tensors far below the gradient scale are cancellation noise);
  * dx (the input gradient: the end of a 16-contraction backward chain, no batch averaging): 5e-3 -- measured 1.1e-3 ... 3.3e-3,
    i.e. NOT always inside the north star's 2e-3;
  * parameter gradients from a handful of sequences (the 6-sequence golden fixture, ragged batches of 1-4): 1e-2 (measured
    <= 7e-3).
The 3xTF32 build of the same kernels (MMX_MLP_TC_FP32=1, FP32 mode) is checked at the FP32 bar below.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mixer_np as O
from tests.golden_util import Golden, check_close, grad_scale
from tests.synthetic import synthetic_pose_windows

pytestmark = pytest.mark.gpu
TOL = 2e-3
TOL_DX = 5e-3
TOL_GRAD_FEW_SEQUENCES = 1e-2


def _model(cfg, params=None, seed=0):
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    torch.manual_seed(seed)
    m = MlpMixer(**cfg)
    if params is not None:
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    return m.cuda().set_precision("tf32")


def _run(model, x, gt):
    from motionmixerconv_b200.functional import mpjpe_error
    model.zero_grad()
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    pred = model(xg)
    loss = mpjpe_error(pred, torch.from_numpy(gt).cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    return pred.detach().cpu().numpy(), float(loss), grads, xg.grad.cpu().numpy()


def _oracle(cfg, params, x, gt):
    o = O.MlpMixerOracle(cfg, params, dtype=np.float64)
    p = o.forward(x)
    l, dp = O.mpjpe(p, gt.astype(np.float64))
    g, dx = o.backward(dp)
    return p, float(l), g, dx


def _compare(pred, loss, grads, dx, want, B):
    p64, l64, g64, dx64 = want
    check_close("pred", pred, p64, rtol=TOL)
    assert abs(loss - l64) <= TOL * abs(l64)
    tol_g = TOL if B >= 256 else TOL_GRAD_FEW_SEQUENCES
    floor = 0.1 * tol_g * grad_scale(g64)         # gradients that are pure cancellation noise next to the others
    for k, w in g64.items():
        if k in grads:
            check_close("grad " + k, grads[k], w, rtol=tol_g, atol=floor)
    check_close("dx", dx, dx64, rtol=TOL_DX)


def test_tc_kernels_are_the_ones_running():
    """precision="tf32" must change the arithmetic (else these tests would silently exercise the FP32 kernels)."""
    g = Golden("mlp_k2")
    m = _model(g.cfg, g.params).eval()
    x = torch.from_numpy(g.x).cuda()
    with torch.no_grad():
        a = m(x)
        b = m.set_precision("fp32")(x)
    d = (a - b).abs().max().item() / b.abs().max().item()
    assert 0.0 < d < TOL, d


def test_golden_k2_tf32():
    g = Golden("mlp_k2")
    model = _model(g.cfg, g.params).train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    check_close("pred vs reference", pred, g.pred, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 0.1 * TOL_GRAD_FEW_SEQUENCES * grad_scale(g.grads)
    for k, want in g.grads.items():
        check_close("grad " + k, grads[k], want, rtol=TOL_GRAD_FEW_SEQUENCES, atol=floor)      # 6 sequences
    check_close("dx", dx, g.dx, rtol=TOL_DX)
    model.eval()
    with torch.no_grad():
        pe = model(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval", pe, g.pred_eval, rtol=TOL)


@pytest.mark.parametrize("B", [1, 2, 3, 4, 333, 1000, 4096])
def test_ragged_batches_vs_oracle(B):
    """Batch sizes around the 3-sequence warp group: partial groups, a single sequence, many groups per warp."""
    g = Golden("mlp_k2")
    c = g.cfg
    x, gt = synthetic_pose_windows(B, c["seq_len"], c["pred_len"], c["input_size"], scale="amass", seed=7)
    model = _model(c, g.params).train()
    _compare(*_run(model, x, gt), _oracle(c, g.params, x, gt), B)


VARIANTS = {
    "gelu_h32_ch40": dict(hidden_dim=32, channels_mlp_dim=40, activation="gelu"),
    "no_se": dict(use_se=False),
    "se_hidden_2": dict(r_se=4),
    "h48_ch24": dict(hidden_dim=48, channels_mlp_dim=24),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_shape_variants_vs_oracle(name):
    cfg = dict(Golden("mlp_k2").cfg, num_blocks=2, **VARIANTS[name])
    model = _model(cfg, None, seed=3).train()
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    x, gt = synthetic_pose_windows(515, 10, 10, 66, scale="amass", seed=9)
    _compare(*_run(model, x, gt), _oracle(cfg, params, x, gt), 515)


def test_unsupported_shapes_fall_back_to_fp32_kernels():
    """TF32 is a permission: H = 128 (K4) is not served by the tensor-core variant and must give the FP32 result."""
    g = Golden("mlp_k4")
    m = _model(g.cfg, g.params).eval()
    x = torch.from_numpy(g.x).cuda()
    with torch.no_grad():
        a = m(x)
        b = m.set_precision("fp32")(x)
    assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["tf32", "fp32_3xtf32"])
def test_dropout_masks_consistent_between_forward_and_backward(mode, monkeypatch):
    """Block level, through the C ABI: with dropout on, <dx, v> must equal the directional derivative of sum(y * dy) along v
    (the backward regenerates exactly the forward's masks); masks change with the step.  The finite difference of the TF32
    build carries the operand-rounding noise of two forward passes (sigma ~ 25 % of the value at this size), so the sharp
    check (1e-2) runs on the 3xTF32 build of the same kernels -- identical mask code -- and the TF32 build gets a loose one."""
    from motionmixerconv_b200 import _lib as L
    from motionmixerconv_b200 import functional as F_
    prec = "tf32"
    if mode == "fp32_3xtf32":
        monkeypatch.setenv("MMX_MLP_TC_FP32", "1")
        prec = "fp32"
    torch.manual_seed(0)
    B, T, H, tok, ch = 4096, 10, 50, 20, 50
    lib = L.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shapes = [(H,), (H,), (tok, T), (tok,), (T, tok), (T,), (H,), (H,), (ch, H), (ch,), (H, ch), (H,), (1, T), (T, 1)]
    params = [torch.randn(*s, device="cuda") * 0.3 for s in shapes]
    params[0] += 1.0
    params[6] += 1.0
    x, dy, v = (torch.randn(B, T, H, device="cuda") for _ in range(3))

    def run(xin, step, bwd=False):
        desc = F_.mlp_block_desc(B, T, H, tok, ch, 1, "mish", True, False, True, 1, 0.25, 1234, step, prec)
        y = torch.empty_like(xin)
        tw = F_.mlp_block_table(params)
        L.check(lib, lib.mmx_mlp_block_fwd(C.byref(desc), C.byref(tw), xin.data_ptr(), y.data_ptr(), st), "fwd")
        if not bwd:
            return y
        grads = [torch.zeros_like(p) for p in params]
        dx = torch.empty_like(xin)
        tg = F_.mlp_block_table(grads)
        L.check(lib, lib.mmx_mlp_block_bwd(C.byref(desc), C.byref(tw), C.byref(tg), xin.data_ptr(), dy.data_ptr(), dx.data_ptr(), st), "bwd")
        return y, dx

    eps = 1e-2
    fd = float(((run(x + eps * v, 3) - run(x - eps * v, 3)).double() * dy.double()).sum() / (2 * eps))
    y3, dx = run(x, 3, bwd=True)
    an = float((dx.double() * v.double()).sum())
    # 3xTF32: what is left at eps = 1e-2 is the O(eps^2) truncation of the central difference (measured 0.5 %); wrong masks
    # at p = 0.25 would be a tens-of-percent error
    assert abs(fd - an) <= (1e-2 if mode == "fp32_3xtf32" else 0.25) * abs(fd), (fd, an)
    assert torch.equal(y3, run(x, 3))
    y4 = run(x, 4)
    assert (y3 - y4).abs().max().item() > 0
    assert (y3 != run(x, 5)).float().mean().item() > 0.5       # different masks touch most outputs


def test_3xtf32_build_meets_the_fp32_bar(monkeypatch):
    """MMX_MLP_TC_FP32=1 routes the FP32 mode through the 3xTF32 (error-compensated) build of the tensor-core kernels:
    same 1e-5 criterion as tests/test_gpu_mlp.py::test_golden."""
    monkeypatch.setenv("MMX_MLP_TC_FP32", "1")
    g = Golden("mlp_k2")
    model = _model(g.cfg, g.params).set_precision("fp32").train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    p64, l64, g64, dx64 = _oracle(g.cfg, g.params, g.x, g.gt)
    check_close("pred", pred, g.pred, p64, rtol=1e-5)
    assert abs(loss - g.loss) <= 1e-5 * abs(g.loss)
    floor = 1e-6 * grad_scale(g.grads)
    for k, want in g.grads.items():
        check_close("grad " + k, grads[k], want, g64[k], rtol=1e-5, atol=floor)
    check_close("dx", dx, g.dx, dx64, rtol=1e-5, atol=1e-6 * float(np.abs(g.dx).max()))
    monkeypatch.setenv("MMX_MLP_TC_FP32", "0")
    pred_simt = _run(model, g.x, g.gt)[0]
    assert not np.array_equal(pred, pred_simt)            # the tensor-core build really ran


def test_mpjpe_after_200_steps_within_0p1mm_of_oracle_tf32():
    """North-star criterion in the TF32 mode: MPJPE after 200 synthetic-data Adam steps within 0.1 mm of the oracle."""
    from motionmixerconv_b200.train import TrainStep
    g = Golden("mlp_k2")
    x, gt = synthetic_pose_windows(256, 10, 10, 66, scale="h36m", seed=5)
    model = _model(g.cfg, g.params).train()
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5)
    xs, gts = torch.from_numpy(x).cuda(), torch.from_numpy(gt).cuda()
    for _ in range(200):
        loss = ts.step(xs, gts)
    want = O.train_steps(O.MlpMixerOracle(g.cfg, g.params), x, gt, 200)[-1]
    assert abs(float(loss) - want) < 0.1, (float(loss), want)           # mm
