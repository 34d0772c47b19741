"""GPU (-m gpu): MlpMixer CUDA path through the product modules / C ABI vs the golden fixtures
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


def _model(cfg, params):
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    m = MlpMixer(**cfg)
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


MLP_CASES = [c for c in golden_cases("mlp") if not c.startswith("mlp_bn")]


@pytest.mark.parametrize("case", MLP_CASES)
def test_golden(case):
    g = Golden(case)
    model = _model(g.cfg, g.params).train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    o64 = O.MlpMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(g.x)
    _, dp64 = O.mpjpe(p64, g.gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    check_close("pred", pred, g.pred, p64, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 1e-6 * grad_scale(g.grads)
    for k, want in g.grads.items():
        check_close("grad " + k, grads[k], want, g64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))
    model.eval()
    with torch.no_grad():
        pe = model(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval", pe, g.pred_eval, rtol=TOL)


@pytest.mark.parametrize("case,B", [("mlp_k2", 333), ("mlp_k4", 70), ("mlp_odd_nose", 257), ("mlp_maxpool", 129)])
def test_vs_oracle_ragged_batch(case, B):
    """Batch sizes that are not multiples of the CTA tile: several tiles per CTA + a partial last tile."""
    g = Golden(case)
    c = g.cfg
    x, gt = synthetic_pose_windows(B, c["seq_len"], c["pred_len"], c["input_size"], scale="amass", seed=7)
    model = _model(c, g.params).train()
    pred, loss, grads, dx = _run(model, x, gt)
    res = {}
    for dt in (np.float32, np.float64):
        o = O.MlpMixerOracle(c, g.params, dtype=dt)
        p = o.forward(x)
        l, dp = O.mpjpe(p, gt.astype(dt))
        gr, dxx = o.backward(dp)
        res[dt] = (p, l, gr, dxx)
    p32, l32, g32, dx32 = res[np.float32]
    p64, l64, g64, dx64 = res[np.float64]
    check_close("pred", pred, p32, p64, rtol=TOL)
    assert abs(loss - float(l64)) <= TOL * abs(float(l64))
    floor = 1e-6 * grad_scale(g32)
    for k in g32:
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))


@pytest.mark.parametrize("case,B", [("mlp_k2", 4096), ("mlp_k2", 4097), ("mlp_k4", 4096)])
def test_vs_oracle_at_the_benchmark_batch(case, B):
    """FP32 mode against the numpy oracle (fp32 + fp64) AT the benchmark size: every warp of the persistent kernels runs
    several loop iterations, the shared-memory accumulator locks are contended, the re-alignment barriers are live
    (B = 4097: plus a ragged last group)."""
    g = Golden(case)
    c = dict(g.cfg, regularization=0)
    x, gt = synthetic_pose_windows(B, c["seq_len"], c["pred_len"], c["input_size"], scale="h36m" if case == "mlp_k2" else "amass", seed=13)
    model = _model(c, g.params).train()
    pred, loss, grads, dx = _run(model, x, gt)
    res = {}
    for dt in (np.float32, np.float64):
        o = O.MlpMixerOracle(c, g.params, dtype=dt)
        p = o.forward(x)
        l, dp = O.mpjpe(p, gt.astype(dt))
        gr, dxx = o.backward(dp)
        res[dt] = (p, l, gr, dxx)
    p32, l32, g32, dx32 = res[np.float32]
    p64, l64, g64, dx64 = res[np.float64]
    check_close("pred", pred, p32, p64, rtol=TOL)
    assert abs(loss - float(l64)) <= TOL * abs(float(l64))
    floor = 1e-6 * grad_scale(g32)
    for k in g32:
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))


def test_state_dict_roundtrip_and_errors():
    g = Golden("mlp_k2")
    model = _model(g.cfg, g.params)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g.params.keys())
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 10, 66))            # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 9, 66).cuda())      # wrong seq_len


def test_large_batch_size_independent_properties():
    """At the BASELINE size (B=4096): shard linearity — the loss/gradients of the whole batch equal the
    mean of the two half-batches' (what data-parallel training relies on)."""
    from motionmixerconv_b200.functional import mpjpe_error
    g = Golden("mlp_k2")
    model = _model(g.cfg, g.params).train()
    x, gt = synthetic_pose_windows(4096, 10, 10, 66, scale="h36m", seed=11)
    xs, gts = torch.from_numpy(x).cuda(), torch.from_numpy(gt).cuda()

    def grads_of(xb, gb):
        model.zero_grad()
        l = mpjpe_error(model(xb), gb)
        l.backward()
        return float(l), torch.cat([p.grad.flatten() for p in model.parameters()]).clone()

    l_all, g_all = grads_of(xs, gts)
    l_a, g_a = grads_of(xs[:2048], gts[:2048])
    l_b, g_b = grads_of(xs[2048:], gts[2048:])
    assert abs(l_all - 0.5 * (l_a + l_b)) <= 1e-5 * abs(l_all)
    err = (g_all - 0.5 * (g_a + g_b)).abs().max().item() / g_all.abs().max().item()
    assert err < 2e-5, err


# ---- BatchNorm1d inside the MLP blocks (regularization == -1; mlp_mixer.py:72-73, sampled by optuna_search/optuna_main.py:189-190)
@pytest.mark.parametrize("case", ["mlp_bn", "mlp_bn_maxpool"])
def test_batchnorm_golden_train_and_eval(case):
    g = Golden(case)
    model = _model(g.cfg, g.params).train()
    pred, loss, grads, dx = _run(model, g.x, g.gt)
    o64 = O.MlpMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o64.forward(g.x)
    _, dp64 = O.mpjpe(p64, g.gt.astype(np.float64))
    g64, dx64 = o64.backward(dp64)
    check_close("pred", pred, g.pred, p64, rtol=TOL)
    assert abs(loss - g.loss) <= TOL * abs(g.loss)
    floor = 5e-6 * grad_scale(g.grads)   # exactly-cancelling gradients (a bias in front of a BatchNorm) are pure rounding noise
    for k, want in g.grads.items():
        check_close("grad " + k, grads[k], want, g64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, g.dx, dx64, rtol=TOL, atol=1e-6 * float(np.abs(g.dx).max()))
    sd = model.state_dict()
    n_bn = 0
    for k in sd:                                    # running statistics after ONE training forward == the reference's
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), g.params1[k], rtol=1e-5, atol=1e-7, err_msg=k)
            n_bn += 1
        if "num_batches_tracked" in k:
            assert int(sd[k]) == int(g.params[k]) + 1
    assert n_bn == 2 * 4 * g.cfg["num_blocks"]
    # eval mode uses the running statistics
    ref = O.MlpMixerOracle(g.cfg, {k: v.cpu().numpy() for k, v in sd.items()}, dtype=np.float64)
    pe64 = ref.forward(g.x, training=False)
    model.eval()
    with torch.no_grad():
        pe = model(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval (after one update)", pe, pe64.astype(np.float32), pe64, rtol=TOL)
    fresh = _model(g.cfg, g.params).eval()
    with torch.no_grad():
        pe0 = fresh(torch.from_numpy(g.x).cuda()).cpu().numpy()
    check_close("pred_eval", pe0, g.pred_eval, rtol=TOL)


@pytest.mark.parametrize("variant", ["mish_maxpool", "no_se_h50", "wide_B333"])
def test_batchnorm_variants_vs_oracle(variant):
    g = Golden("mlp_bn")
    extra = {"mish_maxpool": dict(activation="mish", use_max_pooling=True, r_se=4),
             "no_se_h50": dict(use_se=False, hidden_dim=50, channels_mlp_dim=44, tokens_mlp_dim=12),
             "wide_B333": dict(hidden_dim=128, channels_mlp_dim=96)}[variant]
    cfg = dict(g.cfg, **extra)
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    torch.manual_seed(5)
    model = MlpMixer(**cfg)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if ".reg" in k:                          # non-trivial BatchNorm affine
                p.add_(0.3 * torch.randn_like(p))
    params = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    B = 333 if variant == "wide_B333" else 37
    x, gt = synthetic_pose_windows(B, cfg["seq_len"], cfg["pred_len"], cfg["input_size"], scale="amass", seed=3)
    pred, loss, grads, dx = _run(model, x, gt)
    res = {}
    for dt in (np.float32, np.float64):
        o = O.MlpMixerOracle(cfg, params, dtype=dt)
        p = o.forward(x)
        l, dp = O.mpjpe(p, gt.astype(dt))
        gr, dxx = o.backward(dp)
        res[dt] = (p, l, gr, dxx, o)
    p32, l32, g32, dx32, _ = res[np.float32]
    p64, l64, g64, dx64, o64 = res[np.float64]
    check_close("pred", pred, p32, p64, rtol=TOL)
    assert abs(loss - float(l64)) <= TOL * abs(float(l64))
    floor = 5e-6 * grad_scale(g32)
    for k in g32:
        check_close("grad " + k, grads[k], g32[k], g64[k], rtol=TOL, atol=floor)
    check_close("dx", dx, dx32, dx64, rtol=TOL, atol=1e-6 * float(np.abs(dx32).max()))
    sd = model.state_dict()
    for k in sd:
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), o64.p[k], rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("use_graph", [False, True])
def test_batchnorm_trainstep_three_adam_steps(use_graph):
    from motionmixerconv_b200.train import TrainStep
    g = Golden("mlp_bn")
    model = _model(g.cfg, g.params).train()
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, use_cuda_graph=use_graph)
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    losses = [float(ts.step(x, gt)) for _ in range(3)]
    np.testing.assert_allclose(losses, g.losses, rtol=5e-5)
    sd = model.state_dict()
    assert int(sd["Mixer_Block.0.mlp_block_token_mixing.reg1.num_batches_tracked"]) == 3
    for k in g.params3:
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), g.params3[k], rtol=2e-4, atol=1e-6, err_msg=k)
        elif "num_batches" not in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), g.params3[k], rtol=2e-3, atol=2e-5, err_msg=k)
    p = ts.predict(x)                                 # eval semantics through the same plan: running statistics
    ref = O.MlpMixerOracle(g.cfg, {k: v.cpu().numpy() for k, v in sd.items()}, dtype=np.float64)
    check_close("predict", p.cpu().numpy(), ref.forward(g.x, training=False), rtol=TOL)
