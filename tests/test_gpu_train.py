"""GPU (-m gpu): the fused training step (TrainStep: flat buffers, CUDA graph, fused Adam) against the
oracle's fwd -> MPJPE -> bwd -> Adam, and FusedAdam against torch.optim.Adam."""
import numpy as np
import pytest
import torch

from oracle import mixer_np as O
from tests.golden_util import Golden

pytestmark = pytest.mark.gpu


def _model(cfg, params):
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    m = MlpMixer(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    return m.cuda().train()


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainstep_matches_golden_three_adam_steps(use_graph):
    from motionmixerconv_b200.train import TrainStep
    g = Golden("mlp_k2")
    model = _model(g.cfg, g.params)
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, use_cuda_graph=use_graph)
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    losses = [float(ts.step(x, gt)) for _ in range(3)]
    np.testing.assert_allclose(losses, g.losses, rtol=2e-5)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g.params.keys())       # flat buffers do not disturb the state_dict
    bad = tot = 0
    for k in O.trainable_keys(g.params):
        upd = sd[k].cpu().numpy() - g.params[k]
        want = g.params3[k] - g.params[k]
        bad += int((np.abs(upd - want) > 3e-3 * 1e-2).sum())
        tot += upd.size
    assert bad / tot <= 0.005, (bad, tot)
    assert ts.steps_done == 3


def test_mpjpe_after_200_steps_within_0p1mm_of_oracle():
    """North-star criterion: MPJPE after a fixed number of synthetic-data steps within 0.1 mm."""
    from motionmixerconv_b200.train import TrainStep
    from tests.synthetic import synthetic_pose_windows
    g = Golden("mlp_k2")
    x, gt = synthetic_pose_windows(256, 10, 10, 66, scale="h36m", seed=5)
    model = _model(g.cfg, g.params)
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5)
    xs, gts = torch.from_numpy(x).cuda(), torch.from_numpy(gt).cuda()
    for _ in range(200):
        loss = ts.step(xs, gts)
    ours = float(loss)
    orc = O.MlpMixerOracle(g.cfg, g.params)
    want = O.train_steps(orc, x, gt, 200)[-1]
    assert abs(ours - want) < 0.1, (ours, want)           # mm
    assert want < O.train_steps(O.MlpMixerOracle(g.cfg, g.params), x, gt, 1)[0] - 5.0   # it did train (mm)


def test_fused_adam_matches_torch_adam():
    """Same gradients into torch.optim.Adam and FusedAdam (the gradients of model A are copied into model B, so the
    comparison is of the optimisers alone and not of the atomics' summation order), with a MultiStepLR on both."""
    from motionmixerconv_b200.functional import mpjpe_error
    from motionmixerconv_b200.train import FusedAdam
    g = Golden("mlp_odd_nose")
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    ma, mb = _model(g.cfg, g.params), _model(g.cfg, g.params)
    oa = torch.optim.Adam(ma.parameters(), lr=1e-3, weight_decay=1e-5)
    ob = FusedAdam(mb.parameters(), lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.MultiStepLR(ob, milestones=[2], gamma=0.1)
    scheda = torch.optim.lr_scheduler.MultiStepLR(oa, milestones=[2], gamma=0.1)
    for _ in range(4):
        oa.zero_grad()
        mpjpe_error(ma(x), gt).backward()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            pb.grad = pa.grad.clone()
        oa.step()
        ob.step()
        sched.step()
        scheda.step()
        with torch.no_grad():                      # keep the two models on the same trajectory
            worst = max((pa - pb).abs().max().item() for pa, pb in zip(ma.parameters(), mb.parameters()))
        assert worst <= 2e-6, worst                # one step moves a weight by <= lr = 1e-3
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa - pb).abs().max().item() <= 5e-6, k


def test_dropout_training_mode():
    """regularization=0.1: masks differ between steps, the expected keep rate holds, eval() is deterministic,
    and the backward recomputation uses the same masks as the forward (finite-difference check of one
    parameter direction would be noisy; instead check grads are finite and eval == no-dropout model)."""
    from motionmixerconv_b200.functional import mpjpe_error
    g = Golden("mlp_k2")
    cfg = dict(g.cfg, regularization=0.1)
    model = _model(cfg, g.params)
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    a, b = model(x), model(x)
    assert (a - b).abs().max().item() > 0            # fresh masks per call
    mpjpe_error(a, gt).backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    with torch.no_grad():
        e = model(x).cpu().numpy()
    np.testing.assert_allclose(e, g.pred_eval, rtol=0, atol=1e-5 * np.abs(g.pred_eval).max())


def test_prefetch_pipeline_gives_identical_results():
    """step(x, gt, prefetch=next) (double-buffered H2D on a copy stream) must train exactly like plain step(x, gt)."""
    from motionmixerconv_b200.train import TrainStep
    from tests.synthetic import synthetic_pose_windows
    g = Golden("mlp_k2")
    data = [tuple(torch.from_numpy(a).pin_memory() for a in synthetic_pose_windows(64, 10, 10, 66, scale="h36m", seed=s)) for s in range(4)]
    out = []
    for mode in ("plain", "prefetch"):
        ts = TrainStep(_model(g.cfg, g.params), lr=1e-3, weight_decay=1e-5)
        losses = []
        for i in range(8):
            if mode == "plain":
                losses.append(float(ts.step(*data[i % 4])))
            else:
                losses.append(float(ts.step(*data[i % 4], prefetch=data[(i + 1) % 4])))
        out.append(losses)
    np.testing.assert_allclose(out[0], out[1], rtol=1e-6)


def test_predict_and_step_interleave_at_different_batch_sizes():
    """predict() before / between step() calls, at the training batch size and at another one, must neither break
    the target buffer of the training plan nor throw away its captured graphs."""
    from motionmixerconv_b200.train import TrainStep
    g = Golden("mlp_k2")
    cfg = dict(g.cfg, regularization=0)
    ts = TrainStep(_model(cfg, g.params), lr=1e-3, weight_decay=1e-5)
    ref = TrainStep(_model(cfg, g.params), lr=1e-3, weight_decay=1e-5)
    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    x3 = x[:3].contiguous()
    p0 = ts.predict(x).clone()                       # predict first: no target buffer yet
    np.testing.assert_allclose(p0.cpu().numpy(), g.pred_eval, rtol=0, atol=1e-5 * np.abs(g.pred_eval).max())
    l1, r1 = float(ts.step(x, gt)), float(ref.step(x, gt))
    graph = ts.graph_a
    p3 = ts.predict(x3)                              # another batch size: a second plan, the first one stays
    assert p3.shape[0] == 3
    l2, r2 = float(ts.step(x, gt)), float(ref.step(x, gt))
    assert ts.graph_a is graph                       # the training plan's captured graphs survived
    np.testing.assert_allclose([l1, l2], [r1, r2], rtol=1e-6)      # gradient REDs are not order-deterministic
    assert ts.predict(x).shape == p0.shape
    with pytest.raises(RuntimeError):
        ts.model.eval()
        ts.step(x, gt)
    ts.model.train()
    with pytest.raises(RuntimeError):
        ts.step(x[:, :9], gt)                        # wrong frame count: rejected on the host, no out-of-bounds device write


def test_peer_adam_kernel_world_of_one_equals_the_plain_adam_kernel():
    """mmx_adam_step_peer (all-reduce over peer memory fused into Adam, csrc/mmx_api_peer.cu) with a world of one rank: the
    flag protocol runs against the rank's own flag block, the bucket lives in an mmx_peer_alloc'ed allocation, and the update
    must equal mmx_adam_step bit for bit over several replays (the epoch advances on the device).  The multi-rank behaviour is
    checked by tools/dp_check.py on 2 / 8 GPUs (profiles/)."""
    import ctypes as C
    from motionmixerconv_b200 import _lib as L
    from motionmixerconv_b200.parallel import _RawDeviceArray
    lib = L.load()
    n = 30000 + 4 * 7
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flag_bytes = lib.mmx_peer_flag_bytes(1)
    g_bytes = (n * 4 + 255) // 256 * 256
    ptr = C.c_void_p()
    L.check(lib, lib.mmx_peer_alloc(g_bytes + flag_bytes, C.byref(ptr)), "mmx_peer_alloc")
    handle = C.create_string_buffer(64)
    L.check(lib, lib.mmx_ipc_export(ptr.value, handle), "mmx_ipc_export")
    assert any(handle.raw)
    raw = _RawDeviceArray(ptr.value, n)
    g = torch.as_tensor(raw, device="cuda")
    peer_g = torch.tensor([ptr.value], dtype=torch.int64, device="cuda")
    peer_f = torch.tensor([ptr.value + g_bytes], dtype=torch.int64, device="cuda")
    epoch = torch.zeros(2, dtype=torch.int32, device="cuda")
    torch.manual_seed(0)
    pa = torch.randn(n, device="cuda")
    pb = pa.clone()
    ma, va, mb, vb = (torch.zeros(n, device="cuda") for _ in range(4))
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-5, 1.0, 1.0, 1.0, 1 - 0.9, 1 - 0.999], dtype=torch.float32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(5):
        g.copy_(torch.randn(n, device="cuda"))
        L.check(lib, lib.mmx_adam_advance(hyper.data_ptr(), step.data_ptr(), st), "advance")
        L.check(lib, lib.mmx_adam_step(pa.data_ptr(), g.data_ptr(), ma.data_ptr(), va.data_ptr(), n, hyper.data_ptr(), st), "adam")
        L.check(lib, lib.mmx_adam_step_peer(pb.data_ptr(), mb.data_ptr(), vb.data_ptr(), peer_g.data_ptr(), peer_f.data_ptr(), 0, 1, n,
                                            hyper.data_ptr(), epoch.data_ptr(), st), "adam_peer")
        torch.cuda.synchronize()
        assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb), it
        assert int(epoch[0]) == it + 1 and int(epoch[1]) == 0
    assert lib.mmx_tc5_abort_count() == 0
    del g, raw
    L.check(lib, lib.mmx_peer_free(ptr.value), "mmx_peer_free")
