"""GPU (-m gpu): autoregressive rollout (train_autoreg_mixer_h36m.py:195-258) through the drop-in ConvMixer vs the same
rollout restated with the numpy oracle: loss, stitched predictions and — with BPTT through the predictions — every
parameter gradient.  Config: BASELINE configs[2] shape (AIS-like 33 dims, 10 -> 5 frames, 5 chained passes, step 5)."""
import types

import numpy as np
import pytest
import torch

from oracle import mixer_np as O
from tests.golden_util import Golden, check_close, grad_scale
from tests.synthetic import synthetic_full_windows

pytestmark = pytest.mark.gpu
ARGS = types.SimpleNamespace(input_n_dataset=10, output_n_dataset=25, input_n_model=10, output_n_model=5, step_window=5, loss_type='mpjpe')


def _oracle_rollout(cfg, params, seq, teacher_forcing, dtype):
    """-> (loss, predict [B,25,D], grads) with gradients flowing through the predictions when not teacher forcing."""
    a = ARGS
    seq = seq.astype(dtype)
    starts = list(range(0, a.input_n_dataset + a.output_n_dataset - a.input_n_model - a.output_n_model + 1, a.step_window))
    nwin = a.output_n_dataset // a.step_window
    passes, loss = [], 0.0
    predict = np.zeros((seq.shape[0], a.output_n_dataset, seq.shape[2]), dtype)
    win = seq[:, :a.input_n_model]
    for st in starts:
        et = st + a.input_n_model
        if teacher_forcing:
            win = seq[:, st:et]
        orc = O.ConvMixerOracle(cfg, params, dtype=dtype)
        pred = orc.forward(win)
        l, dpred = O.mpjpe(pred, seq[:, et:et + a.output_n_model])
        loss += float(l)
        predict[:, st:st + a.output_n_model] = pred
        passes.append((orc, dpred))
        if not teacher_forcing:
            win = np.concatenate((win[:, a.step_window:], pred), axis=1)
    assert teacher_forcing, "free-running gradients: use _oracle_rollout_bptt"
    grads = {}
    for orc, dpred in passes:          # teacher forcing: every window sees ground-truth input, the passes are independent
        g, _ = orc.backward(dpred / nwin)
        for k, v in g.items():
            grads[k] = grads.get(k, 0) + v
    return loss / nwin, predict, grads


def _oracle_rollout_bptt(cfg, params, seq, dtype):
    """Free-running rollout with exact BPTT: keeps a per-frame gradient buffer over the stitched input timeline."""
    a = ARGS
    seq = seq.astype(dtype)
    starts = list(range(0, a.input_n_dataset + a.output_n_dataset - a.input_n_model - a.output_n_model + 1, a.step_window))
    nwin = a.output_n_dataset // a.step_window
    B, _, D = seq.shape
    timeline = np.zeros((B, a.input_n_model + a.output_n_dataset, D), dtype)     # frames the model actually sees
    timeline[:, :a.input_n_model] = seq[:, :a.input_n_model]
    passes, loss = [], 0.0
    for st in starts:
        et = st + a.input_n_model
        orc = O.ConvMixerOracle(cfg, params, dtype=dtype)
        pred = orc.forward(timeline[:, st:et])
        l, dpred = O.mpjpe(pred, seq[:, et:et + a.output_n_model])
        loss += float(l)
        timeline[:, et:et + a.output_n_model] = pred
        passes.append((orc, dpred, st, et))
    dline = np.zeros_like(timeline)
    grads = {}
    for orc, dpred, st, et in reversed(passes):
        dout = dpred / nwin + dline[:, et:et + a.output_n_model]
        g, dx = orc.backward(dout)
        for k, v in g.items():
            grads[k] = grads.get(k, 0) + v
        dline[:, st:et] += dx
    return loss / nwin, timeline[:, a.input_n_model:], grads


def _model(cfg, params):
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    m = ConvMixer(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    return m.cuda()


@pytest.mark.parametrize("teacher_forcing", [True, False])
def test_rollout_training_matches_oracle(teacher_forcing):
    from motionmixerconv_b200.rollout import autoregressive_process_batch
    g = Golden("conv_k3")
    cfg = g.cfg
    seq = synthetic_full_windows(6, 35, 33, scale="ais", seed=17)
    model = _model(cfg, g.params).train()
    batch = torch.from_numpy(seq).cuda()
    loss, predict = autoregressive_process_batch(batch, model, ARGS, list(range(33)), teacher_forcing)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
    if teacher_forcing:
        ref32 = _oracle_rollout(cfg, g.params, seq, True, np.float32)
        ref64 = _oracle_rollout(cfg, g.params, seq, True, np.float64)
    else:
        ref32 = _oracle_rollout_bptt(cfg, g.params, seq, np.float32)
        ref64 = _oracle_rollout_bptt(cfg, g.params, seq, np.float64)
    assert abs(float(loss) - ref64[0]) <= 1e-5 * abs(ref64[0])
    check_close("predict", predict.cpu().numpy(), ref32[1], ref64[1], rtol=1e-5)
    floor = 1e-6 * grad_scale(ref32[2])
    for k in O.trainable_keys(g.params):
        check_close("grad " + k, grads[k], ref32[2][k], ref64[2][k], rtol=1e-5, atol=floor)


def test_rollout_executor_cuda_graph_matches_eager():
    from motionmixerconv_b200.rollout import RolloutExecutor, autoregressive_process_batch
    g = Golden("conv_k3")
    model = _model(g.cfg, g.params).eval()
    ex = RolloutExecutor(model, 10, 25, 10, 5, 5)
    for seed in (1, 2):                                   # second call replays the captured graph on new data
        seq = synthetic_full_windows(8, 35, 33, scale="ais", seed=seed)
        batch = torch.from_numpy(seq).cuda()
        with torch.no_grad():
            _, want = autoregressive_process_batch(batch, model, ARGS, list(range(33)), False)
        got = ex(batch)
        assert torch.equal(got, want)
        ref = _oracle_rollout_bptt(g.cfg, g.params, seq, np.float64)[1]
        check_close("predict", got.cpu().numpy(), ref.astype(np.float32), ref, rtol=1e-5)


@pytest.mark.parametrize("teacher_forcing", [True, False])
def test_rollout_trainer_cuda_graph_matches_eager_steps(teacher_forcing):
    """RolloutTrainer (chained passes + BPTT + fused Adam as ONE graph replay, device-side NaN flag) against the eager loop
    body of train_autoregressive (autoregressive_process_batch -> backward -> FusedAdam.step): same losses, same weights."""
    from motionmixerconv_b200.rollout import RolloutTrainer, autoregressive_process_batch
    from motionmixerconv_b200.train import FusedAdam
    g = Golden("conv_k3_bn")
    data = [torch.from_numpy(synthetic_full_windows(8, 35, 33, scale="ais", seed=s)).cuda() for s in (1, 2, 3)]
    eager = _model(g.cfg, g.params).train()
    opt = FusedAdam(eager.parameters(), lr=1e-3, weight_decay=1e-5)
    want = []
    for b in data:
        opt.zero_grad(set_to_none=True)
        loss, _ = autoregressive_process_batch(b, eager, ARGS, list(range(33)), teacher_forcing)
        loss.backward()
        opt.step()
        want.append(float(loss.detach()))
    model = _model(g.cfg, g.params).train()
    tr = RolloutTrainer(model, ARGS, list(range(33)), teacher_forcing=teacher_forcing, lr=1e-3, weight_decay=1e-5)
    got = []
    for b in data:
        loss, predict = tr.step(b)
        got.append(float(loss))
        assert predict.shape == (8, 25, 33) and not bool(tr.nan_flag)
    np.testing.assert_allclose(got, want, rtol=2e-5)
    bad = tot = 0
    for (k, a), b_ in zip(eager.state_dict().items(), model.state_dict().values()):
        if a.dtype.is_floating_point:
            bad += int(((a - b_).abs() > 0.03 * 3e-3).sum())
            tot += a.numel()
        else:
            assert torch.equal(a, b_), k                  # num_batches_tracked
    assert bad / tot <= 0.01, (bad, tot)


def test_angle_l1_loss_kernel_and_rollout_with_the_angle_loss():
    """loss_type == 'angle' (train_mixer_h36m.py:187, train_autoreg_mixer_h36m.py:209-210): the fused L1 kernel against the
    reference's torch expression (value and gradient), and a teacher-forced rollout with it against the same rollout whose
    loss is the torch expression on the same predictions."""
    from motionmixerconv_b200.functional import angle_l1_error
    from motionmixerconv_b200.rollout import autoregressive_process_batch
    torch.manual_seed(3)
    pred = torch.randn(37, 5, 48, device="cuda", requires_grad=True)
    gt = torch.randn(37, 5, 48, device="cuda")
    with torch.no_grad():
        pred[0, 0, :4] = gt[0, 0, :4]                   # exact zeros: sign(0) = 0 as torch.abs' backward
    want = torch.mean(torch.sum(torch.abs(pred.reshape(-1, 5, 48) - gt), dim=2).view(-1))
    (gw,) = torch.autograd.grad(want, pred)
    got = angle_l1_error(pred.reshape(-1, 5, 48), gt)
    (gg,) = torch.autograd.grad(got, pred)
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    assert torch.equal(gg, gw)
    g = Golden("conv_k3")
    seq = synthetic_full_windows(6, 35, 33, scale="ais", seed=17)
    batch = torch.from_numpy(seq).cuda()
    args = types.SimpleNamespace(**dict(vars(ARGS), loss_type="angle"))
    dim_used = list(range(33))
    model = _model(g.cfg, g.params).train()
    loss, predict = autoregressive_process_batch(batch, model, args, dim_used, True)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    model.zero_grad()
    torch_l1 = lambda p, t: torch.mean(torch.sum(torch.abs(p.reshape(-1, args.output_n_model, len(dim_used)) - t), dim=2).view(-1))
    loss2, predict2 = autoregressive_process_batch(batch, model, args, dim_used, True, loss_fn=torch_l1)
    loss2.backward()
    assert abs(float(loss) - float(loss2)) <= 1e-5 * abs(float(loss2))
    assert (predict - predict2).abs().max().item() <= 2e-6 * predict2.abs().max().item()     # (the encoder's split-K partial sums combine in any order)
    scale = max(p.grad.abs().max().item() for p in model.parameters())
    for k, p in model.named_parameters():
        assert (p.grad - grads[k]).abs().max().item() <= 2e-5 * scale, k       # same dL/dpred; the kernels' atomics reorder sums
