"""Checkpoint / wire format: ``torch.save(model.state_dict())`` files are interchangeable with the REAL reference
modules in both directions (strict), incl. the aliased ``se2.*`` keys, BatchNorm buffers and ``encoder.frequencies``; and a
TrainStep run resumes bit-compatibly from (model.state_dict(), TrainStep.state_dict()).
The reference modules come from oracle/_ref (unmodified copies staged by oracle/make_ref.py)."""
import io

import numpy as np
import pytest
import torch

from oracle import make_ref
from tests.golden_util import Golden

needs_ref = pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not staged (python oracle/make_ref.py)")

MLP = dict(num_classes=66, num_blocks=3, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10, pred_len=10, activation="mish",
           regularization=0.1, input_size=66, r_se=8, use_se=True)
CONVS = {
    "twice_se_harmonic": dict(num_blocks=2, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1, conv1_kernel_shape=(1, 3),
                              conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation="mish", regularization=0.1, use_se=True, r_se=8),
    "once_bn_c4": dict(num_blocks=2, dimPosIn=33, dimPosEmb=64, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=4, conv1_kernel_shape=(5, 9),
                       mode_conv="once", activation="mish", regularization=-1.0, use_se=True, r_se=8, encoder_n_harmonic_functions=0, encoder_omega0=0),
    "twice_bn_nose": dict(num_blocks=1, dimPosIn=33, dimPosEmb=32, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=2, conv1_kernel_shape=(3, 3),
                          mode_conv="twice", activation="gelu", regularization=-1.0, use_se=False, encoder_n_harmonic_functions=4, encoder_omega0=0.1),
}


def _roundtrip(sd):
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    return torch.load(buf)


def _check_both_ways(ours, ref):
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    assert [tuple(v.shape) for v in sd_o.values()] == [tuple(v.shape) for v in sd_r.values()]
    assert [v.dtype for v in sd_o.values()] == [v.dtype for v in sd_r.values()]
    # product checkpoint -> reference module (strict), and the values really arrive
    ref.load_state_dict(_roundtrip(sd_o), strict=True)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, sd_o[k]), k
    # reference checkpoint -> product module (strict)
    for p in ref.parameters():
        p.data.add_(1.0)
    ours.load_state_dict(_roundtrip(ref.state_dict()), strict=True)
    for k, v in ours.state_dict().items():
        assert torch.equal(v, ref.state_dict()[k]), k


@needs_ref
def test_mlpmixer_checkpoints_interchange_with_the_reference_module():
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    Ref, _, _ = make_ref.import_reference()
    torch.manual_seed(1)
    ours = MlpMixer(**MLP)
    torch.manual_seed(2)
    ref = Ref(**MLP)
    _check_both_ways(ours, ref)


@needs_ref
@pytest.mark.parametrize("name", sorted(CONVS))
def test_convmixer_checkpoints_interchange_with_the_reference_module(name):
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    _, Ref, _ = make_ref.import_reference()
    torch.manual_seed(1)
    ours = ConvMixer(**CONVS[name])
    torch.manual_seed(2)
    ref = Ref(**CONVS[name])
    _check_both_ways(ours, ref)
    if CONVS[name]["use_se"] and CONVS[name]["mode_conv"] == "twice":
        sd = ours.state_dict()
        assert sd["Mixer_Block.0.se2.excitationBlock.0.weight"].data_ptr() == sd["Mixer_Block.0.se.excitationBlock.0.weight"].data_ptr()


@needs_ref
def test_same_seed_gives_the_reference_modules_initial_weights():
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    Ref, _, _ = make_ref.import_reference()
    torch.manual_seed(7)
    a = MlpMixer(**MLP).state_dict()
    torch.manual_seed(7)
    b = Ref(**MLP).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in b)


@pytest.mark.gpu
def test_trainstep_resumes_from_a_checkpoint():
    """3 steps, save (model + optimiser state) through torch.save, load into fresh objects, 2 more steps == 5 uninterrupted."""
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    from motionmixerconv_b200.train import TrainStep
    g = Golden("mlp_k2")
    cfg = dict(g.cfg, regularization=0)

    def fresh():
        m = MlpMixer(**cfg)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.params.items()}, strict=True)
        return m.cuda().train()

    x, gt = torch.from_numpy(g.x).cuda(), torch.from_numpy(g.gt).cuda()
    a = TrainStep(fresh(), lr=1e-3, weight_decay=1e-5)
    for _ in range(5):
        a.step(x, gt)
    b = TrainStep(fresh(), lr=1e-3, weight_decay=1e-5)
    for _ in range(3):
        b.step(x, gt)
    ck = _roundtrip({"model": b.model.state_dict(), "optim": b.state_dict()})
    assert ck["optim"]["step"] == 3 and set(ck["optim"]["exp_avg"]) == {n for n, _ in b.model.named_parameters()}
    m2 = MlpMixer(**cfg)
    m2.load_state_dict(ck["model"], strict=True)
    c = TrainStep(m2.cuda().train(), lr=1e-3, weight_decay=1e-5)
    c.load_state_dict(ck["optim"])
    for _ in range(2):
        c.step(x, gt)
    assert c.steps_done == 5
    # Adam normalises every gradient to ~lr per step, so an element whose gradient is rounding noise (order of the REDs) may move
    # differently: count elements that disagree by more than 3 % of the 5e-3 the weights travelled in 5 steps
    bad = tot = 0
    for (k, va), vc in zip(a.model.state_dict().items(), c.model.state_dict().values()):
        bad += int(((va - vc).abs() > 0.03 * 5e-3).sum())
        tot += va.numel()
    assert bad / tot <= 0.005, (bad, tot)
    # and resuming WITHOUT the optimiser state is measurably different (the test would not notice a no-op load otherwise)
    d = TrainStep(fresh(), lr=1e-3, weight_decay=1e-5)
    d.model.load_state_dict(ck["model"], strict=True)
    for _ in range(2):
        d.step(x, gt)
    worse = sum(int(((va - vd).abs() > 0.03 * 5e-3).sum()) for va, vd in zip(a.model.state_dict().values(), d.model.state_dict().values()))
    assert worse > 10 * max(bad, 1)
