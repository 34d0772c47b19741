"""CPU suite: the harness that drives the reference's own train() (tests/ref_train.py) works with the reference's own modules
(so the GPU drop-in test, tests/test_gpu_dropin_train.py, compares like with like)."""
import pytest
import torch

from oracle import make_ref


@pytest.mark.skipif(not make_ref.train_script_available(), reason="oracle/_ref training script not staged")
def test_reference_train_runs_with_its_own_modules(tmp_path):
    from tests import ref_train as RT
    tr = RT.load_reference_train()
    MlpMixer, _, _ = make_ref.import_reference()
    torch.manual_seed(0)
    model = MlpMixer(num_classes=66, num_blocks=2, hidden_dim=32, tokens_mlp_dim=20, channels_mlp_dim=32, seq_len=10, pred_len=10,
                     activation="mish", regularization=0, input_size=66, r_se=8, use_se=True)
    out = RT.run_train(tr, model, "ref_cpu", RT.train_args(str(tmp_path), "cpu"))
    assert len(out["train"]) == 2 and out["train"][1] < out["train"][0]
    assert all(0.0 <= a <= 1.0 for a in out["auc_pck"])
    sd = torch.load(out["state_path"])
    model.load_state_dict(sd, strict=True)
