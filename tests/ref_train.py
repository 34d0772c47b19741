"""TEST INFRASTRUCTURE: drive the reference's own ``train()`` (h36m/train_mixer_h36m.py:47-279, staged unmodified under
oracle/_ref by oracle/make_ref.py) with a synthetic dataset, so that the drop-in claim -- "the existing train scripts can swap the
module in" (BASELINE.json north star) -- is tested end to end: DataLoader, dim_used gather, /1000,
model(x), mpjpe_error, backward, clip_grad_norm_, optim.Adam, MultiStepLR, model.eval() validation, test_mpjpe with
auc_pck_metric, torch.save(model.state_dict()).  Nothing of the script is edited; modules that only serve plotting / real
datasets (matplotlib, mpl_toolkits, h5py) are stubbed when they are not installed, and the dataset class is replaced.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch
from torch.utils.data import Dataset

from oracle import make_ref


class SyntheticH36M(Dataset):
    """Same constructor as h36m/datasets/dataset_h36m.py H36M_Dataset; windows of [input_n + output_n, 96] float32 in mm."""

    SIZES = {0: 64, 1: 32, 2: 24}

    def __init__(self, data_dir, input_n, output_n, skip_rate, actions=None, split=0):
        rng = np.random.default_rng(1000 + split)
        n = self.SIZES[split]
        base = 250.0 * rng.standard_normal((n, 1, 96))
        walk = np.cumsum(8.0 * rng.standard_normal((n, input_n + output_n, 96)), axis=1)
        self.data = torch.from_numpy((base + walk).astype(np.float32))

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        return self.data[i]


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference_train():
    if not make_ref.train_script_available():
        raise ImportError("oracle/_ref does not hold the reference training script (run `python oracle/make_ref.py` where /root/reference exists)")
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.animation = _stub("matplotlib.animation")
        tk = _stub("mpl_toolkits")
        tk.mplot3d = _stub("mpl_toolkits.mplot3d", Axes3D=object)
    try:
        import h5py  # noqa: F401
    except ImportError:
        _stub("h5py", File=object)
    if make_ref.DST not in sys.path:
        sys.path.insert(0, make_ref.DST)
    import h36m.train_mixer_h36m as tr
    tr.H36M_Dataset = SyntheticH36M
    return tr


def train_args(save_path, dev, input_n=10, output_n=10, n_epochs=2, clip_grad=1.0):
    return argparse.Namespace(save_path=save_path, dev=dev, lr=1e-3, use_scheduler=True, milestones=[1], gamma=0.1, loss_type="mpjpe",
                              data_dir="", input_n=input_n, output_n=output_n, skip_rate=1, batch_size=16, num_worker=0, n_epochs=n_epochs,
                              pose_dim=66, delta_x=False, clip_grad=clip_grad, actions_to_consider="walking", batch_size_test=8)


def run_train(tr, model, name, args, seed=7):
    torch.manual_seed(seed)               # DataLoader(shuffle=True) draws its permutation from the global generator
    train_loss, val_loss, test_loss, metrics = tr.train(model, name, args)
    f = lambda xs: [float(v) for v in xs]
    return dict(train=f(train_loss), val=f(val_loss), test=f(test_loss), auc_pck=f(metrics["auc_pck"]), mpjpe=f(metrics["mpjpe"]),
                state_path=os.path.join(args.save_path, name, "model.pt"))
