"""Seeded synthetic pose windows shaped like the reference's datasets.

H36M-like (``h36m/train_mixer_h36m.py:117-119,179``): model input in metres (mm/1000), target in mm.
AIS / AMASS-like (``h36m/train_mixer_ais.py:193``, ``amass/train_mixer_amass.py:77-92``): metres both sides.
"""
import numpy as np


def synthetic_pose_windows(B, T, To, D, scale="h36m", seed=1234):
    rng = np.random.default_rng(seed)
    base = 250.0 * rng.standard_normal((B, 1, D))
    walk = np.cumsum(8.0 * rng.standard_normal((B, T + To, D)), axis=1)
    seq_mm = (base + walk).astype(np.float32)
    if scale == "h36m":
        x = seq_mm[:, :T] / np.float32(1000.0)
        gt = seq_mm[:, T:T + To]
    elif scale in ("ais", "amass"):
        seq_m = seq_mm / np.float32(1000.0)
        x = seq_m[:, :T]
        gt = seq_m[:, T:T + To]
    else:
        raise ValueError(scale)
    return np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(gt, dtype=np.float32)


def synthetic_full_windows(B, Ttot, D, scale="ais", seed=1234):
    """Full [B, Ttot, D] windows for the autoregressive rollout (train_autoreg_mixer_h36m.py:208)."""
    rng = np.random.default_rng(seed)
    base = 250.0 * rng.standard_normal((B, 1, D))
    walk = np.cumsum(8.0 * rng.standard_normal((B, Ttot, D)), axis=1)
    seq = (base + walk).astype(np.float32)
    if scale != "h36m_mm":
        seq = seq / np.float32(1000.0)
    return np.ascontiguousarray(seq, dtype=np.float32)
