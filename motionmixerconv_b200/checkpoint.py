"""Checkpoint / wire-format validator: is a ``torch.save(model.state_dict())`` file interchangeable between the
reference modules and this package's drop-ins?  (``train_mixer_h36m.py:276`` writes it, ``test_mixer_h36m.py:216`` loads it strictly.)

    python -m motionmixerconv_b200.checkpoint model.pt --family mlp  --cfg '{"num_classes": 66, "num_blocks": 4, ...}'
    python -m motionmixerconv_b200.checkpoint model.pt --family conv --cfg cfg.json

Checks, against the module built from the constructor arguments: the key set and the key ORDER (registration order, which
``load_state_dict(strict=True)`` and same-seed initialisation rely on), shapes and dtypes, the aliased ``se2.*`` entries of a
``mode_conv="twice"`` ConvMixer (same values as ``se.*``), BatchNorm buffers (``running_var >= 0``, integer
``num_batches_tracked``), ``encoder.frequencies`` (the harmonic ladder ``omega0 * 2**k``) and finiteness.  Runs on the CPU: building
the parameter-holder modules needs no GPU and no ``libmmx.so``.
"""
from __future__ import annotations

import argparse
import json
import sys

import torch


def build_model(family, cfg):
    if family == "mlp":
        from .mlp_mixer import MlpMixer
        return MlpMixer(**cfg)
    if family == "conv":
        from .conv_mixer_model import ConvMixer
        cfg = {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items()}
        return ConvMixer(**cfg)
    raise ValueError("family must be 'mlp' or 'conv', got %r" % (family,))


def validate_state_dict(sd, family, cfg):
    """-> list of human-readable problems (empty: the file loads strictly into both the reference module and the drop-in)."""
    want = build_model(family, cfg).state_dict()
    problems = []
    got_keys, want_keys = list(sd.keys()), list(want.keys())
    missing = [k for k in want_keys if k not in sd]
    extra = [k for k in got_keys if k not in want]
    if missing:
        problems.append("missing keys: %s" % missing)
    if extra:
        problems.append("unexpected keys: %s" % extra)
    if not missing and not extra and got_keys != want_keys:
        first = next(i for i, (a, b) in enumerate(zip(got_keys, want_keys)) if a != b)
        problems.append("key order differs from the registration order at position %d: %r vs %r" % (first, got_keys[first], want_keys[first]))
    for k in want_keys:
        if k not in sd:
            continue
        t, w = sd[k], want[k]
        if not isinstance(t, torch.Tensor):
            problems.append("%s: not a tensor (%s)" % (k, type(t).__name__))
            continue
        if tuple(t.shape) != tuple(w.shape):
            problems.append("%s: shape %s, expected %s" % (k, tuple(t.shape), tuple(w.shape)))
        if t.dtype != w.dtype:
            problems.append("%s: dtype %s, expected %s" % (k, t.dtype, w.dtype))
        if t.is_floating_point() and not bool(torch.isfinite(t).all()):
            problems.append("%s: non-finite values" % k)
        if k.endswith("running_var") and t.numel() and float(t.min()) < 0.0:
            problems.append("%s: negative running variance" % k)
        if ".se2." in k:
            twin = k.replace(".se2.", ".se.")
            if twin in sd and isinstance(sd[twin], torch.Tensor) and sd[twin].shape == t.shape and not torch.equal(sd[twin], t):
                problems.append("%s differs from %s (the reference aliases them: one shared SE layer per block)" % (k, twin))
    if family == "conv" and "encoder.frequencies" in sd and "encoder.frequencies" in want:
        f = sd["encoder.frequencies"]
        if isinstance(f, torch.Tensor) and f.shape == want["encoder.frequencies"].shape and not torch.allclose(f.float(), want["encoder.frequencies"].float(), rtol=1e-6):
            problems.append("encoder.frequencies is not omega0 * 2**arange(n_harmonic_functions) for these constructor arguments")
    return problems


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("file", help="torch.save(model.state_dict()) file")
    ap.add_argument("--family", required=True, choices=["mlp", "conv"])
    ap.add_argument("--cfg", required=True, help="constructor arguments: a JSON object or the path of a JSON file")
    args = ap.parse_args(argv)
    cfg = json.loads(args.cfg) if args.cfg.lstrip().startswith("{") else json.load(open(args.cfg))
    sd = torch.load(args.file, map_location="cpu")
    if not isinstance(sd, dict):
        print("not a state_dict: %s" % type(sd).__name__)
        return 2
    problems = validate_state_dict(sd, args.family, cfg)
    for p in problems:
        print("PROBLEM:", p)
    print("%s: %d tensors, %d parameters/buffers elements -- %s" % (args.file, len(sd), sum(t.numel() for t in sd.values() if isinstance(t, torch.Tensor)),
                                                                   "OK (interchangeable with the reference module)" if not problems else "%d problem(s)" % len(problems)))
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
