"""Per-tensor relative errors (max|got-want| / max|want|) of the MlpMixer TF32 path vs the golden fixture / fp64 oracle."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mixer_np as O
from tests.golden_util import Golden, rel_err, grad_scale
from tests.synthetic import synthetic_pose_windows
from motionmixerconv_b200.mlp_mixer import MlpMixer
from motionmixerconv_b200.functional import mpjpe_error

g = Golden("mlp_k2")
for B in (0, 333, 4096):
    if B == 0:
        x, gt = g.x, g.gt
    else:
        x, gt = synthetic_pose_windows(B, 10, 10, 66, scale="h36m", seed=7)
    o = O.MlpMixerOracle(g.cfg, g.params, dtype=np.float64)
    p64 = o.forward(x); l64, dp = O.mpjpe(p64, gt.astype(np.float64)); g64, dx64 = o.backward(dp)
    gs = grad_scale(g64)
    for prec in ("fp32", "tf32"):
        m = MlpMixer(**g.cfg); m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.params.items()}); m = m.cuda().train().set_precision(prec)
        xg = torch.from_numpy(x).cuda().requires_grad_(True)
        pred = m(xg); loss = mpjpe_error(pred, torch.from_numpy(gt).cuda()); loss.backward()
        print("B=%d %s: pred %.2e loss %.2e dx %.2e  grad_scale %.3e" % (len(x), prec, rel_err(pred.detach().cpu().numpy(), p64), abs(float(loss) - l64) / abs(l64), rel_err(xg.grad.cpu().numpy(), dx64), gs))
        rows = []
        for k, p in m.named_parameters():
            w = g64[k]
            rows.append((rel_err(p.grad.cpu().numpy(), w), float(np.abs(w).max()) / gs, k))
        for e, s, k in sorted(rows, reverse=True)[:8]:
            print("   %.2e  (|want|max / grad_scale %.1e)  %s" % (e, s, k))
