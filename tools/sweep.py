"""Large-batch throughput sweep (BASELINE.json configs[4]): K2-family MlpMixer, hidden 64 and the K2 widths, batch 16K-256K."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200.mlp_mixer import MlpMixer
from motionmixerconv_b200.train import TrainStep


def run(H, B, prec="fp32", steps=5):
    torch.manual_seed(0)
    cfg = dict(num_classes=66, num_blocks=4, hidden_dim=H, tokens_mlp_dim=20, channels_mlp_dim=H, seq_len=10, pred_len=10,
               activation="mish", regularization=0.1, input_size=66, r_se=8, use_se=True)
    model = MlpMixer(**cfg).cuda().train().set_precision(prec)
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5)
    x = torch.randn(B, 10, 66, device="cuda") * 0.3
    gt = torch.randn(B, 10, 66, device="cuda") * 300
    for _ in range(3):
        loss = ts.step(x, gt)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = ts.step(x, gt)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    assert torch.isfinite(loss)
    return dict(H=H, B=B, precision=prec, ms_per_step=ms, seq_per_s=B / ms * 1e3, loss=float(loss))


if __name__ == "__main__":
    out = []
    for H in (50, 64, 128):
        for B in (16384, 65536, 262144):
            if H == 128 and B > 65536:
                continue
            for prec in (("fp32", "tf32") if H == 50 else ("fp32",)):      # the tensor-core kernels serve H, ch <= 50
                r = run(H, B, prec)
                print(r, file=sys.stderr)
                out.append(r)
                torch.cuda.empty_cache()
    print(json.dumps(out))
