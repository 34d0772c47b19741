"""Whole-step timings of ConvMixer configs through TrainStep (CUDA events), for quick iteration on the GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200.conv_mixer_model import ConvMixer
from motionmixerconv_b200.train import TrainStep

CFGS = {
    "k1": (dict(num_blocks=4, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1, conv1_kernel_shape=(1, 3),
                conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation='mish', regularization=0.1, use_se=True, r_se=8), 256),
    "k1_h0_b4096": (dict(num_blocks=4, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1, conv1_kernel_shape=(1, 3),
                         conv1_padding=(0, 1), mode_conv="twice", activation='mish', regularization=0.1, use_se=True, r_se=8,
                         encoder_n_harmonic_functions=0), 4096),
    "k3_nobn": (dict(num_blocks=6, dimPosIn=33, dimPosEmb=192, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=4, conv1_kernel_shape=(5, 9),
                     mode_conv="twice", activation='mish', regularization=0, use_se=True, r_se=8, encoder_n_harmonic_functions=0,
                     encoder_omega0=0), 256),
    "c8_e192_5x5_once": (dict(num_blocks=6, dimPosIn=33, dimPosEmb=192, dimPosOut=33, in_nTP=10, out_nTP=10, conv_nChan=8,
                              conv1_kernel_shape=(5, 5), mode_conv="once", activation='mish', regularization=0, use_se=True, r_se=8,
                              encoder_n_harmonic_functions=0, encoder_omega0=0), 256),
}


def main():
    out = {}
    for name, (cfg, B) in CFGS.items():
        torch.manual_seed(0)
        model = ConvMixer(**cfg).cuda().train()
        ts = TrainStep(model, lr=1e-3, weight_decay=1e-5)
        x = torch.randn(B, cfg["in_nTP"], cfg["dimPosIn"], device="cuda") * 0.3
        gt = torch.randn(B, cfg["out_nTP"], cfg["dimPosOut"], device="cuda") * 300
        for _ in range(3):
            ts.step(x, gt)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        s.record()
        for _ in range(n):
            loss = ts.step(x, gt)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        out[name] = dict(B=B, ms_per_step=ms, seq_per_s=B / ms * 1e3, loss=float(loss))
        print(name, out[name], file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
