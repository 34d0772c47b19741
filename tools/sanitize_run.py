"""One small forward + backward (+ one TrainStep) of every kernel family, for compute-sanitizer (memcheck / racecheck):
MlpMixer fp32 warp kernels (K2 shape), fp32 generic kernels (H = 72), tcgen05 family (precision tf32), ConvMixer K1-shaped
(dropout) and K3-shaped (BatchNorm) models.  env FAMILIES selects a subset (comma separated: warp,generic,tc5,conv,convbn)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200.conv_mixer_model import ConvMixer
from motionmixerconv_b200.functional import mpjpe_error
from motionmixerconv_b200.mlp_mixer import MlpMixer
from motionmixerconv_b200.train import TrainStep

fam = os.environ.get("FAMILIES", "warp,generic,tc5,conv,convbn").split(",")
B = int(os.environ.get("B", 37))
torch.manual_seed(0)


def run(model, D, To, name):
    x = torch.randn(B, 10, D, device="cuda", requires_grad=True)
    gt = torch.randn(B, To, D, device="cuda")
    loss = mpjpe_error(model(x), gt)
    loss.backward()
    ts = TrainStep(model, use_cuda_graph=False)
    l2 = ts.step(x.detach(), gt)
    torch.cuda.synchronize()
    print(name, "ok", float(loss.detach()), float(l2), flush=True)


k2 = dict(num_classes=66, num_blocks=2, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10, pred_len=10, activation="mish",
          regularization=0.1, input_size=66, r_se=8, use_se=True)
if "warp" in fam:
    run(MlpMixer(**k2).cuda().train().set_precision("fp32"), 66, 10, "mlp fp32 warp kernels")
if "generic" in fam:
    run(MlpMixer(**dict(k2, hidden_dim=72, channels_mlp_dim=40, activation="gelu")).cuda().train().set_precision("fp32"), 66, 10, "mlp fp32 generic kernels")
if "tc5" in fam:
    run(MlpMixer(**k2).cuda().train().set_precision("tf32"), 66, 10, "mlp tcgen05 family")
    run(MlpMixer(**dict(k2, hidden_dim=64, channels_mlp_dim=64, activation="gelu")).cuda().train().set_precision("tf32"), 66, 10, "mlp tcgen05 family H=64")
    from motionmixerconv_b200 import _lib as L
    print("tc5 abort count", L.load().mmx_tc5_abort_count())
if "conv" in fam:
    run(ConvMixer(num_blocks=2, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1, conv1_kernel_shape=(1, 3),
                  conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation="mish", regularization=0.1, use_se=True, r_se=8,
                  encoder_n_harmonic_functions=8).cuda().train(), 66, 25, "conv k1-shaped")
if "convbn" in fam:
    run(ConvMixer(num_blocks=2, dimPosIn=33, dimPosEmb=64, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=4, conv1_kernel_shape=(5, 9),
                  mode_conv="twice", activation="mish", regularization=-1.0, use_se=True, r_se=8, encoder_n_harmonic_functions=0,
                  encoder_omega0=0).cuda().train(), 33, 5, "conv k3-shaped (BatchNorm)")
