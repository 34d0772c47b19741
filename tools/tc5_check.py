"""tcgen05 MixerBlock family (precision="tf32"): per-tensor relative errors of the whole MlpMixer vs the fp64 oracle at several
batch sizes / variants, the pipeline abort counter, and block-level kernel timings.  Run on a B200 through gpurun."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mixer_np as O
from tests.golden_util import Golden, rel_err, grad_scale
from tests.synthetic import synthetic_pose_windows
from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import functional as F_
from motionmixerconv_b200.mlp_mixer import MlpMixer
from motionmixerconv_b200.functional import mpjpe_error

lib = L.load()


def model_errors(cfg, params, x, gt, prec, train=True):
    o = O.MlpMixerOracle(cfg, params, dtype=np.float64)
    p64 = o.forward(x)
    l64, dp = O.mpjpe(p64, gt.astype(np.float64))
    g64, dx64 = o.backward(dp)
    gs = grad_scale(g64)
    m = MlpMixer(**cfg)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    m = m.cuda().set_precision(prec)
    m.train() if train else m.eval()
    xg = torch.from_numpy(x).cuda().requires_grad_(True)
    pred = m(xg)
    loss = mpjpe_error(pred, torch.from_numpy(gt).cuda())
    loss.backward()
    torch.cuda.synchronize()
    rows = []
    for k, p in m.named_parameters():
        w = g64[k]
        rows.append((float(np.abs(p.grad.cpu().numpy() - w).max()) / gs, rel_err(p.grad.cpu().numpy(), w), k))
    rows.sort(reverse=True)
    return dict(pred=rel_err(pred.detach().cpu().numpy(), p64), loss=abs(float(loss) - l64) / abs(l64),
                dx=rel_err(xg.grad.cpu().numpy(), dx64), worst_grads=[(k, "%.2e (rel own %.2e)" % (e, r)) for e, r, k in rows[:4]])


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def block_timing(B, H, ch, act, prec, p_drop):
    T, tok = 10, 20
    x = torch.randn(B, T, H, device="cuda")
    dy = torch.randn(B, T, H, device="cuda")
    y, dx = torch.empty_like(x), torch.empty_like(x)
    shapes = [(H,), (H,), (tok, T), (tok,), (T, tok), (T,), (H,), (H,), (ch, H), (ch,), (H, ch), (H,), (1, T), (T, 1)]
    params = [torch.randn(*s, device="cuda") * 0.1 for s in shapes]
    grads = [torch.zeros_like(p) for p in params]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    desc = F_.mlp_block_desc(B, T, H, tok, ch, 1, act, True, False, True, 0, p_drop, 1234, 0, prec)
    tw, tg = F_.mlp_block_table(params), F_.mlp_block_table(grads)
    tf = timeit(lambda: L.check(lib, lib.mmx_mlp_block_fwd(C.byref(desc), C.byref(tw), x.data_ptr(), y.data_ptr(), st), "fwd"))
    tb = timeit(lambda: L.check(lib, lib.mmx_mlp_block_bwd(C.byref(desc), C.byref(tw), C.byref(tg), x.data_ptr(), dy.data_ptr(), dx.data_ptr(), st), "bwd"))
    return dict(B=B, H=H, ch=ch, act=act, prec=prec, p=p_drop, fwd_us=round(tf * 1e3, 1), bwd_us=round(tb * 1e3, 1))


def main():
    g = Golden("mlp_k2")
    what = os.environ.get("WHAT", "errors,timing").split(",")
    if "errors" in what:
        cfg0 = dict(g.cfg, regularization=0)
        for B in (0, 1, 13, 333, 4096, 4097):
            if B == 0:
                x, gt = g.x, g.gt
            else:
                x, gt = synthetic_pose_windows(B, 10, 10, 66, scale="h36m", seed=7)
            for prec in ("tf32",):
                print("B=%d %s" % (len(x), prec), json.dumps(model_errors(cfg0, g.params, x, gt, prec)), flush=True)
            print("   abort_count", lib.mmx_tc5_abort_count(), flush=True)
        for name, var in {"gelu_h32_ch40": dict(hidden_dim=32, channels_mlp_dim=40, activation="gelu"), "no_se": dict(use_se=False),
                          "se_hidden_2": dict(r_se=4), "h48_ch24": dict(hidden_dim=48, channels_mlp_dim=24),
                          "h64": dict(hidden_dim=64, channels_mlp_dim=64)}.items():
            cfg = dict(g.cfg, num_blocks=2, regularization=0, **var)
            torch.manual_seed(3)
            m = MlpMixer(**cfg)
            params = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
            x, gt = synthetic_pose_windows(515, 10, 10, 66, scale="amass", seed=9)
            print(name, json.dumps(model_errors(cfg, params, x, gt, "tf32")), flush=True)
        print("abort_count", lib.mmx_tc5_abort_count(), flush=True)
    if "timing" in what:
        for prec in ("fp32", "tf32"):
            for p in (0.0, 0.1):
                print(json.dumps(block_timing(4096, 50, 50, "mish", prec, p)), flush=True)
        print(json.dumps(block_timing(16384, 50, 50, "mish", "tf32", 0.1)), flush=True)
        print(json.dumps(block_timing(4096, 64, 64, "gelu", "tf32", 0.1)), flush=True)
        print(json.dumps(block_timing(4096, 64, 64, "gelu", "fp32", 0.1)), flush=True)
        print("abort_count", lib.mmx_tc5_abort_count(), flush=True)
    if "timing_wide" in what:
        for B in (4096, 16384):
            for prec in ("tf32", "fp32"):
                if prec == "fp32" and B > 4096:
                    continue
                print(json.dumps(block_timing(B, 128, 128, "gelu", prec, 0.1)), flush=True)
        print(json.dumps(block_timing(4096, 96, 96, "mish", "tf32", 0.1)), flush=True)
        print("abort_count", lib.mmx_tc5_abort_count(), flush=True)


if __name__ == "__main__":
    main()
