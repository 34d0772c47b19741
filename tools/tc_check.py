"""Block-level check of the TF32 tensor-core MixerBlock kernels against the FP32 kernels (same C ABI, precision flag).
python tools/tc_check.py            (env: B, H, CH, ACT, SE, RR)"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import functional as F_


def run(B, H, ch, act, use_se, rr, p_drop, prec, params, x, dy, step=0):
    lib = L.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    T, tok = 10, 20
    desc = F_.mlp_block_desc(B, T, H, tok, ch, rr, act, use_se, False, True, 1, p_drop, 1234, step, prec)
    grads = [torch.zeros_like(p) for p in params]
    y, dx = torch.empty_like(x), torch.empty_like(x)
    tw, tg = F_.mlp_block_table(params), F_.mlp_block_table(grads)
    L.check(lib, lib.mmx_mlp_block_fwd(C.byref(desc), C.byref(tw), x.data_ptr(), y.data_ptr(), st), "fwd")
    torch.cuda.synchronize()
    L.check(lib, lib.mmx_mlp_block_bwd(C.byref(desc), C.byref(tw), C.byref(tg), x.data_ptr(), dy.data_ptr(), dx.data_ptr(), st), "bwd")
    torch.cuda.synchronize()
    return y, dx, grads


NAMES = ["ln1_w", "ln1_b", "tok_w1", "tok_b1", "tok_w2", "tok_b2", "ln2_w", "ln2_b", "ch_w1", "ch_b1", "ch_w2", "ch_b2", "se_w1", "se_w2"]


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    torch.manual_seed(0)
    B = int(os.environ.get("B", 100))
    H = int(os.environ.get("H", 50)); ch = int(os.environ.get("CH", H)); act = os.environ.get("ACT", "mish")
    use_se = bool(int(os.environ.get("SE", 1))); rr = int(os.environ.get("RR", 1))
    T, tok = 10, 20
    dev = "cuda"
    x = torch.randn(B, T, H, device=dev)
    dy = torch.randn(B, T, H, device=dev)
    shapes = [(H,), (H,), (tok, T), (tok,), (T, tok), (T,), (H,), (H,), (ch, H), (ch,), (H, ch), (H,), (rr, T), (T, rr)]
    params = [torch.randn(*s, device=dev) * 0.3 for s in shapes]
    params[0] += 1.0; params[6] += 1.0
    if not use_se:
        params[12] = params[13] = None
    out = {}
    y0, dx0, g0 = run(B, H, ch, act, use_se, rr, 0.0, "fp32", params, x, dy)
    y1, dx1, g1 = run(B, H, ch, act, use_se, rr, 0.0, "tf32", params, x, dy)
    out["y"] = rel(y1, y0); out["dx"] = rel(dx1, dx0)
    for n, a, b in zip(NAMES, g1, g0):
        if a is not None:
            out["g_" + n] = rel(a, b)
    out["finite"] = bool(torch.isfinite(y1).all() and torch.isfinite(dx1).all())
    # dropout: forward/backward mask consistency through a directional derivative of sum(y * dy)
    p = 0.25
    v = torch.randn_like(x)
    eps = 1e-2
    yp, _, _ = run(B, H, ch, act, use_se, rr, p, "tf32", params, x + eps * v, dy, step=3)
    ym, _, _ = run(B, H, ch, act, use_se, rr, p, "tf32", params, x - eps * v, dy, step=3)
    _, dxd, gd = run(B, H, ch, act, use_se, rr, p, "tf32", params, x, dy, step=3)
    fd = float(((yp - ym) * dy).sum() / (2 * eps))
    an = float((dxd * v).sum())
    out["drop_fd"] = fd; out["drop_an"] = an; out["drop_rel"] = abs(fd - an) / max(abs(fd), 1e-30)
    ya, _, _ = run(B, H, ch, act, use_se, rr, p, "tf32", params, x, dy, step=3)
    yb, _, _ = run(B, H, ch, act, use_se, rr, p, "tf32", params, x, dy, step=4)
    out["drop_same_step_equal"] = bool((ya - run(B, H, ch, act, use_se, rr, p, "tf32", params, x, dy, step=3)[0]).abs().max().item() == 0.0)
    out["drop_steps_differ"] = bool((ya - yb).abs().max().item() > 0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
