"""tcgen05 embedding (per-frame Linear) and output head: errors vs an fp64 torch evaluation of the same formulas at several shapes
(ragged tiles included), vs the fp32 kernels, the pipeline abort counter, and kernel timings.  Run on a B200 through gpurun."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import functional as F_

lib = L.load()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


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
    return st.elapsed_time(en) / iters * 1e3


def linear_case(rows_shape, K, N, need_dx):
    torch.manual_seed(1)
    x = (torch.randn(*rows_shape, K, device="cuda") * 0.5).requires_grad_(need_dx)
    w = (torch.randn(N, K, device="cuda") * 0.2).requires_grad_(True)
    b = (torch.randn(N, device="cuda") * 0.1).requires_grad_(True)
    dy = torch.randn(*rows_shape, N, device="cuda")
    out = {}
    ref = None
    for prec in ("fp64", "fp32", "tf32"):
        for t in (x, w, b):
            t.grad = None
        if prec == "fp64":
            y = torch.nn.functional.linear(x.double(), w.double(), b.double())
        else:
            y = F_.linear(x, w, b, prec)
        y.backward(dy.double() if prec == "fp64" else dy)
        cur = [y.detach(), w.grad.clone(), b.grad.clone()] + ([x.grad.clone()] if need_dx else [])
        if ref is None:
            ref = cur
        else:
            out[prec] = ["%.1e" % rel(c, r) for c, r in zip(cur, ref)]
    return out


def head_ref(x, p, T, To):
    ln_w, ln_b, wt, bt, wf, bf = p
    z = torch.nn.functional.layer_norm(x, (x.shape[-1],), ln_w, ln_b, 1e-5)
    pm = torch.einsum("ot,bth->boh", wt.reshape(To, T), z) + bt[None, :, None]
    return torch.nn.functional.linear(pm, wf, bf)


def head_case(B, T, To, H, D):
    torch.manual_seed(2)
    x = (torch.randn(B, T, H, device="cuda") * 2 + 0.3).requires_grad_(True)
    shapes = [(H,), (H,), (To, T, 1), (To,), (D, H), (D,)]
    p = [(torch.randn(*s, device="cuda") * 0.3 + (1.0 if i == 0 else 0.0)).requires_grad_(True) for i, s in enumerate(shapes)]
    dout = torch.randn(B, To, D, device="cuda")
    out = {}
    ref = None
    for prec in ("fp64", "fp32", "tf32"):
        for t in [x] + p:
            t.grad = None
        if prec == "fp64":
            y = head_ref(x.double(), [q.double() for q in p], T, To)
            y.backward(dout.double())
        else:
            y = F_.mlp_head(x, *p, prec)
            y.backward(dout)
        cur = [y.detach(), x.grad.clone()] + [q.grad.clone() for q in p]
        if ref is None:
            ref = cur
        else:
            out[prec] = ["%.1e" % rel(c, r) for c, r in zip(cur, ref)]
    return out


def main():
    print("linear (y, dw, db[, dx]) rel err vs fp64")
    for shape, K, N, dx in [((4096, 10), 66, 50, False), ((4097, 10), 66, 50, True), ((3, 10), 66, 50, True), ((4096, 10), 50, 66, True),
                            ((130, 10), 48, 64, True), ((50, 10), 66, 64, True)]:
        print(shape, K, N, json.dumps(linear_case(shape, K, N, dx)), flush=True)
    print("head (out, dx, dln_w, dln_b, dwt, dbt, dwf, dbf) rel err vs fp64")
    for B, T, To, H, D in [(4096, 10, 10, 50, 66), (4097, 10, 10, 50, 66), (5, 10, 10, 50, 66), (256, 10, 25, 50, 66), (333, 16, 10, 48, 66), (64, 10, 10, 62, 80)]:
        print((B, T, To, H, D), json.dumps(head_case(B, T, To, H, D)), flush=True)
    print("abort count", lib.mmx_tc5_abort_count())
    # timings at the headline shape
    B, T, D, H = 4096, 10, 66, 50
    x = torch.randn(B, T, D, device="cuda")
    w, b = torch.randn(H, D, device="cuda") * 0.1, torch.randn(H, device="cuda")
    y = torch.empty(B, T, H, device="cuda")
    dy = torch.randn(B, T, H, device="cuda")
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    import ctypes as C
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    hp = [torch.randn(*s, device="cuda") * 0.3 for s in [(H,), (H,), (10, T, 1), (10,), (D, H), (D,)]]
    hg = [torch.zeros_like(q) for q in hp]
    hw, hgt = F_.mlp_head_table(hp), F_.mlp_head_table(hg)
    desc = L.MmxMlpHeadDesc(B, T, 10, H, D)
    out = torch.empty(B, 10, D, device="cuda")
    dout = torch.randn(B, 10, D, device="cuda")
    dx = torch.empty(B, T, H, device="cuda")
    for prec in ("fp32", "tf32"):
        pc = L.MMX_PREC[prec]
        r = dict(prec=prec)
        r["embed_fwd_us"] = timeit(lambda: L.check(lib, lib.mmx_linear_fwd_prec(B * T, D, H, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), pc, st), "lf"))
        r["embed_bwd_us"] = timeit(lambda: L.check(lib, lib.mmx_linear_bwd_prec(B * T, D, H, x.data_ptr(), w.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), None, pc, st), "lb"))
        r["head_fwd_us"] = timeit(lambda: L.check(lib, lib.mmx_mlp_head_fwd_prec(C.byref(desc), C.byref(hw), y.data_ptr(), out.data_ptr(), pc, st), "hf"))
        r["head_bwd_us"] = timeit(lambda: L.check(lib, lib.mmx_mlp_head_bwd_prec(C.byref(desc), C.byref(hw), C.byref(hgt), y.data_ptr(), dout.data_ptr(), dx.data_ptr(), pc, st), "hb"))
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
    print("abort count", lib.mmx_tc5_abort_count())


if __name__ == "__main__":
    main()
