"""Per-kernel timings through the C ABI (CUDA events on the launching stream)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import functional as F_

PEAK = 6539.2


def timeit(fn, iters=int(os.environ.get('ITERS', 20)), warm=int(os.environ.get('WARM', 5))):
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


def main():
    B = int(os.environ.get("B", 4096))
    T, H, tok, ch, act = 10, int(os.environ.get("H", 50)), 20, int(os.environ.get("CH", os.environ.get("H", 50))), os.environ.get("ACT", "mish")
    dev = "cuda"
    x = torch.randn(B, T, H, device=dev)
    dy = torch.randn(B, T, H, device=dev)
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    shapes = [(H,), (H,), (tok, T), (tok,), (T, tok), (T,), (H,), (H,), (ch, H), (ch,), (H, ch), (H,), (1, T), (T, 1)]
    params = [torch.randn(*s, device=dev) * 0.1 for s in shapes]
    grads = [torch.zeros_like(p) for p in params]
    lib = L.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for p_drop in [float(v) for v in os.environ.get('PDROP', '0.0,0.1').split(',')]:
        desc = F_.mlp_block_desc(B, T, H, tok, ch, 1, act, True, False, True, 0, p_drop, 1234, 0)
        tw, tg = F_.mlp_block_table(params), F_.mlp_block_table(grads)
        tf = timeit(lambda: L.check(lib, lib.mmx_mlp_block_fwd(C.byref(desc), C.byref(tw), x.data_ptr(), y.data_ptr(), st), "fwd"))
        tb = timeit(lambda: L.check(lib, lib.mmx_mlp_block_bwd(C.byref(desc), C.byref(tw), C.byref(tg), x.data_ptr(), dy.data_ptr(), dx.data_ptr(), st), "bwd"))
        tile = T * H * 4
        out["p%.1f" % p_drop] = dict(fwd_ms=tf, bwd_ms=tb, fwd_GBps=B * 2 * tile / tf / 1e6, bwd_GBps=B * 3 * tile / tb / 1e6,
                                     frac=B * 5 * tile / (tf + tb) / 1e6 / PEAK)
    print(json.dumps(dict(B=B, H=H, ch=ch, act=act, **out)))


if __name__ == "__main__":
    main()
