"""A few MixerBlock forward / backward launches through the C ABI (for ncu).  env: B, H, CH, ACT, P, PREC, ITERS."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import functional as F_
B = int(os.environ.get("B", 4096)); H = int(os.environ.get("H", 50)); ch = int(os.environ.get("CH", H)); T, tok = 10, 20
act = os.environ.get("ACT", "mish"); p = float(os.environ.get("P", 0.1)); prec = os.environ.get("PREC", "tf32")
lib = L.load()
x = torch.randn(B, T, H, device="cuda"); dy = torch.randn(B, T, H, device="cuda"); y = torch.empty_like(x); dx = torch.empty_like(x)
shapes = [(H,), (H,), (tok, T), (tok,), (T, tok), (T,), (H,), (H,), (ch, H), (ch,), (H, ch), (H,), (1, T), (T, 1)]
params = [torch.randn(*s, device="cuda") * 0.1 for s in shapes]; grads = [torch.zeros_like(q) for q in params]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
desc = F_.mlp_block_desc(B, T, H, tok, ch, 1, act, True, False, True, 0, p, 1234, 0, prec)
tw, tg = F_.mlp_block_table(params), F_.mlp_block_table(grads)
x1 = torch.empty_like(x); gate = torch.empty(B, T, device="cuda")
save = bool(lib.mmx_mlp_block_saves(C.byref(desc))) and not int(os.environ.get("NOSAVE", 0))
for _ in range(int(os.environ.get("ITERS", 3))):
    if save:
        L.check(lib, lib.mmx_mlp_block_fwd_save(C.byref(desc), C.byref(tw), x.data_ptr(), y.data_ptr(), x1.data_ptr(), gate.data_ptr(), st), "fwd")
        L.check(lib, lib.mmx_mlp_block_bwd_saved(C.byref(desc), C.byref(tw), C.byref(tg), x.data_ptr(), x1.data_ptr(), gate.data_ptr(), dy.data_ptr(), dx.data_ptr(), st), "bwd")
    else:
        L.check(lib, lib.mmx_mlp_block_fwd(C.byref(desc), C.byref(tw), x.data_ptr(), y.data_ptr(), st), "fwd")
        L.check(lib, lib.mmx_mlp_block_bwd(C.byref(desc), C.byref(tw), C.byref(tg), x.data_ptr(), dy.data_ptr(), dx.data_ptr(), st), "bwd")
torch.cuda.synchronize()
print("ok abort", lib.mmx_tc5_abort_count())
