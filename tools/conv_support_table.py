"""Which ConvMixerBlock shapes do the fused half kernels serve?  Tabulates mmx_conv_half_plan (the planner behind
mmx_conv_half_{fwd,bwd}: shared-memory budget 227 KB, weight-gradient tiling) over the reference's Optuna grid
(optuna_search/conv_optuna_main.py:339-342: C = 8, E = 192, kT in {1,5,9}, kP in {1,5,...,29}; conv2 kernel rule
conv_mixer_model.py:243) and the other configurations the reference uses.  "S=.. ..K": served by the fused kernels (sequences per
CTA tile, dynamic shared memory); "chain": served by the stage-kernel chain (csrc/mmx_api_conv_large.cu, functional.ConvHalfLarge).
Runs on the CPU build of the launch layer (tests/emu: same planner, same budget) when no GPU is present.

    python tools/conv_support_table.py [--markdown]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200 import _lib as L


def lib():
    try:
        import torch
        if torch.cuda.is_available():
            return L.load()
    except Exception:
        pass
    from tests.emu.harness import emu
    return emu()


def plan(lb, B, Cn, T, E, kt, kp, bwd):
    d = L.MmxConvHalfDesc(B, Cn, T, E, kt, kp, (kt - 1) // 2, (kp - 1) // 2, max(T // 8, 1), 1, 1, 0, 1, 0, L.MmxDropout(0.0, 0, 0, None))
    S, smem = C.c_int(0), C.c_int(0)
    rc = lb.mmx_conv_half_plan(C.byref(d), int(bwd), C.byref(S), C.byref(smem))
    if rc == 0:
        return "S=%d %dK" % (S.value, smem.value // 1024)
    msg = lb.mmx_last_error().decode()
    return "chain (%s)" % ("weight-gradient tiling" if "weight-gradient" in msg else "tile > shared memory")


def main():
    lb = lib()
    T = 10
    rows = []
    for Cn, E in ((8, 192), (4, 192), (1, 50)):
        for kt in (1, 5, 9):
            for kp in range(1, 30, 4):
                k2 = (min(kp, T), min(kt, E))
                rows.append((Cn, E, kt, kp, plan(lb, 256, Cn, T, E, kt, kp, False), plan(lb, 256, Cn, T, E, kt, kp, True),
                             k2, plan(lb, 256, Cn, T, E, *k2, True)))
    print("| C | E | conv1 (kT,kP) | forward | backward | conv2 ('twice') | conv2 backward |")
    print("|---|---|---|---|---|---|---|")
    for Cn, E, kt, kp, f, b, k2, b2 in rows:
        print("| %d | %d | (%d,%d) | %s | %s | (%d,%d) | %s |" % (Cn, E, kt, kp, f, b, k2[0], k2[1], b2))


if __name__ == "__main__":
    main()
