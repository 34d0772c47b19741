"""TEST INFRASTRUCTURE: run the real kernel source on the CPU.

``libmmx_emu.so`` is ``motionmixerconv_b200/csrc/mmx_api.cu`` compiled by g++ with
``-DMMX_HOST_EMU``: every kernel body runs phase by phase as a loop over thread ids, with shared
memory poisoned with NaN before each CTA.  It exposes the same C ABI as ``libmmx.so`` and is driven
here with numpy arrays.  It exists so that indexing / math of the kernels can be checked against the
oracle in a container without a GPU; it is never loaded by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from motionmixerconv_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "motionmixerconv_b200", "csrc")
EMU_SO = os.path.join(HERE, "libmmx_emu.so")


def build_emu(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "mmx.h")]
    if not force and os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(s) for s in srcs):
        return EMU_SO
    cmd = ["g++", "-O1", "-std=c++17", "-DMMX_HOST_EMU", "-x", "c++", "-shared", "-fPIC", "-o", EMU_SO,
           *sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))]
    subprocess.run(cmd, check=True)
    return EMU_SO


_emu = None


def emu():
    global _emu
    if _emu is None:
        _emu = L.bind(C.CDLL(build_emu()))
    return _emu


def ptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data


def f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def call(name, *args):
    lib = emu()
    rc = getattr(lib, name)(*args)
    L.check(lib, rc, name)


def dropout_struct(p=0.0, seed=0, step=0):
    return L.MmxDropout(p, seed, step, None)


# ---------------------------------------------------------------------------------------------
# MlpMixer assembled from the C-ABI ops (mirrors motionmixerconv_b200/mlp_mixer.py, numpy arrays)
# ---------------------------------------------------------------------------------------------
_BLOCK_KEYS = [("ln1_w", "LN1.weight"), ("ln1_b", "LN1.bias"),
               ("tok_w1", "mlp_block_token_mixing.fc1.weight"), ("tok_b1", "mlp_block_token_mixing.fc1.bias"),
               ("tok_w2", "mlp_block_token_mixing.fc2.weight"), ("tok_b2", "mlp_block_token_mixing.fc2.bias"),
               ("ln2_w", "LN2.weight"), ("ln2_b", "LN2.bias"),
               ("ch_w1", "mlp_block_channel_mixing.fc1.weight"), ("ch_b1", "mlp_block_channel_mixing.fc1.bias"),
               ("ch_w2", "mlp_block_channel_mixing.fc2.weight"), ("ch_b2", "mlp_block_channel_mixing.fc2.bias"),
               ("se_w1", "se.excitation.0.weight"), ("se_w2", "se.excitation.2.weight")]


class EmuMlpMixer:
    def __init__(self, cfg, params, training=True, dropout=None):
        self.cfg = cfg
        self.p = {k: f32(v) for k, v in params.items() if np.asarray(v).dtype.kind == "f"}
        self.training = training
        self.dropout = dropout or dropout_struct()
        self.use_se = bool(cfg.get("use_se", False))

    def _block_tables(self, i, src):
        t = L.MmxMlpBlockParams()
        for field, key in _BLOCK_KEYS:
            full = "Mixer_Block.%d.%s" % (i, key)
            setattr(t, field, ptr(src[full]) if full in src else None)
        return t

    def _desc(self, i, B):
        c = self.cfg
        return L.MmxMlpBlockDesc(B, c["seq_len"], c["hidden_dim"], c["tokens_mlp_dim"], c["channels_mlp_dim"],
                                 c["seq_len"] // c.get("r_se", 4), L.MMX_ACT[c.get("activation", "gelu")],
                                 int(self.use_se), int(c.get("use_max_pooling", False)), int(self.training), i,
                                 self.dropout)

    def _head(self, src):
        t = L.MmxMlpHeadParams()
        t.ln_w, t.ln_b = ptr(src["LN.weight"]), ptr(src["LN.bias"])
        t.wt, t.bt = ptr(src["conv_out.weight"]), ptr(src["conv_out.bias"])
        t.wf, t.bf = ptr(src["fc_out.weight"]), ptr(src["fc_out.bias"])
        return t

    def forward(self, x):
        c = self.cfg
        x = f32(x)
        B, T, D = x.shape
        H, To = c["hidden_dim"], c["pred_len"]
        self.x = x
        self.acts = [np.empty((B, T, H), np.float32)]
        call("mmx_linear_fwd", B * T, D, H, ptr(x), ptr(self.p["conv.weight"]), ptr(self.p["conv.bias"]),
             ptr(self.acts[0]), None)
        for i in range(c["num_blocks"]):
            y = np.empty((B, T, H), np.float32)
            d = self._desc(i, B)
            call("mmx_mlp_block_fwd", C.byref(d), C.byref(self._block_tables(i, self.p)), ptr(self.acts[-1]), ptr(y), None)
            self.acts.append(y)
        out = np.empty((B, To, c["num_classes"]), np.float32)
        hd = L.MmxMlpHeadDesc(B, T, To, H, c["num_classes"])
        call("mmx_mlp_head_fwd", C.byref(hd), C.byref(self._head(self.p)), ptr(self.acts[-1]), ptr(out), None)
        return out

    def backward(self, dout, need_dx=True):
        c = self.cfg
        dout = f32(dout)
        B, T, D = self.x.shape
        H, To = c["hidden_dim"], c["pred_len"]
        g = {k: np.zeros_like(v) for k, v in self.p.items()}
        d_act = np.empty((B, T, H), np.float32)
        hd = L.MmxMlpHeadDesc(B, T, To, H, c["num_classes"])
        call("mmx_mlp_head_bwd", C.byref(hd), C.byref(self._head(self.p)), C.byref(self._head(g)),
             ptr(self.acts[-1]), ptr(dout), ptr(d_act), None)
        for i in reversed(range(c["num_blocks"])):
            dx = np.empty((B, T, H), np.float32)
            d = self._desc(i, B)
            call("mmx_mlp_block_bwd", C.byref(d), C.byref(self._block_tables(i, self.p)), C.byref(self._block_tables(i, g)),
                 ptr(self.acts[i]), ptr(d_act), ptr(dx), None)
            d_act = dx
        dxin = np.empty_like(self.x) if need_dx else None
        call("mmx_linear_bwd", B * T, D, H, ptr(self.x), ptr(self.p["conv.weight"]), ptr(d_act),
             ptr(g["conv.weight"]), ptr(g["conv.bias"]), ptr(dxin), None)
        return g, dxin


def mpjpe(pred, gt, gscale=1.0):
    pred, gt = f32(pred), f32(gt)
    dpred = np.empty_like(pred)
    loss_sum = np.zeros(1, np.float32)
    n = pred.size // 3
    call("mmx_mpjpe_fwd_bwd", ptr(pred), ptr(gt), ptr(dpred), ptr(loss_sum), n, gscale, None)
    return float(loss_sum[0]) / n, dpred


def adam_step(p, g, m, v, step, lr=1e-3, wd=1e-5, b1=0.9, b2=0.999, eps=1e-8, gscale=1.0):
    hp = np.array([lr, b1, b2, eps, wd, 1 - b1 ** step, np.sqrt(1 - b2 ** step), gscale], np.float32)
    call("mmx_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.size, ptr(hp), None)
