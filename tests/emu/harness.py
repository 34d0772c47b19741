"""TEST INFRASTRUCTURE: run the real kernel source on the CPU.

``libmmx_emu.so`` is ``motionmixerconv_b200/csrc/mmx_api.cu`` compiled by g++ with
``-DMMX_HOST_EMU``: every kernel body runs phase by phase as a loop over thread ids, with shared
memory poisoned with NaN before each CTA.  It exposes the same C ABI as ``libmmx.so`` and is driven
here with numpy arrays.  It exists so that indexing / math of the kernels can be checked against the
oracle in a container without a GPU; it is never loaded by the product package.
"""
import ctypes as C
from ctypes import c_double as C_double, c_float as C_float
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
    hp = np.array([lr, b1, b2, eps, wd, 1 - b1 ** step, np.sqrt(1 - b2 ** step), gscale, 1 - b1, 1 - b2], np.float32)
    call("mmx_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.size, ptr(hp), None)


# ---------------------------------------------------------------------------------------------
# ConvMixer assembled from the C-ABI ops (mirrors motionmixerconv_b200/conv_mixer_model.py, numpy arrays)
# ---------------------------------------------------------------------------------------------
def conv_geometry(cfg):
    """-> (k1, pad1, k2, pad2) with pads as (top, left); k2 / pad2 None for mode_conv='once'."""
    k1 = tuple(cfg.get("conv1_kernel_shape", (1, 3)))
    p1 = cfg.get("conv1_padding")
    pad1 = ((k1[0] - 1) // 2, (k1[1] - 1) // 2) if p1 is None else tuple(p1)
    if cfg.get("mode_conv", "twice") != "twice":
        return k1, pad1, None, None
    k2 = cfg.get("conv2_kernel_shape")
    if k2 is None:
        k2 = (min(k1[1], cfg["in_nTP"]), min(k1[0], cfg["dimPosEmb"]))
    k2 = tuple(k2)
    p2 = cfg.get("conv2_padding")
    pad2 = ((k2[0] - 1) // 2, (k2[1] - 1) // 2) if p2 is None else tuple(p2)
    return k1, pad1, k2, pad2


class EmuConvMixer:
    def __init__(self, cfg, params, training=True, dropout=None):
        self.cfg = dict(cfg)
        self.p = {k: f32(v) for k, v in params.items() if np.asarray(v).dtype.kind == "f"}
        self.training = training
        self.dropout = dropout or dropout_struct()
        c = self.cfg
        self.C = c.get("conv_nChan", 1)
        self.T, self.To, self.E = c["in_nTP"], c["out_nTP"], c["dimPosEmb"]
        self.D, self.Dout = c["dimPosIn"], c["dimPosOut"]
        self.Hn = max(c.get("encoder_n_harmonic_functions", 64), 0)
        self.use_se = bool(c.get("use_se", False))
        self.use_max = bool(c.get("use_max_pooling", False))
        self.rr = self.T // c.get("r_se", 4)
        self.act = L.MMX_ACT[c.get("activation", "gelu")]
        self.twice = c.get("mode_conv", "twice") == "twice"
        self.k1, self.pad1, self.k2, self.pad2 = conv_geometry(c)
        self.bn = c.get("regularization", 0) == -1.0
        self.running = {k: np.array(v, copy=True) for k, v in params.items() if "running_" in k or "num_batches" in k}

    def _half_tables(self, i, half, src):
        t = L.MmxConvHalfParams()
        pre = "Mixer_Block.%d." % i
        ln, cv = ("LN1", "conv1") if half == 0 else ("LN2", "conv2")
        t.ln_w, t.ln_b = ptr(src[pre + ln + ".weight"]), ptr(src[pre + ln + ".bias"])
        t.conv_w, t.conv_b = ptr(src[pre + cv + ".conv.weight"]), ptr(src[pre + cv + ".conv.bias"])
        if self.use_se:
            t.se_w1, t.se_w2 = ptr(src[pre + "se.excitationBlock.0.weight"]), ptr(src[pre + "se.excitationBlock.2.weight"])
        return t

    def _half_desc(self, i, half, B):
        k, pad = (self.k1, self.pad1) if half == 0 else (self.k2, self.pad2)
        return L.MmxConvHalfDesc(B, self.C, self.T, self.E, k[0], k[1], pad[0], pad[1], self.rr, self.act, int(self.use_se),
                                 int(self.use_max), int(self.training), 2 * i + half, self.dropout)

    def _enc(self, src):
        t = L.MmxEncoderParams()
        t.freq = ptr(self.p["encoder.frequencies"]) if self.Hn > 0 else None
        t.w, t.b = ptr(src["encoder.embed_mlp.weight"]), ptr(src["encoder.embed_mlp.bias"])
        t.wc, t.bc = ptr(src["encoder.channelUpscaling.weight"]), ptr(src["encoder.channelUpscaling.bias"])
        return t

    def _head(self, src):
        t = L.MmxConvHeadParams()
        t.ln_w, t.ln_b = ptr(src["LN.weight"]), ptr(src["LN.bias"])
        t.wt, t.bt = ptr(src["conv_out.weight"]), ptr(src["conv_out.bias"])
        t.wp, t.bp = ptr(src["project_channels.weight"]), ptr(src["project_channels.bias"])
        t.wf, t.bf = ptr(src["fc_out.weight"]), ptr(src["fc_out.bias"])
        return t

    def forward(self, x):
        x = f32(x)
        B = x.shape[0]
        C, T, E = self.C, self.T, self.E
        self.x = x
        self.m = np.empty((B * T, E), np.float32)
        y = np.empty((B, C, T, E), np.float32)
        ed = L.MmxEncoderDesc(B, T, self.D, E, C, self.Hn)
        call("mmx_pose_encoder_fwd", _byref(ed), _byref(self._enc(self.p)), ptr(x), ptr(self.m), ptr(y), None)
        self.acts = [y]           # inputs of every half / tail, in execution order
        self.ops = []
        self.bn_saved, self._keep = {}, []
        for i in range(self.cfg["num_blocks"]):
            for half in ((0, 1) if self.twice else (0,)):
                out = np.empty_like(y)
                desc, tw = self._half_desc(i, half, B), self._half_tables(i, half, self.p)
                if self.bn and self.training:
                    pre = "Mixer_Block.%d.conv%d.reg." % (i, half + 1)
                    z = np.empty_like(y)
                    sums = np.zeros(2 * C, np.float64)
                    call("mmx_conv_half_bn_stats", _byref(desc), _byref(tw), ptr(self.acts[-1]), ptr(z), sums.ctypes.data, None)
                    bn = np.empty(4 * C, np.float32)
                    rm, rv = f32(self.running[pre + "running_mean"]), f32(self.running[pre + "running_var"])
                    nbt = np.array(self.running[pre + "num_batches_tracked"], dtype=np.int64).reshape(1)
                    call("mmx_bn_finalize", sums.ctypes.data, C, C_double(B * T * E), ptr(self.p[pre + "weight"]), ptr(self.p[pre + "bias"]),
                         ptr(rm), ptr(rv), nbt.ctypes.data, C_float(0.1), C_float(1e-5), ptr(bn), None)
                    assert not sums.any()
                    call("mmx_conv_half_bn_apply", _byref(desc), _byref(tw), ptr(bn), ptr(self.acts[-1]), ptr(z), ptr(out), None)
                    self.running[pre + "running_mean"], self.running[pre + "running_var"] = rm, rv
                    self.running[pre + "num_batches_tracked"] = nbt[0]
                    self.bn_saved[(i, half)] = (z, bn)
                else:
                    if self.bn:
                        pre = "Mixer_Block.%d.conv%d.reg." % (i, half + 1)
                        sc = self.p[pre + "weight"] / np.sqrt(self.running[pre + "running_var"] + 1e-5)
                        aff = f32(np.concatenate([sc, self.p[pre + "bias"] - self.running[pre + "running_mean"] * sc]))
                        self._keep.append(aff)
                        tw.bn_aff = ptr(aff)
                    call("mmx_conv_half_fwd", _byref(desc), _byref(tw), ptr(self.acts[-1]), ptr(out), None)
                self.acts.append(out)
                self.ops.append(("half", i, half))
            if not self.twice:
                out = np.empty_like(y)
                pre = "Mixer_Block.%d." % i
                s1 = ptr(self.p[pre + "se.excitationBlock.0.weight"]) if self.use_se else None
                s2 = ptr(self.p[pre + "se.excitationBlock.2.weight"]) if self.use_se else None
                call("mmx_se_tail_fwd", B, C, T, E, self.rr, int(self.use_se), int(self.use_max), s1, s2, ptr(self.acts[-1]), ptr(out), None)
                self.acts.append(out)
                self.ops.append(("tail", i, 1))
        pred = np.empty((B, self.To, self.Dout), np.float32)
        hd = L.MmxConvHeadDesc(B, C, T, self.To, E, self.Dout)
        call("mmx_conv_head_fwd", _byref(hd), _byref(self._head(self.p)), ptr(self.acts[-1]), ptr(pred), None)
        return pred

    def backward(self, dout, need_dx=True):
        dout = f32(dout)
        B = self.x.shape[0]
        C, T, E = self.C, self.T, self.E
        g = {k: np.zeros_like(v) for k, v in self.p.items()}
        d_act = np.empty((B, C, T, E), np.float32)
        hd = L.MmxConvHeadDesc(B, C, T, self.To, E, self.Dout)
        call("mmx_conv_head_bwd", _byref(hd), _byref(self._head(self.p)), _byref(self._head(g)), ptr(self.acts[-1]), ptr(dout), ptr(d_act), None)
        for n in reversed(range(len(self.ops))):
            kind, i, half = self.ops[n]
            dx = np.empty_like(d_act)
            if kind == "half" and self.bn:
                desc, tw, tg = self._half_desc(i, half, B), self._half_tables(i, half, self.p), self._half_tables(i, half, g)
                z, bn = self.bn_saved[(i, half)]
                pre = "Mixer_Block.%d.conv%d.reg." % (i, half + 1)
                gd = np.empty((B, T, 2), np.float32)
                sums = np.zeros(2 * C, np.float64)
                call("mmx_conv_half_bn_bwd1", _byref(desc), _byref(tw), _byref(tg), ptr(bn), ptr(z), ptr(d_act), ptr(gd), sums.ctypes.data, None)
                coef = np.empty(3 * C, np.float32)
                call("mmx_bn_coef", sums.ctypes.data, C, C_double(B * T * E), ptr(bn), ptr(coef), ptr(g[pre + "weight"]), ptr(g[pre + "bias"]), None)
                call("mmx_conv_half_bn_bwd2", _byref(desc), _byref(tw), _byref(tg), ptr(bn), ptr(coef), ptr(self.acts[n]), ptr(z),
                     ptr(d_act), ptr(gd), ptr(dx), None)
            elif kind == "half":
                call("mmx_conv_half_bwd", _byref(self._half_desc(i, half, B)), _byref(self._half_tables(i, half, self.p)),
                     _byref(self._half_tables(i, half, g)), ptr(self.acts[n]), ptr(d_act), ptr(dx), None)
            else:
                pre = "Mixer_Block.%d." % i
                k1, k2 = pre + "se.excitationBlock.0.weight", pre + "se.excitationBlock.2.weight"
                s = [ptr(self.p[k1]), ptr(self.p[k2]), ptr(g[k1]), ptr(g[k2])] if self.use_se else [None] * 4
                call("mmx_se_tail_bwd", B, C, T, E, self.rr, int(self.use_se), int(self.use_max), *s, ptr(self.acts[n]), ptr(d_act), ptr(dx), None)
            d_act = dx
        dm = np.empty((B * T, E), np.float32)
        dxin = np.empty_like(self.x) if need_dx else None
        ed = L.MmxEncoderDesc(B, T, self.D, E, C, self.Hn)
        call("mmx_pose_encoder_bwd", _byref(ed), _byref(self._enc(self.p)), _byref(self._enc(g)), ptr(self.x), ptr(self.m),
             ptr(d_act), ptr(dm), ptr(dxin), None)
        for k in list(g):
            if ".se2." in k or "running_" in k:          # aliases of .se. in the reference state_dict; BN buffers
                g.pop(k)
        g.pop("encoder.frequencies", None)
        return g, dxin


def _byref(s):
    return C.byref(s)
