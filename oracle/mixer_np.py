"""TEST INFRASTRUCTURE — numpy restatement of the reference hot path.

This file is the *oracle* (checker) for the CUDA kernels.  It restates, with
hand-derived backward passes, the arithmetic of

  * ``h36m/mlp_mixer.py``            SELayer :30-34, mish :37-41, MlpBlock :87-96,
                                     MixerBlock :138-164, MlpMixer :325-337
  * ``h36m/conv_mixer_model.py``     MultiChanSELayer :47-70, ConvBlock :129-142,
                                     ConvMixerBlock :268-292, ConvMixer :428-465
  * ``conv_mixer/encoding/positional_encoder.py``  PoseEncoder.forward :79-97
  * ``h36m/utils/utils_mixer.py``    mpjpe_error :48-53
  * ``torch.optim.Adam`` (coupled L2) as called at ``h36m/train_mixer_h36m.py:63``

The reference's arithmetic lives in PyTorch ATen (third-party, not vendored in
``/root/reference``; pinned ``torch==1.9.1`` in ``requirements.txt:11``, the
container has 2.11.0).  The reference holds no golden vectors for this path, so
the oracle is pinned against the reference modules themselves, imported from
``/root/reference`` in the build container: ``tests/golden/make_golden.py``
writes the fixtures, ``tests/test_oracle_golden.py`` checks this file against
them.  Parameters are addressed by the reference's ``state_dict`` key names.

Everything is written for whole batches with numpy matmul/einsum; ``dtype`` may
be float32 (the parity mode) or float64 (to measure the fp32 noise floor).
Never imported by the product package.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf as _erf

_SQRT1_2 = 0.7071067811865476
_INV_SQRT_2PI = 0.3989422804014327
LN_EPS = 1e-5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# activations (mlp_mixer.py:37-41,78-81; conv_mixer_model.py:121-124)
# --------------------------------------------------------------------------------------
def gelu(u):
    return 0.5 * u * (1.0 + _erf(u * _SQRT1_2))


def gelu_grad(u):
    return 0.5 * (1.0 + _erf(u * _SQRT1_2)) + u * np.exp(-0.5 * u * u) * _INV_SQRT_2PI


def softplus(u):
    # F.softplus(beta=1, threshold=20)
    with np.errstate(over="ignore"):
        return np.where(u > 20.0, u, np.log1p(np.exp(np.minimum(u, 20.0))))


def mish(u):
    return u * np.tanh(softplus(u))


def mish_grad(u):
    sp = softplus(u)
    t = np.tanh(sp)
    with np.errstate(over="ignore"):
        sig = np.where(u > 20.0, 1.0, 1.0 / (1.0 + np.exp(-u)))
    return t + u * (1.0 - t * t) * sig


def _act(name):
    if name == "gelu":
        return gelu, gelu_grad
    if name == "mish":
        return mish, mish_grad
    raise ValueError("Unknown activation function type: %s" % name)


# --------------------------------------------------------------------------------------
# LayerNorm over the last dim (biased variance, eps 1e-5)
# --------------------------------------------------------------------------------------
def ln_fwd(x, g, b):
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(LN_EPS))
    xhat = xc * rstd
    return xhat * g + b, (xhat, rstd)


def ln_bwd(dy, cache, g):
    xhat, rstd = cache
    n = dy.shape[-1]
    dg = (dy * xhat).reshape(-1, n).sum(0)
    db = dy.reshape(-1, n).sum(0)
    dxh = dy * g
    dx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xhat * (dxh * xhat).mean(-1, keepdims=True))
    return dx, dg, db


# --------------------------------------------------------------------------------------
# BatchNorm (training mode: biased var to normalise, unbiased for the running stat)
# --------------------------------------------------------------------------------------
def bn_fwd(x, w, b, rm, rv, caxis, training):
    """x: any rank; channel axis ``caxis``.  Returns y, cache, (new_rm, new_rv)."""
    axes = tuple(i for i in range(x.ndim) if i != caxis)
    shp = [1] * x.ndim
    shp[caxis] = x.shape[caxis]
    if training:
        mean = x.mean(axes)
        xc = x - mean.reshape(shp)
        var = (xc * xc).mean(axes)
        n = x.size // x.shape[caxis]
        new_rm = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
        new_rv = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
    else:
        mean, var = rm, rv
        xc = x - mean.reshape(shp)
        new_rm, new_rv = rm, rv
    rstd = 1.0 / np.sqrt(var + x.dtype.type(BN_EPS))
    xhat = xc * rstd.reshape(shp)
    y = xhat * w.reshape(shp) + b.reshape(shp)
    return y, (xhat, rstd, axes, shp, training), (new_rm.astype(x.dtype), new_rv.astype(x.dtype))


def bn_bwd(dy, cache, w):
    xhat, rstd, axes, shp, training = cache
    dw = (dy * xhat).sum(axes)
    db = dy.sum(axes)
    dxh = dy * w.reshape(shp)
    if training:
        dx = rstd.reshape(shp) * (
            dxh - dxh.mean(axes).reshape(shp) - xhat * (dxh * xhat).mean(axes).reshape(shp)
        )
    else:
        dx = dxh * rstd.reshape(shp)
    return dx, dw, db


# --------------------------------------------------------------------------------------
# squeeze-excitation over frames (mlp_mixer.py:30-34, conv_mixer_model.py:57-70)
# --------------------------------------------------------------------------------------
def se_fwd(y, S1, S2, pool_axes, taxis, use_max):
    """y: [B,T,H] (pool_axes=(2,), taxis=1) or [B,C,T,E] (pool_axes=(1,3), taxis=2)."""
    if use_max:
        s = y.max(axis=pool_axes)
    else:
        s = y.mean(axis=pool_axes)
    z = s @ S1.T  # [B, r]
    a = np.maximum(z, 0)
    q = a @ S2.T  # [B, T]
    g = 1.0 / (1.0 + np.exp(-q))
    shp = [1] * y.ndim
    shp[0] = y.shape[0]
    shp[taxis] = y.shape[taxis]
    return y * g.reshape(shp), (y, s, z, a, g, shp, pool_axes, taxis, use_max)


def se_bwd(dout, cache, S1, S2):
    y, s, z, a, g, shp, pool_axes, taxis, use_max = cache
    ge = g.reshape(shp)
    dy = dout * ge
    dg = (dout * y).sum(axis=pool_axes)  # [B, T]
    dq = dg * g * (1 - g)
    dS2 = dq.T @ a  # [T, r]
    da = dq @ S2  # [B, r]
    dz = da * (z > 0)
    dS1 = dz.T @ s  # [r, T]
    ds = dz @ S1  # [B, T]
    if use_max:
        # gradient goes to the first maximum (row-major order over the pooled axes)
        ym = np.moveaxis(y, taxis, 1)  # [B, T, ...pooled]
        flat = ym.reshape(ym.shape[0], ym.shape[1], -1)
        idx = flat.argmax(-1)
        dflat = np.zeros_like(flat)
        np.put_along_axis(dflat, idx[..., None], ds[..., None], axis=-1)
        dy = dy + np.moveaxis(dflat.reshape(ym.shape), 1, taxis)
    else:
        npool = 1
        for ax in pool_axes:
            npool *= y.shape[ax]
        dy = dy + (ds / npool).reshape(shp)
    return dy, dS1, dS2


# --------------------------------------------------------------------------------------
# conv2d, stride 1, zero padding (conv_mixer_model.py:112,139)
# --------------------------------------------------------------------------------------
def same_padding(k):
    """PyTorch padding='same': left=(k-1)//2, surplus on the right/bottom."""
    tot = k - 1
    lo = tot // 2
    return lo, tot - lo


def resolve_padding(kernel, padding):
    """-> (top, bottom, left, right)"""
    if padding is None or padding == "same":
        t, b = same_padding(kernel[0])
        l, r = same_padding(kernel[1])
        return t, b, l, r
    return padding[0], padding[0], padding[1], padding[1]


def conv2d_fwd(x, w, b, pad):
    pt, pb, pl, pr = pad
    B, Ci, T, E = x.shape
    Co, _, kT, kP = w.shape
    To = T + pt + pb - kT + 1
    Eo = E + pl + pr - kP + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pt, pb), (pl, pr)))
    out = np.zeros((B, Co, To, Eo), dtype=x.dtype)
    for i in range(kT):
        for j in range(kP):
            out += np.einsum("bcte,oc->bote", xp[:, :, i:i + To, j:j + Eo], w[:, :, i, j], optimize=True)
    out += b.reshape(1, Co, 1, 1)
    return out, (xp, pad, (T, E))


def conv2d_bwd(dout, cache, w):
    xp, (pt, pb, pl, pr), (T, E) = cache
    Co, Ci, kT, kP = w.shape
    To, Eo = dout.shape[2], dout.shape[3]
    dw = np.zeros_like(w)
    dxp = np.zeros_like(xp)
    for i in range(kT):
        for j in range(kP):
            dw[:, :, i, j] = np.einsum("bote,bcte->oc", dout, xp[:, :, i:i + To, j:j + Eo], optimize=True)
            dxp[:, :, i:i + To, j:j + Eo] += np.einsum("bote,oc->bcte", dout, w[:, :, i, j], optimize=True)
    db = dout.sum((0, 2, 3))
    dx = dxp[:, :, pt:pt + T, pl:pl + E]
    return dx, dw, db


# --------------------------------------------------------------------------------------
# MPJPE (utils_mixer.py:48-53) and Adam (train_mixer_h36m.py:63)
# --------------------------------------------------------------------------------------
def mpjpe(pred, gt):
    """-> (loss, dloss/dpred).  Gradient is 0 where the joint distance is 0."""
    d = (gt - pred).reshape(-1, 3)
    nrm = np.sqrt((d * d).sum(1))
    loss = nrm.mean(dtype=pred.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(nrm[:, None] > 0, -d / nrm[:, None], 0.0) / pred.dtype.type(nrm.shape[0])
    return loss, g.reshape(pred.shape).astype(pred.dtype)


def adam_step(p, g, m, v, step, lr, wd=1e-5, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place PyTorch Adam (non-decoupled L2).  ``step`` is the 1-based step count."""
    dt = p.dtype.type
    g = g + dt(wd) * p
    m *= dt(beta1)
    m += dt(1 - beta1) * g
    v *= dt(beta2)
    v += dt(1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / dt(np.sqrt(bc2)) + dt(eps)
    p -= dt(lr / bc1) * (m / denom)


# --------------------------------------------------------------------------------------
# MlpMixer (h36m/mlp_mixer.py)
# --------------------------------------------------------------------------------------
class MlpMixerOracle:
    """Forward/backward of ``MlpMixer`` (mlp_mixer.py:239-337) on numpy arrays.

    ``cfg``: the reference constructor kwargs.  ``params``: arrays keyed by the
    reference ``state_dict`` names (BN running stats included and updated in place).
    ``masks``: optional dict ``'Mixer_Block.{i}.mlp_block_{token,channel}_mixing.reg{1,2}'``
    -> dropout keep-mask already scaled by 1/(1-p).
    """

    def __init__(self, cfg, params, dtype=np.float32):
        self.cfg = dict(cfg)
        self.dtype = dtype
        self.p = {k: (np.asarray(v).astype(dtype) if np.asarray(v).dtype.kind == "f" else np.asarray(v).copy())
                  for k, v in params.items()}
        self.act, self.act_grad = _act(cfg.get("activation", "gelu"))
        self.use_se = bool(cfg.get("use_se", False))
        self.use_max = bool(cfg.get("use_max_pooling", False))
        reg = cfg.get("regularization", 0)
        self.bn = reg == -1.0
        self.dropout = reg > 0.0
        self.nb = cfg["num_blocks"]

    # -- MlpBlock (mlp_mixer.py:87-96) acting on the last dim of ``y`` -----------------
    def _mlp_fwd(self, y, pre, training, masks):
        p = self.p
        u = y @ p[pre + ".fc1.weight"].T + p[pre + ".fc1.bias"]
        gact = self.act(u)
        c = {"y": y, "u": u}
        h = gact
        if self.bn:
            h, c["bn1"], (rm, rv) = bn_fwd(h, p[pre + ".reg1.weight"], p[pre + ".reg1.bias"],
                                           p[pre + ".reg1.running_mean"], p[pre + ".reg1.running_var"], 1, training)
            if training:
                p[pre + ".reg1.running_mean"], p[pre + ".reg1.running_var"] = rm, rv
                p[pre + ".reg1.num_batches_tracked"] = p[pre + ".reg1.num_batches_tracked"] + 1
        elif self.dropout and training and masks is not None:
            c["m1"] = masks[pre + ".reg1"].astype(self.dtype)
            h = h * c["m1"]
        c["h"] = h
        o = h @ p[pre + ".fc2.weight"].T + p[pre + ".fc2.bias"]
        if self.bn:
            o, c["bn2"], (rm, rv) = bn_fwd(o, p[pre + ".reg2.weight"], p[pre + ".reg2.bias"],
                                           p[pre + ".reg2.running_mean"], p[pre + ".reg2.running_var"], 1, training)
            if training:
                p[pre + ".reg2.running_mean"], p[pre + ".reg2.running_var"] = rm, rv
                p[pre + ".reg2.num_batches_tracked"] = p[pre + ".reg2.num_batches_tracked"] + 1
        elif self.dropout and training and masks is not None:
            c["m2"] = masks[pre + ".reg2"].astype(self.dtype)
            o = o * c["m2"]
        return o, c

    def _mlp_bwd(self, do, c, pre, grads):
        p = self.p
        if "bn2" in c:
            do, dw, db = bn_bwd(do, c["bn2"], p[pre + ".reg2.weight"])
            _acc(grads, pre + ".reg2.weight", dw)
            _acc(grads, pre + ".reg2.bias", db)
        elif "m2" in c:
            do = do * c["m2"]
        nin = do.shape[-1]
        nh = c["h"].shape[-1]
        _acc(grads, pre + ".fc2.weight", do.reshape(-1, nin).T @ c["h"].reshape(-1, nh))
        _acc(grads, pre + ".fc2.bias", do.reshape(-1, nin).sum(0))
        dh = do @ p[pre + ".fc2.weight"]
        if "bn1" in c:
            dh, dw, db = bn_bwd(dh, c["bn1"], p[pre + ".reg1.weight"])
            _acc(grads, pre + ".reg1.weight", dw)
            _acc(grads, pre + ".reg1.bias", db)
        elif "m1" in c:
            dh = dh * c["m1"]
        du = dh * self.act_grad(c["u"])
        _acc(grads, pre + ".fc1.weight", du.reshape(-1, nh).T @ c["y"].reshape(-1, nin))
        _acc(grads, pre + ".fc1.bias", du.reshape(-1, nh).sum(0))
        return du @ p[pre + ".fc1.weight"]

    def forward(self, x, training=True, masks=None):
        p = self.p
        x = np.asarray(x).astype(self.dtype)
        H = self.cfg["hidden_dim"]
        Wc = p["conv.weight"].reshape(H, -1)
        X = x @ Wc.T + p["conv.bias"]  # [B,T,H]   (mlp_mixer.py:325-327)
        self.c = {"x": x, "blocks": []}
        for i in range(self.nb):
            pre = "Mixer_Block.%d" % i
            c = {}
            n1, c["ln1"] = ln_fwd(X, p[pre + ".LN1.weight"], p[pre + ".LN1.bias"])
            yt, c["tok"] = self._mlp_fwd(np.swapaxes(n1, 1, 2), pre + ".mlp_block_token_mixing", training, masks)
            y = np.swapaxes(yt, 1, 2)
            if self.use_se:
                y, c["se1"] = se_fwd(y, p[pre + ".se.excitation.0.weight"], p[pre + ".se.excitation.2.weight"],
                                     (2,), 1, self.use_max)
            X = X + y
            n2, c["ln2"] = ln_fwd(X, p[pre + ".LN2.weight"], p[pre + ".LN2.bias"])
            y, c["ch"] = self._mlp_fwd(n2, pre + ".mlp_block_channel_mixing", training, masks)
            if self.use_se:
                y, c["se2"] = se_fwd(y, p[pre + ".se.excitation.0.weight"], p[pre + ".se.excitation.2.weight"],
                                     (2,), 1, self.use_max)
            X = X + y
            self.c["blocks"].append(c)
        z, self.c["ln"] = ln_fwd(X, p["LN.weight"], p["LN.bias"])
        Wt = p["conv_out.weight"][:, :, 0]  # [To,T]
        P = np.einsum("ot,bth->boh", Wt, z) + p["conv_out.bias"][None, :, None]
        self.c["z"], self.c["P"] = z, P
        return P @ p["fc_out.weight"].T + p["fc_out.bias"]

    def backward(self, dout):
        p = self.p
        c0 = self.c
        grads = {}
        dout = np.asarray(dout).astype(self.dtype)
        D = dout.shape[-1]
        H = self.cfg["hidden_dim"]
        _acc(grads, "fc_out.weight", dout.reshape(-1, D).T @ c0["P"].reshape(-1, H))
        _acc(grads, "fc_out.bias", dout.reshape(-1, D).sum(0))
        dP = dout @ p["fc_out.weight"]
        Wt = p["conv_out.weight"][:, :, 0]
        _acc(grads, "conv_out.weight", np.einsum("boh,bth->ot", dP, c0["z"])[:, :, None])
        _acc(grads, "conv_out.bias", dP.sum((0, 2)))
        dz = np.einsum("ot,boh->bth", Wt, dP)
        dX, dg, db = ln_bwd(dz, c0["ln"], p["LN.weight"])
        _acc(grads, "LN.weight", dg)
        _acc(grads, "LN.bias", db)
        for i in reversed(range(self.nb)):
            pre = "Mixer_Block.%d" % i
            c = c0["blocks"][i]
            dy = dX
            if self.use_se:
                dy, dS1, dS2 = se_bwd(dy, c["se2"], p[pre + ".se.excitation.0.weight"], p[pre + ".se.excitation.2.weight"])
                _acc(grads, pre + ".se.excitation.0.weight", dS1)
                _acc(grads, pre + ".se.excitation.2.weight", dS2)
            dn2 = self._mlp_bwd(dy, c["ch"], pre + ".mlp_block_channel_mixing", grads)
            d, dg, db = ln_bwd(dn2, c["ln2"], p[pre + ".LN2.weight"])
            _acc(grads, pre + ".LN2.weight", dg)
            _acc(grads, pre + ".LN2.bias", db)
            dX = dX + d
            dy = dX
            if self.use_se:
                dy, dS1, dS2 = se_bwd(dy, c["se1"], p[pre + ".se.excitation.0.weight"], p[pre + ".se.excitation.2.weight"])
                _acc(grads, pre + ".se.excitation.0.weight", dS1)
                _acc(grads, pre + ".se.excitation.2.weight", dS2)
            dn1t = self._mlp_bwd(np.swapaxes(dy, 1, 2), c["tok"], pre + ".mlp_block_token_mixing", grads)
            d, dg, db = ln_bwd(np.swapaxes(dn1t, 1, 2), c["ln1"], p[pre + ".LN1.weight"])
            _acc(grads, pre + ".LN1.weight", dg)
            _acc(grads, pre + ".LN1.bias", db)
            dX = dX + d
        x = c0["x"]
        _acc(grads, "conv.weight", (dX.reshape(-1, H).T @ x.reshape(-1, x.shape[-1])).reshape(p["conv.weight"].shape))
        _acc(grads, "conv.bias", dX.reshape(-1, H).sum(0))
        dx = dX @ p["conv.weight"].reshape(H, -1)
        return grads, dx


def _acc(grads, key, val):
    if key in grads:
        grads[key] = grads[key] + val
    else:
        grads[key] = val


# --------------------------------------------------------------------------------------
# ConvMixer (h36m/conv_mixer_model.py + conv_mixer/encoding/positional_encoder.py)
# --------------------------------------------------------------------------------------
class ConvMixerOracle:
    """Forward/backward of ``ConvMixer`` (conv_mixer_model.py:295-465).

    ``masks``: optional dict ``'Mixer_Block.{i}.conv{1,2}.reg'`` -> scaled dropout keep-mask.
    """

    def __init__(self, cfg, params, dtype=np.float32):
        cfg = dict(cfg)
        cfg.setdefault("conv_nChan", 1)
        cfg.setdefault("conv1_kernel_shape", (1, 3))
        cfg.setdefault("conv1_stride", (1, 1))
        cfg.setdefault("conv1_padding", None)
        cfg.setdefault("mode_conv", "twice")
        cfg.setdefault("conv2_kernel_shape", None)
        cfg.setdefault("conv2_stride", None)
        cfg.setdefault("conv2_padding", None)
        cfg.setdefault("activation", "gelu")
        cfg.setdefault("regularization", 0)
        cfg.setdefault("use_se", False)
        cfg.setdefault("r_se", 4)
        cfg.setdefault("use_max_pooling", False)
        cfg.setdefault("encoder_n_harmonic_functions", 64)
        cfg.setdefault("encoder_omega0", 0.1)
        self.cfg = cfg
        self.dtype = dtype
        self.p = {k: (np.asarray(v).astype(dtype) if np.asarray(v).dtype.kind == "f" else np.asarray(v).copy())
                  for k, v in params.items()}
        # the harmonic argument is ONE fp32 multiply in the reference (positional_encoder.py:86)
        if "encoder.frequencies" in params:
            self.freq32 = np.asarray(params["encoder.frequencies"]).astype(np.float32)
        self.act, self.act_grad = _act(cfg["activation"])
        self.use_se = bool(cfg["use_se"])
        self.use_max = bool(cfg["use_max_pooling"])
        reg = cfg["regularization"]
        self.bn = reg == -1.0
        self.dropout = reg > 0.0
        self.nb = cfg["num_blocks"]
        self.twice = cfg["mode_conv"] == "twice"
        if cfg["mode_conv"] not in ("once", "twice"):
            raise ValueError("mode_conv %s must be one of 'once' or 'twice'" % cfg["mode_conv"])
        k1 = tuple(cfg["conv1_kernel_shape"])
        self.pad1 = resolve_padding(k1, cfg["conv1_padding"])
        if self.twice:
            k2 = cfg["conv2_kernel_shape"]
            if k2 is None:
                k2 = (min(k1[1], cfg["in_nTP"]), min(k1[0], cfg["dimPosEmb"]))
            self.pad2 = resolve_padding(tuple(k2), cfg["conv2_padding"])

    def _convblock_fwd(self, y, pre, pad, training, masks):
        p = self.p
        z, cc = conv2d_fwd(y, p[pre + ".conv.weight"], p[pre + ".conv.bias"], pad)
        a = self.act(z)
        c = {"conv": cc, "z": z}
        if self.bn:
            a, c["bn"], (rm, rv) = bn_fwd(a, p[pre + ".reg.weight"], p[pre + ".reg.bias"],
                                          p[pre + ".reg.running_mean"], p[pre + ".reg.running_var"], 1, training)
            if training:
                p[pre + ".reg.running_mean"], p[pre + ".reg.running_var"] = rm, rv
                p[pre + ".reg.num_batches_tracked"] = p[pre + ".reg.num_batches_tracked"] + 1
        elif self.dropout and training and masks is not None:
            c["m"] = masks[pre + ".reg"].astype(self.dtype)
            a = a * c["m"]
        return a, c

    def _convblock_bwd(self, da, c, pre, grads):
        p = self.p
        if "bn" in c:
            da, dw, db = bn_bwd(da, c["bn"], p[pre + ".reg.weight"])
            _acc(grads, pre + ".reg.weight", dw)
            _acc(grads, pre + ".reg.bias", db)
        elif "m" in c:
            da = da * c["m"]
        dz = da * self.act_grad(c["z"])
        dx, dw, db = conv2d_bwd(dz, c["conv"], p[pre + ".conv.weight"])
        _acc(grads, pre + ".conv.weight", dw)
        _acc(grads, pre + ".conv.bias", db)
        return dx

    def _se_keys(self, pre):
        return pre + ".se.excitationBlock.0.weight", pre + ".se.excitationBlock.2.weight"

    def forward(self, x, training=True, masks=None):
        p = self.p
        cfg = self.cfg
        x32 = np.asarray(x).astype(np.float32)
        x = x32.astype(self.dtype)
        Hn = cfg["encoder_n_harmonic_functions"]
        c0 = {"x": x}
        if Hn > 0:
            a = (x32[..., None] * self.freq32).reshape(x.shape[0], x.shape[1], -1)  # fp32 product, index d*Hn+h
            a = a.astype(self.dtype)
            c0["sin"], c0["cos"] = np.sin(a), np.cos(a)
            emb = np.concatenate((c0["sin"], c0["cos"]), axis=-1)
        else:
            emb = x
        c0["emb"] = emb
        m = emb @ p["encoder.embed_mlp.weight"].T + p["encoder.embed_mlp.bias"]  # [B,T,E]
        c0["m"] = m
        wc = p["encoder.channelUpscaling.weight"][:, 0]
        bc = p["encoder.channelUpscaling.bias"]
        Y = m[:, None, :, :] * wc[None, :, None, None] + bc[None, :, None, None]  # [B,C,T,E]
        c0["blocks"] = []
        for i in range(self.nb):
            pre = "Mixer_Block.%d" % i
            c = {}
            n1, c["ln1"] = ln_fwd(Y, p[pre + ".LN1.weight"], p[pre + ".LN1.bias"])
            y, c["cb1"] = self._convblock_fwd(n1, pre + ".conv1", self.pad1, training, masks)
            if self.use_se:
                k0, k2 = self._se_keys(pre)
                y, c["se1"] = se_fwd(y, p[k0], p[k2], (1, 3), 2, self.use_max)
            Y = Y + y
            if self.twice:
                n2, c["ln2"] = ln_fwd(Y, p[pre + ".LN2.weight"], p[pre + ".LN2.bias"])
                y, c["cb2"] = self._convblock_fwd(n2, pre + ".conv2", self.pad2, training, masks)
            else:
                y = Y  # LN2/conv2 are Identity (conv_mixer_model.py:259-263) ...
            if self.use_se:
                k0, k2 = self._se_keys(pre)
                y, c["se2"] = se_fwd(y, p[k0], p[k2], (1, 3), 2, self.use_max)  # ... but self.se is still applied (:289)
            Y = Y + y
            c0["blocks"].append(c)
        Q, c0["ln"] = ln_fwd(Y, p["LN.weight"], p["LN.bias"])
        Wt = p["conv_out.weight"][:, :, 0, 0]  # [To,T]
        P = np.einsum("ot,bcte->bcoe", Wt, Q) + p["conv_out.bias"][None, None, :, None]
        wp = p["project_channels.weight"][0, :, 0, 0]  # [C]
        r = np.einsum("c,bcoe->boe", wp, P) + p["project_channels.bias"][0]
        c0["Q"], c0["P"], c0["r"] = Q, P, r
        gr = gelu(r)
        c0["gr"] = gr
        self.c = c0
        return gr @ p["fc_out.weight"].T + p["fc_out.bias"]

    def backward(self, dout):
        p = self.p
        cfg = self.cfg
        c0 = self.c
        grads = {}
        dout = np.asarray(dout).astype(self.dtype)
        D = dout.shape[-1]
        E = cfg["dimPosEmb"]
        _acc(grads, "fc_out.weight", dout.reshape(-1, D).T @ c0["gr"].reshape(-1, E))
        _acc(grads, "fc_out.bias", dout.reshape(-1, D).sum(0))
        dr = (dout @ p["fc_out.weight"]) * gelu_grad(c0["r"])
        wp = p["project_channels.weight"][0, :, 0, 0]
        _acc(grads, "project_channels.weight", np.einsum("boe,bcoe->c", dr, c0["P"]).reshape(1, -1, 1, 1))
        _acc(grads, "project_channels.bias", dr.sum().reshape(1))
        dP = dr[:, None, :, :] * wp[None, :, None, None]
        Wt = p["conv_out.weight"][:, :, 0, 0]
        _acc(grads, "conv_out.weight", np.einsum("bcoe,bcte->ot", dP, c0["Q"])[:, :, None, None])
        _acc(grads, "conv_out.bias", dP.sum((0, 1, 3)))
        dQ = np.einsum("ot,bcoe->bcte", Wt, dP)
        dY, dg, db = ln_bwd(dQ, c0["ln"], p["LN.weight"])
        _acc(grads, "LN.weight", dg)
        _acc(grads, "LN.bias", db)
        for i in reversed(range(self.nb)):
            pre = "Mixer_Block.%d" % i
            c = c0["blocks"][i]
            k0, k2 = self._se_keys(pre)
            dy = dY
            if self.use_se:
                dy, dS1, dS2 = se_bwd(dy, c["se2"], p[k0], p[k2])
                _acc(grads, k0, dS1)
                _acc(grads, k2, dS2)
            if self.twice:
                dn2 = self._convblock_bwd(dy, c["cb2"], pre + ".conv2", grads)
                d, dg, db = ln_bwd(dn2, c["ln2"], p[pre + ".LN2.weight"])
                _acc(grads, pre + ".LN2.weight", dg)
                _acc(grads, pre + ".LN2.bias", db)
            else:
                d = dy
            dY = dY + d
            dy = dY
            if self.use_se:
                dy, dS1, dS2 = se_bwd(dy, c["se1"], p[k0], p[k2])
                _acc(grads, k0, dS1)
                _acc(grads, k2, dS2)
            dn1 = self._convblock_bwd(dy, c["cb1"], pre + ".conv1", grads)
            d, dg, db = ln_bwd(dn1, c["ln1"], p[pre + ".LN1.weight"])
            _acc(grads, pre + ".LN1.weight", dg)
            _acc(grads, pre + ".LN1.bias", db)
            dY = dY + d
        wc = p["encoder.channelUpscaling.weight"][:, 0]
        _acc(grads, "encoder.channelUpscaling.weight", np.einsum("bcte,bte->c", dY, c0["m"])[:, None])
        _acc(grads, "encoder.channelUpscaling.bias", dY.sum((0, 2, 3)))
        dm = np.einsum("bcte,c->bte", dY, wc)
        emb = c0["emb"]
        _acc(grads, "encoder.embed_mlp.weight", dm.reshape(-1, E).T @ emb.reshape(-1, emb.shape[-1]))
        _acc(grads, "encoder.embed_mlp.bias", dm.reshape(-1, E).sum(0))
        demb = dm @ p["encoder.embed_mlp.weight"]
        Hn = cfg["encoder_n_harmonic_functions"]
        if Hn > 0:
            n = demb.shape[-1] // 2
            da = demb[..., :n] * c0["cos"] - demb[..., n:] * c0["sin"]
            x = c0["x"]
            dx = (da.reshape(x.shape[0], x.shape[1], x.shape[2], Hn) * self.freq32.astype(self.dtype)).sum(-1)
        else:
            dx = demb
        return grads, dx


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------------------
def trainable_keys(params):
    """state_dict keys that are nn.Parameters (not buffers / aliases), in state_dict order."""
    out = []
    for k in params:
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")) or k == "encoder.frequencies":
            continue
        if ".se2." in k:  # alias of .se. (conv_mixer_model.py:257)
            continue
        out.append(k)
    return out


def train_steps(oracle, x, gt, n_steps, lr=1e-3, wd=1e-5, loss_scale=1.0):
    """Full-batch fwd -> MPJPE -> bwd -> Adam for ``n_steps``; returns the list of losses."""
    keys = trainable_keys(oracle.p)
    m = {k: np.zeros_like(oracle.p[k]) for k in keys}
    v = {k: np.zeros_like(oracle.p[k]) for k in keys}
    losses = []
    for s in range(1, n_steps + 1):
        pred = oracle.forward(x, training=True)
        loss, dpred = mpjpe(pred, gt.astype(oracle.dtype))
        losses.append(float(loss) * loss_scale)
        grads, _ = oracle.backward(dpred * oracle.dtype(loss_scale))
        for k in keys:
            adam_step(oracle.p[k], grads[k].astype(oracle.dtype), m[k], v[k], s, lr, wd)
        for k in list(oracle.p):
            if ".se2." in k:
                oracle.p[k] = oracle.p[k.replace(".se2.", ".se.")]
    return losses
