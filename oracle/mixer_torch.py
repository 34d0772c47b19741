"""TEST / BASELINE INFRASTRUCTURE — functional torch-CPU port of the reference hot path.

A second restatement of the same arithmetic as ``oracle/mixer_np.py``, written with
``torch.nn.functional`` ops on a plain ``dict`` of weights (the reference ``state_dict`` layout), so
that it runs the *same ATen CPU kernels, multi-threaded,* that the reference modules dispatch to
(the reference's arithmetic lives in PyTorch ATen).  It is what ``bench.py`` times as
the CPU baseline (``cpu_baseline.kind = "port"`` and ``--impl reference``) on the GPU box, where
``/root/reference`` does not exist.  ``tests/test_oracle_golden.py`` pins it against the fixtures
generated from the reference itself.  Never imported by the product package.

Follows: h36m/mlp_mixer.py:30-34,37-41,87-96,138-164,325-337; h36m/conv_mixer_model.py:47-70,
129-142,268-292,428-465; conv_mixer/encoding/positional_encoder.py:79-97;
h36m/utils/utils_mixer.py:48-53; step body h36m/train_mixer_h36m.py:126,180-193.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _act(name):
    if name == "gelu":
        return F.gelu
    if name == "mish":
        return lambda x: x * torch.tanh(F.softplus(x))
    raise ValueError("Unknown activation function type: %s" % name)


def _reg(p, pre, x, cfg, training):
    r = cfg.get("regularization", 0)
    if r > 0.0:
        return F.dropout(x, r, training)
    if r == -1.0:
        return F.batch_norm(x, p[pre + ".running_mean"], p[pre + ".running_var"], p[pre + ".weight"], p[pre + ".bias"],
                            training, 0.1, 1e-5)
    return x


def _se(y, w1, w2, pool_dims, tdim, use_max):
    s = y.amax(pool_dims) if use_max else y.mean(pool_dims)
    g = torch.sigmoid(F.linear(F.relu(F.linear(s, w1)), w2))
    shp = [1] * y.dim()
    shp[0], shp[tdim] = y.shape[0], y.shape[tdim]
    return y * g.view(shp)


def mlpmixer_forward(p, cfg, x, training=True):
    act = _act(cfg.get("activation", "gelu"))
    H = cfg["hidden_dim"]
    use_se, use_max = cfg.get("use_se", False), cfg.get("use_max_pooling", False)
    y = F.linear(x, p["conv.weight"].view(H, -1), p["conv.bias"])

    def mlp(pre, v):
        v = _reg(p, pre + ".reg1", act(F.linear(v, p[pre + ".fc1.weight"], p[pre + ".fc1.bias"])), cfg, training)
        return _reg(p, pre + ".reg2", F.linear(v, p[pre + ".fc2.weight"], p[pre + ".fc2.bias"]), cfg, training)

    for i in range(cfg["num_blocks"]):
        b = "Mixer_Block.%d" % i
        z = F.layer_norm(y, (H,), p[b + ".LN1.weight"], p[b + ".LN1.bias"])
        z = mlp(b + ".mlp_block_token_mixing", z.transpose(1, 2)).transpose(1, 2)
        if use_se:
            z = _se(z, p[b + ".se.excitation.0.weight"], p[b + ".se.excitation.2.weight"], (2,), 1, use_max)
        y = y + z
        z = mlp(b + ".mlp_block_channel_mixing", F.layer_norm(y, (H,), p[b + ".LN2.weight"], p[b + ".LN2.bias"]))
        if use_se:
            z = _se(z, p[b + ".se.excitation.0.weight"], p[b + ".se.excitation.2.weight"], (2,), 1, use_max)
        y = y + z
    y = F.layer_norm(y, (H,), p["LN.weight"], p["LN.bias"])
    y = F.conv1d(y, p["conv_out.weight"], p["conv_out.bias"])
    return F.linear(y, p["fc_out.weight"], p["fc_out.bias"])


def convmixer_forward(p, cfg, x, training=True):
    act = F.gelu if cfg.get("activation", "gelu") == "gelu" else F.mish
    if cfg.get("activation", "gelu") not in ("gelu", "mish"):
        raise ValueError("Unknown activation function type: %s" % cfg["activation"])
    E, T = cfg["dimPosEmb"], cfg["in_nTP"]
    use_se, use_max = cfg.get("use_se", False), cfg.get("use_max_pooling", False)
    twice = cfg.get("mode_conv", "twice") == "twice"
    k1 = tuple(cfg.get("conv1_kernel_shape", (1, 3)))
    pad1 = cfg.get("conv1_padding") or "same"
    pad1 = tuple(pad1) if not isinstance(pad1, str) else pad1
    pad2 = cfg.get("conv2_padding") or "same"
    pad2 = tuple(pad2) if not isinstance(pad2, str) else pad2
    x = x.unsqueeze(1)
    if cfg.get("encoder_n_harmonic_functions", 64) > 0:
        emb = (x[..., None] * p["encoder.frequencies"]).view(*x.shape[:-1], -1)
        emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    else:
        emb = x
    y = F.linear(emb, p["encoder.embed_mlp.weight"], p["encoder.embed_mlp.bias"]).transpose(1, 3)
    y = F.linear(y, p["encoder.channelUpscaling.weight"], p["encoder.channelUpscaling.bias"]).transpose(1, 3)

    def convblock(pre, v, pad):
        v = act(F.conv2d(v, p[pre + ".conv.weight"], p[pre + ".conv.bias"], padding=pad))
        return _reg(p, pre + ".reg", v, cfg, training)

    for i in range(cfg["num_blocks"]):
        b = "Mixer_Block.%d" % i
        z = convblock(b + ".conv1", F.layer_norm(y, (E,), p[b + ".LN1.weight"], p[b + ".LN1.bias"]), pad1)
        if use_se:
            z = _se(z, p[b + ".se.excitationBlock.0.weight"], p[b + ".se.excitationBlock.2.weight"], (1, 3), 2, use_max)
        y = y + z
        if twice:
            z = convblock(b + ".conv2", F.layer_norm(y, (E,), p[b + ".LN2.weight"], p[b + ".LN2.bias"]), pad2)
        else:
            z = y
        if use_se:
            z = _se(z, p[b + ".se.excitationBlock.0.weight"], p[b + ".se.excitationBlock.2.weight"], (1, 3), 2, use_max)
        y = y + z
    y = F.layer_norm(y, (E,), p["LN.weight"], p["LN.bias"])
    y = F.conv2d(y.transpose(1, 2), p["conv_out.weight"], p["conv_out.bias"]).transpose(1, 2)
    y = F.conv2d(y, p["project_channels.weight"], p["project_channels.bias"]).squeeze(1)
    return F.linear(F.gelu(y), p["fc_out.weight"], p["fc_out.bias"])


def mpjpe_error(pred, gt):
    return torch.mean(torch.norm(gt.contiguous().view(-1, 3) - pred.contiguous().view(-1, 3), 2, 1))


def is_trainable(key):
    return not (key.endswith(("running_mean", "running_var", "num_batches_tracked")) or key == "encoder.frequencies"
                or ".se2." in key)


class CpuTrainer:
    """fwd -> mpjpe -> bwd -> torch.optim.Adam(lr, weight_decay=1e-5) on the host cores."""

    def __init__(self, family, cfg, params, lr=1e-3, weight_decay=1e-5, loss_scale=1.0):
        self.forward = mlpmixer_forward if family == "mlp" else convmixer_forward
        self.cfg = cfg
        self.p = {k: torch.as_tensor(v).clone() for k, v in params.items()}
        self.train_keys = [k for k in self.p if is_trainable(k) and self.p[k].dtype.is_floating_point]
        for k in self.train_keys:
            self.p[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.p[k] for k in self.train_keys], lr=lr, weight_decay=weight_decay)
        self.loss_scale = loss_scale

    def step(self, x, gt):
        self.opt.zero_grad()
        loss = mpjpe_error(self.forward(self.p, self.cfg, x, True), gt) * self.loss_scale
        loss.backward()
        self.opt.step()
        return loss


def random_params(family, cfg, seed=0):
    """Random-init weights with the reference's state_dict layout and PyTorch-default distributions,
    built WITHOUT the reference (for bench.py on the GPU box)."""
    import math
    g = torch.Generator().manual_seed(seed)
    p = {}

    def lin(name, out_f, in_f, shape=None, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        p[name + ".weight"] = ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound).view(shape or (out_f, in_f))
        if bias:
            p[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    def ln(name, n):
        p[name + ".weight"], p[name + ".bias"] = torch.ones(n), torch.zeros(n)

    def bn(name, n):
        ln(name, n)
        p[name + ".running_mean"], p[name + ".running_var"] = torch.zeros(n), torch.ones(n)
        p[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    reg = cfg.get("regularization", 0)
    if family == "mlp":
        H, T, D = cfg["hidden_dim"], cfg["seq_len"], cfg["input_size"]
        tok, ch, r = cfg["tokens_mlp_dim"], cfg["channels_mlp_dim"], cfg.get("r_se", 4)
        lin("conv", H, D, (H, 1, 1, D))
        for i in range(cfg["num_blocks"]):
            b = "Mixer_Block.%d" % i
            lin(b + ".mlp_block_token_mixing.fc1", tok, T)
            lin(b + ".mlp_block_token_mixing.fc2", T, tok)
            if reg == -1.0:
                bn(b + ".mlp_block_token_mixing.reg1", H)
                bn(b + ".mlp_block_token_mixing.reg2", H)
            lin(b + ".mlp_block_channel_mixing.fc1", ch, H)
            lin(b + ".mlp_block_channel_mixing.fc2", H, ch)
            if reg == -1.0:
                bn(b + ".mlp_block_channel_mixing.reg1", T)
                bn(b + ".mlp_block_channel_mixing.reg2", T)
            if cfg.get("use_se", False):
                lin(b + ".se.excitation.0", T // r, T, bias=False)
                lin(b + ".se.excitation.2", T, T // r, bias=False)
            ln(b + ".LN1", H)
            ln(b + ".LN2", H)
        ln("LN", H)
        lin("fc_out", cfg["num_classes"], H)
        lin("conv_out", cfg["pred_len"], T, (cfg["pred_len"], T, 1))
        return p
    E, T, D, C = cfg["dimPosEmb"], cfg["in_nTP"], cfg["dimPosIn"], cfg.get("conv_nChan", 1)
    Hn = cfg.get("encoder_n_harmonic_functions", 64)
    r = cfg.get("r_se", 4)
    if Hn > 0:
        p["encoder.frequencies"] = cfg.get("encoder_omega0", 0.1) * (2.0 ** torch.arange(Hn))
    lin("encoder.embed_mlp", E, D * 2 * Hn if Hn > 0 else D)
    lin("encoder.channelUpscaling", C, 1)
    k1 = tuple(cfg.get("conv1_kernel_shape", (1, 3)))
    k2 = cfg.get("conv2_kernel_shape") or (min(k1[1], T), min(k1[0], E))
    twice = cfg.get("mode_conv", "twice") == "twice"
    for i in range(cfg["num_blocks"]):
        b = "Mixer_Block.%d" % i
        lin(b + ".conv1.conv", C, C * k1[0] * k1[1], (C, C, k1[0], k1[1]))
        if reg == -1.0:
            bn(b + ".conv1.reg", C)
        if cfg.get("use_se", False):
            lin(b + ".se.excitationBlock.0", T // r, T, bias=False)
            lin(b + ".se.excitationBlock.2", T, T // r, bias=False)
        ln(b + ".LN1", E)
        if twice:
            lin(b + ".conv2.conv", C, C * k2[0] * k2[1], (C, C, k2[0], k2[1]))
            if reg == -1.0:
                bn(b + ".conv2.reg", C)
            if cfg.get("use_se", False):
                p[b + ".se2.excitationBlock.0.weight"] = p[b + ".se.excitationBlock.0.weight"]
                p[b + ".se2.excitationBlock.2.weight"] = p[b + ".se.excitationBlock.2.weight"]
            ln(b + ".LN2", E)
    ln("LN", E)
    lin("project_channels", 1, C, (1, C, 1, 1))
    lin("conv_out", cfg["out_nTP"], T, (cfg["out_nTP"], T, 1, 1))
    lin("fc_out", cfg["dimPosOut"], E)
    return p
