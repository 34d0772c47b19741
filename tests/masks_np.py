"""TEST INFRASTRUCTURE: numpy twins of the dropout-mask generators of the CUDA kernels (bit-exact), so that the parity tests
can hand the oracle the very masks a kernel family draws for (seed, step, site).

Generators (csrc/mmx_common.cuh, csrc/mmx_chan_tc5.cuh):
  * ``philox``            Philox4x32-N  (Salmon et al., SC'11)
  * generic MixerBlock / ConvMixer kernels: one 32-bit Philox4x32-10 word per element, counter = element / 4
  * warp-per-sequence-pair MixerBlock kernels (T=10, tok=20, H,ch <= 64): Philox4x32-7, 16 bits per element, one call per
    quad of a row pair
  * tcgen05 MixerBlock family: lowbias32 hash, 16 bits per element, one 4-hash call per 8-column chunk
Each ``*_masks`` function returns the dict of scaled keep-masks ``oracle.mixer_np`` expects.
"""
import numpy as np

U32 = np.uint32
M32 = np.uint64(0xFFFFFFFF)


def philox(c0, c1, c2, c3, k0, k1, rounds):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & M32 for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    for _ in range(rounds):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & M32
        n1 = p1 & M32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & M32
        n3 = p0 & M32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & M32, (k1 + W1) & M32
    return c0, c1, c2, c3


def thresh32(p):
    t = float(np.float32(p)) * 4294967296.0
    th = 0xFFFFFFFF if t >= 4294967295.0 else int(t + 0.5)
    return max(th, 1)


def scale_of(p):
    return np.float32(1.0) / (np.float32(1.0) - np.float32(p))


def _seed_keys(seed):
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------- generic kernels (dropout_scale / dropout_quad)
def elementwise_mask(shape, p, seed, site, step):
    """keep-scale tensor of a dropout site whose elements are numbered row-major (generic MixerBlock kernels)."""
    n = int(np.prod(shape))
    e = np.arange(n, dtype=np.uint64)
    k0, k1 = _seed_keys(seed)
    r = philox(e >> np.uint64(2), e >> np.uint64(34), np.uint64(site), np.uint64(step), k0, k1, 10)
    lane = (e & np.uint64(3)).astype(np.int64)
    v = np.choose(lane, r)
    return np.where(v >= np.uint64(thresh32(p)), scale_of(p), np.float32(0)).astype(np.float32).reshape(shape)


def quad_row_mask(rows, W, p, seed, site, step):
    """[rows][W] site whose rows are padded to quads: quad = row * ceil(W/4) + w // 4 (ConvMixer kernels, dropout_quad)."""
    W4 = (W + 3) // 4
    row = np.arange(rows, dtype=np.uint64)[:, None]
    w = np.arange(W, dtype=np.uint64)[None, :]
    quad = row * np.uint64(W4) + (w >> np.uint64(2))
    k0, k1 = _seed_keys(seed)
    r = philox(quad, quad >> np.uint64(32), np.uint64(site), np.uint64(step), k0, k1, 10)
    v = np.choose(np.broadcast_to((w & np.uint64(3)).astype(np.int64), quad.shape), r)
    return np.where(v >= np.uint64(thresh32(p)), scale_of(p), np.float32(0)).astype(np.float32)


def mlp_generic_masks(cfg, B, seed, step=0):
    p = cfg["regularization"]
    T, H, tok, ch = cfg["seq_len"], cfg["hidden_dim"], cfg["tokens_mlp_dim"], cfg["channels_mlp_dim"]
    out = {}
    for i in range(cfg["num_blocks"]):
        pre = "Mixer_Block.%d." % i
        out[pre + "mlp_block_token_mixing.reg1"] = elementwise_mask((B, H, tok), p, seed, 4 * i + 0, step)
        out[pre + "mlp_block_token_mixing.reg2"] = elementwise_mask((B, H, T), p, seed, 4 * i + 1, step)
        out[pre + "mlp_block_channel_mixing.reg1"] = elementwise_mask((B, T, ch), p, seed, 4 * i + 2, step)
        out[pre + "mlp_block_channel_mixing.reg2"] = elementwise_mask((B, T, H), p, seed, 4 * i + 3, step)
    return out


def conv_masks(cfg, B, seed, step=0):
    """ConvMixer dropout sites: one per block half, on [B, C, T, E] (site = 2 * block + half)."""
    p = cfg["regularization"]
    C, T, E = cfg.get("conv_nChan", 1), cfg["in_nTP"], cfg["dimPosEmb"]
    twice = cfg.get("mode_conv", "twice") == "twice"
    out = {}
    for i in range(cfg["num_blocks"]):
        out["Mixer_Block.%d.conv1.reg" % i] = quad_row_mask(B * C * T, E, p, seed, 2 * i, step).reshape(B, C, T, E)
        if twice:
            out["Mixer_Block.%d.conv2.reg" % i] = quad_row_mask(B * C * T, E, p, seed, 2 * i + 1, step).reshape(B, C, T, E)
    return out


# ---------------------------------------------------------------------------------- warp-per-sequence-pair kernels (dropout_rowpair)
def rowpair_mask(nseq, R, W, p, seed, site, step):
    """[nseq][R][W] site: pair = seq * ceil(R/2) + r // 2, counter = pair * ceil(W/4) + w // 4, Philox4x32-7, 16-bit fields:
    row parity selects words (c0, c1) / (c2, c3), w % 4 the half-word."""
    W4, R2 = (W + 3) // 4, (R + 1) // 2
    s = np.arange(nseq, dtype=np.uint64)[:, None, None]
    r = np.arange(R, dtype=np.uint64)[None, :, None]
    w = np.arange(W, dtype=np.uint64)[None, None, :]
    ctr = (s * np.uint64(R2) + (r >> np.uint64(1))) * np.uint64(W4) + (w >> np.uint64(2))
    k0, k1 = _seed_keys(seed)
    c = philox(ctr & M32, ctr >> np.uint64(32), np.uint64(site ^ 0x5bd1e995), np.uint64(step), k0, k1, 7)
    word = (np.uint64(2) * (r & np.uint64(1)) + ((w & np.uint64(3)) >> np.uint64(1))).astype(np.int64)
    v = np.choose(np.broadcast_to(word, ctr.shape), c)
    hi = np.broadcast_to((w & np.uint64(1)).astype(bool), ctr.shape)
    field = np.where(hi, v >> np.uint64(16), v & np.uint64(0xFFFF))
    return np.where(field >= np.uint64(thresh32(p) >> 16), scale_of(p), np.float32(0)).astype(np.float32)


def mlp_warp_masks(cfg, B, seed, step=0):
    p = cfg["regularization"]
    T, H, tok, ch = cfg["seq_len"], cfg["hidden_dim"], cfg["tokens_mlp_dim"], cfg["channels_mlp_dim"]
    out = {}
    for i in range(cfg["num_blocks"]):
        pre = "Mixer_Block.%d." % i
        out[pre + "mlp_block_token_mixing.reg1"] = rowpair_mask(B, tok, H, p, seed, 4 * i + 0, step).transpose(0, 2, 1).copy()
        out[pre + "mlp_block_token_mixing.reg2"] = rowpair_mask(B, T, H, p, seed, 4 * i + 1, step).transpose(0, 2, 1).copy()
        out[pre + "mlp_block_channel_mixing.reg1"] = rowpair_mask(B, T, ch, p, seed, 4 * i + 2, step)
        out[pre + "mlp_block_channel_mixing.reg2"] = rowpair_mask(B, T, H, p, seed, 4 * i + 3, step)
    return out


# ---------------------------------------------------------------------------------- tcgen05 family (chan::keep8)
def mix32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x21f0aaad)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x735a2d97)) & M32
    x ^= x >> np.uint64(15)
    return x


def drop_key(seed, site, step):
    lo, hi = _seed_keys(seed)
    inner = mix32((np.uint64(site) * np.uint64(0x9E3779B1) + np.uint64(step)) & M32)
    return mix32(np.uint64(lo) ^ mix32(np.uint64(hi) ^ inner))


def tc5_mask(rows, W, p, seed, site, step):
    """[rows][W] keep-scale tensor of a tcgen05-family dropout site (chunk c8 of a row = columns [8*c8, 8*c8 + 8))."""
    W8 = (W + 7) // 8
    key = drop_key(seed, site, step)
    row = np.arange(rows, dtype=np.uint64)[:, None]
    w = np.arange(W, dtype=np.uint64)[None, :]
    ctr = ((row * np.uint64(W8) + (w >> np.uint64(3))) * np.uint64(4) + ((w & np.uint64(7)) >> np.uint64(1))) & M32
    r = mix32(ctr ^ key)
    field = np.where(np.broadcast_to((w & np.uint64(1)).astype(bool), r.shape), r >> np.uint64(16), r & np.uint64(0xFFFF))
    return np.where(field >= np.uint64(thresh32(p) >> 16), scale_of(p), np.float32(0)).astype(np.float32)


def mlp_tc5_masks(cfg, B, seed, step=0):
    p = cfg["regularization"]
    T, H, tok, ch = cfg["seq_len"], cfg["hidden_dim"], cfg["tokens_mlp_dim"], cfg["channels_mlp_dim"]
    out = {}
    for i in range(cfg["num_blocks"]):
        pre = "Mixer_Block.%d." % i
        col = tc5_mask(B * H, tok + T, p, seed, 4 * i, step)          # one stream per (sequence, hidden column)
        out[pre + "mlp_block_token_mixing.reg1"] = col[:, :tok].reshape(B, H, tok).copy()
        out[pre + "mlp_block_token_mixing.reg2"] = col[:, tok:].reshape(B, H, T).copy()
        out[pre + "mlp_block_channel_mixing.reg1"] = tc5_mask(B * T, ch, p, seed, 4 * i + 2, step).reshape(B, T, ch)
        out[pre + "mlp_block_channel_mixing.reg2"] = tc5_mask(B * T, H, p, seed, 4 * i + 3, step).reshape(B, T, H)
    return out
