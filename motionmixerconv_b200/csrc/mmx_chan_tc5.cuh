// Channel half of a MixerBlock on the Blackwell tensor cores (tcgen05 + TMEM + bulk-copy engine), sm_100a.
//
//     y = x1 + SE(reg2(fc2(reg1(act(fc1(LN2(x1)))))))          reference: h36m/mlp_mixer.py:157-164 (MixerBlock.forward, second
//     half), MlpBlock :87-96, SELayer :30-34; restated in oracle/mixer_np.py (MlpMixerOracle.forward / backward).
//
// Work unit: a tile of 4 * (32 / T) whole sequences = one 128-row MMA tile.  Row r of the tile is TMEM lane r; the two warps
// that may touch a lane quarter (warp % 4) split the row's 8-column chunks between them, so LayerNorm, the activation,
// dropout, the squeeze sum and the LayerNorm backward are per-thread loops over half a row plus one shared-memory exchange
// of the partial sums.  The T frames of a sequence sit in T consecutive lanes of one warp: the SE excitation is a handful of
// shuffles.  Lanes 32/T*T .. 31 of every warp are padding rows (all-zero operands).
//
// Contractions: tcgen05.mma kind::f16 on bf16 operands with fp32 accumulation in TMEM.  Every fp32 operand x is split as
// x = hi + lo (+ <= 2^-18 |x|), hi = bf16(x), lo = bf16(x - hi), and a product is issued as three MMAs hi*hi + lo*hi + hi*lo
// ("bf16x3"): measured 5e-6 .. 2e-5 on predictions / gradients of the whole model against the fp64 oracle.
// Operands live in shared memory in the 16-bit PANEL layout
//     element (row r, col c)  ->  plane + (c / 8) * (R * 16) + r * 16 + (c % 8) * 2        (hi plane, lo plane)
// which the tensor core reads in both orientations (SWIZZLE_NONE canonical layouts, checked by tools/micro/umma_layout_probe_bf16):
//   K-major  (rows = M/N index, cols = K): SBO = 128, LBO = R*16     -> forward GEMMs   D = A W^T
//   MN-major (rows = K index, cols = M/N): SBO = R*16, LBO = 128     -> weight gradients dW = dY^T A  (K = the tile's rows)
//                                                                       and data gradients dA = dY W with the UNtransposed weight
// so one copy of each activation / weight serves the forward, the data-gradient and the weight-gradient product.
// LN2's affine is folded into fc1 (W1' = W1 * gamma2, b1' = b1 + W1 beta2): the A operand is the plain normalised row and
// dgamma2, dbeta2, dW1 are derived at flush time from ONE accumulated product  Wt = dU^T xhat.  Bias gradients ride along as a
// column of ones in the B operand of the weight-gradient products.  dW1 / dW2 accumulate in TMEM across the CTA's whole
// persistent loop and are flushed once (no shared-memory accumulators, no locks).  Rows a thread needs again later in the
// tile (the residual input, xhat) are parked in spare TMEM columns, so every per-row loop is a rolled loop over chunks
// (small code: the first version, fully unrolled over register-resident rows, was 300 KB of SASS and instruction-fetch bound).
// Activations enter and leave through the bulk-copy engine (cp.async.bulk, SASS UBLKCP) with mbarrier completion.
#pragma once
#include "mmx_common.cuh"
#include "mmx_tc5.cuh"

namespace mmx {
namespace chan {

using namespace tc5;

#ifndef MMX_CHAN_HALVES
#define MMX_CHAN_HALVES 4
#endif
constexpr int kHalves = MMX_CHAN_HALVES;        // warps per TMEM lane quarter (they split a row's 8-column chunks)
constexpr int kThreadsChan = 128 * kHalves;
constexpr int kMaxRR = 4;                      // SE bottleneck width served (T // r_se)

struct ChanArgs {
    const float* x1;               // [B*T, H]   input of the channel half
    const float* dy;               // [B*T, H]   upstream gradient (backward)
    float* out;                    // forward: y; backward: dx1
    const float *ln_g, *ln_b, *w1, *b1, *w2, *b2, *se1, *se2;
    float *g_ln_g, *g_ln_b, *g_w1, *g_b1, *g_w2, *g_b2, *g_se1, *g_se2;
    int B, T, H, ch, rr;           // rr = SE hidden width (0: no SE)
    int site_base;                 // dropout site of the block's first Dropout; the channel MLP uses site_base + 2 / + 3
    Dropout dr;
    int* abort_count;              // device counter, incremented if a pipeline wait times out (tests assert it stays 0)
};

// ------------------------------------------------------------------------------------------ dropout masks of the tcgen05 family
// A dropout site is a [rows][W] tensor; chunk c8 of row `row` covers its columns [8*c8, 8*c8+8).  Four counter-based 32-bit
// hashes (lowbias32 finaliser over counter ^ key) give 16 random bits per element; an element is kept iff its field >=
// thresh16.  The key mixes (seed, site, step).  tests/tc5_masks.py is the numpy twin (bit-exact; pinned by a GPU test), which
// lets the parity tests hand the very same masks to the oracle.
MMX_HD uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}
MMX_HD uint32_t drop_key(uint32_t seed_lo, uint32_t seed_hi, uint32_t site, uint32_t step) {
    return mix32(seed_lo ^ mix32(seed_hi ^ mix32(site * 0x9E3779B1u + step)));
}
// bit j of the result = keep element 8*c8 + j
MMX_HD uint32_t keep8(uint32_t key, uint32_t thresh16, uint32_t row, uint32_t W8, uint32_t c8) {
    const uint32_t base = (row * W8 + c8) * 4u;
    uint32_t bits = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t r = mix32((base + i) ^ key);
        bits |= ((r & 0xffffu) >= thresh16 ? 1u : 0u) << (2 * i);
        bits |= ((r >> 16) >= thresh16 ? 1u : 0u) << (2 * i + 1);
    }
    return bits;
}

// ------------------------------------------------------------------------------------------ small device helpers
MMX_D uint32_t pack_bf16x2(float lo_elem, float hi_elem) {   // lo_elem -> bits [0,16) (lower address)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// 8 fp32 values -> 8 bf16 "hi" (16 bytes) + 8 bf16 "lo" (16 bytes)
MMX_D void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// instruction descriptor: kind::f16, bf16 x bf16 -> fp32
MMX_HD uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
MMX_D void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16-bit panel operand, rows = M/N index, cols = K index; k0 = first column of this K=16 step
MMX_D uint64_t dk(uint32_t plane, uint32_t panel_bytes, int k0) { return smem_desc(plane + (uint32_t)(k0 >> 3) * panel_bytes, panel_bytes, 128u); }
// 16-bit panel operand, rows = K index, cols = M/N index; r0 = first row of this K=16 step
MMX_D uint64_t dmn(uint32_t plane, uint32_t panel_bytes, int r0) { return smem_desc(plane + (uint32_t)r0 * 16u, 128u, panel_bytes); }

// D (+)= A B^T with split operands: hi*hi + lo*hi + hi*lo.  A: buffer (planes a_hi, a_lo, panel bytes a_ps), B likewise.
// A_MN / B_MN: orientation of each operand; ksteps K=16 steps.  Issued by ONE thread.
template <int A_MN, int B_MN>
MMX_D void gemm3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_ps, uint32_t b_hi, uint32_t b_lo, uint32_t b_ps, int N,
                 int ksteps, bool accumulate_first) {
    const uint32_t id = idesc_bf16(128, N, A_MN, B_MN);
#pragma unroll 1
    for (int s = 0; s < ksteps; ++s) {
        const int k0 = 16 * s;
        const uint64_t ah = A_MN ? dmn(a_hi, a_ps, k0) : dk(a_hi, a_ps, k0);
        const uint64_t al = A_MN ? dmn(a_lo, a_ps, k0) : dk(a_lo, a_ps, k0);
        const uint64_t bh = B_MN ? dmn(b_hi, b_ps, k0) : dk(b_hi, b_ps, k0);
        const uint64_t bl = B_MN ? dmn(b_lo, b_ps, k0) : dk(b_lo, b_ps, k0);
        mma_f16(d_tmem, ah, bh, id, (s > 0 || accumulate_first) ? 1u : 0u);
        mma_f16(d_tmem, al, bh, id, 1u);
        mma_f16(d_tmem, ah, bl, id, 1u);
    }
}

// ------------------------------------------------------------------------------------------ shared-memory plan
template <int KP>
struct Plan {
    static constexpr uint32_t PS = 128 * 16;                 // activation panel: 128 rows x 16 B
    static constexpr uint32_t PLANE = (KP / 8) * PS;
    static constexpr uint32_t BUF = 2 * PLANE;               // hi plane + lo plane
    static constexpr uint32_t WPS = KP * 16;                 // weight panel: KP rows x 16 B
    static constexpr uint32_t WPLANE = (KP / 8) * WPS;
    static constexpr uint32_t WBUF = 2 * WPLANE;
    static constexpr uint32_t SMALL = (4 * KP + 2 * 32 * kMaxRR + 4 * kHalves * 128 + 64) * 4;   // c1f, c2, gamma2, spare | se1, se2 | exchange | barriers
};

struct Geo {
    int spw, rpw, seq_per_tile, tile_rows, pitch;
};
MMX_HD Geo make_geo(int T, int H, int vec) {
    Geo g;
    g.spw = 32 / T;
    g.rpw = g.spw * T;
    g.seq_per_tile = 4 * g.spw;
    g.tile_rows = 4 * g.rpw;
    g.pitch = vec == 4 ? H + 4 : H;
    return g;
}
template <int KP>
MMX_HD size_t chan_smem_bytes(int T, int H, int vec, bool bwd) {
    const Geo g = make_geo(T, H, vec);
    size_t stage = (size_t)g.tile_rows * g.pitch * 4 + 64;    // + slack: chunk loads of the last row may run past the row
    stage = (stage + 127) / 128 * 128;
    size_t buf = Plan<KP>::BUF;
    if (buf < stage) buf = stage;                             // the X region doubles as the output staging tile
    return 1024 + (bwd ? 2 : 1) * buf + 2 * Plan<KP>::WBUF + (bwd ? 2 : 1) * stage + Plan<KP>::SMALL;
}

// bulk copy of a tile of activation rows global -> shared (issued by warp 0), completion on `bar`
template <int VEC>
MMX_D void stage_in(float* S, const float* g, size_t row0, int nrows, int H, int pitch, uint64_t* bar, int lane) {
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)nrows * H * 4u);
    __syncwarp();
    if (VEC == 2) {
        if (lane == 0) bulk_g2s(S, g + row0 * H, (uint32_t)nrows * H * 4u, bar);
    } else {
        for (int r = lane; r < nrows; r += 32) bulk_g2s(S + (size_t)r * pitch, g + (row0 + r) * H, (uint32_t)H * 4u, bar);
    }
}
template <int VEC>
MMX_D void stage_out(float* g, const float* S, size_t row0, int nrows, int H, int pitch, int lane) {
    if (VEC == 2) {
        if (lane == 0) bulk_s2g(g + row0 * H, S, (uint32_t)nrows * H * 4u);
    } else {
        for (int r = lane; r < nrows; r += 32) bulk_s2g(g + (row0 + r) * H, S + (size_t)r * pitch, (uint32_t)H * 4u);
    }
    bulk_commit();
}

// 8 consecutive columns [k0, k0+8) of a staged row; columns >= H read as zero (H % VEC == 0)
template <int VEC>
MMX_D void ld8(const float* srow, int k0, int H, float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j += VEC) {
        if (k0 + j < H) {
            if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(srow + k0 + j); v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w; }
            else { const float2 t = *reinterpret_cast<const float2*>(srow + k0 + j); v[j] = t.x; v[j + 1] = t.y; }
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[j + i] = 0.0f;
        }
    }
}
template <int VEC>
MMX_D void st8(float* srow, int k0, int H, const float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j += VEC) {
        if (k0 + j < H) {
            if (VEC == 4) *reinterpret_cast<float4*>(srow + k0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            else *reinterpret_cast<float2*>(srow + k0 + j) = make_float2(v[j], v[j + 1]);
        }
    }
}
MMX_D void ld8s(const float* p, float (&v)[8]) {   // 8 floats from a 16-byte aligned shared array
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

MMX_D void put_chunk(uint8_t* buf, uint32_t plane_bytes, uint32_t ps, int row, int c8, const float (&t)[8]) {
    uint4 hi, lo;
    split8(t, hi, lo);
    *reinterpret_cast<uint4*>(buf + c8 * ps + row * 16) = hi;
    *reinterpret_cast<uint4*>(buf + plane_bytes + c8 * ps + row * 16) = lo;
}

// stage one weight matrix W[rows][cols] (row-major, optional per-column scale) into a weight buffer (panel layout, KP x KP,
// zero padded; the buffer must have been zeroed)
template <int KP>
MMX_D void stage_weight(uint8_t* wbuf, const float* W, int rows, int cols, const float* colscale, int tid) {
    if ((cols & 1) == 0) {
        const int n2 = rows * cols / 2;
        for (int i = tid; i < n2; i += kThreadsChan) {
            const int e = 2 * i, r = e / cols, c = e - r * cols;
            float2 v = *reinterpret_cast<const float2*>(W + e);
            if (colscale) { v.x *= colscale[c]; v.y *= colscale[c + 1]; }
            const uint32_t h = pack_bf16x2(v.x, v.y);
            const uint32_t l = pack_bf16x2(v.x - __uint_as_float(h << 16), v.y - __uint_as_float(h & 0xffff0000u));
            const uint32_t off = (uint32_t)(c >> 3) * Plan<KP>::WPS + (uint32_t)r * 16u + (uint32_t)(c & 7) * 2u;
            *reinterpret_cast<uint32_t*>(wbuf + off) = h;
            *reinterpret_cast<uint32_t*>(wbuf + Plan<KP>::WPLANE + off) = l;
        }
    } else {
        for (int i = tid; i < rows * cols; i += kThreadsChan) {
            const int r = i / cols, c = i - r * cols;
            float v = W[i];
            if (colscale) v *= colscale[c];
            const uint32_t h = pack_bf16x2(v, 0.0f) & 0xffffu;
            const uint32_t l = pack_bf16x2(v - __uint_as_float(h << 16), 0.0f) & 0xffffu;
            const uint32_t off = (uint32_t)(c >> 3) * Plan<KP>::WPS + (uint32_t)r * 16u + (uint32_t)(c & 7) * 2u;
            *reinterpret_cast<uint16_t*>(wbuf + off) = (uint16_t)h;
            *reinterpret_cast<uint16_t*>(wbuf + Plan<KP>::WPLANE + off) = (uint16_t)l;
        }
    }
}

// per-CTA constants: weights (LN2 affine folded into fc1), biases, SE weights.  Contains CTA barriers.
template <int KP>
MMX_D void prologue(const ChanArgs& a, uint8_t* w1b, uint8_t* w2b, float* c1f, float* c2, float* gam, float* se1, float* se2, int tid) {
    const int H = a.H, ch = a.ch, T = a.T, rr = a.rr;
    for (int i = tid; i < (int)(2 * Plan<KP>::WBUF / 16); i += kThreadsChan) reinterpret_cast<uint4*>(w1b)[i] = make_uint4(0, 0, 0, 0);   // w1b, w2b contiguous
    for (int c = tid; c < KP; c += kThreadsChan) {
        c2[c] = c < H ? a.b2[c] : 0.0f;
        gam[c] = c < H ? a.ln_g[c] : 0.0f;
        c1f[c] = 0.0f;
    }
    for (int i = tid; i < 32 * kMaxRR; i += kThreadsChan) {
        se1[i] = (rr > 0 && i < rr * T) ? a.se1[i] : 0.0f;
        se2[i] = (rr > 0 && i < rr * T) ? a.se2[i] : 0.0f;
    }
    __syncthreads();
    stage_weight<KP>(w1b, a.w1, ch, H, a.ln_g, tid);
    stage_weight<KP>(w2b, a.w2, H, ch, nullptr, tid);
    // b1'[c] = b1[c] + sum_h W1[c][h] beta2[h]: one warp per row, lanes over h
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = warp; c < ch; c += kThreadsChan / 32) {
        float s = 0.0f;
        for (int h = lane; h < H; h += 32) s = fmaf(a.w1[(size_t)c * H + h], a.ln_b[h], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) c1f[c] = s + a.b1[c];
    }
}

// SE excitation of the sequence this lane belongs to (lanes seq_base .. seq_base+T-1 hold the T squeeze values)
struct SeOut { float gate; float z[kMaxRR]; };
MMX_D SeOut se_excite(float s, int t, int seq_base, int T, int rr, const float* se1, const float* se2) {
    SeOut o;
#pragma unroll
    for (int k = 0; k < kMaxRR; ++k) o.z[k] = 0.0f;
    for (int tt = 0; tt < T; ++tt) {
        const float v = __shfl_sync(0xffffffffu, s, seq_base + tt);
#pragma unroll
        for (int k = 0; k < kMaxRR; ++k)
            if (k < rr) o.z[k] = fmaf(se1[k * T + tt], v, o.z[k]);
    }
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxRR; ++k)
        if (k < rr) q = fmaf(se2[t * rr + k], fmaxf(o.z[k], 0.0f), q);
    o.gate = sigmoidf_(q);
    return o;
}

// sum the two halves' partial sums of a row (both halves get the totals).  Contains a CTA barrier.
MMX_D void row_exchange(float* ex, int half, int prow, float& a, float& b) {
    if (kHalves == 1) { __syncthreads(); return; }
    ex[(0 * kHalves + half) * 128 + prow] = a;
    ex[(1 * kHalves + half) * 128 + prow] = b;
    __syncthreads();
    a = 0.0f; b = 0.0f;
#pragma unroll
    for (int i = 0; i < kHalves; ++i) { a += ex[(0 * kHalves + i) * 128 + prow]; b += ex[(1 * kHalves + i) * 128 + prow]; }
}

// common carve-up of the dynamic shared memory
template <int KP>
struct Carve {
    uint8_t *bufX, *bufY, *w1b, *w2b;
    float *S1, *S2, *c1f, *c2, *gam, *se1, *se2, *ex;
    uint64_t* bars;
    uint32_t* tslot;
    volatile int* abortf;
    uint32_t buf_bytes;
    MMX_D Carve(uint8_t* raw, const Geo& g, bool bwd) {
        using P = Plan<KP>;
        uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
        const uint32_t stage_bytes = ((uint32_t)g.tile_rows * g.pitch * 4u + 64u + 127u) / 128u * 128u;
        buf_bytes = P::BUF > stage_bytes ? P::BUF : stage_bytes;
        bufX = sm;
        bufY = bwd ? bufX + buf_bytes : bufX;
        w1b = bufY + buf_bytes;
        w2b = w1b + P::WBUF;
        S1 = reinterpret_cast<float*>(w2b + P::WBUF);
        S2 = bwd ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(S1) + stage_bytes) : S1;
        c1f = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(S2) + stage_bytes);
        c2 = c1f + KP;
        gam = c2 + KP;
        se1 = gam + 2 * KP;
        se2 = se1 + 32 * kMaxRR;
        ex = se2 + 32 * kMaxRR;
        bars = reinterpret_cast<uint64_t*>(ex + 4 * kHalves * 128);
        tslot = reinterpret_cast<uint32_t*>(bars + 4);
        abortf = reinterpret_cast<volatile int*>(tslot + 1);
    }
};

// ==========================================================================================
// forward
// ==========================================================================================
template <int ACT, int KP, int VEC>
__global__ void __launch_bounds__(kThreadsChan) chan_fwd_kernel(const ChanArgs a) {
    using P = Plan<KP>;
    extern __shared__ uint8_t smem_raw[];
    const Geo g = make_geo(a.T, a.H, VEC);
    Carve<KP> cv(smem_raw, g, false);
    uint8_t* bufX = cv.bufX;
    float* S = cv.S1;
    uint64_t* bars = cv.bars;
    volatile int* abortf = cv.abortf;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qtr = warp & 3, half = warp >> 2;
    const int prow = qtr * 32 + lane;
    const int H = a.H, ch = a.ch, T = a.T, rr = a.rr;
    const Dropout dr = resolve_dropout(a.dr);
    const uint32_t th16 = dr.thresh >> 16;
    const uint32_t key2 = drop_key(dr.seed_lo, dr.seed_hi, a.site_base + 2, dr.step), key3 = drop_key(dr.seed_lo, dr.seed_hi, a.site_base + 3, dr.step);
    constexpr int TM_COLS = 3 * KP <= 256 ? 256 : 512;
    constexpr int NCH = KP / 8;

    if (tid == 0) {
        mbar_init(&bars[0], 1);   // X1 tile landed
        mbar_init(&bars[1], 1);   // MMA group done
        *abortf = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TM_COLS>(cv.tslot);
    pdl_launch_dependents();
    const int ntiles = (a.B + g.seq_per_tile - 1) / g.seq_per_tile;
    auto tile_nrows = [&](int tile) { return min(g.seq_per_tile, a.B - tile * g.seq_per_tile) * a.T; };
    prologue<KP>(a, cv.w1b, cv.w2b, cv.c1f, cv.c2, cv.gam, cv.se1, cv.se2, tid);   // parameters only: overlaps the previous kernel's tail
    pdl_wait();                // the previous kernel (the token half) has completed: x1 is readable
    if (warp == 0 && (int)blockIdx.x < ntiles)
        stage_in<VEC>(S, a.x1, (size_t)blockIdx.x * g.tile_rows, tile_nrows(blockIdx.x), a.H, g.pitch, &bars[0], lane);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *cv.tslot;
    const uint32_t tU = tmem, tY = tmem + KP, tXR = tmem + 2 * KP;
    const uint32_t xB = smem_u32(bufX), w1B = smem_u32(cv.w1b), w2B = smem_u32(cv.w2b);

    const bool lane_ok = lane < g.rpw;
    const int t = lane_ok ? lane % T : 0;
    const int seq_base = lane_ok ? (lane / T) * T : 0;
    const int drow = qtr * g.rpw + (lane_ok ? lane : 0);
    const int nchH = (H + 7) >> 3;
    const uint32_t ch8 = (uint32_t)(ch + 7) >> 3, H8 = (uint32_t)nchH;
    uint32_t ph_in = 0, ph_mma = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nrows = tile_nrows(tile);
        const bool valid = lane_ok && drow < nrows;
        const uint32_t grow = (uint32_t)((size_t)tile * g.tile_rows + drow);
        const float* srow = S + (size_t)drow * g.pitch;

        // ---------------- P0: LayerNorm statistics (shifted one-pass), xhat -> operand X, raw row -> TMEM
        mbar_wait(&bars[0], ph_in, abortf);
        ph_in ^= 1;
        float mean, rstd;
        {
            const float c0 = valid ? srow[0] : 0.0f;
            float s = 0.0f, ss = 0.0f;
            if (valid) {
#pragma unroll 1
                for (int c8 = half; c8 < nchH; c8 += kHalves) {
                    float v[8];
                    ld8<VEC>(srow, 8 * c8, H, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dv = 8 * c8 + j < H ? v[j] - c0 : 0.0f;
                        s += dv;
                        ss = fmaf(dv, dv, ss);
                    }
                }
            }
            if (warp == 0) bulk_wait_read0();          // the previous tile's output (staged in the X region) has left shared memory
            row_exchange(cv.ex, half, prow, s, ss);
            const float ms = s / (float)H;
            mean = c0 + ms;
            rstd = 1.0f / sqrtf(fmaxf(ss / (float)H - ms * ms, 0.0f) + 1e-5f);
        }
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float v[8], xh[8];
            ld8<VEC>(srow, 8 * c8, valid ? H : 0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) xh[j] = (valid && 8 * c8 + j < H) ? (v[j] - mean) * rstd : 0.0f;
            put_chunk(bufX, P::PLANE, P::PS, prow, c8, xh);
            tmem_st8(tmem_addr(tXR, qtr, 8 * c8), v);
        }
        tmem_wait_st();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();       // operand X complete; S fully consumed
        if (warp == 0) {
            const int next = tile + gridDim.x;
            if (next < ntiles) stage_in<VEC>(S, a.x1, (size_t)next * g.tile_rows, tile_nrows(next), H, g.pitch, &bars[0], lane);
        }
        if (tid == 0) {
            tc_fence_after();
            gemm3<0, 0>(tU, xB, xB + P::PLANE, P::PS, w1B, w1B + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        // ---------------- E1: U -> G = reg1(act(U + b1')) -> operand X
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float u[8], b[8];
            tmem_ld8(tmem_addr(tU, qtr, 8 * c8), u);
            ld8s(cv.c1f + 8 * c8, b);
            const uint32_t kb = th16 ? keep8(key2, th16, grow, ch8, c8) : 0xffu;
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float gv = act_fwd<ACT>(u[j] + b[j]);
                gv = (kb >> j) & 1u ? gv * dr.scale : 0.0f;
                u[j] = (valid && 8 * c8 + j < ch) ? gv : 0.0f;
            }
            put_chunk(bufX, P::PLANE, P::PS, prow, c8, u);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            gemm3<0, 0>(tY, xB, xB + P::PLANE, P::PS, w2B, w2B + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        // ---------------- E2: Y2 -> reg2 -> SE -> + residual -> staged output row
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
        float ssum = 0.0f, dummy = 0.0f;
#pragma unroll 1
        for (int c8 = half; c8 < nchH; c8 += kHalves) {
            float u[8], b[8];
            tmem_ld8(tmem_addr(tY, qtr, 8 * c8), u);
            ld8s(cv.c2 + 8 * c8, b);
            const uint32_t kb = th16 ? keep8(key3, th16, grow, H8, c8) : 0xffu;
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float yv = u[j] + b[j];
                yv = (kb >> j) & 1u ? yv * dr.scale : 0.0f;
                yv = (valid && 8 * c8 + j < H) ? yv : 0.0f;
                u[j] = yv;
                ssum += yv;
            }
            tmem_st8(tmem_addr(tY, qtr, 8 * c8), u);
        }
        tmem_wait_st();
        row_exchange(cv.ex, half, prow, ssum, dummy);
        float gate = 1.0f;
        if (rr > 0) gate = se_excite(ssum / (float)H, t, seq_base, T, rr, cv.se1, cv.se2).gate;
        {
            float* orow = reinterpret_cast<float*>(bufX) + (size_t)drow * g.pitch;
#pragma unroll 1
            for (int c8 = half; c8 < nchH; c8 += kHalves) {
                float y[8], x[8];
                tmem_ld8(tmem_addr(tY, qtr, 8 * c8), y);
                tmem_ld8(tmem_addr(tXR, qtr, 8 * c8), x);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = fmaf(y[j], gate, x[j]);
                if (valid) st8<VEC>(orow, 8 * c8, H, y);
            }
        }
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
        if (warp == 0) stage_out<VEC>(a.out, reinterpret_cast<const float*>(bufX), (size_t)tile * g.tile_rows, nrows, H, g.pitch, lane);
    }
    if (warp == 0) bulk_wait_all0();
    if (tid == 0 && *abortf) atomicAdd(a.abort_count, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(tmem);
}

// ==========================================================================================
// backward (forward recomputed from x1)
// ==========================================================================================
template <int ACT, int KP, int VEC>
__global__ void __launch_bounds__(kThreadsChan) chan_bwd_kernel(const ChanArgs a) {
    using P = Plan<KP>;
    extern __shared__ uint8_t smem_raw[];
    const Geo g = make_geo(a.T, a.H, VEC);
    Carve<KP> cv(smem_raw, g, true);
    uint8_t* bufX = cv.bufX;                  // xhat -> dY2 -> xhat again -> output staging
    uint8_t* bufY = cv.bufY;                  // G2 -> dU2
    float* S1 = cv.S1;                        // x1 tile
    float* S2 = cv.S2;                        // dy tile
    uint64_t* bars = cv.bars;
    volatile int* abortf = cv.abortf;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qtr = warp & 3, half = warp >> 2;
    const int prow = qtr * 32 + lane;
    const int H = a.H, ch = a.ch, T = a.T, rr = a.rr;
    const Dropout dr = resolve_dropout(a.dr);
    const uint32_t th16 = dr.thresh >> 16;
    const uint32_t key2 = drop_key(dr.seed_lo, dr.seed_hi, a.site_base + 2, dr.step), key3 = drop_key(dr.seed_lo, dr.seed_hi, a.site_base + 3, dr.step);
    constexpr int TM_COLS = 512;
    static_assert(5 * KP <= 512, "TMEM: five KP-column regions");
    constexpr int NCH = KP / 8;

    if (tid == 0) {
        mbar_init(&bars[0], 1);   // x1 tile landed
        mbar_init(&bars[1], 1);   // MMA group done
        mbar_init(&bars[2], 1);   // dy tile landed
        *abortf = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TM_COLS>(cv.tslot);
    pdl_launch_dependents();
    const int ntiles = (a.B + g.seq_per_tile - 1) / g.seq_per_tile;
    auto tile_nrows = [&](int tile) { return min(g.seq_per_tile, a.B - tile * g.seq_per_tile) * a.T; };
    prologue<KP>(a, cv.w1b, cv.w2b, cv.c1f, cv.c2, cv.gam, cv.se1, cv.se2, tid);   // parameters only: overlaps the previous kernel's tail
    pdl_wait();                // the previous kernel has completed: x1 / dy are readable, the gradient buffers may be added to
    if (warp == 0 && (int)blockIdx.x < ntiles) {
        stage_in<VEC>(S1, a.x1, (size_t)blockIdx.x * g.tile_rows, tile_nrows(blockIdx.x), a.H, g.pitch, &bars[0], lane);
        stage_in<VEC>(S2, a.dy, (size_t)blockIdx.x * g.tile_rows, tile_nrows(blockIdx.x), a.H, g.pitch, &bars[2], lane);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *cv.tslot;
    // U2 / d xhat | Y2 / dG2 | xhat (parked) | Wt = dU^T xhat | dW2
    const uint32_t tU = tmem, tY = tmem + KP, tXH = tmem + 2 * KP, tDW1 = tmem + 3 * KP, tDW2 = tmem + 4 * KP;
    const uint32_t xB = smem_u32(bufX), yB = smem_u32(bufY), w1B = smem_u32(cv.w1b), w2B = smem_u32(cv.w2b);

    const bool lane_ok = lane < g.rpw;
    const int t = lane_ok ? lane % T : 0;
    const int seq_base = lane_ok ? (lane / T) * T : 0;
    const int drow = qtr * g.rpw + (lane_ok ? lane : 0);
    const int nchH = (H + 7) >> 3;
    const uint32_t ch8 = (uint32_t)(ch + 7) >> 3, H8 = (uint32_t)nchH;
    uint32_t ph_x = 0, ph_dy = 0, ph_mma = 0;
    bool first = true;
    float gS1[kMaxRR], gS2[kMaxRR];
#pragma unroll
    for (int k = 0; k < kMaxRR; ++k) gS1[k] = gS2[k] = 0.0f;


    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nrows = tile_nrows(tile);
        const bool valid = lane_ok && drow < nrows;
        const uint32_t grow = (uint32_t)((size_t)tile * g.tile_rows + drow);
        const int next = tile + gridDim.x;
        const float* srow = S1 + (size_t)drow * g.pitch;
        const float* drw = S2 + (size_t)drow * g.pitch;

        // ---------------- P0: xhat = LN2(x1) without affine (+ ones column at H) -> operand X and TMEM
        mbar_wait(&bars[0], ph_x, abortf);
        ph_x ^= 1;
        float mean, rstd;
        {
            const float c0 = valid ? srow[0] : 0.0f;
            float s = 0.0f, ss = 0.0f;
            if (valid) {
#pragma unroll 1
                for (int c8 = half; c8 < nchH; c8 += kHalves) {
                    float v[8];
                    ld8<VEC>(srow, 8 * c8, H, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dv = 8 * c8 + j < H ? v[j] - c0 : 0.0f;
                        s += dv;
                        ss = fmaf(dv, dv, ss);
                    }
                }
            }
            if (warp == 0) bulk_wait_read0();
            row_exchange(cv.ex, half, prow, s, ss);
            const float ms = s / (float)H;
            mean = c0 + ms;
            rstd = 1.0f / sqrtf(fmaxf(ss / (float)H - ms * ms, 0.0f) + 1e-5f);
        }
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float v[8];
            ld8<VEC>(srow, 8 * c8, valid ? H : 0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = 8 * c8 + j;
                v[j] = !valid ? 0.0f : (k < H ? (v[j] - mean) * rstd : (k == H ? 1.0f : 0.0f));
            }
            put_chunk(bufX, P::PLANE, P::PS, prow, c8, v);
            tmem_st8(tmem_addr(tXH, qtr, 8 * c8), v);
        }
        tmem_wait_st();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();   // S1 consumed
        if (warp == 0 && next < ntiles) stage_in<VEC>(S1, a.x1, (size_t)next * g.tile_rows, tile_nrows(next), H, g.pitch, &bars[0], lane);
        if (tid == 0) {
            tc_fence_after();
            gemm3<0, 0>(tU, xB, xB + P::PLANE, P::PS, w1B, w1B + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        // ---------------- E1: G2 = reg1(act(U2 + b1')) -> operand Y (ones column at ch); TMEM keeps reg1'(.) * act'(U2 + b1')
        // in place of U2 (what the backward needs of it), so the activation is evaluated once per element and tile
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float u[8], b[8], dact[8];
            tmem_ld8(tmem_addr(tU, qtr, 8 * c8), u);
            ld8s(cv.c1f + 8 * c8, b);
            const uint32_t kb = th16 ? keep8(key2, th16, grow, ch8, c8) : 0xffu;
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * c8 + j;
                float gv;
                const float da = act_fwd_grad<ACT>(u[j] + b[j], &gv);
                const float ks = (kb >> j) & 1u ? dr.scale : 0.0f;
                dact[j] = (valid && c < ch) ? da * ks : 0.0f;
                u[j] = !valid ? 0.0f : (c < ch ? gv * ks : (c == ch ? 1.0f : 0.0f));
            }
            put_chunk(bufY, P::PLANE, P::PS, prow, c8, u);
            tmem_st8(tmem_addr(tU, qtr, 8 * c8), dact);
        }
        tmem_wait_st();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            gemm3<0, 0>(tY, yB, yB + P::PLANE, P::PS, w2B, w2B + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        // ---------------- E2: y2, SE forward + backward, dY2 -> operand X
        mbar_wait(&bars[2], ph_dy, abortf);
        ph_dy ^= 1;
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
        {
            float ssum = 0.0f, dgate = 0.0f;
            unsigned long long keepbits = 0ull;
            int i = 0;
#pragma unroll 1
            for (int c8 = half; c8 < nchH; c8 += kHalves, ++i) {
                float u[8], b[8], dyv[8];
                tmem_ld8(tmem_addr(tY, qtr, 8 * c8), u);
                ld8s(cv.c2 + 8 * c8, b);
                ld8<VEC>(drw, 8 * c8, valid ? H : 0, dyv);
                const uint32_t kb = th16 ? keep8(key3, th16, grow, H8, c8) : 0xffu;
                keepbits |= (unsigned long long)kb << (8 * i);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float yv = u[j] + b[j];
                    yv = (kb >> j) & 1u ? yv * dr.scale : 0.0f;
                    yv = (valid && 8 * c8 + j < H) ? yv : 0.0f;
                    ssum += yv;
                    dgate = fmaf(dyv[j], yv, dgate);
                }
            }
            row_exchange(cv.ex, half, prow, ssum, dgate);
            float gate = 1.0f, dsq = 0.0f;
            if (rr > 0) {
                const float sq = ssum / (float)H;
                const SeOut se = se_excite(sq, t, seq_base, T, rr, cv.se1, cv.se2);
                gate = se.gate;
                const float dq = valid ? dgate * gate * (1.0f - gate) : 0.0f;
                float da[kMaxRR];
#pragma unroll
                for (int k = 0; k < kMaxRR; ++k) da[k] = 0.0f;
                for (int tt = 0; tt < T; ++tt) {
                    const float dqt = __shfl_sync(0xffffffffu, dq, seq_base + tt);
#pragma unroll
                    for (int k = 0; k < kMaxRR; ++k)
                        if (k < rr) da[k] = fmaf(dqt, cv.se2[tt * rr + k], da[k]);
                }
#pragma unroll
                for (int k = 0; k < kMaxRR; ++k)
                    if (k < rr) {
                        const float dz = se.z[k] > 0.0f ? da[k] : 0.0f;
                        dsq = fmaf(dz, cv.se1[k * T + t], dsq);
                        if (valid && half == 0) {
                            gS2[k] = fmaf(dq, fmaxf(se.z[k], 0.0f), gS2[k]);
                            gS1[k] = fmaf(dz, sq, gS1[k]);
                        }
                    }
                dsq /= (float)H;
            }
            // dY2 = reg2'( dy * gate + dsq )
            i = 0;
#pragma unroll 1
            for (int c8 = half; c8 < NCH; c8 += kHalves, ++i) {
                float dyv[8];
                ld8<VEC>(drw, 8 * c8, valid ? H : 0, dyv);
                const uint32_t kb = (uint32_t)(keepbits >> (8 * i)) & 0xffu;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float d = fmaf(dyv[j], gate, dsq);
                    d = (kb >> j) & 1u ? d * dr.scale : 0.0f;
                    dyv[j] = (valid && 8 * c8 + j < H) ? d : 0.0f;
                }
                put_chunk(bufX, P::PLANE, P::PS, prow, c8, dyv);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // dG2 = dY2 W2          (A = X K-major, B = W2 [H rows][ch cols] read MN-major: K = h)
            gemm3<0, 1>(tY, xB, xB + P::PLANE, P::PS, w2B, w2B + P::WPLANE, P::WPS, KP, KP / 16, false);
            // dW2[h][c] += sum_r dY2[r][h] G2[r][c]   (both MN-major, K = the 128 rows); column ch of G2 is all ones -> db2
            gemm3<1, 1>(tDW2, xB, xB + P::PLANE, P::PS, yB, yB + P::PLANE, P::PS, KP, 128 / 16, !first);
            mma_commit(&bars[1]);
        }
        // ---------------- E3: dU2 = dG2 * (reg1' act')(stored in TMEM by E1) -> operand Y; xhat -> operand X again
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float u[8], dg[8], xh[8];
            tmem_ld8(tmem_addr(tU, qtr, 8 * c8), u);          // reg1' * act' (stored by E1)
            tmem_ld8(tmem_addr(tY, qtr, 8 * c8), dg);
            tmem_ld8(tmem_addr(tXH, qtr, 8 * c8), xh);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] *= dg[j];
            put_chunk(bufY, P::PLANE, P::PS, prow, c8, u);
            put_chunk(bufX, P::PLANE, P::PS, prow, c8, xh);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // d xhat = dU2 W1'      (A = Y K-major, B = W1' [ch rows][H cols] read MN-major: K = c)
            gemm3<0, 1>(tU, yB, yB + P::PLANE, P::PS, w1B, w1B + P::WPLANE, P::WPS, KP, KP / 16, false);
            // Wt[c][h] += sum_r dU2[r][c] xhat[r][h]; column H of xhat is all ones -> db1'
            gemm3<1, 1>(tDW1, yB, yB + P::PLANE, P::PS, xB, xB + P::PLANE, P::PS, KP, 128 / 16, !first);
            mma_commit(&bars[1]);
        }
        first = false;
        // ---------------- E4: LayerNorm backward + residual -> staged dx1 row
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
        {
            float m1 = 0.0f, m2 = 0.0f;
#pragma unroll 1
            for (int c8 = half; c8 < nchH; c8 += kHalves) {
                float u[8], xh[8];
                tmem_ld8(tmem_addr(tU, qtr, 8 * c8), u);
                tmem_ld8(tmem_addr(tXH, qtr, 8 * c8), xh);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dv = 8 * c8 + j < H ? u[j] : 0.0f;
                    m1 += dv;
                    m2 = fmaf(dv, xh[j], m2);
                }
            }
            row_exchange(cv.ex, half, prow, m1, m2);
            m1 /= (float)H;
            m2 /= (float)H;
            float* orow = reinterpret_cast<float*>(bufX) + (size_t)drow * g.pitch;
#pragma unroll 1
            for (int c8 = half; c8 < nchH; c8 += kHalves) {
                float u[8], xh[8], dyv[8];
                tmem_ld8(tmem_addr(tU, qtr, 8 * c8), u);
                tmem_ld8(tmem_addr(tXH, qtr, 8 * c8), xh);
                ld8<VEC>(drw, 8 * c8, valid ? H : 0, dyv);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = fmaf(rstd, u[j] - m1 - xh[j] * m2, dyv[j]);
                if (valid) st8<VEC>(orow, 8 * c8, H, u);
            }
        }
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            stage_out<VEC>(a.out, reinterpret_cast<const float*>(bufX), (size_t)tile * g.tile_rows, nrows, H, g.pitch, lane);
            if (next < ntiles) stage_in<VEC>(S2, a.dy, (size_t)next * g.tile_rows, tile_nrows(next), H, g.pitch, &bars[2], lane);
        }
    }
    if (warp == 0) bulk_wait_all0();
    __syncthreads();

    // ---------------- flush: Wt, dW2 (TMEM, lane = output row) -> global gradients
    // lane c (< ch) holds row c of Wt [ch][H | ones]; lane h (< H) holds row h of dW2 [H][ch | ones]
    if (!first) {
        float* stg = reinterpret_cast<float*>(bufX);          // [KP][KP+1] staging for lane-contiguous REDs
        constexpr int SP = KP + 1;
        tc_fence_after();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float u[8];
            tmem_ld8(tmem_addr(tDW1, qtr, 8 * c8), u);
            tmem_wait_ld();
            if (prow < KP)
#pragma unroll
                for (int j = 0; j < 8; ++j) stg[prow * SP + 8 * c8 + j] = u[j];
        }
        __syncthreads();
        for (int i = tid; i < ch * H; i += kThreadsChan) {
            const int c = i / H, h = i - c * H;
            red_add(a.g_w1 + i, stg[c * SP + h] * cv.gam[h]);
        }
        for (int h = tid; h < H; h += kThreadsChan) {
            float sg = 0.0f, sb = 0.0f;
#pragma unroll 8
            for (int c = 0; c < ch; ++c) {
                const float w = a.w1[(size_t)c * H + h];
                sg = fmaf(stg[c * SP + h], w, sg);
                sb = fmaf(stg[c * SP + H], w, sb);
            }
            red_add(a.g_ln_g + h, sg);
            red_add(a.g_ln_b + h, sb);
        }
        for (int c = tid; c < ch; c += kThreadsChan) red_add(a.g_b1 + c, stg[c * SP + H]);
        __syncthreads();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float u[8];
            tmem_ld8(tmem_addr(tDW2, qtr, 8 * c8), u);
            tmem_wait_ld();
            if (prow < KP)
#pragma unroll
                for (int j = 0; j < 8; ++j) stg[prow * SP + 8 * c8 + j] = u[j];
        }
        __syncthreads();
        for (int i = tid; i < H * ch; i += kThreadsChan) {
            const int h = i / ch, c = i - h * ch;
            red_add(a.g_w2 + i, stg[h * SP + c]);
        }
        for (int h = tid; h < H; h += kThreadsChan) red_add(a.g_b2 + h, stg[h * SP + ch]);
        if (rr > 0) {
            __syncthreads();
            float* acc = stg;                                  // [2][rr*T]
            for (int i = tid; i < 2 * rr * T; i += kThreadsChan) acc[i] = 0.0f;
            __syncthreads();
            if (lane_ok && half == 0)
                for (int k = 0; k < rr; ++k) {
                    atomicAdd(acc + k * T + t, gS1[k]);              // dS1[k][t]
                    atomicAdd(acc + rr * T + t * rr + k, gS2[k]);    // dS2[t][k]
                }
            __syncthreads();
            for (int i = tid; i < rr * T; i += kThreadsChan) {
                red_add(a.g_se1 + i, acc[i]);
                red_add(a.g_se2 + i, acc[rr * T + i]);
            }
        }
    }
    if (tid == 0 && *abortf) atomicAdd(a.abort_count, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(tmem);
}

}  // namespace chan
}  // namespace mmx
