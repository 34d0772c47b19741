// Channel half of a MixerBlock on the Blackwell tensor cores (tcgen05 + TMEM + bulk-copy engine), sm_100a.
//
//     y = x1 + SE(reg2(fc2(reg1(act(fc1(LN2(x1)))))))          reference: h36m/mlp_mixer.py:157-164 (MixerBlock.forward, second
//     half), MlpBlock :87-96, SELayer :30-34; restated in oracle/mixer_np.py (MlpMixerOracle.forward / backward).
//
// Work unit: a tile of 4 * (32 / T) whole sequences = one 128-row MMA tile, one thread per row (thread <-> TMEM lane): LayerNorm,
// the activation, dropout, the squeeze sum and the LayerNorm backward are thread-local loops over the row; the T frames of a
// sequence sit in T consecutive lanes of one warp, so the SE excitation is a handful of shuffles.  Lanes 32/T*T .. 31 of every
// warp are padding rows (all-zero operands).
//
// Contractions: tcgen05.mma kind::f16 on bf16 operands with fp32 accumulation in TMEM.  Every fp32 operand x is split as
// x = hi + lo (+ <= 2^-18 |x|), hi = bf16(x), lo = bf16(x - hi), and a product is issued as three MMAs hi*hi + lo*hi + hi*lo
// ("bf16x3"): products carry ~2^-17 relative error, two orders of magnitude inside the 2e-3 bar of the reduced-precision mode.
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
// persistent loop and are flushed once (no shared-memory accumulators, no locks).
// Activations enter and leave through the bulk-copy engine (cp.async.bulk, SASS UBLKCP) with mbarrier completion.
#pragma once
#include "mmx_common.cuh"
#include "mmx_tc5.cuh"

namespace mmx {
namespace chan {

using namespace tc5;

constexpr int kThreadsChan = 128;
constexpr int kMaxRR = 4;          // SE bottleneck width served (T // r_se)

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
MMX_D void split1(float v, uint16_t& hi, uint16_t& lo) {
    const uint32_t h = pack_bf16x2(v, 0.0f) & 0xffffu;
    hi = (uint16_t)h;
    lo = (uint16_t)(pack_bf16x2(v - __uint_as_float(h << 16), 0.0f) & 0xffffu);
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

// D (+)= A B^T with split operands: hi*hi + lo*hi + hi*lo.  A: activation buffer (planes a_hi, a_lo, panel bytes a_ps),
// B: buffer (b_hi, b_lo, b_ps).  a_mn / b_mn: orientation of each operand; ksteps K=16 steps.
template <int A_MN, int B_MN>
MMX_D void gemm3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_ps, uint32_t b_hi, uint32_t b_lo, uint32_t b_ps, int N,
                 int ksteps, bool accumulate_first) {
    const uint32_t id = idesc_bf16(128, N, A_MN, B_MN);
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

// keep-scales of 8 consecutive columns [8*c8, 8*c8+8) of row `grow` of a [rows][W] dropout site: one Philox4x32-7 call,
// 16 random bits per element (the masks only need to be uncorrelated).  Identical in forward and backward.
MMX_D void drop8(const Dropout& d, uint32_t site, uint32_t grow, uint32_t W8, uint32_t c8, float (&ks)[8]) {
    const uint64_t ctr = (uint64_t)grow * W8 + c8;
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = site ^ 0x2545f491u, c3 = d.step, k0 = d.seed_lo, k1 = d.seed_hi;
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += W0; k1 += W1;
    }
    const uint32_t th = d.thresh >> 16;
    ks[0] = (c0 & 0xffffu) >= th ? d.scale : 0.0f; ks[1] = (c0 >> 16) >= th ? d.scale : 0.0f;
    ks[2] = (c1 & 0xffffu) >= th ? d.scale : 0.0f; ks[3] = (c1 >> 16) >= th ? d.scale : 0.0f;
    ks[4] = (c2 & 0xffffu) >= th ? d.scale : 0.0f; ks[5] = (c2 >> 16) >= th ? d.scale : 0.0f;
    ks[6] = (c3 & 0xffffu) >= th ? d.scale : 0.0f; ks[7] = (c3 >> 16) >= th ? d.scale : 0.0f;
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
    static constexpr uint32_t SMALL = (4 * KP + 2 * 32 * kMaxRR + 64) * 4;   // c1f, c2, gamma2, spare | se1, se2 | barriers, slots
};

// geometry of a tile
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
    size_t stage = (size_t)g.tile_rows * g.pitch * 4;
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

template <int KP, int VEC>
MMX_D void load_row(const float* srow, int H, float (&x)[KP]) {
#pragma unroll
    for (int k = 0; k < KP; k += VEC) {
        if (k < H) {
            if (VEC == 4) { const float4 v = *reinterpret_cast<const float4*>(srow + k); x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w; }
            else { const float2 v = *reinterpret_cast<const float2*>(srow + k); x[k] = v.x; x[k + 1] = v.y; }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) x[k + j] = 0.0f;
        }
    }
}
template <int KP, int VEC>
MMX_D void store_row(float* srow, int H, const float (&x)[KP]) {
#pragma unroll
    for (int k = 0; k < KP; k += VEC) {
        if (k < H) {
            if (VEC == 4) *reinterpret_cast<float4*>(srow + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
            else *reinterpret_cast<float2*>(srow + k) = make_float2(x[k], x[k + 1]);
        }
    }
}

// write one row of an operand buffer: values v[KP] (already including the ones column / zero padding) -> hi and lo planes
template <int KP>
MMX_D void put_row(uint8_t* buf, int row, const float (&v)[KP]) {
#pragma unroll
    for (int c8 = 0; c8 < KP / 8; ++c8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = v[8 * c8 + j];
        uint4 hi, lo;
        split8(t, hi, lo);
        *reinterpret_cast<uint4*>(buf + c8 * Plan<KP>::PS + row * 16) = hi;
        *reinterpret_cast<uint4*>(buf + Plan<KP>::PLANE + c8 * Plan<KP>::PS + row * 16) = lo;
    }
}
MMX_D void put_chunk(uint8_t* buf, uint32_t plane_bytes, uint32_t ps, int row, int c8, const float (&t)[8]) {
    uint4 hi, lo;
    split8(t, hi, lo);
    *reinterpret_cast<uint4*>(buf + c8 * ps + row * 16) = hi;
    *reinterpret_cast<uint4*>(buf + plane_bytes + c8 * ps + row * 16) = lo;
}

// stage one weight matrix W[rows][cols] (row-major, optional per-column scale) into a weight buffer (panel layout, KP x KP, zero padded)
template <int KP>
MMX_D void stage_weight(uint8_t* wbuf, const float* W, int rows, int cols, const float* colscale, int tid) {
    for (int i = tid; i < (int)(Plan<KP>::WBUF / 16); i += kThreadsChan) reinterpret_cast<uint4*>(wbuf)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < rows * cols; i += kThreadsChan) {
        const int r = i / cols, c = i - r * cols;
        float v = W[i];
        if (colscale) v *= colscale[c];
        uint16_t hi, lo;
        split1(v, hi, lo);
        const uint32_t off = (uint32_t)(c >> 3) * Plan<KP>::WPS + (uint32_t)r * 16u + (uint32_t)(c & 7) * 2u;
        *reinterpret_cast<uint16_t*>(wbuf + off) = hi;
        *reinterpret_cast<uint16_t*>(wbuf + Plan<KP>::WPLANE + off) = lo;
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

// ==========================================================================================
// forward
// ==========================================================================================
template <int ACT, int KP, int VEC>
__global__ void __launch_bounds__(kThreadsChan) chan_fwd_kernel(const ChanArgs a) {
    using P = Plan<KP>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Geo g = make_geo(a.T, a.H, VEC);
    const uint32_t stage_bytes = ((uint32_t)g.tile_rows * g.pitch * 4u + 127u) / 128u * 128u;
    const uint32_t buf_bytes = P::BUF > stage_bytes ? P::BUF : stage_bytes;
    uint8_t* bufX = sm;
    uint8_t* w1b = bufX + buf_bytes;
    uint8_t* w2b = w1b + P::WBUF;
    float* S = reinterpret_cast<float*>(w2b + P::WBUF);
    float* c1f = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(S) + stage_bytes);
    float* c2 = c1f + KP;
    float* se1 = c2 + 3 * KP;
    float* se2 = se1 + 32 * kMaxRR;
    uint64_t* bars = reinterpret_cast<uint64_t*>(se2 + 32 * kMaxRR);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* abortf = reinterpret_cast<volatile int*>(tslot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, ch = a.ch, T = a.T, rr = a.rr;
    const Dropout dr = resolve_dropout(a.dr);
    constexpr int TM_COLS = 2 * KP <= 128 ? 128 : (2 * KP <= 256 ? 256 : 512);

    // ---------------- prologue: barriers, TMEM, weights (LN2 affine folded into fc1)
    if (tid == 0) {
        mbar_init(&bars[0], 1);   // X1 tile landed
        mbar_init(&bars[1], 1);   // MMA group done
        *abortf = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TM_COLS>(tslot);
    stage_weight<KP>(w1b, a.w1, ch, H, a.ln_g, tid);
    stage_weight<KP>(w2b, a.w2, H, ch, nullptr, tid);
    for (int c = tid; c < KP; c += kThreadsChan) {
        float s = 0.0f;
        if (c < ch) {
            s = a.b1[c];
            for (int h = 0; h < H; ++h) s = fmaf(a.w1[(size_t)c * H + h], a.ln_b[h], s);
        }
        c1f[c] = s;
        c2[c] = c < H ? a.b2[c] : 0.0f;
    }
    for (int i = tid; i < 32 * kMaxRR; i += kThreadsChan) {
        se1[i] = (rr > 0 && i < rr * T) ? a.se1[i] : 0.0f;
        se2[i] = (rr > 0 && i < rr * T) ? a.se2[i] : 0.0f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tU = tmem, tY = tmem + KP;
    const uint32_t xB = smem_u32(bufX), w1B = smem_u32(w1b), w2B = smem_u32(w2b);

    const int ntiles = (a.B + g.seq_per_tile - 1) / g.seq_per_tile;
    const bool lane_ok = lane < g.rpw;
    const int t = lane_ok ? lane % T : 0;
    const int seq_base = lane_ok ? (lane / T) * T : 0;
    const int drow = warp * g.rpw + (lane_ok ? lane : 0);
    uint32_t ph_in = 0, ph_mma = 0;
    bool store_pending = false;

    auto tile_nrows = [&](int tile) {
        const int nseq = min(g.seq_per_tile, a.B - tile * g.seq_per_tile);
        return nseq * T;
    };
    if (warp == 0 && (int)blockIdx.x < ntiles)
        stage_in<VEC>(S, a.x1, (size_t)blockIdx.x * g.tile_rows, tile_nrows(blockIdx.x), H, g.pitch, &bars[0], lane);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nrows = tile_nrows(tile);
        const bool valid = lane_ok && drow < nrows;
        const uint32_t grow = (uint32_t)((size_t)tile * g.tile_rows + drow);

        // ---------------- P0: LayerNorm of the row -> operand X (hi/lo planes)
        mbar_wait(&bars[0], ph_in, abortf);
        ph_in ^= 1;
        float x[KP];
        load_row<KP, VEC>(S + (size_t)drow * g.pitch, valid ? H : 0, x);
        float mean = 0.0f, rstd = 0.0f;
        {
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < KP; ++k) s += x[k];
            mean = s / (float)H;
            float ss = 0.0f;
#pragma unroll
            for (int k = 0; k < KP; ++k) { const float dv = k < H ? x[k] - mean : 0.0f; ss = fmaf(dv, dv, ss); }
            rstd = 1.0f / sqrtf(ss / (float)H + 1e-5f);
        }
        if (store_pending) {   // the previous tile's output (staged in the X region) must have left shared memory
            if (warp == 0) bulk_wait_read0();
            store_pending = false;
        }
        __syncthreads();       // S fully read; X region free
        if (warp == 0) {
            const int next = tile + gridDim.x;
            if (next < ntiles) stage_in<VEC>(S, a.x1, (size_t)next * g.tile_rows, tile_nrows(next), H, g.pitch, &bars[0], lane);
        }
        {
            float v[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) v[k] = !valid ? 0.0f : (k < H ? (x[k] - mean) * rstd : 0.0f);
            put_row<KP>(bufX, tid, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            gemm3<0, 0>(tU, xB, xB + P::PLANE, P::PS, w1B, w1B + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        // ---------------- E1: U -> G = reg1(act(U + b1')) -> operand X
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c8 = 0; c8 < KP / 8; ++c8) {
            float u[8], ks[8];
            tmem_ld8(tmem_addr(tU, warp, 8 * c8), u);
            tmem_wait_ld();
            if (dr.thresh) drop8(dr, a.site_base + 2, grow, (uint32_t)(ch + 7) >> 3, c8, ks);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * c8 + j;
                float gv = act_fwd<ACT>(u[j] + c1f[c]);
                if (dr.thresh) gv *= ks[j];
                u[j] = (!valid || c >= ch) ? 0.0f : gv;
            }
            put_chunk(bufX, P::PLANE, P::PS, tid, c8, u);
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
        float y[KP];
        float ssum = 0.0f;
#pragma unroll
        for (int c8 = 0; c8 < KP / 8; ++c8) {
            float u[8], ks[8];
            tmem_ld8(tmem_addr(tY, warp, 8 * c8), u);
            tmem_wait_ld();
            if (dr.thresh) drop8(dr, a.site_base + 3, grow, (uint32_t)(H + 7) >> 3, c8, ks);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int h = 8 * c8 + j;
                float yv = u[j] + c2[h];
                if (dr.thresh) yv *= ks[j];
                yv = h < H ? yv : 0.0f;
                y[h] = yv;
                ssum += yv;
            }
        }
        float gate = 1.0f;
        if (rr > 0) gate = se_excite(valid ? ssum / (float)H : 0.0f, t, seq_base, T, rr, se1, se2).gate;
#pragma unroll
        for (int k = 0; k < KP; ++k) y[k] = fmaf(y[k], gate, x[k]);
        tc_fence_before();
        if (valid) store_row<KP, VEC>(reinterpret_cast<float*>(bufX) + (size_t)drow * g.pitch, H, y);
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            stage_out<VEC>(a.out, reinterpret_cast<const float*>(bufX), (size_t)tile * g.tile_rows, nrows, H, g.pitch, lane);
            store_pending = true;
        }
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
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Geo g = make_geo(a.T, a.H, VEC);
    const uint32_t stage_bytes = ((uint32_t)g.tile_rows * g.pitch * 4u + 127u) / 128u * 128u;
    const uint32_t buf_bytes = P::BUF > stage_bytes ? P::BUF : stage_bytes;
    uint8_t* bufX = sm;                       // N2 -> dY2 -> N2 again -> output staging
    uint8_t* bufY = bufX + buf_bytes;         // G2 -> dU2
    uint8_t* w1b = bufY + buf_bytes;
    uint8_t* w2b = w1b + P::WBUF;
    float* S1 = reinterpret_cast<float*>(w2b + P::WBUF);                                  // x1 tile
    float* S2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(S1) + stage_bytes);   // dy tile
    float* c1f = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(S2) + stage_bytes);
    float* c2 = c1f + KP;
    float* gam = c2 + KP;
    float* se1 = gam + 2 * KP;
    float* se2 = se1 + 32 * kMaxRR;
    uint64_t* bars = reinterpret_cast<uint64_t*>(se2 + 32 * kMaxRR);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* abortf = reinterpret_cast<volatile int*>(tslot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.H, ch = a.ch, T = a.T, rr = a.rr;
    const Dropout dr = resolve_dropout(a.dr);
    constexpr int TM_COLS = 4 * KP <= 256 ? 256 : 512;
    static_assert(4 * KP <= 512, "TMEM: four KP-column accumulators");

    if (tid == 0) {
        mbar_init(&bars[0], 1);   // x1 tile landed
        mbar_init(&bars[1], 1);   // MMA group done
        mbar_init(&bars[2], 1);   // dy tile landed
        *abortf = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TM_COLS>(tslot);
    stage_weight<KP>(w1b, a.w1, ch, H, a.ln_g, tid);
    stage_weight<KP>(w2b, a.w2, H, ch, nullptr, tid);
    for (int c = tid; c < KP; c += kThreadsChan) {
        float s = 0.0f;
        if (c < ch) {
            s = a.b1[c];
            for (int h = 0; h < H; ++h) s = fmaf(a.w1[(size_t)c * H + h], a.ln_b[h], s);
        }
        c1f[c] = s;
        c2[c] = c < H ? a.b2[c] : 0.0f;
        gam[c] = c < H ? a.ln_g[c] : 0.0f;
    }
    for (int i = tid; i < 32 * kMaxRR; i += kThreadsChan) {
        se1[i] = (rr > 0 && i < rr * T) ? a.se1[i] : 0.0f;
        se2[i] = (rr > 0 && i < rr * T) ? a.se2[i] : 0.0f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tU = tmem, tY = tmem + KP, tDW1 = tmem + 2 * KP, tDW2 = tmem + 3 * KP;   // U2 / dN2 | Y2 / dG2 | Wt = dU^T xhat | dW2
    const uint32_t xB = smem_u32(bufX), yB = smem_u32(bufY), w1B = smem_u32(w1b), w2B = smem_u32(w2b);

    const int ntiles = (a.B + g.seq_per_tile - 1) / g.seq_per_tile;
    const bool lane_ok = lane < g.rpw;
    const int t = lane_ok ? lane % T : 0;
    const int seq_base = lane_ok ? (lane / T) * T : 0;
    const int drow = warp * g.rpw + (lane_ok ? lane : 0);
    uint32_t ph_x = 0, ph_dy = 0, ph_mma = 0;
    bool store_pending = false;
    bool first = true;
    float gS1[kMaxRR], gS2[kMaxRR];
#pragma unroll
    for (int k = 0; k < kMaxRR; ++k) gS1[k] = gS2[k] = 0.0f;

    auto tile_nrows = [&](int tile) {
        const int nseq = min(g.seq_per_tile, a.B - tile * g.seq_per_tile);
        return nseq * T;
    };
    if (warp == 0 && (int)blockIdx.x < ntiles) {
        stage_in<VEC>(S1, a.x1, (size_t)blockIdx.x * g.tile_rows, tile_nrows(blockIdx.x), H, g.pitch, &bars[0], lane);
        stage_in<VEC>(S2, a.dy, (size_t)blockIdx.x * g.tile_rows, tile_nrows(blockIdx.x), H, g.pitch, &bars[2], lane);
    }

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nrows = tile_nrows(tile);
        const bool valid = lane_ok && drow < nrows;
        const uint32_t grow = (uint32_t)((size_t)tile * g.tile_rows + drow);
        const int next = tile + gridDim.x;

        // ---------------- P0: xhat = LN2(x1) without affine -> operand X
        mbar_wait(&bars[0], ph_x, abortf);
        ph_x ^= 1;
        float xh[KP];     // xhat row, kept for the whole tile (ones column at H)
        float rstd;
        {
            load_row<KP, VEC>(S1 + (size_t)drow * g.pitch, valid ? H : 0, xh);
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < KP; ++k) s += xh[k];
            const float mean = s / (float)H;
            float ss = 0.0f;
#pragma unroll
            for (int k = 0; k < KP; ++k) { const float dv = k < H ? xh[k] - mean : 0.0f; ss = fmaf(dv, dv, ss); }
            rstd = 1.0f / sqrtf(ss / (float)H + 1e-5f);
#pragma unroll
            for (int k = 0; k < KP; ++k) xh[k] = !valid ? 0.0f : (k < H ? (xh[k] - mean) * rstd : (k == H ? 1.0f : 0.0f));
        }
        if (store_pending) {
            if (warp == 0) bulk_wait_read0();
            store_pending = false;
        }
        __syncthreads();   // S1 consumed, X region free
        if (warp == 0 && next < ntiles) stage_in<VEC>(S1, a.x1, (size_t)next * g.tile_rows, tile_nrows(next), H, g.pitch, &bars[0], lane);
        put_row<KP>(bufX, tid, xh);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            gemm3<0, 0>(tU, xB, xB + P::PLANE, P::PS, w1B, w1B + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        // ---------------- E1: G2 = reg1(act(U2 + b1')) -> operand Y (ones column at ch); U2 stays in TMEM
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c8 = 0; c8 < KP / 8; ++c8) {
            float u[8], ks[8];
            tmem_ld8(tmem_addr(tU, warp, 8 * c8), u);
            tmem_wait_ld();
            if (dr.thresh) drop8(dr, a.site_base + 2, grow, (uint32_t)(ch + 7) >> 3, c8, ks);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * c8 + j;
                float gv = act_fwd<ACT>(u[j] + c1f[c]);
                if (dr.thresh) gv *= ks[j];
                u[j] = !valid ? 0.0f : (c < ch ? gv : (c == ch ? 1.0f : 0.0f));
            }
            put_chunk(bufY, P::PLANE, P::PS, tid, c8, u);
        }
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
            float v[KP];      // y2, then dY2
            float dyr[KP];
            load_row<KP, VEC>(S2 + (size_t)drow * g.pitch, valid ? H : 0, dyr);
            float ssum = 0.0f, dgate = 0.0f;
#pragma unroll
            for (int c8 = 0; c8 < KP / 8; ++c8) {
                float u[8], ks[8];
                tmem_ld8(tmem_addr(tY, warp, 8 * c8), u);
                tmem_wait_ld();
                if (dr.thresh) drop8(dr, a.site_base + 3, grow, (uint32_t)(H + 7) >> 3, c8, ks);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int h = 8 * c8 + j;
                    float yv = u[j] + c2[h];
                    if (dr.thresh) yv *= ks[j];
                    yv = (valid && h < H) ? yv : 0.0f;
                    v[h] = yv;
                    ssum += yv;
                    dgate = fmaf(dyr[h], yv, dgate);
                }
            }
            float gate = 1.0f, dsq = 0.0f;
            if (rr > 0) {
                const float sq = ssum / (float)H;
                const SeOut se = se_excite(sq, t, seq_base, T, rr, se1, se2);
                gate = se.gate;
                const float dq = valid ? dgate * gate * (1.0f - gate) : 0.0f;
                float da[kMaxRR];
#pragma unroll
                for (int k = 0; k < kMaxRR; ++k) da[k] = 0.0f;
                for (int tt = 0; tt < T; ++tt) {
                    const float dqt = __shfl_sync(0xffffffffu, dq, seq_base + tt);
#pragma unroll
                    for (int k = 0; k < kMaxRR; ++k)
                        if (k < rr) da[k] = fmaf(dqt, se2[tt * rr + k], da[k]);
                }
#pragma unroll
                for (int k = 0; k < kMaxRR; ++k)
                    if (k < rr) {
                        const float dz = se.z[k] > 0.0f ? da[k] : 0.0f;
                        dsq = fmaf(dz, se1[k * T + t], dsq);
                        if (valid) {
                            gS2[k] = fmaf(dq, fmaxf(se.z[k], 0.0f), gS2[k]);
                            gS1[k] = fmaf(dz, sq, gS1[k]);
                        }
                    }
                dsq /= (float)H;
            }
            // dY2 = reg2'( dy * gate + dsq )
#pragma unroll
            for (int c8 = 0; c8 < KP / 8; ++c8) {
                float ks[8], o[8];
                if (dr.thresh) drop8(dr, a.site_base + 3, grow, (uint32_t)(H + 7) >> 3, c8, ks);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int h = 8 * c8 + j;
                    float d = fmaf(dyr[h], gate, dsq);
                    if (dr.thresh) d *= ks[j];
                    o[j] = (valid && h < H) ? d : 0.0f;
                }
                put_chunk(bufX, P::PLANE, P::PS, tid, c8, o);
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
        // ---------------- E3: dU2 = reg1'(dG2) * act'(U2 + b1') -> operand Y; xhat -> operand X again
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c8 = 0; c8 < KP / 8; ++c8) {
            float u[8], dg[8], ks[8];
            tmem_ld8(tmem_addr(tU, warp, 8 * c8), u);
            tmem_ld8(tmem_addr(tY, warp, 8 * c8), dg);
            tmem_wait_ld();
            if (dr.thresh) drop8(dr, a.site_base + 2, grow, (uint32_t)(ch + 7) >> 3, c8, ks);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * c8 + j;
                float av;
                const float dact = act_fwd_grad<ACT>(u[j] + c1f[c], &av);
                float d = dg[j] * dact;
                if (dr.thresh) d *= ks[j];
                u[j] = (valid && c < ch) ? d : 0.0f;
            }
            put_chunk(bufY, P::PLANE, P::PS, tid, c8, u);
        }
        put_row<KP>(bufX, tid, xh);
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
            float d[KP];
            float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
            for (int c8 = 0; c8 < KP / 8; ++c8) {
                float u[8];
                tmem_ld8(tmem_addr(tU, warp, 8 * c8), u);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int h = 8 * c8 + j;
                    const float dv = h < H ? u[j] : 0.0f;
                    d[h] = dv;
                    m1 += dv;
                    m2 = fmaf(dv, xh[h], m2);
                }
            }
            m1 /= (float)H;
            m2 /= (float)H;
            float dyr[KP];
            load_row<KP, VEC>(S2 + (size_t)drow * g.pitch, valid ? H : 0, dyr);
#pragma unroll
            for (int k = 0; k < KP; ++k) d[k] = fmaf(rstd, d[k] - m1 - xh[k] * m2, dyr[k]);
            tc_fence_before();
            if (valid) store_row<KP, VEC>(reinterpret_cast<float*>(bufX) + (size_t)drow * g.pitch, H, d);
        }
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            stage_out<VEC>(a.out, reinterpret_cast<const float*>(bufX), (size_t)tile * g.tile_rows, nrows, H, g.pitch, lane);
            store_pending = true;
            if (next < ntiles) stage_in<VEC>(S2, a.dy, (size_t)next * g.tile_rows, tile_nrows(next), H, g.pitch, &bars[2], lane);
        }
    }
    if (warp == 0) bulk_wait_all0();
    __syncthreads();

    // ---------------- flush: Wt, dW2 (TMEM, lane = output row) -> global gradients
    // thread c (< ch) owns row c of Wt [ch][H | ones]; thread h (< H) owns row h of dW2 [H][ch | ones]
    if (!first) {
        float* stg = reinterpret_cast<float*>(bufX);          // [KP][KP+1] staging for lane-contiguous REDs
        constexpr int SP = KP + 1;
        tc_fence_after();
        // Wt
#pragma unroll
        for (int c8 = 0; c8 < KP / 8; ++c8) {
            float u[8];
            tmem_ld8(tmem_addr(tDW1, warp, 8 * c8), u);
            tmem_wait_ld();
            if (tid < KP)
#pragma unroll
                for (int j = 0; j < 8; ++j) stg[tid * SP + 8 * c8 + j] = u[j];
        }
        __syncthreads();
        for (int i = tid; i < ch * H; i += kThreadsChan) {
            const int c = i / H, h = i - c * H;
            red_add(a.g_w1 + i, stg[c * SP + h] * gam[h]);
        }
        for (int h = tid; h < H; h += kThreadsChan) {
            float sg = 0.0f, sb = 0.0f;
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
        // dW2
#pragma unroll
        for (int c8 = 0; c8 < KP / 8; ++c8) {
            float u[8];
            tmem_ld8(tmem_addr(tDW2, warp, 8 * c8), u);
            tmem_wait_ld();
            if (tid < KP)
#pragma unroll
                for (int j = 0; j < 8; ++j) stg[tid * SP + 8 * c8 + j] = u[j];
        }
        __syncthreads();
        for (int i = tid; i < H * ch; i += kThreadsChan) {
            const int h = i / ch, c = i - h * ch;
            red_add(a.g_w2 + i, stg[h * SP + c]);
        }
        for (int h = tid; h < H; h += kThreadsChan) red_add(a.g_b2 + h, stg[h * SP + ch]);
        // SE gradients: every valid thread holds partial sums for its frame t
        if (rr > 0) {
            __syncthreads();
            float* acc = stg;                                  // [2][rr*T]
            for (int i = tid; i < 2 * rr * T; i += kThreadsChan) acc[i] = 0.0f;
            __syncthreads();
            if (lane_ok)
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
