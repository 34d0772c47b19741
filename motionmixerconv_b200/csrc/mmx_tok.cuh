// Token half of a MixerBlock (fp32 CUDA cores, sm_100a), the companion of the tensor-core channel half (mmx_chan_tc5.cuh):
//
//     x1 = x + SE(reg2(fc2(reg1(act(fc1(LN1(x)^T)))))^T)       reference: h36m/mlp_mixer.py:146-155 (MixerBlock.forward, first
//     half), MlpBlock :87-96, SELayer :30-34; restated in oracle/mixer_np.py.
//
// The token MLP contracts over the T frames (K = T and K = tokens_mlp_dim, 10 and 20 in the reference's configurations): far
// too shallow for a tensor-core tile, so it runs as register FMAs, ONE THREAD PER (sequence, hidden column): the thread holds
// its column's T frames in registers, both contractions are thread-local, and the weights are warp-uniform broadcast reads
// from shared memory.  What crosses threads -- LayerNorm statistics, the SE squeeze, the LayerNorm backward sums (all sums over
// the hidden dim of a (sequence, frame) row) -- goes through the shared tile with one thread per ROW doing the sum.
// Tiles of S whole sequences enter and leave through the bulk-copy engine (cp.async.bulk + mbarrier).
// Backward: forward recomputed; the token weight gradients (a [tok x T] product whose contraction runs over ALL columns of
// ALL sequences) are computed from operands staged in shared memory by "owner" threads with register-resident 4x4 accumulator
// tiles that persist across the CTA's loop over tiles (one flush per CTA).
#pragma once
#include "mmx_common.cuh"
#include "mmx_tc5.cuh"
#include "mmx_chan_tc5.cuh"   // drop8

namespace mmx {
namespace tok {

using namespace tc5;

constexpr int kMaxT = 16;      // frames per sequence served (register arrays)
constexpr int kMaxTok = 32;    // tokens_mlp_dim served
constexpr int kMaxRRt = 4;

struct TokArgs {
    const float* x;              // [B,T,H] block input
    const float* dx1;            // backward: gradient wrt x1 (output of the token half) [B,T,H]
    float* out;                  // forward: x1; backward: dx
    const float *ln_g, *ln_b, *w1, *b1, *w2, *b2, *se1, *se2;
    float *g_ln_g, *g_ln_b, *g_w1, *g_b1, *g_w2, *g_b2, *g_se1, *g_se2;
    int B, T, H, tok, rr;
    int S;                       // sequences per tile; S*T*H % 4 == 0
    int site_base;               // token MLP dropout sites: site_base + 0 (after act), + 1 (after fc2)
    Dropout dr;
    int* abort_count;
};

struct TokSmem {                 // offsets in floats
    int x, d, y, stat, sq, gate, w1, b1, w2, b2, lg, lb, se1, se2, stg, misc, total;
    int cols_pad, rows;
};
MMX_HD TokSmem tok_smem(int T, int H, int tok, int S, bool bwd) {
    TokSmem m;
    const int tile = (S * T * H + 3) / 4 * 4;
    m.rows = S * T;
    m.cols_pad = (S * H + 3) / 4 * 4 + 4;
    int o = 0;
    m.x = o; o += tile;
    m.d = o; o += bwd ? tile : 0;
    m.y = o; o += tile;
    m.stat = o; o += 2 * m.rows;            // mean, rstd
    m.sq = o; o += 2 * m.rows;              // squeeze / dgate   (bwd: also m1, m2 of the LayerNorm backward)
    m.gate = o; o += 2 * m.rows;            // gate / ds
    m.w1 = o; o += (tok * T + 3) / 4 * 4;
    m.b1 = o; o += (tok + 3) / 4 * 4;
    m.w2 = o; o += (T * tok + 3) / 4 * 4;
    m.b2 = o; o += (T + 3) / 4 * 4;
    m.lg = o; o += (H + 3) / 4 * 4;
    m.lb = o; o += (H + 3) / 4 * 4;
    m.se1 = o; o += 32 * kMaxRRt;
    m.se2 = o; o += 32 * kMaxRRt;
    m.stg = o; o += bwd ? (2 * T + 1 + 2 * tok + 1) * m.cols_pad : 0;   // N (T+1 rows: ones last), DYT (T), DU (tok), G (tok+1: ones last)
    m.misc = o; o += 32;                     // barriers, abort flag
    m.total = o;
    return m;
}

// sum over the H columns of row `row` of a dense [rows][H] tile (rotated start: conflict-free for any H)
MMX_D float row_sum(const float* tile, int row, int H) {
    const float* p = tile + (size_t)row * H;
    float s = 0.0f;
    int k = row % H;
    for (int i = 0; i < H; ++i) { s += p[k]; k = k + 1 == H ? 0 : k + 1; }
    return s;
}

// excitation of one sequence (T squeeze values at sq[0..T)) for frame t; returns the gate and the pre-activations z
MMX_D float excite(const float* sq, int t, int T, int rr, const float* se1, const float* se2, float (&z)[kMaxRRt]) {
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxRRt; ++k) {
        z[k] = 0.0f;
        if (k < rr) {
            for (int tt = 0; tt < T; ++tt) z[k] = fmaf(se1[k * T + tt], sq[tt], z[k]);
            q = fmaf(se2[t * rr + k], fmaxf(z[k], 0.0f), q);
        }
    }
    return sigmoidf_(q);
}

template <int NW>
MMX_D void load_params(float* sm, const TokSmem& m, const TokArgs& a, int tid) {
    const int nt = NW * 32;
    for (int i = tid; i < a.tok * a.T; i += nt) { sm[m.w1 + i] = a.w1[i]; sm[m.w2 + i] = a.w2[i]; }
    for (int i = tid; i < a.tok; i += nt) sm[m.b1 + i] = a.b1[i];
    for (int i = tid; i < a.T; i += nt) sm[m.b2 + i] = a.b2[i];
    for (int i = tid; i < a.H; i += nt) { sm[m.lg + i] = a.ln_g[i]; sm[m.lb + i] = a.ln_b[i]; }
    for (int i = tid; i < 32 * kMaxRRt; i += nt) {
        sm[m.se1 + i] = (a.rr > 0 && i < a.rr * a.T) ? a.se1[i] : 0.0f;
        sm[m.se2 + i] = (a.rr > 0 && i < a.rr * a.T) ? a.se2[i] : 0.0f;
    }
}

// forward of one column: n[t] (normalised, affine) -> y[t] (token MLP output after reg2).  TT, TOK compile-time bounds.
template <int ACT, int TT>
MMX_D void column_fwd(const float* sm, const TokSmem& m, int T, int tok, const float (&n)[TT], float (&y)[TT], const Dropout& dr,
                      uint32_t site_base, uint32_t colrow) {
#pragma unroll
    for (int t = 0; t < TT; ++t) y[t] = t < T ? sm[m.b2 + t] : 0.0f;
    const uint32_t tok8 = (uint32_t)(tok + 7) >> 3;
    for (int j8 = 0; j8 < (int)tok8; ++j8) {
        float ks[8];
        if (dr.thresh) chan::drop8(dr, site_base + 0, colrow, tok8, j8, ks);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = 8 * j8 + jj;
            if (j < tok) {
                float u = sm[m.b1 + j];
                const float* w = sm + m.w1 + j * T;
#pragma unroll
                for (int t = 0; t < TT; ++t)
                    if (t < T) u = fmaf(w[t], n[t], u);
                float gv = act_fwd<ACT>(u);
                if (dr.thresh) gv *= ks[jj];
#pragma unroll
                for (int t = 0; t < TT; ++t)
                    if (t < T) y[t] = fmaf(sm[m.w2 + t * tok + j], gv, y[t]);
            }
        }
    }
    if (dr.thresh) {
        const uint32_t t8 = (uint32_t)(T + 7) >> 3;
        for (int c8 = 0; c8 < (int)t8; ++c8) {
            float ks[8];
            chan::drop8(dr, site_base + 1, colrow, t8, c8, ks);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
                if (8 * c8 + jj < TT) y[8 * c8 + jj] *= ks[jj];
        }
    }
}

// ==========================================================================================
// forward:  x -> x1
// ==========================================================================================
template <int ACT, int TT, int NW>
__global__ void __launch_bounds__(NW * 32) tok_fwd_kernel(const TokArgs a) {
    extern __shared__ float4 tok_smem_raw[];
    float* sm = reinterpret_cast<float*>(tok_smem_raw);
    const TokSmem m = tok_smem(a.T, a.H, a.tok, a.S, false);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + m.misc);
    volatile int* abortf = reinterpret_cast<volatile int*>(sm + m.misc + 8);
    const int tid = threadIdx.x, T = a.T, H = a.H, S = a.S, rr = a.rr;
    const Dropout dr = resolve_dropout(a.dr);
    if (tid == 0) { mbar_init(&bars[0], 1); *abortf = 0; fence_mbar_init(); }
    load_params<NW>(sm, m, a, tid);
    __syncthreads();
    const int ntiles = (a.B + S - 1) / S;
    const int s_l = tid / H, h = tid - s_l * H;
    const bool col_ok = tid < S * H;
    uint32_t ph = 0;
    auto tile_bytes = [&](int tile) { return (uint32_t)(min(S, a.B - tile * S) * T * H) * 4u; };
    if (tid == 0 && (int)blockIdx.x < ntiles) {
        mbar_expect_tx(&bars[0], tile_bytes(blockIdx.x));
        bulk_g2s(sm + m.x, a.x + (size_t)blockIdx.x * S * T * H, tile_bytes(blockIdx.x), &bars[0]);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nseq = min(S, a.B - tile * S);
        const int nrows = nseq * T;
        mbar_wait(&bars[0], ph, abortf);
        ph ^= 1;
        // ---- A: LayerNorm statistics, one thread per row
        if (tid < nrows) {
            const float mean = row_sum(sm + m.x, tid, H) / (float)H;
            const float* p = sm + m.x + (size_t)tid * H;
            float ss = 0.0f;
            int k = tid % H;
            for (int i = 0; i < H; ++i) { const float dv = p[k] - mean; ss = fmaf(dv, dv, ss); k = k + 1 == H ? 0 : k + 1; }
            sm[m.stat + 2 * tid] = mean;
            sm[m.stat + 2 * tid + 1] = 1.0f / sqrtf(ss / (float)H + 1e-5f);
        }
        __syncthreads();
        // ---- B: token MLP of the thread's column
        const bool act_col = col_ok && s_l < nseq;
        float x[TT], y[TT];
        if (act_col) {
            float n[TT];
            const float gmm = sm[m.lg + h], bta = sm[m.lb + h];
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t < T) {
                    const int r = s_l * T + t;
                    x[t] = sm[m.x + (size_t)r * H + h];
                    n[t] = fmaf((x[t] - sm[m.stat + 2 * r]) * sm[m.stat + 2 * r + 1], gmm, bta);
                } else { x[t] = 0.0f; n[t] = 0.0f; }
            }
            column_fwd<ACT, TT>(sm, m, T, a.tok, n, y, dr, a.site_base, (uint32_t)((size_t)(tile * S + s_l) * H + h));
            if (rr > 0) {
#pragma unroll
                for (int t = 0; t < TT; ++t)
                    if (t < T) sm[m.y + (size_t)(s_l * T + t) * H + h] = y[t];
            }
        }
        if (rr > 0) {
            __syncthreads();
            // ---- C: squeeze (mean over the hidden dim of every row)
            if (tid < nrows) sm[m.sq + tid] = row_sum(sm + m.y, tid, H) / (float)H;
            __syncthreads();
            if (tid < nrows) {
                float z[kMaxRRt];
                const int sq0 = (tid / T) * T;
                sm[m.gate + tid] = excite(sm + m.sq + sq0, tid - sq0, T, rr, sm + m.se1, sm + m.se2, z);
            }
            __syncthreads();
        }
        // ---- D: gate + residual, in place in the x tile
        if (act_col) {
#pragma unroll
            for (int t = 0; t < TT; ++t)
                if (t < T) {
                    const int r = s_l * T + t;
                    const float gte = rr > 0 ? sm[m.gate + r] : 1.0f;
                    sm[m.x + (size_t)r * H + h] = fmaf(y[t], gte, x[t]);
                }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(a.out + (size_t)tile * S * T * H, sm + m.x, (uint32_t)nrows * H * 4u);
            bulk_commit();
            const int next = tile + gridDim.x;
            bulk_wait_read0();                 // the x tile is both the store source and the next load's destination
            if (next < ntiles) {
                mbar_expect_tx(&bars[0], tile_bytes(next));
                bulk_g2s(sm + m.x, a.x + (size_t)next * S * T * H, tile_bytes(next), &bars[0]);
            }
        }
    }
    if (tid == 0) {
        bulk_wait_all0();
        if (*abortf) atomicAdd(a.abort_count, 1);
    }
}

// ==========================================================================================
// backward:  (x, dx1) -> dx, parameter gradients of the token half
// ==========================================================================================
template <int ACT, int TT, int NW>
__global__ void __launch_bounds__(NW * 32) tok_bwd_kernel(const TokArgs a) {
    extern __shared__ float4 tok_smem_raw[];
    float* sm = reinterpret_cast<float*>(tok_smem_raw);
    const TokSmem m = tok_smem(a.T, a.H, a.tok, a.S, true);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + m.misc);
    volatile int* abortf = reinterpret_cast<volatile int*>(sm + m.misc + 8);
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, T = a.T, H = a.H, S = a.S, rr = a.rr, tok = a.tok;
    const Dropout dr = resolve_dropout(a.dr);
    if (tid == 0) { mbar_init(&bars[0], 1); *abortf = 0; fence_mbar_init(); }
    load_params<NW>(sm, m, a, tid);
    // staging area: rows [0,T] = N (row T: ones), [T+1, 2T] = DYT, [2T+1, 2T+tok] = DU, [2T+tok+1, 2T+2tok+1] = G (last: ones)
    const int CP = m.cols_pad;
    float* sN = sm + m.stg;
    float* sDYT = sN + (T + 1) * CP;
    float* sDU = sDYT + T * CP;
    float* sG = sDU + tok * CP;
    for (int i = tid; i < (2 * T + 2 * tok + 2) * CP; i += NT) sm[m.stg + i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < S * H; i += NT) { sN[T * CP + i] = 1.0f; sG[tok * CP + i] = 1.0f; }
    __syncthreads();

    const int ntiles = (a.B + S - 1) / S;
    const int s_l = tid / H, h = tid - s_l * H;
    const bool col_ok = tid < S * H;
    // ---- owner threads of the weight-gradient tiles
    //   product 0: dW1ext[j][t'] = sum_col DU[j][col] * N[t'][col]       (j < tok, t' <= T; column T = db1)
    //   product 1: dW2ext[t][j'] = sum_col DYT[t][col] * G[j'][col]      (t < T, j' <= tok; column tok = db2)
    const int nb0r = (tok + 3) / 4, nb0c = (T + 1 + 3) / 4, nb1r = (T + 3) / 4, nb1c = (tok + 1 + 3) / 4;
    const int nblocks = nb0r * nb0c + nb1r * nb1c;
    const int KS = max(1, NT / nblocks);
    const int ob = tid / KS, oks = tid - ob * KS;
    const bool owner = ob < nblocks;
    const int oprod = ob < nb0r * nb0c ? 0 : 1;
    const int obb = oprod ? ob - nb0r * nb0c : ob;
    const int obr = oprod ? obb / nb1c : obb / nb0c, obc = oprod ? obb - obr * nb1c : obb - obr * nb0c;
    const float* oA = oprod ? sDYT : sDU;
    const float* oB = oprod ? sG : sN;
    const int oAr = oprod ? T : tok, oBr = oprod ? tok + 1 : T + 1;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    float g_lg = 0.0f, g_lb = 0.0f;                  // dLN1.weight[h], dLN1.bias[h] of the thread's column
    float gS1[kMaxRRt], gS2[kMaxRRt];
#pragma unroll
    for (int k = 0; k < kMaxRRt; ++k) gS1[k] = gS2[k] = 0.0f;

    uint32_t ph = 0;
    auto tile_bytes = [&](int tile) { return (uint32_t)(min(S, a.B - tile * S) * T * H) * 4u; };
    if (tid == 0 && (int)blockIdx.x < ntiles) {
        mbar_expect_tx(&bars[0], 2 * tile_bytes(blockIdx.x));
        bulk_g2s(sm + m.x, a.x + (size_t)blockIdx.x * S * T * H, tile_bytes(blockIdx.x), &bars[0]);
        bulk_g2s(sm + m.d, a.dx1 + (size_t)blockIdx.x * S * T * H, tile_bytes(blockIdx.x), &bars[0]);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nseq = min(S, a.B - tile * S);
        const int nrows = nseq * T;
        mbar_wait(&bars[0], ph, abortf);
        ph ^= 1;
        // ---- A: LayerNorm statistics
        if (tid < nrows) {
            const float mean = row_sum(sm + m.x, tid, H) / (float)H;
            const float* p = sm + m.x + (size_t)tid * H;
            float ss = 0.0f;
            int k = tid % H;
            for (int i = 0; i < H; ++i) { const float dv = p[k] - mean; ss = fmaf(dv, dv, ss); k = k + 1 == H ? 0 : k + 1; }
            sm[m.stat + 2 * tid] = mean;
            sm[m.stat + 2 * tid + 1] = 1.0f / sqrtf(ss / (float)H + 1e-5f);
        }
        __syncthreads();
        // ---- B: forward of the column (n, y)
        const bool act_col = col_ok && s_l < nseq;
        const uint32_t colrow = (uint32_t)((size_t)(tile * S + s_l) * H + h);
        float n[TT], xh[TT], dyt[TT];
        const float gmm = col_ok ? sm[m.lg + h] : 0.0f, bta = col_ok ? sm[m.lb + h] : 0.0f;
        if (act_col) {
            float y[TT];
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t < T) {
                    const int r = s_l * T + t;
                    xh[t] = (sm[m.x + (size_t)r * H + h] - sm[m.stat + 2 * r]) * sm[m.stat + 2 * r + 1];
                    n[t] = fmaf(xh[t], gmm, bta);
                } else { xh[t] = 0.0f; n[t] = 0.0f; }
            }
            if (rr > 0) {
                column_fwd<ACT, TT>(sm, m, T, tok, n, y, dr, a.site_base, colrow);
#pragma unroll
                for (int t = 0; t < TT; ++t)
                    if (t < T) sm[m.y + (size_t)(s_l * T + t) * H + h] = y[t];
            }
        }
        if (rr > 0) {
            __syncthreads();
            // ---- C: squeeze and d(gate) per row
            if (tid < nrows) {
                const float* py = sm + m.y + (size_t)tid * H;
                const float* pd = sm + m.d + (size_t)tid * H;
                float s = 0.0f, dg = 0.0f;
                int k = tid % H;
                for (int i = 0; i < H; ++i) { s += py[k]; dg = fmaf(pd[k], py[k], dg); k = k + 1 == H ? 0 : k + 1; }
                sm[m.sq + 2 * tid] = s / (float)H;
                sm[m.sq + 2 * tid + 1] = dg;
            }
            __syncthreads();
            // ---- C2: excitation forward + backward of the row's sequence
            if (tid < nrows) {
                const int sq0 = (tid / T) * T, t = tid - sq0;
                float sqv[kMaxT], z[kMaxRRt];
                for (int tt = 0; tt < T; ++tt) sqv[tt] = sm[m.sq + 2 * (sq0 + tt)];
                const float gate = excite(sqv, t, T, rr, sm + m.se1, sm + m.se2, z);
                float da[kMaxRRt];
#pragma unroll
                for (int k = 0; k < kMaxRRt; ++k) da[k] = 0.0f;
                float dq_own = 0.0f;
                for (int tt = 0; tt < T; ++tt) {
                    float q = 0.0f;
#pragma unroll
                    for (int k = 0; k < kMaxRRt; ++k)
                        if (k < rr) q = fmaf(sm[m.se2 + tt * rr + k], fmaxf(z[k], 0.0f), q);
                    const float gt = sigmoidf_(q);
                    const float dq = sm[m.sq + 2 * (sq0 + tt) + 1] * gt * (1.0f - gt);
                    if (tt == t) dq_own = dq;
#pragma unroll
                    for (int k = 0; k < kMaxRRt; ++k)
                        if (k < rr) da[k] = fmaf(dq, sm[m.se2 + tt * rr + k], da[k]);
                }
                float ds = 0.0f;
#pragma unroll
                for (int k = 0; k < kMaxRRt; ++k)
                    if (k < rr) {
                        const float dz = z[k] > 0.0f ? da[k] : 0.0f;
                        ds = fmaf(dz, sm[m.se1 + k * T + t], ds);
                        gS2[k] = fmaf(dq_own, fmaxf(z[k], 0.0f), gS2[k]);
                        gS1[k] = fmaf(dz, sqv[t], gS1[k]);
                    }
                sm[m.gate + 2 * tid] = gate;
                sm[m.gate + 2 * tid + 1] = ds / (float)H;
            }
            __syncthreads();
        }
        // ---- E: backward of the column
        float dnh[TT];
        if (act_col) {
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t < T) {
                    const int r = s_l * T + t;
                    const float d1 = sm[m.d + (size_t)r * H + h];
                    dyt[t] = rr > 0 ? fmaf(d1, sm[m.gate + 2 * r], sm[m.gate + 2 * r + 1]) : d1;
                } else dyt[t] = 0.0f;
            }
            if (dr.thresh) {
                const uint32_t t8 = (uint32_t)(T + 7) >> 3;
                for (int c8 = 0; c8 < (int)t8; ++c8) {
                    float ks[8];
                    chan::drop8(dr, a.site_base + 1, colrow, t8, c8, ks);
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        if (8 * c8 + jj < TT) dyt[8 * c8 + jj] *= ks[jj];
                }
            }
            float dn[TT];
#pragma unroll
            for (int t = 0; t < TT; ++t) dn[t] = 0.0f;
            const uint32_t tok8 = (uint32_t)(tok + 7) >> 3;
            for (int j8 = 0; j8 < (int)tok8; ++j8) {
                float ks[8];
                if (dr.thresh) chan::drop8(dr, a.site_base + 0, colrow, tok8, j8, ks);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = 8 * j8 + jj;
                    if (j < tok) {
                        float u = sm[m.b1 + j];
                        const float* w = sm + m.w1 + j * T;
                        float dgj = 0.0f;
#pragma unroll
                        for (int t = 0; t < TT; ++t)
                            if (t < T) { u = fmaf(w[t], n[t], u); dgj = fmaf(sm[m.w2 + t * tok + j], dyt[t], dgj); }
                        float av;
                        const float dact = act_fwd_grad<ACT>(u, &av);
                        float du = dgj * dact;
                        if (dr.thresh) { av *= ks[jj]; du *= ks[jj]; }
#pragma unroll
                        for (int t = 0; t < TT; ++t)
                            if (t < T) dn[t] = fmaf(w[t], du, dn[t]);
                        sDU[j * CP + tid] = du;
                        sG[j * CP + tid] = av;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < TT; ++t)
                if (t < T) {
                    sN[t * CP + tid] = n[t];
                    sDYT[t * CP + tid] = dyt[t];
                    g_lg = fmaf(dn[t], xh[t], g_lg);
                    g_lb += dn[t];
                    dnh[t] = dn[t] * gmm;
                    sm[m.y + (size_t)(s_l * T + t) * H + h] = dnh[t];
                }
        } else if (col_ok) {
            for (int j = 0; j < tok; ++j) { sDU[j * CP + tid] = 0.0f; sG[j * CP + tid] = 0.0f; }
            for (int t = 0; t < T; ++t) { sN[t * CP + tid] = 0.0f; sDYT[t * CP + tid] = 0.0f; }
        }
        __syncthreads();
        // ---- F: LayerNorm backward row sums; weight-gradient tiles
        if (tid < nrows) {
            const float* pd = sm + m.y + (size_t)tid * H;
            const float* px = sm + m.x + (size_t)tid * H;
            const float mean = sm[m.stat + 2 * tid], rstd = sm[m.stat + 2 * tid + 1];
            float s1 = 0.0f, s2 = 0.0f;
            int k = tid % H;
            for (int i = 0; i < H; ++i) { s1 += pd[k]; s2 = fmaf(pd[k], (px[k] - mean) * rstd, s2); k = k + 1 == H ? 0 : k + 1; }
            sm[m.sq + 2 * tid] = s1 / (float)H;
            sm[m.sq + 2 * tid + 1] = s2 / (float)H;
        }
        if (owner) {
            const float* ar[4];
            const float* br[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ar[i] = oA + (size_t)min(4 * obr + i, oAr - 1) * CP;
                br[i] = oB + (size_t)min(4 * obc + i, oBr - 1) * CP;
            }
            for (int q = oks; q < CP / 4; q += KS) {
                float4 av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { av[i] = *reinterpret_cast<const float4*>(ar[i] + 4 * q); bv[i] = *reinterpret_cast<const float4*>(br[i] + 4 * q); }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                        acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                        acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                        acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                    }
            }
        }
        __syncthreads();
        // ---- G: dx = dx1 + LN1 backward, in place in the dx1 tile
        if (act_col) {
#pragma unroll
            for (int t = 0; t < TT; ++t)
                if (t < T) {
                    const int r = s_l * T + t;
                    const float rstd = sm[m.stat + 2 * r + 1];
                    const float v = rstd * (dnh[t] - sm[m.sq + 2 * r] - xh[t] * sm[m.sq + 2 * r + 1]);
                    sm[m.d + (size_t)r * H + h] += v;
                }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(a.out + (size_t)tile * S * T * H, sm + m.d, (uint32_t)nrows * H * 4u);
            bulk_commit();
            const int next = tile + gridDim.x;
            bulk_wait_read0();
            if (next < ntiles) {
                mbar_expect_tx(&bars[0], 2 * tile_bytes(next));
                bulk_g2s(sm + m.x, a.x + (size_t)next * S * T * H, tile_bytes(next), &bars[0]);
                bulk_g2s(sm + m.d, a.dx1 + (size_t)next * S * T * H, tile_bytes(next), &bars[0]);
            }
        }
    }
    if (tid == 0) bulk_wait_all0();
    __syncthreads();
    // ---------------- flush
    float* red = sm + m.stg;           // staging area is free now
    const int n0 = tok * (T + 1), n1 = T * (tok + 1);
    for (int i = tid; i < n0 + n1 + 2 * H + 2 * 32 * kMaxRRt; i += NT) red[i] = 0.0f;
    __syncthreads();
    if (owner) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = 4 * obr + i, c = 4 * obc + j;
                if (oprod == 0) { if (r < tok && c <= T) atomicAdd(red + r * (T + 1) + c, acc[i][j]); }
                else { if (r < T && c <= tok) atomicAdd(red + n0 + r * (tok + 1) + c, acc[i][j]); }
            }
    }
    if (col_ok) {
        atomicAdd(red + n0 + n1 + h, g_lg);
        atomicAdd(red + n0 + n1 + H + h, g_lb);
    }
    if (rr > 0 && tid < S * T) {
        const int t = tid % T;
        float* rs = red + n0 + n1 + 2 * H;
        for (int k = 0; k < rr; ++k) {
            atomicAdd(rs + k * T + t, gS1[k]);
            atomicAdd(rs + 32 * kMaxRRt + t * rr + k, gS2[k]);
        }
    }
    __syncthreads();
    for (int i = tid; i < tok * T; i += NT) {
        const int j = i / T, t = i - j * T;
        red_add(a.g_w1 + i, red[j * (T + 1) + t]);
        const int t2 = i / tok, j2 = i - t2 * tok;
        red_add(a.g_w2 + i, red[n0 + t2 * (tok + 1) + j2]);
    }
    for (int j = tid; j < tok; j += NT) red_add(a.g_b1 + j, red[j * (T + 1) + T]);
    for (int t = tid; t < T; t += NT) red_add(a.g_b2 + t, red[n0 + t * (tok + 1) + tok]);
    for (int i = tid; i < H; i += NT) { red_add(a.g_ln_g + i, red[n0 + n1 + i]); red_add(a.g_ln_b + i, red[n0 + n1 + H + i]); }
    if (rr > 0)
        for (int i = tid; i < rr * T; i += NT) {
            red_add(a.g_se1 + i, red[n0 + n1 + 2 * H + i]);
            red_add(a.g_se2 + i, red[n0 + n1 + 2 * H + 32 * kMaxRRt + i]);
        }
    if (tid == 0 && *abortf) atomicAdd(a.abort_count, 1);
}

}  // namespace tok
}  // namespace mmx
