// Fused ConvMixer kernels: ConvMixerBlock halves (forward / backward) and the "once"-mode SE tail.
//
// Reference arithmetic: h36m/conv_mixer_model.py  MultiChanSELayer :47-70, ConvBlock :129-142,
// ConvMixerBlock.forward :268-292 (restated in SURVEY.md Appendix A.2 and oracle/mixer_np.py).
//
// A ConvMixerBlock is two structurally identical halves
//     y = x + SE(reg(act(conv2d(LN(x)))))                 (conv_mixer_model.py:279-284 and :287-292)
// with their own LN / conv weights and a shared SE.  One kernel runs one half: a CTA owns S whole
// sequences ([C,T,E] each), normalises them into a zero-padded shared-memory tile, runs the
// convolution as a register-tiled stencil over that tile (every thread: all output channels x EW
// consecutive embedding positions), and applies activation, dropout, squeeze-excitation and the
// residual before the single store.  The backward recomputes the forward from the half's input.
#pragma once
#include "mmx_common.cuh"
#include "mmx_mlp.cuh"

namespace mmx {

struct ConvDims {
    int B, C, T, E;
    int kT, kP, pT, pP;   // kernel (time, embedding) and top / left zero padding ('same': (k-1)/2)
    int rr, S;            // SE hidden width (0: no SE), sequences per CTA tile
    int use_se, use_max, training;
    int site;             // dropout site of this half (2*block_index + half)
    int x_in_smem;        // backward: X and dY tiles live in shared memory (else re-read from global)
    int bn_mode;          // 0: fused half.  forward 1: BatchNorm statistics pass (writes Z, accumulates batch sums).
                          // backward 2: BatchNorm second pass (Z, gates and BN coefficients come from the first pass)
};

// BatchNorm2d (regularization == -1, conv_mixer_model.py:115-116,141) sits between the activation and the SE layer and
// needs batch-global statistics, so a training half runs as
//   forward : stats pass (LN -> conv -> Z to HBM, sum(A), sum(A^2) per channel)  ->  apply pass (mmx_conv_io.cuh)
//   backward: pass 1 (SE backward, sum(dR), sum(dR*xhat) per channel; mmx_conv_io.cuh)  ->  pass 2 (this file: BN backward
//             folded into the dZ phase, then the usual conv / LN backward; the conv is NOT recomputed, Z was saved)
// The per-channel vectors are [scale | shift | xs | xo] (R = A*scale + shift, xhat = A*xs + xo) and
// [k1 | k2 | k3] (dA = k1*(dR - k2 - xhat*k3)), computed from the sums by the host layer with tiny tensor ops.
// In eval mode BN is a fixed per-channel affine: the fused kernels take it as an optional epilogue (`aff`).

struct ConvHalfW {        // parameter (or gradient) pointers of one half, reference layouts
    float *ln_g, *ln_b;   // LN{1,2}.weight / bias                 [E]
    float *cw, *cb;       // conv{1,2}.conv.weight / bias          [C,C,kT,kP], [C]
    float *se1, *se2;     // se.excitationBlock.{0,2}.weight       [rr,T], [T,rr]
};

MMX_HD int conv_cp(int C) { return C <= 1 ? 1 : C <= 2 ? 2 : C <= 4 ? 4 : 8; }
MMX_HD int conv_ew(int CP) { return CP <= 4 ? 8 : 4; }

struct ConvSmem {
    int CP, EW, PE, TP, EP, R, ST;
    int ln_g, ln_b, wtab, wtab2, cb, se1, se2;
    int part, part2, mean, rstd, pool, gate, amax, z, dq, dz, ds;
    int a_lng, a_lnb, a_cb, a_se1, a_se2, a_cw, a_bn;
    int bX, bD, nPad, zPad, bA, dN;
    int total;
};

MMX_HD ConvSmem conv_smem(const ConvDims& d, bool bwd) {
    ConvSmem L;
    const int C = d.C, T = d.T, E = d.E, S = d.S, rr = imax(d.rr, 1);
    L.CP = conv_cp(C);
    L.EW = conv_ew(L.CP);
    L.PE = pitch_of(E);
    L.TP = T + d.kT - 1;
    // padded row: wide enough for the stencil's aligned segment loads and the weight-gradient's 8-wide windows
    L.EP = round_up(imax(round_up(E, L.EW) + d.kP - 1, round_up(E, 4) + 4 * ((d.kP + 3) / 4) + 4), 4);
    if (((L.EP >> 2) & 1) == 0) L.EP += 4;
    L.R = S * C * T;
    L.ST = S * T;
    int o = 0;
    auto take = [&](int n) { int r = o; o += round_up(n, 4); return r; };
    L.ln_g = take(E); L.ln_b = take(E);
    L.wtab = take(C * d.kT * d.kP * L.CP);
    L.wtab2 = bwd ? take(C * d.kT * d.kP * L.CP) : -1;
    L.cb = take(L.CP);
    L.a_bn = take(16);   // BatchNorm statistics pass: per-CTA sum(A)[8], sum(A^2)[8]
    L.se1 = take(rr * T); L.se2 = take(T * rr);
    const int nred = imax(L.R, L.ST) * kParts;
    L.part = take(nred); L.part2 = take(nred);
    L.mean = take(L.R); L.rstd = take(L.R);
    L.pool = take(L.ST); L.gate = take(L.ST); L.amax = take(L.ST);
    L.z = take(S * rr); L.dq = take(L.ST); L.dz = take(S * rr); L.ds = take(L.ST);
    if (bwd) {
        L.a_lng = take(E); L.a_lnb = take(E); L.a_cb = take(L.CP);
        L.a_se1 = take(rr * T); L.a_se2 = take(T * rr);
        L.a_cw = take(C * C * d.kT * d.kP);
    } else {
        L.a_lng = L.a_lnb = L.a_cb = L.a_se1 = L.a_se2 = L.a_cw = -1;
    }
    const int tile = L.R * L.PE, ptile = S * C * L.TP * L.EP;
    if (!bwd) {
        L.bX = take(tile); L.nPad = take(ptile); L.bA = take(tile);
        L.bD = L.zPad = L.dN = -1;
    } else {
        if (d.x_in_smem) { L.bX = take(tile); L.bD = take(tile); } else { L.bX = L.bD = -1; }
        L.nPad = take(ptile); L.zPad = take(ptile);
        L.dN = L.nPad;   // dN overwrites the (dead) normalised input after the weight-gradient phase
        L.bA = -1;
    }
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------
// the stencil:  out[s][co][t][e] = sum_{ci,i,j} wtab[ci][i][j][co] * in[s][ci][t+i][e+j]
// `in` is the zero-padded tile [ns][Cin][TP][EP]; wtab rows are CP floats (output channels, zero
// padded).  A thread owns all CP output channels of EW consecutive positions; the epilogue is called
// per aligned quad: epi(s, co, t, e, v[4], n) for co < Cout, e % 4 == 0, n = valid values (e + n <= E).
// ------------------------------------------------------------------------------------------
template <int CP>
MMX_D void load_wvec(const float* w, float (&wv)[CP]) {
    if (CP >= 4) {
        MMX_UNROLL
        for (int q = 0; q < CP / 4; ++q) { f4 v = ld4(w + 4 * q); wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w; }
    } else {
        MMX_UNROLL
        for (int c = 0; c < CP; ++c) wv[c] = w[c];
    }
}

template <int CP, int EW, int KP, class Epi>
MMX_D void conv_corr_k(int tid, int nthr, const float* in, const float* wtab, int Cin, int Cout, int kT, int kP,
                       int TP, int EP, int ns, int T, int E, Epi&& epi) {
    const int nEB = (E + EW - 1) / EW;
    const int items = ns * T * nEB;
    constexpr int SEG = KP > 0 ? ((EW + KP - 1 + 3) / 4) * 4 : 4;
    for (int it = tid; it < items; it += nthr) {
        const int eb = it % nEB, st = it / nEB, t = st % T, s = st / T;
        const int e0 = eb * EW;
        float acc[CP][EW];
        MMX_UNROLL
        for (int c = 0; c < CP; ++c)
            MMX_UNROLL
            for (int e = 0; e < EW; ++e) acc[c][e] = 0.0f;
        for (int ci = 0; ci < Cin; ++ci) {
            const float* base = in + ((size_t)(s * Cin + ci) * TP + t) * EP + e0;
            const float* wc = wtab + (size_t)ci * kT * kP * CP;
            for (int i = 0; i < kT; ++i) {
                const float* row = base + (size_t)i * EP;
                const float* w = wc + (size_t)i * kP * CP;
                if (KP > 0) {
                    float seg[SEG];
                    MMX_UNROLL
                    for (int q = 0; q < SEG / 4; ++q) {
                        f4 v = ld4(row + 4 * q);
                        seg[4 * q] = v.x; seg[4 * q + 1] = v.y; seg[4 * q + 2] = v.z; seg[4 * q + 3] = v.w;
                    }
                    MMX_UNROLL
                    for (int j = 0; j < (KP > 0 ? KP : 1); ++j) {
                        float wv[CP];
                        load_wvec<CP>(w + j * CP, wv);
                        MMX_UNROLL
                        for (int c = 0; c < CP; ++c)
                            MMX_UNROLL
                            for (int e = 0; e < EW; ++e) acc[c][e] = fmaf(wv[c], seg[e + j], acc[c][e]);
                    }
                } else {
                    for (int j = 0; j < kP; ++j) {
                        float wv[CP];
                        load_wvec<CP>(w + j * CP, wv);
                        MMX_UNROLL
                        for (int e = 0; e < EW; ++e) {
                            const float v = row[e + j];
                            MMX_UNROLL
                            for (int c = 0; c < CP; ++c) acc[c][e] = fmaf(wv[c], v, acc[c][e]);
                        }
                    }
                }
            }
        }
        MMX_UNROLL
        for (int c = 0; c < CP; ++c) {
            if (c < Cout) {
                MMX_UNROLL
                for (int q = 0; q < EW / 4; ++q)
                    if (e0 + 4 * q < E) epi(s, c, t, e0 + 4 * q, &acc[c][4 * q], imin(4, E - e0 - 4 * q));
            }
        }
    }
}

template <int CP, int EW, class Epi>
MMX_D void conv_corr(int tid, int nthr, const float* in, const float* wtab, int Cin, int Cout, int kT, int kP,
                     int TP, int EP, int ns, int T, int E, Epi&& epi) {
    switch (kP) {
        case 1: conv_corr_k<CP, EW, 1>(tid, nthr, in, wtab, Cin, Cout, kT, kP, TP, EP, ns, T, E, epi); break;
        case 3: conv_corr_k<CP, EW, 3>(tid, nthr, in, wtab, Cin, Cout, kT, kP, TP, EP, ns, T, E, epi); break;
        case 5: conv_corr_k<CP, EW, 5>(tid, nthr, in, wtab, Cin, Cout, kT, kP, TP, EP, ns, T, E, epi); break;
        case 9: conv_corr_k<CP, EW, 9>(tid, nthr, in, wtab, Cin, Cout, kT, kP, TP, EP, ns, T, E, epi); break;
        default: conv_corr_k<CP, EW, 0>(tid, nthr, in, wtab, Cin, Cout, kT, kP, TP, EP, ns, T, E, epi); break;
    }
}

// Weight gradient with a thread-owned accumulator that persists across tiles:
//     acc[co][jj] += sum_{s,t,e} dz[s][co][t][e] * n[s][ci][t+i][e+j0+jj]
// thread -> (slice, ci, i, jb): j0 = 4*jb; the (s,t) pairs are dealt round-robin to the slices.
// zpad holds dz at offset (qT, qP) of its padded tile (zero halo), npad the normalised input.
template <int CP, int DLT>
MMX_D void conv_wgrad_acc_d(float (&acc)[CP][4], int tid, int nthr, const float* zpad, const float* npad, int C,
                            int kT, int kP, int qT, int qP, int TP, int EP, int ns, int T, int E) {
    const int nJB = (kP + 3) >> 2;
    const int n_items = C * kT * nJB;
    const int nsl = imax(1, nthr / n_items);
    const int sl = tid / n_items, item = tid - sl * n_items;
    if (sl >= nsl) return;
    const int jb = item % nJB, ii = (item / nJB) % kT, ci = item / (nJB * kT);
    for (int st = sl; st < ns * T; st += nsl) {
        const int s = st / T, t = st - s * T;
        const float* nrow = npad + ((size_t)(s * C + ci) * TP + t + ii) * EP + 4 * jb;
        const float* zrow = zpad + ((size_t)(s * C) * TP + qT + t) * EP + qP - DLT;    // 16-byte aligned: DLT = qP & 3
        for (int e = 0; e < E; e += 4) {
            const f4 n0 = ld4(nrow + e), n1 = ld4(nrow + e + 4);
            const float nv[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
            MMX_UNROLL
            for (int c = 0; c < CP; ++c) {
                if (c < C) {
                    // dz[e .. e+3] of channel c: one aligned quad, or two when the interior is not quad-aligned (beyond E: zero halo)
                    const float* za = zrow + (size_t)c * TP * EP + e;
                    const f4 a0 = ld4(za);
                    float zq[8] = {a0.x, a0.y, a0.z, a0.w, 0.0f, 0.0f, 0.0f, 0.0f};
                    if (DLT > 0) { const f4 a1 = ld4(za + 4); zq[4] = a1.x; zq[5] = a1.y; zq[6] = a1.z; zq[7] = a1.w; }
                    MMX_UNROLL
                    for (int k = 0; k < 4; ++k) {
                        const float zv = zq[DLT + k];
                        MMX_UNROLL
                        for (int jj = 0; jj < 4; ++jj) acc[c][jj] = fmaf(zv, nv[k + jj], acc[c][jj]);
                    }
                }
            }
        }
    }
}
template <int CP>
MMX_D void conv_wgrad_acc(float (&acc)[CP][4], int tid, int nthr, const float* zpad, const float* npad, int C,
                          int kT, int kP, int qT, int qP, int TP, int EP, int ns, int T, int E) {
    switch (qP & 3) {
        case 0: conv_wgrad_acc_d<CP, 0>(acc, tid, nthr, zpad, npad, C, kT, kP, qT, qP, TP, EP, ns, T, E); break;
        case 1: conv_wgrad_acc_d<CP, 1>(acc, tid, nthr, zpad, npad, C, kT, kP, qT, qP, TP, EP, ns, T, E); break;
        case 2: conv_wgrad_acc_d<CP, 2>(acc, tid, nthr, zpad, npad, C, kT, kP, qT, qP, TP, EP, ns, T, E); break;
        default: conv_wgrad_acc_d<CP, 3>(acc, tid, nthr, zpad, npad, C, kT, kP, qT, qP, TP, EP, ns, T, E); break;
    }
}

// SE squeeze over (c, e) of one (sequence, frame): partial of part p.  `rows` points at row (s, c=0, t);
// consecutive channels are `cstride` floats apart.  mean: partial sum; max: partial max + first argmax (index c*E+e).
MMX_D void se_pool_part(const float* rows, int cstride, int C, int E, int p, int use_max, float* val, float* arg) {
    if (use_max) {
        float m = -INFINITY; int am = 0x7fffffff;
        for (int c = 0; c < C; ++c)
            for (int h = 4 * p; h < E; h += 4 * kParts) {
                const int n = imin(4, E - h);
                for (int k = 0; k < n; ++k) { const float v = rows[(size_t)c * cstride + h + k]; if (v > m) { m = v; am = c * E + h + k; } }
            }
        *val = m; *arg = (float)am;
    } else {
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s += row_part_sum(rows + (size_t)c * cstride, E, p);
        *val = s;
    }
}
// combine the kParts partials of one (sequence, frame)
MMX_D void se_pool_combine(const float* val, const float* arg, int n, int use_max, float* pool, float* amax) {
    if (use_max) {
        float m = val[0], a = arg[0];
        for (int p = 1; p < kParts; ++p)
            if (val[p] > m || (val[p] == m && arg[p] < a)) { m = val[p]; a = arg[p]; }
        *pool = m; *amax = a;
    } else {
        *pool = sum_parts(val) / (float)n;
    }
}

// SE backward for one use.  In: sm[o_dq + st] = dgate, gate, pool, z.  Out: sm[o_ds + st]; accumulates a_se1/a_se2.
template <class ExecT>
MMX_D void se_bwd_phases(ExecT& ex, float* sm, int T, int rr, int ns, int o_se1, int o_se2, int o_dq, int o_dz,
                         int o_gate, int o_pool, int o_z, int o_ds, int o_a1, int o_a2) {
    const int nthr = ex.nthr, nr = ns * T;
    ex.phase([&](int tid) {
        for (int r = tid; r < nr; r += nthr) { const float g = sm[o_gate + r]; sm[o_dq + r] *= g * (1.0f - g); }
    });
    ex.phase([&](int tid) {
        for (int i = tid; i < ns * rr; i += nthr) {
            const int s = i / rr, j = i - s * rr;
            float da = 0.0f;
            for (int t = 0; t < T; ++t) da = fmaf(sm[o_se2 + t * rr + j], sm[o_dq + s * T + t], da);
            sm[o_dz + i] = sm[o_z + i] > 0.0f ? da : 0.0f;
        }
    });
    ex.phase([&](int tid) {
        for (int r = tid; r < nr; r += nthr) {
            const int s = r / T, t = r - s * T;
            float v = 0.0f;
            for (int j = 0; j < rr; ++j) v = fmaf(sm[o_se1 + j * T + t], sm[o_dz + s * rr + j], v);
            sm[o_ds + r] = v;
        }
        for (int i = tid; i < T * rr; i += nthr) {
            {
                const int t = i / rr, j = i - t * rr;
                float v = 0.0f;
                for (int s = 0; s < ns; ++s) v = fmaf(sm[o_dq + s * T + t], fmaxf(sm[o_z + s * rr + j], 0.0f), v);
                sm[o_a2 + i] += v;
            }
            {
                const int j = i / T, t = i - j * T;
                float v = 0.0f;
                for (int s = 0; s < ns; ++s) v = fmaf(sm[o_dz + s * rr + j], sm[o_pool + s * T + t], v);
                sm[o_a1 + i] += v;
            }
        }
    });
}

// stage the stencil tables: forward wtab[ci][i][j][co] = K[co][ci][i][j]; backward-data
// wtab2[co][i'][j'][ci] = K[co][ci][kT-1-i'][kP-1-j'] (correlation with the flipped kernel)
MMX_D void conv_stage_tables(int tid, int nthr, float* sm, const ConvSmem& L, const ConvDims& d, const float* cw, bool bwd) {
    const int C = d.C, kT = d.kT, kP = d.kP, CP = L.CP;
    for (int i = tid; i < C * kT * kP * CP; i += nthr) {
        const int co = i % CP, rest = i / CP, j = rest % kP, ii = (rest / kP) % kT, ci = rest / (kP * kT);
        sm[L.wtab + i] = co < C ? cw[((size_t)(co * C + ci) * kT + ii) * kP + j] : 0.0f;
    }
    if (bwd)
        for (int i = tid; i < C * kT * kP * CP; i += nthr) {
            const int ci = i % CP, rest = i / CP, j = rest % kP, ii = (rest / kP) % kT, co = rest / (kP * kT);
            sm[L.wtab2 + i] = ci < C ? cw[((size_t)(co * C + ci) * kT + (kT - 1 - ii)) * kP + (kP - 1 - j)] : 0.0f;
        }
}

// ------------------------------------------------------------------------------------------
// half forward:  y = x + SE(drop(act(conv(LN(x)))))
// ------------------------------------------------------------------------------------------
struct ConvHalfFwdArgs {
    ConvDims d;
    Dropout dr;
    ConvHalfW w;
    const float* x;
    float* y;
    const float* aff;   // optional eval-mode BatchNorm affine after the activation: [scale[C] | shift[C]] (null: none)
    float* zout;        // bn_mode 1: pre-activation Z [B,C,T,E] (saved for the apply pass and the backward)
    double* bnsum;      // bn_mode 1: [sum(A)[C] | sum(A^2)[C]] accumulated over the batch
};

template <int ACT, int CP>
MMX_D void conv_half_fwd_body(Exec& ex, const ConvHalfFwdArgs& a) {
    const ConvDims& d = a.d;
    const ConvSmem L = conv_smem(d, false);
    constexpr int EW = CP <= 4 ? 8 : 4;
    float* sm = ex.smem;
    const int nthr = ex.nthr;
    const int C = d.C, T = d.T, E = d.E, S = d.S, rr = d.rr, PE = L.PE, TP = L.TP, EP = L.EP;
    const int E4 = (E + 3) >> 2;
    const Dropout dr = resolve_dropout(a.dr);
    const bool drop = d.training && dr.thresh != 0u;

    ex.phase([&](int tid) {
        conv_stage_tables(tid, nthr, sm, L, d, a.w.cw, false);
        for (int i = tid; i < CP; i += nthr) sm[L.cb + i] = i < C ? a.w.cb[i] : 0.0f;
        for (int i = tid; i < 16; i += nthr) sm[L.a_bn + i] = 0.0f;
        copy_vec(tid, nthr, sm + L.ln_g, a.w.ln_g, E); copy_vec(tid, nthr, sm + L.ln_b, a.w.ln_b, E);
        if (d.use_se) { copy_vec(tid, nthr, sm + L.se1, a.w.se1, rr * T); copy_vec(tid, nthr, sm + L.se2, a.w.se2, T * rr); }
    });
    const bool stats_pass = d.bn_mode == 1;

    const int ntiles = (d.B + S - 1) / S;
    for (int tile = ex.bid; tile < ntiles; tile += ex.nblk) {
        const long long seq0 = (long long)tile * S;
        const int ns = imin(S, d.B - (int)seq0), nr = ns * C * T;
        const float* xg = a.x + (size_t)seq0 * C * T * E;
        float* yg = stats_pass ? nullptr : a.y + (size_t)seq0 * C * T * E;
        float* zg = stats_pass ? a.zout + (size_t)seq0 * C * T * E : nullptr;

        ex.phase([&](int tid) {
            load_tile(tid, nthr, sm + L.bX, xg, nr, E, PE);
            for (int i = tid; i < (ns * C * TP * EP) >> 2; i += nthr) st4(sm + L.nPad + 4 * i, make_f4(0.f, 0.f, 0.f, 0.f));
        });
        ln_stats_phases(ex, sm, L.part, L.part2, L.mean, L.rstd, sm + L.bX, PE, nr, E);
        ex.phase([&](int tid) {
            for (int i = tid; i < nr * E; i += nthr) {
                const int r = i / E, e = i - r * E, sc = r / T, t = r - sc * T;
                sm[L.nPad + ((size_t)sc * TP + d.pT + t) * EP + d.pP + e] =
                    (sm[L.bX + r * PE + e] - sm[L.mean + r]) * sm[L.rstd + r] * sm[L.ln_g + e] + sm[L.ln_b + e];
            }
        });
        ex.phase([&](int tid) {
            conv_corr<CP, EW>(tid, nthr, sm + L.nPad, sm + L.wtab, C, C, d.kT, d.kP, TP, EP, ns, T, E,
                              [&](int s, int co, int t, int e, const float* v, int n) {
                                  const int r = (s * C + co) * T + t;
                                  const float b = sm[L.cb + co];
                                  float ks[4] = {1.0f, 1.0f, 1.0f, 1.0f};
                                  if (drop) dropout_quad(dr, d.site, ((uint64_t)(seq0 * C * T) + r) * E4 + (e >> 2), ks);
                                  const float sc = a.aff ? a.aff[co] : 1.0f, sh = a.aff ? a.aff[C + co] : 0.0f;
                                  for (int k = 0; k < n; ++k) {
                                      const float z = v[k] + b;
                                      if (stats_pass) zg[(size_t)r * E + e + k] = z;
                                      sm[L.bA + r * PE + e + k] = fmaf(act_fwd<ACT>(z) * ks[k], sc, sh);
                                  }
                              });
        });
        if (stats_pass) {
            // per-channel batch sums of A = act(Z): row partials, then one (channel, part) owner per accumulator
            ex.phase([&](int tid) {
                for (int i = tid; i < nr * kParts; i += nthr) {
                    const float* row = sm + L.bA + (i / kParts) * PE;
                    const int p = i % kParts;
                    float s1 = 0.0f, s2 = 0.0f;
                    for (int h = 4 * p; h < E; h += 4 * kParts) {
                        const int n = imin(4, E - h);
                        for (int k = 0; k < n; ++k) { const float v = row[h + k]; s1 += v; s2 = fmaf(v, v, s2); }
                    }
                    sm[L.part + i] = s1; sm[L.part2 + i] = s2;
                }
            });
            ex.phase([&](int tid) {
                for (int i = tid; i < C * kParts; i += nthr) {
                    const int co = i / kParts, p = i - co * kParts;
                    float s1 = 0.0f, s2 = 0.0f;
                    for (int s = 0; s < ns; ++s)
                        for (int t = 0; t < T; ++t) {
                            const int o = ((s * C + co) * T + t) * kParts + p;
                            s1 += sm[L.part + o]; s2 += sm[L.part2 + o];
                        }
                    smem_add(sm + L.a_bn + co, s1); smem_add(sm + L.a_bn + 8 + co, s2);
                }
            });
            continue;
        }
        if (d.use_se) {
            ex.phase([&](int tid) {
                for (int i = tid; i < ns * T * kParts; i += nthr) {
                    const int st = i / kParts, p = i - st * kParts, s = st / T, t = st - s * T;
                    se_pool_part(sm + L.bA + ((s * C) * T + t) * PE, T * PE, C, E, p, d.use_max, sm + L.part + i, sm + L.part2 + i);
                }
            });
            ex.phase([&](int tid) {
                for (int st = tid; st < ns * T; st += nthr)
                    se_pool_combine(sm + L.part + st * kParts, sm + L.part2 + st * kParts, C * E, d.use_max, sm + L.pool + st, sm + L.amax + st);
            });
            ex.phase([&](int tid) {
                for (int st = tid; st < ns * T; st += nthr) {
                    const int s = st / T, t = st - s * T;
                    sm[L.gate + st] = se_excite(sm + L.se1, sm + L.se2, sm + L.pool + s * T, T, rr, t, nullptr);
                }
            });
        }
        ex.phase([&](int tid) {
            if ((E & 3) == 0) {
                for (int i = tid; i < nr * E4; i += nthr) {
                    const int r = i / E4, q = i - r * E4, sc = r / T, t = r - sc * T, s = sc / C;
                    const float g = d.use_se ? sm[L.gate + s * T + t] : 1.0f;
                    const f4 xv = ld4(sm + L.bX + r * PE + 4 * q), av = ld4(sm + L.bA + r * PE + 4 * q);
                    st4(yg + (size_t)r * E + 4 * q, make_f4(fmaf(g, av.x, xv.x), fmaf(g, av.y, xv.y), fmaf(g, av.z, xv.z), fmaf(g, av.w, xv.w)));
                }
            } else {
                for (int i = tid; i < nr * E; i += nthr) {
                    const int r = i / E, e = i - r * E, sc = r / T, t = r - sc * T, s = sc / C;
                    const float g = d.use_se ? sm[L.gate + s * T + t] : 1.0f;
                    yg[(size_t)r * E + e] = fmaf(g, sm[L.bA + r * PE + e], sm[L.bX + r * PE + e]);
                }
            }
        });
    }
    if (stats_pass)
        ex.phase([&](int tid) {
            for (int i = tid; i < C; i += nthr) {
                red_add_f64(a.bnsum + i, (double)sm[L.a_bn + i]);
                red_add_f64(a.bnsum + C + i, (double)sm[L.a_bn + 8 + i]);
            }
        });
}

// ------------------------------------------------------------------------------------------
// half backward (forward recomputed from x)
// ------------------------------------------------------------------------------------------
struct ConvHalfBwdArgs {
    ConvDims d;
    Dropout dr;
    ConvHalfW w;     // parameters
    ConvHalfW g;     // gradient accumulators (global, +=)
    const float* x;  // half input            [B,C,T,E]
    const float* dy; // dL/d(half output)
    float* dx;       // dL/d(half input)
    const float* aff;   // fused mode: optional eval-mode BatchNorm affine [scale[C] | shift[C]]
    // bn_mode 2 (BatchNorm backward, second pass):
    const float* z;     // pre-activation saved by the forward statistics pass  [B,C,T,E]
    const float* gd;    // per (sequence, frame): SE gate and d(pool) from backward pass 1   [B,T,2]
    const float* bn;    // [scale | shift | xs | xo][C]
    const float* coef;  // [k1 | k2 | k3][C]
};

template <int CP>
struct ConvBwdRegs { float dK[CP][4]; };

template <int ACT, int CP>
MMX_D void conv_half_bwd_body(Exec& ex, const ConvHalfBwdArgs& a) {
    const ConvDims& d = a.d;
    const ConvSmem L = conv_smem(d, true);
    constexpr int EW = CP <= 4 ? 8 : 4;
    float* sm = ex.smem;
    const int nthr = ex.nthr;
    const int C = d.C, T = d.T, E = d.E, S = d.S, rr = d.rr, PE = L.PE, TP = L.TP, EP = L.EP;
    const int kT = d.kT, kP = d.kP, qT = kT - 1 - d.pT, qP = kP - 1 - d.pP;
    const int E4 = (E + 3) >> 2;
    const float invE = 1.0f / (float)E, invCE = 1.0f / (float)(C * E);
    const Dropout dr = resolve_dropout(a.dr);
    const bool drop = d.training && dr.thresh != 0u;
    const int nJB = (kP + 3) >> 2, n_items = C * kT * nJB, nsl_w = imax(1, nthr / n_items);
    const bool bn2 = d.bn_mode == 2;
    const bool se_here = d.use_se && !bn2;     // BatchNorm pass 2: the SE backward already ran in pass 1
    const float* asc = sm + L.a_bn;            // eval-mode BatchNorm affine (identity when absent), staged below
    const float* ash = sm + L.a_bn + 8;

    PerThread<ConvBwdRegs<CP>> regs(ex);

    ex.phase([&](int tid) {
        conv_stage_tables(tid, nthr, sm, L, d, a.w.cw, true);
        for (int i = tid; i < CP; i += nthr) { sm[L.cb + i] = i < C ? a.w.cb[i] : 0.0f; sm[L.a_cb + i] = 0.0f; }
        for (int i = tid; i < 8; i += nthr) {
            sm[L.a_bn + i] = (a.aff && i < C) ? a.aff[i] : 1.0f;
            sm[L.a_bn + 8 + i] = (a.aff && i < C) ? a.aff[C + i] : 0.0f;
        }
        copy_vec(tid, nthr, sm + L.ln_g, a.w.ln_g, E); copy_vec(tid, nthr, sm + L.ln_b, a.w.ln_b, E);
        zero_vec(tid, nthr, sm + L.a_lng, E); zero_vec(tid, nthr, sm + L.a_lnb, E);
        zero_vec(tid, nthr, sm + L.a_cw, C * C * kT * kP);
        if (d.use_se) {
            copy_vec(tid, nthr, sm + L.se1, a.w.se1, rr * T); copy_vec(tid, nthr, sm + L.se2, a.w.se2, T * rr);
            zero_vec(tid, nthr, sm + L.a_se1, rr * T); zero_vec(tid, nthr, sm + L.a_se2, T * rr);
        }
        ConvBwdRegs<CP>& rg = regs[tid];
        MMX_UNROLL
        for (int c = 0; c < CP; ++c)
            MMX_UNROLL
            for (int j = 0; j < 4; ++j) rg.dK[c][j] = 0.0f;
    });

    const int ntiles = (d.B + S - 1) / S;
    for (int tile = ex.bid; tile < ntiles; tile += ex.nblk) {
        const long long seq0 = (long long)tile * S;
        const int ns = imin(S, d.B - (int)seq0), nr = ns * C * T;
        const float* xg = a.x + (size_t)seq0 * C * T * E;
        const float* dyg = a.dy + (size_t)seq0 * C * T * E;
        float* dxg = a.dx + (size_t)seq0 * C * T * E;
        // X / dY rows: shared tile (pitch PE) or straight from global (pitch E)
        const float* xt = d.x_in_smem ? sm + L.bX : xg;
        const float* dt = d.x_in_smem ? sm + L.bD : dyg;
        const int xp = d.x_in_smem ? PE : E;

        ex.phase([&](int tid) {
            if (d.x_in_smem) {
                load_tile(tid, nthr, sm + L.bX, xg, nr, E, PE);
                load_tile(tid, nthr, sm + L.bD, dyg, nr, E, PE);
            }
            for (int i = tid; i < (ns * C * TP * EP) >> 2; i += nthr) {
                st4(sm + L.nPad + 4 * i, make_f4(0.f, 0.f, 0.f, 0.f));
                st4(sm + L.zPad + 4 * i, make_f4(0.f, 0.f, 0.f, 0.f));
            }
        });
        // ---------------- recompute the forward up to the pre-activation ----------------
        ln_stats_phases(ex, sm, L.part, L.part2, L.mean, L.rstd, xt, xp, nr, E);
        ex.phase([&](int tid) {
            for (int i = tid; i < nr * E; i += nthr) {
                const int r = i / E, e = i - r * E, sc = r / T, t = r - sc * T;
                sm[L.nPad + ((size_t)sc * TP + d.pT + t) * EP + d.pP + e] =
                    (xt[(size_t)r * xp + e] - sm[L.mean + r]) * sm[L.rstd + r] * sm[L.ln_g + e] + sm[L.ln_b + e];
            }
        });
        ex.phase([&](int tid) {
            if (bn2) {      // Z was saved by the forward statistics pass: no conv recompute
                const float* zg = a.z + (size_t)seq0 * C * T * E;
                if ((E & 3) == 0) {
                    const int E4q = E >> 2;
                    for (int i = tid; i < nr * E4q; i += nthr) {
                        const int r = i / E4q, q = i - r * E4q, sc = r / T, t = r - sc * T;
                        const f4 v = ld4(zg + (size_t)r * E + 4 * q);
                        float* o = sm + L.zPad + ((size_t)sc * TP + qT + t) * EP + qP + 4 * q;
                        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                    }
                } else {
                    for (int i = tid; i < nr * E; i += nthr) {
                        const int r = i / E, e = i - r * E, sc = r / T, t = r - sc * T;
                        sm[L.zPad + ((size_t)sc * TP + qT + t) * EP + qP + e] = zg[i];
                    }
                }
            } else {
                conv_corr<CP, EW>(tid, nthr, sm + L.nPad, sm + L.wtab, C, C, kT, kP, TP, EP, ns, T, E,
                                  [&](int s, int co, int t, int e, const float* v, int n) {
                                      float* zr = sm + L.zPad + ((size_t)(s * C + co) * TP + qT + t) * EP + qP + e;
                                      const float b = sm[L.cb + co];
                                      for (int k = 0; k < n; ++k) zr[k] = v[k] + b;
                                  });
            }
        });
        // ---------------- SE: squeeze of A = drop(act(Z)) and dgate = sum dY*A, per (s,t) ----------------
        if (se_here) {
            ex.phase([&](int tid) {
                for (int i = tid; i < ns * T * kParts; i += nthr) {
                    const int st = i / kParts, p = i - st * kParts, s = st / T, t = st - s * T;
                    float pv = d.use_max ? -INFINITY : 0.0f, dg = 0.0f;
                    int am = 0x7fffffff;
                    for (int c = 0; c < C; ++c) {
                        const int r = (s * C + c) * T + t;
                        const float* zr = sm + L.zPad + ((size_t)(s * C + c) * TP + qT + t) * EP + qP;
                        const float* dyr = dt + (size_t)r * xp;
                        for (int h = 4 * p; h < E; h += 4 * kParts) {
                            float ks[4] = {1.0f, 1.0f, 1.0f, 1.0f};
                            if (drop) dropout_quad(dr, d.site, ((uint64_t)(seq0 * C * T) + r) * E4 + (h >> 2), ks);
                            const int n = imin(4, E - h);
                            for (int k = 0; k < n; ++k) {
                                const float av = fmaf(act_fwd<ACT>(zr[h + k]) * ks[k], asc[c], ash[c]);
                                dg = fmaf(dyr[h + k], av, dg);
                                if (d.use_max) { if (av > pv) { pv = av; am = c * E + h + k; } }
                                else pv += av;
                            }
                        }
                    }
                    sm[L.part + i] = pv;
                    // part2: dgate partial (mean squeeze) or the argmax (max squeeze: dgate gets its own pass below)
                    sm[L.part2 + i] = d.use_max ? (float)am : dg;
                }
            });
            if (d.use_max) {
                // max squeeze: part2 carries the argmax, so the dgate partial sums need their own pass
                ex.phase([&](int tid) {
                    for (int st = tid; st < ns * T; st += nthr)
                        se_pool_combine(sm + L.part + st * kParts, sm + L.part2 + st * kParts, C * E, 1, sm + L.pool + st, sm + L.amax + st);
                });
                ex.phase([&](int tid) {
                    for (int i = tid; i < ns * T * kParts; i += nthr) {
                        const int st = i / kParts, p = i - st * kParts, s = st / T, t = st - s * T;
                        float dg = 0.0f;
                        for (int c = 0; c < C; ++c) {
                            const int r = (s * C + c) * T + t;
                            const float* zr = sm + L.zPad + ((size_t)(s * C + c) * TP + qT + t) * EP + qP;
                            const float* dyr = dt + (size_t)r * xp;
                            for (int h = 4 * p; h < E; h += 4 * kParts) {
                                float ks[4] = {1.0f, 1.0f, 1.0f, 1.0f};
                                if (drop) dropout_quad(dr, d.site, ((uint64_t)(seq0 * C * T) + r) * E4 + (h >> 2), ks);
                                const int n = imin(4, E - h);
                                for (int k = 0; k < n; ++k) dg = fmaf(dyr[h + k], fmaf(act_fwd<ACT>(zr[h + k]) * ks[k], asc[c], ash[c]), dg);
                            }
                        }
                        sm[L.part2 + i] = dg;
                    }
                });
                ex.phase([&](int tid) {
                    for (int st = tid; st < ns * T; st += nthr) sm[L.dq + st] = sum_parts(sm + L.part2 + st * kParts);
                });
            } else {
                ex.phase([&](int tid) {
                    for (int st = tid; st < ns * T; st += nthr) {
                        sm[L.pool + st] = sum_parts(sm + L.part + st * kParts) * invCE;
                        sm[L.dq + st] = sum_parts(sm + L.part2 + st * kParts);
                    }
                });
            }
            ex.phase([&](int tid) {
                for (int st = tid; st < ns * T; st += nthr) {
                    const int s = st / T, t = st - s * T;
                    sm[L.gate + st] = se_excite(sm + L.se1, sm + L.se2, sm + L.pool + s * T, T, rr, t, sm + L.z + s * rr);
                }
            });
            se_bwd_phases(ex, sm, T, rr, ns, L.se1, L.se2, L.dq, L.dz, L.gate, L.pool, L.z, L.ds, L.a_se1, L.a_se2);
        }
        // ---------------- dZ = (dY*gate + dpool) * mask * act'(Z), in place ----------------
        ex.phase([&](int tid) {
            for (int i = tid; i < nr * E4; i += nthr) {
                const int r = i / E4, q = i - r * E4, sc = r / T, t = r - sc * T, s = sc / C, c = sc - s * C;
                float* zr = sm + L.zPad + ((size_t)sc * TP + qT + t) * EP + qP + 4 * q;
                const float* dyr = dt + (size_t)r * xp + 4 * q;
                float g = se_here ? sm[L.gate + s * T + t] : 1.0f;
                float dsv = se_here ? sm[L.ds + s * T + t] : 0.0f;
                if (bn2 && d.use_se) { g = a.gd[((size_t)(seq0 + s) * T + t) * 2]; dsv = a.gd[((size_t)(seq0 + s) * T + t) * 2 + 1]; }
                const int amax = (se_here && d.use_max) ? (int)sm[L.amax + s * T + t] : -1;
                float ks[4] = {1.0f, 1.0f, 1.0f, 1.0f};
                if (drop) dropout_quad(dr, d.site, ((uint64_t)(seq0 * C * T) + r) * E4 + q, ks);
                const int n = imin(4, E - 4 * q);
                for (int k = 0; k < n; ++k) {
                    float av;
                    const float gp = act_fwd_grad<ACT>(zr[k], &av);
                    float da = dyr[k] * g;
                    if (d.use_se) {
                        if (se_here && d.use_max) { if (c * E + 4 * q + k == amax) da += dsv; }
                        else da = fmaf(dsv, invCE, da);
                    }
                    if (bn2) {          // BatchNorm backward: dA = k1 * (dR - mean(dR) - xhat * mean(dR*xhat))
                        const float xh = fmaf(av, a.bn[2 * C + c], a.bn[3 * C + c]);
                        da = a.coef[c] * (da - a.coef[C + c] - xh * a.coef[2 * C + c]);
                    } else {
                        da *= asc[c];
                    }
                    zr[k] = da * ks[k] * gp;
                }
            }
        });
        // ---------------- conv bias / weight gradients ----------------
        ex.phase([&](int tid) {
            for (int i = tid; i < nr * kParts; i += nthr) {
                const int r = i / kParts, p = i - r * kParts, sc = r / T, t = r - sc * T;
                sm[L.part + i] = row_part_sum(sm + L.zPad + ((size_t)sc * TP + qT + t) * EP + qP, E, p);
            }
            conv_wgrad_acc<CP>(regs[tid].dK, tid, nthr, sm + L.zPad, sm + L.nPad, C, kT, kP, qT, qP, TP, EP, ns, T, E);
        });
        // ---------------- dN = corr(dZ, flipped K)  (overwrites the normalised input) ----------------
        ex.phase([&](int tid) {
            for (int i = tid; i < C * kParts; i += nthr) {
                const int co = i / kParts, p = i - co * kParts;
                float s_ = 0.0f;
                for (int s = 0; s < ns; ++s)
                    for (int t = 0; t < T; ++t) s_ += sm[L.part + ((s * C + co) * T + t) * kParts + p];
                smem_add(sm + L.a_cb + co, s_);
            }
            conv_corr<CP, EW>(tid, nthr, sm + L.zPad, sm + L.wtab2, C, C, kT, kP, TP, EP, ns, T, E,
                              [&](int s, int ci, int t, int e, const float* v, int n) {
                                  float* o = sm + L.dN + ((s * C + ci) * T + t) * PE + e;
                                  for (int k = 0; k < n; ++k) o[k] = v[k];
                              });
        });
        // ---------------- LN backward ----------------
        ex.phase([&](int tid) {
            const int nsl = imax(1, nthr / E);   // row slices per column
            for (int it = tid; it < E * nsl; it += nthr) {
                const int sl = it / E, e = it - sl * E;
                float sg = 0.0f, sb = 0.0f;
                for (int r = sl; r < nr; r += nsl) {
                    const float dn = sm[L.dN + r * PE + e];
                    const float xh = (xt[(size_t)r * xp + e] - sm[L.mean + r]) * sm[L.rstd + r];
                    sg = fmaf(dn, xh, sg); sb += dn;
                }
                smem_add(sm + L.a_lng + e, sg); smem_add(sm + L.a_lnb + e, sb);
            }
            for (int i = tid; i < nr * kParts; i += nthr) {
                const int r = i / kParts, p = i - r * kParts;
                const float mu = sm[L.mean + r], rs = sm[L.rstd + r];
                const float* dn = sm + L.dN + r * PE;
                const float* xr = xt + (size_t)r * xp;
                float m1 = 0.0f, m2 = 0.0f;
                for (int h = 4 * p; h < E; h += 4 * kParts) {
                    const int n = imin(4, E - h);
                    for (int k = 0; k < n; ++k) {
                        const float dxh = dn[h + k] * sm[L.ln_g + h + k];
                        m1 += dxh; m2 = fmaf(dxh, (xr[h + k] - mu) * rs, m2);
                    }
                }
                sm[L.part + i] = m1;     // (the bias partials in `part` were consumed in the previous phase)
                sm[L.part2 + i] = m2;
            }
        });
        ex.phase([&](int tid) {
            for (int i = tid; i < nr * E; i += nthr) {
                const int r = i / E, e = i - r * E;
                const float mu = sm[L.mean + r], rs = sm[L.rstd + r];
                const float m1 = sum_parts(sm + L.part + r * kParts) * invE, m2 = sum_parts(sm + L.part2 + r * kParts) * invE;
                const float dxh = sm[L.dN + r * PE + e] * sm[L.ln_g + e];
                const float xh = (xt[(size_t)r * xp + e] - mu) * rs;
                dxg[(size_t)r * E + e] = dt[(size_t)r * xp + e] + rs * (dxh - m1 - xh * m2);
            }
        });
    }

    // ---------------- flush the CTA's gradient accumulators ----------------
    ex.phase([&](int tid) {
        const int sl = tid / n_items, item = tid - sl * n_items;
        if (sl < nsl_w) {
            const int jb = item % nJB, ii = (item / nJB) % kT, ci = item / (nJB * kT);
            ConvBwdRegs<CP>& rg = regs[tid];
            MMX_UNROLL
            for (int c = 0; c < CP; ++c)
                MMX_UNROLL
                for (int jj = 0; jj < 4; ++jj)
                    if (c < C && 4 * jb + jj < kP) smem_add(sm + L.a_cw + ((c * C + ci) * kT + ii) * kP + 4 * jb + jj, rg.dK[c][jj]);
        }
    });
    ex.phase([&](int tid) {
        for (int i = tid; i < C * C * kT * kP; i += nthr) red_add(a.g.cw + i, sm[L.a_cw + i]);
        for (int i = tid; i < C; i += nthr) red_add(a.g.cb + i, sm[L.a_cb + i]);
        for (int e = tid; e < E; e += nthr) { red_add(a.g.ln_g + e, sm[L.a_lng + e]); red_add(a.g.ln_b + e, sm[L.a_lnb + e]); }
        if (se_here)
            for (int i = tid; i < T * rr; i += nthr) { red_add(a.g.se1 + i, sm[L.a_se1 + i]); red_add(a.g.se2 + i, sm[L.a_se2 + i]); }
    });
}

}  // namespace mmx
