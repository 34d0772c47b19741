// MixerBlock forward / backward on the tensor cores (TF32 operands, fp32 accumulation): the "bf16/TF32" mode of the north
// star, selected with MmxMlpBlockDesc.precision = MMX_PREC_TF32 (NX = 1).  The NX = 3 build of the same kernels splits every
// operand into hi + lo TF32 parts ("3xTF32", fp32-grade results, opt-in for the FP32 mode with MMX_MLP_TC_FP32=1).
//
// Reference arithmetic: h36m/mlp_mixer.py:138-164 (MixerBlock.forward), :6-34 (SELayer), :44-96 (MlpBlock).
//
// Why warp-level MMA and not tcgen05 here: at the shapes this variant serves (T = 10, tok = 20, H, ch <= 50) one
// sequence is a 10 x 50 tile and the four contractions have K = 10 / 20 / 50 / 50.  The tensor work is < 5 % of the
// block's instruction stream; what costs time is everything between the contractions (LayerNorm, Mish, dropout, SE,
// the residuals).  mma.sync keeps every intermediate in the registers of the warp that produced it -- the accumulator
// fragment of one contraction IS the A fragment of the next (with the K index permuted on the weight side) -- so a whole
// block runs with four small shared-memory round trips and no data dependency between the warps of a CTA.  A tcgen05 formulation would need 128-row
// tiles staged in shared memory for every A operand and a TMEM -> register load in front of every epilogue.
//
// Execution model: one WARP owns a group of 3 sequences (30 rows, padded to two m16 tiles) from load to store.
//   token half  ("T orientation", rows = hidden column h, cols = frame t): X^T fragments are read from the warp's shared
//                x tile (forward) or straight from global memory (backward: three shared tiles per warp instead of four ->
//                up to 8 warps per SM), LN1 statistics are column sums (register adds + 3 shuffles), token fc1 -> act -> fc2 chain in
//                registers, SE squeeze / excitation in registers, X1 = X + g*Y, LN2 statistics, xhat2 scattered to shared.
//   channel half ("H orientation", rows = (sequence, frame), cols = h): A fragments of xhat2 from shared, LN2's affine is
//                folded into the weights (V1' = V1*gamma2, c1' = c1 + V1 beta2), fc1 -> act -> fc2 chain in registers,
//                SE through a 32-float shared exchange, residual, coalesced 64-bit stores.
//   backward:    forward recomputed; dG2 / dxhat2 chained in registers; the weight gradients are MMAs whose K dimension
//                is the group's 32 rows (operands transposed through the warp's shared tiles), added to CTA-shared
//                accumulators under one lock per 16-row slice; bias gradients ride along as a column of ones in the
//                B operand; dgamma2 / dbeta2 / dV1 are derived from the folded gradient at flush time.  The token
//                weight gradients (K = h) accumulate in 24 registers per lane for the whole kernel.
//   scheduling:  one CTA per SM, the warps needed for the batch spread over all SMs; every warp of a CTA runs the same number
//                of iterations (a warp without a group runs a dead one) and the CTA re-aligns ONCE per iteration
//                (TC_ALIGN points, swept on the GPU: profiles/r1zc_alignment_sweeps.txt) -- the loop body is ~270 KB of SASS,
//                warps that walk it together share the instruction-cache lines they fetch.
//
// This file is plain CUDA (no host emulator build): the CPU suite cannot execute mma.sync.  GPU parity tests:
// tests/test_gpu_mlp_tc.py.
#pragma once
#include "mmx_common.cuh"
#include "mmx_mlp.cuh"

#if !defined(MMX_HOST_EMU)
namespace mmx {
namespace tc {

constexpr int kT = 10, kTok = 20;
constexpr int kSeq = 3;            // sequences per warp group
constexpr int kRows = 32;          // rows of a group tile (30 valid)
constexpr int kPA = 52;            // activation tile pitch: == 4 (mod 8) -> conflict-free A-fragment reads and T-layout scatters
constexpr int kPW = 60;            // channel weight pitch (zero padded to 56 + 4)
constexpr int kHP = 56;            // padded H / ch extent (7 n8 tiles)
constexpr int kNT = 7;
constexpr int kPAcc = 56;          // pitch of the CTA-shared dV accumulators
constexpr int kAccRows = 52;       // their rows (H, ch <= 50)
constexpr int kTile = kRows * kPA; // 1664 floats
constexpr int kPT = 24;            // pitch of the token weight-gradient operand tiles [64][24]
constexpr int kMaxRR = 2;          // SE bottleneck widths served by this variant (seq_len // r_se)
constexpr int kFwdWarps = 8, kBwdWarps = 8;

struct Smem {
    int v1, v2, c1, c2, g1, b1, tb1, tb2, w1f, w2f, w2g, w1g, se1, se2;
    int accV1, accV2, wln1, locks;
    int warp0, wstride;
    int xs, ns, gs, ds, pool, gate, dsh, rs2, wse, kb;   // offsets inside a warp's region
    int total;
};

MMX_HD Smem smem_layout(bool bwd, int nwarp) {
    Smem L;
    int o = 0;
    auto take = [&](int n) { int r = o; o += round_up(n, 4); return r; };
    L.v1 = take(kHP * kPW); L.v2 = take(kHP * kPW);
    L.c1 = take(64); L.c2 = take(64); L.g1 = take(64); L.b1 = take(64);
    L.tb1 = take(24); L.tb2 = take(16);
    L.w1f = take(6 * 64); L.w2f = take(6 * 64);
    L.se1 = take(kMaxRR * kT); L.se2 = take(kT * kMaxRR);
    if (bwd) {
        L.w2g = take(6 * 64); L.w1g = take(6 * 64);
        L.accV1 = take(kAccRows * kPAcc); L.accV2 = take(kAccRows * kPAcc); L.wln1 = take(128); L.locks = take(12);
    } else L.w2g = L.w1g = L.accV1 = L.accV2 = L.wln1 = L.locks = -1;
    L.warp0 = o;
    int w = 0;
    auto wtake = [&](int n) { int r = w; w += round_up(n, 4); return r; };
    // forward: xs (X, then X1) + ns (xhat2).  backward: X^T fragments come straight from global memory (L2), tiles ns, gs, ds
    if (bwd) { L.xs = -1; L.ns = wtake(kTile); L.gs = wtake(kTile); L.ds = wtake(kTile); }
    else { L.xs = wtake(kTile); L.ns = wtake(kTile); L.gs = L.ds = -1; }
    L.pool = wtake(32); L.gate = wtake(32);
    if (bwd) {
        L.dsh = wtake(32); L.rs2 = wtake(32);
        L.wse = wtake(kSeq * 2 * kMaxRR * kT);
    } else L.dsh = L.rs2 = L.wse = -1;
    L.kb = -1;
    L.wstride = w;
    L.total = o + nwarp * w + 64;      // slack: fragment reads run up to 12 floats past the last tile
    return L;
}

// ------------------------------------------------------------------------------------------ primitives
// round-to-nearest (ties away) TF32: the tensor core ignores the 13 low mantissa bits, so adding half a TF32 ulp to the bit
// pattern is the whole conversion (cvt.rna.tf32.f32 compiles to 4 instructions: it also handles Inf/NaN, which never occur here)
MMX_D uint32_t tf32(float x) { return __float_as_uint(x) + 0x1000u; }
MMX_D float tf32f(float x) { return __uint_as_float(tf32(x)); }

// D += A(16x8, row) * B(8x8, col), TF32 in, fp32 accumulate.  Lane (g = lane>>2, t4 = lane&3) holds
//   A: a0 (g, t4)  a1 (g+8, t4)  a2 (g, t4+4)  a3 (g+8, t4+4);   B: b0 (k = t4, n = g)  b1 (k = t4+4, n = g)
//   C: c0 (g, 2t4)  c1 (g, 2t4+1)  c2 (g+8, 2t4)  c3 (g+8, 2t4+1)
MMX_D void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Operand fragments.  NX = 1: one MMA on TF32-rounded operands (2e-3 mode).  NX = 3: "3xTF32" error compensation -- every
// operand is split x = hi + lo (hi = x truncated to TF32, lo = x - hi, exact in fp32) and the product is accumulated as
// lo*hi + hi*lo + hi*hi (the dropped lo*lo term and the TF32 truncation of lo are both ~2^-22 relative): fp32-grade
// contractions on the tensor cores, for the 1e-5 mode.
template <int NX> struct AFrag { uint32_t hi[4]; uint32_t lo[NX == 3 ? 4 : 1]; };
template <int NX> struct BFrag { uint32_t hi[2]; uint32_t lo[NX == 3 ? 2 : 1]; };
MMX_D void split3(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
// v0..v3 in A-fragment order
template <int NX>
MMX_D void prep_a(float v0, float v1, float v2, float v3, AFrag<NX>& a) {
    if (NX == 1) { a.hi[0] = tf32(v0); a.hi[1] = tf32(v1); a.hi[2] = tf32(v2); a.hi[3] = tf32(v3); }
    else { split3(v0, a.hi[0], a.lo[0]); split3(v1, a.hi[1], a.lo[1]); split3(v2, a.hi[2], a.lo[2 % (NX == 3 ? 4 : 1)]); split3(v3, a.hi[3], a.lo[3 % (NX == 3 ? 4 : 1)]); }
}
// weights: already rounded to TF32 by stage() when NX == 1
template <int NX>
MMX_D void prep_b_w(float b0, float b1, BFrag<NX>& b) {
    if (NX == 1) { b.hi[0] = __float_as_uint(b0); b.hi[1] = __float_as_uint(b1); }
    else { split3(b0, b.hi[0], b.lo[0]); split3(b1, b.hi[1], b.lo[1 % (NX == 3 ? 2 : 1)]); }
}
// activations used as a B operand
template <int NX>
MMX_D void prep_b_a(float b0, float b1, BFrag<NX>& b) {
    if (NX == 1) { b.hi[0] = tf32(b0); b.hi[1] = tf32(b1); }
    else { split3(b0, b.hi[0], b.lo[0]); split3(b1, b.hi[1], b.lo[1 % (NX == 3 ? 2 : 1)]); }
}
template <int NX>
MMX_D void mmaX(float (&d)[4], const AFrag<NX>& a, const BFrag<NX>& b) {
    if (NX == 3) {
        uint32_t alo[4] = {a.lo[0], a.lo[1 % (NX == 3 ? 4 : 1)], a.lo[2 % (NX == 3 ? 4 : 1)], a.lo[3 % (NX == 3 ? 4 : 1)]};
        mma8(d, alo, b.hi[0], b.hi[1]);
        mma8(d, a.hi, b.lo[0], b.lo[1 % (NX == 3 ? 2 : 1)]);
    }
    mma8(d, a.hi, b.hi[0], b.hi[1]);
}
// An accumulator tile re-used as the A operand of the next contraction: K slot j < 4 carries the tile's column 2j,
// slot j+4 column 2j+1 -- the B operand of that contraction is staged with the same permutation.
template <int NX>
MMX_D void a_from_c(const float (&c)[4], AFrag<NX>& a) { prep_a<NX>(c[0], c[2], c[1], c[3], a); }

MMX_D float ldf2x(const float* p) { return p[0]; }

// Philox4x32-7 -> 128 random bits
MMX_D u4 philox7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    MMX_UNROLL
    for (int i = 0; i < 7; ++i) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += W0; k1 += W1;
    }
    u4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
    return r;
}

// keep-bits of NCALL*8 elements held by this lane (16 random bits per element); `base` identifies the (site-local) tile
template <int NCALL>
MMX_D void keep_bits(const Dropout& dr, bool drop, uint32_t site, uint32_t base, int lane, uint32_t (&bits)[(NCALL + 3) / 4]) {
    MMX_UNROLL
    for (int w = 0; w < (NCALL + 3) / 4; ++w) bits[w] = 0xffffffffu;
    if (!drop) return;
    const uint32_t th = dr.thresh >> 16;
    MMX_UNROLL
    for (int w = 0; w < (NCALL + 3) / 4; ++w) bits[w] = 0u;
    MMX_UNROLL
    for (int c = 0; c < NCALL; ++c) {
        const u4 r = philox7((base * (uint32_t)NCALL + (uint32_t)c) * 32u + (uint32_t)lane, 0x7c0de5u, site ^ 0x5bd1e995u, dr.step, dr.seed_lo, dr.seed_hi);
        uint32_t b = 0;
        b |= ((r.x & 0xffffu) >= th) ? 1u : 0u;   b |= ((r.x >> 16) >= th) ? 2u : 0u;
        b |= ((r.y & 0xffffu) >= th) ? 4u : 0u;   b |= ((r.y >> 16) >= th) ? 8u : 0u;
        b |= ((r.z & 0xffffu) >= th) ? 16u : 0u;  b |= ((r.z >> 16) >= th) ? 32u : 0u;
        b |= ((r.w & 0xffffu) >= th) ? 64u : 0u;  b |= ((r.w >> 16) >= th) ? 128u : 0u;
        bits[c >> 2] |= b << (8 * (c & 3));
    }
}
template <int NW>
MMX_D float keepf(const uint32_t (&bits)[NW], int i, float scale) { return ((bits[i >> 5] >> (i & 31)) & 1u) ? scale : 0.0f; }

// column sums (over the 64 hidden rows held by the warp) of a T-orientation tile; column c <-> t = 8*(c>>1) + 2*t4 + (c&1)
MMX_D void colsum(const float (&v)[4][2][4], float (&s)[4]) {
    MMX_UNROLL
    for (int c = 0; c < 4; ++c) {
        float t = 0.0f;
        MMX_UNROLL
        for (int mt = 0; mt < 4; ++mt) t += v[mt][c >> 1][c & 1] + v[mt][c >> 1][2 + (c & 1)];
        s[c] = t;
    }
    MMX_UNROLL
    for (int c = 0; c < 4; ++c) {
        s[c] += __shfl_xor_sync(0xffffffffu, s[c], 4);
        s[c] += __shfl_xor_sync(0xffffffffu, s[c], 8);
        s[c] += __shfl_xor_sync(0xffffffffu, s[c], 16);
    }
}

// LayerNorm statistics over h of a T-orientation tile (invalid entries hold 0)
MMX_D void col_stats(const float (&x)[4][2][4], int H, float invH, int g, float (&mu)[4], float (&rs)[4]) {
    float s[4];
    colsum(x, s);
    float q[4][2][4];
    MMX_UNROLL
    for (int mt = 0; mt < 4; ++mt)
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) {
                const int h = 16 * mt + g + 8 * (r >> 1), c = nt * 2 + (r & 1);
                const float dv = x[mt][nt][r] - s[c] * invH;
                q[mt][nt][r] = h < H ? dv * dv : 0.0f;
            }
    float ss[4];
    colsum(q, ss);
    MMX_UNROLL
    for (int c = 0; c < 4; ++c) { mu[c] = s[c] * invH; rs[c] = 1.0f / sqrtf(ss[c] * invH + 1e-5f); }
}

struct Ctx {
    const float* sm;      // CTA shared base
    Smem L;
    int H, ch, rr, use_se;
    float invH;
    int lane, g, t4;
    bool drop;
    Dropout dr;
    int site_base;
};

MMX_D int col_t(int c, int t4) { return 8 * (c >> 1) + 2 * t4 + (c & 1); }

// SE excitation in the T orientation: every lane holds pool[] of its 4 frames; z_j reduced over the quad
MMX_D void se_gates_T(const Ctx& cx, const float (&pool)[4], float (&z)[kMaxRR], float (&gate)[4]) {
    float q[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    MMX_UNROLL
    for (int j = 0; j < kMaxRR; ++j) {
        z[j] = 0.0f;
        if (j < cx.rr) {
            float p = 0.0f;
            MMX_UNROLL
            for (int c = 0; c < 4; ++c) { const int t = col_t(c, cx.t4); if (t < kT) p = fmaf(cx.sm[cx.L.se1 + j * kT + t], pool[c], p); }
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            p += __shfl_xor_sync(0xffffffffu, p, 2);
            z[j] = p;
            const float rz = fmaxf(p, 0.0f);
            MMX_UNROLL
            for (int c = 0; c < 4; ++c) { const int t = col_t(c, cx.t4); if (t < kT) q[c] = fmaf(cx.sm[cx.L.se2 + t * cx.rr + j], rz, q[c]); }
        }
    }
    MMX_UNROLL
    for (int c = 0; c < 4; ++c) gate[c] = col_t(c, cx.t4) < kT ? sigmoidf_(q[c]) : 0.0f;
}

// X^T fragments of one sequence from the warp's x tile (rows rb..rb+9): x[mt][nt][r] = X[t][h], h = 16mt+g+8(r>>1), t = 8nt+2t4+(r&1)
MMX_D void load_xT(const float* tile, int rb, int H, int g, int t4, float (&x)[4][2][4]) {
    MMX_UNROLL
    for (int mt = 0; mt < 4; ++mt)
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) {
                const int h = 16 * mt + g + 8 * (r >> 1), t = 8 * nt + 2 * t4 + (r & 1);
                x[mt][nt][r] = (h < H && t < kT) ? tile[(rb + t) * kPA + h] : 0.0f;
            }
}

// the same fragments straight from global memory (xseq = first float of the sequence, null for a dead sequence): for a
// fixed register the 8 lanes of a row-group read 32 contiguous bytes -> every sector fetched is fully used
MMX_D void load_xT_global(const float* xseq, int H, int g, int t4, float (&x)[4][2][4]) {
    MMX_UNROLL
    for (int mt = 0; mt < 4; ++mt)
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) {
                const int h = 16 * mt + g + 8 * (r >> 1), t = 8 * nt + 2 * t4 + (r & 1);
                x[mt][nt][r] = (xseq != nullptr && h < H && t < kT) ? __ldg(xseq + t * H + h) : 0.0f;
            }
}

struct TokW {            // token-MLP B fragments + biases of this lane
    float w1[2][3][2], w2[3][2][2];
    float b1v[3][2], b2v[2][2];
};
MMX_D void load_tokw(const Ctx& cx, TokW& w) {
    MMX_UNROLL
    for (int kk = 0; kk < 2; ++kk)
        MMX_UNROLL
        for (int nt = 0; nt < 3; ++nt) {
            const float2 v = *reinterpret_cast<const float2*>(cx.sm + cx.L.w1f + (kk * 3 + nt) * 64 + cx.lane * 2);
            w.w1[kk][nt][0] = v.x; w.w1[kk][nt][1] = v.y;
        }
    MMX_UNROLL
    for (int kk = 0; kk < 3; ++kk)
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt) {
            const float2 v = *reinterpret_cast<const float2*>(cx.sm + cx.L.w2f + (kk * 2 + nt) * 64 + cx.lane * 2);
            w.w2[kk][nt][0] = v.x; w.w2[kk][nt][1] = v.y;
        }
    MMX_UNROLL
    for (int nt = 0; nt < 3; ++nt) { w.b1v[nt][0] = cx.sm[cx.L.tb1 + 8 * nt + 2 * cx.t4]; w.b1v[nt][1] = cx.sm[cx.L.tb1 + 8 * nt + 2 * cx.t4 + 1]; }
    MMX_UNROLL
    for (int nt = 0; nt < 2; ++nt) { w.b2v[nt][0] = cx.sm[cx.L.tb2 + 8 * nt + 2 * cx.t4]; w.b2v[nt][1] = cx.sm[cx.L.tb2 + 8 * nt + 2 * cx.t4 + 1]; }
}

// token fc1 for one m16 tile of hidden rows: u = b1 + N1 W1^T, N1 = xhat*gamma1 + beta1 (xhat already normalised)
template <int NX>
MMX_D void token_fc1(const Ctx& cx, const TokW& w, const float (&xh)[2][4], int mt, float (&u)[3][4]) {
    const float ga0 = cx.sm[cx.L.g1 + 16 * mt + cx.g], ga1 = cx.sm[cx.L.g1 + 16 * mt + cx.g + 8];
    const float be0 = cx.sm[cx.L.b1 + 16 * mt + cx.g], be1 = cx.sm[cx.L.b1 + 16 * mt + cx.g + 8];
    MMX_UNROLL
    for (int nt = 0; nt < 3; ++nt) { u[nt][0] = w.b1v[nt][0]; u[nt][1] = w.b1v[nt][1]; u[nt][2] = w.b1v[nt][0]; u[nt][3] = w.b1v[nt][1]; }
    MMX_UNROLL
    for (int kk = 0; kk < 2; ++kk) {
        AFrag<NX> a;
        prep_a<NX>(fmaf(xh[kk][0], ga0, be0), fmaf(xh[kk][2], ga1, be1), fmaf(xh[kk][1], ga0, be0), fmaf(xh[kk][3], ga1, be1), a);
        MMX_UNROLL
        for (int nt = 0; nt < 3; ++nt) { BFrag<NX> b; prep_b_w<NX>(w.w1[kk][nt][0], w.w1[kk][nt][1], b); mmaX<NX>(u[nt], a, b); }
    }
}
template <int NX>
MMX_D void token_fc2(const TokW& w, const float (&gv)[3][4], float (&y)[2][4]) {
    MMX_UNROLL
    for (int nt = 0; nt < 2; ++nt) { y[nt][0] = w.b2v[nt][0]; y[nt][1] = w.b2v[nt][1]; y[nt][2] = w.b2v[nt][0]; y[nt][3] = w.b2v[nt][1]; }
    MMX_UNROLL
    for (int kk = 0; kk < 3; ++kk) {
        AFrag<NX> a;
        a_from_c<NX>(gv[kk], a);
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt) { BFrag<NX> b; prep_b_w<NX>(w.w2[kk][nt][0], w.w2[kk][nt][1], b); mmaX<NX>(y[nt], a, b); }
    }
}

// Token half forward of one sequence.  in: x tile rows rb..; out: x = X1 (T orientation), LN2 statistics of its 4 frames.
template <int ACT, int NX>
MMX_D void token_fwd(const Ctx& cx, const TokW& w, const float* xtile, int rb, const float* xglobal, bool from_global, uint32_t seq,
                     float (&x)[4][2][4], float (&mu2)[4], float (&rs2)[4]) {
    if (from_global) load_xT_global(xglobal, cx.H, cx.g, cx.t4, x);
    else load_xT(xtile, rb, cx.H, cx.g, cx.t4, x);
    float mu[4], rs[4];
    col_stats(x, cx.H, cx.invH, cx.g, mu, rs);
    uint32_t bits0[2], bits1[1];
    keep_bits<6>(cx.dr, cx.drop, cx.site_base + 0, seq, cx.lane, bits0);
    keep_bits<4>(cx.dr, cx.drop, cx.site_base + 1, seq, cx.lane, bits1);
    float y[4][2][4];
    MMX_UNROLL
    for (int mt = 0; mt < 4; ++mt) {
        float xh[2][4];
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) { const int c = nt * 2 + (r & 1); xh[nt][r] = (x[mt][nt][r] - mu[c]) * rs[c]; }
        float u[3][4];
        token_fc1<NX>(cx, w, xh, mt, u);
        MMX_UNROLL
        for (int nt = 0; nt < 3; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) u[nt][r] = act_fwd<ACT>(u[nt][r]) * keepf(bits0, mt * 12 + nt * 4 + r, cx.dr.scale);
        token_fc2<NX>(w, u, y[mt]);
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) {
                const int h = 16 * mt + cx.g + 8 * (r >> 1);
                y[mt][nt][r] = h < cx.H ? y[mt][nt][r] * keepf(bits1, mt * 8 + nt * 4 + r, cx.dr.scale) : 0.0f;
            }
    }
    float gate[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    if (cx.use_se) {
        float pool[4], z[kMaxRR];
        colsum(y, pool);
        MMX_UNROLL
        for (int c = 0; c < 4; ++c) pool[c] *= cx.invH;
        se_gates_T(cx, pool, z, gate);
    }
    MMX_UNROLL
    for (int mt = 0; mt < 4; ++mt)
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) x[mt][nt][r] = fmaf(gate[nt * 2 + (r & 1)], y[mt][nt][r], x[mt][nt][r]);
    col_stats(x, cx.H, cx.invH, cx.g, mu2, rs2);
}

// scatter a T-orientation tile into rows rb.. of a [row][h] shared tile (valid entries only)
MMX_D void scatter_T(float* tile, int rb, int H, int g, int t4, const float (&v)[4][2][4]) {
    MMX_UNROLL
    for (int mt = 0; mt < 4; ++mt)
        MMX_UNROLL
        for (int nt = 0; nt < 2; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) {
                const int h = 16 * mt + g + 8 * (r >> 1), t = 8 * nt + 2 * t4 + (r & 1);
                if (h < H && t < kT) tile[(rb + t) * kPA + h] = v[mt][nt][r];
            }
}

// coalesced load of a group's x rows (nseq*10*H contiguous floats) into the pitched tile; dead sequences -> 0
MMX_D void load_group(float* tile, const float* src, int nvalid /*floats*/, int H, int lane) {
    const int total = kSeq * kT * H;       // H even: 64-bit accesses
    for (int e = 2 * lane; e < total; e += 64) {
        const int row = e / H, col = e - row * H;
        float2 v = make_float2(0.0f, 0.0f);
        if (e < nvalid) v = __ldg(reinterpret_cast<const float2*>(src + e));
        *reinterpret_cast<float2*>(tile + row * kPA + col) = v;
    }
}

// stage weights (all threads of the CTA).  Channel weights are rounded to TF32 once here.
MMX_D void stage(float* sm, const Smem& L, const MlpDims& d, const MlpBlockW& w, bool bwd, bool round_w, int tid, int nthr) {
    const int H = d.H, ch = d.ch, rr = d.rr;
    auto rw = [&](float v) { return round_w ? tf32f(v) : v; };
    for (int i = tid; i < L.total; i += nthr) sm[i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < kHP * kPW; i += nthr) {
        const int r = i / kPW, c = i - r * kPW;
        if (r < ch && c < H) sm[L.v1 + i] = rw(w.cw1[r * H + c] * w.ln2_g[c]);      // V1'[c][h] = V1[c][h] * gamma2[h]
        if (r < H && c < ch) sm[L.v2 + i] = rw(w.cw2[r * ch + c]);
    }
    for (int i = tid; i < 64; i += nthr) {
        if (i < ch) {
            float s = w.cb1[i];
            for (int h = 0; h < H; ++h) s = fmaf(w.cw1[i * H + h], w.ln2_b[h], s);       // c1' = c1 + V1 beta2
            sm[L.c1 + i] = s;
        }
        if (i < H) { sm[L.c2 + i] = w.cb2[i]; sm[L.g1 + i] = w.ln1_g[i]; sm[L.b1 + i] = w.ln1_b[i]; }
    }
    for (int i = tid; i < 24; i += nthr) sm[L.tb1 + i] = i < kTok ? w.tb1[i] : 0.0f;
    for (int i = tid; i < 16; i += nthr) sm[L.tb2 + i] = i < kT ? w.tb2[i] : 0.0f;
    for (int i = tid; i < 6 * 64; i += nthr) {
        const int f = i >> 6, lane = (i >> 1) & 31, j = i & 1, g = lane >> 2, t4 = lane & 3;
        {   // fc1 B: (kk < 2, nt < 3): W1[k = 8nt+g][t = 8kk+2t4+j]
            const int kk = f / 3, nt = f - kk * 3, k = 8 * nt + g, t = 8 * kk + 2 * t4 + j;
            sm[L.w1f + i] = (k < kTok && t < kT) ? rw(w.tw1[k * kT + t]) : 0.0f;
        }
        {   // fc2 B: (kk < 3, nt < 2): W2[t = 8nt+g][k = 8kk+2t4+j]
            const int kk = f / 2, nt = f - kk * 2, t = 8 * nt + g, k = 8 * kk + 2 * t4 + j;
            sm[L.w2f + i] = (t < kT && k < kTok) ? rw(w.tw2[t * kTok + k]) : 0.0f;
        }
        if (bwd) {
            {   // dG1 = dYt W2: (kk < 2, nt < 3): W2[t = 8kk+2t4+j][k = 8nt+g]
                const int kk = f / 3, nt = f - kk * 3, t = 8 * kk + 2 * t4 + j, k = 8 * nt + g;
                sm[L.w2g + i] = (t < kT && k < kTok) ? rw(w.tw2[t * kTok + k]) : 0.0f;
            }
            {   // dN1 = dU1 W1: (kk < 3, nt < 2): W1[k = 8kk+2t4+j][t = 8nt+g]
                const int kk = f / 2, nt = f - kk * 2, k = 8 * kk + 2 * t4 + j, t = 8 * nt + g;
                sm[L.w1g + i] = (k < kTok && t < kT) ? rw(w.tw1[k * kT + t]) : 0.0f;
            }
        }
    }
    if (d.use_se) {
        for (int i = tid; i < rr * kT; i += nthr) { sm[L.se1 + i] = w.se1[i]; sm[L.se2 + i] = w.se2[i]; }
    }
    __syncthreads();
}

// SE for the H orientation: lane L < kSeq handles sequence L serially.  pool/dg: [10] of that sequence in shared.
// Writes gate[t]; with BWD also dsh[t] = d pool / H and accumulates the SE weight gradients into the lane's slot.
template <bool BWD>
MMX_D void se_rows(const Ctx& cx, const float* pool, const float* dg, float* gate, float* dsh, bool accumulate, float* wse) {
    float z[kMaxRR], da[kMaxRR];
    MMX_UNROLL
    for (int j = 0; j < kMaxRR; ++j) {
        z[j] = 0.0f; da[j] = 0.0f;
        if (j < cx.rr)
            for (int t = 0; t < kT; ++t) z[j] = fmaf(cx.sm[cx.L.se1 + j * kT + t], pool[t], z[j]);
    }
    for (int t = 0; t < kT; ++t) {
        float q = 0.0f;
        MMX_UNROLL
        for (int j = 0; j < kMaxRR; ++j)
            if (j < cx.rr) q = fmaf(cx.sm[cx.L.se2 + t * cx.rr + j], fmaxf(z[j], 0.0f), q);
        const float gt = sigmoidf_(q);
        gate[t] = gt;
        if (BWD) {
            const float dq = dg[t] * gt * (1.0f - gt);
            dsh[t] = dq;
            MMX_UNROLL
            for (int j = 0; j < kMaxRR; ++j)
                if (j < cx.rr) da[j] = fmaf(cx.sm[cx.L.se2 + t * cx.rr + j], dq, da[j]);
        }
    }
    if (BWD) {
        float dz[kMaxRR];
        MMX_UNROLL
        for (int j = 0; j < kMaxRR; ++j) dz[j] = z[j] > 0.0f ? da[j] : 0.0f;
        for (int t = 0; t < kT; ++t) {
            const float dq = dsh[t];
            float ds = 0.0f;
            MMX_UNROLL
            for (int j = 0; j < kMaxRR; ++j)
                if (j < cx.rr) {
                    ds = fmaf(cx.sm[cx.L.se1 + j * kT + t], dz[j], ds);
                    if (accumulate) {
                        wse[j * kT + t] += dz[j] * pool[t];                                   // dS1[j][t]
                        wse[kMaxRR * kT + t * cx.rr + j] += dq * fmaxf(z[j], 0.0f);            // dS2[t][j]
                    }
                }
            dsh[t] = ds * cx.invH;
        }
    }
}

// ------------------------------------------------------------------------------------------ channel-half building blocks
// u[mt][nt] = c1' + xhat2 V1'^T      (A from the ns tile, natural K order)
template <int NX>
MMX_D void channel_fc1(const Ctx& cx, const float* ns, float (&u)[2][kNT][4]) {
    MMX_UNROLL
    for (int nt = 0; nt < kNT; ++nt) {
        const float2 b = *reinterpret_cast<const float2*>(cx.sm + cx.L.c1 + 8 * nt + 2 * cx.t4);
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt) { u[mt][nt][0] = b.x; u[mt][nt][1] = b.y; u[mt][nt][2] = b.x; u[mt][nt][3] = b.y; }
    }
    MMX_UNROLL
    for (int kk = 0; kk < kNT; ++kk) {
        AFrag<NX> a[2];
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt) {
            const float* p = ns + (16 * mt + cx.g) * kPA + 8 * kk + cx.t4;
            prep_a<NX>(p[0], p[8 * kPA], p[4], p[8 * kPA + 4], a[mt]);
        }
        MMX_UNROLL
        for (int nt = 0; nt < kNT; ++nt) {
            const float* q = cx.sm + cx.L.v1 + (8 * nt + cx.g) * kPW + 8 * kk + cx.t4;
            BFrag<NX> b;
            prep_b_w<NX>(q[0], q[4], b);
            mmaX<NX>(u[0][nt], a[0], b);
            mmaX<NX>(u[1][nt], a[1], b);
        }
    }
}
// out[mt][nt] (+)= in[mt][kk] (chained) * W, with B(k, n) = W[(8nt+g)*kPW + 8kk+2t4 (+1)]     ("NT": W rows are the outputs)
template <int NX>
MMX_D void chain_nt(const Ctx& cx, const float* W, const float (&in)[2][kNT][4], float (&out)[2][kNT][4]) {
    MMX_UNROLL
    for (int kk = 0; kk < kNT; ++kk) {
        AFrag<NX> a[2];
        a_from_c<NX>(in[0][kk], a[0]); a_from_c<NX>(in[1][kk], a[1]);
        MMX_UNROLL
        for (int nt = 0; nt < kNT; ++nt) {
            const float2 bv = *reinterpret_cast<const float2*>(W + (8 * nt + cx.g) * kPW + 8 * kk + 2 * cx.t4);
            BFrag<NX> b;
            prep_b_w<NX>(bv.x, bv.y, b);
            mmaX<NX>(out[0][nt], a[0], b);
            mmaX<NX>(out[1][nt], a[1], b);
        }
    }
}
// out[mt][nt] += in[mt][kk] (chained) * W, with B(k, n) = W[(8kk+2t4 (+1))*kPW + 8nt+g]        ("NN": W rows are the contraction)
template <int NX>
MMX_D void chain_nn(const Ctx& cx, const float* W, const float (&in)[2][kNT][4], float (&out)[2][kNT][4]) {
    MMX_UNROLL
    for (int kk = 0; kk < kNT; ++kk) {
        AFrag<NX> a[2];
        a_from_c<NX>(in[0][kk], a[0]); a_from_c<NX>(in[1][kk], a[1]);
        MMX_UNROLL
        for (int nt = 0; nt < kNT; ++nt) {
            const float* q = W + (8 * kk + 2 * cx.t4) * kPW + 8 * nt + cx.g;
            BFrag<NX> b;
            prep_b_w<NX>(q[0], q[kPW], b);
            mmaX<NX>(out[0][nt], a[0], b);
            mmaX<NX>(out[1][nt], a[1], b);
        }
    }
}
// H-orientation tile -> shared [row][col] (cols < kPA), optional column of ones at `ones_col`
MMX_D void store_H(float* tile, const Ctx& cx, const float (&v)[2][kNT][4], int ones_col) {
    MMX_UNROLL
    for (int mt = 0; mt < 2; ++mt)
        MMX_UNROLL
        for (int nt = 0; nt < kNT; ++nt) {
            const int col = 8 * nt + 2 * cx.t4;
            if (col < kPA) {
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) {
                    float2 o = make_float2(v[mt][nt][2 * hi], v[mt][nt][2 * hi + 1]);
                    if (col == ones_col) o.x = 1.0f;
                    if (col + 1 == ones_col) o.y = 1.0f;
                    *reinterpret_cast<float2*>(tile + (16 * mt + cx.g + 8 * hi) * kPA + col) = o;
                }
            }
        }
}
// row sums over the 56 columns held by the quad
MMX_D void rowsum(float (&p)[2][2]) {
    MMX_UNROLL
    for (int mt = 0; mt < 2; ++mt)
        MMX_UNROLL
        for (int hi = 0; hi < 2; ++hi) {
            p[mt][hi] += __shfl_xor_sync(0xffffffffu, p[mt][hi], 1);
            p[mt][hi] += __shfl_xor_sync(0xffffffffu, p[mt][hi], 2);
        }
}

// ------------------------------------------------------------------------------------------ forward kernel
template <int ACT, int NX>
__global__ void __launch_bounds__(kFwdWarps * 32, 1) mlp_block_fwd_tc_kernel(const MlpBlockFwdArgs a) {
    extern __shared__ float4 mmx_tc_smem_raw[];
    float* sm = reinterpret_cast<float*>(mmx_tc_smem_raw);
    const MlpDims& d = a.d;
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Smem L = smem_layout(false, nwarp);
    stage(sm, L, d, a.w, false, NX == 1, threadIdx.x, blockDim.x);

    Ctx cx;
    cx.sm = sm; cx.L = L; cx.H = d.H; cx.ch = d.ch; cx.rr = d.rr; cx.use_se = d.use_se; cx.invH = 1.0f / (float)d.H;
    cx.lane = lane; cx.g = lane >> 2; cx.t4 = lane & 3;
    cx.dr = resolve_dropout(a.dr); cx.drop = d.training && cx.dr.thresh != 0u; cx.site_base = d.site_base;
    const int H = d.H;
    float* ws = sm + L.warp0 + warp * L.wstride;
    float* xs = ws + L.xs; float* ns = ws + L.ns; float* pools = ws + L.pool; float* gates = ws + L.gate;
    TokW tw;
    load_tokw(cx, tw);

    const int amask = d.align_mask >> 16;
#define TC_ALIGN(i) do { if (amask >> (i) & 1) __syncthreads(); else __syncwarp(); } while (0)
    const int groups = (d.B + kSeq - 1) / kSeq;
    // Every warp of the CTA runs the same number of iterations (a warp without a group runs a dead one: all loads and
    // stores predicated off) and the CTA re-aligns at the phase boundaries: the loop body is far larger than the
    // instruction caches, so warps that walk it together share every fetched line instead of missing on it one by one.
    const int per_iter = gridDim.x * nwarp, n_iter = (groups + per_iter - 1) / per_iter;
    MMX_NOUNROLL
    for (int it = 0; it < n_iter; ++it) {
        const int grp = it * per_iter + blockIdx.x * nwarp + warp;
        const int seq0 = grp * kSeq, nseq = imax(0, imin(kSeq, d.B - seq0));
        TC_ALIGN(0);
        load_group(xs, a.x + (size_t)seq0 * kT * H, nseq * kT * H, H, lane);
        TC_ALIGN(1);
        MMX_NOUNROLL
        for (int s = 0; s < kSeq; ++s) {
            float x[4][2][4], mu2[4], rs2[4];
            token_fwd<ACT, NX>(cx, tw, xs, s * kT, nullptr, false, (uint32_t)(seq0 + s), x, mu2, rs2);
            scatter_T(xs, s * kT, H, cx.g, cx.t4, x);                                   // X1 (same lane reads / writes each address)
            MMX_UNROLL
            for (int mt = 0; mt < 4; ++mt)
                MMX_UNROLL
                for (int nt = 0; nt < 2; ++nt)
                    MMX_UNROLL
                    for (int r = 0; r < 4; ++r) { const int c = nt * 2 + (r & 1); x[mt][nt][r] = (x[mt][nt][r] - mu2[c]) * rs2[c]; }
            scatter_T(ns, s * kT, H, cx.g, cx.t4, x);                                   // xhat2
            TC_ALIGN(2);
        }
        TC_ALIGN(3);
        // ---------------- channel half on the group's 32 rows ----------------
        float u[2][kNT][4];
        channel_fc1<NX>(cx, ns, u);
        {
            uint32_t bits2[2];
            keep_bits<7>(cx.dr, cx.drop, cx.site_base + 2, (uint32_t)grp, lane, bits2);
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int nt = 0; nt < kNT; ++nt)
                    MMX_UNROLL
                    for (int r = 0; r < 4; ++r) u[mt][nt][r] = act_fwd<ACT>(u[mt][nt][r]) * keepf(bits2, (mt * kNT + nt) * 4 + r, cx.dr.scale);
        }
        float y2[2][kNT][4];
        MMX_UNROLL
        for (int nt = 0; nt < kNT; ++nt) {
            const float2 b = *reinterpret_cast<const float2*>(sm + L.c2 + 8 * nt + 2 * cx.t4);
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt) { y2[mt][nt][0] = b.x; y2[mt][nt][1] = b.y; y2[mt][nt][2] = b.x; y2[mt][nt][3] = b.y; }
        }
        chain_nt<NX>(cx, sm + L.v2, u, y2);
        {
            uint32_t bits3[2];
            keep_bits<7>(cx.dr, cx.drop, cx.site_base + 3, (uint32_t)grp, lane, bits3);
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int nt = 0; nt < kNT; ++nt)
                    MMX_UNROLL
                    for (int r = 0; r < 4; ++r) y2[mt][nt][r] *= keepf(bits3, (mt * kNT + nt) * 4 + r, cx.dr.scale);
        }
        float gt[2][2] = {{1.0f, 1.0f}, {1.0f, 1.0f}};
        if (d.use_se) {
            float p[2][2];
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) {
                    float t = 0.0f;
                    MMX_UNROLL
                    for (int nt = 0; nt < kNT; ++nt) t += y2[mt][nt][2 * hi] + y2[mt][nt][2 * hi + 1];
                    p[mt][hi] = t;
                }
            rowsum(p);
            if (cx.t4 == 0) {
                MMX_UNROLL
                for (int mt = 0; mt < 2; ++mt)
                    MMX_UNROLL
                    for (int hi = 0; hi < 2; ++hi) pools[16 * mt + cx.g + 8 * hi] = p[mt][hi] * cx.invH;
            }
            TC_ALIGN(4);
            if (lane < kSeq) se_rows<false>(cx, pools + lane * kT, nullptr, gates + lane * kT, nullptr, false, nullptr);
            TC_ALIGN(5);
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) gt[mt][hi] = gates[16 * mt + cx.g + 8 * hi];
        }
        // OUT = X1 + gate * Y2
        float* yg = a.y + (size_t)seq0 * kT * H;
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt)
            MMX_UNROLL
            for (int hi = 0; hi < 2; ++hi) {
                const int row = 16 * mt + cx.g + 8 * hi;
                if (row < nseq * kT) {
                    MMX_UNROLL
                    for (int nt = 0; nt < kNT; ++nt) {
                        const int col = 8 * nt + 2 * cx.t4;
                        if (col < H) {
                            const float2 x1 = *reinterpret_cast<const float2*>(xs + row * kPA + col);
                            float2 o;
                            o.x = fmaf(gt[mt][hi], y2[mt][nt][2 * hi], x1.x);
                            o.y = fmaf(gt[mt][hi], y2[mt][nt][2 * hi + 1], x1.y);
                            *reinterpret_cast<float2*>(yg + (size_t)row * H + col) = o;
                        }
                    }
                }
            }
        TC_ALIGN(6);
    }
}

#undef TC_ALIGN
// ------------------------------------------------------------------------------------------ backward pieces
// acc[m][n] += sum_{row < 32} At[row][m] * Bt[row][n]   (m, n < 56; CTA-shared accumulator, one lock per 16 rows of m)
template <int NX>
MMX_D void wgrad_rows(float* acc, unsigned int* locks, const float* At, const float* Bt, const Ctx& cx, int warp) {
    MMX_NOUNROLL
    for (int i = 0; i < 4; ++i) {
        const int mt = (i + warp) & 3;
        float c[kNT][4];
        MMX_UNROLL
        for (int nt = 0; nt < kNT; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.0f; }
        MMX_UNROLL
        for (int kk = 0; kk < 4; ++kk) {
            const float* pa = At + (8 * kk + cx.t4) * kPA + 16 * mt + cx.g;
            AFrag<NX> af;
            prep_a<NX>(pa[0], pa[8], pa[4 * kPA], pa[4 * kPA + 8], af);
            const float* pb = Bt + (8 * kk + cx.t4) * kPA + cx.g;
            MMX_UNROLL
            for (int nt = 0; nt < kNT; ++nt) { BFrag<NX> b; prep_b_a<NX>(pb[8 * nt], pb[4 * kPA + 8 * nt], b); mmaX<NX>(c[nt], af, b); }
        }
        warp_lock(locks + mt, cx.lane);
        MMX_UNROLL
        for (int hi = 0; hi < 2; ++hi) {
            const int row = 16 * mt + cx.g + 8 * hi;
            if (row < kAccRows) {
                MMX_UNROLL
                for (int nt = 0; nt < kNT; ++nt) {
                    float2* p = reinterpret_cast<float2*>(acc + row * kPAcc + 8 * nt + 2 * cx.t4);
                    float2 v = *p;
                    v.x += c[nt][2 * hi]; v.y += c[nt][2 * hi + 1];
                    *p = v;
                }
            }
        }
        warp_unlock(locks + mt, cx.lane);
    }
}

// acc[nt] += sum_{h < 64} A[h][m = t] * B[h][n = k]   (token weight gradients; operand tiles [64][kPT])
template <int NX>
MMX_D void wgrad_tok(float (&acc)[3][4], const float* A, const float* Bm, const Ctx& cx) {
    MMX_UNROLL
    for (int kk = 0; kk < 8; ++kk) {
        const float* pa = A + (8 * kk + cx.t4) * kPT + cx.g;
        AFrag<NX> af;
        prep_a<NX>(pa[0], pa[8], pa[4 * kPT], pa[4 * kPT + 8], af);
        const float* pb = Bm + (8 * kk + cx.t4) * kPT + cx.g;
        MMX_UNROLL
        for (int nt = 0; nt < 3; ++nt) { BFrag<NX> b; prep_b_a<NX>(pb[8 * nt], pb[4 * kPT + 8 * nt], b); mmaX<NX>(acc[nt], af, b); }
    }
}
// token tile [h][k or t] (C layout, NT n8 tiles) -> shared [64][kPT]; optional ones column
template <int NTL>
MMX_D void store_tok(float* buf, const Ctx& cx, int mt, const float (&v)[NTL][4], int ones_col) {
    MMX_UNROLL
    for (int nt = 0; nt < NTL; ++nt)
        MMX_UNROLL
        for (int hi = 0; hi < 2; ++hi) {
            const int col = 8 * nt + 2 * cx.t4;
            float2 o = make_float2(v[nt][2 * hi], v[nt][2 * hi + 1]);
            if (col == ones_col) o.x = 1.0f;
            if (col + 1 == ones_col) o.y = 1.0f;
            *reinterpret_cast<float2*>(buf + (16 * mt + cx.g + 8 * hi) * kPT + col) = o;
        }
}

// ------------------------------------------------------------------------------------------ backward kernel
template <int ACT, int NX>
__global__ void __launch_bounds__(kBwdWarps * 32, 1) mlp_block_bwd_tc_kernel(const MlpBlockBwdArgs a) {
    extern __shared__ float4 mmx_tc_smem_raw[];
    float* sm = reinterpret_cast<float*>(mmx_tc_smem_raw);
    const MlpDims& d = a.d;
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Smem L = smem_layout(true, nwarp);
    stage(sm, L, d, a.w, true, NX == 1, threadIdx.x, blockDim.x);

    Ctx cx;
    cx.sm = sm; cx.L = L; cx.H = d.H; cx.ch = d.ch; cx.rr = d.rr; cx.use_se = d.use_se; cx.invH = 1.0f / (float)d.H;
    cx.lane = lane; cx.g = lane >> 2; cx.t4 = lane & 3;
    cx.dr = resolve_dropout(a.dr); cx.drop = d.training && cx.dr.thresh != 0u; cx.site_base = d.site_base;
    const int H = d.H, ch = d.ch, g = cx.g, t4 = cx.t4;
    float* ws = sm + L.warp0 + warp * L.wstride;
    float* ns = ws + L.ns; float* gs = ws + L.gs; float* ds = ws + L.ds;
    float* pools = ws + L.pool; float* gates = ws + L.gate; float* dshs = ws + L.dsh; float* rs2s = ws + L.rs2;
    float* wse = ws + L.wse; float* wln1 = sm + L.wln1;      // LN1 weight gradients: CTA-shared, under locks[8]
    unsigned int* locks = reinterpret_cast<unsigned int*>(sm + L.locks);
    if (!d.use_se) { gates[lane] = lane < kSeq * kT ? 1.0f : 0.0f; }      // dsh stays 0
    TokW tw;
    load_tokw(cx, tw);
    float aW2[3][4], aW1[3][4];        // dW2[t][k] (+ db2 in column 20), dW1^T[t][k] (+ db1 in row 10): whole-kernel accumulators
    MMX_UNROLL
    for (int nt = 0; nt < 3; ++nt)
        MMX_UNROLL
        for (int r = 0; r < 4; ++r) { aW2[nt][r] = 0.0f; aW1[nt][r] = 0.0f; }
    __syncwarp();

    const int amask = d.align_mask & 0xffff;
#define TC_ALIGN(i) do { if (amask >> (i) & 1) __syncthreads(); else __syncwarp(); } while (0)
    const int groups = (d.B + kSeq - 1) / kSeq;
    // uniform iteration count + CTA re-alignment at the phase boundaries: see the forward kernel
    const int per_iter = gridDim.x * nwarp, n_iter = (groups + per_iter - 1) / per_iter;
    MMX_NOUNROLL
    for (int it = 0; it < n_iter; ++it) {
        const int grp = it * per_iter + blockIdx.x * nwarp + warp;
        const int seq0 = grp * kSeq, nseq = imax(0, imin(kSeq, d.B - seq0));
        TC_ALIGN(1);
        // ---------------- A: token half forward -> xhat2 (ns), rstd2 ----------------
        MMX_NOUNROLL
        for (int s = 0; s < kSeq; ++s) {
            float x[4][2][4], mu2[4], rs2[4];
            token_fwd<ACT, NX>(cx, tw, nullptr, 0, s < nseq ? a.x + (size_t)(seq0 + s) * kT * H : nullptr, true, (uint32_t)(seq0 + s), x, mu2, rs2);
            MMX_UNROLL
            for (int mt = 0; mt < 4; ++mt)
                MMX_UNROLL
                for (int nt = 0; nt < 2; ++nt)
                    MMX_UNROLL
                    for (int r = 0; r < 4; ++r) { const int c = nt * 2 + (r & 1); x[mt][nt][r] = (x[mt][nt][r] - mu2[c]) * rs2[c]; }
            scatter_T(ns, s * kT, H, g, t4, x);
            if (g == 0) {
                MMX_UNROLL
                for (int c = 0; c < 4; ++c) { const int t = col_t(c, t4); if (t < kT) rs2s[s * kT + t] = rs2[c]; }
            }
            TC_ALIGN(2);
        }
        ns[lane * kPA + H] = 1.0f;          // column of ones: W~[c][H] = sum_rows dU2[row][c] = dc1
        TC_ALIGN(3);
        // ---------------- B: channel half forward + backward ----------------
        const float* dyg = a.dy + (size_t)seq0 * kT * H;
        float u2[2][kNT][4];
        channel_fc1<NX>(cx, ns, u2);
        uint32_t bits2[2], bits3[2];
        keep_bits<7>(cx.dr, cx.drop, cx.site_base + 2, (uint32_t)grp, lane, bits2);
        keep_bits<7>(cx.dr, cx.drop, cx.site_base + 3, (uint32_t)grp, lane, bits3);
        float dO[2][kNT][4];
        float gt[2][2], dsh[2][2];
        {
            float g2[2][kNT][4];
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int nt = 0; nt < kNT; ++nt)
                    MMX_UNROLL
                    for (int r = 0; r < 4; ++r) g2[mt][nt][r] = act_fwd<ACT>(u2[mt][nt][r]) * keepf(bits2, (mt * kNT + nt) * 4 + r, cx.dr.scale);
            store_H(gs, cx, g2, ch);        // G2 (+ ones column at c = ch: dV2[h][ch] = dc2[h])
            float y2[2][kNT][4];
            MMX_UNROLL
            for (int nt = 0; nt < kNT; ++nt) {
                const float2 b = *reinterpret_cast<const float2*>(sm + L.c2 + 8 * nt + 2 * t4);
                MMX_UNROLL
                for (int mt = 0; mt < 2; ++mt) { y2[mt][nt][0] = b.x; y2[mt][nt][1] = b.y; y2[mt][nt][2] = b.x; y2[mt][nt][3] = b.y; }
            }
            chain_nt<NX>(cx, sm + L.v2, g2, y2);
            float p1[2][2], p2[2][2];
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) {
                    const int row = 16 * mt + g + 8 * hi;
                    float s1 = 0.0f, s2 = 0.0f;
                    MMX_UNROLL
                    for (int nt = 0; nt < kNT; ++nt) {
                        const int col = 8 * nt + 2 * t4;
                        float2 dv = make_float2(0.0f, 0.0f);
                        if (row < nseq * kT && col < H) dv = __ldg(reinterpret_cast<const float2*>(dyg + (size_t)row * H + col));
                        dO[mt][nt][2 * hi] = dv.x; dO[mt][nt][2 * hi + 1] = dv.y;
                        const float ya = y2[mt][nt][2 * hi] * keepf(bits3, (mt * kNT + nt) * 4 + 2 * hi, cx.dr.scale);
                        const float yb = y2[mt][nt][2 * hi + 1] * keepf(bits3, (mt * kNT + nt) * 4 + 2 * hi + 1, cx.dr.scale);
                        s1 += ya + yb;
                        s2 = fmaf(dv.x, ya, s2); s2 = fmaf(dv.y, yb, s2);
                    }
                    p1[mt][hi] = s1; p2[mt][hi] = s2;
                }
            if (d.use_se) {
                rowsum(p1); rowsum(p2);
                if (t4 == 0) {
                    MMX_UNROLL
                    for (int mt = 0; mt < 2; ++mt)
                        MMX_UNROLL
                        for (int hi = 0; hi < 2; ++hi) { pools[16 * mt + g + 8 * hi] = p1[mt][hi] * cx.invH; dshs[16 * mt + g + 8 * hi] = p2[mt][hi]; }
                }
                TC_ALIGN(4);
                if (lane < kSeq)
                    se_rows<true>(cx, pools + lane * kT, dshs + lane * kT, gates + lane * kT, dshs + lane * kT,      // dg in, d pool / H out (in place)
                                  seq0 + lane < d.B, wse + lane * 2 * kMaxRR * kT);
                TC_ALIGN(5);
            }
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) { gt[mt][hi] = gates[16 * mt + g + 8 * hi]; dsh[mt][hi] = dshs[16 * mt + g + 8 * hi]; }
        }
        // dY2 = (dOut*gate2 + ds2/H) * mask3   (in dO)
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt)
            MMX_UNROLL
            for (int nt = 0; nt < kNT; ++nt)
                MMX_UNROLL
                for (int r = 0; r < 4; ++r) {
                    const int col = 8 * nt + 2 * t4 + (r & 1);
                    const float v = fmaf(dO[mt][nt][r], gt[mt][r >> 1], dsh[mt][r >> 1]) * keepf(bits3, (mt * kNT + nt) * 4 + r, cx.dr.scale);
                    dO[mt][nt][r] = col < H ? v : 0.0f;
                }
        store_H(ds, cx, dO, -1);
        float dg2[2][kNT][4];
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt)
            MMX_UNROLL
            for (int nt = 0; nt < kNT; ++nt) { dg2[mt][nt][0] = dg2[mt][nt][1] = dg2[mt][nt][2] = dg2[mt][nt][3] = 0.0f; }
        chain_nn<NX>(cx, sm + L.v2, dO, dg2);                 // dG2 = dY2 V2
        TC_ALIGN(6);
        wgrad_rows<NX>(sm + L.accV2, locks + 0, ds, gs, cx, warp);        // dV2[h][c] += dY2^T G2
        // dU2 = dG2 * mask2 * act'(U2)   (in dg2)
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt)
            MMX_UNROLL
            for (int nt = 0; nt < kNT; ++nt)
                MMX_UNROLL
                for (int r = 0; r < 4; ++r) {
                    float av;
                    const float gp = act_fwd_grad<ACT>(u2[mt][nt][r], &av);
                    dg2[mt][nt][r] = dg2[mt][nt][r] * keepf(bits2, (mt * kNT + nt) * 4 + r, cx.dr.scale) * gp;
                }
        TC_ALIGN(7);
        store_H(ds, cx, dg2, -1);
        float dxh[2][kNT][4];
        MMX_UNROLL
        for (int mt = 0; mt < 2; ++mt)
            MMX_UNROLL
            for (int nt = 0; nt < kNT; ++nt) { dxh[mt][nt][0] = dxh[mt][nt][1] = dxh[mt][nt][2] = dxh[mt][nt][3] = 0.0f; }
        chain_nn<NX>(cx, sm + L.v1, dg2, dxh);                // d xhat2 = dU2 V1'
        TC_ALIGN(8);
        wgrad_rows<NX>(sm + L.accV1, locks + 4, ds, ns, cx, warp);        // W~[c][h] += dU2^T xhat2
        // LN2 backward: dX1 = dOut + rstd2 * (dxh - mean(dxh) - xhat2 * mean(dxh * xhat2))  -> gs
        {
            float m1[2][2], m2[2][2];
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) {
                    const int row = 16 * mt + g + 8 * hi;
                    float s1 = 0.0f, s2 = 0.0f;
                    MMX_UNROLL
                    for (int nt = 0; nt < kNT; ++nt) {
                        const int col = 8 * nt + 2 * t4;
                        float2 xv = make_float2(0.0f, 0.0f);
                        if (col < H) xv = *reinterpret_cast<const float2*>(ns + row * kPA + col);
                        u2[mt][nt][2 * hi] = xv.x; u2[mt][nt][2 * hi + 1] = xv.y;      // u2 is dead: holds xhat2 now
                        s1 += dxh[mt][nt][2 * hi] + dxh[mt][nt][2 * hi + 1];
                        s2 = fmaf(dxh[mt][nt][2 * hi], xv.x, s2); s2 = fmaf(dxh[mt][nt][2 * hi + 1], xv.y, s2);
                    }
                    m1[mt][hi] = s1; m2[mt][hi] = s2;
                }
            rowsum(m1); rowsum(m2);
            MMX_UNROLL
            for (int mt = 0; mt < 2; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) {
                    const int row = 16 * mt + g + 8 * hi;
                    const bool rv = row < nseq * kT;
                    const float rs = rs2s[row < kSeq * kT ? row : 0], a1 = m1[mt][hi] * cx.invH, a2 = m2[mt][hi] * cx.invH;
                    MMX_UNROLL
                    for (int nt = 0; nt < kNT; ++nt) {
                        const int col = 8 * nt + 2 * t4;
                        float2 dv = make_float2(0.0f, 0.0f);
                        if (rv && col < H) {
                            dv = __ldg(reinterpret_cast<const float2*>(dyg + (size_t)row * H + col));
                            dv.x += rs * (dxh[mt][nt][2 * hi] - a1 - u2[mt][nt][2 * hi] * a2);
                            dv.y += rs * (dxh[mt][nt][2 * hi + 1] - a1 - u2[mt][nt][2 * hi + 1] * a2);
                        }
                        if (col < kPA) *reinterpret_cast<float2*>(gs + row * kPA + col) = dv;
                    }
                }
        }
        TC_ALIGN(9);
        // ---------------- C: token half backward, one sequence at a time ----------------
        float gsum[4][2], bsum[4][2];
        MMX_UNROLL
        for (int mt = 0; mt < 4; ++mt) { gsum[mt][0] = gsum[mt][1] = bsum[mt][0] = bsum[mt][1] = 0.0f; }
        float* bufA = ds;      // [64][kPT]: G1 (+ ones at k = 20), later dU1
        float* bufB = ns;      // [64][kPT]: dYt, later N1 (+ ones at t = 10)
        MMX_NOUNROLL
        for (int s = 0; s < kSeq; ++s) {
            const int rb = s * kT;
            const bool live = seq0 + s < d.B;
            float mu1[4], rs1[4];
            float u1[4][3][4], y[4][2][4];
            const float* xseq = live ? a.x + (size_t)(seq0 + s) * kT * H : nullptr;
            uint32_t bits0[2], bits1[1];
            keep_bits<6>(cx.dr, cx.drop, cx.site_base + 0, (uint32_t)(seq0 + s), lane, bits0);
            keep_bits<4>(cx.dr, cx.drop, cx.site_base + 1, (uint32_t)(seq0 + s), lane, bits1);
            {
                float x[4][2][4];
                load_xT_global(xseq, H, g, t4, x);
                col_stats(x, H, cx.invH, g, mu1, rs1);
                MMX_UNROLL
                for (int mt = 0; mt < 4; ++mt) {
                    float xh[2][4];
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) { const int c = nt * 2 + (r & 1); xh[nt][r] = (x[mt][nt][r] - mu1[c]) * rs1[c]; }
                    token_fc1<NX>(cx, tw, xh, mt, u1[mt]);
                    float gv[3][4];
                    MMX_UNROLL
                    for (int nt = 0; nt < 3; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) gv[nt][r] = act_fwd<ACT>(u1[mt][nt][r]) * keepf(bits0, mt * 12 + nt * 4 + r, cx.dr.scale);
                    store_tok<3>(bufA, cx, mt, gv, kTok);
                    token_fc2<NX>(tw, gv, y[mt]);
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const int h = 16 * mt + g + 8 * (r >> 1);
                            y[mt][nt][r] = h < H ? y[mt][nt][r] * keepf(bits1, mt * 8 + nt * 4 + r, cx.dr.scale) : 0.0f;
                        }
                }
            }
            float gate[4] = {1.0f, 1.0f, 1.0f, 1.0f}, dsv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            {
                float d1[4][2][4];
                load_xT(gs, rb, H, g, t4, d1);            // dX1^T
                if (cx.use_se) {
                    float pool[4], z[kMaxRR];
                    colsum(y, pool);
                    MMX_UNROLL
                    for (int c = 0; c < 4; ++c) pool[c] *= cx.invH;
                    se_gates_T(cx, pool, z, gate);
                    float q[4][2][4], dgt[4];
                    MMX_UNROLL
                    for (int mt = 0; mt < 4; ++mt)
                        MMX_UNROLL
                        for (int nt = 0; nt < 2; ++nt)
                            MMX_UNROLL
                            for (int r = 0; r < 4; ++r) q[mt][nt][r] = d1[mt][nt][r] * y[mt][nt][r];
                    colsum(q, dgt);
                    float dq[4];
                    MMX_UNROLL
                    for (int c = 0; c < 4; ++c) dq[c] = dgt[c] * gate[c] * (1.0f - gate[c]);      // gate = 0 on invalid frames
                    float* wse_s = wse + s * 2 * kMaxRR * kT;
                    MMX_UNROLL
                    for (int j = 0; j < kMaxRR; ++j) {
                        if (j < cx.rr) {
                            float da = 0.0f;
                            MMX_UNROLL
                            for (int c = 0; c < 4; ++c) { const int t = col_t(c, t4); if (t < kT) da = fmaf(sm[L.se2 + t * cx.rr + j], dq[c], da); }
                            da += __shfl_xor_sync(0xffffffffu, da, 1);
                            da += __shfl_xor_sync(0xffffffffu, da, 2);
                            const float dz = z[j] > 0.0f ? da : 0.0f, rz = fmaxf(z[j], 0.0f);
                            MMX_UNROLL
                            for (int c = 0; c < 4; ++c) {
                                const int t = col_t(c, t4);
                                if (t < kT) {
                                    dsv[c] = fmaf(sm[L.se1 + j * kT + t], dz, dsv[c]);
                                    if (g == 0 && live) { wse_s[j * kT + t] += dz * pool[c]; wse_s[kMaxRR * kT + t * cx.rr + j] += dq[c] * rz; }
                                }
                            }
                        }
                    }
                }
                // dYt = (dX1*gate1 + ds1/H) * mask1   (in y)
                MMX_UNROLL
                for (int mt = 0; mt < 4; ++mt)
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const int h = 16 * mt + g + 8 * (r >> 1), c = nt * 2 + (r & 1);
                            const bool ok = h < H && col_t(c, t4) < kT;
                            y[mt][nt][r] = ok ? fmaf(d1[mt][nt][r], gate[c], dsv[c] * cx.invH) * keepf(bits1, mt * 8 + nt * 4 + r, cx.dr.scale) : 0.0f;
                        }
            }
            MMX_UNROLL
            for (int mt = 0; mt < 4; ++mt) store_tok<2>(bufB, cx, mt, y[mt], -1);
            TC_ALIGN(10);
            wgrad_tok<NX>(aW2, bufB, bufA, cx);               // dW2[t][k] += dYt^T G1 ; column 20 = db2
            // pass 2: dG1, dU1, dN1
            float dn[4][2][4];
            {
                float w2g[2][3][2], w1g[3][2][2];
                MMX_UNROLL
                for (int kk = 0; kk < 2; ++kk)
                    MMX_UNROLL
                    for (int nt = 0; nt < 3; ++nt) {
                        const float2 v = *reinterpret_cast<const float2*>(sm + L.w2g + (kk * 3 + nt) * 64 + lane * 2);
                        w2g[kk][nt][0] = v.x; w2g[kk][nt][1] = v.y;
                    }
                MMX_UNROLL
                for (int kk = 0; kk < 3; ++kk)
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt) {
                        const float2 v = *reinterpret_cast<const float2*>(sm + L.w1g + (kk * 2 + nt) * 64 + lane * 2);
                        w1g[kk][nt][0] = v.x; w1g[kk][nt][1] = v.y;
                    }
                MMX_UNROLL
                for (int mt = 0; mt < 4; ++mt) {
                    float dg1[3][4];
                    MMX_UNROLL
                    for (int nt = 0; nt < 3; ++nt) { dg1[nt][0] = dg1[nt][1] = dg1[nt][2] = dg1[nt][3] = 0.0f; }
                    MMX_UNROLL
                    for (int kk = 0; kk < 2; ++kk) {
                        AFrag<NX> af;
                        a_from_c<NX>(y[mt][kk], af);
                        MMX_UNROLL
                        for (int nt = 0; nt < 3; ++nt) { BFrag<NX> b; prep_b_w<NX>(w2g[kk][nt][0], w2g[kk][nt][1], b); mmaX<NX>(dg1[nt], af, b); }
                    }
                    MMX_UNROLL
                    for (int nt = 0; nt < 3; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            float av;
                            const float gp = act_fwd_grad<ACT>(u1[mt][nt][r], &av);
                            u1[mt][nt][r] = dg1[nt][r] * keepf(bits0, mt * 12 + nt * 4 + r, cx.dr.scale) * gp;      // dU1
                        }
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt) { dn[mt][nt][0] = dn[mt][nt][1] = dn[mt][nt][2] = dn[mt][nt][3] = 0.0f; }
                    MMX_UNROLL
                    for (int kk = 0; kk < 3; ++kk) {
                        AFrag<NX> af;
                        a_from_c<NX>(u1[mt][kk], af);
                        MMX_UNROLL
                        for (int nt = 0; nt < 2; ++nt) { BFrag<NX> b; prep_b_w<NX>(w1g[kk][nt][0], w1g[kk][nt][1], b); mmaX<NX>(dn[mt][nt], af, b); }
                    }
                }
            }
            TC_ALIGN(11);                                 // the dW2 MMAs have read bufA / bufB
            float xh[4][2][4];
            load_xT_global(xseq, H, g, t4, xh);
            MMX_UNROLL
            for (int mt = 0; mt < 4; ++mt) {
                store_tok<3>(bufA, cx, mt, u1[mt], -1);   // dU1
                const float ga0 = sm[L.g1 + 16 * mt + g], ga1 = sm[L.g1 + 16 * mt + g + 8];
                const float be0 = sm[L.b1 + 16 * mt + g], be1 = sm[L.b1 + 16 * mt + g + 8];
                float n1[2][4];
                MMX_UNROLL
                for (int nt = 0; nt < 2; ++nt)
                    MMX_UNROLL
                    for (int r = 0; r < 4; ++r) {
                        const int c = nt * 2 + (r & 1);
                        const bool ok = 16 * mt + g + 8 * (r >> 1) < H && col_t(c, t4) < kT;
                        xh[mt][nt][r] = ok ? (xh[mt][nt][r] - mu1[c]) * rs1[c] : 0.0f;
                        n1[nt][r] = ok ? fmaf(xh[mt][nt][r], (r >> 1) ? ga1 : ga0, (r >> 1) ? be1 : be0) : 0.0f;
                    }
                store_tok<2>(bufB, cx, mt, n1, kT);       // N1 (+ ones at t = 10: dW1^T[10][k] = db1[k])
            }
            TC_ALIGN(12);
            wgrad_tok<NX>(aW1, bufB, bufA, cx);               // dW1^T[t][k] += N1^T dU1
            // LN1 backward
            {
                float dxv[4][2][4], q[4][2][4], m1[4], m2[4];
                MMX_UNROLL
                for (int mt = 0; mt < 4; ++mt) {
                    const float ga0 = sm[L.g1 + 16 * mt + g], ga1 = sm[L.g1 + 16 * mt + g + 8];
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const bool ok = 16 * mt + g + 8 * (r >> 1) < H && col_t(nt * 2 + (r & 1), t4) < kT;
                            const float dnv = ok ? dn[mt][nt][r] : 0.0f;
                            gsum[mt][r >> 1] = fmaf(dnv, xh[mt][nt][r], gsum[mt][r >> 1]);
                            bsum[mt][r >> 1] += dnv;
                            dxv[mt][nt][r] = dnv * ((r >> 1) ? ga1 : ga0);
                            q[mt][nt][r] = dxv[mt][nt][r] * xh[mt][nt][r];
                        }
                }
                colsum(dxv, m1); colsum(q, m2);
                MMX_UNROLL
                for (int mt = 0; mt < 4; ++mt)
                    MMX_UNROLL
                    for (int nt = 0; nt < 2; ++nt)
                        MMX_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const int h = 16 * mt + g + 8 * (r >> 1), c = nt * 2 + (r & 1), t = col_t(c, t4);
                            if (h < H && t < kT) {
                                float* p = gs + (rb + t) * kPA + h;
                                *p += rs1[c] * (dxv[mt][nt][r] - m1[c] * cx.invH - xh[mt][nt][r] * m2[c] * cx.invH);
                            }
                        }
            }
            TC_ALIGN(13);                              // bufA / bufB are rewritten by the next sequence; CTA re-alignment
        }
        // group epilogue: LN1 weight gradients, dx copy-out
        MMX_UNROLL
        for (int mt = 0; mt < 4; ++mt)
            MMX_UNROLL
            for (int hi = 0; hi < 2; ++hi) {
                gsum[mt][hi] += __shfl_xor_sync(0xffffffffu, gsum[mt][hi], 1); gsum[mt][hi] += __shfl_xor_sync(0xffffffffu, gsum[mt][hi], 2);
                bsum[mt][hi] += __shfl_xor_sync(0xffffffffu, bsum[mt][hi], 1); bsum[mt][hi] += __shfl_xor_sync(0xffffffffu, bsum[mt][hi], 2);
            }
        warp_lock(locks + 8, lane);
        if (t4 == 0) {
            MMX_UNROLL
            for (int mt = 0; mt < 4; ++mt)
                MMX_UNROLL
                for (int hi = 0; hi < 2; ++hi) { wln1[16 * mt + g + 8 * hi] += gsum[mt][hi]; wln1[64 + 16 * mt + g + 8 * hi] += bsum[mt][hi]; }
        }
        warp_unlock(locks + 8, lane);
        {
            float* dxg = a.dx + (size_t)seq0 * kT * H;
            const int nvalid = nseq * kT * H;
            for (int e = 2 * lane; e < nvalid; e += 64) {
                const int row = e / H, col = e - row * H;
                *reinterpret_cast<float2*>(dxg + e) = *reinterpret_cast<const float2*>(gs + row * kPA + col);
            }
        }
        TC_ALIGN(14);
    }

#undef TC_ALIGN
    // ---------------- flush ----------------
    {   // per-warp register accumulators -> the warp's ns tile: [2][16][24]
        MMX_UNROLL
        for (int nt = 0; nt < 3; ++nt)
            MMX_UNROLL
            for (int r = 0; r < 4; ++r) {
                const int row = g + 8 * (r >> 1), col = 8 * nt + 2 * t4 + (r & 1);
                ns[row * 24 + col] = aW2[nt][r];
                ns[384 + row * 24 + col] = aW1[nt][r];
            }
    }
    __syncthreads();
    const int tid = threadIdx.x, nthr = blockDim.x;
    const float* accV1 = sm + L.accV1;
    const float* accV2 = sm + L.accV2;
    for (int i = tid; i < H * ch; i += nthr) { const int h = i / ch, c = i - h * ch; red_add(a.g.cw2 + i, accV2[h * kPAcc + c]); }
    for (int i = tid; i < ch * H; i += nthr) {
        const int c = i / H, h = i - c * H;
        red_add(a.g.cw1 + i, fmaf(a.w.ln2_g[h], accV1[c * kPAcc + h], a.w.ln2_b[h] * accV1[c * kPAcc + H]));
    }
    for (int h = tid; h < H; h += nthr) {
        red_add(a.g.cb2 + h, accV2[h * kPAcc + ch]);
        float sg = 0.0f, sb = 0.0f;
        for (int c = 0; c < ch; ++c) { const float v = a.w.cw1[c * H + h]; sg = fmaf(v, accV1[c * kPAcc + h], sg); sb = fmaf(v, accV1[c * kPAcc + H], sb); }
        red_add(a.g.ln2_g + h, sg); red_add(a.g.ln2_b + h, sb);
        red_add(a.g.ln1_g + h, wln1[h]); red_add(a.g.ln1_b + h, wln1[64 + h]);
    }
    for (int c = tid; c < ch; c += nthr) red_add(a.g.cb1 + c, accV1[c * kPAcc + H]);
    for (int i = tid; i < 2 * 384; i += nthr) {
        const int which = i / 384, j = i - which * 384, t = j / 24, k = j - t * 24;
        float v = 0.0f;
        for (int w = 0; w < nwarp; ++w) v += sm[L.warp0 + w * L.wstride + L.ns + i];
        if (which == 0) {
            if (t < kT && k < kTok) red_add(a.g.tw2 + t * kTok + k, v);
            else if (t < kT && k == kTok) red_add(a.g.tb2 + t, v);
        } else {
            if (t < kT && k < kTok) red_add(a.g.tw1 + k * kT + t, v);
            else if (t == kT && k < kTok) red_add(a.g.tb1 + k, v);
        }
    }
    if (d.use_se) {
        const int rr = d.rr;
        for (int i = tid; i < 2 * rr * kT; i += nthr) {
            const int which = i / (rr * kT), j = i - which * rr * kT;
            float v = 0.0f;
            for (int w = 0; w < nwarp; ++w)
                for (int s = 0; s < kSeq; ++s) v += sm[L.warp0 + w * L.wstride + L.wse + s * 2 * kMaxRR * kT + which * kMaxRR * kT + j];
            red_add((which == 0 ? a.g.se1 : a.g.se2) + j, v);
        }
    }
}

}  // namespace tc
}  // namespace mmx
#endif
