// The two callers either side of model(x) in the reference loops:
//   * step input: batch[:, :T, dim_used] (scaled) and batch[:, T:T+To, dim_used]   train_mixer_h36m.py:117-120,179
//   * evaluation: PCK histogram over the 299 thresholds of auc_pck_metric           utils/utils_mixer.py:20-45
#pragma once
#include "mmx_common.cuh"

namespace mmx {

struct WindowSplitArgs {
    const float* batch;     // [B, Ttot, Dfull]
    const int* dim_used;    // [D]
    float *x, *gt;          // [B, T, D], [B, To, D]
    int B, Ttot, Dfull, D, T, To;
    float x_scale, gt_scale;
};

MMX_D void window_split_body(Exec& ex, const WindowSplitArgs& a) {
    const int nthr = ex.nthr;
    ex.phase([&](int tid) {
        const long long per = (long long)(a.T + a.To) * a.D, total = (long long)a.B * per;
        for (long long i = (long long)ex.bid * nthr + tid; i < total; i += (long long)ex.nblk * nthr) {
            const long long b = i / per;
            const int r = (int)(i - b * per), t = r / a.D, d = r - t * a.D;
            const float v = a.batch[((size_t)b * a.Ttot + t) * a.Dfull + a.dim_used[d]];
            if (t < a.T) a.x[((size_t)b * a.T + t) * a.D + d] = v * a.x_scale;
            else a.gt[((size_t)b * a.To + (t - a.T)) * a.D + d] = v * a.gt_scale;
        }
    });
}

// hist[k] += #joints whose distance d satisfies thresh[k-1] < d <= thresh[k]  (k = 0: d <= thresh[0]; k = n: d > thresh[n-1]).
// PCK(thresh[k]) = cumsum(hist)[k] / n_joints.
struct PckHistArgs {
    const float *pred, *gt;   // [n_joints, 3]
    const float* thresh;      // [n] ascending
    int* hist;                // [n + 1]
    long long n_joints;
    int n;
};

MMX_D void pck_hist_body(Exec& ex, const PckHistArgs& a) {
    const int nthr = ex.nthr;
    int* sh = reinterpret_cast<int*>(ex.smem);
    float* st = ex.smem + a.n + 4;
    ex.phase([&](int tid) {
        for (int i = tid; i <= a.n; i += nthr) sh[i] = 0;
        for (int i = tid; i < a.n; i += nthr) st[i] = a.thresh[i];
    });
    ex.phase([&](int tid) {
        for (long long j = (long long)ex.bid * nthr + tid; j < a.n_joints; j += (long long)ex.nblk * nthr) {
            const float* p = a.pred + 3 * j;
            const float* g = a.gt + 3 * j;
            const float dx = p[0] - g[0], dy = p[1] - g[1], dz = p[2] - g[2];
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);      // torch.sqrt(torch.sum(diff**2)) : no fma contraction
            int lo = 0, hi = a.n;                                       // first k with dist <= thresh[k]
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (dist <= st[mid]) hi = mid; else lo = mid + 1; }
#if defined(MMX_HOST_EMU)
            sh[lo] += 1;
#else
            atomicAdd(sh + lo, 1);
#endif
        }
    });
    ex.phase([&](int tid) {
        for (int i = tid; i <= a.n; i += nthr)
            if (sh[i]) {
#if defined(MMX_HOST_EMU)
                a.hist[i] += sh[i];
#else
                atomicAdd(a.hist + i, sh[i]);
#endif
            }
    });
}


// ------------------------------------------------------------------------------------------
// BatchNorm2d bookkeeping between the two passes of a BatchNorm half (C <= 8 channels; one tiny CTA each):
//   finalize: batch sums -> [scale | shift | xs | xo], running statistics (momentum, unbiased variance), counter
//   coef    : backward sums -> [k1 | k2 | k3], d weight += sum dR*xhat, d bias += sum dR
// Both zero the sums they consumed, so the accumulating kernels always start from zero.
// (nn.BatchNorm2d semantics: torch/nn/modules/batchnorm.py; used at conv_mixer_model.py:115-116,141)
// ------------------------------------------------------------------------------------------
struct BnFinalizeArgs {
    double* sums;               // [sum A | sum A^2][C], zeroed on exit
    const float *w, *b;         // BatchNorm weight / bias [C]
    float *rm, *rv;             // running_mean / running_var [C] (updated)
    long long* nbt;             // num_batches_tracked (incremented)
    float* bn;                  // out: [scale | shift | xs | xo][C]
    double n;                   // elements per channel (B*T*E)
    float momentum, eps;
    int C;
};
MMX_D void bn_finalize_body(Exec& ex, const BnFinalizeArgs& a) {
    ex.phase([&](int tid) {
        const int c = ex.bid * ex.nthr + tid, C = a.C;
        if (c < C) {
            const double mean = a.sums[c] / a.n;
            double var = a.sums[C + c] / a.n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double rstd = 1.0 / sqrt(var + (double)a.eps);
            const double scale = (double)a.w[c] * rstd;
            a.bn[c] = (float)scale;
            a.bn[C + c] = (float)((double)a.b[c] - mean * scale);
            a.bn[2 * C + c] = (float)rstd;
            a.bn[3 * C + c] = (float)(-mean * rstd);
            a.rm[c] = a.rm[c] * (1.0f - a.momentum) + (float)mean * a.momentum;
            a.rv[c] = a.rv[c] * (1.0f - a.momentum) + (float)(var * (a.n / (a.n > 1.0 ? a.n - 1.0 : 1.0))) * a.momentum;
            a.sums[c] = 0.0; a.sums[C + c] = 0.0;
            if (c == 0) *a.nbt += 1;
        }
    });
}

struct BnCoefArgs {
    double* sums;               // [sum dR | sum dR*xhat][C], zeroed on exit
    const float* bn;            // [scale | shift | xs | xo][C]
    float* coef;                // out: [k1 | k2 | k3][C]
    float *gw, *gb;             // BatchNorm weight / bias gradient accumulators [C]
    double n;
    int C;
};
MMX_D void bn_coef_body(Exec& ex, const BnCoefArgs& a) {
    ex.phase([&](int tid) {
        const int c = ex.bid * ex.nthr + tid, C = a.C;
        if (c < C) {
            a.coef[c] = a.bn[c];
            a.coef[C + c] = (float)(a.sums[c] / a.n);
            a.coef[2 * C + c] = (float)(a.sums[C + c] / a.n);
            a.gw[c] += (float)a.sums[C + c];
            a.gb[c] += (float)a.sums[c];
            a.sums[c] = 0.0; a.sums[C + c] = 0.0;
        }
    });
}

}  // namespace mmx
