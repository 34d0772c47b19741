// Fused MPJPE loss forward+backward and multi-tensor (flat-buffer) Adam.
//
// Reference: h36m/utils/utils_mixer.py:48-53 (mpjpe_error) and torch.optim.Adam with coupled
// L2 weight decay as constructed at h36m/train_mixer_h36m.py:63.
#pragma once
#include "mmx_common.cuh"

namespace mmx {

// loss_sum += sum_j ||gt_j - pred_j||_2 ;  dpred = gscale * (pred - gt) / (||.|| * n_joints), 0 at zero distance.
// `loss_sum` must be zeroed by the caller; the mean is loss_sum / n_joints (done by `mpjpe_finish`
// or by the host wrapper).  dpred may be null (evaluation).
struct MpjpeArgs { const float *pred, *gt; float* dpred; float* loss_sum; long long n_joints; float gscale; };

MMX_D void mpjpe_body(Exec& ex, const MpjpeArgs& a) {
    float* sm = ex.smem;
    const int nthr = ex.nthr;
    const float inv_n = a.gscale / (float)a.n_joints;
    ex.phase([&](int tid) {
        float acc = 0.0f;
        for (long long j = (long long)ex.bid * nthr + tid; j < a.n_joints; j += (long long)ex.nblk * nthr) {
            const float* p = a.pred + 3 * j;
            const float* g = a.gt + 3 * j;
            const float dx = p[0] - g[0], dy = p[1] - g[1], dz = p[2] - g[2];
            const float nrm = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            acc += nrm;
            if (a.dpred) {
                const float s = nrm > 0.0f ? inv_n / nrm : 0.0f;
                float* d = a.dpred + 3 * j;
                d[0] = dx * s; d[1] = dy * s; d[2] = dz * s;
            }
        }
        sm[tid] = acc;
    });
    ex.phase([&](int tid) {   // tree levels handled by the first warp's worth of threads
        if (tid < 32) {
            float s = 0.0f;
            for (int i = tid; i < nthr; i += 32) s += sm[i];
            sm[tid] = s;
        }
    });
    ex.phase([&](int tid) {
        if (tid == 0) {
            float s = 0.0f;
            for (int i = 0; i < 32 && i < nthr; ++i) s += sm[i];
            red_add(a.loss_sum, s);
        }
    });
}

// Adam on flat fp32 buffers (every parameter of the model is a view into `p`):
//   g' = gscale*g + wd*p ; m += (g'-m)(1-b1) ; v = b2 v + (1-b2) g'^2 ;
//   p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// Hyper-parameters come from a small DEVICE array so the launch can sit inside a CUDA graph
// while lr / bias corrections change every step:
//   hp = { lr, beta1, beta2, eps, weight_decay, 1-beta1^t, sqrt(1-beta2^t), gscale, 1-beta1, 1-beta2 }
// (1-beta computed by the host in double and rounded once, as torch.optim.Adam does: 1.0f-0.999f != fl32(0.001))
struct AdamArgs { float *p, *m, *v; const float* g; const float* hp; long long n; };

MMX_D void adam_body(Exec& ex, const AdamArgs& a) {
    const int nthr = ex.nthr;
    ex.phase([&](int tid) {
        const float lr = a.hp[0], b1 = a.hp[1], b2 = a.hp[2], eps = a.hp[3], wd = a.hp[4];
        const float bc1 = a.hp[5], bc2s = a.hp[6], gs = a.hp[7], omb1 = a.hp[8], omb2 = a.hp[9];
        const float step = lr / bc1;
        const long long n4 = a.n >> 2;
        for (long long i = (long long)ex.bid * nthr + tid; i < n4; i += (long long)ex.nblk * nthr) {
            f4 p = ld4(a.p + 4 * i), g = ld4(a.g + 4 * i), m = ld4(a.m + 4 * i), v = ld4(a.v + 4 * i);
            float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g.x, g.y, g.z, g.w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
            MMX_UNROLL
            for (int k = 0; k < 4; ++k) {
                const float gr = fmaf(wd, pp[k], gs * gg[k]);
                mm[k] = mm[k] + (gr - mm[k]) * omb1;
                vv[k] = fmaf(vv[k], b2, omb2 * gr * gr);
                pp[k] -= step * (mm[k] / (sqrtf(vv[k]) / bc2s + eps));
            }
            st4(a.p + 4 * i, make_f4(pp[0], pp[1], pp[2], pp[3]));
            st4(a.m + 4 * i, make_f4(mm[0], mm[1], mm[2], mm[3]));
            st4(a.v + 4 * i, make_f4(vv[0], vv[1], vv[2], vv[3]));
        }
        if (ex.bid == 0)
            for (long long i = 4 * n4 + tid; i < a.n; i += nthr) {
                const float gr = fmaf(wd, a.p[i], gs * a.g[i]);
                const float m = a.m[i] + (gr - a.m[i]) * omb1;
                const float v = fmaf(a.v[i], b2, omb2 * gr * gr);
                a.m[i] = m; a.v[i] = v;
                a.p[i] -= step * (m / (sqrtf(v) / bc2s + eps));
            }
    });
}

// Advance the device-resident optimiser clock (one thread): t = ++*step; refresh the bias
// corrections hp[5] = 1-beta1^t, hp[6] = sqrt(1-beta2^t).  Lets lr-independent per-step state
// live on the device so the whole step can be replayed from a CUDA graph without host writes.
struct AdamAdvanceArgs { float* hp; unsigned int* step; };

MMX_D void adam_advance_body(Exec& ex, const AdamAdvanceArgs& a) {
    ex.phase([&](int tid) {
        if (ex.bid == 0 && tid == 0) {
            const unsigned int t = *a.step + 1u;
            *a.step = t;
            const double b1 = (double)a.hp[1], b2 = (double)a.hp[2];
            a.hp[5] = (float)(1.0 - pow(b1, (double)t));
            a.hp[6] = (float)sqrt(1.0 - pow(b2, (double)t));
        }
    });
}

}  // namespace mmx
