// mmx_conv_half_fwd, mmx_se_tail_{fwd,bwd} (include/mmx.h).
#include "mmx_conv_host.cuh"
#include "mmx_conv_io.cuh"

using namespace mmx;

namespace mmx_tu_conv_fwd {
template <int ACT, int CP>
struct ConvFwdBody { static MMX_D void run(Exec& ex, const ConvHalfFwdArgs& a) { conv_half_fwd_body<ACT, CP>(ex, a); } };
struct SeTailFwdBody { static MMX_D void run(Exec& ex, const SeTailArgs& a) { se_tail_fwd_body(ex, a); } };
struct SeTailBwdBody { static MMX_D void run(Exec& ex, const SeTailArgs& a) { se_tail_bwd_body(ex, a); } };

template <int ACT>
int dispatch(const ConvHalfFwdArgs& a, int grid, size_t smem, void* stream) {
    switch (conv_cp(a.d.C)) {
        case 1: return launch<ConvFwdBody<ACT, 1>>(a, grid, kThreads, smem, stream, 1);
        case 2: return launch<ConvFwdBody<ACT, 2>>(a, grid, kThreads, smem, stream, 1);
        case 4: return launch<ConvFwdBody<ACT, 4>>(a, grid, kThreads, smem, stream, 1);
        default: return launch<ConvFwdBody<ACT, 8>>(a, grid, kThreads, smem, stream, 1);
    }
}

int plan_se_tail(int B, int C, int T, int E, int se_hidden, int use_se, int use_max, SeTailDims* out, size_t* smem, int* grid) {
    if (B <= 0 || C <= 0 || T <= 0 || E <= 0) return fail(MMX_E_INVALID, "non-positive dimension");
    if (T > 32) return fail(MMX_E_UNSUPPORTED, "in_nTP %d > 32", T);
    if (use_se && se_hidden < 1) return fail(MMX_E_UNSUPPORTED, "in_nTP // r_se == 0: empty SE bottleneck");
    SeTailDims d; d.B = B; d.C = C; d.T = T; d.E = E; d.rr = use_se ? se_hidden : 0; d.use_se = use_se; d.use_max = use_max;
    d.S = imin(B, imax(1, kConvTileElems / (C * T * E)));
    const DevInfo di = dev_info();
    *out = d; *smem = (size_t)se_tail_smem(d).total * 4;
    *grid = balanced_grid((B + d.S - 1) / d.S, di.sms * 4);
    return MMX_OK;
}
}  // namespace mmx_tu_conv_fwd
using namespace mmx_tu_conv_fwd;

extern "C" int mmx_conv_half_fwd(const MmxConvHalfDesc* d, const MmxConvHalfParams* w, const float* x, float* y, void* stream) {
    if (!x || !y) return fail(MMX_E_INVALID, "mmx_conv_half_fwd: null tensor");
    ConvHalfFwdArgs a;
    size_t smem; int grid;
    int rc = plan_conv_half(d, false, &a.d, &smem, &grid);
    if (rc) return rc;
    if ((rc = check_conv_params(w, d->use_se, "mmx_conv_half_fwd"))) return rc;
    a.dr = make_dropout(d->dropout, d->training);
    a.w = to_cw(w); a.x = x; a.y = y;
    return d->act == MMX_ACT_GELU ? dispatch<ACT_GELU>(a, grid, smem, stream) : dispatch<ACT_MISH>(a, grid, smem, stream);
}

extern "C" int mmx_se_tail_fwd(int B, int C, int T, int E, int se_hidden, int use_se, int use_max_pooling,
                               const float* se_w1, const float* se_w2, const float* x, float* y, void* stream) {
    if (!x || !y) return fail(MMX_E_INVALID, "mmx_se_tail_fwd: null tensor");
    if (use_se && (!se_w1 || !se_w2)) return fail(MMX_E_INVALID, "mmx_se_tail_fwd: use_se set but SE weights are null");
    SeTailArgs a; size_t smem; int grid;
    int rc = plan_se_tail(B, C, T, E, se_hidden, use_se, use_max_pooling, &a.d, &smem, &grid);
    if (rc) return rc;
    a.se1 = se_w1; a.se2 = se_w2; a.g_se1 = a.g_se2 = nullptr; a.x = x; a.dy = nullptr; a.out = y;
    return launch<SeTailFwdBody>(a, grid, kThreads, smem, stream, 1);
}

extern "C" int mmx_se_tail_bwd(int B, int C, int T, int E, int se_hidden, int use_se, int use_max_pooling,
                               const float* se_w1, const float* se_w2, float* g_se_w1, float* g_se_w2,
                               const float* x, const float* dy, float* dx, void* stream) {
    if (!x || !dy || !dx) return fail(MMX_E_INVALID, "mmx_se_tail_bwd: null tensor");
    if (use_se && (!se_w1 || !se_w2 || !g_se_w1 || !g_se_w2)) return fail(MMX_E_INVALID, "mmx_se_tail_bwd: use_se set but SE weights / grads are null");
    SeTailArgs a; size_t smem; int grid;
    int rc = plan_se_tail(B, C, T, E, se_hidden, use_se, use_max_pooling, &a.d, &smem, &grid);
    if (rc) return rc;
    a.se1 = se_w1; a.se2 = se_w2; a.g_se1 = g_se_w1; a.g_se2 = g_se_w2; a.x = x; a.dy = dy; a.out = dx;
    return launch<SeTailBwdBody>(a, grid, kThreads, smem, stream, 1);
}
