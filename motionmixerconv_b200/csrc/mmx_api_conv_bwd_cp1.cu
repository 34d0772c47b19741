// mmx_conv_half_bwd: fused ConvMixerBlock-half backward (include/mmx.h).
#define MMX_CONV_CP 1
#include "mmx_api_conv_bwd.inl"

int mmx_conv_bwd_launch_cp2(const mmx::ConvHalfBwdArgs& a, int act, int grid, size_t smem, void* stream);
int mmx_conv_bwd_launch_cp4(const mmx::ConvHalfBwdArgs& a, int act, int grid, size_t smem, void* stream);
int mmx_conv_bwd_launch_cp8(const mmx::ConvHalfBwdArgs& a, int act, int grid, size_t smem, void* stream);

extern "C" int mmx_conv_half_bwd(const MmxConvHalfDesc* d, const MmxConvHalfParams* w, const MmxConvHalfParams* grads,
                                 const float* x, const float* dy, float* dx, void* stream) {
    if (!x || !dy || !dx) return fail(MMX_E_INVALID, "mmx_conv_half_bwd: null tensor");
    ConvHalfBwdArgs a;
    size_t smem; int grid;
    int rc = plan_conv_half(d, true, &a.d, &smem, &grid);
    if (rc) return rc;
    if ((rc = check_conv_params(w, d->use_se, "mmx_conv_half_bwd"))) return rc;
    if ((rc = check_conv_params(grads, d->use_se, "mmx_conv_half_bwd(grads)"))) return rc;
    a.dr = make_dropout(d->dropout, d->training);
    a.w = to_cw(w); a.g = to_cw(grads); a.x = x; a.dy = dy; a.dx = dx;
    a.aff = w->bn_aff; a.z = nullptr; a.gd = nullptr; a.bn = nullptr; a.coef = nullptr;
    switch (conv_cp(d->C)) {
        case 1: return mmx_conv_bwd_launch_cp1(a, d->act, grid, smem, stream);
        case 2: return mmx_conv_bwd_launch_cp2(a, d->act, grid, smem, stream);
        case 4: return mmx_conv_bwd_launch_cp4(a, d->act, grid, smem, stream);
        default: return mmx_conv_bwd_launch_cp8(a, d->act, grid, smem, stream);
    }
}

extern "C" int mmx_conv_half_bn_bwd2(const MmxConvHalfDesc* d, const MmxConvHalfParams* w, const MmxConvHalfParams* grads, const float* bn,
                                     const float* coef, const float* x, const float* z, const float* dy, const float* gd, float* dx,
                                     void* stream) {
    if (!bn || !coef || !x || !z || !dy || !dx) return fail(MMX_E_INVALID, "mmx_conv_half_bn_bwd2: null tensor");
    if (d && d->use_se && !gd) return fail(MMX_E_INVALID, "mmx_conv_half_bn_bwd2: use_se set but the (gate, ds) workspace is null");
    ConvHalfBwdArgs a;
    size_t smem; int grid;
    int rc = plan_conv_half(d, true, &a.d, &smem, &grid);
    if (rc) return rc;
    if (d->use_se && d->use_max_pooling) return fail(MMX_E_UNSUPPORTED, "BatchNorm halves support the mean squeeze only");
    if ((rc = check_conv_params(w, 0, "mmx_conv_half_bn_bwd2"))) return rc;
    if ((rc = check_conv_params(grads, 0, "mmx_conv_half_bn_bwd2(grads)"))) return rc;
    a.d.bn_mode = 2; a.d.training = 0;
    a.dr = make_dropout(d->dropout, 0);
    a.w = to_cw(w); a.g = to_cw(grads); a.x = x; a.dy = dy; a.dx = dx;
    a.aff = nullptr; a.z = z; a.gd = gd; a.bn = bn; a.coef = coef;
    switch (conv_cp(d->C)) {
        case 1: return mmx_conv_bwd_launch_cp1(a, d->act, grid, smem, stream);
        case 2: return mmx_conv_bwd_launch_cp2(a, d->act, grid, smem, stream);
        case 4: return mmx_conv_bwd_launch_cp4(a, d->act, grid, smem, stream);
        default: return mmx_conv_bwd_launch_cp8(a, d->act, grid, smem, stream);
    }
}
