// mmx_mlp_block_bwd: fused MixerBlock backward (include/mmx.h) -- the C entry point, plus the gelu / generic-WT1 kernels.
#define MMX_BWD_ACT mmx::ACT_GELU
#define MMX_BWD_NAME mmx_mlp_bwd_launch_gelu_wt1
#define MMX_BWD_NS mmx_tu_bwd_gelu_wt1
#define MMX_BWD_PART 0
#include "mmx_api_mlp_bwd.inl"

#define MMX_DECL_BWD(name) int name(const mmx::MlpBlockBwdArgs& a, int wt1, int grid, size_t smem, void* stream)
MMX_DECL_BWD(mmx_mlp_bwd_launch_gelu_wt4); MMX_DECL_BWD(mmx_mlp_bwd_launch_gelu_warp);
MMX_DECL_BWD(mmx_mlp_bwd_launch_mish_wt1); MMX_DECL_BWD(mmx_mlp_bwd_launch_mish_wt4); MMX_DECL_BWD(mmx_mlp_bwd_launch_mish_warp);

bool mmx_mlp_tc5_ok(const MmxMlpBlockDesc* d);
int mmx_mlp_tc5_bwd(const MmxMlpBlockDesc* d, const MmxMlpBlockParams* w, const MmxMlpBlockParams* grads,
                    const float* x, const float* dy, float* dx, void* stream);

extern "C" int mmx_mlp_block_bwd(const MmxMlpBlockDesc* d, const MmxMlpBlockParams* w, const MmxMlpBlockParams* grads,
                                 const float* x, const float* dy, float* dx, void* stream) {
    if (!x || !dy || !dx) return fail(MMX_E_INVALID, "mmx_mlp_block_bwd: null tensor");
    MlpBlockBwdArgs a;
    size_t smem; int grid, nwarp = 0;
    if (!d) return fail(MMX_E_INVALID, "null descriptor");
    if (mmx_mlp_tc5_ok(d)) return mmx_mlp_tc5_bwd(d, w, grads, x, dy, dx, stream);   // tcgen05 / TMEM family (sm_100a)
    const bool warp_variant = mlp_warp_variant_ok(d);
    int rc = warp_variant ? plan_mlp_block_warp(d, true, &a.d, &smem, &grid, &nwarp) : plan_mlp_block(d, true, &a.d, &smem, &grid);
    if (rc) return rc;
    if ((rc = check_block_params(w, d->use_se, "mmx_mlp_block_bwd"))) return rc;
    if ((rc = check_block_params(grads, d->use_se, "mmx_mlp_block_bwd(grads)"))) return rc;
    a.dr = make_dropout(d->dropout, d->training);
    a.w = to_w(w); a.g = to_w(grads); a.x = x; a.dy = dy; a.dx = dx;
    const int tiles = imax(((d->ch + 3) / 4) * ((d->H + 3) / 4), 1);
    const int wt1 = warp_variant ? -nwarp : (tiles <= kThreads);
    const bool gelu = d->act == MMX_ACT_GELU;
    if (warp_variant) return gelu ? mmx_mlp_bwd_launch_gelu_warp(a, wt1, grid, smem, stream) : mmx_mlp_bwd_launch_mish_warp(a, wt1, grid, smem, stream);
    if (wt1) return gelu ? mmx_mlp_bwd_launch_gelu_wt1(a, wt1, grid, smem, stream) : mmx_mlp_bwd_launch_mish_wt1(a, wt1, grid, smem, stream);
    return gelu ? mmx_mlp_bwd_launch_gelu_wt4(a, wt1, grid, smem, stream) : mmx_mlp_bwd_launch_mish_wt4(a, wt1, grid, smem, stream);
}
