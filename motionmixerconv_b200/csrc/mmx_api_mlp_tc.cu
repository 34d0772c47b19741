// TF32 tensor-core MixerBlock kernels (mmx_mlp_tc.cuh): launch layer.  Reached from mmx_mlp_block_fwd / mmx_mlp_block_bwd
// when MmxMlpBlockDesc.precision == MMX_PREC_TF32 and the shape is one this variant serves.
#include "mmx_mlp_host.cuh"
#include "mmx_mlp_tc.cuh"

using namespace mmx;

#if defined(MMX_HOST_EMU)
// the CPU emulator (test infrastructure) cannot execute mma.sync: the tensor-core variant does not exist there
bool mmx_mlp_tc_ok(const MmxMlpBlockDesc*) { return false; }
int mmx_mlp_tc_fwd(const MmxMlpBlockDesc*, const MmxMlpBlockParams*, const float*, float*, void*) { return fail(MMX_E_UNSUPPORTED, "no tensor cores in the emulator"); }
int mmx_mlp_tc_bwd(const MmxMlpBlockDesc*, const MmxMlpBlockParams*, const MmxMlpBlockParams*, const float*, const float*, float*, void*) {
    return fail(MMX_E_UNSUPPORTED, "no tensor cores in the emulator");
}
#else
bool mmx_mlp_tc_ok(const MmxMlpBlockDesc* d) {
    const bool want = d->precision == MMX_PREC_TF32 || env_int("MMX_MLP_TC_FP32", 0);   // opt-in: FP32 mode on the 3xTF32 build of the same kernels (1e-5 parity, measured no faster than the SIMT kernels)
    return want && d->T == tc::kT && d->tok == tc::kTok && d->H >= 8 && d->H <= 50 && d->ch >= 8 && d->ch <= 50 &&
           (d->H & 1) == 0 && !d->use_max_pooling && (!d->use_se || (d->se_hidden >= 1 && d->se_hidden <= tc::kMaxRR)) &&
           !env_int("MMX_MLP_NO_TC", 0);
}

static int tc_dims(const MmxMlpBlockDesc* d, MlpDims* m) {
    if (d->B <= 0) return fail(MMX_E_INVALID, "non-positive dimension");
    if (d->act != MMX_ACT_GELU && d->act != MMX_ACT_MISH) return fail(MMX_E_INVALID, "Unknown activation function type: %d", d->act);
    m->B = d->B; m->T = d->T; m->H = d->H; m->tok = d->tok; m->ch = d->ch; m->rr = d->use_se ? d->se_hidden : 0;
    m->use_se = d->use_se; m->use_max = 0; m->training = d->training; m->site_base = d->block_index * 4;
    m->S = tc::kSeq; m->w_in_smem = 1; m->align_mask = env_int("MMX_TC_ALIGN_MASK", 2) | (env_int("MMX_TC_ALIGN_MASK_FWD", 8) << 16);   // bwd: bits 0-15, fwd: bits 16+ (measured: one re-alignment per iteration is best)
    return MMX_OK;
}

// One CTA per SM.  The unit of work is a warp's group of 3 sequences; with `waves` groups per warp the job needs
// ceil(groups / waves) warps -- spread over ALL SMs (fewer warps per CTA) rather than packed into fewer CTAs: a warp runs
// faster with fewer co-resident warps, and the makespan is waves x the per-group latency either way.
static void tc_shape(int B, int max_warps_per_cta, int* grid, int* nwarp) {
    const DevInfo di = dev_info();
    const int groups = (B + tc::kSeq - 1) / tc::kSeq;
    const int max_warps = di.sms * max_warps_per_cta;
    const int waves = (groups + max_warps - 1) / max_warps;
    const int warps_needed = (groups + waves - 1) / waves;
    int nw = imax(1, imin(max_warps_per_cta, (warps_needed + di.sms - 1) / di.sms));
    nw = imax(nw, imin(max_warps_per_cta, env_int("MMX_TC_MIN_WARPS", 1)));
    *nwarp = nw;
    *grid = imax(1, imin(di.sms, (warps_needed + nw - 1) / nw));
}

template <class K, class A>
static int tc_launch(K kern, const A& a, int grid, int nwarp, size_t smem, void* stream) {
    // cudaFuncSetAttribute once per (kernel, device): several kernels share this instantiation (same pointer type)
    struct Conf { const void* fn; int dev; size_t smem; };
    static Conf conf[64];
    static int nconf = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    int slot = -1;
    for (int i = 0; i < nconf; ++i)
        if (conf[i].fn == (const void*)kern && conf[i].dev == dev) slot = i;
    if (slot < 0 || conf[slot].smem < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(MMX_E_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (slot < 0 && nconf < 64) slot = nconf++;
        if (slot >= 0) { conf[slot].fn = (const void*)kern; conf[slot].dev = dev; conf[slot].smem = smem; }
    }
    kern<<<grid, nwarp * 32, smem, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MMX_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    return MMX_OK;
}

int mmx_mlp_tc_fwd(const MmxMlpBlockDesc* d, const MmxMlpBlockParams* w, const float* x, float* y, void* stream) {
    MlpBlockFwdArgs a;
    int rc = tc_dims(d, &a.d);
    if (rc) return rc;
    if ((rc = check_block_params(w, d->use_se, "mmx_mlp_block_fwd"))) return rc;
    a.dr = make_dropout(d->dropout, d->training);
    a.w = to_w(w); a.x = x; a.y = y;
    int grid, nwarp;
    tc_shape(d->B, tc::kFwdWarps, &grid, &nwarp);
    const size_t smem = (size_t)tc::smem_layout(false, nwarp).total * 4;
    if (d->precision == MMX_PREC_TF32)
        return d->act == MMX_ACT_GELU ? tc_launch(tc::mlp_block_fwd_tc_kernel<ACT_GELU, 1>, a, grid, nwarp, smem, stream)
                                      : tc_launch(tc::mlp_block_fwd_tc_kernel<ACT_MISH, 1>, a, grid, nwarp, smem, stream);
    return d->act == MMX_ACT_GELU ? tc_launch(tc::mlp_block_fwd_tc_kernel<ACT_GELU, 3>, a, grid, nwarp, smem, stream)
                                  : tc_launch(tc::mlp_block_fwd_tc_kernel<ACT_MISH, 3>, a, grid, nwarp, smem, stream);
}

int mmx_mlp_tc_bwd(const MmxMlpBlockDesc* d, const MmxMlpBlockParams* w, const MmxMlpBlockParams* grads,
                   const float* x, const float* dy, float* dx, void* stream) {
    MlpBlockBwdArgs a;
    int rc = tc_dims(d, &a.d);
    if (rc) return rc;
    if ((rc = check_block_params(w, d->use_se, "mmx_mlp_block_bwd"))) return rc;
    if ((rc = check_block_params(grads, d->use_se, "mmx_mlp_block_bwd(grads)"))) return rc;
    a.dr = make_dropout(d->dropout, d->training);
    a.w = to_w(w); a.g = to_w(grads); a.x = x; a.dy = dy; a.dx = dx;
    int grid, nwarp;
    tc_shape(d->B, tc::kBwdWarps, &grid, &nwarp);
    const size_t smem = (size_t)tc::smem_layout(true, nwarp).total * 4;
    if (smem > (size_t)dev_info().max_smem) return fail(MMX_E_UNSUPPORTED, "MixerBlock (tensor-core variant) does not fit shared memory");
    if (d->precision == MMX_PREC_TF32)
        return d->act == MMX_ACT_GELU ? tc_launch(tc::mlp_block_bwd_tc_kernel<ACT_GELU, 1>, a, grid, nwarp, smem, stream)
                                      : tc_launch(tc::mlp_block_bwd_tc_kernel<ACT_MISH, 1>, a, grid, nwarp, smem, stream);
    return d->act == MMX_ACT_GELU ? tc_launch(tc::mlp_block_bwd_tc_kernel<ACT_GELU, 3>, a, grid, nwarp, smem, stream)
                                  : tc_launch(tc::mlp_block_bwd_tc_kernel<ACT_MISH, 3>, a, grid, nwarp, smem, stream);
}
#endif
