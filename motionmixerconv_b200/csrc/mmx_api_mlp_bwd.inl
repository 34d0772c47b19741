// Included by mmx_api_mlp_bwd_{gelu,mish}{,_wt4,_warp}.cu with MMX_BWD_ACT / MMX_BWD_NAME / MMX_BWD_NS / MMX_BWD_PART defined:
// one translation unit per (activation, kernel family) keeps the (large) backward kernels compiling in parallel.
//   MMX_BWD_PART 0: generic CTA-per-tile kernels, one 4x4 weight-gradient tile per thread   (WT = 1)
//   MMX_BWD_PART 1: generic CTA-per-tile kernels, four tiles per thread                      (WT = 4)
//   MMX_BWD_PART 2: warp-per-sequence-pair variant (mmx_mlp_warp.cuh)
#include "mmx_mlp_host.cuh"

using namespace mmx;

namespace MMX_BWD_NS {
#if MMX_BWD_PART == 2
template <int ACT, int TC, int TOKC>
struct MlpBwdWarpBody { static MMX_D void run(Exec& ex, const MlpBlockBwdArgs& a) { mlp_block_bwd_warp_body<ACT, TC, TOKC>(ex, a); } };
#else
template <int ACT, int TC, int TOKC, int WT>
struct MlpBwdBody { static MMX_D void run(Exec& ex, const MlpBlockBwdArgs& a) { mlp_block_bwd_body<ACT, TC, TOKC, WT>(ex, a); } };

template <int ACT, int WT>
int dispatch_mlp_bwd(const MlpBlockBwdArgs& a, int grid, size_t smem, void* stream) {
    if (a.d.T == 10 && a.d.tok == 20) return launch<MlpBwdBody<ACT, 10, 20, WT>>(a, grid, kThreads, smem, stream, 1);
    return launch<MlpBwdBody<ACT, 0, 0, WT>>(a, grid, kThreads, smem, stream, 1);
}
#endif
}  // namespace MMX_BWD_NS
using namespace MMX_BWD_NS;

// wt1 < 0: warp variant with -wt1 warps per CTA; otherwise the generic kernels
int MMX_BWD_NAME(const mmx::MlpBlockBwdArgs& a, int wt1, int grid, size_t smem, void* stream) {
#if MMX_BWD_PART == 2
    return launch<MlpBwdWarpBody<MMX_BWD_ACT, 10, 20>>(a, grid, -wt1 * 32, smem, stream, 1);
#elif MMX_BWD_PART == 1
    (void)wt1;
    return dispatch_mlp_bwd<MMX_BWD_ACT, 4>(a, grid, smem, stream);
#else
    (void)wt1;
    return dispatch_mlp_bwd<MMX_BWD_ACT, 1>(a, grid, smem, stream);
#endif
}
