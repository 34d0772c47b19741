// Included by mmx_api_mlp_bwd_{gelu,mish}.cu with MMX_BWD_ACT / MMX_BWD_NAME defined: one translation
// unit per activation keeps the (large) backward kernels compiling in parallel.
#include "mmx_mlp_host.cuh"

using namespace mmx;

namespace MMX_BWD_NS {
template <int ACT, int TC, int TOKC, int WT>
struct MlpBwdBody { static MMX_D void run(Exec& ex, const MlpBlockBwdArgs& a) { mlp_block_bwd_body<ACT, TC, TOKC, WT>(ex, a); } };

template <int ACT, int TC, int TOKC>
struct MlpBwdWarpBody { static MMX_D void run(Exec& ex, const MlpBlockBwdArgs& a) { mlp_block_bwd_warp_body<ACT, TC, TOKC>(ex, a); } };

template <int ACT, int WT>
int dispatch_mlp_bwd(const MlpBlockBwdArgs& a, int grid, size_t smem, void* stream) {
    if (a.d.T == 10 && a.d.tok == 20) return launch<MlpBwdBody<ACT, 10, 20, WT>>(a, grid, kThreads, smem, stream, 1);
    return launch<MlpBwdBody<ACT, 0, 0, WT>>(a, grid, kThreads, smem, stream, 1);
}
}  // namespace MMX_BWD_NS
using namespace MMX_BWD_NS;

int MMX_BWD_NAME(const mmx::MlpBlockBwdArgs& a, int wt1, int grid, size_t smem, void* stream) {
    if (wt1 < 0)   // warp-per-sequence-pair variant (mmx_mlp_warp.cuh)
        return launch<MlpBwdWarpBody<MMX_BWD_ACT, 10, 20>>(a, grid, -wt1 * 32, smem, stream, 1);
    return wt1 ? dispatch_mlp_bwd<MMX_BWD_ACT, 1>(a, grid, smem, stream) : dispatch_mlp_bwd<MMX_BWD_ACT, 4>(a, grid, smem, stream);
}
