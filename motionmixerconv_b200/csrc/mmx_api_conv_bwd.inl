// Included by mmx_api_conv_bwd_cp{1,2,4,8}.cu with MMX_CONV_CP defined: one translation unit per padded
// channel count keeps the (large) backward kernels compiling in parallel.
#include "mmx_conv_host.cuh"

using namespace mmx;

#define MMX_CAT2(a, b) a##b
#define MMX_CAT(a, b) MMX_CAT2(a, b)

namespace MMX_CAT(mmx_tu_conv_bwd_cp, MMX_CONV_CP) {
template <int ACT>
struct ConvBwdBody { static MMX_D void run(Exec& ex, const ConvHalfBwdArgs& a) { conv_half_bwd_body<ACT, MMX_CONV_CP>(ex, a); } };
}

int MMX_CAT(mmx_conv_bwd_launch_cp, MMX_CONV_CP)(const mmx::ConvHalfBwdArgs& a, int act, int grid, size_t smem, void* stream) {
    using namespace MMX_CAT(mmx_tu_conv_bwd_cp, MMX_CONV_CP);
    return act == MMX_ACT_GELU ? launch<ConvBwdBody<ACT_GELU>>(a, grid, kThreads, smem, stream, 1)
                               : launch<ConvBwdBody<ACT_MISH>>(a, grid, kThreads, smem, stream, 1);
}
