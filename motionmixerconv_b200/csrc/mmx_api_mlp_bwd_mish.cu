#define MMX_BWD_ACT mmx::ACT_MISH
#define MMX_BWD_NAME mmx_mlp_bwd_launch_mish
#define MMX_BWD_NS mmx_tu_bwd_mish
#include "mmx_api_mlp_bwd.inl"
