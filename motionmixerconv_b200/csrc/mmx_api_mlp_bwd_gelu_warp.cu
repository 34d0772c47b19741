// mmx_mlp_block_bwd, gelu activation, kernel family "warp" (see mmx_api_mlp_bwd.inl).
#define MMX_BWD_ACT mmx::ACT_GELU
#define MMX_BWD_NAME mmx_mlp_bwd_launch_gelu_warp
#define MMX_BWD_NS mmx_tu_bwd_gelu_warp
#define MMX_BWD_PART 2
#include "mmx_api_mlp_bwd.inl"
