#define MMX_CONV_CP 2
#include "mmx_api_conv_bwd.inl"
