#define MMX_CONV_CP 8
#include "mmx_api_conv_bwd.inl"
