#define MMX_CONV_CP 4
#include "mmx_api_conv_bwd.inl"
