// Fused ConvMixer kernels (encoder, ConvMixerBlock forward/backward, head).
#pragma once
#include "mmx_common.cuh"
#include "mmx_mlp.cuh"

namespace mmx {
}  // namespace mmx
