// Launch layer shared by every translation unit of libmmx.so.
//
// nvcc build  -> motionmixerconv_b200/libmmx.so            (the product; CUDA only)
// g++ -DMMX_HOST_EMU -x c++ -> tests/emu/libmmx_emu.so     (test infrastructure: runs the very same
//                                                           kernel bodies phase by phase on the CPU)
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mmx.h"
#include "mmx_common.cuh"

// defined in mmx_api_misc.cu: records the thread-local message returned by mmx_last_error()
int mmx_fail(int code, const char* fmt, ...);
#define fail mmx_fail

namespace mmx {

static inline int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

#if defined(MMX_HOST_EMU)
// ----------------------------------------------------------------------------- CPU emulator
struct DevInfo { int sms = 4; int max_smem = 227 * 1024; };
static inline DevInfo dev_info() { return DevInfo(); }

template <class Body, class Args, int OCC = 1>
static int launch(const Args& a, int grid, int block, size_t smem_bytes, void*, int /*min_blocks*/) {
    std::vector<float> buf(smem_bytes / 4 + 8);
    float* sm = buf.data();
    while (((uintptr_t)sm) & 15) ++sm;
    for (int b = 0; b < grid; ++b) {
        for (size_t i = 0; i < smem_bytes / 4; ++i) sm[i] = NAN;   // poison: catches reads of unwritten smem
        Exec ex{block, b, grid, sm};
        Body::run(ex, a);
    }
    return MMX_OK;
}
#else
// ----------------------------------------------------------------------------- CUDA
struct DevInfo { int sms; int max_smem; };
static inline DevInfo dev_info() {
    static DevInfo cache[64];
    static bool have[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!have[dev]) {
        cudaDeviceGetAttribute(&cache[dev].sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&cache[dev].max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        have[dev] = true;
    }
    return cache[dev];
}

template <class Body, class Args>
static __global__ void __launch_bounds__(256) mmx_kernel(const Args a) {
    extern __shared__ float4 mmx_smem_raw[];
    Exec ex{(int)blockDim.x, (int)blockIdx.x, (int)gridDim.x, reinterpret_cast<float*>(mmx_smem_raw)};
    Body::run(ex, a);
}

// same kernel compiled for two resident CTAs per SM (<= 128 registers per thread)
template <class Body, class Args>
static __global__ void __launch_bounds__(256, 2) mmx_kernel_occ2(const Args a) {
    extern __shared__ float4 mmx_smem_raw[];
    Exec ex{(int)blockDim.x, (int)blockIdx.x, (int)gridDim.x, reinterpret_cast<float*>(mmx_smem_raw)};
    Body::run(ex, a);
}

template <class Body, class Args, int OCC = 1>
static int launch(const Args& a, int grid, int block, size_t smem_bytes, void* stream, int /*min_blocks*/) {
    constexpr int min_blocks = OCC;
    auto kern = OCC >= 2 ? mmx_kernel_occ2<Body, Args> : mmx_kernel<Body, Args>;
    static size_t configured[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    size_t& conf = configured[min_blocks >= 2 ? 1 : 0][dev];
    if (smem_bytes > 48 * 1024 && conf < smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return fail(MMX_E_CUDA, "cudaFuncSetAttribute(%zu): %s", smem_bytes, cudaGetErrorString(e));
        // the planners size tiles so that two CTAs fit one SM: ask for the full shared-memory carve-out, otherwise the driver
        // may pick an L1/shared split that holds only one
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        conf = smem_bytes;
    }
    kern<<<grid, block, smem_bytes, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MMX_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    return MMX_OK;
}
#endif

// grid for a persistent loop over `ntiles` tiles with at most `max_ctas` resident CTAs: balance the
// tiles so every CTA runs the same number of iterations (no ragged last wave)
static inline int balanced_grid(int ntiles, int max_ctas) {
    if (ntiles <= max_ctas) return ntiles > 0 ? ntiles : 1;
    const int waves = (ntiles + max_ctas - 1) / max_ctas;
    return (ntiles + waves - 1) / waves;
}

// Tile size in [lo, hi] (multiples of `step`) for a loop over n rows / sequences: CTAs are dealt round-robin to the SMs, so
// the makespan is (tiles on the busiest SM) x (tile size) -- e.g. 4096 sequences in tiles of 9 give 456 tiles = 4 rounds of
// 148 SMs at 9 sequences, tiles of 7 give 586 tiles = 4 rounds at 7.  Ties go to the larger tile (better weight reuse).
static inline int balanced_tile(int n, int hi, int lo, int step, int sms) {
    if (env_int("MMX_NO_BALANCED_TILE", 0)) return hi;
    int best = hi;
    long long best_cost = -1;
    for (int S = hi; S >= imax(lo, 1); S -= step) {
        const int tiles = (n + S - 1) / S;
        const long long cost = (long long)((tiles + sms - 1) / sms) * S;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = S; }
    }
    return best;
}

static inline Dropout make_dropout(const MmxDropout& s, int training) {
    Dropout d;
    d.seed_lo = (uint32_t)(s.seed & 0xffffffffull);
    d.seed_hi = (uint32_t)(s.seed >> 32);
    d.step = s.step;
    d.step_ptr = s.step_dev;
    if (training && s.p > 0.0f) {
        double t = (double)s.p * 4294967296.0;
        d.thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)(t + 0.5);
        if (d.thresh == 0u) d.thresh = 1u;
        d.scale = 1.0f / (1.0f - s.p);
    } else {
        d.thresh = 0u;
        d.scale = 1.0f;
    }
    return d;
}




constexpr int kThreads = 256;

// zero a device buffer on the stream (memset node: legal inside CUDA-graph capture)
static inline int zero_async(void* p, size_t bytes, void* stream) {
#if defined(MMX_HOST_EMU)
    (void)stream;
    memset(p, 0, bytes);
    return MMX_OK;
#else
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(MMX_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    return MMX_OK;
#endif
}

}  // namespace mmx
