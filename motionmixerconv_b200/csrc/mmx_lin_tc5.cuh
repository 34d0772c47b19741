// Row-wise linear layer on the Blackwell tensor cores (tcgen05 + TMEM + bulk-copy engine), sm_100a:
//
//     y[r, :] = W x[r, :] + b          W: [N, K] (nn.Linear layout), rows r = 0 .. R-1
//
// Serves MlpMixer.conv (Conv2d(1,H,(1,D)) == per-frame Linear D -> H, h36m/mlp_mixer.py:268,325-327) and fc_out (Linear H -> D,
// mlp_mixer.py:300,335) in the reduced-precision mode, with the machinery of mmx_chan_tc5.cuh: 128-row tiles, one thread per
// (row, chunk set), bf16 hi+lo split operands in the 16-bit panel layout (one copy, both orientations), three MMAs per
// product, accumulators in TMEM.  Backward: dW = dY^T X (+ db from a ones column of X) accumulates in TMEM across the CTA's
// persistent loop and is flushed once; dx = dY W (optional) uses the untransposed weight as an MN-major operand.
#pragma once
#include "mmx_chan_tc5.cuh"

namespace mmx {
namespace lin {

using namespace tc5;
using chan::kHalves;
using chan::kThreadsChan;

struct LinArgs {
    const float* x;      // [R, K]
    const float* dy;     // backward: [R, N]
    float* out;          // forward: y [R, N]; backward: dx [R, K] (nullable)
    const float *w, *b;  // [N, K], [N]
    float *g_w, *g_b;    // backward, accumulated
    long long R;
    int K, N;
    int* abort_count;
};

// KP: padded width of BOTH operand buffers (multiple of 16, > K so that column K can carry the ones of the bias gradient, >= N)
template <int KP>
struct LPlan {
    static constexpr uint32_t PS = 128 * 16, PLANE = (KP / 8) * PS, BUF = 2 * PLANE;
    static constexpr uint32_t WPS = KP * 16, WPLANE = (KP / 8) * WPS, WBUF = 2 * WPLANE;
};
template <int KP>
MMX_HD size_t lin_smem_bytes(int K, int N, bool bwd) {
    const size_t sx = ((size_t)128 * K * 4 + 64 + 127) / 128 * 128, sy = ((size_t)128 * N * 4 + 64 + 127) / 128 * 128;
    size_t bx = LPlan<KP>::BUF;
    if (bx < sy) bx = sy;
    if (bx < sx) bx = sx;
    return 1024 + (bwd ? 2 : 1) * bx + LPlan<KP>::WBUF + sx + (bwd ? sy : 0) + (KP + 64) * 4;
}

// ------------------------------------------------------------------------------------------ forward
template <int KP, int VEC>
__global__ void __launch_bounds__(kThreadsChan) lin_fwd_kernel(const LinArgs a) {
    using P = LPlan<KP>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int K = a.K, N = a.N;
    const uint32_t sx = ((uint32_t)128 * K * 4 + 64 + 127) / 128 * 128, sy = ((uint32_t)128 * N * 4 + 64 + 127) / 128 * 128;
    uint32_t bx = P::BUF;
    if (bx < sy) bx = sy;
    if (bx < sx) bx = sx;
    uint8_t* bufX = sm;                                       // operand X, later the output staging tile
    uint8_t* wb = bufX + bx;
    float* S = reinterpret_cast<float*>(wb + P::WBUF);
    float* bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(S) + sx);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias + KP);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* abortf = reinterpret_cast<volatile int*>(tslot + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qtr = warp & 3, half = warp >> 2;
    const int prow = qtr * 32 + lane;
    constexpr int TM_COLS = KP <= 64 ? 64 : 128;
    constexpr int NCH = KP / 8;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); *abortf = 0; fence_mbar_init(); }
    if (warp == 0) tmem_alloc<TM_COLS>(tslot);
    pdl_launch_dependents();
    for (int i = tid; i < (int)(P::WBUF / 16); i += kThreadsChan) reinterpret_cast<uint4*>(wb)[i] = make_uint4(0, 0, 0, 0);
    for (int c = tid; c < KP; c += kThreadsChan) bias[c] = c < N ? a.b[c] : 0.0f;
    __syncthreads();
    chan::stage_weight<KP>(wb, a.w, N, K, nullptr, tid);
    pdl_wait();
    const long long ntiles = (a.R + 127) / 128;
    auto tile_rows = [&](long long t) { return (int)min((long long)128, a.R - t * 128); };
    if (warp == 0 && (long long)blockIdx.x < ntiles)
        chan::stage_in<VEC>(S, a.x, (size_t)blockIdx.x * 128, tile_rows(blockIdx.x), K, K, &bars[0], lane);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t xB = smem_u32(bufX), wB = smem_u32(wb);
    uint32_t ph_in = 0, ph_mma = 0;
    const int nchK = (K + 7) >> 3, nchN = (N + 7) >> 3;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nrows = tile_rows(tile);
        const bool valid = prow < nrows;
        const float* srow = S + (size_t)prow * K;
        mbar_wait(&bars[0], ph_in, abortf);
        ph_in ^= 1;
        if (warp == 0) bulk_wait_read0();              // previous output (staged in the X region) has left shared memory
        __syncthreads();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float v[8];
            chan::ld8<VEC>(srow, 8 * c8, (valid && c8 < nchK) ? K : 0, v);
            chan::put_chunk(bufX, P::PLANE, P::PS, prow, c8, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            const long long next = tile + gridDim.x;
            if (next < ntiles) chan::stage_in<VEC>(S, a.x, (size_t)next * 128, tile_rows(next), K, K, &bars[0], lane);
        }
        if (tid == 0) {
            tc_fence_after();
            chan::gemm3<0, 0>(tmem, xB, xB + P::PLANE, P::PS, wB, wB + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
        {
            float* orow = reinterpret_cast<float*>(bufX) + (size_t)prow * N;
#pragma unroll 1
            for (int c8 = half; c8 < nchN; c8 += kHalves) {
                float u[8], b[8];
                tmem_ld8(tmem_addr(tmem, qtr, 8 * c8), u);
                chan::ld8s(bias + 8 * c8, b);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] += b[j];
                if (valid) chan::st8<2>(orow, 8 * c8, N, u);
            }
        }
        tc_fence_before();
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) bulk_s2g(a.out + (size_t)tile * 128 * N, bufX, (uint32_t)nrows * N * 4u);
            bulk_commit();
        }
    }
    if (warp == 0) bulk_wait_all0();
    if (tid == 0 && *abortf) atomicAdd(a.abort_count, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------ backward
template <int KP, int VEC>
__global__ void __launch_bounds__(kThreadsChan) lin_bwd_kernel(const LinArgs a) {
    using P = LPlan<KP>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int K = a.K, N = a.N;
    const bool want_dx = a.out != nullptr;
    const uint32_t sx = ((uint32_t)128 * K * 4 + 64 + 127) / 128 * 128, sy = ((uint32_t)128 * N * 4 + 64 + 127) / 128 * 128;
    uint32_t bx = P::BUF;
    if (bx < sy) bx = sy;
    if (bx < sx) bx = sx;
    uint8_t* bufX = sm;                                       // operand X (+ ones column at K), later the dx staging tile
    uint8_t* bufY = bufX + bx;                                // operand dY
    uint8_t* wb = bufY + bx;
    float* SX = reinterpret_cast<float*>(wb + P::WBUF);
    float* SY = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(SX) + sx);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(SY) + sy);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* abortf = reinterpret_cast<volatile int*>(tslot + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qtr = warp & 3, half = warp >> 2;
    const int prow = qtr * 32 + lane;
    constexpr int TM_COLS = 2 * KP <= 128 ? 128 : 256;
    constexpr int NCH = KP / 8;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); *abortf = 0; fence_mbar_init(); }
    if (warp == 0) tmem_alloc<TM_COLS>(tslot);
    pdl_launch_dependents();
    for (int i = tid; i < (int)(P::WBUF / 16); i += kThreadsChan) reinterpret_cast<uint4*>(wb)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (want_dx) chan::stage_weight<KP>(wb, a.w, N, K, nullptr, tid);
    pdl_wait();
    const long long ntiles = (a.R + 127) / 128;
    auto tile_rows = [&](long long t) { return (int)min((long long)128, a.R - t * 128); };
    auto load_tile = [&](long long t) {
        const int nr = tile_rows(t);
        if (lane == 0) mbar_expect_tx(&bars[0], (uint32_t)nr * (K + N) * 4u);
        __syncwarp();
        if (lane == 0) {
            bulk_g2s(SX, a.x + (size_t)t * 128 * K, (uint32_t)nr * K * 4u, &bars[0]);
            bulk_g2s(SY, a.dy + (size_t)t * 128 * N, (uint32_t)nr * N * 4u, &bars[0]);
        }
    };
    if (warp == 0 && (long long)blockIdx.x < ntiles) load_tile(blockIdx.x);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tDW = tmem, tDX = tmem + KP;
    const uint32_t xB = smem_u32(bufX), yB = smem_u32(bufY), wB = smem_u32(wb);
    uint32_t ph_in = 0, ph_mma = 0;
    bool first = true;
    const int nchK = (K + 7) >> 3, nchN = (N + 7) >> 3;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nrows = tile_rows(tile);
        const bool valid = prow < nrows;
        mbar_wait(&bars[0], ph_in, abortf);
        ph_in ^= 1;
        if (warp == 0) bulk_wait_read0();
        __syncthreads();
        const float* xrow = SX + (size_t)prow * K;
        const float* yrow = SY + (size_t)prow * N;
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float v[8];
            chan::ld8<VEC>(xrow, 8 * c8, (valid && c8 < nchK) ? K : 0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (valid && 8 * c8 + j == K) v[j] = 1.0f;          // ones column: dW[:, K] = db
            chan::put_chunk(bufX, P::PLANE, P::PS, prow, c8, v);
            chan::ld8<2>(yrow, 8 * c8, (valid && c8 < nchN) ? N : 0, v);
            chan::put_chunk(bufY, P::PLANE, P::PS, prow, c8, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            const long long next = tile + gridDim.x;
            if (next < ntiles) load_tile(next);
        }
        if (tid == 0) {
            tc_fence_after();
            // dW[n][k] += sum_r dY[r][n] X[r][k]   (both MN-major, K = the 128 rows)
            chan::gemm3<1, 1>(tDW, yB, yB + P::PLANE, P::PS, xB, xB + P::PLANE, P::PS, KP, 128 / 16, !first);
            // dx[r][k] = sum_n dY[r][n] W[n][k]    (A = dY K-major, B = W [N rows][K cols] read MN-major)
            if (want_dx) chan::gemm3<0, 1>(tDX, yB, yB + P::PLANE, P::PS, wB, wB + P::WPLANE, P::WPS, KP, KP / 16, false);
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph_mma, abortf);
        ph_mma ^= 1;
        tc_fence_after();
        if (want_dx) {
            float* orow = reinterpret_cast<float*>(bufX) + (size_t)prow * K;
#pragma unroll 1
            for (int c8 = half; c8 < nchK; c8 += kHalves) {
                float u[8];
                tmem_ld8(tmem_addr(tDX, qtr, 8 * c8), u);
                tmem_wait_ld();
                if (valid) chan::st8<VEC>(orow, 8 * c8, K, u);
            }
            tc_fence_before();
            fence_async_smem();
            __syncthreads();
            if (warp == 0) {
                if (lane == 0) bulk_s2g(a.out + (size_t)tile * 128 * K, bufX, (uint32_t)nrows * K * 4u);
                bulk_commit();
            }
        }
    }
    if (warp == 0) bulk_wait_all0();
    __syncthreads();
    if (!first) {
        float* stg = reinterpret_cast<float*>(bufX);          // [KP][KP+1]
        constexpr int SP = KP + 1;
        tc_fence_after();
#pragma unroll 1
        for (int c8 = half; c8 < NCH; c8 += kHalves) {
            float u[8];
            tmem_ld8(tmem_addr(tDW, qtr, 8 * c8), u);
            tmem_wait_ld();
            if (prow < KP)
#pragma unroll
                for (int j = 0; j < 8; ++j) stg[prow * SP + 8 * c8 + j] = u[j];
        }
        __syncthreads();
        for (int i = tid; i < N * K; i += kThreadsChan) {
            const int n = i / K, k = i - n * K;
            red_add(a.g_w + i, stg[n * SP + k]);
        }
        for (int n = tid; n < N; n += kThreadsChan) red_add(a.g_b + n, stg[n * SP + K]);
    }
    if (tid == 0 && *abortf) atomicAdd(a.abort_count, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(tmem);
}

}  // namespace lin
}  // namespace mmx
