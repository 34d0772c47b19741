// Common device-side building blocks for the fused mixer kernels (sm_100a).
//
// Every kernel body in this directory is written as a sequence of *phases*:
//
//     ex.phase([&](int tid) { ... });        // implicit CTA barrier after each phase
//
// A phase may read anything written by earlier phases and may write shared / global
// memory, but threads never communicate inside a phase.  On the GPU a phase is
// "run the lambda for threadIdx.x, then __syncthreads()".  The same source also builds
// with g++ (-DMMX_HOST_EMU) into tests/emu/libmmx_emu.so, where a phase is a loop over
// tid: that emulator is TEST INFRASTRUCTURE (it lets the CPU test-suite run the real
// kernel source against the oracle in a container without a GPU); the product library
// libmmx.so contains only the CUDA build and has no CPU path.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(MMX_HOST_EMU)
#include <cstring>
#include <vector>
#define MMX_HD inline
#define MMX_D inline
#define MMX_UNROLL
#define MMX_NOUNROLL
#define MMX_NOINLINE
#else
#include <cuda_runtime.h>
#define MMX_HD __host__ __device__ __forceinline__
#define MMX_D __device__ __forceinline__
#define MMX_UNROLL _Pragma("unroll")
#define MMX_NOUNROLL _Pragma("unroll 1")
#define MMX_NOINLINE __noinline__
#endif

namespace mmx {

// ------------------------------------------------------------------------------------------
// 128-bit shared/global access
// ------------------------------------------------------------------------------------------
#if defined(MMX_HOST_EMU)
struct alignas(16) f4 { float x, y, z, w; };
#else
typedef float4 f4;
#endif

MMX_D f4 ld4(const float* p) { return *reinterpret_cast<const f4*>(p); }
MMX_D void st4(float* p, const f4& v) { *reinterpret_cast<f4*>(p) = v; }
MMX_D f4 make_f4(float x, float y, float z, float w) { f4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// global-memory reduction (fire-and-forget RED on the GPU)
MMX_D void red_add(float* p, float v) {
#if defined(MMX_HOST_EMU)
    *p += v;
#else
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#endif
}

// shared-memory accumulate from several threads of the CTA (ATOMS on the GPU)
MMX_D void smem_add(float* p, float v) {
#if defined(MMX_HOST_EMU)
    *p += v;
#else
    atomicAdd(p, v);
#endif
}

// CTA-scope spin lock in shared memory taken by one WARP: lane 0 acquires, the warp's other lanes wait at the
// __syncwarp.  Protects short plain read-modify-write sections on accumulators shared by the CTA's warps (shared
// memory has no native fp32 atomic add: atomicAdd(float) on shared compiles to a CAS loop per element).
MMX_D void warp_lock(unsigned int* l, int lane) {
#if defined(MMX_HOST_EMU)
    (void)l; (void)lane;
#else
    if (lane == 0) {
        while (atomicCAS(l, 0u, 1u) != 0u) {}
        __threadfence_block();
    }
    __syncwarp();
#endif
}
MMX_D void warp_unlock(unsigned int* l, int lane) {
#if defined(MMX_HOST_EMU)
    (void)l; (void)lane;
#else
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicExch(l, 0u);
    }
#endif
}

// global double-precision accumulate (BatchNorm batch sums)
MMX_D void red_add_f64(double* p, double v) {
#if defined(MMX_HOST_EMU)
    *p += v;
#else
    atomicAdd(p, v);
#endif
}

MMX_HD int round_up(int x, int m) { return (x + m - 1) / m * m; }
MMX_HD int imin(int a, int b) { return a < b ? a : b; }
MMX_HD int imax(int a, int b) { return a > b ? a : b; }

// Row pitch (floats) for a row-major shared tile of logical width w: a multiple of 4 (so rows
// are float4-aligned) whose quad count is odd, so that the same column of 8 consecutive rows
// falls into 8 different 4-bank groups (conflict-free float4 access by one-thread-per-row
// phases and by the strided GEMM tiles below).
MMX_HD int pitch_of(int w) {
    int p = round_up(w, 4);
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}

// ------------------------------------------------------------------------------------------
// activations.  ACT: 0 = exact GELU (nn.GELU(), approximate='none'), 1 = Mish.
// Reference: h36m/mlp_mixer.py:37-41,78-81; h36m/conv_mixer_model.py:121-124.
// ------------------------------------------------------------------------------------------
enum { ACT_GELU = 0, ACT_MISH = 1 };

// exp / divide of the activations: one MUFU each on the GPU (ex2.approx.ftz / rcp.approx.ftz: <= 2 ulp, no denormal
// fix-up instructions; two orders of magnitude inside the 1e-5 parity bar, checked by the GPU parity tests), libm in the
// emulator
#if defined(MMX_HOST_EMU)
MMX_D float fast_exp(float x) { return expf(x); }
MMX_D float fast_div(float a, float b) { return a / b; }
MMX_D float fast_rcp(float b) { return 1.0f / b; }
#else
MMX_D float fast_exp(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f)); return r; }
MMX_D float fast_rcp(float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
MMX_D float fast_div(float a, float b) { return a * fast_rcp(b); }
#endif

template <int ACT>
MMX_D float act_fwd(float u) {
    if (ACT == ACT_GELU) {
        return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f));
    } else {
        // u * tanh(softplus(u)); tanh(log(1+e)) = n/(n+2) with n = e*(e+2).  softplus threshold 20: beyond it e is
        // clamped to exp(20), where n/(n+2) rounds to exactly 1 -> returns u (branch-free)
        float e = fast_exp(fminf(u, 20.0f));
        float n = e * (e + 2.0f);
        return u * (n * fast_rcp(n + 2.0f));
    }
}

// returns act(u) in *a and d act/du as the return value
template <int ACT>
MMX_D float act_fwd_grad(float u, float* a) {
    if (ACT == ACT_GELU) {
        float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
        *a = u * cdf;
        return cdf + u * fast_exp(-0.5f * u * u) * 0.39894228040143268f;
    } else {
        // mish'(u) = e * w / (n+2)^2,  w = 4(u+1) + e*(4u + 6 + e*(4 + e))   (two MUFUs, branch-free; e clamped at exp(20),
        // where the value rounds to 1 like the reference's softplus threshold branch)
        float uc = fminf(u, 20.0f);
        float e = fast_exp(uc);
        float n = e * (e + 2.0f);
        float inv = fast_rcp(n + 2.0f);
        float w = fmaf(e, fmaf(e, 4.0f + e, fmaf(4.0f, uc, 6.0f)), 4.0f * (uc + 1.0f));
        *a = u * (n * inv);
        return (e * inv) * (w * inv);
    }
}

MMX_D float sigmoidf_(float q) { return fast_div(1.0f, 1.0f + fast_exp(-q)); }

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (dropout masks).  Counter = (c0,c1,c2,c3), key = (k0,k1).
// tests/philox_np.py holds the numpy twin used to check masks bit-for-bit.
// ------------------------------------------------------------------------------------------
struct u4 { uint32_t x, y, z, w; };

MMX_D u4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    MMX_UNROLL
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    u4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
    return r;
}

// Dropout descriptor carried by every kernel that has a Dropout site.
struct Dropout {
    uint32_t seed_lo, seed_hi;   // Philox key
    uint32_t step;               // training step counter (c3)
    uint32_t thresh;             // keep iff r >= thresh, thresh = round(p * 2^32) (0 => no dropout)
    float scale;                 // 1/(1-p)
    const uint32_t* step_ptr;    // optional device-resident addend to `step` (CUDA-graph replays)
};

MMX_D Dropout resolve_dropout(Dropout d) {
    if (d.step_ptr) d.step += *d.step_ptr;
    return d;
}

// keep-scale (0 or 1/(1-p)) for element `elem` of dropout site `site`
#if defined(MMX_HOST_EMU)
inline
#else
static __device__ __noinline__
#endif
float dropout_scale(const Dropout& d, uint32_t site, uint64_t elem) {
    u4 r = philox4x32_10((uint32_t)(elem >> 2), (uint32_t)(elem >> 34), site, d.step, d.seed_lo, d.seed_hi);
    uint32_t lane = (uint32_t)elem & 3u;
    uint32_t v = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
    return v >= d.thresh ? d.scale : 0.0f;
}

// Two quads per Philox call (16 random bits per element): `pair` numbers consecutive row pairs of a [rows][W] site,
// ks0 / ks1 are the keep-scales of quad `q` of the pair's first / second row.  Philox4x32-7 (Crush-resistant per
// Salmon et al., SC'11): the masks only need to be uncorrelated, not cryptographic.
MMX_D void dropout_rowpair(const Dropout& d, uint32_t site, uint64_t pair, int q, int W4, float (&ks0)[4], float (&ks1)[4]) {
    const uint64_t ctr = pair * (uint64_t)W4 + (uint64_t)q;
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = site ^ 0x5bd1e995u, c3 = d.step, k0 = d.seed_lo, k1 = d.seed_hi;
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    MMX_UNROLL
    for (int i = 0; i < 7; ++i) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += W0; k1 += W1;
    }
    const uint32_t th = d.thresh >> 16;   // keep iff the 16-bit field >= th
    ks0[0] = (c0 & 0xffffu) >= th ? d.scale : 0.0f; ks0[1] = (c0 >> 16) >= th ? d.scale : 0.0f;
    ks0[2] = (c1 & 0xffffu) >= th ? d.scale : 0.0f; ks0[3] = (c1 >> 16) >= th ? d.scale : 0.0f;
    ks1[0] = (c2 & 0xffffu) >= th ? d.scale : 0.0f; ks1[1] = (c2 >> 16) >= th ? d.scale : 0.0f;
    ks1[2] = (c3 & 0xffffu) >= th ? d.scale : 0.0f; ks1[3] = (c3 >> 16) >= th ? d.scale : 0.0f;
}

// Quad-granular masks: a dropout site is a [rows][W] tensor with W4 = ceil(W/4) quads per row; quad number
// row*W4 + w/4 is one Philox counter and yields the keep-scales of its 4 elements.
#if defined(MMX_HOST_EMU)
inline
#else
static __device__ __noinline__
#endif
void dropout_quad(const Dropout& d, uint32_t site, uint64_t quad, float (&s)[4]) {
    u4 r = philox4x32_10((uint32_t)quad, (uint32_t)(quad >> 32), site, d.step, d.seed_lo, d.seed_hi);
    s[0] = r.x >= d.thresh ? d.scale : 0.0f; s[1] = r.y >= d.thresh ? d.scale : 0.0f;
    s[2] = r.z >= d.thresh ? d.scale : 0.0f; s[3] = r.w >= d.thresh ? d.scale : 0.0f;
}

// ------------------------------------------------------------------------------------------
// executors
// ------------------------------------------------------------------------------------------
#if defined(MMX_HOST_EMU)
// warp-level executor: sub-phases of ONE warp, separated by __syncwarp() on the GPU (lanes may exchange data
// through shared memory between sub-phases); the emulator runs each sub-phase as a loop over the 32 lanes
struct WarpExec {
    int warp, nwarp;
    template <class F>
    void phase(F&& f) {
        for (int l = 0; l < 32; ++l) f(l);
    }
    // CTA-wide re-alignment point: carries no data dependency (the warps of a CTA stay independent), it only keeps them
    // walking the (instruction-cache-sized) loop body together on the GPU.  Every warp of the CTA must reach it the same
    // number of times.  Nothing to do in the emulator, which runs the warps one after the other.
    void align() {}
};
struct Exec {
    int nthr, bid, nblk;
    float* smem;
    template <class F>
    void phase(F&& f) {
        for (int t = 0; t < nthr; ++t) f(t);
    }
    // run body(WarpExec&) once per warp of the CTA; warps do not communicate inside (no CTA barrier implied)
    template <class F>
    void warps(F&& body) {
        for (int w = 0; w < nthr / 32; ++w) { WarpExec wx{w, nthr / 32}; body(wx); }
    }
};
template <class T>
struct PerThread {
    std::vector<T> v;
    explicit PerThread(const Exec& ex) : v(ex.nthr) {}
    T& operator[](int tid) { return v[tid]; }
};
#else
struct WarpExec {
    int warp, nwarp;
    template <class F>
    __device__ __forceinline__ void phase(F&& f) {
        f((int)(threadIdx.x & 31));
        __syncwarp();
    }
    __device__ __forceinline__ void align() { __syncthreads(); }
};
struct Exec {
    int nthr, bid, nblk;
    float* smem;
    template <class F>
    __device__ __forceinline__ void phase(F&& f) {
        f((int)threadIdx.x);
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void warps(F&& body) {
        WarpExec wx{(int)(threadIdx.x >> 5), (int)(blockDim.x >> 5)};
        body(wx);
    }
};
template <class T>
struct PerThread {
    T v;
    __device__ __forceinline__ explicit PerThread(const Exec&) {}
    __device__ __forceinline__ T& operator[](int) { return v; }
};
#endif

// ------------------------------------------------------------------------------------------
// CTA-wide register-tiled GEMMs on shared-memory operands (fp32 SIMT).
//
// All operands are fp32 in shared memory (B may also be a global pointer: the access
// pattern is the same).  K-extent of row-major operands must be zero-padded to a multiple
// of 4.  Thread tiles are TM x TN with *strided* ownership so that neighbouring lanes touch
// neighbouring rows (pitch_of() then makes every float4 access conflict-free):
//     rows  m = rt + i * n_rt  (i < TM),   cols  n = ct + j * n_ct  (j < TN)
// The epilogue functor is called as epi(m, n, value) for every in-range output element.
// ------------------------------------------------------------------------------------------

// C[m][n] = sum_k A[m*lda + k] * B[n*ldb + k]      ("NT": both operands k-contiguous)
template <int TM, int TN, class Epi>
MMX_D void gemm_nt(int tid, int nthr, const float* A, int lda, const float* B, int ldb,
                   int M, int N, int K, Epi&& epi) {
    const int n_rt = (M + TM - 1) / TM, n_ct = (N + TN - 1) / TN;
    const int K4 = (K + 3) >> 2;
    for (int tile = tid; tile < n_rt * n_ct; tile += nthr) {
        const int rt = tile / n_ct, ct = tile - rt * n_ct;
        const float* ap[TM];
        const float* bp[TN];
        MMX_UNROLL
        for (int i = 0; i < TM; ++i) ap[i] = A + (size_t)imin(rt + i * n_rt, M - 1) * lda;
        MMX_UNROLL
        for (int j = 0; j < TN; ++j) bp[j] = B + (size_t)imin(ct + j * n_ct, N - 1) * ldb;
        float acc[TM][TN];
        MMX_UNROLL
        for (int i = 0; i < TM; ++i)
            MMX_UNROLL
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
        for (int k4 = 0; k4 < K4; ++k4) {
            f4 av[TM], bv[TN];
            MMX_UNROLL
            for (int i = 0; i < TM; ++i) av[i] = ld4(ap[i] + 4 * k4);
            MMX_UNROLL
            for (int j = 0; j < TN; ++j) bv[j] = ld4(bp[j] + 4 * k4);
            MMX_UNROLL
            for (int i = 0; i < TM; ++i)
                MMX_UNROLL
                for (int j = 0; j < TN; ++j) {
                    acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                }
        }
        MMX_UNROLL
        for (int i = 0; i < TM; ++i) {
            const int m = rt + i * n_rt;
            if (m < M) {
                MMX_UNROLL
                for (int j = 0; j < TN; ++j) {
                    const int n = ct + j * n_ct;
                    if (n < N) epi(m, n, acc[i][j]);
                }
            }
        }
    }
}

// C[m][n] = sum_k A[m*lda + k] * B[k*ldb + n]      ("NN": A row-major, B k-major)
// cols owned by a thread are 4 consecutive: n = 4*ct + j  (TN == 4); ldb % 4 == 0; columns of B
// beyond N (up to the next multiple of 4) must be readable (zero padded).
template <int TM, class Epi>
MMX_D void gemm_nn(int tid, int nthr, const float* A, int lda, const float* B, int ldb,
                   int M, int N, int K, Epi&& epi) {
    const int n_rt = (M + TM - 1) / TM, n_ct = (N + 3) >> 2;
    const int K4 = K >> 2;
    for (int tile = tid; tile < n_rt * n_ct; tile += nthr) {
        const int rt = tile / n_ct, ct = tile - rt * n_ct;
        const float* ap[TM];
        MMX_UNROLL
        for (int i = 0; i < TM; ++i) ap[i] = A + (size_t)imin(rt + i * n_rt, M - 1) * lda;
        const float* bp = B + 4 * ct;
        float acc[TM][4];
        MMX_UNROLL
        for (int i = 0; i < TM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
        for (int k4 = 0; k4 < K4; ++k4) {
            f4 av[TM];
            MMX_UNROLL
            for (int i = 0; i < TM; ++i) av[i] = ld4(ap[i] + 4 * k4);
            const f4 b0 = ld4(bp + (size_t)(4 * k4 + 0) * ldb);
            const f4 b1 = ld4(bp + (size_t)(4 * k4 + 1) * ldb);
            const f4 b2 = ld4(bp + (size_t)(4 * k4 + 2) * ldb);
            const f4 b3 = ld4(bp + (size_t)(4 * k4 + 3) * ldb);
            MMX_UNROLL
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(av[i].x, b0.x, acc[i][0]); acc[i][1] = fmaf(av[i].x, b0.y, acc[i][1]);
                acc[i][2] = fmaf(av[i].x, b0.z, acc[i][2]); acc[i][3] = fmaf(av[i].x, b0.w, acc[i][3]);
                acc[i][0] = fmaf(av[i].y, b1.x, acc[i][0]); acc[i][1] = fmaf(av[i].y, b1.y, acc[i][1]);
                acc[i][2] = fmaf(av[i].y, b1.z, acc[i][2]); acc[i][3] = fmaf(av[i].y, b1.w, acc[i][3]);
                acc[i][0] = fmaf(av[i].z, b2.x, acc[i][0]); acc[i][1] = fmaf(av[i].z, b2.y, acc[i][1]);
                acc[i][2] = fmaf(av[i].z, b2.z, acc[i][2]); acc[i][3] = fmaf(av[i].z, b2.w, acc[i][3]);
                acc[i][0] = fmaf(av[i].w, b3.x, acc[i][0]); acc[i][1] = fmaf(av[i].w, b3.y, acc[i][1]);
                acc[i][2] = fmaf(av[i].w, b3.z, acc[i][2]); acc[i][3] = fmaf(av[i].w, b3.w, acc[i][3]);
            }
        }
        for (int k = 4 * K4; k < K; ++k) {   // K tail (B rows are not padded in k)
            const f4 b0 = ld4(bp + (size_t)k * ldb);
            MMX_UNROLL
            for (int i = 0; i < TM; ++i) {
                const float a = ap[i][k];
                acc[i][0] = fmaf(a, b0.x, acc[i][0]); acc[i][1] = fmaf(a, b0.y, acc[i][1]);
                acc[i][2] = fmaf(a, b0.z, acc[i][2]); acc[i][3] = fmaf(a, b0.w, acc[i][3]);
            }
        }
        MMX_UNROLL
        for (int i = 0; i < TM; ++i) {
            const int m = rt + i * n_rt;
            if (m < M) {
                MMX_UNROLL
                for (int j = 0; j < 4; ++j) {
                    const int n = 4 * ct + j;
                    if (n < N) epi(m, n, acc[i][j]);
                }
            }
        }
    }
}

// Weight-gradient GEMM with a thread-owned accumulator tile that persists across calls:
//     acc[i][j] += sum_{k<K} A[k*lda + 4*mt + i] * B[k*ldb + 4*nt + j]     ("TN", 4x4 tiles)
// Tile `tile` (< n_mt*n_nt, n_nt = ceil(N/4)) is owned by the caller; columns of A / B up to the
// next multiple of 4 must be readable and finite.
MMX_D void gemm_tn_acc4x4(float (&acc)[4][4], int tile, int n_nt, const float* A, int lda,
                          const float* B, int ldb, int K) {
    const int mt = tile / n_nt, nt = tile - mt * n_nt;
    const float* ap = A + 4 * mt;
    const float* bp = B + 4 * nt;
#if !defined(MMX_HOST_EMU)
#pragma unroll 4
#endif
    for (int k = 0; k < K; ++k) {
        const f4 a = ld4(ap + (size_t)k * lda);
        const f4 b = ld4(bp + (size_t)k * ldb);
        acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
        acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
        acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
        acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
        acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
        acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
        acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
        acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
    }
}

// store an accumulator tile into a row-major SHARED staging matrix [M][ld] (then one coalesced RED pass flushes it:
// lane-contiguous addresses instead of 16-byte-strided ones)
MMX_D void stage_acc4x4(const float (&acc)[4][4], int tile, int n_nt, float* S, int ld, int M, int N) {
    const int mt = tile / n_nt, nt = tile - mt * n_nt;
    MMX_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int m = 4 * mt + i;
        if (m < M) {
            MMX_UNROLL
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * nt + j;
                if (n < N) S[m * ld + n] = acc[i][j];
            }
        }
    }
}

// add an accumulator tile into a row-major SHARED matrix [M][ld] with shared atomics (combines the K-split slices of a CTA
// before ONE global RED per element: many threads RED-ing the same few addresses serialise in L2)
MMX_D void smem_add_acc4x4(const float (&acc)[4][4], int tile, int n_nt, float* S, int ld, int M, int N) {
    const int mt = tile / n_nt, nt = tile - mt * n_nt;
    MMX_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int m = 4 * mt + i;
        if (m < M) {
            MMX_UNROLL
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * nt + j;
                if (n < N) smem_add(S + m * ld + n, acc[i][j]);
            }
        }
    }
}

// flush an accumulator tile into a row-major global gradient [M][ldg] with RED.ADD
MMX_D void flush_acc4x4(const float (&acc)[4][4], int tile, int n_nt, float* G, int ldg, int M, int N) {
    const int mt = tile / n_nt, nt = tile - mt * n_nt;
    MMX_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int m = 4 * mt + i;
        if (m < M) {
            MMX_UNROLL
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * nt + j;
                if (n < N) red_add(G + (size_t)m * ldg + n, acc[i][j]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// row helpers (one thread per row of a row-major shared tile; pitch from pitch_of())
// ------------------------------------------------------------------------------------------
// biased-variance LayerNorm statistics of one row; pad columns must be zero
MMX_D void row_stats(const float* row, int W, float* mean, float* rstd, float eps) {
    float s = 0.0f;
    const int W4 = (W + 3) >> 2;
    for (int q = 0; q < W4; ++q) { f4 v = ld4(row + 4 * q); s += (v.x + v.y) + (v.z + v.w); }
    const float mu = s / (float)W;
    float ss = 0.0f;
    for (int h = 0; h < W; ++h) { float d = row[h] - mu; ss = fmaf(d, d, ss); }
    *mean = mu;
    *rstd = 1.0f / sqrtf(ss / (float)W + eps);
}

}  // namespace mmx
