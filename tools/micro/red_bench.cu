// Microbenchmark: cost of flushing per-CTA gradient partials with red.global.add.f32.
// G CTAs each add n floats (from shared memory) into the SAME n global addresses.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void flush(float* g, int n, int rot) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = 1.0f;
    __syncthreads();
    const int off = rot ? (int)(((long long)blockIdx.x * n) / gridDim.x) : 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int j = i + off; if (j >= n) j -= n;
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + j), "f"(sm[j]) : "memory");
    }
}
__global__ void flush_v4(float* g, int n) {   // red.global.add.v4.f32 (sm_90+)
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = 1.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
        float4 v = reinterpret_cast<float4*>(sm)[i];
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}
int main() {
    float* g; cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int ns[] = {3300, 5200, 16384};
    int gs[] = {37, 74, 148, 228, 296, 592};
    for (int n : ns) for (int G : gs) for (int mode = 0; mode < 3; ++mode) {
        for (int w = 0; w < 3; ++w) { if (mode < 2) flush<<<G, 256, n * 4>>>(g, n, mode); else flush_v4<<<G, 256, n * 4>>>(g, n); }
        cudaEventRecord(a);
        for (int r = 0; r < 20; ++r) { if (mode < 2) flush<<<G, 256, n * 4>>>(g, n, mode); else flush_v4<<<G, 256, n * 4>>>(g, n); }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("n=%5d G=%3d mode=%s : %7.2f us per launch (%.1f G red-elem/s)\n", n, G, mode == 0 ? "plain " : mode == 1 ? "rotate" : "v4    ", ms / 20 * 1e3,
               (double)n * G / (ms / 20 * 1e-3) / 1e9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
