// Micro-benchmark: does packed FP32x2 (FFMA2/FADD2, sm_100a) relieve an ISSUE-bound FP32 kernel?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float b, float c) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    int ia[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ia[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v = __ffma2_rn(make_float2(a[2 * i], a[2 * i + 1]), make_float2(b, b), make_float2(c, c));
                a[2 * i] = v.x; a[2 * i + 1] = v.y;
            }
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) ia[i] = (ia[i] ^ it) + (ia[(i + 1) & 7] >> 1);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    int si = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) si += ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)si;
}

template <int MODE>
float run(float* d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const double fma_total = 148.0 * 8 * 256 * 16.0 * iters;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    printf("scalar FFMA          : %.3f ms  %.1f TFLOP/s\n", t0, 2 * fma_total / t0 / 1e9);
    printf("packed FFMA2         : %.3f ms  %.1f TFLOP/s\n", t1, 2 * fma_total / t1 / 1e9);
    printf("scalar FFMA  + 16 int: %.3f ms\n", t2);
    printf("packed FFMA2 + 16 int: %.3f ms\n", t3);
    return 0;
}
