// FP32 issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (__ffma2_rn) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
    for (int r = 0; r < iters; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) acc[i] = __ffma2_rn(acc[i], x, y);
            else { acc[i].x = fmaf(acc[i].x, x.x, y.x); acc[i].y = fmaf(acc[i].y, x.y, y.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int pass = 0; pass < 2; ++pass)
    for (int packed = 0; packed < 2; ++packed) {
        cudaEventRecord(e0);
        if (packed) k<true><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f); else k<false><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * 8 * 256 * (double)iters * 16;
        printf("%s: %.3f ms, %.1f TFLOP/s (2 flop per fma)\n", packed ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9);
    }
    return 0;
}
