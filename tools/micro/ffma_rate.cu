// FP32 FMA issue-rate microbenchmark on sm_100a: which operand forms run at full rate?
//   reg    : FFMA  R, R, R, R      (three register sources)
//   reg2   : FFMA2 (packed pair)
//   cbank  : FFMA  R, R, c[0x3][imm], R   (weight in the constant bank, compile-time offset)
//   ureg   : FFMA  R, R, UR, R            (weight in a uniform register: warp-uniform value loaded with a uniform load)
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float cw[64];
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* wg, int iters) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float x[4]; for (int j = 0; j < 4; ++j) x[j] = wg[40 + j] + threadIdx.x * 1e-6f;
    float wr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[i] = wg[(MODE == 3 ? blockIdx.x % 4 * 8 : threadIdx.x % 4 * 8) + i];   // MODE 3: warp-uniform index
    for (int r = 0; r < iters; ++r) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 2) acc[i] = fmaf(x[j], cw[i + 8 * j], acc[i]);
                else acc[i] = fmaf(x[j], wr[i], acc[i]);
            }
        x[0] += 1e-7f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k2(float* out, const float* wg, int iters) {
    float2 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i);
    float x[4]; for (int j = 0; j < 4; ++j) x[j] = wg[40 + j] + threadIdx.x * 1e-6f;
    float2 wr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wr[i] = make_float2(wg[threadIdx.x % 4 * 8 + i], wg[threadIdx.x % 4 * 8 + i + 4]);
    for (int r = 0; r < iters; ++r) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = __ffma2_rn(make_float2(x[j], x[j]), wr[i], acc[i]);
        x[0] += 1e-7f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
}
int main() {
    float *d, *w; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&w, 64 * 4);
    float hw[64]; for (int i = 0; i < 64; ++i) hw[i] = 0.001f * i;
    cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice); cudaMemcpyToSymbol(cw, hw, sizeof(hw));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[] = {"reg  FFMA ", "reg  FFMA2", "cbank FFMA", "ureg FFMA "};
    for (int pass = 0; pass < 2; ++pass)
        for (int m = 0; m < 4; ++m) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<148 * 8, 256>>>(d, w, iters);
            if (m == 1) k2<<<148 * 8, 256>>>(d, w, iters);
            if (m == 2) k<2><<<148 * 8, 256>>>(d, w, iters);
            if (m == 3) k<3><<<148 * 8, 256>>>(d, w, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = 148.0 * 8 * 256 * (double)iters * 32;
            if (pass) printf("%s: %.3f ms, %.1f TFLOP/s\n", names[m], ms, 2 * fma / ms / 1e9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
