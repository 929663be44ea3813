// Experiment: does a K-major SW128 A descriptor whose start is shifted by r rows (128 B each, not 1024-aligned)
// read rows r..r+127 correctly when base_offset = (r & 7)?   Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../dnpy_b200/csrc/tc_common.cuh"
using namespace dnb::tc;

__device__ __forceinline__ uint32_t sw_off(int row, int k) {
    return (uint32_t)(row * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}
__global__ void __launch_bounds__(128, 1) shift_kernel(const float* A, const float* Bm, float* D, int shift, int use_base_off) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 256 rows x 128 B
    uint8_t* sB = smem + 256 * 128;     // 32 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 256 * 32; e += 128) { int r = e / 32, k = e % 32; *reinterpret_cast<float*>(sA + sw_off(r, k)) = A[e]; }
    for (int e = tid; e < 32 * 32; e += 128) { int r = e / 32, k = e % 32; *reinterpret_cast<float*>(sB + sw_off(r, k)) = Bm[e]; }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 32); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, 32, 0, 0);
        const uint32_t a = smem_u32(sA) + shift * 128, b = smem_u32(sB);
        for (int j = 0; j < 4; ++j) {
            uint64_t ad = make_smem_desc(a + j * 32, 16, 1024);
            if (use_base_off) ad |= (uint64_t)(shift & 7) << 49;
            mma_tf32(tm, ad, make_smem_desc(b + j * 32, 16, 1024), idesc, j > 0);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + (tid & 31)) * 32 + j] = __uint_as_float(v[j]);
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 32);
}
int main() {
    std::vector<float> A(256 * 32), B(32 * 32), D(128 * 32);
    for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
    for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    size_t smem = 256 * 128 + 32 * 128 + 64 + 1024;
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int ub = 0; ub < 2; ++ub)
        for (int shift : {0, 1, 2, 3, 7, 8, 9, 14, 15, 16, 30, 33}) {
            shift_kernel<<<1, 128, smem>>>(dA, dB, dD, shift, ub);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift %d base_off %d: CUDA error %s\n", shift, ub, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 32; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 32; ++k) ref += (double)A[(m + shift) * 32 + k] * B[n * 32 + k];
                    maxerr = fmax(maxerr, fabs(ref - D[m * 32 + n])); maxref = fmax(maxref, fabs(ref));
                }
            printf("shift %2d base_off %d: rel err %.2e %s\n", shift, ub, maxerr / maxref, maxerr / maxref < 5e-3 ? "OK" : "FAIL");
        }
    return 0;
}
