// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, N, K=8) in SS mode for aligned / shifted A descriptors.
#include <cstdio>
#include <cuda_runtime.h>
#include "../dnpy_b200/csrc/tc_common.cuh"
using namespace dnb::tc;
template <int BN>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int shift, int reps, int ksteps) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem; uint8_t* sB = smem + 512 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int e = threadIdx.x; e < (512 + 256) * 32; e += 128) reinterpret_cast<float*>(smem)[e] = 0.001f * (e % 97);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(slot, 256); tmem_relinquish(); }
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
        const uint32_t a = smem_u32(sA) + shift * 128, b = smem_u32(sB);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int j = 0; j < ksteps; ++j)
                mma_tf32(tm, make_smem_desc(a + (j & 3) * 32 + (j >> 2) * 128 * 16, 16, 1024), make_smem_desc(b + (j & 3) * 32, 16, 1024), idesc, 1);
        mma_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc));
}
template <int BN>
__global__ void __launch_bounds__(128, 1) rate_kernel2(long long* out, int shift, int reps, int ksteps) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem; uint8_t* sB = smem + 512 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int e = threadIdx.x; e < (512 + 256) * 32; e += 128) reinterpret_cast<float*>(smem)[e] = 0.001f * (e % 97);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(slot, 256); tmem_relinquish(); }
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x < 32) {          // whole warp runs the loop, one elected lane issues
        const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
        const uint64_t a0 = make_smem_desc(smem_u32(sA) + shift * 128, 16, 1024), b0 = make_smem_desc(smem_u32(sB), 16, 1024);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int kb = 0; kb < ksteps / 4; ++kb) {
                const uint64_t a = a0 + (uint64_t)(kb * 128 * 8 / 16), b = b0;
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_acc(tm, a + 2 * j, b + 2 * j, idesc);
                }
            }
        if (elect_one()) mma_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}
template <int BN> void run2(long long* d, int shift) {
    size_t smem = (512 + 256) * 128 + 64 + 1024;
    cudaFuncSetAttribute(rate_kernel2<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 200, ks = 36;
    rate_kernel2<BN><<<1, 128, smem>>>(d, shift, reps, ks);
    cudaDeviceSynchronize();
    rate_kernel2<BN><<<1, 128, smem>>>(d, shift, reps, ks);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("v2 N=%3d shift=%2d: %.1f cycles / MMA (%s)\n", BN, shift, (double)c / (reps * ks), cudaGetErrorString(e));
}
template <int BN> void run(long long* d, int shift) {
    size_t smem = (512 + 256) * 128 + 64 + 1024;
    cudaFuncSetAttribute(rate_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 200, ks = 36;
    rate_kernel<BN><<<1, 128, smem>>>(d, shift, reps, ks);
    cudaDeviceSynchronize();
    rate_kernel<BN><<<1, 128, smem>>>(d, shift, reps, ks);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d shift=%2d: %.1f cycles / MMA (%s)\n", BN, shift, (double)c / (reps * ks), cudaGetErrorString(e));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    for (int shift : {0, 1, 3, 8, 14, 29}) { run<64>(d, shift); }
    for (int shift : {0, 1, 14}) { run2<16>(d, shift); run2<32>(d, shift); run2<64>(d, shift); run2<128>(d, shift); run2<256>(d, shift); }
    return 0;
}
