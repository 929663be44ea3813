// Microbenchmark: steady-state cycles per tcgen05.mma (SS mode, cta_group::1) with a minimal issue loop —
// fixed, precomputed descriptors, 8 MMAs unrolled per iteration inside one elected-thread region.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mma_rate2 scratch/mma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../dnpy_b200/csrc/tc_common.cuh"
using namespace dnb::tc;
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
// kind::f16 with bf16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int M, int BN, bool BF16, bool MN = false>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int shift) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem; uint8_t* sB = smem + 512 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 512 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int e = threadIdx.x; e < (512 + 512) * 32; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = BF16 ? (make_idesc_bf16(M, BN) | (MN ? (3u << 15) : 0u)) : make_idesc_tf32(M, BN, MN, MN);
        // MN-major: tf32 = 32B-atom swizzle (4-row atoms 512 B apart along K, MN chunks of 32 elements LBO apart);
        //           bf16 = 128B swizzle (8-row atoms 1024 B apart along K, MN chunks of 64 elements LBO apart)
        const uint64_t a0 = !MN ? make_smem_desc(smem_u32(sA) + shift * 128, 16, 1024)
                          : BF16 ? make_smem_desc(smem_u32(sA) + shift * 128, 16384, 1024)
                                 : make_smem_desc(smem_u32(sA) + shift * 128, 16384, 512, LAYOUT_SW128_BASE32B);
        const uint64_t b0 = !MN ? make_smem_desc(smem_u32(sB), 16, 1024)
                          : BF16 ? make_smem_desc(smem_u32(sB), shift ? 256 : 8192, 1024)
                                 : make_smem_desc(smem_u32(sB), shift ? 128 : 8192, 512, LAYOUT_SW128_BASE32B);
        const int kadv = MN ? (BF16 ? 128 : 64) : 2;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int r = 0; r < iters; ++r) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // two accumulators alternate (like a double-buffered epilogue); 4 k-steps inside a 128-byte row
                    if (BF16) mma_f16(tm + (j & 4 ? BN : 0), a0 + kadv * (j & 3), b0 + kadv * (j & 3), idesc);
                    else mma_tf32_acc(tm + (j & 4 ? BN : 0), a0 + kadv * (j & 3), b0 + kadv * (j & 3), idesc);
                }
            }
            mma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
        long long tt = __shfl_sync(0xffffffff, t0, 0);   // elected lane is lane 0
        if (threadIdx.x == 0) out[0] = t1 - tt;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}
template <int M, int BN, bool BF16, bool MN = false> void run(long long* d, int shift) {
    size_t smem = (512 + 512) * 128 + 64 + 1024;
    cudaFuncSetAttribute(rate<M, BN, BF16, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    rate<M, BN, BF16, MN><<<1, 128, smem>>>(d, iters, shift);
    cudaDeviceSynchronize();
    rate<M, BN, BF16, MN><<<1, 128, smem>>>(d, iters, shift);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%s %s M=%3d N=%3d shift=%2d: %6.1f cycles / MMA (%s)\n", MN ? "MN-major" : "K-major ", BF16 ? "bf16" : "tf32", M, BN, shift, (double)c / (iters * 8.0), cudaGetErrorString(e));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    run<128, 16, false>(d, 0); run<128, 32, false>(d, 0); run<128, 64, false>(d, 0); run<128, 128, false>(d, 0); run<128, 256, false>(d, 0);
    run<128, 32, false>(d, 3); run<128, 64, false>(d, 3); run<128, 64, false>(d, 14);
    run<64, 32, false>(d, 0); run<64, 64, false>(d, 0); run<64, 128, false>(d, 0); run<64, 256, false>(d, 0);
    run<128, 32, true>(d, 0); run<128, 64, true>(d, 0); run<128, 128, true>(d, 0); run<128, 256, true>(d, 0);
    run<128, 64, true>(d, 3); run<64, 64, true>(d, 0); run<64, 256, true>(d, 0);
    run<64, 96, false, true>(d, 0); run<64, 96, false, true>(d, 1); run<128, 96, false, true>(d, 0); run<128, 96, false, true>(d, 1);
    run<64, 64, false, true>(d, 0); run<128, 128, false, true>(d, 0); run<128, 192, false, true>(d, 0);
    run<64, 96, true, true>(d, 0); run<64, 96, true, true>(d, 1); run<128, 96, true, true>(d, 0); run<128, 96, true, true>(d, 1);
    run<64, 64, true, true>(d, 0); run<128, 128, true, true>(d, 0); run<128, 192, true, true>(d, 0); run<64, 192, true, true>(d, 0);
    run<128, 192, false>(d, 0); run<128, 192, true>(d, 0); run<64, 192, true>(d, 0); run<64, 96, true>(d, 0); run<64, 96, false>(d, 0);
    return 0;
}
