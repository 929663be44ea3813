// Interface of the tcgen05/TMEM/TMA TF32 GEMM (tc_gemm.cu) used by Dense and the 1x1 convolutions.
#pragma once
#include <cuda_runtime.h>

namespace dnb {

enum : int {
    TC_A_KMAJOR = 0,      // A stored [M,K] row-major (K contiguous)
    TC_A_MMAJOR = 1,      // A stored [K,M] row-major (M contiguous)
    TC_B_KMAJOR = 0,      // B stored [N,K] row-major (K contiguous)
    TC_B_NMAJOR = 2,      // B stored [K,N] row-major (N contiguous)
    TC_ACCUM_GRAD = 4,    // C += (acc + reg'(W)) * inv_m   (otherwise C = acc + bias[n])
};

// True when the shape/alignment can be served by the tensor-core kernel (TMA needs 16-byte
// aligned row pitches; tiny problems stay on the SIMT path).
bool tc_gemm_supported(int M, int N, int K, bool a_kmajor, bool b_kmajor);

// C[M,N] (row-major fp32) = A·B with TF32 operands and fp32 accumulation in TMEM.
int tc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int flags,
            const float* bias, const float* W, float inv_m, float l1, float l2, cudaStream_t st);

}  // namespace dnb
