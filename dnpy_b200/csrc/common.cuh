// Shared helpers for the dnpy_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/dnpy_b200.h"

namespace dnb {

// SM count of the current device (B200: 148 = 2 dies x 74), queried once per process; grids are sized in multiples of it
int num_sms();
#define kNumSMs (dnb::num_sms())

// thread-local error string behind dnb_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

// Every entry point converts its stream handle through S(); the last one is remembered so the launch
// bookkeeping below (launch counter, optional per-kernel event chain) knows where the kernel went.
extern thread_local cudaStream_t g_cur_stream;
inline cudaStream_t S(dnb_stream_t s) { g_cur_stream = reinterpret_cast<cudaStream_t>(s); return g_cur_stream; }

// Called after every kernel launch: counts it and, while profiling (dnb_prof_begin/end), records a
// CUDA event on the launching stream so consecutive events bracket each kernel.
void note_launch(const char* name);

#define DNB_CHECK_ARG(cond, ...) \
    do { if (!(cond)) return dnb::set_err(DNB_ERR_INVALID, __VA_ARGS__); } while (0)

#define DNB_LAUNCH_CHECK(name) \
    do { cudaError_t e_ = cudaGetLastError(); \
         if (e_ != cudaSuccess) return dnb::set_err(DNB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_)); \
         dnb::note_launch(name); \
    } while (0)

// Grid for a grid-stride elementwise kernel: enough CTAs for >= 2 waves at 8 CTAs/SM, capped.
inline int ew_grid(size_t n_vec, int block) {
    size_t need = (n_vec + block - 1) / block;
    size_t cap = (size_t)kNumSMs * 16;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast through smem to all).
template <int BLOCK>
__device__ __forceinline__ float block_sum(float v, float* smem /* >= 32 floats */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    float r = (threadIdx.x < BLOCK / 32) ? smem[threadIdx.x] : 0.f;
    if (wid == 0) r = warp_sum(r);
    if (threadIdx.x == 0) smem[0] = r;
    __syncthreads();
    return smem[0];
}

__device__ __forceinline__ float reg_term(float p, float l1, float l2) {
    // lambda_l1*sign(p) + lambda_l2*2p  (dnpy/regularizers.py:25-26,38-39)
    float s = (p > 0.f) ? 1.f : ((p < 0.f) ? -1.f : 0.f);
    return l1 * s + l2 * 2.f * p;
}

}  // namespace dnb
