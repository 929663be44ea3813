// Data-parallel gradient exchange over NVLink / NVSwitch PEER MEMORY (Net.fit's minibatch sharded over the GPUs of one box,
// dnpy/net.py:176-195; the reference has no distributed code — SURVEY.md 8e).
//
// The flat gradient arena of every rank lives in CUDA-IPC memory that all ranks of the node map (dnb_ipc_*).  The exchange
// is ONE kernel per step on every rank (dnb_dp_allreduce), captured in the step graph like any other kernel — no NCCL call,
// no host round trip, no communicator stream:
//     1. CTA 0 publishes "my gradients are final" (epoch number, st.release.sys) in every peer's flag block and waits until
//        every peer has published the same; the other CTAs of the grid are released through a local flag;
//     2. every CTA adds its slice of the N arenas — its own and the peers' through NVLink (16-byte system-scope loads) —
//        in rank order 0..N-1 (the same order on every rank: the replicas stay bit-identical) into a local scratch;
//     3. the last CTA to finish publishes "I have read everybody" and waits for the same from every peer — only then may an
//        arena be overwritten;
//     4. every CTA copies its slice of the scratch back into the arena: grads now hold the global sum, as after an all-reduce.
// A 140 KB arena (the CNN) costs a few microseconds of NVLink round trips instead of a latency-bound ncclAllReduce (~30-60 us
// at 2-8 ranks), which was the whole scaling loss of the 0.6 ms step.  Spins are bounded: a protocol error traps (the
// launch fails) instead of hanging the GPU.
#include "common.cuh"
#include <algorithm>

namespace dnb {

constexpr int DP_MAX_RANKS = 8;
constexpr int DP_CTAS = 32;                  // co-resident by construction (<< 148 SMs, nothing else on the stream)
// flag block of a rank (unsigned words, IPC memory, zeroed at allocation):
//   [0..7]  arrive1[src]  gradients of rank src are final (epoch)        [8..15] arrive2[src]  rank src has read every arena (epoch)
//   [16]    go1           local release of phase 2                       [17]    go2           local release of phase 4
//   [18]    done          CTAs of this rank that finished phase 2 (monotonic)
//   [32+b]  epoch of CTA b (each CTA owns its word)
constexpr int DP_FLAG_WORDS = 32 + DP_CTAS;

struct DpComm {
    float* bufs[DP_MAX_RANKS];
    unsigned* flags[DP_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// epochs only grow; compare with wrap-around tolerance
__device__ __forceinline__ bool reached(unsigned have, unsigned want) { return (int)(have - want) >= 0; }
template <bool SYS>
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned want) {
    for (unsigned spins = 0; !reached(SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p), want); ++spins) {
        if (spins > (1u << 27)) __trap();           // ~seconds: a rank died or the protocol is broken
        __nanosleep(20);
    }
}

__global__ void __launch_bounds__(256) dp_allreduce_kernel(DpComm c, float* __restrict__ scratch, size_t n4, int rank, int world) {
    __shared__ unsigned s_epoch, s_last;
    unsigned* my = c.flags[rank];
    const int tid = threadIdx.x;
    if (tid == 0) s_epoch = my[32 + blockIdx.x] + 1u;
    __syncthreads();
    const unsigned epoch = s_epoch;
    // 1. every rank's gradients are final
    if (blockIdx.x == 0) {
        if (tid < world) {
            __threadfence_system();
            st_release_sys(c.flags[tid] + rank, epoch);
            spin_until<true>(my + tid, epoch);
        }
        __syncthreads();
        if (tid == 0) st_release_gpu(my + 16, epoch);
    } else {
        if (tid == 0) spin_until<false>(my + 16, epoch);
        __syncthreads();
    }
    // 2. the sum of the N arenas, ranks in order
    float4* s4 = reinterpret_cast<float4*>(scratch);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r0 = 0; r0 < world; r0 += 4) {                  // up to 4 peer loads in flight
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = r0 + u < world ? ld_sys_f4(reinterpret_cast<const float4*>(c.bufs[r0 + u]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        s4[i] = acc;
    }
    // 3. everybody has read everybody
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(my + 18, 1u) + 1u == epoch * gridDim.x) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        if (tid < world) {
            st_release_sys(c.flags[tid] + 8 + rank, epoch);
            spin_until<true>(my + 8 + tid, epoch);
        }
        __syncthreads();
        if (tid == 0) st_release_gpu(my + 17, epoch);
    } else {
        if (tid == 0) spin_until<false>(my + 17, epoch);
        __syncthreads();
    }
    // 4. the arena holds the global sum
    float4* g4 = reinterpret_cast<float4*>(c.bufs[rank]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < n4; i += stride) g4[i] = s4[i];
    if (tid == 0) my[32 + blockIdx.x] = epoch;
}

}  // namespace dnb

using namespace dnb;

extern "C" {

size_t dnb_dp_flag_bytes(void) { return DP_FLAG_WORDS * sizeof(unsigned); }

/* memory every rank of the node can map: cudaMalloc + cudaIpcGetMemHandle (zero filled); handle_out: 64 bytes */
int dnb_ipc_alloc(void** ptr, size_t bytes, void* handle_out) {
    DNB_CHECK_ARG(ptr && handle_out && bytes > 0, "ipc_alloc: bad arguments");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "ipc_alloc: %s", cudaGetErrorString(e));
    cudaMemset(*ptr, 0, bytes);
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return set_err(DNB_ERR_CUDA, "ipc_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    memcpy(handle_out, &h, sizeof(h));
    return DNB_OK;
}
int dnb_ipc_open(const void* handle, void** ptr) {
    DNB_CHECK_ARG(ptr && handle, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "ipc_open: %s", cudaGetErrorString(e));
    return DNB_OK;
}
int dnb_ipc_close(void* ptr) { return ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess ? set_err(DNB_ERR_CUDA, "ipc_close failed") : DNB_OK; }
int dnb_ipc_free(void* ptr) { return ptr && cudaFree(ptr) != cudaSuccess ? set_err(DNB_ERR_CUDA, "ipc_free failed") : DNB_OK; }

/* In-place sum over the ranks of a float arena that lives in IPC memory on every rank.  bufs[r] / flags[r]: rank r's arena and
 * flag block (dnb_dp_flag_bytes, zeroed) as mapped in THIS process; n must be a multiple of 4; scratch: n floats of local
 * memory.  Every rank must make the same sequence of calls (one kernel each, stream ordered, graph capturable). */
int dnb_dp_allreduce(void* const* bufs, void* const* flags, float* scratch, size_t n, int rank, int world, dnb_stream_t stream) {
    DNB_CHECK_ARG(bufs && flags && scratch && world >= 1 && world <= DP_MAX_RANKS && rank >= 0 && rank < world, "dp_allreduce: bad arguments");
    DNB_CHECK_ARG(n % 4 == 0, "dp_allreduce: the arena must be a whole number of 16-byte groups");
    if (n == 0 || world == 1) return DNB_OK;
    DpComm c{};
    for (int r = 0; r < world; ++r) {
        DNB_CHECK_ARG(bufs[r] && flags[r] && !(reinterpret_cast<uintptr_t>(bufs[r]) & 15), "dp_allreduce: null / misaligned arena of rank %d", r);
        c.bufs[r] = static_cast<float*>(bufs[r]);
        c.flags[r] = static_cast<unsigned*>(flags[r]);
    }
    const size_t n4 = n / 4;
    const int grid = (int)std::min<size_t>(DP_CTAS, (n4 + 255) / 256);
    dp_allreduce_kernel<<<grid, 256, 0, S(stream)>>>(c, scratch, n4, rank, world);
    DNB_LAUNCH_CHECK("dp_allreduce");
    return DNB_OK;
}

}  // extern "C"
