// Classifier heads: Dense with out <= 16 on wide inputs (dnpy/layers/core.py:116-135), e.g. the 10-class
// layer on 1600 (MNIST CNN) or 16384 (residual net) features.  These are GEMV-like and HBM bound: the only
// large operand is X (forward, dW) or dX (backward), and each kernel streams it exactly once with 128-bit
// accesses, OUT known at compile time so every accumulator and weight index is a register.
//
//   forward : CTA = 32 rows x 8 feature lanes; 8 consecutive threads read 128 contiguous bytes of one row
//             (16 LDG.128 in flight per thread), W arrives in 512-feature chunks through a double-buffered
//             cp.async.bulk (TMA) stage and is read back as the thread's 4*OUT contiguous floats; the 8 lanes
//             of a row are reduced with 3 shuffle steps.  Deterministic (no atomics).
//   dX      : a thread keeps the 4 x OUT weights of its feature quad in registers and walks 32 rows whose dY
//             sits in shared memory; one STG.128 per row.
//   dW / db : a warp owns 32 feature quads (512 contiguous bytes per row), the 8 warps of a CTA take
//             interleaved rows of a 256-row block; X rows are fetched 8 deep, dY rows are broadcast from shared
//             memory; the warps are summed in shared memory and one atomic per (feature, out) leaves the CTA.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {

constexpr int HD_MAXOUT = 16;
constexpr int HF_ROWS = 32, HF_CH = 512;          // forward: rows per CTA, features per W chunk
constexpr int HX_ROWS = 32;                       // dX: rows per thread
constexpr int HW_ROWS = 256;                      // dW: rows per CTA

__device__ __forceinline__ void hd_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// ---- forward ---------------------------------------------------------------------------------------
template <int OUT>
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, int B, int in) {
    extern __shared__ __align__(128) uint8_t hsm[];
    float (*sw)[HF_CH * OUT] = reinterpret_cast<float (*)[HF_CH * OUT]>(hsm);          // [2][HF_CH * OUT]
    uint64_t* full = reinterpret_cast<uint64_t*>(hsm + 2 * HF_CH * OUT * sizeof(float));
    const int tid = threadIdx.x, ks = tid & 7, rl = tid >> 3;
    const int row = blockIdx.x * HF_ROWS + rl;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)min(row, B - 1) * in);
    const int in4 = in >> 2;
    const int nch = (in + HF_CH - 1) / HF_CH;
    if (tid == 0) {
        tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
        tc::fence_barrier_init();
        const uint32_t bytes = (uint32_t)min(HF_CH, in) * OUT * 4u;
        tc::mbar_arrive_expect_tx(&full[0], bytes);
        hd_bulk_g2s(sw[0], w, bytes, &full[0]);
    }
    __syncthreads();
    float acc[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc[o] = 0.f;
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) {
            __syncthreads();                               // everyone is done with buffer (c+1)&1 (chunk c-1)
            if (tid == 0) {
                const int c1 = (c + 1) * HF_CH;
                const uint32_t bytes = (uint32_t)min(HF_CH, in - c1) * OUT * 4u;
                tc::mbar_arrive_expect_tx(&full[(c + 1) & 1], bytes);
                hd_bulk_g2s(sw[(c + 1) & 1], w + (size_t)c1 * OUT, bytes, &full[(c + 1) & 1]);
            }
        }
        const int q0 = c * (HF_CH / 4);                    // first quad of the chunk
        float4 xv[HF_CH / 32];
#pragma unroll
        for (int t = 0; t < HF_CH / 32; ++t) {
            const int q = q0 + ks + 8 * t;
            xv[t] = q < in4 ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tc::mbar_wait(&full[c & 1], (c >> 1) & 1);
        const float* swc = sw[c & 1];
#pragma unroll
        for (int t = 0; t < HF_CH / 32; ++t) {
            if (q0 + ks + 8 * t < in4) {
                const float4* wq = reinterpret_cast<const float4*>(swc + (ks + 8 * t) * 4 * OUT);
                const float xs[4] = {xv[t].x, xv[t].y, xv[t].z, xv[t].w};
#pragma unroll
                for (int m = 0; m < OUT; ++m) {
                    const float4 w4 = wq[m];
                    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const int f = 4 * m + n;           // flat index into [4 features][OUT]
                        acc[f % OUT] = fmaf(xs[f / OUT], wv[n], acc[f % OUT]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        float v = acc[o];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        acc[o] = v;
    }
    if (ks == 0 && row < B) {
#pragma unroll
        for (int o = 0; o < OUT; ++o) y[(size_t)row * OUT + o] = acc[o] + (bias ? bias[o] : 0.f);
    }
}

// ---- dX ----------------------------------------------------------------------------------------------
template <int OUT>
__global__ void __launch_bounds__(256) head_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                      float* __restrict__ dx, int B, int in) {
    __shared__ __align__(16) float sd[HX_ROWS][HD_MAXOUT];
    const int b0 = blockIdx.y * HX_ROWS, nr = min(HX_ROWS, B - b0);
    for (int e = threadIdx.x; e < HX_ROWS * HD_MAXOUT; e += 256) {
        const int r = e / HD_MAXOUT, o = e % HD_MAXOUT;
        sd[r][o] = (r < nr && o < OUT) ? __ldg(dy + (size_t)(b0 + r) * OUT + o) : 0.f;
    }
    __syncthreads();
    const int in4 = in >> 2;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= in4) return;
    float wr[4 * OUT];                                      // [4 features][OUT], contiguous in W
    const float4* wq = reinterpret_cast<const float4*>(w + (size_t)q * 4 * OUT);
#pragma unroll
    for (int m = 0; m < OUT; ++m) {
        const float4 w4 = __ldg(wq + m);
        wr[4 * m] = w4.x; wr[4 * m + 1] = w4.y; wr[4 * m + 2] = w4.z; wr[4 * m + 3] = w4.w;
    }
    float4* xo = reinterpret_cast<float4*>(dx + (size_t)b0 * in) + q;
    for (int r = 0; r < nr; ++r) {
        float d[HD_MAXOUT];
#pragma unroll
        for (int o4 = 0; o4 < (OUT + 3) / 4; ++o4) {
            const float4 t = *reinterpret_cast<const float4*>(&sd[r][4 * o4]);
            d[4 * o4] = t.x; d[4 * o4 + 1] = t.y; d[4 * o4 + 2] = t.z; d[4 * o4 + 3] = t.w;
        }
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < OUT; ++o)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = fmaf(d[o], wr[j * OUT + o], a[j]);
        xo[(size_t)r * in4] = make_float4(a[0], a[1], a[2], a[3]);
    }
}

// ---- dW / db -------------------------------------------------------------------------------------------
template <int OUT>
__global__ void __launch_bounds__(256) head_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                      const float* __restrict__ w, float* __restrict__ gw,
                                                      const float* __restrict__ bvec, float* __restrict__ gb,
                                                      int B, int in, float inv_m, float l1, float l2, float bl1, float bl2) {
    constexpr int PITCH = 4 * OUT + 1;
    extern __shared__ __align__(128) uint8_t hsm[];
    float (*sd)[HD_MAXOUT] = reinterpret_cast<float (*)[HD_MAXOUT]>(hsm);                 // [HW_ROWS][16]
    float (*red)[32][PITCH] = reinterpret_cast<float (*)[32][PITCH]>(hsm + HW_ROWS * HD_MAXOUT * sizeof(float));   // [8][32][PITCH]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b0 = blockIdx.y * HW_ROWS, nr = min(HW_ROWS, B - b0);
    for (int e = threadIdx.x; e < HW_ROWS * HD_MAXOUT; e += 256) {
        const int r = e / HD_MAXOUT, o = e % HD_MAXOUT;
        sd[r][o] = (r < nr && o < OUT) ? __ldg(dy + (size_t)(b0 + r) * OUT + o) : 0.f;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < OUT) {             // db: column sums of this row block
        float s = 0.f;
        for (int r = 0; r < nr; ++r) s += sd[r][threadIdx.x];
        float g = s * inv_m;
        if (blockIdx.y == 0) g += reg_term(bvec[threadIdx.x], bl1, bl2) * inv_m;
        atomicAdd(gb + threadIdx.x, g);
    }
    const int in4 = in >> 2;
    const int q = blockIdx.x * 32 + lane;
    const bool qok = q < in4;
    float acc[4 * OUT];                                     // [4 features][OUT]
#pragma unroll
    for (int k = 0; k < 4 * OUT; ++k) acc[k] = 0.f;
    const float4* xq = reinterpret_cast<const float4*>(x + (size_t)b0 * in) + min(q, in4 - 1);
    for (int r0 = wid; r0 < nr; r0 += 64) {                 // rows wid, wid+8, ... : 8 in flight
        float4 xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + 8 * u;
            xv[u] = (r < nr && qok) ? __ldg(xq + (size_t)r * in4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = min(r0 + 8 * u, HW_ROWS - 1);     // rows >= nr hold zeros in sd
            float d[HD_MAXOUT];
#pragma unroll
            for (int o4 = 0; o4 < (OUT + 3) / 4; ++o4) {
                const float4 t = *reinterpret_cast<const float4*>(&sd[r][4 * o4]);
                d[4 * o4] = t.x; d[4 * o4 + 1] = t.y; d[4 * o4 + 2] = t.z; d[4 * o4 + 3] = t.w;
            }
            const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int o = 0; o < OUT; ++o) acc[j * OUT + o] = fmaf(xs[j], d[o], acc[j * OUT + o]);
        }
    }
#pragma unroll
    for (int k = 0; k < 4 * OUT; ++k) red[wid][lane][k] = acc[k];
    __syncthreads();
    // 32 quads x 4*OUT values of this CTA: feature-major in W, so (quad, k) -> gw[(4*quad)*OUT + k]
    for (int e = threadIdx.x; e < 32 * 4 * OUT; e += 256) {
        const int ql = e / (4 * OUT), k = e - ql * (4 * OUT);
        const int qg = blockIdx.x * 32 + ql;
        if (qg < in4) {
            float s = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += red[ww][ql][k];
            const size_t gi = (size_t)qg * 4 * OUT + k;
            float g = s * inv_m;
            if (blockIdx.y == 0) g += reg_term(w[gi], l1, l2) * inv_m;
            atomicAdd(gw + gi, g);
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool head_supported(int B, int in, int out) {
    return out >= 1 && out <= HD_MAXOUT && B >= 64 && (in & 3) == 0 && in >= 64 && !getenv("DNB_NO_HEAD");
}

#define HD_DISPATCH(OUTV, ...)                                                                         \
    switch (OUTV) {                                                                                      \
        case 1: { constexpr int O = 1; __VA_ARGS__; } break;   case 2: { constexpr int O = 2; __VA_ARGS__; } break;    \
        case 3: { constexpr int O = 3; __VA_ARGS__; } break;   case 4: { constexpr int O = 4; __VA_ARGS__; } break;    \
        case 5: { constexpr int O = 5; __VA_ARGS__; } break;   case 6: { constexpr int O = 6; __VA_ARGS__; } break;    \
        case 7: { constexpr int O = 7; __VA_ARGS__; } break;   case 8: { constexpr int O = 8; __VA_ARGS__; } break;    \
        case 9: { constexpr int O = 9; __VA_ARGS__; } break;   case 10: { constexpr int O = 10; __VA_ARGS__; } break;  \
        case 11: { constexpr int O = 11; __VA_ARGS__; } break; case 12: { constexpr int O = 12; __VA_ARGS__; } break;  \
        case 13: { constexpr int O = 13; __VA_ARGS__; } break; case 14: { constexpr int O = 14; __VA_ARGS__; } break;  \
        case 15: { constexpr int O = 15; __VA_ARGS__; } break; default: { constexpr int O = 16; __VA_ARGS__; } break;  \
    }

// returns false when the pointers are not 16-byte aligned (caller falls back)
bool head_fwd(const float* x, const float* w, const float* b, float* y, int B, int in, int out, cudaStream_t st) {
    if (!al16(x) || !al16(w)) return false;
    const int grid = (B + HF_ROWS - 1) / HF_ROWS;
    HD_DISPATCH(out, {
        const size_t smem = 2 * HF_CH * O * sizeof(float) + 16;
        if (smem > 48 * 1024) cudaFuncSetAttribute(head_fwd_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        head_fwd_kernel<O><<<grid, 256, smem, st>>>(x, w, b, y, B, in);
    });
    return true;
}

bool head_dx(const float* dy, const float* w, float* dx, int B, int in, int out, cudaStream_t st) {
    if (!al16(dx) || !al16(w)) return false;
    dim3 grid(((in >> 2) + 255) / 256, (B + HX_ROWS - 1) / HX_ROWS);
    HD_DISPATCH(out, (head_dx_kernel<O><<<grid, 256, 0, st>>>(dy, w, dx, B, in)));
    return true;
}

bool head_dw(const float* x, const float* dy, const float* w, float* gw, const float* b, float* gb, int B, int in, int out,
             float inv_m, float l1, float l2, float bl1, float bl2, cudaStream_t st) {
    if (!al16(x)) return false;
    dim3 grid(((in >> 2) + 31) / 32, (B + HW_ROWS - 1) / HW_ROWS);
    HD_DISPATCH(out, {
        const size_t smem = (HW_ROWS * HD_MAXOUT + 8 * 32 * (4 * O + 1)) * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(head_dw_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        head_dw_kernel<O><<<grid, 256, smem, st>>>(x, dy, w, gw, b, gb, B, in, inv_m, l1, l2, bl1, bl2);
    });
    return true;
}

}  // namespace dnb
