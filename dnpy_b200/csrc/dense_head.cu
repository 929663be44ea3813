// Classifier heads: Dense with out <= 16 on wide inputs (dnpy/layers/core.py:116-135), e.g. the 10-class
// layer on 1600 (MNIST CNN) or 16384 (residual net) features.  These are GEMV-like and HBM bound: the only
// large operand is X (forward, dW) or dX (backward), and each kernel streams it exactly once with 128-bit
// accesses, OUT known at compile time so every accumulator and weight index is a register.
//
//   forward : a warp owns 8 rows, its lanes the feature quads lane, lane+32, ... (512 contiguous bytes per row and load,
//             16 LDG.128 in flight per thread); W arrives in 512-feature chunks through a double-buffered
//             cp.async.bulk (TMA) stage, the 4*OUT weights of a quad are read once and reused for the 8 rows; the lanes
//             of a row are reduced with 5 shuffle steps.  Deterministic (no atomics).
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
constexpr int HF_CH = 512;                        // forward: features per W chunk
constexpr int HX_ROWS = 32;                       // dX: rows per thread
constexpr int HW_ROWS = 512;                      // dW: rows per CTA

__device__ __forceinline__ void hd_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// ---- forward ---------------------------------------------------------------------------------------
// A warp owns HF_RPW rows; its 32 lanes take the feature quads lane, lane+32, ... of every row (512 contiguous bytes per
// row and load).  The 4*OUT weights of a quad are read ONCE from shared memory (OUT LDS.128) and reused for all HF_RPW
// rows — the previous one-row-per-thread version re-read them for every row and was bound by shared-memory wavefronts
// (LDS.128 is served a quarter-warp at a time): 43 us for 52 MB on B200.
constexpr int HF_RPW = 4;                         // rows per warp
constexpr int HF_WARPS = 8;
template <int OUT>
__global__ void __launch_bounds__(HF_WARPS * 32, 2) head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y, int B, int in) {
    extern __shared__ __align__(128) uint8_t hsm[];
    float (*sw)[HF_CH * OUT] = reinterpret_cast<float (*)[HF_CH * OUT]>(hsm);          // [2][HF_CH * OUT]
    uint64_t* full = reinterpret_cast<uint64_t*>(hsm + 2 * HF_CH * OUT * sizeof(float));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = (blockIdx.x * HF_WARPS + warp) * HF_RPW;
    const int in4 = in >> 2;
    const int nch = (in + HF_CH - 1) / HF_CH;
    const float4* xr[HF_RPW];
#pragma unroll
    for (int r = 0; r < HF_RPW; ++r) xr[r] = reinterpret_cast<const float4*>(x + (size_t)min(row0 + r, B - 1) * in);
    if (tid == 0) {
        tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
        tc::fence_barrier_init();
        const uint32_t bytes = (uint32_t)min(HF_CH, in) * OUT * 4u;
        tc::mbar_arrive_expect_tx(&full[0], bytes);
        hd_bulk_g2s(sw[0], w, bytes, &full[0]);
    }
    __syncthreads();
    float acc[HF_RPW][OUT];
#pragma unroll
    for (int r = 0; r < HF_RPW; ++r)
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[r][o] = 0.f;
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
        const int q0 = c * (HF_CH / 4);                    // first quad of the chunk (HF_CH / 4 = 128 quads = 4 per lane)
        const float* swc = sw[c & 1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {                      // two quads per lane at a time: 2 * HF_RPW loads in flight
            float4 xv[2][HF_RPW];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int q = q0 + lane + 32 * (2 * h + u);
#pragma unroll
                for (int r = 0; r < HF_RPW; ++r) xv[u][r] = q < in4 ? __ldg(xr[r] + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (h == 0) tc::mbar_wait(&full[c & 1], (c >> 1) & 1);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int ql = lane + 32 * (2 * h + u);
                if (q0 + ql < in4) {
                    float wr[4 * OUT];                      // [4 features][OUT], contiguous in W
                    const float4* wq = reinterpret_cast<const float4*>(swc + ql * 4 * OUT);
#pragma unroll
                    for (int m = 0; m < OUT; ++m) {
                        const float4 w4 = wq[m];
                        wr[4 * m] = w4.x; wr[4 * m + 1] = w4.y; wr[4 * m + 2] = w4.z; wr[4 * m + 3] = w4.w;
                    }
#pragma unroll
                    for (int r = 0; r < HF_RPW; ++r) {
                        const float xs[4] = {xv[u][r].x, xv[u][r].y, xv[u][r].z, xv[u][r].w};
#pragma unroll
                        for (int f = 0; f < 4 * OUT; ++f) acc[r][f % OUT] = fmaf(xs[f / OUT], wr[f], acc[r][f % OUT]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < HF_RPW; ++r)
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
            float v = acc[r][o];
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            acc[r][o] = v;
        }
    // lane r writes row r: select its row's sums without dynamic register indexing
#pragma unroll
    for (int r = 0; r < HF_RPW; ++r) {
        if (lane == r && row0 + r < B) {
#pragma unroll
            for (int o = 0; o < OUT; ++o) y[(size_t)(row0 + r) * OUT + o] = acc[r][o] + (bias ? bias[o] : 0.f);
        }
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
                                                      int B, int in, int rows, float inv_m, float l1, float l2, float bl1, float bl2) {
    constexpr int PITCH = 4 * OUT + 1;
    extern __shared__ __align__(128) uint8_t hsm[];
    float (*sd)[HD_MAXOUT] = reinterpret_cast<float (*)[HD_MAXOUT]>(hsm);                 // [HW_ROWS][16]
    float (*red)[32][PITCH] = reinterpret_cast<float (*)[32][PITCH]>(hsm + HW_ROWS * HD_MAXOUT * sizeof(float));   // [8][32][PITCH]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b0 = blockIdx.y * rows, nr = min(rows, B - b0);          // rows <= HW_ROWS batch rows per CTA
    const int in4 = in >> 2;
    const int q = blockIdx.x * 32 + lane;
    const bool qok = q < in4;
    const float4* xq = reinterpret_cast<const float4*>(x + (size_t)b0 * in) + min(q, in4 - 1);
    // the first batch of X rows is requested BEFORE dY is staged: both DRAM latencies overlap
    float4 xn[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int r = wid + 8 * u;
        xn[u] = (r < nr && qok) ? __ldg(xq + (size_t)r * in4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int e = threadIdx.x; e < rows * HD_MAXOUT; e += 256) {
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
    float acc[4 * OUT];                                     // [4 features][OUT]
#pragma unroll
    for (int k = 0; k < 4 * OUT; ++k) acc[k] = 0.f;
    for (int r0 = wid; r0 < nr; r0 += 64) {                 // rows wid, wid+8, ... : 8 in use, the next 8 in flight
        float4 xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = xn[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + 64 + 8 * u;
            xn[u] = (r < nr && qok) ? __ldg(xq + (size_t)r * in4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = min(r0 + 8 * u, rows - 1);        // rows >= nr hold zeros in sd
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
    const int grid = (B + HF_WARPS * HF_RPW - 1) / (HF_WARPS * HF_RPW);
    HD_DISPATCH(out, {
        const size_t smem = 2 * HF_CH * O * sizeof(float) + 16;
        if (smem > 48 * 1024) cudaFuncSetAttribute(head_fwd_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        head_fwd_kernel<O><<<grid, HF_WARPS * 32, smem, st>>>(x, w, b, y, B, in);
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
    // rows per CTA: 512 on large batches (one atomic per (feature, out) and CTA), fewer when the grid would not fill the GPU
    // (256 -> 10 at B = 1024: 4 CTAs of 512 rows took 20 us)
    const int gx = ((in >> 2) + 31) / 32;
    int rows = HW_ROWS;
    while (rows > 64 && (long)gx * ((B + rows - 1) / rows) < 2L * kNumSMs) rows >>= 1;
    dim3 grid(gx, (B + rows - 1) / rows);
    HD_DISPATCH(out, {
        const size_t smem = (HW_ROWS * HD_MAXOUT + 8 * 32 * (4 * O + 1)) * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(head_dw_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        head_dw_kernel<O><<<grid, 256, smem, st>>>(x, dy, w, gw, b, gb, B, in, rows, inv_m, l1, l2, bl1, bl2);
    });
    return true;
}

}  // namespace dnb
