// Dense layer (dnpy/layers/core.py:116-135): Y = XW + b, dX = dY W^T, gW += (X^T dY + reg)/m,
// gb += (sum_b dY + reg)/m.
//
// FP32 path: register-tiled SIMT SGEMM (64x64x16 tiles, 4x4 per thread) with generic operand
// strides so the NN / NT / TN contractions share one kernel.  The TF32 tensor-core path lives
// in tc_gemm.cu (tcgen05 + TMEM + TMA) and is selected with math == DNB_MATH_TF32.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace dnb {

constexpr int GBM = 64, GBN = 64, GBK = 16, GPAD = 4;

enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_GRAD = 2 };

struct GemmArgs {
    const float* A; long sAm, sAk;    // A(m,k) = A[m*sAm + k*sAk]
    const float* B; long sBk, sBn;    // B(k,n) = B[k*sBk + n*sBn]
    float* C;                         // row-major M x N
    int M, N, K;
    const float* bias;                // EPI_BIAS: length N (may be null)
    const float* W;                   // EPI_GRAD: same layout as C, regulariser argument
    float inv_m, l1, l2;              // EPI_GRAD (l1,l2 already multiplied by reg_scale)
    int k_per_split;                  // EPI_GRAD: K range of one z-slice (split-K with atomic accumulate)
};

template <int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs a) {
    __shared__ __align__(16) float As[GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[GBK][GBN + GPAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const bool a_kfast = (a.sAk == 1);      // pick the loader that walks the contiguous dimension
    const bool b_nfast = (a.sBn == 1);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int k_begin = (EPI == EPI_GRAD) ? blockIdx.z * a.k_per_split : 0;
    const int k_end = (EPI == EPI_GRAD) ? min(a.K, k_begin + a.k_per_split) : a.K;
    for (int k0 = k_begin; k0 < k_end; k0 += GBK) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int idx = tid + r * 256;
            int k, m;
            if (a_kfast) { k = idx & 15; m = idx >> 4; } else { m = idx & 63; k = idx >> 6; }
            int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < a.M && gk < k_end) ? __ldg(a.A + (long)gm * a.sAm + (long)gk * a.sAk) : 0.f;
            int kk, n;
            if (b_nfast) { n = idx & 63; kk = idx >> 6; } else { kk = idx & 15; n = idx >> 4; }
            int gn = n0 + n, gk2 = k0 + kk;
            Bs[kk][n] = (gn < a.N && gk2 < k_end) ? __ldg(a.B + (long)gk2 * a.sBk + (long)gn * a.sBn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= a.N) continue;
            size_t o = (size_t)gm * a.N + gn;
            if (EPI == EPI_STORE) a.C[o] = acc[i][j];
            else if (EPI == EPI_BIAS) a.C[o] = acc[i][j] + (a.bias ? a.bias[gn] : 0.f);
            else {
                float g = acc[i][j] * a.inv_m;
                if (blockIdx.z == 0) g += reg_term(a.W[o], a.l1, a.l2) * a.inv_m;
                if (gridDim.z == 1) a.C[o] += g; else atomicAdd(a.C + o, g);
            }
        }
    }
}

template <int EPI>
static int launch_sgemm(const char* name, GemmArgs a, cudaStream_t st) {
    if (a.M == 0 || a.N == 0) return DNB_OK;
    dim3 grid((a.N + GBN - 1) / GBN, (a.M + GBM - 1) / GBM);
    a.k_per_split = a.K;
    if (EPI == EPI_GRAD) {            // skinny weight-gradient shapes: split the batch reduction over the SMs
        int tiles = grid.x * grid.y;
        int splits = max(1, min(a.K / (4 * GBK), (2 * kNumSMs + tiles - 1) / tiles));
        int per = ((a.K + splits - 1) / splits + GBK - 1) / GBK * GBK;
        a.k_per_split = per;
        grid.z = (a.K + per - 1) / per;
    }
    sgemm_kernel<EPI><<<grid, 256, 0, st>>>(a);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

// gb[n] += (sum_b dy[b,n] + reg'(b[n])) * inv_m — 32 columns x 8 row lanes per CTA, coalesced.
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, const float* __restrict__ b,
                                                        float* __restrict__ gb, int B, int N, float inv_m,
                                                        float l1, float l2) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (n < N)
        for (int r = ty; r < B; r += 8) acc += dy[(size_t)r * N + n];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][tx];
        gb[n] += (s + reg_term(b[n], l1, l2)) * inv_m;
    }
}

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_dense_fwd(const float* x, const float* w, const float* b, float* y,
                  int B, int in, int out, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && in > 0 && out > 0, "dense_fwd: bad shape B=%d in=%d out=%d", B, in, out);
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && y, "dense_fwd: null pointer");
    if (math == DNB_MATH_TF32 && tc_gemm_supported(B, out, in, /*a_kmajor*/ true, /*b_kmajor*/ false)) {
        // Y[B,out] = X[B,in] . W[in,out]: A is K-major, B is N-major (row-major [K,N])
        return tc_gemm(x, w, y, B, out, in, TC_A_KMAJOR | TC_B_NMAJOR, b, nullptr, 0.f, 0.f, 0.f, S(stream));
    }
    GemmArgs a{x, in, 1, w, out, 1, y, B, out, in, b, nullptr, 0.f, 0.f, 0.f, 0};
    return launch_sgemm<EPI_BIAS>("dense_fwd", a, S(stream));
}

int dnb_dense_bwd(const float* x, const float* w, const float* b, const float* dy,
                  float* dx, float* gw, float* gb, int B, int in, int out, float inv_m,
                  dnb_reg_t reg_w, dnb_reg_t reg_b, float reg_scale, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && in > 0 && out > 0, "dense_bwd: bad shape B=%d in=%d out=%d", B, in, out);
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && dy && gw && gb, "dense_bwd: null pointer");
    cudaStream_t st = S(stream);
    int rc;
    if (dx) {
        // dX[B,in] = dY[B,out] . W^T : B(k=o, n=i) = W[i*out + o]  -> K-major B
        if (math == DNB_MATH_TF32 && tc_gemm_supported(B, in, out, true, true)) {
            rc = tc_gemm(dy, w, dx, B, in, out, TC_A_KMAJOR | TC_B_KMAJOR, nullptr, nullptr, 0.f, 0.f, 0.f, st);
        } else {
            GemmArgs a{dy, out, 1, w, 1, out, dx, B, in, out, nullptr, nullptr, 0.f, 0.f, 0.f, 0};
            rc = launch_sgemm<EPI_STORE>("dense_bwd_dx", a, st);
        }
        if (rc) return rc;
    }
    // gW[in,out] += (X^T[in,B] . dY[B,out] + reg) * inv_m : A(m=i,k=b) = X[b*in + i]  -> M-major A
    if (math == DNB_MATH_TF32 && tc_gemm_supported(in, out, B, false, false)) {
        rc = tc_gemm(x, dy, gw, in, out, B, TC_A_MMAJOR | TC_B_NMAJOR | TC_ACCUM_GRAD, nullptr, w, inv_m,
                     reg_w.l1 * reg_scale, reg_w.l2 * reg_scale, st);
    } else {
        GemmArgs a{x, 1, in, dy, out, 1, gw, in, out, B, nullptr, w, inv_m, reg_w.l1 * reg_scale, reg_w.l2 * reg_scale, 0};
        rc = launch_sgemm<EPI_GRAD>("dense_bwd_dw", a, st);
    }
    if (rc) return rc;
    bias_grad_kernel<<<(out + 31) / 32, 256, 0, st>>>(dy, b, gb, B, out, inv_m, reg_b.l1 * reg_scale, reg_b.l2 * reg_scale);
    DNB_LAUNCH_CHECK("dense_bwd_db");
    return DNB_OK;
}

}  // extern "C"
