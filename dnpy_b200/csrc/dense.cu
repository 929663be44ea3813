// Dense layer (dnpy/layers/core.py:116-135): Y = XW + b, dX = dY W^T, gW += (X^T dY + reg)/m,
// gb += (sum_b dY + reg)/m.
//
// FP32 path: register-tiled SIMT SGEMM (64x64x16 tiles, 4x4 per thread) with generic operand
// strides so the NN / NT / TN contractions share one kernel.  The TF32 tensor-core path lives
// in tc_gemm.cu (tcgen05 + TMEM + TMA) and is selected with math == DNB_MATH_TF32.
#include "common.cuh"
#include "tc_gemm.cuh"
#include <algorithm>

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

// gb[n] += (sum_b dy[b,n] + reg'(b[n])) * inv_m — a CTA owns 32 columns x BG_ROWS rows (8 row lanes, 4 independent loads
// in flight per lane), coalesced; the row blocks of a column meet through one atomicAdd each (a single CTA per column strip
// walks B rows with one dependent chain per thread: 10 us for 1024 x 512 on B200, launch-latency territory otherwise).
constexpr int BG_ROWS = 128;
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, const float* __restrict__ b,
                                                        float* __restrict__ gb, int B, int N, float inv_m,
                                                        float l1, float l2) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int r_lo = blockIdx.y * BG_ROWS, r_hi = min(B, r_lo + BG_ROWS);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
        int r = r_lo + ty;
        for (; r + 24 < r_hi; r += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] += __ldg(dy + (size_t)(r + 8 * u) * N + n);
        }
        for (; r < r_hi; r += 8) acc[0] += __ldg(dy + (size_t)r * N + n);
    }
    red[ty][tx] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][tx];
        if (blockIdx.y == 0) s += reg_term(b[n], l1, l2);
        if (gridDim.y == 1) gb[n] += s * inv_m;
        else atomicAdd(gb + n, s * inv_m);
    }
}

// ------------------------------------------------------------------------------------------
// Skinny heads (out <= 16, e.g. the 10-class classifier on 1600 or 16384 features): bandwidth-bound
// GEMV-like shapes where a 64x64 tile wastes 84% of its columns.  Each kernel streams X exactly once.
// ------------------------------------------------------------------------------------------
constexpr int SKN = 16;      // padded output width

// Y[b, o] = sum_i X[b,i] W[i,o] + bias[o]: a warp owns RPW rows, lanes stride over i.  W is staged through
// shared memory in chunks of SKC features (coalesced global reads; 80-byte pitch -> conflict-free LDS.128), and each
// lane issues RPW x 8 independent X loads before the first FMA of a chunk step.
constexpr int RPW = 2, SKC = 512, SKP = SKN + 4;
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y,
                                                         int B, int in, int out) {
    __shared__ __align__(16) float sw[SKC * SKP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row0 = (blockIdx.x * 8 + wid) * RPW;
    float acc[RPW][SKN];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int o = 0; o < SKN; ++o) acc[r][o] = 0.f;
    const float* xr[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) xr[r] = x + (size_t)min(row0 + r, B - 1) * in;
    for (int c0 = 0; c0 < in; c0 += SKC) {
        const int cn = min(SKC, in - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < cn * out; e += 256) {          // W[c0 .. c0+cn) is contiguous: coalesced
            const int i = e / out, o = e - i * out;
            sw[i * SKP + o] = __ldg(w + (size_t)c0 * out + e);
        }
        if (out < SKN)
            for (int e = threadIdx.x; e < cn * (SKN - out); e += 256) {
                const int i = e / (SKN - out), o = out + e - i * (SKN - out);
                sw[i * SKP + o] = 0.f;
            }
        __syncthreads();
        for (int i0 = lane; i0 < cn; i0 += 256) {
            float xv[8][RPW];
#pragma unroll
            for (int t = 0; t < 8; ++t)
#pragma unroll
                for (int r = 0; r < RPW; ++r) xv[t][r] = (i0 + 32 * t < cn) ? __ldg(xr[r] + c0 + i0 + 32 * t) : 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (i0 + 32 * t < cn) {
                    const float4* wr = reinterpret_cast<const float4*>(sw + (i0 + 32 * t) * SKP);
#pragma unroll
                    for (int q = 0; q < SKN / 4; ++q) {
                        const float4 w4 = wr[q];
#pragma unroll
                        for (int r = 0; r < RPW; ++r) {
                            acc[r][4 * q] = fmaf(xv[t][r], w4.x, acc[r][4 * q]); acc[r][4 * q + 1] = fmaf(xv[t][r], w4.y, acc[r][4 * q + 1]);
                            acc[r][4 * q + 2] = fmaf(xv[t][r], w4.z, acc[r][4 * q + 2]); acc[r][4 * q + 3] = fmaf(xv[t][r], w4.w, acc[r][4 * q + 3]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
        for (int o = 0; o < SKN; ++o) acc[r][o] = warp_sum(acc[r][o]);
        if (lane < out && row0 + r < B) {
            float v = 0.f;
#pragma unroll
            for (int o = 0; o < SKN; ++o) if (o == lane) v = acc[r][o];
            y[(size_t)(row0 + r) * out + lane] = v + (bias ? bias[lane] : 0.f);
        }
    }
}

// dX[b, i] = sum_o dY[b,o] W[i,o]: W^T staged in shared memory ([o][in], conflict-free LDS.128 along i), a thread owns
// 4 consecutive features of TWO rows (each W quad is reused for both), stores are 128-bit and coalesced.
__global__ void __launch_bounds__(256) skinny_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                        float* __restrict__ dx, int B, int in, int out, int rows_per_cta) {
    extern __shared__ __align__(16) float sdx[];
    float* swt = sdx;                         // [out][in]
    float* sdy = sdx + (size_t)out * in;      // [rows_per_cta][SKN]
    const int b0 = blockIdx.x * rows_per_cta, nr = min(rows_per_cta, B - b0);
    for (int e = threadIdx.x; e < in * out; e += 256) {
        const int i = e / out, o = e - i * out;
        swt[o * in + i] = __ldg(w + e);
    }
    for (int e = threadIdx.x; e < rows_per_cta * SKN; e += 256) {
        const int r = e / SKN, o = e % SKN;
        sdy[e] = (r < nr && o < out) ? __ldg(dy + (size_t)(b0 + r) * out + o) : 0.f;
    }
    __syncthreads();
    const int in4 = in >> 2;
    for (int item = threadIdx.x; item < (rows_per_cta >> 1) * in4; item += 256) {
        const int rp = item / in4, c4 = item - rp * in4;
        const int r0 = 2 * rp;
        if (r0 >= nr) break;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        for (int o = 0; o < out; ++o) {
            const float4 w4 = *reinterpret_cast<const float4*>(swt + o * in + 4 * c4);
            const float d0 = sdy[r0 * SKN + o], d1 = sdy[(r0 + 1) * SKN + o];
            a0.x = fmaf(d0, w4.x, a0.x); a0.y = fmaf(d0, w4.y, a0.y); a0.z = fmaf(d0, w4.z, a0.z); a0.w = fmaf(d0, w4.w, a0.w);
            a1.x = fmaf(d1, w4.x, a1.x); a1.y = fmaf(d1, w4.y, a1.y); a1.z = fmaf(d1, w4.z, a1.z); a1.w = fmaf(d1, w4.w, a1.w);
        }
        reinterpret_cast<float4*>(dx + (size_t)(b0 + r0) * in)[c4] = a0;
        if (r0 + 1 < nr) reinterpret_cast<float4*>(dx + (size_t)(b0 + r0 + 1) * in)[c4] = a1;
    }
}

// gW[i, o] += inv_m * (sum_b X[b,i] dY[b,o] + reg'(W[i,o])): a CTA owns a slice of the batch and 256*IPT features;
// lanes run along i (coalesced X rows), dY rows are broadcast from shared memory; 8 rows of X loads are in flight per
// thread before the first FMA; partial tiles are added atomically.
template <int IPT>
__global__ void __launch_bounds__(256) skinny_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ w, float* __restrict__ gw,
                                                        const float* __restrict__ bvec, float* __restrict__ gb,
                                                        int B, int in, int out, int rows_per_cta, float inv_m, float l1, float l2,
                                                        float bl1, float bl2) {
    __shared__ __align__(16) float sd[64][SKN];
    float accb = 0.f;                 // blockIdx.x == 0, thread o < out: column sum of dY -> gb
    const int i0 = blockIdx.x * 256 * IPT;
    const int b0 = blockIdx.y * rows_per_cta, b1 = min(B, b0 + rows_per_cta);
    float acc[IPT][SKN];
#pragma unroll
    for (int t = 0; t < IPT; ++t)
#pragma unroll
        for (int o = 0; o < SKN; ++o) acc[t][o] = 0.f;
    bool iv[IPT];
#pragma unroll
    for (int t = 0; t < IPT; ++t) iv[t] = i0 + (int)threadIdx.x + t * 256 < in;
    for (int bb = b0; bb < b1; bb += 64) {
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * SKN; e += 256) {
            int r = e / SKN, o = e % SKN;
            sd[r][o] = (bb + r < b1 && o < out) ? __ldg(dy + (size_t)(bb + r) * out + o) : 0.f;
        }
        __syncthreads();
        const int nr = min(64, b1 - bb);
        if (blockIdx.x == 0 && threadIdx.x < SKN)
            for (int r = 0; r < nr; ++r) accb += sd[r][threadIdx.x];
        for (int r0 = 0; r0 < nr; r0 += 8) {
            float xv[8][IPT];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* xr = x + (size_t)(bb + r0 + u) * in + i0 + threadIdx.x;
#pragma unroll
                for (int t = 0; t < IPT; ++t) xv[u][t] = (r0 + u < nr && iv[t]) ? __ldg(xr + t * 256) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int q = 0; q < SKN / 4; ++q) {
                    const float4 d4 = *reinterpret_cast<const float4*>(&sd[min(r0 + u, 63)][4 * q]);
#pragma unroll
                    for (int t = 0; t < IPT; ++t) {
                        acc[t][4 * q] = fmaf(xv[u][t], d4.x, acc[t][4 * q]); acc[t][4 * q + 1] = fmaf(xv[u][t], d4.y, acc[t][4 * q + 1]);
                        acc[t][4 * q + 2] = fmaf(xv[u][t], d4.z, acc[t][4 * q + 2]); acc[t][4 * q + 3] = fmaf(xv[u][t], d4.w, acc[t][4 * q + 3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
        const int i = i0 + threadIdx.x + t * 256;
        if (i < in) {
#pragma unroll
            for (int o = 0; o < SKN; ++o)
                if (o < out) {
                    float g = acc[t][o] * inv_m;
                    if (blockIdx.y == 0) g += reg_term(w[(size_t)i * out + o], l1, l2) * inv_m;
                    atomicAdd(gw + (size_t)i * out + o, g);
                }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < out) {
        float g = accb * inv_m;
        if (blockIdx.y == 0) g += reg_term(bvec[threadIdx.x], bl1, bl2) * inv_m;
        atomicAdd(gb + threadIdx.x, g);
    }
}

// Narrow inputs (in <= 32, e.g. the 32 -> 1 classifier behind the RNN): lanes run along the features, a warp takes 8 batch rows,
// a CTA 64; the 8 warps are summed in shared memory and one atomic per (feature, out) leaves the CTA.  skinny_dw_kernel maps
// 512 features to a CTA: at in = 32 an eighth of its threads worked and 8 CTAs walked 128 rows each (19.6 us at B = 1024).
__global__ void __launch_bounds__(256) narrow_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ w, float* __restrict__ gw,
                                                        const float* __restrict__ bvec, float* __restrict__ gb,
                                                        int B, int in, int out, float inv_m, float l1, float l2, float bl1, float bl2) {
    __shared__ float red[8][32][SKN + 1];
    __shared__ float redb[8][SKN];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b0 = blockIdx.x * 64 + wid * 8;
    float acc[SKN], accb = 0.f;
#pragma unroll
    for (int o = 0; o < SKN; ++o) acc[o] = 0.f;
    float xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = (b0 + u < B && lane < in) ? __ldg(x + (size_t)(b0 + u) * in + lane) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (b0 + u < B) {
            const float* dr = dy + (size_t)(b0 + u) * out;
#pragma unroll
            for (int o = 0; o < SKN; ++o)
                if (o < out) acc[o] = fmaf(xv[u], __ldg(dr + o), acc[o]);
            if (lane < out) accb += __ldg(dr + lane);
        }
    }
#pragma unroll
    for (int o = 0; o < SKN; ++o) red[wid][lane][o] = acc[o];
    if (lane < SKN) redb[wid][lane] = accb;
    __syncthreads();
    for (int e = threadIdx.x; e < in * out; e += 256) {
        const int i = e / out, o = e - i * out;
        float g = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) g += red[k][i][o];
        g *= inv_m;
        if (blockIdx.x == 0) g += reg_term(w[e], l1, l2) * inv_m;
        atomicAdd(gw + e, g);
    }
    if (threadIdx.x < out) {
        float g = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) g += redb[k][threadIdx.x];
        g *= inv_m;
        if (blockIdx.x == 0) g += reg_term(bvec[threadIdx.x], bl1, bl2) * inv_m;
        atomicAdd(gb + threadIdx.x, g);
    }
}

// dense_head.cu
bool head_supported(int B, int in, int out);
bool head_fwd(const float* x, const float* w, const float* b, float* y, int B, int in, int out, cudaStream_t st);
bool head_dx(const float* dy, const float* w, float* dx, int B, int in, int out, cudaStream_t st);
bool head_dw(const float* x, const float* dy, const float* w, float* gw, const float* b, float* gb, int B, int in, int out,
             float inv_m, float l1, float l2, float bl1, float bl2, cudaStream_t st);

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_dense_fwd(const float* x, const float* w, const float* b, float* y,
                  int B, int in, int out, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && in > 0 && out > 0, "dense_fwd: bad shape B=%d in=%d out=%d", B, in, out);
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && y, "dense_fwd: null pointer");
    if (head_supported(B, in, out) && head_fwd(x, w, b, y, B, in, out, S(stream))) {
        DNB_LAUNCH_CHECK("dense_fwd");
        return DNB_OK;
    }
    if (out <= SKN && B >= 64) {
        skinny_fwd_kernel<<<(B + 8 * RPW - 1) / (8 * RPW), 256, 0, S(stream)>>>(x, w, b, y, B, in, out);
        DNB_LAUNCH_CHECK("dense_fwd");
        return DNB_OK;
    }
    if (math >= DNB_MATH_TF32 && tc_gemm_supported(B, out, in, /*a_kmajor*/ true, /*b_kmajor*/ false)) {
        // Y[B,out] = X[B,in] . W[in,out]: A is K-major, B is N-major (row-major [K,N])
        return tc_gemm(x, w, y, B, out, in, TC_A_KMAJOR | TC_B_NMAJOR, b, nullptr, 0.f, 0.f, 0.f, S(stream));
    }
    GemmArgs a{x, in, 1, w, out, 1, y, B, out, in, b, nullptr, 0.f, 0.f, 0.f, 0};
    return launch_sgemm<EPI_BIAS>("dense_fwd", a, S(stream));
}

// dX[B,in] = dY[B,out] . W^T (core.py:121), shared by dnb_dense_bwd and dnb_dense_bwd_data
static int dense_dx(const float* dy, const float* w, float* dx, int B, int in, int out, int math, cudaStream_t st) {
    int rc;
    if (out <= SKN && B >= 64) {
        const size_t dx_smem = ((size_t)out * in + 32 * SKN) * sizeof(float);
        if (head_supported(B, in, out) && head_dx(dy, w, dx, B, in, out, st)) {
            DNB_LAUNCH_CHECK("dense_bwd_dx");
        } else if ((in & 3) == 0 && dx_smem <= 160 * 1024 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
            cudaFuncSetAttribute(skinny_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dx_smem);
            skinny_dx_kernel<<<(B + 31) / 32, 256, dx_smem, st>>>(dy, w, dx, B, in, out, 32);
            DNB_LAUNCH_CHECK("dense_bwd_dx");
        } else {
            GemmArgs a{dy, out, 1, w, 1, out, dx, B, in, out, nullptr, nullptr, 0.f, 0.f, 0.f, 0};
            rc = launch_sgemm<EPI_STORE>("dense_bwd_dx", a, st);
            if (rc) return rc;
        }
        return DNB_OK;
    }
    // B(k=o, n=i) = W[i*out + o]  -> K-major B
    if (math >= DNB_MATH_TF32 && tc_gemm_supported(B, in, out, true, true))
        return tc_gemm(dy, w, dx, B, in, out, TC_A_KMAJOR | TC_B_KMAJOR, nullptr, nullptr, 0.f, 0.f, 0.f, st);
    GemmArgs a{dy, out, 1, w, 1, out, dx, B, in, out, nullptr, nullptr, 0.f, 0.f, 0.f, 0};
    return launch_sgemm<EPI_STORE>("dense_bwd_dx", a, st);
}

// The data gradient alone: for a caller that wants the parent's delta but not the parameter gradients (Net materialises
// the delta of an Input layer lazily).
int dnb_dense_bwd_data(const float* w, const float* dy, float* dx, int B, int in, int out, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && in > 0 && out > 0, "dense_bwd_data: bad shape B=%d in=%d out=%d", B, in, out);
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(w && dy && dx, "dense_bwd_data: null pointer");
    return dense_dx(dy, w, dx, B, in, out, math, S(stream));
}

int dnb_dense_bwd(const float* x, const float* w, const float* b, const float* dy,
                  float* dx, float* gw, float* gb, int B, int in, int out, float inv_m,
                  dnb_reg_t reg_w, dnb_reg_t reg_b, float reg_scale, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && in > 0 && out > 0, "dense_bwd: bad shape B=%d in=%d out=%d", B, in, out);
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && dy && gw && gb, "dense_bwd: null pointer");
    cudaStream_t st = S(stream);
    int rc;
    if (out <= SKN && B >= 64) {
        const bool head = head_supported(B, in, out);
        if (dx) {
            rc = dense_dx(dy, w, dx, B, in, out, math, st);
            if (rc) return rc;
        }
        if (head && head_dw(x, dy, w, gw, b, gb, B, in, out, inv_m, reg_w.l1 * reg_scale, reg_w.l2 * reg_scale,
                            reg_b.l1 * reg_scale, reg_b.l2 * reg_scale, st)) {
            DNB_LAUNCH_CHECK("dense_bwd_dw");
            return DNB_OK;
        }
        if (in <= 32) {
            narrow_dw_kernel<<<(B + 63) / 64, 256, 0, st>>>(x, dy, w, gw, b, gb, B, in, out, inv_m, reg_w.l1 * reg_scale,
                                                            reg_w.l2 * reg_scale, reg_b.l1 * reg_scale, reg_b.l2 * reg_scale);
            DNB_LAUNCH_CHECK("dense_bwd_dw");
            return DNB_OK;
        }
        const int itiles = (in + 511) / 512;
        int bsplit = std::max(1, std::min((B + 127) / 128, (4 * kNumSMs + itiles - 1) / itiles));
        int rows = ((B + bsplit - 1) / bsplit + 63) / 64 * 64;
        dim3 grid(itiles, (B + rows - 1) / rows);
        skinny_dw_kernel<2><<<grid, 256, 0, st>>>(x, dy, w, gw, b, gb, B, in, out, rows, inv_m, reg_w.l1 * reg_scale,
                                                   reg_w.l2 * reg_scale, reg_b.l1 * reg_scale, reg_b.l2 * reg_scale);
        DNB_LAUNCH_CHECK("dense_bwd_dw");
        return DNB_OK;
    }
    if (dx) {
        rc = dense_dx(dy, w, dx, B, in, out, math, st);
        if (rc) return rc;
    }
    // gW[in,out] += (X^T[in,B] . dY[B,out] + reg) * inv_m : A(m=i,k=b) = X[b*in + i]  -> M-major A
    if (math >= DNB_MATH_TF32 && tc_gemm_supported(in, out, B, false, false)) {
        rc = tc_gemm(x, dy, gw, in, out, B, TC_A_MMAJOR | TC_B_NMAJOR | TC_ACCUM_GRAD, nullptr, w, inv_m,
                     reg_w.l1 * reg_scale, reg_w.l2 * reg_scale, st);
    } else {
        GemmArgs a{x, 1, in, dy, out, 1, gw, in, out, B, nullptr, w, inv_m, reg_w.l1 * reg_scale, reg_w.l2 * reg_scale, 0};
        rc = launch_sgemm<EPI_GRAD>("dense_bwd_dw", a, st);
    }
    if (rc) return rc;
    bias_grad_kernel<<<dim3((out + 31) / 32, (B + BG_ROWS - 1) / BG_ROWS), 256, 0, st>>>(dy, b, gb, B, out, inv_m, reg_b.l1 * reg_scale,
                                                                                    reg_b.l2 * reg_scale);
    DNB_LAUNCH_CHECK("dense_bwd_db");
    return DNB_OK;
}

}  // extern "C"
