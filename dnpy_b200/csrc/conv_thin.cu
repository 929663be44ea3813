// "Thin" 3x3 stride-1 convolutions: few input channels (C <= 4), e.g. the first layer of the MNIST
// CNN (1 -> F) or of the residual net (3 -> F).  K = C*9 is far below a tensor-core tile, the work is
// bandwidth / FMA bound, and the expensive side is the F-channel tensor (output in forward, dY in
// backward).  Three register-tiled fp32 kernels, each touching that tensor exactly once:
//
//   forward : a thread owns a 1x4 strip of output pixels, keeps the C x 3 x 6 input window in registers
//             and sweeps the filters 16 at a time (weights from shared memory as broadcast float4).
//   wgrad   : a CTA owns FT filters and a slice of the pixel strips; each thread keeps the window in
//             registers, reads the 4 dY values of each of its filters once and accumulates the
//             C*9 (+1 for gb) partial sums privately; one shuffle/shared reduction + atomics at the end.
//   dgrad   : a thread owns a 4x4 block of input pixels for all C channels and walks the filters,
//             loading the 6x6 dY window once per filter (2.25 loads per pixel-filter instead of 9).
//
// Literal semantics as in conv.cu: bias enters as K*b (Q1), gW/gb are (1/B) * full sums.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {

struct ThinArgs {
    const float* x; const float* w; const float* b; float* y;      // forward
    const float* dy; float* dx; float* gw; float* gb;              // backward
    int B, H, W, F, OH, OW, pt, pl;
    float inv_m;
    long strips_per_cta;
};

// VEC2: x0 is even, W is even and the tensor base is 8-byte aligned, so each window row is three 8-byte loads
template <int C, bool VEC2>
__device__ __forceinline__ void load_window(float (&win)[C][3][6], const float* __restrict__ xb, int H, int W, int y0, int x0) {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int iy = y0 + i;
            const bool rok = iy >= 0 && iy < H;
            const float* r = xb + ((long)c * H + iy) * W;
            if (VEC2) {
#pragma unroll
                for (int j = 0; j < 6; j += 2) {
                    const int ix = x0 + j;
                    float2 v = make_float2(0.f, 0.f);
                    if (rok && ix >= 0 && ix + 1 < W) v = __ldg(reinterpret_cast<const float2*>(r + ix));
                    else if (rok && ix >= 0 && ix < W) v.x = __ldg(r + ix);
                    win[c][i][j] = v.x; win[c][i][j + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int ix = x0 + j;
                    win[c][i][j] = (rok && ix >= 0 && ix < W) ? __ldg(r + ix) : 0.f;
                }
            }
        }
}

// ---- forward ---------------------------------------------------------------------------------------
template <int C, bool VEC2>
__global__ void __launch_bounds__(256) thin_fwd_kernel(ThinArgs a) {
    constexpr int K = C * 9;
    extern __shared__ __align__(16) float sm[];
    const int FP = (a.F + 15) & ~15;
    float* sw = sm;                  // [K][FP]
    float* sb = sm + K * FP;         // [FP]  K*b
    for (int e = threadIdx.x; e < K * FP; e += 256) {
        int k = e / FP, f = e % FP;
        sw[e] = f < a.F ? a.w[(long)f * K + k] : 0.f;
    }
    for (int f = threadIdx.x; f < FP; f += 256) sb[f] = f < a.F ? (float)K * a.b[f] : 0.f;
    __syncthreads();
    const int spr = (a.OW + 3) >> 2;
    const long total = (long)a.B * a.OH * spr;
    const long strip = (long)blockIdx.x * 256 + threadIdx.x;
    if (strip >= total) return;
    const int sx = (int)(strip % spr);
    long t = strip / spr;
    const int oy = (int)(t % a.OH);
    const int b = (int)(t / a.OH);
    const int xs = sx * 4;
    float win[C][3][6];
    load_window<C, VEC2>(win, a.x + (long)b * C * a.H * a.W, a.H, a.W, oy - a.pt, xs - a.pl);
    const long plane = (long)a.OH * a.OW;
    float* yb = a.y + (long)b * a.F * plane + (long)oy * a.OW + xs;
    for (int f0 = 0; f0 < a.F; f0 += 16) {
        float acc[4][16];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int f = 0; f < 16; ++f) acc[p][f] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float* wr = sw + ((c * 3 + i) * 3 + j) * FP + f0;
#pragma unroll
                    for (int f4 = 0; f4 < 4; ++f4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * f4);
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float xv = win[c][i][j + p];
                            acc[p][4 * f4 + 0] = fmaf(xv, w4.x, acc[p][4 * f4 + 0]);
                            acc[p][4 * f4 + 1] = fmaf(xv, w4.y, acc[p][4 * f4 + 1]);
                            acc[p][4 * f4 + 2] = fmaf(xv, w4.z, acc[p][4 * f4 + 2]);
                            acc[p][4 * f4 + 3] = fmaf(xv, w4.w, acc[p][4 * f4 + 3]);
                        }
                    }
                }
#pragma unroll
        for (int f = 0; f < 16; ++f) {
            if (f0 + f < a.F) {
                const float bb = sb[f0 + f];
                float* yr = yb + (long)(f0 + f) * plane;
                if (VEC2) {          // OW even, xs % 4 == 0: pairs are either fully inside or fully outside the row
                    if (xs + 1 < a.OW) *reinterpret_cast<float2*>(yr) = make_float2(acc[0][f] + bb, acc[1][f] + bb);
                    if (xs + 3 < a.OW) *reinterpret_cast<float2*>(yr + 2) = make_float2(acc[2][f] + bb, acc[3][f] + bb);
                } else {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if (xs + p < a.OW) yr[p] = acc[p][f] + bb;
                }
            }
        }
    }
}

// ---- weight gradient ------------------------------------------------------------------------------
template <int C, int FT, bool VEC2>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(ThinArgs a) {
    constexpr int K = C * 9, NA = FT * (K + 1);
    __shared__ float red[8][NA];
    const int f0 = blockIdx.y * FT;
    const int spr = (a.OW + 3) >> 2;
    const long total = (long)a.B * a.OH * spr;
    const long s_begin = (long)blockIdx.x * a.strips_per_cta;
    const long s_end = min(total, s_begin + a.strips_per_cta);
    const long plane = (long)a.OH * a.OW;
    float acc[FT][K + 1];
#pragma unroll
    for (int f = 0; f < FT; ++f)
#pragma unroll
        for (int k = 0; k <= K; ++k) acc[f][k] = 0.f;

    for (long strip = s_begin + threadIdx.x; strip < s_end; strip += 256) {
        const int sx = (int)(strip % spr);
        long t = strip / spr;
        const int oy = (int)(t % a.OH);
        const int b = (int)(t / a.OH);
        const int xs = sx * 4;
        float win[C][3][6];
        load_window<C, VEC2>(win, a.x + (long)b * C * a.H * a.W, a.H, a.W, oy - a.pt, xs - a.pl);
        const float* db = a.dy + ((long)b * a.F + f0) * plane + (long)oy * a.OW + xs;
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            float d[4];
            if (VEC2) {
                float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
                if (f0 + f < a.F) {
                    if (xs + 1 < a.OW) lo = __ldg(reinterpret_cast<const float2*>(db + (long)f * plane));
                    if (xs + 3 < a.OW) hi = __ldg(reinterpret_cast<const float2*>(db + (long)f * plane + 2));
                }
                d[0] = lo.x; d[1] = lo.y; d[2] = hi.x; d[3] = hi.y;
            } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) d[p] = (f0 + f < a.F && xs + p < a.OW) ? __ldg(db + (long)f * plane + p) : 0.f;
            }
            acc[f][K] += (d[0] + d[1]) + (d[2] + d[3]);
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        float s = acc[f][(c * 3 + i) * 3 + j];
#pragma unroll
                        for (int p = 0; p < 4; ++p) s = fmaf(d[p], win[c][i][j + p], s);
                        acc[f][(c * 3 + i) * 3 + j] = s;
                    }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int f = 0; f < FT; ++f)
#pragma unroll
        for (int k = 0; k <= K; ++k) {
            float v = warp_sum(acc[f][k]);
            if (lane == 0) red[wid][f * (K + 1) + k] = v;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < NA; e += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][e];
        const int f = f0 + e / (K + 1), k = e % (K + 1);
        if (f < a.F) {
            if (k < K) atomicAdd(a.gw + (long)f * K + k, s * a.inv_m);
            else atomicAdd(a.gb + f, s * a.inv_m);
        }
    }
}

// ---- data gradient --------------------------------------------------------------------------------
// dx[c,iy,ix] = sum_f sum_{i,j} dY[f, iy+pt-i, ix+pl-j] * W[f,c,i,j]
template <int C, bool VEC2>
__global__ void __launch_bounds__(256) thin_dgrad_kernel(ThinArgs a) {
    constexpr int K = C * 9;
    extern __shared__ __align__(16) float sm[];       // [F][K] weights
    for (int e = threadIdx.x; e < a.F * K; e += 256) sm[e] = a.w[e];
    __syncthreads();
    const int bw = (a.W + 3) >> 2, bh = (a.H + 3) >> 2;
    const long total = (long)a.B * bh * bw;
    const long blk = (long)blockIdx.x * 256 + threadIdx.x;
    if (blk >= total) return;
    const int bx = (int)(blk % bw);
    long t = blk / bw;
    const int by = (int)(t % bh);
    const int b = (int)(t / bh);
    const int iy0 = by * 4, ix0 = bx * 4;
    const int r0 = iy0 + a.pt - 2, c0 = ix0 + a.pl - 2;      // top-left of the 6x6 dY window
    const long plane = (long)a.OH * a.OW;
    const float* db = a.dy + (long)b * a.F * plane;
    float acc[C][4][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[c][y][x] = 0.f;
    // per-thread validity of the 6 window rows / columns (independent of f)
    bool rok[6], cok[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) { rok[r] = (r0 + r) >= 0 && (r0 + r) < a.OH; cok[r] = (c0 + r) >= 0 && (c0 + r) < a.OW; }
    for (int f = 0; f < a.F; ++f) {
        float dw[6][6];
        const float* df = db + (long)f * plane;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const float* dr = df + (long)(r0 + r) * a.OW + c0;
            if (VEC2) {          // c0 even, OW even: a pair is inside or outside as a whole
#pragma unroll
                for (int s = 0; s < 6; s += 2) {
                    float2 v = make_float2(0.f, 0.f);
                    if (rok[r] && cok[s]) v = __ldg(reinterpret_cast<const float2*>(dr + s));
                    dw[r][s] = v.x; dw[r][s + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int s = 0; s < 6; ++s) dw[r][s] = (rok[r] && cok[s]) ? __ldg(dr + s) : 0.f;
            }
        }
        const float* wf = sm + f * K;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float wv = wf[(c * 3 + i) * 3 + j];
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int x = 0; x < 4; ++x) acc[c][y][x] = fmaf(dw[y + 2 - i][x + 2 - j], wv, acc[c][y][x]);
                }
    }
    float* xb = a.dx + (long)b * C * a.H * a.W;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int y = 0; y < 4; ++y)
            if (iy0 + y < a.H) {
                float* xr = xb + ((long)c * a.H + iy0 + y) * a.W + ix0;
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (ix0 + x < a.W) xr[x] = acc[c][y][x];
            }
}


// =====================================================================================================
// "Flat" variants for C <= 2 when every F-channel plane is a multiple of 4 floats (e.g. 26x26 = 676): the
// F-channel tensor is then a dense array of 16-byte quads and is streamed with 128-bit accesses that never
// care about the (8-byte aligned, 104-byte) image rows.  A thread owns one quad q of one image = 4 flat
// output positions (which may straddle an image row) and keeps their C*9 taps in registers.
//
//   forward : taps from L1 (36 scalar loads per thread, reused for all F filters), one STG.128 per filter.
//   dgrad   : T[tap][q] = sum_f W[f][tap] * dY[f][q] per pixel (dY read ONCE with LDG.128, 9*F FMAs per pixel),
//             staged in shared memory, then dx[p] = sum_taps T[tap][p - tap]: 9 shared loads per dx pixel.
//   wgrad   : dY planes arrive through a cp.async.bulk (TMA) ring, (image, 8-filter chunk) per stage, a
//             producer thread keeps the ring full; consumers read dY with LDS.128 and accumulate
//             privately; one block reduction + atomics per CTA at the end.
// =====================================================================================================
template <int C>
__device__ __forceinline__ void flat_taps(float (&xin)[4][C * 9], const float* __restrict__ xb, int H, int W, int OW,
                                          int pt, int pl, int p0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int p = p0 + e;
        const int oy = p / OW, ox = p - oy * OW;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int iy = oy - pt + i;
                const bool rok = iy >= 0 && iy < H;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int ix = ox - pl + j;
                    xin[e][(c * 3 + i) * 3 + j] = (rok && ix >= 0 && ix < W) ? __ldg(xb + ((long)c * H + iy) * W + ix) : 0.f;
                }
            }
    }
}

template <int C>
__global__ void __launch_bounds__(256) thin_fwd_flat_kernel(ThinArgs a) {
    constexpr int K = C * 9, K4 = (K + 3) & ~3;
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                   // [F][K4]
    float* sb = sm + a.F * K4;        // [F]  K*b
    for (int e = threadIdx.x; e < a.F * K4; e += 256) {
        const int f = e / K4, k = e - f * K4;
        sw[e] = k < K ? a.w[(long)f * K + k] : 0.f;
    }
    for (int f = threadIdx.x; f < a.F; f += 256) sb[f] = (float)K * a.b[f];
    __syncthreads();
    const int plane = a.OH * a.OW, qpi = plane >> 2;
    const long item = (long)blockIdx.x * 256 + threadIdx.x;
    if (item >= (long)a.B * qpi) return;
    const int b = (int)(item / qpi), q = (int)(item - (long)b * qpi);
    float xin[4][K];
    flat_taps<C>(xin, a.x + (long)b * C * a.H * a.W, a.H, a.W, a.OW, a.pt, a.pl, 4 * q);
    float4* yq = reinterpret_cast<float4*>(a.y + (long)b * a.F * plane) + q;
#pragma unroll 4
    for (int f = 0; f < a.F; ++f) {
        const float bb = sb[f];
        float acc[4] = {bb, bb, bb, bb};
        const float4* wr = reinterpret_cast<const float4*>(sw + f * K4);
#pragma unroll
        for (int k4 = 0; k4 < K4 / 4; ++k4) {
            const float4 w4 = wr[k4];
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * k4 + u < K) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] = fmaf(xin[e][4 * k4 + u], wv[u], acc[e]);
                }
        }
        yq[(long)f * qpi] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

constexpr int FLAT_DG_THREADS = 192;
template <int C>
__global__ void __launch_bounds__(FLAT_DG_THREADS) thin_dgrad_flat_kernel(ThinArgs a) {
    constexpr int K = C * 9, K4 = (K + 3) & ~3;
    extern __shared__ __align__(16) float sm[];
    const int plane = a.OH * a.OW, qpi = plane >> 2;
    float* sT = sm;                   // [K][plane]
    float* sw = sm + K * plane;       // [F][K4]
    for (int e = threadIdx.x; e < a.F * K4; e += FLAT_DG_THREADS) {
        const int f = e / K4, k = e - f * K4;
        sw[e] = k < K ? a.w[(long)f * K + k] : 0.f;
    }
    __syncthreads();
    const int hw4 = (a.H * a.W) >> 2, w4n = a.W >> 2;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        // ---- phase 1: per-pixel contraction over the filters
        const float4* dq = reinterpret_cast<const float4*>(a.dy + (long)b * a.F * plane);
        for (int q = threadIdx.x; q < qpi; q += FLAT_DG_THREADS) {
            float T[K][4];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e) T[k][e] = 0.f;
            for (int f0 = 0; f0 < a.F; f0 += 8) {
                float4 d[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) d[u] = (f0 + u < a.F) ? __ldg(dq + (long)(f0 + u) * qpi + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (f0 + u < a.F) {
                        const float4* wr = reinterpret_cast<const float4*>(sw + (f0 + u) * K4);
#pragma unroll
                        for (int k4 = 0; k4 < K4 / 4; ++k4) {
                            const float4 w4 = wr[k4];
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                if (4 * k4 + v < K) {
                                    T[4 * k4 + v][0] = fmaf(wv[v], d[u].x, T[4 * k4 + v][0]);
                                    T[4 * k4 + v][1] = fmaf(wv[v], d[u].y, T[4 * k4 + v][1]);
                                    T[4 * k4 + v][2] = fmaf(wv[v], d[u].z, T[4 * k4 + v][2]);
                                    T[4 * k4 + v][3] = fmaf(wv[v], d[u].w, T[4 * k4 + v][3]);
                                }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
                *reinterpret_cast<float4*>(sT + k * plane + 4 * q) = make_float4(T[k][0], T[k][1], T[k][2], T[k][3]);
        }
        __syncthreads();
        // ---- phase 2: dx[c, iy, ix] = sum_{i,j} T[c,i,j][iy + pt - i, ix + pl - j]
        float4* xq = reinterpret_cast<float4*>(a.dx + (long)b * C * a.H * a.W);
        for (int t = threadIdx.x; t < C * hw4; t += FLAT_DG_THREADS) {
            const int c = t / hw4, r = t - c * hw4;
            const int iy = r / w4n, ix0 = (r - iy * w4n) * 4;
            float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int oy = iy + a.pt - i;
                if (oy < 0 || oy >= a.OH) continue;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float* tr = sT + ((c * 3 + i) * 3 + j) * plane + oy * a.OW;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ox = ix0 + e + a.pl - j;
                        if (ox >= 0 && ox < a.OW) o[e] += tr[ox];
                    }
                }
            }
            xq[t] = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
    }
}

// ---- wgrad: bulk-copy (TMA) ring ---------------------------------------------------------------------
constexpr int FLAT_WG_CONSUMERS = 192, FLAT_WG_THREADS = FLAT_WG_CONSUMERS + 32, FLAT_WG_STAGES = 4;
__host__ __device__ constexpr int flat_wg_ft(int C) { return C == 1 ? 8 : 4; }      // filters per CTA: FT*(9C+1) private accumulators per thread

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// ONEQ: plane/4 <= consumers, so a thread owns ONE quad for the whole kernel and its window geometry (4 shared
// offsets + validity bits) is computed once, outside the image loop; NOPAD additionally drops the validity tests.
template <int C, bool ONEQ, bool NOPAD>
__global__ void __launch_bounds__(FLAT_WG_THREADS, 2) thin_wgrad_flat_kernel(ThinArgs a) {
    constexpr int K = C * 9, FT = flat_wg_ft(C), ST = FLAT_WG_STAGES;
    extern __shared__ __align__(128) uint8_t smraw[];
    const int plane = a.OH * a.OW, qpi = plane >> 2, xsz = C * a.H * a.W;
    const int f0 = blockIdx.y * FT, nf = min(FT, a.F - f0);
    const uint32_t x_bytes = (uint32_t)xsz * 4, d_bytes = (uint32_t)nf * plane * 4;
    const uint32_t stage_bytes = ((x_bytes + (uint32_t)FT * plane * 4) + 127u) & ~127u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + ST;
    uint8_t* stages = smraw + 128;
    __shared__ float red[FLAT_WG_CONSUMERS / 32][FT * (K + 1)];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], FLAT_WG_CONSUMERS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int nimg = (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // images b = blockIdx.x + n*gridDim.x

    if (wid == FLAT_WG_CONSUMERS / 32) {
        // ===== producer: one elected lane keeps ST stages of (x image, dY chunk) in flight =====
        if (lane == 0) {
            for (int n = 0; n < nimg; ++n) {
                const int s = n % ST;
                if (n >= ST) tc::mbar_wait(&empty[s], ((n / ST) - 1) & 1);
                const long b = (long)blockIdx.x + (long)n * gridDim.x;
                uint8_t* st = stages + (size_t)s * stage_bytes;
                tc::mbar_arrive_expect_tx(&full[s], x_bytes + d_bytes);
                bulk_g2s(st, a.x + b * xsz, x_bytes, &full[s]);
                bulk_g2s(st + x_bytes, a.dy + (b * a.F + f0) * plane, d_bytes, &full[s]);
            }
        }
        return;
    }
    // ===== consumers =====
    // accumulators as float2 pairs (k, k+1): the hot loop issues packed FFMA2 (__ffma2_rn, sm_100: two IEEE fp32 FMAs per
    // instruction, bit-identical to scalar FMAs) — this kernel is bound by FP32 issue, and FFMA2 sustains ~1.55x the scalar
    // FMA rate on B200 (scratch/ffma2.cu).  Slot K is the bias gradient: its "window value" is the constant 1.
    constexpr int P2 = (K + 2) / 2;
    float2 acc2[FT][P2];
#define WG_ACC(f, k) (((k) & 1) ? acc2[f][(k) >> 1].y : acc2[f][(k) >> 1].x)
#pragma unroll
    for (int f = 0; f < FT; ++f)
#pragma unroll
        for (int p2 = 0; p2 < P2; ++p2) acc2[f][p2] = make_float2(0.f, 0.f);
    // ONEQ geometry of this thread's quad (image independent)
    int xo[4] = {0, 0, 0, 0};
    uint32_t vm[4] = {0u, 0u, 0u, 0u};
    if (ONEQ) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int p = min(4 * tid + e, plane - 1);
            const int oy = p / a.OW, ox = p - oy * a.OW;
            xo[e] = (oy - a.pt) * a.W + (ox - a.pl);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int iy = oy - a.pt + i, ix = ox - a.pl + j;
                    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) vm[e] |= 1u << (i * 3 + j);
                }
        }
    }
    const int hw = a.H * a.W;
    for (int n = 0; n < nimg; ++n) {
        const int s = n % ST;
        tc::mbar_wait(&full[s], (n / ST) & 1);
        const float* sx = reinterpret_cast<const float*>(stages + (size_t)s * stage_bytes);
        const float* sd = sx + xsz;
        if (ONEQ) {
            // two pixel pairs per quad: 2 x K window values live at a time (acc + window fit the register budget)
            if (tid < qpi) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float2 xin[2][P2];
#define WG_XIN(e, k) (((k) & 1) ? xin[e][(k) >> 1].y : xin[e][(k) >> 1].x)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        WG_XIN(e, K) = 1.f;
                        if (2 * P2 > K + 1) WG_XIN(e, 2 * P2 - 1) = 0.f;
                        const float* xe = sx + xo[2 * h + e];
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int i = 0; i < 3; ++i)
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    if (NOPAD) WG_XIN(e, (c * 3 + i) * 3 + j) = xe[c * hw + i * a.W + j];
                                    else WG_XIN(e, (c * 3 + i) * 3 + j) = ((vm[2 * h + e] >> (i * 3 + j)) & 1u) ? xe[c * hw + i * a.W + j] : 0.f;
                                }
                    }
#pragma unroll
                    for (int f = 0; f < FT; ++f) {
                        if (f < nf) {
                            const float2 d = *reinterpret_cast<const float2*>(sd + f * plane + 4 * tid + 2 * h);
                            const float2 d0 = make_float2(d.x, d.x), d1 = make_float2(d.y, d.y);
#pragma unroll
                            for (int p2 = 0; p2 < P2; ++p2) acc2[f][p2] = __ffma2_rn(d1, xin[1][p2], __ffma2_rn(d0, xin[0][p2], acc2[f][p2]));
                        }
                    }
                }
            }
        } else {
        for (int q = tid; q < qpi; q += FLAT_WG_CONSUMERS) {
            float xin[4][K];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = 4 * q + e;
                const int oy = p / a.OW, ox = p - oy * a.OW;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int iy = oy - a.pt + i;
                        const bool rok = iy >= 0 && iy < a.H;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int ix = ox - a.pl + j;
                            xin[e][(c * 3 + i) * 3 + j] = (rok && ix >= 0 && ix < a.W) ? sx[(c * a.H + iy) * a.W + ix] : 0.f;
                        }
                    }
            }
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                if (f < nf) {
                    const float4 d = *reinterpret_cast<const float4*>(sd + f * plane + 4 * q);
                    WG_ACC(f, K) += (d.x + d.y) + (d.z + d.w);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        float t = WG_ACC(f, k);
                        t = fmaf(d.x, xin[0][k], t); t = fmaf(d.y, xin[1][k], t);
                        t = fmaf(d.z, xin[2][k], t); t = fmaf(d.w, xin[3][k], t);
                        WG_ACC(f, k) = t;
                    }
                }
            }
        }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);
    }
    // ===== block reduction + atomics =====
#pragma unroll
    for (int f = 0; f < FT; ++f)
#pragma unroll
        for (int k = 0; k <= K; ++k) {
            const float v = warp_sum(WG_ACC(f, k));
            if (lane == 0) red[wid][f * (K + 1) + k] = v;
        }
    // consumers only: named barrier 1 (the producer warp has returned)
    asm volatile("bar.sync 1, %0;" ::"n"(FLAT_WG_CONSUMERS) : "memory");
    for (int e = tid; e < FT * (K + 1); e += FLAT_WG_CONSUMERS) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < FLAT_WG_CONSUMERS / 32; ++w) sum += red[w][e];
        const int f = f0 + e / (K + 1), k = e % (K + 1);
        if (f < a.F) {
            if (k < K) atomicAdd(a.gw + (long)f * K + k, sum * a.inv_m);
            else if (a.gb) atomicAdd(a.gb + f, sum * a.inv_m);
        }
    }
}

// ---- dgrad with the same bulk-copy ring (plane/4 <= 192: one quad per thread) ----------------------------
// stage = (image, 8-filter chunk of dY); T accumulates over the chunks of an image in registers.
constexpr int FLAT_DG_STAGES = 3, FLAT_DG_FT = 8;
template <int C>
__global__ void __launch_bounds__(FLAT_WG_THREADS, 2) thin_dgrad_ring_kernel(ThinArgs a) {
    constexpr int K = C * 9, K4 = (K + 3) & ~3, FT = FLAT_DG_FT, ST = FLAT_DG_STAGES, NC = FLAT_WG_CONSUMERS;
    extern __shared__ __align__(128) uint8_t smraw[];
    const int plane = a.OH * a.OW, qpi = plane >> 2;
    const int G = (a.F + FT - 1) / FT;
    const uint32_t stage_bytes = (uint32_t)FT * plane * 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + ST;
    uint8_t* stages = smraw + 128;
    float* sT = reinterpret_cast<float*>(stages + (size_t)ST * stage_bytes);     // [K][plane]
    float* sw = sT + K * plane;                                                  // [F][K4]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int e = tid; e < a.F * K4; e += FLAT_WG_THREADS) {
        const int f = e / K4, k = e - f * K4;
        sw[e] = k < K ? a.w[(long)f * K + k] : 0.f;
    }
    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], NC / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int nimg = (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (wid == NC / 32) {
        if (lane == 0) {
            int it = 0;
            for (int n = 0; n < nimg; ++n) {
                const long b = (long)blockIdx.x + (long)n * gridDim.x;
                for (int g = 0; g < G; ++g, ++it) {
                    const int s = it % ST;
                    if (it >= ST) tc::mbar_wait(&empty[s], ((it / ST) - 1) & 1);
                    const uint32_t bytes = (uint32_t)min(FT, a.F - g * FT) * plane * 4;
                    tc::mbar_arrive_expect_tx(&full[s], bytes);
                    bulk_g2s(stages + (size_t)s * stage_bytes, a.dy + (b * a.F + (long)g * FT) * plane, bytes, &full[s]);
                }
            }
        }
        return;
    }
    // phase-2 geometry of this thread's dx quads (image independent): up to 2 quads (C*H*W/4 <= 2*NC)
    const int hw4 = (a.H * a.W) >> 2, w4n = a.W >> 2, nxq = C * hw4;
    int x_base[2]; uint32_t x_rm[2], x_cm[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int t = min(tid + r * NC, nxq - 1);
        const int c = t / hw4, rr = t - c * hw4;
        const int iy = rr / w4n, ix0 = (rr - iy * w4n) * 4;
        x_base[r] = (c * 9) * plane + (iy + a.pt) * a.OW + ix0 + a.pl;       // tap (i,j), element e: + (i*3+j)*plane - i*OW - j + e
        uint32_t rm = 0, cm = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const int oy = iy + a.pt - i; if (oy >= 0 && oy < a.OH) rm |= 1u << i; }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { const int ox = ix0 + e + a.pl - j; if (ox >= 0 && ox < a.OW) cm |= 1u << (j * 4 + e); }
        x_rm[r] = rm; x_cm[r] = cm;
    }
    const bool active = tid < qpi;
    int it = 0;
    for (int n = 0; n < nimg; ++n) {
        const long b = (long)blockIdx.x + (long)n * gridDim.x;
        float T[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) T[k][e] = 0.f;
        for (int g = 0; g < G; ++g, ++it) {
            const int s = it % ST;
            tc::mbar_wait(&full[s], (it / ST) & 1);
            if (active) {
                const float4* sd = reinterpret_cast<const float4*>(stages + (size_t)s * stage_bytes) + tid;
                const int nf = min(FT, a.F - g * FT);
#pragma unroll
                for (int u = 0; u < FT; ++u) {
                    if (u < nf) {
                        const float4 d = sd[u * qpi];
                        const float4* wr = reinterpret_cast<const float4*>(sw + (g * FT + u) * K4);
#pragma unroll
                        for (int k4 = 0; k4 < K4 / 4; ++k4) {
                            const float4 w4 = wr[k4];
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                if (4 * k4 + v < K) {
                                    T[4 * k4 + v][0] = fmaf(wv[v], d.x, T[4 * k4 + v][0]);
                                    T[4 * k4 + v][1] = fmaf(wv[v], d.y, T[4 * k4 + v][1]);
                                    T[4 * k4 + v][2] = fmaf(wv[v], d.z, T[4 * k4 + v][2]);
                                    T[4 * k4 + v][3] = fmaf(wv[v], d.w, T[4 * k4 + v][3]);
                                }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&empty[s]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");        // phase 2 of the previous image has finished reading sT
        if (active) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                *reinterpret_cast<float4*>(sT + k * plane + 4 * tid) = make_float4(T[k][0], T[k][1], T[k][2], T[k][3]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
        float4* xq = reinterpret_cast<float4*>(a.dx + b * C * a.H * a.W);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = tid + r * NC;
            if (t < nxq) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float* tr = sT + x_base[r] + (i * 3 + j) * plane - i * a.OW - j;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (((x_rm[r] >> i) & 1u) && ((x_cm[r] >> (j * 4 + e)) & 1u)) o[e] += tr[e];
                    }
                xq[t] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool flat_ok(const ThinArgs& a, int C) {
    return C <= 2 && ((a.OH * a.OW) & 3) == 0 && !getenv("DNB_NO_THIN_FLAT");
}

template <int C>
static int thin_fwd_flat(const ThinArgs& a, cudaStream_t st) {
    constexpr int K4 = (C * 9 + 3) & ~3;
    const size_t smem = (size_t)a.F * (K4 + 1) * sizeof(float);
    const long total = (long)a.B * ((a.OH * a.OW) >> 2);
    if (smem > 48 * 1024) cudaFuncSetAttribute(thin_fwd_flat_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    thin_fwd_flat_kernel<C><<<(unsigned)((total + 255) / 256), 256, smem, st>>>(a);
    DNB_LAUNCH_CHECK("thin_conv2d_fwd");
    return DNB_OK;
}

template <int C>
static int thin_dgrad_flat(const ThinArgs& a, cudaStream_t st) {
    constexpr int K = C * 9, K4 = (K + 3) & ~3;
    const size_t smem = ((size_t)K * a.OH * a.OW + (size_t)a.F * K4) * sizeof(float);
    cudaFuncSetAttribute(thin_dgrad_flat_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, thin_dgrad_flat_kernel<C>, FLAT_DG_THREADS, smem);
    const int grid = std::min<long>(a.B, (long)kNumSMs * std::max(1, per_sm));
    thin_dgrad_flat_kernel<C><<<grid, FLAT_DG_THREADS, smem, st>>>(a);
    DNB_LAUNCH_CHECK("thin_conv2d_bwd_data");
    return DNB_OK;
}

template <int C>
static int thin_dgrad_ring(const ThinArgs& a, cudaStream_t st) {
    constexpr int K = C * 9, K4 = (K + 3) & ~3;
    const int plane = a.OH * a.OW;
    const size_t smem = 128 + (size_t)FLAT_DG_STAGES * FLAT_DG_FT * plane * 4 + ((size_t)K * plane + (size_t)a.F * K4) * 4;
    cudaFuncSetAttribute(thin_dgrad_ring_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(thin_dgrad_ring_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    const int grid = (int)std::min<long>(a.B, 2L * kNumSMs);
    thin_dgrad_ring_kernel<C><<<grid, FLAT_WG_THREADS, smem, st>>>(a);
    DNB_LAUNCH_CHECK("thin_conv2d_bwd_data");
    return DNB_OK;
}
static bool dgrad_ring_ok(const dnb_win_t* g) {
    const int plane = g->OH * g->OW, K = g->C * 9;
    const size_t smem = 128 + (size_t)FLAT_DG_STAGES * FLAT_DG_FT * plane * 4 + ((size_t)K * plane + (size_t)g->F * ((K + 3) & ~3)) * 4;
    return (plane >> 2) <= FLAT_WG_CONSUMERS && g->C * ((g->H * g->W) >> 2) <= 2 * FLAT_WG_CONSUMERS && smem <= 110 * 1024 &&
           !getenv("DNB_NO_THIN_RING");
}

template <int C>
static int thin_wgrad_flat(const ThinArgs& a, cudaStream_t st) {
    constexpr int FT = flat_wg_ft(C);
    const size_t stage = (((size_t)C * a.H * a.W + (size_t)FT * a.OH * a.OW) * 4 + 127) & ~(size_t)127;
    const size_t smem = 128 + FLAT_WG_STAGES * stage;
    const int fgroups = (a.F + FT - 1) / FT;
    const int splits = (int)std::max<long>(1, std::min<long>(a.B, (2L * kNumSMs + fgroups - 1) / fgroups));
    dim3 grid(splits, fgroups);
    const bool oneq = ((a.OH * a.OW) >> 2) <= FLAT_WG_CONSUMERS && !getenv("DNB_NO_THIN_ONEQ");
    const bool nopad = a.pt == 0 && a.pl == 0 && a.OH + 2 <= a.H && a.OW + 2 <= a.W;
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (!getenv("DNB_NO_CARVEOUT")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        kern<<<grid, FLAT_WG_THREADS, smem, st>>>(a);
    };
    if (!oneq) launch(thin_wgrad_flat_kernel<C, false, false>);
    else if (nopad) launch(thin_wgrad_flat_kernel<C, true, true>);
    else launch(thin_wgrad_flat_kernel<C, true, false>);
    DNB_LAUNCH_CHECK("thin_conv2d_bwd_weight");
    return DNB_OK;
}

bool thin_conv_supported(const dnb_win_t* g) {
    return g->C >= 1 && g->C <= 4 && g->kh == 3 && g->kw == 3 && g->sy == 1 && g->sx == 1 && g->F <= 1024;
}

static inline bool al8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

template <int C>
static int thin_fwd_c(const ThinArgs& a, cudaStream_t st) {
    const int FP = (a.F + 15) & ~15;
    const size_t smem = (size_t)(C * 9 + 1) * FP * sizeof(float);
    const long total = (long)a.B * a.OH * ((a.OW + 3) >> 2);
    const bool v2 = (a.W % 2 == 0) && (a.OW % 2 == 0) && (a.pl % 2 == 0) && al8(a.x) && al8(a.y);
    if (v2) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(thin_fwd_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        thin_fwd_kernel<C, true><<<(unsigned)((total + 255) / 256), 256, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(thin_fwd_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        thin_fwd_kernel<C, false><<<(unsigned)((total + 255) / 256), 256, smem, st>>>(a);
    }
    DNB_LAUNCH_CHECK("thin_conv2d_fwd");
    return DNB_OK;
}

template <int C, int FT>
static int thin_wgrad_c(ThinArgs a, cudaStream_t st) {
    const long total = (long)a.B * a.OH * ((a.OW + 3) >> 2);
    const int fgroups = (a.F + FT - 1) / FT;
    long splits = std::max<long>(1, std::min<long>((total + 1023) / 1024, (2L * kNumSMs + fgroups - 1) / fgroups));
    a.strips_per_cta = (total + splits - 1) / splits;
    splits = (total + a.strips_per_cta - 1) / a.strips_per_cta;
    dim3 grid((unsigned)splits, fgroups);
    const bool v2 = (a.W % 2 == 0) && (a.OW % 2 == 0) && (a.pl % 2 == 0) && al8(a.x) && al8(a.dy);
    if (v2) thin_wgrad_kernel<C, FT, true><<<grid, 256, 0, st>>>(a);
    else thin_wgrad_kernel<C, FT, false><<<grid, 256, 0, st>>>(a);
    DNB_LAUNCH_CHECK("thin_conv2d_bwd_weight");
    return DNB_OK;
}

template <int C>
static int thin_dgrad_c(const ThinArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.F * C * 9 * sizeof(float);
    const long total = (long)a.B * ((a.H + 3) >> 2) * ((a.W + 3) >> 2);
    // the 6x6 dY window starts at column ix0 + pl - 2 (ix0 % 4 == 0): even iff pl is even
    const bool v2 = (a.OW % 2 == 0) && (a.pl % 2 == 0) && al8(a.dy);
    if (v2) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(thin_dgrad_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        thin_dgrad_kernel<C, true><<<(unsigned)((total + 255) / 256), 256, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(thin_dgrad_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        thin_dgrad_kernel<C, false><<<(unsigned)((total + 255) / 256), 256, smem, st>>>(a);
    }
    DNB_LAUNCH_CHECK("thin_conv2d_bwd_data");
    return DNB_OK;
}

int thin_conv_fwd(const float* x, const float* w, const float* b, float* y, const dnb_win_t* g, cudaStream_t st) {
    ThinArgs a{x, w, b, y, nullptr, nullptr, nullptr, nullptr, g->B, g->H, g->W, g->F, g->OH, g->OW, g->pt, g->pl, 0.f, 0};
    if (flat_ok(a, g->C) && al16(a.y)) return g->C == 1 ? thin_fwd_flat<1>(a, st) : thin_fwd_flat<2>(a, st);
    switch (g->C) {
        case 1: return thin_fwd_c<1>(a, st);
        case 2: return thin_fwd_c<2>(a, st);
        case 3: return thin_fwd_c<3>(a, st);
        default: return thin_fwd_c<4>(a, st);
    }
}

int thin_conv_bwd(const float* x, const float* w, const float* dy, float* dx, float* gw, float* gb,
                  const dnb_win_t* g, float inv_m, bool skip_wgrad, cudaStream_t st) {
    ThinArgs a{x, w, nullptr, nullptr, dy, dx, gw, gb, g->B, g->H, g->W, g->F, g->OH, g->OW, g->pt, g->pl, inv_m, 0};
    int rc;
    const bool flat = flat_ok(a, g->C) && al16(a.dy);
    const size_t dg_smem = ((size_t)g->C * 9 * g->OH * g->OW + (size_t)g->F * 12 * g->C) * 4;
    if (dx && flat && (g->W & 3) == 0 && al16(dx) && dg_smem <= 200 * 1024) {
        if (dgrad_ring_ok(g)) rc = g->C == 1 ? thin_dgrad_ring<1>(a, st) : thin_dgrad_ring<2>(a, st);
        else rc = g->C == 1 ? thin_dgrad_flat<1>(a, st) : thin_dgrad_flat<2>(a, st);
        if (rc) return rc;
    } else if (dx) {
        switch (g->C) {
            case 1: rc = thin_dgrad_c<1>(a, st); break;
            case 2: rc = thin_dgrad_c<2>(a, st); break;
            case 3: rc = thin_dgrad_c<3>(a, st); break;
            default: rc = thin_dgrad_c<4>(a, st); break;
        }
        if (rc) return rc;
    }
    if (skip_wgrad) return DNB_OK;          // the tensor-core wgrad takes it (math == tf32)
    const size_t wg_stage = ((size_t)g->C * g->H * g->W + (size_t)flat_wg_ft(g->C) * g->OH * g->OW) * 4 + 128;
    if (flat && al16(a.x) && ((g->C * g->H * g->W) & 3) == 0 && 128 + FLAT_WG_STAGES * wg_stage <= 110 * 1024)
        return g->C == 1 ? thin_wgrad_flat<1>(a, st) : thin_wgrad_flat<2>(a, st);
    switch (g->C) {
        case 1: return thin_wgrad_c<1, 8>(a, st);
        case 2: return thin_wgrad_c<2, 4>(a, st);
        case 3: return thin_wgrad_c<3, 4>(a, st);
        default: return thin_wgrad_c<4, 2>(a, st);
    }
}

}  // namespace dnb
