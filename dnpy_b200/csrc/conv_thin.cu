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
#include <algorithm>

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
    if (dx) {
        switch (g->C) {
            case 1: rc = thin_dgrad_c<1>(a, st); break;
            case 2: rc = thin_dgrad_c<2>(a, st); break;
            case 3: rc = thin_dgrad_c<3>(a, st); break;
            default: rc = thin_dgrad_c<4>(a, st); break;
        }
        if (rc) return rc;
    }
    if (skip_wgrad) return DNB_OK;          // the tensor-core wgrad takes it (math == tf32)
    switch (g->C) {
        case 1: return thin_wgrad_c<1, 8>(a, st);
        case 2: return thin_wgrad_c<2, 4>(a, st);
        case 3: return thin_wgrad_c<3, 4>(a, st);
        default: return thin_wgrad_c<4, 2>(a, st);
    }
}

}  // namespace dnb
