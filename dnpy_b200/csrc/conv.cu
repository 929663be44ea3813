// Conv2D / PointwiseConv2D / DepthwiseConv2D — FP32 SIMT path
// (dnpy/layers/convolution.py:193-273, 276-284, 323-388), literal semantics:
//   * bias enters the output as K*b, K = window element count (Q1);
//   * gW / gb are per-position batch means summed over positions, i.e. (1/B) * full sums;
//   * the regulariser gradient is added once per output position (Q2).
// No im2col / col2im buffer is ever materialised: forward and data-gradient are the same
// register-tiled gather kernel (the data-gradient is a correlation of dY with the flipped,
// channel-transposed filter), the weight-gradient is a shared-memory micro-GEMM over gathered
// pixel stages.  The TF32 tensor-core implicit-GEMM path is in tc_conv.cu.
#include "common.cuh"
#include <cstdlib>
#include "tc_conv.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace dnb {

// conv_thin.cu: register-tiled kernels for 3x3 stride-1 convolutions with <= 4 input channels
bool thin_conv_supported(const dnb_win_t* g);
int thin_conv_fwd(const float* x, const float* w, const float* b, float* y, const dnb_win_t* g, cudaStream_t st);
int thin_conv_bwd(const float* x, const float* w, const float* dy, float* dx, float* gw, float* gb,
                  const dnb_win_t* g, float inv_m, bool skip_wgrad, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// Gather convolution: dst[b,d,y,x] = sum_{s,i,j} src[b,s,(y*mul+i-off_y)/div,(x*mul+j-off_x)/div] * w(d,s,i,j)
//   forward : src=x, dst=y, mul=stride, div=1, off=pad_top/left,  w(d,s,i,j)=W[d][s][i][j]
//   dgrad   : src=dy, dst=dx, mul=1, div=stride, off=k-1-pad,     w(d,s,i,j)=W[s][d][kh-1-i][kw-1-j]
// One thread = one output pixel x TF output channels (accumulators in registers); weights for a
// chunk of source channels are staged in shared memory and read as broadcast float4.
// ------------------------------------------------------------------------------------------
struct GatherArgs {
    const float* src; const float* w; const float* bias; float* dst;
    int B, Cs, Hs, Ws, Cd, Hd, Wd, kh, kw;
    int mul_y, mul_x, div_y, div_x, off_y, off_x;
    int cch;            // source channels per shared-memory weight stage
    float bias_k;       // forward: K (Q1); dgrad: unused
};

template <int TF, bool DGRAD>
__global__ void __launch_bounds__(128) conv_gather_kernel(GatherArgs a) {
    extern __shared__ __align__(16) float sw[];          // [cch*kh*kw][TF]
    const int tid = threadIdx.x;
    const long npix = (long)a.B * a.Hd * a.Wd;
    const long p = (long)blockIdx.x * 128 + tid;
    const bool valid = p < npix;
    const int d0 = blockIdx.y * TF;
    int b = 0, y = 0, x = 0;
    if (valid) { x = (int)(p % a.Wd); long t = p / a.Wd; y = (int)(t % a.Hd); b = (int)(t / a.Hd); }
    const int khw = a.kh * a.kw;
    float acc[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) acc[f] = 0.f;

    for (int c0 = 0; c0 < a.Cs; c0 += a.cch) {
        const int cn = min(a.cch, a.Cs - c0);
        const int nw = cn * khw * TF;
        for (int e = tid; e < nw; e += 128) {            // stage weights: k fastest in global
            int f = e / (cn * khw), k = e % (cn * khw);
            int cl = k / khw, t = k % khw, i = t / a.kw, j = t % a.kw;
            int d = d0 + f, s = c0 + cl;
            float v = 0.f;
            if (d < a.Cd) {
                if (!DGRAD) v = a.w[(((long)d * a.Cs + s) * a.kh + i) * a.kw + j];
                else        v = a.w[(((long)s * a.Cd + d) * a.kh + (a.kh - 1 - i)) * a.kw + (a.kw - 1 - j)];
            }
            sw[k * TF + f] = v;
        }
        __syncthreads();
        if (valid) {
            for (int cl = 0; cl < cn; ++cl) {
                const float* sp = a.src + ((long)b * a.Cs + (c0 + cl)) * a.Hs * a.Ws;
                for (int i = 0; i < a.kh; ++i) {
                    int ty = y * a.mul_y + i - a.off_y;
                    if (ty < 0) continue;
                    if (DGRAD && (ty % a.div_y)) continue;
                    int ys = DGRAD ? ty / a.div_y : ty;
                    if (ys >= a.Hs) continue;
                    for (int j = 0; j < a.kw; ++j) {
                        int tx = x * a.mul_x + j - a.off_x;
                        if (tx < 0) continue;
                        if (DGRAD && (tx % a.div_x)) continue;
                        int xs = DGRAD ? tx / a.div_x : tx;
                        if (xs >= a.Ws) continue;
                        float v = __ldg(sp + (long)ys * a.Ws + xs);
                        const float* wr = sw + ((cl * a.kh + i) * a.kw + j) * TF;
#pragma unroll
                        for (int f = 0; f < TF; f += 4) {
                            float4 w4 = *reinterpret_cast<const float4*>(wr + f);
                            acc[f] = fmaf(v, w4.x, acc[f]);
                            acc[f + 1] = fmaf(v, w4.y, acc[f + 1]);
                            acc[f + 2] = fmaf(v, w4.z, acc[f + 2]);
                            acc[f + 3] = fmaf(v, w4.w, acc[f + 3]);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (valid) {
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            int d = d0 + f;
            if (d < a.Cd) {
                float r = acc[f];
                if (!DGRAD && a.bias) r += a.bias_k * a.bias[d];
                a.dst[(((long)b * a.Cd + d) * a.Hd + y) * a.Wd + x] = r;
            }
        }
    }
}

template <bool DGRAD>
static int launch_gather(const char* name, GatherArgs a, cudaStream_t st) {
    const long npix = (long)a.B * a.Hd * a.Wd;
    if (npix == 0) return DNB_OK;
    const int khw = a.kh * a.kw;
    int tf = a.Cd <= 4 ? 4 : (a.Cd <= 8 ? 8 : (a.Cd <= 16 ? 16 : 32));
    a.cch = max(1, min(a.Cs, 8192 / (khw * tf)));          // <= 32 KB of staged weights
    size_t smem = (size_t)a.cch * khw * tf * sizeof(float);
    if (smem > 48 * 1024) return set_err(DNB_ERR_UNSUPPORTED, "%s: window %dx%d too large for the SIMT path", name, a.kh, a.kw);
    dim3 grid((unsigned)((npix + 127) / 128), (a.Cd + tf - 1) / tf);
    switch (tf) {
        case 4:  conv_gather_kernel<4, DGRAD><<<grid, 128, smem, st>>>(a); break;
        case 8:  conv_gather_kernel<8, DGRAD><<<grid, 128, smem, st>>>(a); break;
        case 16: conv_gather_kernel<16, DGRAD><<<grid, 128, smem, st>>>(a); break;
        default: conv_gather_kernel<32, DGRAD><<<grid, 128, smem, st>>>(a); break;
    }
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

// ------------------------------------------------------------------------------------------
// Weight gradient: gW[f,k] += inv_m * sum_p dY[p,f] * col[p,k],  k=(c,i,j), p=(b,y,x);
// gb[f] += inv_m * sum_p dY[p,f].  Shared-memory micro-GEMM: a CTA owns a 32(f) x TK(k) tile,
// split into GROUPS thread groups that walk disjoint pixel sub-stages; each thread keeps a 4x4
// register tile.  The im2col column tile exists only in shared memory, one stage at a time.
// ------------------------------------------------------------------------------------------
struct WgradArgs {
    const float* x; const float* dy; float* gw; float* gb;
    int B, C, H, W, F, OH, OW, kh, kw, sy, sx, pt, pl;
    int K;              // C*kh*kw
    float inv_m;
    long pix_per_cta;   // pixels walked by one CTA (multiple of the stage size)
};

constexpr int WG_TF = 32, WG_PIXG = 32;   // filters per CTA tile, pixels per group per stage

template <int TK>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(WgradArgs a) {
    constexpr int TPG = (WG_TF / 4) * (TK / 4);      // threads per group
    constexpr int GROUPS = 256 / TPG;
    constexpr int PIX = WG_PIXG * GROUPS;            // pixels per stage
    constexpr int DYP = WG_TF + 4, COLP = TK + 4;
    extern __shared__ __align__(16) float wg_smem[];
    float (*s_dy)[DYP] = reinterpret_cast<float (*)[DYP]>(wg_smem);                    // [PIX][DYP]
    float (*s_col)[COLP] = reinterpret_cast<float (*)[COLP]>(wg_smem + PIX * DYP);     // [PIX][COLP]
    float (*s_red)[TPG][17] = reinterpret_cast<float (*)[TPG][17]>(wg_smem);           // aliases the tiles (used after the loop)
    __shared__ int s_pb[PIX], s_py[PIX], s_px[PIX];  // per-pixel batch index / top-left input coords
    __shared__ int s_kc[TK], s_ki[TK], s_kj[TK];

    const int tid = threadIdx.x;
    const int f0 = blockIdx.y * WG_TF, k0 = blockIdx.z * TK;
    const int grp = tid / TPG, tig = tid % TPG;
    const int tf = tig / (TK / 4), tk = tig % (TK / 4);     // micro-tile coordinates
    const long npix = (long)a.B * a.OH * a.OW;
    const long p_begin = (long)blockIdx.x * a.pix_per_cta;
    const long p_end = min(npix, p_begin + a.pix_per_cta);
    const int khw = a.kh * a.kw;

    if (tid < TK) {
        int k = k0 + tid;
        if (k < a.K) { s_kc[tid] = k / khw; int t = k % khw; s_ki[tid] = t / a.kw; s_kj[tid] = t % a.kw; }
        else { s_kc[tid] = -1; s_ki[tid] = 0; s_kj[tid] = 0; }
    }
    float acc[4][4], accb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { accb[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f; }

    for (long ps = p_begin; ps < p_end; ps += PIX) {
        __syncthreads();
        if (tid < PIX) {
            long p = ps + tid;
            if (p < p_end) {
                int x = (int)(p % a.OW); long t = p / a.OW; int y = (int)(t % a.OH); int b = (int)(t / a.OH);
                s_pb[tid] = b; s_py[tid] = y * a.sy - a.pt; s_px[tid] = x * a.sx - a.pl;
            } else { s_pb[tid] = -1; s_py[tid] = 0; s_px[tid] = 0; }
        }
        __syncthreads();
        // dY tile: pixel fastest -> coalesced per filter plane
        for (int e = tid; e < PIX * WG_TF; e += 256) {
            int pl = e % PIX, fl = e / PIX;
            long p = ps + pl;
            int f = f0 + fl;
            float v = 0.f;
            if (p < p_end && f < a.F) {
                int b = s_pb[pl];
                long r = p - (long)b * a.OH * a.OW;
                v = __ldg(a.dy + ((long)b * a.F + f) * a.OH * a.OW + r);
            }
            s_dy[pl][fl] = v;
        }
        // im2col tile (shared memory only)
        for (int e = tid; e < PIX * TK; e += 256) {
            int pl = e % PIX, kl = e / PIX;
            float v = 0.f;
            int c = s_kc[kl], b = s_pb[pl];
            if (c >= 0 && b >= 0) {
                int iy = s_py[pl] + s_ki[kl], ix = s_px[pl] + s_kj[kl];
                if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                    v = __ldg(a.x + (((long)b * a.C + c) * a.H + iy) * a.W + ix);
            }
            s_col[pl][kl] = v;
        }
        __syncthreads();
        const int pbase = grp * WG_PIXG;
#pragma unroll 4
        for (int pp = 0; pp < WG_PIXG; ++pp) {
            float4 d4 = *reinterpret_cast<const float4*>(&s_dy[pbase + pp][tf * 4]);
            float4 c4 = *reinterpret_cast<const float4*>(&s_col[pbase + pp][tk * 4]);
            float dr[4] = {d4.x, d4.y, d4.z, d4.w}, cr[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (tk == 0) accb[i] += dr[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dr[i], cr[j], acc[i][j]);
            }
        }
    }
    // reduce the GROUPS partial tiles, then one atomic per output element per CTA
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s_red[grp][tig][i * 4 + j] = acc[i][j];
    __syncthreads();
    if (grp == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = 0.f;
                for (int g = 0; g < GROUPS; ++g) s += s_red[g][tig][i * 4 + j];
                int f = f0 + tf * 4 + i, k = k0 + tk * 4 + j;
                if (f < a.F && k < a.K) atomicAdd(a.gw + (long)f * a.K + k, s * a.inv_m);
            }
    }
    if (blockIdx.z == 0 && a.gb) {
        __syncthreads();
        if (tk == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s_red[grp][tf][i] = accb[i];
        }
        __syncthreads();
        if (grp == 0 && tk == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = 0.f;
                for (int g = 0; g < GROUPS; ++g) s += s_red[g][tf][i];
                int f = f0 + tf * 4 + i;
                if (f < a.F) atomicAdd(a.gb + f, s * a.inv_m);
            }
        }
    }
}

// regulariser terms, added OH*OW times (Q2)
__global__ void conv_reg_kernel(const float* __restrict__ w, float* __restrict__ gw, long nw,
                                const float* __restrict__ b, float* __restrict__ gb, int nb,
                                float cnt, float wl1, float wl2, float bl1, float bl2) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw && (wl1 != 0.f || wl2 != 0.f)) gw[i] += cnt * reg_term(w[i], wl1, wl2);
    if (i < nb && (bl1 != 0.f || bl2 != 0.f)) gb[i] += cnt * reg_term(b[i], bl1, bl2);
}

int launch_wgrad(WgradArgs a, cudaStream_t st) {
    const long npix = (long)a.B * a.OH * a.OW;
    if (npix == 0) return DNB_OK;
    int tk = a.K <= 16 ? 16 : (a.K <= 32 ? 32 : 64);
    int ftiles = (a.F + WG_TF - 1) / WG_TF, ktiles = (a.K + tk - 1) / tk;
    int tpg = (WG_TF / 4) * (tk / 4), stage = WG_PIXG * (256 / tpg);
    long max_split = (npix + stage - 1) / stage;
    long want = max(1L, (long)(kNumSMs * 4) / (ftiles * ktiles));
    long split = min(max_split, want);
    long per = ((npix + split - 1) / split + stage - 1) / stage * stage;
    a.pix_per_cta = per;
    dim3 grid((unsigned)((npix + per - 1) / per), ftiles, ktiles);
    size_t tiles = (size_t)stage * (WG_TF + 4 + tk + 4) * sizeof(float);
    size_t red = (size_t)(256 / tpg) * tpg * 17 * sizeof(float);
    size_t smem = tiles > red ? tiles : red;
    if (tk == 16) {
        cudaFuncSetAttribute(conv_wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_wgrad_kernel<16><<<grid, 256, smem, st>>>(a);
    } else if (tk == 32) {
        cudaFuncSetAttribute(conv_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_wgrad_kernel<32><<<grid, 256, smem, st>>>(a);
    } else {
        cudaFuncSetAttribute(conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_wgrad_kernel<64><<<grid, 256, smem, st>>>(a);
    }
    DNB_LAUNCH_CHECK("conv2d_bwd_weight");
    return DNB_OK;
}

// ------------------------------------------------------------------------------------------
// Depthwise (bandwidth-bound direct kernels; 18 FLOP per 8 B — not a tensor-core op)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, dnb_win_t g) {
    long n = (long)g.B * g.C * g.OH * g.OW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ox = (int)(i % g.OW); long t = i / g.OW; int oy = (int)(t % g.OH); t /= g.OH; int c = (int)(t % g.C); int b = (int)(t / g.C);
        const float* xp = x + ((long)b * g.C + c) * g.H * g.W;
        const float* wp = w + (long)c * g.kh * g.kw;
        float acc = 0.f;
        for (int ii = 0; ii < g.kh; ++ii) {
            int iy = oy * g.sy + ii - g.pt;
            if (iy < 0 || iy >= g.H) continue;
            for (int jj = 0; jj < g.kw; ++jj) {
                int ix = ox * g.sx + jj - g.pl;
                if (ix < 0 || ix >= g.W) continue;
                acc = fmaf(__ldg(xp + (long)iy * g.W + ix), wp[ii * g.kw + jj], acc);
            }
        }
        y[i] = acc + (float)(g.kh * g.kw) * bias[c];     // (ky*kx)*b, convolution.py:344-347
    }
}

__global__ void __launch_bounds__(256) dw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                       float* __restrict__ dx, dnb_win_t g) {
    long n = (long)g.B * g.C * g.H * g.W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ix = (int)(i % g.W); long t = i / g.W; int iy = (int)(t % g.H); t /= g.H; int c = (int)(t % g.C); int b = (int)(t / g.C);
        const float* dp = dy + ((long)b * g.C + c) * g.OH * g.OW;
        const float* wp = w + (long)c * g.kh * g.kw;
        float acc = 0.f;
        for (int ii = 0; ii < g.kh; ++ii) {
            int ty = iy + g.pt - ii;
            if (ty < 0 || (ty % g.sy)) continue;
            int oy = ty / g.sy;
            if (oy >= g.OH) continue;
            for (int jj = 0; jj < g.kw; ++jj) {
                int tx = ix + g.pl - jj;
                if (tx < 0 || (tx % g.sx)) continue;
                int ox = tx / g.sx;
                if (ox >= g.OW) continue;
                acc = fmaf(__ldg(dp + (long)oy * g.OW + ox), wp[ii * g.kw + jj], acc);
            }
        }
        dx[i] = acc;
    }
}

// gW[c,i,j] += inv_m * sum_{b,y,x} dy*xpad ; gb[c] += inv_m*sum dy.  grid = (splits, C)
template <int KP>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ gw, float* __restrict__ gb, dnb_win_t g, float inv_m) {
    __shared__ float red[32];
    const int c = blockIdx.y;
    const int khw = g.kh * g.kw;
    const long npix = (long)g.B * g.OH * g.OW;
    float acc[KP], accb = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.f;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long)gridDim.x * blockDim.x) {
        int ox = (int)(p % g.OW); long t = p / g.OW; int oy = (int)(t % g.OH); int b = (int)(t / g.OH);
        float d = __ldg(dy + (((long)b * g.C + c) * g.OH + oy) * g.OW + ox);
        accb += d;
        const float* xp = x + ((long)b * g.C + c) * g.H * g.W;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < khw) {
                int iy = oy * g.sy + k / g.kw - g.pt, ix = ox * g.sx + k % g.kw - g.pl;
                if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) acc[k] = fmaf(d, __ldg(xp + (long)iy * g.W + ix), acc[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (k < khw) {
            float s = block_sum<256>(acc[k], red);
            if (threadIdx.x == 0) atomicAdd(gw + (long)c * khw + k, s * inv_m);
        }
    }
    float sb = block_sum<256>(accb, red);
    if (threadIdx.x == 0) atomicAdd(gb + c, sb * inv_m);
}

// ------------------------------------------------------------------------------------------
// Depthwise 3x3, stride 1, pad 1 ("same"): ring kernels.  All global traffic is bulk-asynchronous: (b,c) planes arrive
// by 1-D bulk copies into a shared-memory ring (one producer lane keeps it full), consumers read 4 consecutive pixels of
// a row with one LDS.128 and take the two halo values from the neighbouring lanes with shuffles (a scalar halo load
// would be a 4-way bank conflict), finished planes leave by bulk stores.  The element-per-thread kernels above run at
// ~5% of HBM peak on B200 (9 L1 loads and 64-bit index arithmetic per output).
//   forward / data gradient: y = corr(x, W[c]) (+ 9*bias), dx = corr(dy, flip(W[c]))       (one kernel, `flip`)
//   weight gradient        : CTA (split, c) walks the planes (b, c) of ITS channel: 9 + 1 private accumulators per
//                            thread, one block reduction and 10 atomics per CTA at the end
// ------------------------------------------------------------------------------------------
constexpr int DWR_CONS = 256, DWR_THREADS = DWR_CONS + 32, DWR_ST = 3;

__device__ __forceinline__ void dwr_store_plane(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_u32(ssrc)), "r"(bytes) : "memory");
}
// the 3 x 6 input window of a pixel quad (rows oy-1..oy+1, columns ox0-1..ox0+4) of plane `in`; all lanes of the warp call
__device__ __forceinline__ void dwr_window(const float* in, int H, int W, int oy, int ox0, bool live, float (&v)[3][6]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int iy = oy + i - 1;
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && iy >= 0 && iy < H) m = *reinterpret_cast<const float4*>(in + iy * W + ox0);
        // neighbours along the row: lane-1 holds columns ox0-4..ox0-1, lane+1 holds ox0+4..ox0+7 (same row unless at its edge)
        const float l = __shfl_up_sync(0xffffffffu, m.w, 1), r = __shfl_down_sync(0xffffffffu, m.x, 1);
        v[i][0] = (ox0 > 0 && (threadIdx.x & 31) > 0) ? l : 0.f;
        v[i][1] = m.x; v[i][2] = m.y; v[i][3] = m.z; v[i][4] = m.w;
        v[i][5] = (ox0 + 4 < W && (threadIdx.x & 31) < 31) ? r : 0.f;
    }
}
// lanes 0 / 31 have no neighbour lane: fetch their halo value from shared memory directly
__device__ __forceinline__ void dwr_fix_edges(const float* in, int H, int W, int oy, int ox0, bool live, float (&v)[3][6]) {
    const int lane = threadIdx.x & 31;
    if (!live || (lane != 0 && lane != 31)) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int iy = oy + i - 1;
        if (iy < 0 || iy >= H) continue;
        if (lane == 0 && ox0 > 0) v[i][0] = in[iy * W + ox0 - 1];
        if (lane == 31 && ox0 + 4 < W) v[i][5] = in[iy * W + ox0 + 4];
    }
}

__global__ void __launch_bounds__(DWR_THREADS, 2) dw3_ring_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ y,
                                                                   int planes, int C, int H, int W, int flip) {
    extern __shared__ __align__(128) uint8_t dsm[];
    const int plane = H * W;
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm);
    uint64_t* empty = full + DWR_ST;
    float* sin = reinterpret_cast<float*>(dsm + 128);                   // [ST][plane]
    float* sout = sin + (size_t)DWR_ST * plane;                         // [2][plane]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < DWR_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], DWR_CONS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int mine = (planes - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (wid == DWR_CONS / 32) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)plane * 4u;
            for (int n = 0; n < mine; ++n) {
                const int s = n % DWR_ST;
                if (n >= DWR_ST) tc::mbar_wait(&empty[s], ((n / DWR_ST) - 1) & 1);
                const long p = (long)blockIdx.x + (long)n * gridDim.x;
                tc::mbar_arrive_expect_tx(&full[s], bytes);
                tc::bulk_load_1d(sin + (size_t)s * plane, x + p * plane, bytes, &full[s]);
            }
        }
        return;
    }
    const int qrow = W >> 2, nq = plane >> 2;
    for (int n = 0; n < mine; ++n) {
        const int s = n % DWR_ST, ob = n & 1;
        const long p = (long)blockIdx.x + (long)n * gridDim.x;
        const int c = (int)(p % C);
        float wk[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) wk[k] = __ldg(w + c * 9 + (flip ? 8 - k : k));
        const float b9 = bias ? 9.f * __ldg(bias + c) : 0.f;            // (ky*kx)*b, convolution.py:344-347
        tc::mbar_wait(&full[s], (n / DWR_ST) & 1);
        const float* in = sin + (size_t)s * plane;
        float* out = sout + (size_t)ob * plane;
        for (int q0 = wid * 32; q0 < nq; q0 += DWR_CONS) {
            const int q = q0 + lane;
            const bool live = q < nq;
            const int oy = q / qrow, ox0 = (q - oy * qrow) << 2;
            float v[3][6];
            dwr_window(in, H, W, oy, ox0, live, v);
            dwr_fix_edges(in, H, W, oy, ox0, live, v);
            if (live) {
                float o[4] = {b9, b9, b9, b9};
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = fmaf(v[i][e + j], wk[i * 3 + j], o[e]);
                *reinterpret_cast<float4*>(out + (q << 2)) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);
        tc::fence_proxy_async_smem();
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the other buffer's store has drained
        asm volatile("bar.sync 1, %0;" ::"n"(DWR_CONS) : "memory");
        if (tid == 0) {
            dwr_store_plane(y + p * plane, out, (uint32_t)plane * 4u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(DWR_THREADS, 2) dw3_wgrad_ring_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                         float* __restrict__ gw, float* __restrict__ gb,
                                                                         int B, int C, int H, int W, float inv_m) {
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ float red[32];
    const int plane = H * W;
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm);
    uint64_t* empty = full + DWR_ST;
    float* sin = reinterpret_cast<float*>(dsm + 128);                   // [ST][2 * plane]: x plane, dY plane
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c = blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < DWR_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], DWR_CONS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int mine = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // samples blockIdx.x + n*gridDim.x
    float acc[9], accb = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    if (wid == DWR_CONS / 32) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)plane * 4u;
            for (int n = 0; n < mine; ++n) {
                const int s = n % DWR_ST;
                if (n >= DWR_ST) tc::mbar_wait(&empty[s], ((n / DWR_ST) - 1) & 1);
                const long p = ((long)blockIdx.x + (long)n * gridDim.x) * C + c;
                tc::mbar_arrive_expect_tx(&full[s], 2 * bytes);
                tc::bulk_load_1d(sin + (size_t)s * 2 * plane, x + p * plane, bytes, &full[s]);
                tc::bulk_load_1d(sin + (size_t)s * 2 * plane + plane, dy + p * plane, bytes, &full[s]);
            }
        }
    } else {
        const int qrow = W >> 2, nq = plane >> 2;
        for (int n = 0; n < mine; ++n) {
            const int s = n % DWR_ST;
            tc::mbar_wait(&full[s], (n / DWR_ST) & 1);
            const float* in = sin + (size_t)s * 2 * plane;
            const float* sd = in + plane;
            for (int q0 = wid * 32; q0 < nq; q0 += DWR_CONS) {
                const int q = q0 + lane;
                const bool live = q < nq;
                const int oy = q / qrow, ox0 = (q - oy * qrow) << 2;
                float v[3][6];
                dwr_window(in, H, W, oy, ox0, live, v);
                dwr_fix_edges(in, H, W, oy, ox0, live, v);
                if (live) {
                    const float4 d4 = *reinterpret_cast<const float4*>(sd + (q << 2));
                    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
                    accb += (d[0] + d[1]) + (d[2] + d[3]);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[i * 3 + j] = fmaf(d[e], v[i][e + j], acc[i * 3 + j]);
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&empty[s]);
        }
    }
    // every thread (the producer warp with zeros) joins the block reductions
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float sum = block_sum<DWR_THREADS>(acc[k], red);
        if (tid == 0 && mine > 0) atomicAdd(gw + c * 9 + k, sum * inv_m);
    }
    const float sb = block_sum<DWR_THREADS>(accb, red);
    if (tid == 0 && mine > 0) atomicAdd(gb + c, sb * inv_m);
}

static bool dw3_ring_ok(const dnb_win_t* g) {
    return g->kh == 3 && g->kw == 3 && g->sy == 1 && g->sx == 1 && g->pt == 1 && g->pb == 1 && g->pl == 1 && g->pr == 1 &&
           g->OH == g->H && g->OW == g->W && g->W % 4 == 0 && g->W >= 8 && (long)g->H * g->W * 4 <= 32 * 1024 &&
           (long)g->B * g->C <= 0x7fffffffL && !getenv("DNB_NO_DW_RING");
}

static int check_win(const char* name, const dnb_win_t* g) {
    DNB_CHECK_ARG(g, "%s: null geometry", name);
    DNB_CHECK_ARG(g->B >= 0 && g->C > 0 && g->H > 0 && g->W > 0 && g->F > 0 && g->OH > 0 && g->OW > 0 &&
                  g->kh > 0 && g->kw > 0 && g->sy > 0 && g->sx > 0, "%s: non-positive dimension", name);
    DNB_CHECK_ARG(g->pt >= 0 && g->pb >= 0 && g->pl >= 0 && g->pr >= 0, "%s: negative padding", name);
    DNB_CHECK_ARG((g->OH - 1) * g->sy + g->kh <= g->H + g->pt + g->pb && (g->OW - 1) * g->sx + g->kw <= g->W + g->pl + g->pr,
                  "%s: windows exceed the padded input", name);
    return DNB_OK;
}


// ---- fp32 mode on the tensor core: 3 x TF32 -----------------------------------------------------------------------------
// math = FP32 promises rel 1e-5 against the float64 reference.  A TF32 operand keeps 11 significant bits, so every fp32 value
// is split as v = hi + lo with hi = v truncated to TF32 (exactly representable: the tensor core's own conversion cannot change
// it) and lo = v - hi (exact in fp32; its TF32 conversion loses <= 2^-11 of lo = 2^-22 of v).  With fp32 accumulation
//     a . b  =  a_hi . b_hi  +  a_lo . b_hi  +  a_hi . b_lo        (the dropped a_lo . b_lo term is 2^-22 relative)
// so an fp32-grade convolution is THREE launches of the parity-tested TF32 kernels on the split tensors (convolution.py:193-273
// arithmetic, 1e-6 measured), instead of the SIMT fp32 kernels (conv2 of the MNIST CNN at B = 8192: 2.5 / 3.7 / 3.7 ms).
// Workspace behind the TF32 kernels' own: hi / lo of the filters, of the gathered tensor and of dY, two partial outputs.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(in)[i];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        const float h = __uint_as_float(__float_as_uint(in[i]) & 0xffffe000u);
        hi[i] = h; lo[i] = in[i] - h;
    }
}
// y += t1 + t2 (the two cross terms are added to each other first: they are ~2^-11 of the main term)
__global__ void __launch_bounds__(256) add_cross_terms_kernel(float* __restrict__ y, const float* __restrict__ t1, const float* __restrict__ t2, size_t n) {
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(y)[i];
        const float4 a = reinterpret_cast<const float4*>(t1)[i], b = reinterpret_cast<const float4*>(t2)[i];
        v.x += a.x + b.x; v.y += a.y + b.y; v.z += a.z + b.z; v.w += a.w + b.w;
        reinterpret_cast<float4*>(y)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        y[i] += t1[i] + t2[i];
    }
}

struct SplitWs {               // layout of the fp32-mode workspace (all offsets multiples of 256 bytes)
    uint8_t* tc;               // the TF32 kernels' own workspace (prepared filters, dilated dY)
    float *wh, *wl, *gbd, *ah, *al, *bh, *bl, *t1, *t2;
    size_t bytes;
};
static SplitWs split_ws(const dnb_win_t* g, void* ws) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t nx = (size_t)g->B * g->C * g->H * g->W, ny = (size_t)g->B * g->F * g->OH * g->OW;
    const size_t nw = (size_t)g->F * g->C * g->kh * g->kw;
    SplitWs s{};
    uint8_t* p = static_cast<uint8_t*>(ws);
    size_t o = 0;
    s.tc = p; o += up(tc_conv_ws_bytes(g));
    s.wh = reinterpret_cast<float*>(p + o); o += up(nw * 4);
    s.wl = reinterpret_cast<float*>(p + o); o += up(nw * 4);
    s.gbd = reinterpret_cast<float*>(p + o); o += up((size_t)g->F * 4);
    s.ah = reinterpret_cast<float*>(p + o); o += up(nx * 4);
    s.al = reinterpret_cast<float*>(p + o); o += up(nx * 4);
    s.bh = reinterpret_cast<float*>(p + o); o += up(ny * 4);
    s.bl = reinterpret_cast<float*>(p + o); o += up(ny * 4);
    s.t1 = reinterpret_cast<float*>(p + o); o += up(std::max(nx, ny) * 4);
    s.t2 = reinterpret_cast<float*>(p + o); o += up(std::max(nx, ny) * 4);
    s.bytes = o;
    return s;
}
// eligible: a tensor-core shape with enough work to amortise the six extra passes (DNB_FP32_SIMT=1 keeps the SIMT kernels)
static bool split_ok(const dnb_win_t* g, int kind) {
    if (getenv("DNB_FP32_SIMT") || thin_conv_supported(g)) return false;
    const double macs = (double)g->B * g->F * g->OH * g->OW * g->C * g->kh * g->kw;
    return macs >= (double)(1 << 26) && tc_conv_supported(g, kind);
}
static int split_tensor(const char* name, const float* in, float* hi, float* lo, size_t n, cudaStream_t st) {
    split_tf32_kernel<<<ew_grid((n + 3) / 4, 256), 256, 0, st>>>(in, hi, lo, n);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}
static int add_cross(float* y, const float* t1, const float* t2, size_t n, cudaStream_t st) {
    add_cross_terms_kernel<<<ew_grid((n + 3) / 4, 256), 256, 0, st>>>(y, t1, t2, n);
    DNB_LAUNCH_CHECK("conv2d_add_cross_terms");
    return DNB_OK;
}
static int conv_fwd_split(const float* x, const float* w, const float* b, float* y, const dnb_win_t* g, void* ws, cudaStream_t st) {
    if (!ws) return set_err(DNB_ERR_INVALID, "conv2d_fwd(fp32 on the tensor core): workspace required");
    const SplitWs s = split_ws(g, ws);
    const size_t nx = (size_t)g->B * g->C * g->H * g->W, ny = (size_t)g->B * g->F * g->OH * g->OW;
    int rc = split_tensor("conv2d_split_w", w, s.wh, s.wl, (size_t)g->F * g->C * g->kh * g->kw, st);
    if (!rc) rc = split_tensor("conv2d_split_x", x, s.ah, s.al, nx, st);
    if (!rc) rc = tc_conv_fwd(s.ah, s.wh, b, y, g, s.tc, false, st);
    if (!rc) rc = tc_conv_fwd(s.al, s.wh, nullptr, s.t1, g, s.tc, false, st);
    if (!rc) rc = tc_conv_fwd(s.ah, s.wl, nullptr, s.t2, g, s.tc, false, st);
    if (!rc) rc = add_cross(y, s.t1, s.t2, ny, st);
    return rc;
}
// dy_split: dY is already split into s.bh / s.bl (dnb_conv2d_bwd splits it once for both gradients)
static int conv_dgrad_split(const float* w, const float* dy, float* dx, const dnb_win_t* g, void* ws, bool dy_split, cudaStream_t st) {
    if (!ws) return set_err(DNB_ERR_INVALID, "conv2d_bwd(fp32 on the tensor core): workspace required");
    const SplitWs s = split_ws(g, ws);
    const size_t nx = (size_t)g->B * g->C * g->H * g->W, ny = (size_t)g->B * g->F * g->OH * g->OW;
    int rc = split_tensor("conv2d_split_w", w, s.wh, s.wl, (size_t)g->F * g->C * g->kh * g->kw, st);
    if (!rc && !dy_split) rc = split_tensor("conv2d_split_dy", dy, s.bh, s.bl, ny, st);
    if (!rc) rc = tc_conv_dgrad(s.bh, s.wh, dx, g, s.tc, false, st);
    if (!rc) rc = tc_conv_dgrad(s.bl, s.wh, s.t1, g, s.tc, false, st);
    if (!rc) rc = tc_conv_dgrad(s.bh, s.wl, s.t2, g, s.tc, false, st);
    if (!rc) rc = add_cross(dx, s.t1, s.t2, nx, st);
    return rc;
}
static int conv_wgrad_split(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, void* ws,
                            bool dy_split, cudaStream_t st) {
    if (!ws) return set_err(DNB_ERR_INVALID, "conv2d_bwd(fp32 on the tensor core): workspace required");
    const SplitWs s = split_ws(g, ws);
    const size_t nx = (size_t)g->B * g->C * g->H * g->W, ny = (size_t)g->B * g->F * g->OH * g->OW;
    int rc = split_tensor("conv2d_split_x", x, s.ah, s.al, nx, st);
    if (!rc && !dy_split) rc = split_tensor("conv2d_split_dy", dy, s.bh, s.bl, ny, st);
    if (!rc) rc = tc_conv_wgrad(s.ah, s.bh, gw, gb, g, inv_m, s.tc, false, st);        // gb += sum dY_hi
    if (!rc) rc = tc_conv_wgrad(s.al, s.bh, gw, nullptr, g, inv_m, s.tc, false, st);
    if (!rc) rc = tc_conv_wgrad(s.ah, s.bl, gw, gb, g, inv_m, s.tc, false, st);        // gb += sum dY_lo
    return rc;
}

}  // namespace dnb

using namespace dnb;

extern "C" {

size_t dnb_conv2d_ws_bytes(const dnb_win_t* g, int math) {
    if (!g) return 0;
    if (math >= DNB_MATH_TF32) return tc_conv_ws_bytes(g);
    if (split_ok(g, TC_CONV_FWD) || split_ok(g, TC_CONV_DGRAD) || split_ok(g, TC_CONV_WGRAD)) return split_ws(g, nullptr).bytes;
    return 0;
}

int dnb_conv2d_fwd(const float* x, const float* w, const float* b, float* y,
                   const dnb_win_t* g, void* ws, int math, dnb_stream_t stream) {
    int rc = check_win("conv2d_fwd", g);
    if (rc) return rc;
    if (g->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && y, "conv2d_fwd: null pointer");
    if (thin_conv_supported(g)) return thin_conv_fwd(x, w, b, y, g, S(stream));
    if (math == DNB_MATH_FP32 && split_ok(g, TC_CONV_FWD)) return conv_fwd_split(x, w, b, y, g, ws, S(stream));
    if (math >= DNB_MATH_TF32 && tc_conv_supported(g, TC_CONV_FWD))
        return tc_conv_fwd(x, w, b, y, g, ws, math == DNB_MATH_BF16, S(stream));
    GatherArgs a{x, w, b, y, g->B, g->C, g->H, g->W, g->F, g->OH, g->OW, g->kh, g->kw,
                 g->sy, g->sx, 1, 1, g->pt, g->pl, 0, (float)(g->C * g->kh * g->kw)};
    return launch_gather<false>("conv2d_fwd", a, S(stream));
}

int dnb_conv2d_bwd(const float* x, const float* w, const float* b, const float* dy,
                   float* dx, float* gw, float* gb, const dnb_win_t* g, float inv_m,
                   dnb_reg_t reg_w, dnb_reg_t reg_b, float reg_scale,
                   void* ws, int math, dnb_stream_t stream) {
    int rc = check_win("conv2d_bwd", g);
    if (rc) return rc;
    if (g->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && dy && gw && gb, "conv2d_bwd: null pointer");
    cudaStream_t st = S(stream);
    const bool thin = thin_conv_supported(g);
    // measured on B200 (conv 1->32, 28x28, B=8192): the register-tiled fp32 wgrad (0.35 ms) edges out the tensor-core
    // thin-K variant (0.40 ms) — both are dY-stream bound — so thin layers keep the fp32 kernel in either math mode
    // (C >= 3 on large planes is the exception: the thin-K tensor-core variant wins there — measured on the 3 x 64 x 64 stem
    //  of the residual net, DNB_THIN_TC_WGRAD=0 restores the fp32 kernel)
    const bool thin_tc = thin && g->C >= 3 && (long)g->OH * g->OW >= 1024 && !(getenv("DNB_THIN_TC_WGRAD") && atoi(getenv("DNB_THIN_TC_WGRAD")) == 0);
    const bool tc_wgrad = (!thin || thin_tc) && math >= DNB_MATH_TF32 && tc_conv_supported(g, TC_CONV_WGRAD);
    // ... except the single-channel case, where dY can feed the tensor core straight from TMA (tc_conv_thin.cu)
    const bool tc_thin = thin && math >= DNB_MATH_TF32 && tc_thin_wgrad_supported(g) &&
                         !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15);
    const bool thin_tc_dgrad = thin_tc && dx && math >= DNB_MATH_TF32 && tc_conv_supported(g, TC_CONV_DGRAD);
    if (thin) {
        rc = thin_conv_bwd(x, w, dy, thin_tc_dgrad ? nullptr : dx, gw, gb, g, inv_m, /*skip_wgrad=*/tc_wgrad || tc_thin, st);
        if (rc) return rc;
        if (thin_tc_dgrad) {
            rc = tc_conv_dgrad(dy, w, dx, g, ws, math == DNB_MATH_BF16, st);
            if (rc) return rc;
        }
        if (tc_thin) {
            rc = tc_thin_wgrad(x, dy, gw, gb, g, inv_m, st);
            if (rc) return rc;
        }
    }
    const bool sp_d = math == DNB_MATH_FP32 && dx && !thin && split_ok(g, TC_CONV_DGRAD);
    const bool sp_w = math == DNB_MATH_FP32 && !thin && split_ok(g, TC_CONV_WGRAD);
    if (sp_d) {
        rc = conv_dgrad_split(w, dy, dx, g, ws, false, st);
        if (rc) return rc;
    }
    if (dx && !thin && !sp_d) {
        if (math >= DNB_MATH_TF32 && tc_conv_supported(g, TC_CONV_DGRAD)) {
            rc = tc_conv_dgrad(dy, w, dx, g, ws, math == DNB_MATH_BF16, st);
        } else {
            GatherArgs a{dy, w, nullptr, dx, g->B, g->F, g->OH, g->OW, g->C, g->H, g->W, g->kh, g->kw,
                         1, 1, g->sy, g->sx, g->kh - 1 - g->pt, g->kw - 1 - g->pl, 0, 0.f};
            rc = launch_gather<true>("conv2d_bwd_data", a, st);
        }
        if (rc) return rc;
    }
    if (thin && !tc_wgrad) {
        // gradients already accumulated above
    } else if (sp_w) {
        rc = conv_wgrad_split(x, dy, gw, gb, g, inv_m, ws, /*dy_split=*/sp_d, st);
    } else if (tc_wgrad) {
        rc = tc_conv_wgrad(x, dy, gw, gb, g, inv_m, ws, math == DNB_MATH_BF16, st);
    } else {
        WgradArgs wa{x, dy, gw, gb, g->B, g->C, g->H, g->W, g->F, g->OH, g->OW, g->kh, g->kw, g->sy, g->sx,
                     g->pt, g->pl, g->C * g->kh * g->kw, inv_m, 0};
        rc = launch_wgrad(wa, st);
    }
    if (rc) return rc;
    if (reg_w.l1 != 0.f || reg_w.l2 != 0.f || reg_b.l1 != 0.f || reg_b.l2 != 0.f) {
        long nw = (long)g->F * g->C * g->kh * g->kw;
        long n = nw > g->F ? nw : g->F;
        conv_reg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, gw, nw, b, gb, g->F, (float)(g->OH * g->OW) * reg_scale,
                                                                   reg_w.l1, reg_w.l2, reg_b.l1, reg_b.l2);
        DNB_LAUNCH_CHECK("conv2d_bwd_reg");
    }
    return DNB_OK;
}

// Conv2D (tensor-core shifted-view shapes) -> MaxPool2D -> LeakyRelu in one launch: convolution.py:193-225 + pooling.py:43-99 +
// activations.py:17-19; the convolution output and the pooled pre-activation never reach HBM (idx_gate: csrc/fused_cpa.cu format)
int dnb_conv2d_pool_act_supported(const dnb_win_t* conv, const dnb_win_t* pool, int math) {
    if (!conv || !pool || math != DNB_MATH_BF16 || getenv("DNB_NO_SV") || check_win("conv2d_pool_act", conv)) return 0;
    return tc_conv_svp_supported(conv, pool) ? 1 : 0;
}

int dnb_conv2d_pool_act_fwd(const float* x, const float* w, const float* b, float* act_out, uint8_t* idx_gate,
                            const dnb_win_t* conv, const dnb_win_t* pool, float alpha, void* ws, int math,
                            dnb_stream_t stream) {
    DNB_CHECK_ARG(dnb_conv2d_pool_act_supported(conv, pool, math), "conv2d_pool_act_fwd: unsupported geometry / math mode");
    if (conv->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && act_out && idx_gate, "conv2d_pool_act_fwd: null pointer");
    return tc_conv_fwd_pool(x, w, b, act_out, idx_gate, alpha, conv, pool, ws, S(stream));
}

// Data gradient alone (the second half of Conv2D.backward, convolution.py:244-250,271-273): what a caller runs when the
// parent's delta is wanted but the parameter gradients are not (Net materialises the delta of an Input layer lazily —
// nothing on the training path consumes it).  Same kernel selection as dnb_conv2d_bwd.
int dnb_conv2d_bwd_data(const float* w, const float* dy, float* dx, const dnb_win_t* g, void* ws, int math,
                        dnb_stream_t stream) {
    int rc = check_win("conv2d_bwd_data", g);
    if (rc) return rc;
    if (g->B == 0) return DNB_OK;
    DNB_CHECK_ARG(w && dy && dx, "conv2d_bwd_data: null pointer");
    cudaStream_t st = S(stream);
    const bool thin = thin_conv_supported(g);
    const bool thin_tc = thin && g->C >= 3 && (long)g->OH * g->OW >= 1024 && !(getenv("DNB_THIN_TC_WGRAD") && atoi(getenv("DNB_THIN_TC_WGRAD")) == 0);
    const bool tc = math >= DNB_MATH_TF32 && tc_conv_supported(g, TC_CONV_DGRAD);
    if (thin && !(thin_tc && tc)) return thin_conv_bwd(nullptr, w, dy, dx, nullptr, nullptr, g, 0.f, /*skip_wgrad=*/true, st);
    if (tc) return tc_conv_dgrad(dy, w, dx, g, ws, math == DNB_MATH_BF16, st);
    if (math == DNB_MATH_FP32 && split_ok(g, TC_CONV_DGRAD)) return conv_dgrad_split(w, dy, dx, g, ws, false, st);
    GatherArgs a{dy, w, nullptr, dx, g->B, g->F, g->OH, g->OW, g->C, g->H, g->W, g->kh, g->kw,
                 1, 1, g->sy, g->sx, g->kh - 1 - g->pt, g->kw - 1 - g->pl, 0, 0.f};
    return launch_gather<true>("conv2d_bwd_data", a, st);
}

int dnb_dwconv2d_fwd(const float* x, const float* w, const float* b, float* y, const dnb_win_t* g, dnb_stream_t stream) {
    int rc = check_win("dwconv2d_fwd", g);
    if (rc) return rc;
    DNB_CHECK_ARG(g->F == g->C, "dwconv2d_fwd: depthwise needs F == C");
    if (g->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && y, "dwconv2d_fwd: null pointer");
    long n = (long)g->B * g->C * g->OH * g->OW;
    if (dw3_ring_ok(g) && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15)) {
        const int planes = g->B * g->C;
        const size_t smem = 128 + (size_t)(DWR_ST + 2) * g->H * g->W * 4;
        cudaFuncSetAttribute(dw3_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dw3_ring_kernel<<<std::min(planes, 2 * kNumSMs), DWR_THREADS, smem, S(stream)>>>(x, w, b, y, planes, g->C, g->H, g->W, 0);
        DNB_LAUNCH_CHECK("dwconv2d_fwd");
        return DNB_OK;
    }
    dw_fwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, w, b, y, *g);
    DNB_LAUNCH_CHECK("dwconv2d_fwd");
    return DNB_OK;
}

int dnb_dwconv2d_bwd(const float* x, const float* w, const float* b, const float* dy,
                     float* dx, float* gw, float* gb, const dnb_win_t* g, float inv_m,
                     dnb_reg_t reg_b, float reg_scale, dnb_stream_t stream) {
    int rc = check_win("dwconv2d_bwd", g);
    if (rc) return rc;
    DNB_CHECK_ARG(g->F == g->C, "dwconv2d_bwd: depthwise needs F == C");
    if (g->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && dy && gw && gb, "dwconv2d_bwd: null pointer");
    DNB_CHECK_ARG(g->kh * g->kw <= 64, "dwconv2d_bwd: window larger than 64 taps");
    cudaStream_t st = S(stream);
    const bool ring = dw3_ring_ok(g) && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15);
    if (dx) {
        long n = (long)g->B * g->C * g->H * g->W;
        if (ring && !(reinterpret_cast<uintptr_t>(dx) & 15)) {
            const int planes = g->B * g->C;
            const size_t smem = 128 + (size_t)(DWR_ST + 2) * g->H * g->W * 4;
            cudaFuncSetAttribute(dw3_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            dw3_ring_kernel<<<std::min(planes, 2 * kNumSMs), DWR_THREADS, smem, st>>>(dy, w, nullptr, dx, planes, g->C, g->H, g->W, 1);
        } else {
            dw_dgrad_kernel<<<ew_grid(n, 256), 256, 0, st>>>(dy, w, dx, *g);
        }
        DNB_LAUNCH_CHECK("dwconv2d_bwd_data");
    }
    if (ring) {
        const int splits = std::max(1, std::min(g->B, (2 * kNumSMs + g->C - 1) / g->C));
        const size_t smem = 128 + (size_t)DWR_ST * 2 * g->H * g->W * 4;
        cudaFuncSetAttribute(dw3_wgrad_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dw3_wgrad_ring_kernel<<<dim3(splits, g->C), DWR_THREADS, smem, st>>>(x, dy, gw, gb, g->B, g->C, g->H, g->W, inv_m);
        DNB_LAUNCH_CHECK("dwconv2d_bwd_weight");
        if (reg_b.l1 != 0.f || reg_b.l2 != 0.f) {
            conv_reg_kernel<<<(g->C + 255) / 256, 256, 0, st>>>(w, gw, 0, b, gb, g->C, (float)(g->OH * g->OW) * reg_scale,
                                                               0.f, 0.f, reg_b.l1, reg_b.l2);
            DNB_LAUNCH_CHECK("dwconv2d_bwd_reg");
        }
        return DNB_OK;
    }
    long npix = (long)g->B * g->OH * g->OW;
    int splits = (int)max(1L, min((npix + 255) / 256, (long)(kNumSMs * 8 / g->C + 1)));
    dim3 grid(splits, g->C);
    int khw = g->kh * g->kw;
    if (khw <= 4) dw_wgrad_kernel<4><<<grid, 256, 0, st>>>(x, dy, gw, gb, *g, inv_m);
    else if (khw <= 9) dw_wgrad_kernel<9><<<grid, 256, 0, st>>>(x, dy, gw, gb, *g, inv_m);
    else if (khw <= 25) dw_wgrad_kernel<25><<<grid, 256, 0, st>>>(x, dy, gw, gb, *g, inv_m);
    else dw_wgrad_kernel<64><<<grid, 256, 0, st>>>(x, dy, gw, gb, *g, inv_m);
    DNB_LAUNCH_CHECK("dwconv2d_bwd_weight");
    if (reg_b.l1 != 0.f || reg_b.l2 != 0.f) {
        conv_reg_kernel<<<(g->C + 255) / 256, 256, 0, st>>>(w, gw, 0, b, gb, g->C, (float)(g->OH * g->OW) * reg_scale,
                                                           0.f, 0.f, reg_b.l1, reg_b.l2);
        DNB_LAUNCH_CHECK("dwconv2d_bwd_reg");
    }
    return DNB_OK;
}

}  // extern "C"
