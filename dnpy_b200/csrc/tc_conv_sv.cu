// "Shifted-view" implicit-GEMM convolution for stride-1 layers whose gathered tensor has a multiple of 32
// channels (the tensor-core shapes of the BASELINE configs: conv2 of the MNIST CNN, the 32-channel layers
// of the residual net) — forward and data-gradient.
//
// Idea: stage each (image, row band) ONCE in shared memory as zero-padded NHWC — one 128-byte row per
// padded pixel and 32-channel block, stored in the K-major 128B-swizzled UMMA layout.  Because the
// hardware swizzle is a function of the absolute shared-memory address, a UMMA descriptor may start at
// ANY 128-byte row: the A operand of filter tap (i, j) is then simply the same staged image read from a
// start address shifted by (i*PW + j) rows.  All kh*kw taps reuse the one staged copy; no im2col tile is
// ever built, in global OR shared memory.  (Verified on B200: scratch/shift_test.cu.)
//
//   M index  m = unit*PH*PW + oy*PW + ox   (columns ox >= OW / rows oy >= TH are garbage, never stored)
//   A(tap,cb) = region[cb] rows [m0 + i*PW + j, +128)        B(tap,cb) = filters [BN x 32], resident in smem
//
// Persistent, warp-specialised CTA (1 per SM): 8 loader warps (NCHW -> swizzled NHWC, 4 coalesced loads +
// one 16-byte store per task, double-buffered units), 1 MMA-issuing thread (tcgen05.mma.kind::tf32, two
// TMEM accumulator stages), 4 epilogue warps (tcgen05.ld -> bias*K -> NCHW stores) running concurrently
// on the previous tile.  Filters are fetched once per CTA by TMA and stay resident.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {
namespace tc {

struct SvParams {
    const float* src; float* dst; const float* bias;
    int B, Cs, Hs, Ws, Cd, Hd, Wd, kh, kw;
    int off_y, off_x;            // placement of the source inside the padded plane
    int TH, nbands;              // output rows per band, bands per image
    int PH, PW;                  // padded band: TH + kh - 1, Wd + kw - 1
    int NI;                      // units (image bands) per group
    int rows_group;              // NI * PH * PW
    int tiles;                   // ceil(rows_group / 128)
    int region_rows;             // rows per 32-channel region of a buffer: tiles*128 + (kh-1)*PW + kw
    int nkb;                     // kh*kw*(Cs/32)
    long units;                  // B * nbands
    int debug;                   // experiments: 1 = no epilogue stores, 2 = no MMAs, 4 = no loads
    float bias_k;
};

constexpr int SV_LOADERS = 8, SV_EPI = 4;
constexpr int SV_THREADS = (SV_LOADERS + SV_EPI + 1) * 32;

// K33: 3x3 window known at compile time -> the MMA issue loop is fully unrolled, every descriptor is
// (uniform base + immediate); CB = 32-channel blocks of the gathered tensor (0 = run-time).
template <int BN, bool K33, int CB>
__global__ void __launch_bounds__(SV_THREADS, 1)
tc_conv_sv_kernel(const __grid_constant__ CUtensorMap tmW, SvParams p) {
    constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    constexpr int W_KB = BN * 128;                        // bytes of one K block of filters
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int cblks = p.Cs >> 5;
    const uint32_t region_bytes = (uint32_t)p.region_rows * 128;
    const uint32_t buf_bytes = ((region_bytes * cblks + 1023) / 1024) * 1024;
    uint8_t* sW = smem;                                    // [nkb][BN x 128 B]
    uint8_t* sX = sW + (size_t)p.nkb * W_KB;               // 2 buffers x cblks regions
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * (size_t)buf_bytes);
    uint64_t* w_full = bars;
    uint64_t* x_full = bars + 1;       // [2]
    uint64_t* x_empty = bars + 3;      // [2]
    uint64_t* acc_full = bars + 5;     // [2]
    uint64_t* acc_empty = bars + 7;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* s_bias = reinterpret_cast<float*>(bars + 10);   // [BN]  K*bias (0 without a bias): no global load in the store loop

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const long ngroups = (p.units + p.NI - 1) / p.NI;
    const int PHPW = p.PH * p.PW;

    // zero both image buffers once (slack rows and never-written garbage must stay finite)
    for (uint32_t e = tid; e < 2 * buf_bytes / 16; e += SV_THREADS) reinterpret_cast<float4*>(sX)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int f = tid; f < (BN < 32 ? 32 : BN); f += SV_THREADS) s_bias[f] = (p.bias && f < p.Cd) ? p.bias_k * __ldg(p.bias + f) : 0.f;
    if (tid == 0) {
        tma_prefetch_desc(&tmW);
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&x_full[s], SV_LOADERS); mbar_init(&x_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], SV_EPI);
        }
        fence_barrier_init();
    }
    if (warp == SV_LOADERS + SV_EPI) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < SV_LOADERS) {
        // ===================== loaders: NCHW global -> swizzled NHWC shared =====================
        const long plane_s = (long)p.Hs * p.Ws;
        const int tasks_unit = PHPW * 8 * cblks;           // (pixel, channel quad, channel block)
        // The task list of a thread is the same for every unit (fixed geometry): decode it ONCE.
        //   t_src : element offset of channel (cb*32+quad*4) at padded row 0 / this column, -1 if the column is padding
        //   t_dst : byte offset inside the buffer (swizzled 16-byte chunk), t_yy : padded row
        constexpr int TB = 14, LT = SV_LOADERS * 32;       // up to 14 tasks per thread per unit (3584 tasks)
        int t_src[TB], t_yy[TB];
        uint32_t t_dst[TB];
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const int t = tid + r * LT;
            t_src[r] = -1; t_yy[r] = 0; t_dst[r] = 0xffffffffu;
            if (t < tasks_unit) {
                const int pp = t % PHPW, qc = t / PHPW;          // pixel index fastest across lanes -> coalesced along x
                const int quad = qc & 7, cb = qc >> 3;
                const int yy = pp / p.PW, xx = pp - yy * p.PW;
                const int x = xx - p.off_x;
                t_yy[r] = yy;
                if (x >= 0 && x < p.Ws) t_src[r] = (cb * 32 + quad * 4) * (int)plane_s + x;
                t_dst[r] = cb * region_bytes + (uint32_t)pp * 128 + (((quad ^ (pp & 7)) & 7) << 4);
            }
        }
        const bool unit_rows_mult8 = (PHPW & 7) == 0;
        // Software pipeline (NI == 1 fast path): the loads of the NEXT unit are issued right after this unit has
        // been handed to the MMA warp, so their DRAM latency is hidden behind the current unit's MMAs and the
        // wait for the ring slot.
        float4 v[TB];
        auto issue_loads = [&](long u) {
            const bool uv = u < p.units;
            const int b = uv ? (int)(u / p.nbands) : 0;
            const int band = uv ? (int)(u % p.nbands) : 0;
            const int y_base = band * p.TH - p.off_y;       // source row of padded row 0
            const float* sp = p.src + (long)b * p.Cs * plane_s;
#pragma unroll
            for (int r = 0; r < TB; ++r) {                  // every load of the unit is in flight before the first store
                v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int y = y_base + t_yy[r];
                if (uv && t_src[r] >= 0 && y >= 0 && y < p.Hs && !(p.debug & 4)) {
                    const float* gp = sp + t_src[r] + y * p.Ws;
                    v[r].x = __ldg(gp); v[r].y = __ldg(gp + plane_s);
                    v[r].z = __ldg(gp + 2 * plane_s); v[r].w = __ldg(gp + 3 * plane_s);
                }
            }
        };
        auto store_unit = [&](uint32_t buf_s, int k) {
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                if (t_dst[r] != 0xffffffffu) {
                    uint32_t d = t_dst[r] + (uint32_t)k * PHPW * 128;
                    if (!unit_rows_mult8 && k) {
                        // unit k starts at row k*PHPW: redo the swizzle phase for the shifted row index
                        const uint32_t row = (t_dst[r] % region_bytes) / 128 + (uint32_t)k * PHPW;
                        const uint32_t chunk0 = ((t_dst[r] >> 4) & 7) ^ ((t_dst[r] >> 7) & 7);     // = quad
                        d = (t_dst[r] / region_bytes) * region_bytes + row * 128 + (((chunk0 ^ (row & 7)) & 7) << 4);
                    }
                    sts128(buf_s + d, v[r]);
                }
            }
        };
        int it = 0;
        if (p.NI == 1 && (long)blockIdx.x < ngroups) issue_loads(blockIdx.x);
        for (long g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
            const int s = it & 1;
            if (it >= 2) mbar_wait(&x_empty[s], ((it >> 1) - 1) & 1);
            const uint32_t buf_s = smem_u32(sX + (size_t)s * buf_bytes);
            if (p.NI == 1) {
                store_unit(buf_s, 0);
            } else {
                for (int k = 0; k < p.NI; ++k) { issue_loads(g * p.NI + k); store_unit(buf_s, k); }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_full[s]);       // one arrival per loader warp
            if (p.NI == 1 && g + gridDim.x < ngroups) issue_loads(g + gridDim.x);
        }
    } else if (warp < SV_LOADERS + SV_EPI) {
        // ===================== epilogue: TMEM -> NCHW global =====================
        const int q = warp & 3;                             // TMEM lane quadrant of this warp
        const size_t plane_d = (size_t)p.Hd * p.Wd;
        const uint32_t s_bias_u32 = smem_u32(s_bias);
        int tcount = 0;
        for (long g = blockIdx.x; g < ngroups; g += gridDim.x) {
            for (int t = 0; t < p.tiles; ++t, ++tcount) {
                const int as = tcount & 1;
                const int m = t * 128 + q * 32 + lane;
                const int k = m / PHPW, rem = m - k * PHPW;
                const int oy = rem / p.PW, ox = rem - oy * p.PW;
                const long u = g * p.NI + k;
                const int band = (int)(u % p.nbands);
                const int y = band * p.TH + oy;
                const bool valid = k < p.NI && u < p.units && oy < p.TH && y < p.Hd && ox < p.Wd;
                const int b = valid ? (int)(u / p.nbands) : 0;
                float* dp = p.dst + ((size_t)b * p.Cd) * plane_d + (size_t)y * p.Wd + ox;
                mbar_wait(&acc_full[as], (tcount >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), v);
                    float bv[32];                           // K*bias of these 32 channels: 8 broadcast LDS.128 under the TMEM load
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b4 = lds128_ro(s_bias_u32 + (uint32_t)(c0 + 4 * j4) * 4);
                        bv[4 * j4] = b4.x; bv[4 * j4 + 1] = b4.y; bv[4 * j4 + 2] = b4.z; bv[4 * j4 + 3] = b4.w;
                    }
                    tmem_ld_wait();
                    if (valid && !(p.debug & 1)) {
                        float* o = dp + (size_t)c0 * plane_d;
                        const int nvalid = min(32, min(BN - c0, p.Cd - c0));
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) { *o = __uint_as_float(v[j]) + bv[j]; o += plane_d; }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
            }
        }
    } else {
        // ===================== filter TMA (once) + MMA issue =====================
        // The whole warp walks the loop (uniform control flow and descriptor arithmetic); one elected lane
        // issues.  Measured on B200: 43-49 cycles per 128xNx8 TF32 MMA this way vs 75 with a one-thread loop.
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(p.nkb * W_KB));
            for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sW + (size_t)kb * W_KB, &tmW, kb * 32, 0, w_full);
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        const uint32_t idesc = make_idesc_tf32(128, BN, 0, 0);
        const uint64_t w_d0 = make_smem_desc(smem_u32(sW), 16, 1024);
        const uint32_t pw8 = (uint32_t)p.PW * 8, region16 = region_bytes >> 4;
        int it = 0, tcount = 0;
        for (long g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
            const int s = it & 1;
            mbar_wait(&x_full[s], (it >> 1) & 1);
            tc_fence_after();
            const uint64_t x_d0 = make_smem_desc(smem_u32(sX + (size_t)s * buf_bytes), 16, 1024);
            for (int t = 0; t < p.tiles; ++t, ++tcount) {
                const int as = tcount & 1;
                if (tcount >= 2) { mbar_wait(&acc_empty[as], ((tcount >> 1) - 1) & 1); tc_fence_after(); }
                const uint32_t d_t = tmem_base + as * BN;
                // descriptor address fields only ever grow by < 256 KB >> 4: a plain 64-bit add never carries out
                const uint64_t a_t = x_d0 + (uint64_t)((uint32_t)t * (128 * 8));
                // ONE elected-thread region per tile (ptxas keeps the operands in uniform registers only when the
                // branch is on the elect itself); inside it the 3x3 case is a straight line of MMAs whose
                // descriptors are (tile base + immediate)
                if (elect_one()) {
                    if (!(p.debug & 2)) {
                        if (K33 && CB > 0) {
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                const uint64_t a_i = a_t + (uint64_t)((uint32_t)i * pw8);
#pragma unroll
                                for (int j = 0; j < 3; ++j)
#pragma unroll
                                    for (int cb = 0; cb < CB; ++cb) {
                                        const int kb = (i * 3 + j) * CB + cb;
                                        const uint64_t a = a_i + (uint64_t)(j * 8) + (uint64_t)((uint32_t)cb * region16);
                                        const uint64_t bw = w_d0 + (uint64_t)(kb * (W_KB >> 4));
                                        if (kb == 0) mma_tf32_init(d_t, a, bw, idesc); else mma_tf32_acc(d_t, a, bw, idesc);
                                        mma_tf32_acc(d_t, a + 2, bw + 2, idesc);
                                        mma_tf32_acc(d_t, a + 4, bw + 4, idesc);
                                        mma_tf32_acc(d_t, a + 6, bw + 6, idesc);
                                    }
                            }
                        } else {
                            int kb = 0;
                            for (int i = 0; i < p.kh; ++i)
                                for (int j = 0; j < p.kw; ++j) {
                                    const uint64_t a_ij = a_t + (uint64_t)((uint32_t)i * pw8 + (uint32_t)(j * 8));
                                    for (int cb = 0; cb < cblks; ++cb, ++kb) {
                                        const uint64_t a = a_ij + (uint64_t)((uint32_t)cb * region16);
                                        const uint64_t bw = w_d0 + (uint64_t)(kb * (W_KB >> 4));
                                        if (kb == 0) mma_tf32_init(d_t, a, bw, idesc); else mma_tf32_acc(d_t, a, bw, idesc);
                                        mma_tf32_acc(d_t, a + 2, bw + 2, idesc);
                                        mma_tf32_acc(d_t, a + 4, bw + 4, idesc);
                                        mma_tf32_acc(d_t, a + 6, bw + 6, idesc);
                                    }
                                }
                        }
                    }
                    mma_commit(&acc_full[as]);
                    if (t == p.tiles - 1) mma_commit(&x_empty[s]);   // fires when every MMA that read this buffer is done
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SV_LOADERS + SV_EPI) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t sv_smem_bytes(const SvParams& p, int bn) {
    const size_t region = (size_t)p.region_rows * 128;
    const size_t buf = ((region * (p.Cs / 32) + 1023) / 1024) * 1024;
    return (size_t)p.nkb * bn * 128 + 2 * buf + 128 + (size_t)(bn < 32 ? 32 : bn) * 4 + 1024;
}

// choose band height and units per group so that two image buffers + the resident filters fit in 227 KB
static bool sv_plan(SvParams& p, int bn) {
    p.PW = p.Wd + p.kw - 1;
    p.nkb = p.kh * p.kw * (p.Cs / 32);
    const size_t budget = 225 * 1024;
    for (int th = p.Hd; th >= 1; th = (th > 8 ? th / 2 : th - 1)) {
        p.TH = th;
        p.nbands = (p.Hd + th - 1) / th;
        p.PH = th + p.kh - 1;
        const int unit_rows = p.PH * p.PW;
        if ((long)unit_rows * 8 * (p.Cs / 32) > 14L * SV_LOADERS * 32) continue;     // loader task table: 14 per thread
        int best_ni = 0; double best_eff = 0.0;
        for (int ni = 1; ni <= 8; ++ni) {
            p.NI = ni; p.rows_group = ni * unit_rows; p.tiles = (p.rows_group + 127) / 128;
            // multiple of 8 rows: every 32-channel region must start on a 1024-byte swizzle-pattern boundary
            p.region_rows = (p.tiles * 128 + (p.kh - 1) * p.PW + p.kw + 7) & ~7;
            if (sv_smem_bytes(p, bn) > budget) break;
            const double eff = (double)p.rows_group / (p.tiles * 128);
            if (eff > best_eff + 0.02) { best_eff = eff; best_ni = ni; }
        }
        if (best_ni) {
            p.NI = best_ni; p.rows_group = best_ni * unit_rows; p.tiles = (p.rows_group + 127) / 128;
            p.region_rows = (p.tiles * 128 + (p.kh - 1) * p.PW + p.kw + 7) & ~7;
            p.units = (long)p.B * p.nbands;
            return true;
        }
    }
    return false;
}

static int pick_bn_sv(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }

template <int BN, bool K33, int CB>
static int sv_launch_k(const char* name, const CUtensorMap& tw, const SvParams& p, cudaStream_t st) {
    const size_t smem = sv_smem_bytes(p, BN);
    cudaFuncSetAttribute(tc_conv_sv_kernel<BN, K33, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long ngroups = (p.units + p.NI - 1) / p.NI;
    const int grid = (int)std::min<long>(ngroups, kNumSMs);
    tc_conv_sv_kernel<BN, K33, CB><<<grid, SV_THREADS, smem, st>>>(tw, p);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

template <int BN>
static int sv_launch(const char* name, const CUtensorMap& tw, const SvParams& p, cudaStream_t st) {
    if (p.kh == 3 && p.kw == 3 && p.Cs == 32) return sv_launch_k<BN, true, 1>(name, tw, p, st);
    if (p.kh == 3 && p.kw == 3 && p.Cs == 64) return sv_launch_k<BN, true, 2>(name, tw, p, st);
    return sv_launch_k<BN, false, 0>(name, tw, p, st);
}

}  // namespace tc

using namespace tc;

// dims of the gathered tensor / produced tensor for the two uses
static SvParams sv_make(const dnb_win_t* g, bool dgrad) {
    SvParams p{};
    p.B = g->B; p.kh = g->kh; p.kw = g->kw;
    if (!dgrad) { p.Cs = g->C; p.Hs = g->H; p.Ws = g->W; p.Cd = g->F; p.Hd = g->OH; p.Wd = g->OW; p.off_y = g->pt; p.off_x = g->pl; }
    else { p.Cs = g->F; p.Hs = g->OH; p.Ws = g->OW; p.Cd = g->C; p.Hd = g->H; p.Wd = g->W; p.off_y = g->kh - 1 - g->pt; p.off_x = g->kw - 1 - g->pl; }
    return p;
}

bool tc_conv_sv_supported(const dnb_win_t* g, bool dgrad) {
    if (g->sy != 1 || g->sx != 1) return false;
    SvParams p = sv_make(g, dgrad);
    if (p.Cs % 32 || p.Cd > 128 || p.off_y < 0 || p.off_x < 0) return false;
    // the produced extent must equal (gathered + pads - k + 1): true for forward by construction; for dgrad it
    // requires that the forward windows tile the padded input exactly (no unused trailing rows/columns)
    if (dgrad && ((g->OH - 1) + g->kh != g->H + g->pt + g->pb || (g->OW - 1) + g->kw != g->W + g->pl + g->pr)) return false;
    return sv_plan(p, pick_bn_sv(p.Cd));
}

int tc_conv_sv_run(const char* name, const float* src, const float* wprep, int n_pad, int Kp, const float* bias, float bias_k,
                   float* dst, const dnb_win_t* g, bool dgrad, cudaStream_t st) {
    SvParams p = sv_make(g, dgrad);
    p.src = src; p.dst = dst; p.bias = bias; p.bias_k = bias_k;
    p.debug = getenv("DNB_SV_DEBUG") ? atoi(getenv("DNB_SV_DEBUG")) : 0;
    const int bn = pick_bn_sv(p.Cd);
    if (!sv_plan(p, bn)) return set_err(DNB_ERR_UNSUPPORTED, "%s: shape does not fit the shifted-view kernel", name);
    CUtensorMap tw;
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)n_pad}, str[1] = {(uint64_t)Kp * 4};
    uint32_t box[2] = {32, (uint32_t)bn};
    if (!make_tmap_f32(&tw, wprep, 2, dims, str, box)) return set_err(DNB_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", name);
    switch (bn) {
        case 16: return sv_launch<16>(name, tw, p, st);
        case 32: return sv_launch<32>(name, tw, p, st);
        case 64: return sv_launch<64>(name, tw, p, st);
        default: return sv_launch<128>(name, tw, p, st);
    }
}

}  // namespace dnb
