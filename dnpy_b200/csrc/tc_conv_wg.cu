// "Shifted-view" weight gradient for 3x3 stride-1 convolutions with 32 input channels and small images (conv2 of
// the MNIST CNN: 32 x 12 x 12 -> 64 x 12 x 12): gW[f][c][i][j] = (1/m) sum_{b,oy,ox} dY[b,f,oy,ox] * Xp[b,c,oy+i,ox+j].
//
// The reduction runs over pixels, so both operands are staged PIXEL-major (one 128-byte row per pixel of the
// zero-padded PW-wide grid, 32 channels / filters per row) and fed to the tensor core as MN-major TF32 operands
// (128B swizzle with 32-byte atoms):
//     A = dY  rows m = oy*PW + ox (zero in the pad columns)      M = filters (32-element chunks, LBO = region pitch)
//     B = Xp  rows m + i*PW + j                                   N = (j, c): the three column taps of a kernel row are
//                                                                  three 32-channel chunks ONE ROW APART (LBO = 128 B)
//     D_i[f][(j,c)] += A[k][f] * B[k][(j,c)]                      K = 8 pixel rows per MMA, 3 MMAs (i = 0,1,2) per K step
// Like the forward kernel nothing is ever im2col'ed: a tap is a start-address shift of the one staged image (the
// swizzle is a function of the absolute address), and here even the j taps are folded into the N extent of one MMA.
// The three accumulators D_i (128 x 96 fp32, TMEM) persist over ALL images of the CTA; one epilogue at the end.
//
//   warp 0      TMA producer: raw NCHW x image + dY image (4-D tiled maps, no swizzle) per unit
//   warp 1      MMA issue (tcgen05.mma.kind::tf32, M=128 N=96, both operands MN-major)
//   warps 2-9   transposers raw -> swizzled pixel-major (conflict-free LDS, STS.128) + running sum of dY for gb;
//               afterwards the same warps read TMEM and add the CTA's partial gW / gb atomically.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {
namespace tc {

struct WgSvParams {
    float* gw; float* gb;
    int B, H, W, F, OH, OW, pt, pl;
    int PW;                  // padded row pitch in pixels: OW + 2
    int TH, nbands;          // output rows per unit (band), bands per image
    int KR;                  // reduction rows per unit: round_up(TH*PW, 8)
    int xs_rows;             // rows of one staged X buffer (padded image + tap/chunk overhang), multiple of 8
    int fr;                  // 32-filter regions of dY: F / 32
    int nraw;                // raw TMA stages (2..4)
    int flat;                // tensor maps are (H*W, C, B): a band of a channel is one contiguous box row
    int debug;               // timing experiments (DNB_WG_DEBUG): 1 no epilogue atomics, 2 no MMAs, 4 no transposition
    float inv_m;
};

constexpr int WG_TR = 8;                                   // transposer warps
constexpr int WGSV_THREADS = (2 + WG_TR) * 32;
constexpr int WG_TBX = 6, WG_TBD = 10;                     // task slots per transposer thread (x / dY)

__device__ __forceinline__ void wg_tma_load_4d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ float wg_lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
// byte offset of 4 consecutive MN elements (16 bytes, quad q8 of 8) of K row `row` in the 32B-atom 128B swizzle
__device__ __forceinline__ uint32_t sw32_off(int row, int q8) {
    return (uint32_t)row * 128 + ((((q8 >> 1) ^ (row & 3)) & 3) << 5) + ((q8 & 1) << 4);
}

// M64: F <= 64 -> M=64 MMAs (half the A-operand shared-memory traffic); D row r then lives in TMEM lane (r/16)*32 + r%16
template <bool M64>
__global__ void __launch_bounds__(WGSV_THREADS, 1)
tc_conv_wgrad_sv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, WgSvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t dreg_bytes = (uint32_t)p.KR * 128;                    // one 32-filter region of a dY buffer
    const uint32_t dbuf_bytes = dreg_bytes * p.fr;
    const uint32_t xbuf_bytes = (uint32_t)p.xs_rows * 128;
    const int XH = p.TH + 2;                                             // x rows of a band (with halo)
    const uint32_t rawx_bytes = (uint32_t)XH * p.W * 32 * 4, rawd_box = (uint32_t)p.TH * p.OW * 32 * 4;
    const uint32_t raw_bytes = rawx_bytes + rawd_box * p.fr;             // one raw stage: x band, then fr dY bands
    uint8_t* sD = smem;                                    // 2 x fr regions (A operand; chunks >= fr of the M=128 MMA alias what follows)
    uint8_t* sX = sD + 2 * (size_t)dbuf_bytes;             // 2 buffers (B operand)
    uint8_t* sR = sX + 2 * (size_t)xbuf_bytes;             // 2 raw stages: x band [32][TH+2][W], dY band [F][TH][OW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + (size_t)p.nraw * raw_bytes);
    uint64_t* raw_full = bars;         // [4] tx
    uint64_t* raw_empty = bars + 4;    // [4] WG_TR
    uint64_t* xd_full = bars + 8;      // [2] WG_TR
    uint64_t* xd_empty = bars + 10;    // [2] commit
    uint64_t* acc_full = bars + 12;    // commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* s_gb = reinterpret_cast<float*>(bars + 14);     // [128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const long units = (long)p.B * p.nbands;
    const int nimg = (int)((units - (long)blockIdx.x + (long)gridDim.x - 1) / (long)gridDim.x);     // units of this CTA

    // zero the staged buffers once: pad ring / pad columns / overhang rows are never written afterwards
    for (uint32_t e = tid; e < (2 * dbuf_bytes + 2 * xbuf_bytes) / 16; e += WGSV_THREADS)
        reinterpret_cast<float4*>(sD)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 128) s_gb[tid] = 0.f;
    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmD);
        for (int s = 0; s < 4; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], WG_TR); }
        for (int s = 0; s < 2; ++s) { mbar_init(&xd_full[s], WG_TR); mbar_init(&xd_empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int s = 0, rph = 0;
            for (int n = 0; n < nimg; ++n) {
                if (n >= p.nraw) mbar_wait(&raw_empty[s], rph ^ 1);
                const long u = (long)blockIdx.x + (long)n * gridDim.x;
                const int b = (int)(u / p.nbands), band = (int)(u % p.nbands);
                uint8_t* rs = sR + (size_t)s * raw_bytes;
                mbar_arrive_expect_tx(&raw_full[s], raw_bytes);
                // x rows [band*TH - pt, +TH+2) and dY rows [band*TH, +TH): whatever falls outside the image is zero filled
                if (p.flat) {
                    tma_load_3d(rs, &tmX, (band * p.TH - p.pt) * p.W, 0, b, &raw_full[s]);
                    for (int r = 0; r < p.fr; ++r) tma_load_3d(rs + rawx_bytes + (size_t)r * rawd_box, &tmD, band * p.TH * p.OW, r * 32, b, &raw_full[s]);
                } else {
                    wg_tma_load_4d(rs, &tmX, 0, band * p.TH - p.pt, 0, b, &raw_full[s]);
                    for (int r = 0; r < p.fr; ++r) wg_tma_load_4d(rs + rawx_bytes + (size_t)r * rawd_box, &tmD, 0, band * p.TH, r * 32, b, &raw_full[s]);
                }
                if (++s == p.nraw) { s = 0; rph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issue =====================
        const uint32_t idesc = make_idesc_tf32(M64 ? 64 : 128, 96, 1, 1);
        const int ksteps = p.KR >> 3;
        const uint32_t pw8 = (uint32_t)p.PW * 8;
        for (int n = 0; n < nimg; ++n) {
            const int s = n & 1;
            mbar_wait(&xd_full[s], (n >> 1) & 1);
            tc_fence_after();
            // A: M chunks of 32 filters dreg_bytes apart, 4-row atoms 512 B apart; B: N chunks = the same rows + j
            const uint64_t a0 = make_smem_desc(smem_u32(sD + (size_t)s * dbuf_bytes), dreg_bytes, 512, LAYOUT_SW128_BASE32B);
            const uint64_t b0 = make_smem_desc(smem_u32(sX + (size_t)s * xbuf_bytes), 128, 512, LAYOUT_SW128_BASE32B);
            if (elect_one()) {
                for (int ks = 0; ks < ((p.debug & 2) ? 0 : ksteps); ++ks) {
                    const uint64_t a = a0 + (uint64_t)((uint32_t)ks * 64);                 // 8 rows = 1024 B
                    const uint64_t bq = b0 + (uint64_t)((uint32_t)ks * 64);
                    const uint32_t acc = (n > 0 || ks > 0) ? 1u : 0u;
                    mma_tf32(tmem_base, a, bq, idesc, acc);
                    mma_tf32(tmem_base + 96, a, bq + (uint64_t)pw8, idesc, acc);
                    mma_tf32(tmem_base + 192, a, bq + (uint64_t)(2 * pw8), idesc, acc);
                }
                mma_commit(&xd_empty[s]);
                if (n == nimg - 1) mma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ===================== transposers =====================
        const int ttid = tid - 64;
        constexpr int LT = WG_TR * 32;
        const int npx_x = XH * p.W, npx_d = p.TH * p.OW;
        const int tasks_x = npx_x * 8, tasks_d = npx_d * 8 * p.fr;
        const uint32_t chx = (uint32_t)npx_x * 4, chd = (uint32_t)npx_d * 4;        // bytes between channels of the raw images
        uint32_t xs_src[WG_TBX], xs_dst[WG_TBX], ds_src[WG_TBD], ds_dst[WG_TBD];
#pragma unroll
        for (int r = 0; r < WG_TBX; ++r) {
            const int t = ttid + r * LT;
            xs_src[r] = 0; xs_dst[r] = 0xffffffffu;
            if (t < tasks_x) {
                const int pp = t % npx_x, quad = t / npx_x;
                const int y = pp / p.W, x = pp - y * p.W;
                const int row = y * p.PW + x + p.pl;            // band-local padded row (the halo / pad rows arrive zero filled)
                xs_src[r] = (uint32_t)(quad * 4) * chx + (uint32_t)pp * 4;
                xs_dst[r] = sw32_off(row, quad);
            }
        }
#pragma unroll
        for (int r = 0; r < WG_TBD; ++r) {
            const int t = ttid + r * LT;
            ds_src[r] = 0; ds_dst[r] = 0xffffffffu;
            if (t < tasks_d) {
                const int pp = t % npx_d, qf = t / npx_d;               // qf: filter quad 0 .. 8*fr-1
                const int oy = pp / p.OW, ox = pp - oy * p.OW;
                const int row = oy * p.PW + ox;
                ds_src[r] = (uint32_t)(qf * 4) * chd + (uint32_t)pp * 4;
                ds_dst[r] = (uint32_t)(qf >> 3) * dreg_bytes + sw32_off(row, qf & 7);
            }
        }
        float gbs[WG_TBD][4];
#pragma unroll
        for (int r = 0; r < WG_TBD; ++r) { gbs[r][0] = 0.f; gbs[r][1] = 0.f; gbs[r][2] = 0.f; gbs[r][3] = 0.f; }
        int rs = 0, rph = 0;
        for (int n = 0; n < nimg; ++n) {
            const int s = n & 1;
            mbar_wait(&raw_full[rs], rph);
            if (n >= 2) mbar_wait(&xd_empty[s], ((n >> 1) - 1) & 1);
            const uint32_t rx = smem_u32(sR + (size_t)rs * raw_bytes), rd = rx + rawx_bytes;
            const uint32_t xb = smem_u32(sX + (size_t)s * xbuf_bytes), db = smem_u32(sD + (size_t)s * dbuf_bytes);
            if (!(p.debug & 4)) {
#pragma unroll
            for (int r0 = 0; r0 < WG_TBX; r0 += 3) {
                float4 v[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (xs_dst[r0 + r] != 0xffffffffu) {
                        const uint32_t a = rx + xs_src[r0 + r];
                        v[r].x = wg_lds32(a); v[r].y = wg_lds32(a + chx); v[r].z = wg_lds32(a + 2 * chx); v[r].w = wg_lds32(a + 3 * chx);
                    }
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (xs_dst[r0 + r] != 0xffffffffu) sts128(xb + xs_dst[r0 + r], v[r]);
            }
#pragma unroll
            for (int r0 = 0; r0 < WG_TBD; r0 += 5) {
                float4 v[5];
#pragma unroll
                for (int r = 0; r < 5; ++r)
                    if (ds_dst[r0 + r] != 0xffffffffu) {
                        const uint32_t a = rd + ds_src[r0 + r];
                        v[r].x = wg_lds32(a); v[r].y = wg_lds32(a + chd); v[r].z = wg_lds32(a + 2 * chd); v[r].w = wg_lds32(a + 3 * chd);
                    }
#pragma unroll
                for (int r = 0; r < 5; ++r)
                    if (ds_dst[r0 + r] != 0xffffffffu) {
                        sts128(db + ds_dst[r0 + r], v[r]);
                        gbs[r0 + r][0] += v[r].x; gbs[r0 + r][1] += v[r].y; gbs[r0 + r][2] += v[r].z; gbs[r0 + r][3] += v[r].w;
                    }
            }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&xd_full[s]); mbar_arrive(&raw_empty[rs]); }
            if (++rs == p.nraw) { rs = 0; rph ^= 1; }
        }
        // ---- bias gradient: per-slot sums -> shared -> global
        if (p.gb) {
#pragma unroll
            for (int r = 0; r < WG_TBD; ++r) {
                const int t = ttid + r * LT;
                if (t < tasks_d) {
                    const int f0 = (t / npx_d) * 4;
#pragma unroll
                    for (int u = 0; u < 4; ++u) atomicAdd(&s_gb[f0 + u], gbs[r][u]);
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WG_TR * 32) : "memory");
        if (p.gb && ttid < p.F && nimg > 0) atomicAdd(p.gb + ttid, s_gb[ttid] * p.inv_m);
        // ---- epilogue: D_i[f][(j,c)] -> gW[f][c][i][j]
        if (nimg > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int q = warp & 3, half = (warp - 2) >> 2;           // lane quadrant, column half (144 columns each)
            const int f = M64 ? (lane < 16 ? q * 16 + lane : p.F) : q * 32 + lane;
            if (M64 || q * 32 < p.F) {
#pragma unroll 1
                for (int c0 = half * 144; c0 < half * 144 + 144; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (f < p.F && !(p.debug & 1)) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const int col = c0 + u;
                            const int i = col / 96, jc = col - i * 96, j = jc >> 5, c = jc & 31;
                            atomicAdd(p.gw + (((size_t)f * 32 + c) * 3 + i) * 3 + j, __uint_as_float(v[u]) * p.inv_m);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t wgsv_raw(const WgSvParams& p) { return (size_t)(p.TH + 2) * p.W * 32 * 4 + (size_t)p.TH * p.OW * 32 * 4 * p.fr; }
static size_t wgsv_smem(const WgSvParams& p) {
    return 2 * (size_t)p.KR * 128 * p.fr + 2 * (size_t)p.xs_rows * 128 + (size_t)p.nraw * wgsv_raw(p) + 128 + 128 * 4 + 1024;
}

static bool wgsv_plan(const dnb_win_t* g, WgSvParams& p) {
    if (g->kh != 3 || g->kw != 3 || g->sy != 1 || g->sx != 1 || g->C != 32) return false;
    if (g->F % 32 || g->F > 128 || g->W % 4 || g->OW % 4) return false;
    if (g->OH != g->H + g->pt + g->pb - 2 || g->OW != g->W + g->pl + g->pr - 2) return false;
    p.B = g->B; p.H = g->H; p.W = g->W; p.F = g->F; p.OH = g->OH; p.OW = g->OW; p.pt = g->pt; p.pl = g->pl;
    p.PW = g->OW + 2;
    p.fr = g->F / 32;
    if (g->W > 256 || g->OW > 256) return false;
    // tallest band (fewest units, least halo) whose double-buffered staging fits
    for (int th = g->OH; th >= 1; --th) {
        p.TH = th; p.nbands = (g->OH + th - 1) / th; p.nraw = 2;
        if (p.nbands > 1 && th * p.nbands - g->OH >= th / 2 + 1) continue;          // badly unbalanced last band
        p.KR = (th * p.PW + 7) & ~7;
        p.xs_rows = (p.KR + 2 * p.PW + 2 + 8 + 7) & ~7;
        if ((th + 2) * p.PW > p.xs_rows || th + 2 > 256) continue;
        if ((th + 2) * g->W * 8 > WG_TBX * WG_TR * 32 || th * g->OW * 8 * p.fr > WG_TBD * WG_TR * 32) continue;
        // chunks fr.. of the A operand (M=128: 4 chunks, M=64: 2) alias the memory behind the dY buffers: it must exist
        const size_t behind = 2 * (size_t)p.xs_rows * 128 + 2 * wgsv_raw(p);
        if ((size_t)3 * p.KR * 128 > behind) continue;
        if (wgsv_smem(p) <= 225 * 1024) {
            while (p.nraw < 4) { ++p.nraw; if (wgsv_smem(p) > 225 * 1024) { --p.nraw; break; } }
            return true;
        }
    }
    return false;
}

}  // namespace tc

using namespace tc;

bool tc_conv_wgrad_sv_supported(const dnb_win_t* g) {
    WgSvParams p{};
    return get_encode_tiled() && !getenv("DNB_NO_WGSV") && wgsv_plan(g, p);
}

int tc_conv_wgrad_sv(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, cudaStream_t st) {
    WgSvParams p{};
    if (!wgsv_plan(g, p)) return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(sv): unsupported shape");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15))
        return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(sv): tensors must be 16-byte aligned");
    p.gw = gw; p.gb = gb; p.inv_m = inv_m;
    p.debug = getenv("DNB_WG_DEBUG") ? atoi(getenv("DNB_WG_DEBUG")) : 0;
    CUtensorMap tx, td;
    p.flat = ((p.TH + 2) * g->W <= 256 && p.TH * g->OW <= 256) && !getenv("DNB_SV_NOFLAT");
    bool ok1, ok2;
    if (p.flat) {
        // a band of a channel is contiguous in NCHW: maps over (H*W, C, B), one contiguous box row per channel
        uint64_t xd[3] = {(uint64_t)g->W * g->H, 32, (uint64_t)g->B};
        uint64_t xs[2] = {(uint64_t)g->W * g->H * 4, (uint64_t)g->W * g->H * 32 * 4};
        uint32_t xb[3] = {(uint32_t)((p.TH + 2) * g->W), 32, 1};
        ok1 = make_tmap_f32(&tx, x, 3, xd, xs, xb, 2);
        uint64_t dd[3] = {(uint64_t)g->OW * g->OH, (uint64_t)g->F, (uint64_t)g->B};
        uint64_t ds[2] = {(uint64_t)g->OW * g->OH * 4, (uint64_t)g->OW * g->OH * g->F * 4};
        uint32_t dbx[3] = {(uint32_t)(p.TH * g->OW), 32, 1};
        ok2 = make_tmap_f32(&td, dy, 3, dd, ds, dbx, 2);
    } else {
        uint64_t xd[4] = {(uint64_t)g->W, (uint64_t)g->H, 32, (uint64_t)g->B};
        uint64_t xs[3] = {(uint64_t)g->W * 4, (uint64_t)g->W * g->H * 4, (uint64_t)g->W * g->H * 32 * 4};
        uint32_t xb[4] = {(uint32_t)g->W, (uint32_t)(p.TH + 2), 32, 1};
        ok1 = make_tmap_f32(&tx, x, 4, xd, xs, xb, 2);
        uint64_t dd[4] = {(uint64_t)g->OW, (uint64_t)g->OH, (uint64_t)g->F, (uint64_t)g->B};
        uint64_t ds[3] = {(uint64_t)g->OW * 4, (uint64_t)g->OW * g->OH * 4, (uint64_t)g->OW * g->OH * g->F * 4};
        uint32_t dbx[4] = {(uint32_t)g->OW, (uint32_t)p.TH, 32, 1};
        ok2 = make_tmap_f32(&td, dy, 4, dd, ds, dbx, 2);
    }
    if (!ok1 || !ok2) return set_err(DNB_ERR_CUDA, "conv2d_bwd_weight(sv): cuTensorMapEncodeTiled failed");
    const size_t smem = wgsv_smem(p);
    const int grid = (int)std::min<long>((long)g->B * p.nbands, kNumSMs);
    if (g->F <= 64 && !getenv("DNB_WGSV_M128")) {
        cudaFuncSetAttribute(tc_conv_wgrad_sv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_conv_wgrad_sv_kernel<true><<<grid, WGSV_THREADS, smem, st>>>(tx, td, p);
    } else {
        cudaFuncSetAttribute(tc_conv_wgrad_sv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_conv_wgrad_sv_kernel<false><<<grid, WGSV_THREADS, smem, st>>>(tx, td, p);
    }
    DNB_LAUNCH_CHECK("tc_conv2d_bwd_weight_sv");
    return DNB_OK;
}

}  // namespace dnb
