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
#include <cstdio>

namespace dnb {
namespace tc {

struct WgSvParams {
    float* gw; float* gb; const float* x; const float* dy;
    int B, H, W, F, OH, OW, pt, pl;
    int PW;                  // padded row pitch in pixels: OW + 2
    int TH, nbands;          // output rows per unit (band), bands per image
    int KR;                  // reduction rows per unit: round_up(TH*PW, 8)
    int xs_rows;             // rows of one staged X buffer (padded image + tap/chunk overhang), multiple of 8
    int fr;                  // 32-filter regions of dY: F / 32
    int nraw;                // raw TMA stages (2..4)
    int flat;                // tensor maps are (H*W, C, B): a band of a channel is one contiguous box row
    int bf;                  // operands staged as BF16: 64-byte pixel rows (64B swizzle, MN-major), K = 16 pixel rows per MMA
    int stage;               // the CTA's gW fits the staging layout [F][292] in shared memory and gw is 16-byte aligned: bulk-reduce epilogue
    int bulk;                // one band per image: x and dY of a unit are two contiguous 1-D bulk copies (halo rows never staged)
    int debug;               // timing experiments (DNB_WG_DEBUG): 1 no epilogue atomics, 2 no MMAs, 4 no transposition
    float inv_m;
    unsigned long long* prof;    // DNB_SV_PROF: per-role clock64 totals of CTA 0
};
#define WG_T0() const long long _t0 = p.prof ? clock64() : 0
#define WG_ACC(var) do { if (p.prof) var += clock64() - _t0; } while (0)

constexpr int WG_TR = 8;                                   // transposer warps
constexpr int WGSV_THREADS = (2 + WG_TR) * 32;
constexpr int WG_TBX = 6, WG_TBD = 10;                     // task slots per transposer thread (x / dY); half of them with BF16 (8 channels per task)

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

// byte offset of 8 consecutive MN elements (16 bytes, chunk o4 of 4) of K row `row` in the 64B swizzle (BF16 operands)
__device__ __forceinline__ uint32_t sw64_off(int row, int o4) {
    return (uint32_t)row * 64 + (((o4 ^ (row >> 1)) & 3) << 4);
}

// M64: F <= 64 -> M=64 MMAs (half the A-operand shared-memory traffic); D row r then lives in TMEM lane (r/16)*32 + r%16
// BF : BF16 operands — 64-byte pixel rows, K = 16 pixel rows per MMA: half the operand traffic and half the MMAs
template <bool M64, bool BF>
__global__ void __launch_bounds__(WGSV_THREADS, 1)
tc_conv_wgrad_sv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, WgSvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int ROWB = BF ? 64 : 128;                                  // bytes of a staged pixel row (32 channels / filters)
    constexpr int KROWS = BF ? 16 : 8;                                   // pixel rows per MMA
    constexpr int CPT = BF ? 8 : 4;                                      // channels per transposer task (one 16-byte chunk)
    constexpr int NCH = 32 / CPT;
    constexpr int TBX = BF ? WG_TBX / 2 : WG_TBX, TBD = BF ? WG_TBD / 2 : WG_TBD;
    const uint32_t dreg_bytes = (uint32_t)p.KR * ROWB;                   // one 32-filter region of a dY buffer
    const uint32_t dbuf_bytes = dreg_bytes * p.fr;
    const uint32_t xbuf_bytes = (uint32_t)p.xs_rows * ROWB;
    const int XH = p.bulk ? p.H : p.TH + 2;                              // x rows of a raw band (with halo unless bulk)
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
    if (p.prof && tid == 0) {
        unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        atomicMin(p.prof + 10, g);
        if (blockIdx.x == 0) p.prof[12] = g;
    }
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
                if (p.bulk) {
                    bulk_load_1d(rs, p.x + (size_t)b * 32 * p.H * p.W, rawx_bytes, &raw_full[s]);
                    bulk_load_1d(rs + rawx_bytes, p.dy + (size_t)b * p.F * p.OH * p.OW, rawd_box * p.fr, &raw_full[s]);
                } else if (p.flat) {
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
        const uint32_t idesc = BF ? make_idesc_bf16(M64 ? 64 : 128, 96, 1, 1) : make_idesc_tf32(M64 ? 64 : 128, 96, 1, 1);
        const int ksteps = p.KR / KROWS;
        const uint32_t pw8 = (uint32_t)p.PW * (ROWB >> 4);
        constexpr uint32_t LAYOUT = BF ? LAYOUT_SW64 : LAYOUT_SW128_BASE32B;
        long long w_xf = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        if (p.prof && blockIdx.x == 0 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.prof[14] = g; }
        for (int n = 0; n < nimg; ++n) {
            const int s = n & 1;
            { WG_T0(); mbar_wait(&xd_full[s], (n >> 1) & 1); WG_ACC(w_xf); }
            tc_fence_after();
            // A: M chunks of 32 filters dreg_bytes apart, 4-row atoms 512 B apart; B: N chunks = the same rows + j
            // (TF32: 8 rows x 128 B per MMA = two 4-row atoms; BF16: 16 rows x 64 B = two 8-row atoms — 1024 B per k step either way)
            const uint64_t a0 = make_smem_desc(smem_u32(sD + (size_t)s * dbuf_bytes), dreg_bytes, 512, LAYOUT);
            const uint64_t b0 = make_smem_desc(smem_u32(sX + (size_t)s * xbuf_bytes), ROWB, 512, LAYOUT);
            if (elect_one()) {
                for (int ks = 0; ks < ((p.debug & 2) ? 0 : ksteps); ++ks) {
                    const uint64_t a = a0 + (uint64_t)((uint32_t)ks * 64);                 // 8 rows = 1024 B
                    const uint64_t bq = b0 + (uint64_t)((uint32_t)ks * 64);
                    const uint32_t acc = (n > 0 || ks > 0) ? 1u : 0u;
                    if constexpr (BF) {
                        mma_bf16(tmem_base, a, bq, idesc, acc);
                        mma_bf16(tmem_base + 96, a, bq + (uint64_t)pw8, idesc, acc);
                        mma_bf16(tmem_base + 192, a, bq + (uint64_t)(2 * pw8), idesc, acc);
                    } else {
                        mma_tf32(tmem_base, a, bq, idesc, acc);
                        mma_tf32(tmem_base + 96, a, bq + (uint64_t)pw8, idesc, acc);
                        mma_tf32(tmem_base + 192, a, bq + (uint64_t)(2 * pw8), idesc, acc);
                    }
                }
                mma_commit(&xd_empty[s]);
                if (n == nimg - 1) mma_commit(acc_full);
            }
            __syncwarp();
        }
        if (p.prof && blockIdx.x == 0 && lane == 0) { p.prof[1] = w_xf; p.prof[3] = clock64() - t_begin;
            unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.prof[15] = g; }
    } else {
        // ===================== transposers =====================
        const int ttid = tid - 64;
        constexpr int LT = WG_TR * 32;
        const int npx_x = XH * p.W, npx_d = p.TH * p.OW;
        const int tasks_x = npx_x * NCH, tasks_d = npx_d * NCH * p.fr;
        const uint32_t chx = (uint32_t)npx_x * 4, chd = (uint32_t)npx_d * 4;        // bytes between channels of the raw images
        uint32_t xs_src[TBX], xs_dst[TBX], ds_src[TBD], ds_dst[TBD];
#pragma unroll
        for (int r = 0; r < TBX; ++r) {
            const int t = ttid + r * LT;
            xs_src[r] = 0; xs_dst[r] = 0xffffffffu;
            if (t < tasks_x) {
                const int pp = t % npx_x, quad = t / npx_x;
                const int y = pp / p.W, x = pp - y * p.W;
                const int row = (y + (p.bulk ? p.pt : 0)) * p.PW + x + p.pl;   // band-local padded row (halo / pad rows: zero filled, or never staged)
                xs_src[r] = (uint32_t)(quad * CPT) * chx + (uint32_t)pp * 4;
                xs_dst[r] = BF ? sw64_off(row, quad) : sw32_off(row, quad);
            }
        }
#pragma unroll
        for (int r = 0; r < TBD; ++r) {
            const int t = ttid + r * LT;
            ds_src[r] = 0; ds_dst[r] = 0xffffffffu;
            if (t < tasks_d) {
                const int pp = t % npx_d, qf = t / npx_d;               // qf: filter chunk 0 .. NCH*fr-1
                const int oy = pp / p.OW, ox = pp - oy * p.OW;
                const int row = oy * p.PW + ox;
                ds_src[r] = (uint32_t)(qf * CPT) * chd + (uint32_t)pp * 4;
                ds_dst[r] = (uint32_t)(qf / NCH) * dreg_bytes + (BF ? sw64_off(row, qf % NCH) : sw32_off(row, qf % NCH));
            }
        }
        float gbs[TBD][CPT];
#pragma unroll
        for (int r = 0; r < TBD; ++r)
#pragma unroll
            for (int c = 0; c < CPT; ++c) gbs[r][c] = 0.f;
        int rs = 0, rph = 0;
        long long w_rf = 0, w_xe = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        for (int n = 0; n < nimg; ++n) {
            const int s = n & 1;
            { WG_T0(); mbar_wait(&raw_full[rs], rph); WG_ACC(w_rf); }
            if (n >= 2) { WG_T0(); mbar_wait(&xd_empty[s], ((n >> 1) - 1) & 1); WG_ACC(w_xe); }
            const uint32_t rx = smem_u32(sR + (size_t)rs * raw_bytes), rd = rx + rawx_bytes;
            const uint32_t xb = smem_u32(sX + (size_t)s * xbuf_bytes), db = smem_u32(sD + (size_t)s * dbuf_bytes);
            if (!(p.debug & 4)) {
            constexpr int GX = 3, GD = 5;                            // tasks per batch (loads in flight before the stores)
#pragma unroll
            for (int r0 = 0; r0 + GX <= TBX; r0 += GX) {
                float v[GX][CPT];
#pragma unroll
                for (int r = 0; r < GX; ++r)
                    if (xs_dst[r0 + r] != 0xffffffffu) {
                        const uint32_t a = rx + xs_src[r0 + r];
#pragma unroll
                        for (int c = 0; c < CPT; ++c) v[r][c] = wg_lds32(a + c * chx);
                    }
#pragma unroll
                for (int r = 0; r < GX; ++r)
                    if (xs_dst[r0 + r] != 0xffffffffu) {
                        if constexpr (BF)
                            sts128u(xb + xs_dst[r0 + r], pack_bf16x2(v[r][0], v[r][1]), pack_bf16x2(v[r][2], v[r][3]),
                                    pack_bf16x2(v[r][4 % CPT], v[r][5 % CPT]), pack_bf16x2(v[r][6 % CPT], v[r][7 % CPT]));
                        else sts128(xb + xs_dst[r0 + r], make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
                    }
            }
#pragma unroll
            for (int r0 = 0; r0 + GD <= TBD; r0 += GD) {
                float v[GD][CPT];
#pragma unroll
                for (int r = 0; r < GD; ++r)
                    if (ds_dst[r0 + r] != 0xffffffffu) {
                        const uint32_t a = rd + ds_src[r0 + r];
#pragma unroll
                        for (int c = 0; c < CPT; ++c) v[r][c] = wg_lds32(a + c * chd);
                    }
#pragma unroll
                for (int r = 0; r < GD; ++r)
                    if (ds_dst[r0 + r] != 0xffffffffu) {
                        if constexpr (BF)
                            sts128u(db + ds_dst[r0 + r], pack_bf16x2(v[r][0], v[r][1]), pack_bf16x2(v[r][2], v[r][3]),
                                    pack_bf16x2(v[r][4 % CPT], v[r][5 % CPT]), pack_bf16x2(v[r][6 % CPT], v[r][7 % CPT]));
                        else sts128(db + ds_dst[r0 + r], make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
#pragma unroll
                        for (int c = 0; c < CPT; ++c) gbs[r0 + r][c] += v[r][c];
                    }
            }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&xd_full[s]); mbar_arrive(&raw_empty[rs]); }
            if (++rs == p.nraw) { rs = 0; rph ^= 1; }
        }
        if (p.prof && blockIdx.x == 0 && ttid == 0) { p.prof[4] = w_rf; p.prof[5] = w_xe; p.prof[6] = clock64() - t_begin; p.prof[9] = nimg; }
        // ---- bias gradient: per-thread partial sums (pixel, filter chunk) -> shared [F][npx_d] -> row sums -> global.
        // (Shared-memory float atomics on the 64 filter addresses serialise 32 ways per warp: tens of microseconds.)
        asm volatile("bar.sync 1, %0;" ::"n"(WG_TR * 32) : "memory");      // every warp is done with the raw stages
        float* part = reinterpret_cast<float*>(sR);
        if (p.gb) {
#pragma unroll
            for (int r = 0; r < TBD; ++r) {
                const int t = ttid + r * LT;
                if (t < tasks_d) {
                    const int pp = t % npx_d, f0 = (t / npx_d) * CPT;
#pragma unroll
                    for (int u = 0; u < CPT; ++u) part[(f0 + u) * npx_d + pp] = gbs[r][u];
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WG_TR * 32) : "memory");
        if (p.gb && nimg > 0) {
            const int tpf = p.F <= 32 ? 8 : (p.F <= 64 ? 4 : 2);            // threads per filter (WG_TR * 32 = 256 threads)
            const int f = ttid / tpf, sub = ttid % tpf;
            float sum = 0.f;
            if (f < p.F)
                for (int pp = sub; pp < npx_d; pp += tpf) sum += part[f * npx_d + pp];
            for (int o = tpf >> 1; o; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
            if (f < p.F && sub == 0) atomicAdd(p.gb + f, sum * p.inv_m);
        }
        // ---- epilogue: D_i[f][(j,c)] -> gW[f][c][i][j].  The CTA's partial gW is staged in shared memory in the gradient's own
        // layout (row pitch 292 floats: 16-byte aligned rows, 2-way bank conflicts) and added to global memory with one bulk
        // reduction (cp.reduce.async.bulk ... add.f32) per filter row instead of 288 scattered atomics per filter.
        if (nimg > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int q = warp & 3, half = (warp - 2) >> 2;           // lane quadrant, column half (144 columns each)
            const int f = M64 ? (lane < 16 ? q * 16 + lane : p.F) : q * 32 + lane;
            constexpr int SP = 292;
            float* stg = reinterpret_cast<float*>(smem);
            if (p.stage) asm volatile("bar.sync 1, %0;" ::"n"(WG_TR * 32) : "memory");   // part[] consumed, all MMAs complete: the staging buffers are free
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
                            if (p.stage) stg[f * SP + (c * 3 + i) * 3 + j] = __uint_as_float(v[u]) * p.inv_m;
                            else atomicAdd(p.gw + (((size_t)f * 32 + c) * 3 + i) * 3 + j, __uint_as_float(v[u]) * p.inv_m);
                        }
                    }
                }
            }
            if (p.stage) {
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(WG_TR * 32) : "memory");
                if (ttid < p.F && !(p.debug & 1)) {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 ::"l"(p.gw + (size_t)ttid * 288), "r"(smem_u32(stg + ttid * SP)), "r"(288 * 4) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.prof && tid == 0) {
        unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        atomicMax(p.prof + 11, g);
        if (blockIdx.x == 0) p.prof[13] = g;
    }
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t wgsv_raw(const WgSvParams& p) { return (size_t)(p.bulk ? p.H : p.TH + 2) * p.W * 32 * 4 + (size_t)p.TH * p.OW * 32 * 4 * p.fr; }
static size_t wgsv_smem(const WgSvParams& p) {
    const size_t rowb = p.bf ? 64 : 128;
    return 2 * (size_t)p.KR * rowb * p.fr + 2 * (size_t)p.xs_rows * rowb + (size_t)p.nraw * wgsv_raw(p) + 128 + 128 * 4 + 1024;
}

static bool wgsv_plan(const dnb_win_t* g, WgSvParams& p, bool bf) {
    p.bf = bf ? 1 : 0;
    const size_t rowb = bf ? 64 : 128;
    const int kq = bf ? 16 : 8;                                 // pixel rows per MMA
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
        p.bulk = p.nbands == 1 && !getenv("DNB_SV_NOBULK");
        if (p.nbands > 1 && th * p.nbands - g->OH >= th / 2 + 1) continue;          // badly unbalanced last band
        p.KR = (th * p.PW + kq - 1) / kq * kq;
        p.xs_rows = (p.KR + 2 * p.PW + 2 + 8 + 7) & ~7;
        if ((th + 2) * p.PW > p.xs_rows || th + 2 > 256) continue;
        if ((p.bulk ? g->H : th + 2) * g->W * 8 > WG_TBX * WG_TR * 32 || th * g->OW * 8 * p.fr > WG_TBD * WG_TR * 32) continue;
        // chunks fr.. of the A operand (M=128: 4 chunks, M=64: 2) alias the memory behind the dY buffers: it must exist
        const size_t behind = 2 * (size_t)p.xs_rows * rowb + 2 * wgsv_raw(p);
        if ((size_t)3 * p.KR * rowb > behind) continue;
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
    return get_encode_tiled() && !getenv("DNB_NO_WGSV") && wgsv_plan(g, p, false);
}

int tc_conv_wgrad_sv(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, bool bf16, cudaStream_t st) {
    WgSvParams p{};
    if (bf16) { WgSvParams q{}; if (getenv("DNB_SV_NOBF16") || !wgsv_plan(g, q, true)) bf16 = false; }
    if (!wgsv_plan(g, p, bf16)) return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(sv): unsupported shape");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15))
        return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(sv): tensors must be 16-byte aligned");
    p.gw = gw; p.gb = gb; p.inv_m = inv_m; p.x = x; p.dy = dy;
    p.stage = (size_t)g->F * 292 * 4 + 2048 <= wgsv_smem(p) && !(reinterpret_cast<uintptr_t>(gw) & 15) && !getenv("DNB_WG_NOSTAGE");
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
    if (getenv("DNB_SV_PROF")) {
        static unsigned long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dprof, 0, 16 * sizeof(unsigned long long), st);
        cudaMemsetAsync(dprof + 10, 0xff, sizeof(unsigned long long), st);
        p.prof = dprof;
    }
    const bool m64 = g->F <= 64 && !getenv("DNB_WGSV_M128");
#define DNB_WGSV(M64_, BF_) do { \
        cudaFuncSetAttribute(tc_conv_wgrad_sv_kernel<M64_, BF_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        tc_conv_wgrad_sv_kernel<M64_, BF_><<<grid, WGSV_THREADS, smem, st>>>(tx, td, p); } while (0)
    if (bf16) { if (m64) DNB_WGSV(true, true); else DNB_WGSV(false, true); }
    else { if (m64) DNB_WGSV(true, false); else DNB_WGSV(false, false); }
#undef DNB_WGSV
    DNB_LAUNCH_CHECK("tc_conv2d_bwd_weight_sv");
    if (p.prof) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[wg prof] mma wait xd_full %llu total %llu | transposer wait raw_full %llu wait xd_empty %llu total %llu | images %llu | "
                        "ns: grid %llu, cta0 %llu (loop starts +%llu, ends +%llu)\n",
                h[1], h[3], h[4], h[5], h[6], h[9], h[11] - h[10], h[13] - h[12], h[14] - h[12], h[15] - h[12]);
    }
    return DNB_OK;
}

}  // namespace dnb
