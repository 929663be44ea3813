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
//   M index  m = oy*PW + ox   (columns ox >= OW / rows oy >= TH are garbage, never stored)
//   A(tap,cb) = region[cb] rows [m0 + i*PW + j, +128)        B(tap,cb) = filters [BN x 32], resident in smem
//
// Column-tap folding (JF, 3x3 kernels, padded pitch PW = 16 or 32, BN <= 64): the kernel is bound by the shared-memory
// traffic of the A operand, which every tap re-reads.  With JF the three column taps j of a kernel row are folded into
// the N extent instead: ONE MMA per (i, 32-channel block, k-step) computes E_j[m][f] = sum_c X[m + i*PW][c] W[f][c][i][j]
// for j = 0,1,2 (N = 3*BN; the three filter blocks are contiguous in shared memory) and the epilogue adds the row-shifted
// blocks y[m] = E_0[m] + E_1[m+1] + E_2[m+2] with two warp shuffles per value.  Because PW divides 32, the rows m whose
// m+1 / m+2 fall into the next TMEM lane quadrant are exactly the garbage columns ox >= PW-2: nothing is lost.
// A 3x less often read A operand; the tail tile of a unit uses M=64.
//
// Persistent, warp-specialised CTA (1 per SM), every stage decoupled by mbarriers:
//   warp 0      TMA producer: the raw NCHW band [32 c][PH rows][W] of a unit per 32-channel block (4-D tiled
//               tensor map, rows outside the image zero-filled by the hardware), double buffered; global
//               latency never touches a register.
//   warp 1      filter TMA (once, resident) + MMA issue: one elected-thread region per tile, the 3x3 case a
//               straight line of tcgen05.mma.kind::tf32 whose descriptors are (tile base + immediate).
//   warps 2-5   transposers: raw NCHW (shared) -> swizzled NHWC (shared): 4 conflict-free LDS + one STS.128
//               per (pixel, channel quad); the padding ring of the image buffers is zeroed once and never written.
//   warps 6-13  epilogue: two warps per TMEM lane quadrant split the BN columns; tcgen05.ld -> +K*bias -> NCHW.
// Two image buffers, two raw buffers and two TMEM accumulator stages keep all four roles busy at once.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstdio>

namespace dnb {
namespace tc {

struct SvParams {
    float* dst; const float* bias; const float* src;
    int B, Cs, Hs, Ws, Cd, Hd, Wd, kh, kw;
    int off_y, off_x;            // placement of the source inside the padded plane
    int TH, nbands;              // output rows per band, bands per image
    int PH, PW;                  // padded band: TH + kh - 1, Wd + kw - 1
    int tiles;                   // M tiles per unit
    int last64;                  // the last tile of a unit is an M=64 MMA (<= 64 rows left)
    int region_rows;             // rows per 32-channel region of an image buffer: tiles*128 + (kh-1)*PW + kw (multiple of 8)
    int nkb;                     // kh*kw*(Cs/32)
    int units;                   // B * nbands (< 2^31, checked by sv_plan)
    int nraw;                    // raw TMA stages in flight (2..4): as many as fit, the loads are latency bound
    int flat;                    // source map is (H*W, C, B): a band of a channel is ONE contiguous row of the box
    int bf;                      // operands staged as BF16: 64-byte rows (64B swizzle), K = 16 per MMA — half the shared-memory
                                 // operand traffic (what these kernels are bound by) and half the MMA issue count
    int nout;                    // 2: the unit's output image [Cd][Hd*Wd] is staged in shared memory (double buffered) and written
                                 // with ONE bulk store per image instead of 4-byte scattered stores (single-band units only); 0: direct
    int bulk;                    // one band per image: the unit's source (all channels) is ONE contiguous 1-D bulk copy, the
                                 // halo rows are never staged (they are part of the zero ring of the image buffers)
    int debug;                   // timing experiments (DNB_SV_DEBUG): 1 no stores, 2 no MMAs, 4 no row-shift shuffles, 8 no transposition
    float bias_k;
    unsigned long long* prof;    // DNB_SV_PROF: per-role clock64 totals of CTA 0 (wait vs work), see tc_conv_sv_run
};
#define SV_T0() const long long _t0 = p.prof ? clock64() : 0
#define SV_ACC(var) do { if (p.prof) var += clock64() - _t0; } while (0)

#ifndef DNB_SV_TR
#define DNB_SV_TR 6
#endif
#ifndef DNB_SV_EPI
#define DNB_SV_EPI 8
#endif
constexpr int SV_TR = DNB_SV_TR, SV_EPI = DNB_SV_EPI;     // transposer / epilogue warps (epilogue: a multiple of 4)
constexpr int SV_W_TMA = 0, SV_W_MMA = 1, SV_W_TR0 = 2, SV_W_EPI0 = SV_W_TR0 + SV_TR;
constexpr int SV_THREADS = (2 + SV_TR + SV_EPI) * 32;
constexpr int SV_TB = ((2048 / (SV_TR * 32)) + 3) & ~3;                                  // transposer task slots per thread (<= 2048 tasks per unit)

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
// NC fp32 columns of this warp's 32 TMEM lanes
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[NC]) {
    static_assert(NC == 8 || NC == 16 || NC == 32, "tcgen05.ld shapes used here");
    if constexpr (NC == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr) : "memory");
    } else if constexpr (NC == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr) : "memory");
    } else {
        tmem_ld_32x32(taddr, v);
    }
}

template <bool BF> __device__ __forceinline__ void sv_mma_init(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    if constexpr (BF) mma_bf16_init(d, a, b, idesc); else mma_tf32_init(d, a, b, idesc);
}
template <bool BF> __device__ __forceinline__ void sv_mma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    if constexpr (BF) mma_bf16_acc(d, a, b, idesc); else mma_tf32_acc(d, a, b, idesc);
}

template <int BN, bool K33, int CB, bool JF, bool BF>
__global__ void __launch_bounds__(SV_THREADS, 1)
tc_conv_sv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, SvParams p) {
    constexpr int ACC = JF ? 3 * BN : BN;                 // TMEM columns of one accumulator stage
    constexpr uint32_t TMEM_COLS = (2 * ACC <= 32) ? 32 : (2 * ACC <= 64) ? 64 : (2 * ACC <= 128) ? 128 : (2 * ACC <= 256) ? 256 : 512;
    static_assert(2 * ACC <= 512, "two accumulator stages must fit TMEM");
    static_assert(!JF || (K33 && CB > 0), "column-tap folding is a 3x3 specialisation");
    constexpr int ROWB = BF ? 64 : 128;                   // bytes of a staged row: 32 channels of one pixel / one filter
    constexpr int ROWU = ROWB >> 4;                       // ... in descriptor address units
    constexpr int KS = BF ? 2 : 4;                        // MMAs per 32-channel block (K = 16 bf16 / 8 tf32 = 32 bytes each)
    constexpr uint32_t SBO = BF ? 512 : 1024, LAYOUT = BF ? LAYOUT_SW64 : LAYOUT_SW128;
    static_assert(!BF || (K33 && CB > 0), "the BF16 variant is instantiated for the 3x3 cases only");
    constexpr int W_KB = BN * ROWB;                       // bytes of one K block of filters
    // epilogue: SV_EPI / 4 warps per TMEM lane quadrant split the BN columns into PARTS slices of CW >= 8 columns
    constexpr int PARTS = (BN / 8 < SV_EPI / 4) ? BN / 8 : SV_EPI / 4;
    constexpr int CW = BN / PARTS;
    constexpr int NCMAX = JF ? 16 : 32;
    constexpr int NC = CW > NCMAX ? NCMAX : CW;           // TMEM columns per epilogue pass
    constexpr int NPASS = CW / NC;                        // passes per epilogue warp
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int cblks = p.Cs >> 5;
    const uint32_t region_bytes = (uint32_t)p.region_rows * ROWB;
    const uint32_t buf_bytes = ((region_bytes * cblks + 1023) / 1024) * 1024;
    const int raw_rows = p.bulk ? p.Hs : p.PH;                                 // rows of a raw band
    const uint32_t box_bytes = (uint32_t)p.Ws * raw_rows * 32 * 4;             // one 32-channel raw block
    const uint32_t raw_bytes = ((box_bytes * cblks + 127) / 128) * 128;
    uint8_t* sW = smem;                                    // [nkb][BN x 128 B]
    uint8_t* sX = sW + (size_t)p.nkb * W_KB;               // 2 buffers x cblks regions (swizzled NHWC)
    uint8_t* sR = sX + 2 * (size_t)buf_bytes;              // nraw raw NCHW bands
    const uint32_t out_bytes = ((uint32_t)p.Cd * p.Hd * p.Wd * 4 + 127) & ~127u;   // one staged output image
    uint8_t* sO = sR + (size_t)p.nraw * raw_bytes;         // nout output images
    uint64_t* bars = reinterpret_cast<uint64_t*>(sO + (size_t)p.nout * out_bytes);
    uint64_t* w_full = bars;
    uint64_t* raw_full = bars + 1;     // [4]
    uint64_t* raw_empty = bars + 5;    // [4]
    uint64_t* x_full = bars + 9;       // [2]
    uint64_t* x_empty = bars + 11;     // [2]
    uint64_t* acc_full = bars + 13;    // [2]
    uint64_t* acc_empty = bars + 15;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    float* s_bias = reinterpret_cast<float*>(bars + 18);   // [max(BN,32)]  K*bias (0 without a bias)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

    // zero both image buffers once (padding ring, slack rows and never-written garbage must stay finite)
    for (uint32_t e = tid; e < 2 * buf_bytes / 16; e += SV_THREADS) reinterpret_cast<float4*>(sX)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int f = tid; f < (BN < 32 ? 32 : BN); f += SV_THREADS) s_bias[f] = (p.bias && f < p.Cd) ? p.bias_k * __ldg(p.bias + f) : 0.f;
    if (tid == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmX);
        mbar_init(w_full, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], SV_TR); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&x_full[s], SV_TR); mbar_init(&x_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], SV_EPI);
        }
        fence_barrier_init();
    }
    if (warp == SV_W_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == SV_W_TMA) {
        // ===================== TMA producer: raw NCHW band of every unit =====================
        if (elect_one()) {
            int it = 0, rs = 0, rph = 0;                              // raw stage and its phase parity
            long long w_re = 0;
            for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
                if (it >= p.nraw) { SV_T0(); mbar_wait(&raw_empty[rs], rph ^ 1); SV_ACC(w_re); }
                const int b = p.nbands == 1 ? u : u / p.nbands, band = p.nbands == 1 ? 0 : u - b * p.nbands;
                const int y0 = band * p.TH - p.off_y;                     // may be negative / run past Hs: zero filled
                mbar_arrive_expect_tx(&raw_full[rs], box_bytes * cblks);
                if (p.bulk) bulk_load_1d(sR + (size_t)rs * raw_bytes, p.src + (size_t)b * p.Cs * p.Hs * p.Ws, box_bytes * cblks, &raw_full[rs]);
                else
                for (int cb = 0; cb < cblks; ++cb) {
                    void* dstp = sR + (size_t)rs * raw_bytes + (size_t)cb * box_bytes;
                    if (p.flat) tma_load_3d(dstp, &tmX, y0 * p.Ws, cb * 32, b, &raw_full[rs]);
                    else tma_load_4d(dstp, &tmX, 0, y0, cb * 32, b, &raw_full[rs]);
                }
                if (++rs == p.nraw) { rs = 0; rph ^= 1; }
            }
            if (p.prof && blockIdx.x == 0) p.prof[0] = w_re;
        }
    } else if (warp == SV_W_MMA) {
        // ===================== filter TMA (once) + MMA issue =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(p.nkb * W_KB));
            // global K order is (tap, channel block); JF keeps the three j blocks of an (i, cb) contiguous: slot (i*CB+cb)*3 + j
            for (int kb = 0; kb < p.nkb; ++kb) {
                int slot = kb;
                if (JF) { const int tap = kb / CB, cb = kb - tap * CB; slot = ((tap / 3) * CB + cb) * 3 + tap % 3; }
                tma_load_2d(sW + (size_t)slot * W_KB, &tmW, kb * 32, 0, w_full);
            }
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        const uint32_t idesc = BF ? make_idesc_bf16(128, ACC, 0, 0) : make_idesc_tf32(128, ACC, 0, 0);
        const uint32_t idesc64 = BF ? make_idesc_bf16(64, ACC, 0, 0) : make_idesc_tf32(64, ACC, 0, 0);
        const uint64_t w_d0 = make_smem_desc(smem_u32(sW), 16, SBO, LAYOUT);
        const uint32_t pw8 = (uint32_t)p.PW * ROWU, region16 = region_bytes >> 4;
        int it = 0, tcount = 0;
        long long w_xf = 0, w_ae = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
            const int s = it & 1;
            { SV_T0(); mbar_wait(&x_full[s], (it >> 1) & 1); SV_ACC(w_xf); }
            tc_fence_after();
            const uint64_t x_d0 = make_smem_desc(smem_u32(sX + (size_t)s * buf_bytes), 16, SBO, LAYOUT);
            for (int t = 0; t < p.tiles; ++t, ++tcount) {
                const int as = tcount & 1;
                if (tcount >= 2) { SV_T0(); mbar_wait(&acc_empty[as], ((tcount >> 1) - 1) & 1); tc_fence_after(); SV_ACC(w_ae); }
                const uint32_t d_t = tmem_base + as * ACC;
                const uint32_t idc = (p.last64 && t == p.tiles - 1) ? idesc64 : idesc;
                // descriptor address fields only ever grow by < 256 KB >> 4: a plain 64-bit add never carries out
                const uint64_t a_t = x_d0 + (uint64_t)((uint32_t)t * (128 * ROWU));
                // ONE elected-thread region per tile (ptxas keeps the operands in uniform registers only when the
                // branch is on the elect itself); inside it the 3x3 case is a straight line of MMAs
                if (elect_one()) {
                    if (p.debug & 2) {
                    } else if (JF) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int cb = 0; cb < CB; ++cb) {
                                const uint64_t a = a_t + (uint64_t)((uint32_t)i * pw8) + (uint64_t)((uint32_t)cb * region16);
                                const uint64_t bw = w_d0 + (uint64_t)((i * CB + cb) * 3 * (W_KB >> 4));
                                if (i == 0 && cb == 0) sv_mma_init<BF>(d_t, a, bw, idc); else sv_mma_acc<BF>(d_t, a, bw, idc);
#pragma unroll
                                for (int ks = 1; ks < KS; ++ks) sv_mma_acc<BF>(d_t, a + 2 * ks, bw + 2 * ks, idc);
                            }
                    } else if (K33 && CB > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const uint64_t a_i = a_t + (uint64_t)((uint32_t)i * pw8);
#pragma unroll
                            for (int j = 0; j < 3; ++j)
#pragma unroll
                                for (int cb = 0; cb < CB; ++cb) {
                                    const int kb = (i * 3 + j) * CB + cb;
                                    const uint64_t a = a_i + (uint64_t)(j * ROWU) + (uint64_t)((uint32_t)cb * region16);
                                    const uint64_t bw = w_d0 + (uint64_t)(kb * (W_KB >> 4));
                                    if (kb == 0) sv_mma_init<BF>(d_t, a, bw, idc); else sv_mma_acc<BF>(d_t, a, bw, idc);
#pragma unroll
                                    for (int ks = 1; ks < KS; ++ks) sv_mma_acc<BF>(d_t, a + 2 * ks, bw + 2 * ks, idc);
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
                                    if (kb == 0) mma_tf32_init(d_t, a, bw, idc); else mma_tf32_acc(d_t, a, bw, idc);
                                    mma_tf32_acc(d_t, a + 2, bw + 2, idc);
                                    mma_tf32_acc(d_t, a + 4, bw + 4, idc);
                                    mma_tf32_acc(d_t, a + 6, bw + 6, idc);
                                }
                            }
                    }
                    mma_commit(&acc_full[as]);
                    if (t == p.tiles - 1) mma_commit(&x_empty[s]);   // fires when every MMA that read this buffer is done
                }
                __syncwarp();
            }
        }
        if (p.prof && blockIdx.x == 0 && lane == 0) { p.prof[1] = w_xf; p.prof[2] = w_ae; p.prof[3] = clock64() - t_begin; }
    } else if (warp < SV_W_EPI0) {
        // ===================== transposers: raw NCHW (shared) -> swizzled NHWC (shared) =====================
        // task = (pixel pp of the PH x Ws band, channel quad, channel block); lanes run along pp: the four LDS of a task
        // are conflict-free, its 16-byte store lands in the swizzled chunk.  The task list of a thread is the same for
        // every unit (fixed geometry): decode it ONCE.
        const int ttid = tid - SV_W_TR0 * 32;
        constexpr int LT = SV_TR * 32;
        constexpr int CPT = BF ? 8 : 4;                      // channels per task: one 16-byte chunk of the staged row
        constexpr int NCH = 32 / CPT;                        // chunks per 32-channel block
        constexpr int GRP = 16 / CPT;                        // tasks per batch: 16 loads in flight, then GRP stores
        const int npix = raw_rows * p.Ws;
        const int tasks = npix * NCH * cblks;
        const uint32_t ch_stride = (uint32_t)npix * 4;       // bytes between channels of the raw band
        uint32_t t_src[SV_TB], t_dst[SV_TB];
#pragma unroll
        for (int r = 0; r < SV_TB; ++r) {
            const int t = ttid + r * LT;
            t_src[r] = 0; t_dst[r] = 0xffffffffu;
            if (t < tasks) {
                const int pp = t % npix, qc = t / npix;
                const int quad = qc % NCH, cb = qc / NCH;
                const int rr = pp / p.Ws, x = pp - rr * p.Ws;
                const int row = (rr + (p.bulk ? p.off_y : 0)) * p.PW + x + p.off_x;     // row of the padded band inside a region
                t_src[r] = (uint32_t)cb * box_bytes + (uint32_t)(quad * CPT) * ch_stride + (uint32_t)pp * 4;
                t_dst[r] = (uint32_t)cb * region_bytes + (uint32_t)row * ROWB
                         + (BF ? (((quad ^ (row >> 1)) & 3) << 4) : (((quad ^ (row & 7)) & 7) << 4));
            }
        }
        int it = 0, rs = 0, rph = 0;
        long long w_rf = 0, w_xe = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
            const int s = it & 1;
            { SV_T0(); mbar_wait(&raw_full[rs], rph); SV_ACC(w_rf); }
            if (it >= 2) { SV_T0(); mbar_wait(&x_empty[s], ((it >> 1) - 1) & 1); SV_ACC(w_xe); }
            const uint32_t raw_s = smem_u32(sR + (size_t)rs * raw_bytes), buf_s = smem_u32(sX + (size_t)s * buf_bytes);
            if (!(p.debug & 8))
#pragma unroll
            for (int r0 = 0; r0 < SV_TB; r0 += GRP) {             // 16 loads in flight, then GRP stores
                float v[GRP][CPT];
#pragma unroll
                for (int r = 0; r < GRP; ++r) {
                    if (t_dst[r0 + r] != 0xffffffffu) {
                        const uint32_t a = raw_s + t_src[r0 + r];
#pragma unroll
                        for (int c = 0; c < CPT; ++c) v[r][c] = lds32(a + c * ch_stride);
                    }
                }
#pragma unroll
                for (int r = 0; r < GRP; ++r)
                    if (t_dst[r0 + r] != 0xffffffffu) {
                        if constexpr (BF)
                            sts128u(buf_s + t_dst[r0 + r], pack_bf16x2(v[r][0], v[r][1]), pack_bf16x2(v[r][2], v[r][3]),
                                    pack_bf16x2(v[r][4 % CPT], v[r][5 % CPT]), pack_bf16x2(v[r][6 % CPT], v[r][7 % CPT]));
                        else sts128(buf_s + t_dst[r0 + r], make_float4(v[r][0], v[r][1], v[r][2], v[r][3]));
                    }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&x_full[s]); mbar_arrive(&raw_empty[rs]); }
            if (++rs == p.nraw) { rs = 0; rph ^= 1; }
        }
        if (p.prof && blockIdx.x == 0 && ttid == 0) { p.prof[4] = w_rf; p.prof[5] = w_xe; p.prof[6] = clock64() - t_begin; }
    } else {
        // ===================== epilogue: TMEM -> NCHW global (8 warps: lane quadrant x column half) =====================
        const int q = warp & 3;                             // TMEM lane quadrant this warp may access
        const int part = (warp - SV_W_EPI0) >> 2;           // which slice of the BN columns (slices >= PARTS idle along)
        const size_t plane_d = (size_t)p.Hd * p.Wd;
        const uint32_t s_bias_u32 = smem_u32(s_bias);
        int tcount = 0, it = 0;
        long long w_af = 0, e_ld = 0, e_sh = 0, e_st = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        const bool staged = p.nout == 2;
        int t_oy[4], t_ox[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool t64 = p.last64 && t == p.tiles - 1;
            const int m = t * 128 + (t64 ? q * 16 + lane : q * 32 + lane);
            t_oy[t] = m / p.PW; t_ox[t] = m - t_oy[t] * p.PW;
        }
        for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
            const int b = p.nbands == 1 ? u : u / p.nbands, band = p.nbands == 1 ? 0 : u - b * p.nbands;
            float* so = reinterpret_cast<float*>(sO + (size_t)(it & 1) * out_bytes);
            // staging buffer (it & 1) was copied out by every epilogue thread before it reached the barrier of unit it-1
            for (int t = 0; t < p.tiles; ++t, ++tcount) {
                const int as = tcount & 1;
                // M=128 tile: D row r = TMEM lane r; M=64 tile: row r = lane (r/16)*32 + r%16 (16 rows per quadrant)
                const bool t64 = p.last64 && t == p.tiles - 1;
                // geometry of the first tiles is unit independent (select chain: no dynamic register indexing)
                int oy = t == 0 ? t_oy[0] : t == 1 ? t_oy[1] : t == 2 ? t_oy[2] : t_oy[3];
                int ox = t == 0 ? t_ox[0] : t == 1 ? t_ox[1] : t == 2 ? t_ox[2] : t_ox[3];
                if (t >= 4) {
                    const int m = t * 128 + (t64 ? q * 16 + lane : q * 32 + lane);
                    oy = m / p.PW; ox = m - oy * p.PW;
                }
                const int y = band * p.TH + oy;
                const bool valid = (!t64 || lane < 16) && oy < p.TH && y < p.Hd && ox < p.Wd;
                float* dp = p.dst + ((size_t)b * p.Cd) * plane_d + (valid ? (size_t)y * p.Wd + ox : 0);
                { SV_T0(); mbar_wait(&acc_full[as], (tcount >> 1) & 1); SV_ACC(w_af); }
                tc_fence_after();
#pragma unroll
                for (int ps = 0; ps < (part < PARTS ? NPASS : 0); ++ps) {
                    long long tq0 = p.prof ? clock64() : 0, tq1 = 0, tq2 = 0;
                    const int c0 = part * CW + ps * NC;
                    const uint32_t tad = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC + c0);
                    uint32_t v[NC];
                    tmem_ld_cols<NC>(tad, v);
                    float bv[NC];                           // K*bias of these channels: broadcast LDS.128 under the TMEM load
#pragma unroll
                    for (int j4 = 0; j4 < NC / 4; ++j4) {
                        const float4 b4 = lds128_ro(s_bias_u32 + (uint32_t)(c0 + 4 * j4) * 4);
                        bv[4 * j4] = b4.x; bv[4 * j4 + 1] = b4.y; bv[4 * j4 + 2] = b4.z; bv[4 * j4 + 3] = b4.w;
                    }
                    if (JF) {
                        // y[m] = E_0[m] + E_1[m+1] + E_2[m+2]: the partner rows sit one / two lanes up (same quadrant for every
                        // row that is stored: PW divides 32, the lanes that would cross are the garbage columns)
                        uint32_t v1[NC], v2[NC];
                        tmem_ld_cols<NC>(tad + BN, v1);
                        tmem_ld_cols<NC>(tad + 2 * BN, v2);
                        tmem_ld_wait();
                        if (p.prof) tq1 = clock64();
                        if (p.debug & 4) {
#pragma unroll
                            for (int j = 0; j < NC; ++j) bv[j] += (__uint_as_float(v[j]) + __uint_as_float(v1[j])) + __uint_as_float(v2[j]);
                        } else {
#pragma unroll
                        for (int j = 0; j < NC; ++j) {
                            const float e1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[j]), 1);
                            const float e2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 2);
                            bv[j] += (__uint_as_float(v[j]) + e1) + e2;
                        }
                        }
                    } else {
                        tmem_ld_wait();
                        if (p.prof) tq1 = clock64();
#pragma unroll
                        for (int j = 0; j < NC; ++j) bv[j] += __uint_as_float(v[j]);
                    }
                    if (p.prof) tq2 = clock64();
                    if (staged) {
                        if (valid) {
                            float* o = so + (size_t)c0 * plane_d + (size_t)y * p.Wd + ox;
#pragma unroll
                            for (int j = 0; j < NC; ++j)
                                if (c0 + NC <= p.Cd || c0 + j < p.Cd) { *o = bv[j]; o += plane_d; }
                        }
                    } else if (valid && !(p.debug & 1)) {
                        float* o = dp + (size_t)c0 * plane_d;
                        if (c0 + NC <= p.Cd) {
#pragma unroll
                            for (int j = 0; j < NC; ++j) { *o = bv[j]; o += plane_d; }
                        } else {
#pragma unroll
                            for (int j = 0; j < NC; ++j)
                                if (c0 + j < p.Cd) { *o = bv[j]; o += plane_d; }
                        }
                    }
                    if (p.prof) { const long long tq3 = clock64(); e_ld += tq1 - tq0; e_sh += tq2 - tq1; e_st += tq3 - tq2; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
            }
            if (staged) {
                // the image [Cd][Hd*Wd] is contiguous in shared AND in global memory: a straight, fully coalesced copy
                // (16 bytes per lane, 512 contiguous bytes per warp instruction) instead of 4-byte stores scattered over
                // the channel planes.  (A bulk store of the same buffer is slower: its read of shared memory completes late.)
                asm volatile("bar.sync 2, %0;" ::"n"(SV_EPI * 32) : "memory");
                if (!(p.debug & 1)) {
                    const int n4 = (int)(p.Cd * plane_d) >> 2;
                    float4* g4 = reinterpret_cast<float4*>(p.dst + (size_t)b * p.Cd * plane_d);
                    const float4* s4 = reinterpret_cast<const float4*>(so);
                    for (int e = tid - SV_W_EPI0 * 32; e < n4; e += SV_EPI * 32) g4[e] = s4[e];
                }
            }
        }
        if (p.prof && blockIdx.x == 0 && lane == 0 && warp == SV_W_EPI0) { p.prof[10] = e_ld; p.prof[11] = e_sh; p.prof[12] = e_st; }
        if (p.prof && blockIdx.x == 0 && lane == 0 && warp == SV_W_EPI0) { p.prof[7] = w_af; p.prof[8] = clock64() - t_begin; p.prof[9] = tcount; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SV_W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t sv_smem_bytes(const SvParams& p, int bn) {
    const int cblks = p.Cs / 32;
    const size_t rowb = p.bf ? 64 : 128;
    const size_t region = (size_t)p.region_rows * rowb;
    const size_t buf = ((region * cblks + 1023) / 1024) * 1024;
    const size_t raw = (((size_t)p.Ws * (p.bulk ? p.Hs : p.PH) * 32 * 4) * cblks + 127) / 128 * 128;
    const size_t outb = p.nout ? (size_t)p.nout * (((size_t)p.Cd * p.Hd * p.Wd * 4 + 127) & ~(size_t)127) : 0;
    return outb + (size_t)p.nkb * bn * rowb + 2 * buf + (size_t)p.nraw * raw + 192 + (size_t)(bn < 32 ? 32 : bn) * 4 + 1024;
}

// column-tap folding applies to 3x3 kernels whose padded row fits a pitch that divides 32, with 3*BN*2 <= 512 TMEM columns
// (BF16 operands with BN >= 64: the un-folded form wins — N = 64 MMAs are cheap at K = 16 and the epilogue loses its
//  3x TMEM reads and 2 shuffles per value, which is what bounds the folded kernel)
static bool sv_jfold(const SvParams& p, int bn) {
    if (p.bf && bn >= 64 && !getenv("DNB_SV_BF_JF")) return false;
    return p.kh == 3 && p.kw == 3 && (p.Cs == 32 || p.Cs == 64) && bn <= 64 && p.Wd + 2 <= 32 && !getenv("DNB_SV_NOJF");
}

// choose the band height: fewest MMA tile-halves per image that fit in 227 KB (ties -> the taller band: less halo)
static bool sv_plan(SvParams& p, int bn) {
    const bool jf = sv_jfold(p, bn);
    p.PW = jf ? (p.Wd + 2 <= 16 ? 16 : 32) : p.Wd + p.kw - 1;
    p.nkb = p.kh * p.kw * (p.Cs / 32);
    const size_t budget = 225 * 1024;
    long best_cost = -1; int best_th = 0;
    auto shape = [&](SvParams& q, int th) {
        q.TH = th; q.nbands = (p.Hd + th - 1) / th; q.PH = th + p.kh - 1;
        const int rows = th * q.PW;
        q.tiles = (rows + 127) / 128;
        // an M=64 tail tile keeps 16 rows per lane quadrant: with folding its rows 14,15 (mod 16) must be garbage columns
        q.last64 = (rows - (q.tiles - 1) * 128) <= 64 && !(jf && q.PW != 16);
        q.region_rows = ((q.tiles - 1) * 128 + (q.last64 ? 64 : 128) + (p.kh - 1) * q.PW + p.kw + 7) & ~7;
        q.nraw = 2;
        q.nout = 0;
        q.bulk = q.nbands == 1 && ((long)p.Cs * p.Hs * p.Ws) % 4 == 0 && !getenv("DNB_SV_NOBULK");
    };
    for (int th = p.Hd; th >= 1; --th) {
        SvParams q = p;
        shape(q, th);
        if (q.PH > 256 || p.Ws > 256) continue;                                            // TMA box limits
        if ((long)(q.bulk ? p.Hs : q.PH) * p.Ws * (p.bf ? 4 : 8) * (p.Cs / 32) > (long)SV_TB * SV_TR * 32) continue;      // transposer task table
        if (sv_smem_bytes(q, bn) > budget) continue;
        const long cost = (long)q.nbands * (2 * q.tiles - q.last64);                       // in units of 64 MMA rows
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_th = th; }
    }
    if (!best_th) return false;
    shape(p, best_th);
    // staged output (bulk store per image): single-band units whose image is a whole number of 16-byte quads, if two
    // staging buffers fit next to at least two raw stages
    // EXPERIMENT, off by default (DNB_SV_STAGE=1).  Measured on B200, conv2 of the CNN (bf16): direct 4-byte stores 113 us
    // forward / 110 us dgrad; staged + coalesced 16-byte copy-out by the epilogue warps 147 / 132 us; staged + cp.async.bulk
    // store 130 / 123 us.  The per-image barrier among the epilogue warps costs more than the scattered stores.
    if (p.nbands == 1 && ((long)p.Cd * p.Hd * p.Wd) % 4 == 0 && getenv("DNB_SV_STAGE")) {
        p.nout = 2;
        if (sv_smem_bytes(p, bn) > budget) p.nout = 0;
    }
    while (p.nraw < 4) { ++p.nraw; if (sv_smem_bytes(p, bn) > budget) { --p.nraw; break; } }
    if ((long)p.B * p.nbands > 0x7fffffffL) return false;
    p.units = p.B * p.nbands;
    return true;
}

static int pick_bn_sv(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }

template <int BN, bool K33, int CB, bool JF, bool BF = false>
static int sv_launch_k(const char* name, const CUtensorMap& tw, const CUtensorMap& tx, const SvParams& p, cudaStream_t st) {
    const size_t smem = sv_smem_bytes(p, BN);
    cudaFuncSetAttribute(tc_conv_sv_kernel<BN, K33, CB, JF, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = std::min(p.units, kNumSMs);
    tc_conv_sv_kernel<BN, K33, CB, JF, BF><<<grid, SV_THREADS, smem, st>>>(tw, tx, p);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

template <int BN>
static int sv_launch(const char* name, const CUtensorMap& tw, const CUtensorMap& tx, const SvParams& p, cudaStream_t st) {
    if constexpr (BN <= 64) {
        if (sv_jfold(p, BN)) {
            if (p.bf) {
                if (p.Cs == 32) return sv_launch_k<BN, true, 1, true, true>(name, tw, tx, p, st);
                return sv_launch_k<BN, true, 2, true, true>(name, tw, tx, p, st);
            }
            if (p.Cs == 32) return sv_launch_k<BN, true, 1, true>(name, tw, tx, p, st);
            return sv_launch_k<BN, true, 2, true>(name, tw, tx, p, st);
        }
    }
    if (p.bf) {
        if (p.Cs == 32) return sv_launch_k<BN, true, 1, false, true>(name, tw, tx, p, st);
        return sv_launch_k<BN, true, 2, false, true>(name, tw, tx, p, st);
    }
    if (p.kh == 3 && p.kw == 3 && p.Cs == 32) return sv_launch_k<BN, true, 1, false>(name, tw, tx, p, st);
    if (p.kh == 3 && p.kw == 3 && p.Cs == 64) return sv_launch_k<BN, true, 2, false>(name, tw, tx, p, st);
    return sv_launch_k<BN, false, 0, false>(name, tw, tx, p, st);
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

static bool sv_geom_ok(const dnb_win_t* g, bool dgrad, SvParams& p);

// BF16 staging halves the transposer task count and the staged bytes: some geometries (wide, 64-channel bands) fit ONLY this way
bool tc_conv_sv_bf16_supported(const dnb_win_t* g, bool dgrad) {
    SvParams p;
    if (getenv("DNB_SV_NOBF16") || !sv_geom_ok(g, dgrad, p)) return false;
    p.bf = 1;
    return p.kh == 3 && p.kw == 3 && (p.Cs == 32 || p.Cs == 64) && sv_plan(p, pick_bn_sv(p.Cd));
}

bool tc_conv_sv_supported(const dnb_win_t* g, bool dgrad) {
    SvParams p;
    return sv_geom_ok(g, dgrad, p) && sv_plan(p, pick_bn_sv(p.Cd));
}

static bool sv_geom_ok(const dnb_win_t* g, bool dgrad, SvParams& p) {
    if (g->sy != 1 || g->sx != 1) return false;
    p = sv_make(g, dgrad);
    if (p.Cs % 32 || p.Cd > 128 || p.off_y < 0 || p.off_x < 0) return false;
    if (p.Ws % 4) return false;                                   // TMA box rows must be a multiple of 16 bytes
    // the produced extent must equal (gathered + pads - k + 1): true for forward by construction; for dgrad it
    // requires that the forward windows tile the padded input exactly (no unused trailing rows/columns)
    if (dgrad && ((g->OH - 1) + g->kh != g->H + g->pt + g->pb || (g->OW - 1) + g->kw != g->W + g->pl + g->pr)) return false;
    // every source column must land inside the padded row (the ring is zeroed once and never rewritten)
    if (p.off_x + p.Ws > p.Wd + p.kw - 1) return false;
    return true;
}

int tc_conv_sv_run(const char* name, const float* src, const void* wprep, int n_pad, int Kp, const float* bias, float bias_k,
                   float* dst, const dnb_win_t* g, bool dgrad, bool bf16, cudaStream_t st) {
    SvParams p = sv_make(g, dgrad);
    p.bf = bf16 ? 1 : 0;
    p.dst = dst; p.bias = bias; p.bias_k = bias_k; p.src = src;
    p.debug = getenv("DNB_SV_DEBUG") ? atoi(getenv("DNB_SV_DEBUG")) : 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15) != 0) return set_err(DNB_ERR_UNSUPPORTED, "%s: source tensor is not 16-byte aligned", name);
    const int bn = pick_bn_sv(p.Cd);
    if (!sv_plan(p, bn)) return set_err(DNB_ERR_UNSUPPORTED, "%s: shape does not fit the shifted-view kernel", name);
    if (reinterpret_cast<uintptr_t>(dst) & 15) p.nout = 0;                     // the bulk store needs a 16-byte aligned image
    CUtensorMap tw, tx;
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)n_pad}, str[1] = {(uint64_t)Kp * (bf16 ? 2 : 4)};
    uint32_t box[2] = {32, (uint32_t)bn};
    if (!(bf16 ? make_tmap_bf16_sw64(&tw, wprep, 2, dims, str, box) : make_tmap_f32(&tw, wprep, 2, dims, str, box)))
        return set_err(DNB_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", name);
    // one box = a PH-row band of 32 channels, unswizzled, zero filled outside the image.  The rows of a band are
    // contiguous in NCHW: when they fit a box row (<= 256 elements) the map is (H*W, C, B) and a channel's band is ONE
    // contiguous box row (TMA moves 48-byte image rows an order of magnitude slower); otherwise (W, H, C, B).
    p.flat = (p.PH * p.Ws <= 256) && !getenv("DNB_SV_NOFLAT");
    bool ok;
    if (p.flat) {
        uint64_t xd[3] = {(uint64_t)p.Ws * p.Hs, (uint64_t)p.Cs, (uint64_t)p.B};
        uint64_t xs[2] = {(uint64_t)p.Ws * p.Hs * 4, (uint64_t)p.Ws * p.Hs * p.Cs * 4};
        uint32_t xb[3] = {(uint32_t)(p.PH * p.Ws), 32, 1};
        ok = make_tmap_f32(&tx, src, 3, xd, xs, xb, 2);
    } else {
        uint64_t xd[4] = {(uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.Cs, (uint64_t)p.B};
        uint64_t xs[3] = {(uint64_t)p.Ws * 4, (uint64_t)p.Ws * p.Hs * 4, (uint64_t)p.Ws * p.Hs * p.Cs * 4};
        uint32_t xb[4] = {(uint32_t)p.Ws, (uint32_t)p.PH, 32, 1};
        ok = make_tmap_f32(&tx, src, 4, xd, xs, xb, 2);
    }
    if (!ok) return set_err(DNB_ERR_CUDA, "%s: cuTensorMapEncodeTiled (source) failed", name);
    if (getenv("DNB_SV_PROF")) {               // debugging aid (synchronises!): where each role of CTA 0 spends its clocks
        static unsigned long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dprof, 0, 16 * sizeof(unsigned long long), st);
        p.prof = dprof;
    }
    int rc;
    switch (bn) {
        case 16: rc = sv_launch<16>(name, tw, tx, p, st); break;
        case 32: rc = sv_launch<32>(name, tw, tx, p, st); break;
        case 64: rc = sv_launch<64>(name, tw, tx, p, st); break;
        default: rc = sv_launch<128>(name, tw, tx, p, st); break;
    }
    if (p.prof && rc == DNB_OK) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[sv prof] %s: tma wait raw_empty %llu | mma wait x_full %llu wait acc_empty %llu total %llu | "
                        "transposer wait raw_full %llu wait x_empty %llu total %llu | epilogue wait acc_full %llu total %llu tiles %llu "
                        "(tmem ld %llu, shuffle+add %llu, store %llu)\n",
                name, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12]);
    }
    return rc;
}

}  // namespace dnb
