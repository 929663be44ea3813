// Tensor-core weight gradient for the single-channel 3x3 stride-1 convolution (conv1 of the MNIST CNN, 1 -> F filters):
//     gW[f][tap] = (1/m) sum_{b,px} dY[b,f,px] * Xp[b, px + tap],     gb[f] = (1/m) sum_{b,px} dY[b,f,px]
// The FP32 register-tiled kernel (conv_thin.cu) spends 10 FMA-pipe instructions per dY element and is bound by FP32
// issue (67% of HBM peak on B200).  Here dY never touches a register:
//   A = dY[b]  [F rows][K = pixels]   K-major, exactly the NCHW layout: one TMA box {32 px, F} per 32-pixel chunk, 128B swizzle
//   B = Xcol   [16 rows][K = pixels]  rows 0-8 the nine shifted copies of the (3 KB) input image, row 9 all ones (its
//                                      column of D is the bias gradient), rows 10-15 zero; built in shared memory by the
//                                      transposer warps from the raw image (one 1-D bulk copy per image)
//   D[f][16]   EIGHT TMEM accumulators (M=64, N=16) used round-robin — back-to-back MMAs into one accumulator serialise on
//              its read-modify-write (~100 clk each at N=16) — persistent over ALL images of the CTA; one tiny epilogue.
// Roles: warp 0 TMA producer, warp 1 MMA issue (tcgen05.mma.kind::tf32), warps 2-9 im2col transposers (+ epilogue).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstdio>

namespace dnb {
namespace tc {

struct ThinWgParams {
    const float* x; const float* dy; float* gw; float* gb;
    int B, H, W, F, OH, OW, pt, pl;
    int npx;                 // OH * OW
    int nchunks;             // 32-pixel K chunks per image
    float inv_m;
    int debug;               // DNB_TW_DEBUG: 1 no MMAs, 2 no dY loads, 4 no im2col
    unsigned long long* prof;    // DNB_SV_PROF: per-role clock64 totals of CTA 0
};
#define TW_T0() const long long _t0 = p.prof ? clock64() : 0
#define TW_ACC(var) do { if (p.prof) var += clock64() - _t0; } while (0)

constexpr int TW_TR = 8;                                   // transposer warps
constexpr int TW_THREADS = (2 + TW_TR) * 32;
constexpr int TW_SLOTS = 6;                                // A ring: dY slots in flight
constexpr int TW_CPS = 4;                                  // 32-pixel chunks (TMA boxes) per slot: one barrier round trip per 128 px
constexpr int TW_TB = 8;                                   // im2col task slots per transposer thread
constexpr int TW_XR = 4;                                   // raw input images in flight: requested TW_XR-1 images ahead, so the
                                                           // load + im2col latency of image n+1 hides behind the MMAs of image n

__global__ void __launch_bounds__(TW_THREADS, 1)
tc_thin_wgrad_kernel(const __grid_constant__ CUtensorMap tmD, ThinWgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t box_bytes = (uint32_t)p.F * 128;                        // one dY chunk: F rows x 32 px
    const uint32_t slot_bytes = TW_CPS * box_bytes;
    const uint32_t xc_bytes = (uint32_t)p.nchunks * 2048;                  // one Xcol buffer: 16 rows x 32 px per chunk
    const uint32_t xr_bytes = ((uint32_t)p.H * p.W * 4 + 127) & ~127u;     // one raw input image
    uint8_t* sA = smem;                                    // TW_SLOTS chunks (+4 KB slack: an M=64 MMA over F=32 rows reads on)
    uint8_t* sXc = sA + (size_t)TW_SLOTS * slot_bytes + 4096;              // 2 Xcol buffers
    uint8_t* sXr = sXc + 2 * (size_t)xc_bytes;                             // TW_XR raw images
    uint64_t* bars = reinterpret_cast<uint64_t*>(sXr + TW_XR * (size_t)xr_bytes);
    uint64_t* a_full = bars;                    // [TW_SLOTS] tx
    uint64_t* a_empty = bars + TW_SLOTS;        // [TW_SLOTS] commit
    uint64_t* xr_full = bars + 2 * TW_SLOTS;    // [TW_XR] tx
    uint64_t* xr_empty = xr_full + TW_XR;       // [TW_XR] TW_TR
    uint64_t* xc_full = xr_full + 2 * TW_XR;    // [2] TW_TR
    uint64_t* xc_empty = xc_full + 2;           // [2] commit
    uint64_t* acc_full = xc_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xc_full + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int nimg = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;       // images blockIdx.x + n*gridDim.x

    // Xcol buffers: rows 0-8 are rewritten per image, row 9 = 1 for real pixels (0 in the K padding), rows 10-15 = 0
    for (uint32_t e = tid; e < 2 * xc_bytes / 16; e += TW_THREADS) {
        const uint32_t off = (e * 16) % xc_bytes;
        const int chunk = off >> 11, row = (off >> 7) & 15, q8 = ((off >> 4) & 7) ^ (row & 7);    // logical 16-byte chunk
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row == 9) {
            const int m0 = chunk * 32 + q8 * 4;
            v = make_float4(m0 < p.npx ? 1.f : 0.f, m0 + 1 < p.npx ? 1.f : 0.f, m0 + 2 < p.npx ? 1.f : 0.f, m0 + 3 < p.npx ? 1.f : 0.f);
        }
        reinterpret_cast<float4*>(sXc)[e] = v;
    }
    // the slack behind the ring is read (rows 32-63 of the last slot when F == 32): keep it finite
    for (uint32_t e = tid; e < 4096 / 16; e += TW_THREADS)
        reinterpret_cast<float4*>(sA + (size_t)TW_SLOTS * slot_bytes)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tma_prefetch_desc(&tmD);
        for (int s = 0; s < TW_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TW_XR; ++s) { mbar_init(&xr_full[s], 1); mbar_init(&xr_empty[s], TW_TR); }
        for (int s = 0; s < 2; ++s) { mbar_init(&xc_full[s], TW_TR); mbar_init(&xc_empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: the raw image, then its dY chunks =====================
        if (elect_one()) {
            int it = 0;
            long long w_ae = 0, w_xre = 0;
            const long long t_begin = p.prof ? clock64() : 0;
            // The 128-byte box rows of a chunk sit 4*npx bytes apart: fetched straight from HBM they open a DRAM page each.
            // Pull every image into L2 as ONE contiguous stream two images ahead; the boxes then hit L2.
            const uint32_t img_bytes = (uint32_t)p.npx * p.F * 4;
            constexpr int PF = 2;
            for (int n = 0; n < min(PF, nimg); ++n)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.dy + (size_t)((int)blockIdx.x + n * (int)gridDim.x) * p.npx * p.F), "r"(img_bytes) : "memory");
            auto issue_raw = [&](int n) {
                const int b = (int)blockIdx.x + n * (int)gridDim.x, buf = n % TW_XR;
                if (n >= TW_XR) { TW_T0(); mbar_wait(&xr_empty[buf], ((n / TW_XR) - 1) & 1); TW_ACC(w_xre); }
                mbar_arrive_expect_tx(&xr_full[buf], (uint32_t)p.H * p.W * 4);
                bulk_load_1d(sXr + (size_t)buf * xr_bytes, p.x + (size_t)b * p.H * p.W, (uint32_t)p.H * p.W * 4, &xr_full[buf]);
            };
            for (int n = 0; n < min(TW_XR - 1, nimg); ++n) issue_raw(n);
            for (int n = 0; n < nimg; ++n) {
                const int b = (int)blockIdx.x + n * (int)gridDim.x;
                if (n + PF < nimg)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.dy + (size_t)(b + PF * (int)gridDim.x) * p.npx * p.F), "r"(img_bytes) : "memory");
                if (n + TW_XR - 1 < nimg) issue_raw(n + TW_XR - 1);
                for (int c = 0; c < p.nchunks; c += TW_CPS, ++it) {
                    const int s = it % TW_SLOTS;
                    if (it >= TW_SLOTS) { TW_T0(); mbar_wait_spin(&a_empty[s], ((it / TW_SLOTS) - 1) & 1); TW_ACC(w_ae); }
                    if (p.debug & 2) { mbar_arrive(&a_full[s]); continue; }
                    mbar_arrive_expect_tx(&a_full[s], slot_bytes);
#pragma unroll
                    for (int k = 0; k < TW_CPS; ++k)          // chunks past npx arrive as zeros (and still count their bytes)
                        tma_load_3d(sA + (size_t)s * slot_bytes + k * box_bytes, &tmD, (c + k) * 32, 0, b, &a_full[s]);
                }
            }
            if (p.prof && blockIdx.x == 0) { p.prof[0] = w_ae; p.prof[1] = w_xre; p.prof[2] = clock64() - t_begin; }
        }
    } else if (warp == 1) {
        // ===================== MMA issue =====================
        const uint32_t idesc = make_idesc_tf32(64, 16, 0, 0);
        int it = 0;
        long long w_xcf = 0, w_af = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        for (int n = 0; n < nimg; ++n) {
            const int buf = n & 1;
            { TW_T0(); mbar_wait(&xc_full[buf], (n >> 1) & 1); TW_ACC(w_xcf); }
            tc_fence_after();
            const uint64_t b0 = make_smem_desc(smem_u32(sXc + (size_t)buf * xc_bytes), 16, 1024);
            for (int c = 0; c < p.nchunks; c += TW_CPS, ++it) {
                const int s = it % TW_SLOTS;
                { TW_T0(); mbar_wait_spin(&a_full[s], (it / TW_SLOTS) & 1); TW_ACC(w_af); }
                tc_fence_after();
                const uint64_t a0 = make_smem_desc(smem_u32(sA + (size_t)s * slot_bytes), 16, 1024);
                if (elect_one()) {
                    if (!(p.debug & 1)) {
#pragma unroll
                        for (int k = 0; k < TW_CPS; ++k) {
                            const uint64_t a = a0 + (uint64_t)(k * (box_bytes >> 4));
                            const uint64_t bq = b0 + (uint64_t)((uint32_t)(c + k) * (2048 >> 4));
                            const uint32_t d0 = tmem_base + (uint32_t)(k & 1) * 64, acc = (it > 0 || k > 1) ? 1u : 0u;
                            mma_tf32(d0, a, bq, idesc, acc);
                            mma_tf32(d0 + 16, a + 2, bq + 2, idesc, acc);
                            mma_tf32(d0 + 32, a + 4, bq + 4, idesc, acc);
                            mma_tf32(d0 + 48, a + 6, bq + 6, idesc, acc);
                        }
                    }
                    mma_commit(&a_empty[s]);
                    if (c + TW_CPS >= p.nchunks) mma_commit(&xc_empty[buf]);
                }
                __syncwarp();
            }
        }
        if (elect_one()) mma_commit(acc_full);
        __syncwarp();
        if (p.prof && blockIdx.x == 0 && lane == 0) { p.prof[3] = w_xcf; p.prof[4] = w_af; p.prof[5] = clock64() - t_begin; }
    } else {
        // ===================== im2col transposers: raw image -> nine shifted rows of Xcol =====================
        const int ttid = tid - 64;
        constexpr int LT = TW_TR * 32;
        const int nquad = (p.npx + 3) >> 2;
        const int tasks = 9 * nquad;
        int t_src[TW_TB][4];
        uint32_t t_dst[TW_TB];
#pragma unroll
        for (int r = 0; r < TW_TB; ++r) {
            const int t = ttid + r * LT;
            t_dst[r] = 0xffffffffu;
#pragma unroll
            for (int e = 0; e < 4; ++e) t_src[r][e] = -1;
            if (t < tasks) {
                const int tap = t / nquad, q = t - tap * nquad;
                const int i = tap / 3, j = tap - i * 3;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int m = 4 * q + e;
                    if (m < p.npx) {
                        const int oy = m / p.OW, ox = m - oy * p.OW;
                        const int iy = oy + i - p.pt, ix = ox + j - p.pl;
                        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) t_src[r][e] = (iy * p.W + ix) * 4;
                    }
                }
                const int m0 = 4 * q;
                t_dst[r] = (uint32_t)(m0 >> 5) * 2048 + (uint32_t)tap * 128 + (((((m0 & 31) >> 2) ^ (tap & 7)) & 7) << 4);
            }
        }
        long long w_xrf = 0, w_xce = 0;
        const long long t_begin = p.prof ? clock64() : 0;
        for (int n = 0; n < nimg; ++n) {
            const int buf = n & 1, rbuf = n % TW_XR;
            { TW_T0(); mbar_wait(&xr_full[rbuf], (n / TW_XR) & 1); TW_ACC(w_xrf); }
            if (n >= 2) { TW_T0(); mbar_wait(&xc_empty[buf], ((n >> 1) - 1) & 1); TW_ACC(w_xce); }
            const uint32_t raw_s = smem_u32(sXr + (size_t)rbuf * xr_bytes), xc_s = smem_u32(sXc + (size_t)buf * xc_bytes);
#pragma unroll
            for (int r = 0; r < TW_TB; ++r) {
                if (t_dst[r] != 0xffffffffu && !(p.debug & 4)) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[e] = 0.f;
                        if (t_src[r][e] >= 0) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[e]) : "r"(raw_s + (uint32_t)t_src[r][e]) : "memory");
                    }
                    sts128(xc_s + t_dst[r], make_float4(v[0], v[1], v[2], v[3]));
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&xc_full[buf]); mbar_arrive(&xr_empty[rbuf]); }
        }
        if (p.prof && blockIdx.x == 0 && ttid == 0) { p.prof[6] = w_xrf; p.prof[7] = w_xce; p.prof[8] = clock64() - t_begin; }
        // ---- epilogue: D[f][tap] -> gW, D[f][9] -> gb (M=64: row r lives in TMEM lane (r/16)*32 + r%16) ----
        if (warp < 6 && nimg > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int nacc = 8;                                       // every slot writes all eight (TW_CPS = 4 chunks)
            float sum[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) sum[k] = 0.f;
            for (int a8 = 0; a8 < nacc; ++a8) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a8 * 16) : "memory");
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 10; ++k) sum[k] += __uint_as_float(v[k]);
            }
            const int f = q * 16 + lane;
            if (lane < 16 && f < p.F) {
#pragma unroll
                for (int k = 0; k < 9; ++k) atomicAdd(p.gw + (size_t)f * 9 + k, sum[k] * p.inv_m);
                if (p.gb) atomicAdd(p.gb + f, sum[9] * p.inv_m);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

static size_t tw_smem(const dnb_win_t* g) {
    const size_t npx = (size_t)g->OH * g->OW, nchunks = ((npx + 31) / 32 + TW_CPS - 1) / TW_CPS * TW_CPS;
    return (size_t)TW_SLOTS * TW_CPS * g->F * 128 + 4096 + 2 * nchunks * 2048 + TW_XR * (((size_t)g->H * g->W * 4 + 127) & ~(size_t)127) +
           (2 * TW_SLOTS + 2 * TW_XR + 8) * 8 + 1024;
}

}  // namespace tc

using namespace tc;

bool tc_thin_wgrad_supported(const dnb_win_t* g) {
    // EXPERIMENT, off by default (DNB_TC_THIN=1 enables it).  Measured on B200 (B=8192, F=32, 26x26): 168 us — the same as
    // the FP32 register-tiled kernel (with one barrier round trip per 4 KB chunk it was 260-310 us: ~350 clk of TMA -> full
    // barrier -> fence -> commit -> empty barrier latency per chunk; 128-pixel slots fixed that).  Debug switches: no MMAs
    // 117 us, bare skeleton 98 us — the 96 M=64/N=16/K=8 MMAs per image cost ~60 clk each and are the bound.  A K=9
    // contraction with 32 filters uses 1/8 of an MMA's rows x columns: too thin to beat FP32 FMAs.
    if (!get_encode_tiled() || !getenv("DNB_TC_THIN")) return false;
    if (g->C != 1 || g->kh != 3 || g->kw != 3 || g->sy != 1 || g->sx != 1) return false;
    if (g->F != 32 && g->F != 64) return false;
    const long npx = (long)g->OH * g->OW;
    if (npx % 4 || (g->H * g->W) % 4 || g->B < 1) return false;               // TMA row pitch / bulk copy granularity
    if (9 * ((npx + 3) / 4) > (long)TW_TB * TW_TR * 32) return false;          // im2col task table
    return tw_smem(g) <= 225 * 1024;
}

int tc_thin_wgrad(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, cudaStream_t st) {
    if (!tc_thin_wgrad_supported(g)) return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(thin tc): unsupported shape");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15))
        return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(thin tc): tensors must be 16-byte aligned");
    // chunks per image rounded up to whole slots: the extra chunks are all-zero K padding on both operands
    ThinWgParams p{x, dy, gw, gb, g->B, g->H, g->W, g->F, g->OH, g->OW, g->pt, g->pl, g->OH * g->OW,
                   ((g->OH * g->OW + 31) / 32 + TW_CPS - 1) / TW_CPS * TW_CPS, inv_m,
                   getenv("DNB_TW_DEBUG") ? atoi(getenv("DNB_TW_DEBUG")) : 0, nullptr};
    if (getenv("DNB_SV_PROF")) {
        static unsigned long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dprof, 0, 16 * sizeof(unsigned long long), st);
        p.prof = dprof;
    }
    CUtensorMap td;
    uint64_t dd[3] = {(uint64_t)p.npx, (uint64_t)g->F, (uint64_t)g->B};
    uint64_t ds[2] = {(uint64_t)p.npx * 4, (uint64_t)p.npx * g->F * 4};
    uint32_t dbx[3] = {32, (uint32_t)g->F, 1};
    if (!make_tmap_f32(&td, dy, 3, dd, ds, dbx, 0)) return set_err(DNB_ERR_CUDA, "conv2d_bwd_weight(thin tc): cuTensorMapEncodeTiled failed");
    const size_t smem = tw_smem(g);
    cudaFuncSetAttribute(tc_thin_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_thin_wgrad_kernel<<<std::min(g->B, kNumSMs), TW_THREADS, smem, st>>>(td, p);
    DNB_LAUNCH_CHECK("tc_thin_conv2d_bwd_weight");
    if (p.prof) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[tw prof] tma: wait a_empty %llu wait xr_empty %llu total %llu | mma: wait xc_full %llu wait a_full %llu total %llu | "
                        "im2col: wait xr_full %llu wait xc_empty %llu total %llu\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
    }
    return DNB_OK;
}

}  // namespace dnb

// =====================================================================================================
// Pointwise (1x1) convolution weight gradient: gW[f][c] = (1/m) sum_{b,px} dY[b,f,px] * X[b,c,px], gb[f] = (1/m) sum dY.
// Both operands are K-major (pixels contiguous) exactly as they lie in NCHW memory: A = dY[b] (F rows) and B = X[b] (C rows)
// arrive as TMA boxes {32 px, rows} with the 128B swizzle — no transposer, no register ever holds an operand.  A constant
// all-ones B tile yields the bias gradient from the same A tiles.  Two TMEM accumulators (gW: N = C columns, gb: N = 16)
// persist over all (image, 64-pixel slot) units of the CTA; one small epilogue of atomics at the end.
//   warp 0: TMA producer (64-pixel slots: two boxes per operand, one barrier)   warp 1: MMA issue   warps 2-5: epilogue
// =====================================================================================================
namespace dnb {
namespace tc {

struct PwWgParams {
    float* gw; float* gb;
    int B, C, F, npx;
    int slots_per_img;           // ceil(npx / 64)
    long units;                  // B * slots_per_img
    float inv_m;
};
constexpr int PW_THREADS = 6 * 32;
constexpr int PW_SLOTS = 6;

template <int M>
__global__ void __launch_bounds__(PW_THREADS, 1)
tc_pointwise_wgrad_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX, PwWgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_blk = (uint32_t)p.F * 128, b_blk = (uint32_t)p.C * 128;       // one 32-pixel block of dY / X
    const uint32_t a_pitch = a_blk > (uint32_t)M * 128 ? a_blk : (uint32_t)M * 128; // an M-row MMA reads M rows: keep them in the slot
    const uint32_t slot_bytes = 2 * a_pitch + 2 * b_blk;
    uint8_t* sOnes = smem;                                  // [16 rows x 32 px] of 1.0f (swizzle invariant)
    uint8_t* sS = smem + 2048;                              // PW_SLOTS slots: A block 0, A block 1, B block 0, B block 1
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + (size_t)PW_SLOTS * slot_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + PW_SLOTS;
    uint64_t* acc_full = bars + 2 * PW_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PW_SLOTS + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const long mine = (p.units - (long)blockIdx.x + (long)gridDim.x - 1) / (long)gridDim.x;
    const uint32_t ncols = (uint32_t)p.C + 16;              // gW columns, then 16 for the bias gradient
    const uint32_t tmem_cols = ncols <= 32 ? 32 : (ncols <= 64 ? 64 : (ncols <= 128 ? 128 : 256));

    for (int e = tid; e < 2048 / 4; e += PW_THREADS) reinterpret_cast<float*>(sOnes)[e] = 1.f;
    // rows F..M-1 of an A block are never written by TMA when F < M: zero them once (they only feed unused D rows)
    for (uint32_t e = tid; e < PW_SLOTS * slot_bytes / 16; e += PW_THREADS) reinterpret_cast<float4*>(sS)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tma_prefetch_desc(&tmD); tma_prefetch_desc(&tmX);
        for (int s = 0; s < PW_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (long n = 0; n < mine; ++n) {
                const int s = (int)(n % PW_SLOTS);
                if (n >= PW_SLOTS) mbar_wait_spin(&empty[s], (uint32_t)((n / PW_SLOTS) - 1) & 1);
                const long u = (long)blockIdx.x + n * gridDim.x;
                const int b = (int)(u / p.slots_per_img), px0 = (int)(u % p.slots_per_img) * 64;
                uint8_t* st = sS + (size_t)s * slot_bytes;
                mbar_arrive_expect_tx(&full[s], 2 * (a_blk + b_blk));
                tma_load_3d(st, &tmD, px0, 0, b, &full[s]);                       // pixels past npx arrive as zeros
                tma_load_3d(st + a_pitch, &tmD, px0 + 32, 0, b, &full[s]);
                tma_load_3d(st + 2 * a_pitch, &tmX, px0, 0, b, &full[s]);
                tma_load_3d(st + 2 * a_pitch + b_blk, &tmX, px0 + 32, 0, b, &full[s]);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_w = make_idesc_tf32(M, p.C, 0, 0), idesc_b = make_idesc_tf32(M, 16, 0, 0);
        const uint64_t ones = make_smem_desc(smem_u32(sOnes), 16, 1024);
        for (long n = 0; n < mine; ++n) {
            const int s = (int)(n % PW_SLOTS);
            mbar_wait_spin(&full[s], (uint32_t)(n / PW_SLOTS) & 1);
            tc_fence_after();
            const uint32_t st = smem_u32(sS + (size_t)s * slot_bytes);
            if (elect_one()) {
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const uint64_t a = make_smem_desc(st + blk * a_pitch, 16, 1024);
                    const uint64_t bx = make_smem_desc(st + 2 * a_pitch + blk * b_blk, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t acc = (n > 0 || blk > 0 || k > 0) ? 1u : 0u;
                        mma_tf32(tmem_base, a + 2 * k, bx + 2 * k, idesc_w, acc);
                        mma_tf32(tmem_base + (uint32_t)p.C, a + 2 * k, ones + 2 * k, idesc_b, acc);
                    }
                }
                mma_commit(&empty[s]);
                if (n == mine - 1) mma_commit(acc_full);
            }
            __syncwarp();
        }
    } else if (mine > 0) {
        // ---- epilogue: D[f][c] -> gW[f][c], D[f][C] -> gb[f] ----
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int f = M == 64 ? (lane < 16 ? q * 16 + lane : p.F) : q * 32 + lane;
        for (int c0 = 0; c0 < (int)ncols; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0) : "memory");
            tmem_ld_wait();
            if (f < p.F) {
                if (c0 < p.C) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(p.gw + (size_t)f * p.C + c0 + j, __uint_as_float(v[j]) * p.inv_m);
                } else if (p.gb) {
                    atomicAdd(p.gb + f, __uint_as_float(v[0]) * p.inv_m);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

static size_t pw_smem(const dnb_win_t* g, int M) {
    const size_t a_pitch = std::max<size_t>((size_t)g->F * 128, (size_t)M * 128);
    return 2048 + (size_t)PW_SLOTS * (2 * a_pitch + 2 * (size_t)g->C * 128) + (2 * PW_SLOTS + 4) * 8 + 1024;
}

}  // namespace tc

using namespace tc;

bool tc_pointwise_wgrad_supported(const dnb_win_t* g) {
    if (!get_encode_tiled() || getenv("DNB_NO_TC_PW")) return false;
    if (g->kh != 1 || g->kw != 1 || g->sy != 1 || g->sx != 1 || g->pt || g->pb || g->pl || g->pr) return false;
    if (g->F != 32 && g->F != 64 && g->F != 128) return false;
    if (g->C % 16 || g->C < 16 || g->C > 128) return false;
    const long npx = (long)g->H * g->W;
    if (npx % 4 || npx < 64) return false;
    return pw_smem(g, g->F <= 64 ? 64 : 128) <= 225 * 1024;
}

int tc_pointwise_wgrad(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, cudaStream_t st) {
    if (!tc_pointwise_wgrad_supported(g)) return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(pointwise tc): unsupported shape");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15))
        return set_err(DNB_ERR_UNSUPPORTED, "conv2d_bwd_weight(pointwise tc): tensors must be 16-byte aligned");
    const int npx = g->H * g->W;
    PwWgParams p{gw, gb, g->B, g->C, g->F, npx, (npx + 63) / 64, (long)g->B * ((npx + 63) / 64), inv_m};
    CUtensorMap td, tx;
    uint64_t dd[3] = {(uint64_t)npx, (uint64_t)g->F, (uint64_t)g->B}, ds[2] = {(uint64_t)npx * 4, (uint64_t)npx * g->F * 4};
    uint32_t dbx[3] = {32, (uint32_t)g->F, 1};
    uint64_t xd[3] = {(uint64_t)npx, (uint64_t)g->C, (uint64_t)g->B}, xs[2] = {(uint64_t)npx * 4, (uint64_t)npx * g->C * 4};
    uint32_t xbx[3] = {32, (uint32_t)g->C, 1};
    if (!make_tmap_f32(&td, dy, 3, dd, ds, dbx, 0) || !make_tmap_f32(&tx, x, 3, xd, xs, xbx, 0))
        return set_err(DNB_ERR_CUDA, "conv2d_bwd_weight(pointwise tc): cuTensorMapEncodeTiled failed");
    const int grid = (int)std::min<long>(p.units, kNumSMs);
    if (g->F <= 64) {
        const size_t smem = pw_smem(g, 64);
        cudaFuncSetAttribute(tc_pointwise_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_pointwise_wgrad_kernel<64><<<grid, PW_THREADS, smem, st>>>(td, tx, p);
    } else {
        const size_t smem = pw_smem(g, 128);
        cudaFuncSetAttribute(tc_pointwise_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_pointwise_wgrad_kernel<128><<<grid, PW_THREADS, smem, st>>>(td, tx, p);
    }
    DNB_LAUNCH_CHECK("tc_pointwise_bwd_weight");
    return DNB_OK;
}

}  // namespace dnb
