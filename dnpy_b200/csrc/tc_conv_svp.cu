// Conv2D(3x3, stride 1, 'same', 32 / 64 input channels, <= 64 filters) -> MaxPool2D -> LeakyRelu/Relu in ONE launch
// (dnpy/layers/convolution.py:193-225, pooling.py:43-99, activations.py:17-19) — conv2 + pool2 + relu2 of the MNIST CNN.
//
// Same staging as tc_conv_sv.cu (each image arrives by ONE 1-D bulk copy, transposer warps turn the raw NCHW planes into a
// zero-ringed NHWC BF16 image in the 64B-swizzled K-major layout, a filter tap is that image read from a shifted start
// row), but with the OPERAND ROLES SWAPPED:  A = the filters of a tap (M = 64 rows), B = the shifted image (N = all padded
// pixel rows of the image, <= 256), D[f][m] in TMEM.  A TMEM lane then holds one filter's WHOLE output image along its
// columns, so an epilogue thread pools its channel in registers — no shared-memory staging of the convolution output (the
// shared-memory/L1 data path is what bounds the un-swapped kernel: DNB_SV_PROF shows its staging stores at ~26 clk each
// under the operand reads of the tensor core), no shuffles, and the convolution output never exists outside TMEM.
// Operand traffic per image: 18 MMAs x (2 KB filters + 5.5 KB pixels) = 137 KB instead of 180 KB; the M = 64 MMAs run at
// half the tensor rate (88 clk each at N = 176), which is the bound: 18 x 88 = 1 584 clk per image.
//
//   warp 0      bulk-copy producer (raw NCHW image, up to 4 in flight)
//   warp 1      filter TMA (once, resident) + MMA issue: 9 taps x channel blocks x 2 k-steps per image, one accumulator
//               stage per image, two stages
//   warps 2-7   transposers (raw fp32 NCHW -> swizzled BF16 NHWC, two image buffers)
//   warps 8-19  epilogue: three groups of four warps (one per TMEM lane quadrant = 16 filters), one per accumulator stage
//               (3 x 168 TMEM columns), take the images round robin;
//               tcgen05.ld of the conv rows a pooled row needs (rolling), + K*bias, np.argmax as max.NaN + first equal,
//               LeakyRelu, results into the group's 8 KB staging buffers that leave by two bulk stores per image
//               (activation output, one byte per pooled element: argmax offset | gate << 4).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {
namespace tc {

struct SvpParams {
    const float* src; const float* bias; float bias_k;
    float* act; unsigned char* idxg; float alpha;
    int B, Cs, Hs, Ws, Cd;
    int off_y, off_x;            // placement of the source inside the padded image (the top / left padding)
    int region_rows;             // rows per 32-channel region of an image buffer (multiple of 8)
    int nraw;                    // raw stages in flight (2..4)
};

constexpr int SVP_TR = 6, SVP_MAXST = 3;                    // transposer warps; accumulator stages = epilogue groups (<= 3)
constexpr int SVP_W_TMA = 0, SVP_W_MMA = 1, SVP_W_TR0 = 2, SVP_W_EPI0 = SVP_W_TR0 + SVP_TR;
constexpr int SVP_THREADS = (2 + SVP_TR + 4 * SVP_MAXST) * 32;
constexpr int SVP_TB = 12;                                   // transposer task slots per thread
static_assert(SVP_W_EPI0 % 4 == 0, "epilogue warp w must own TMEM lane quadrant w % 4");

__device__ __forceinline__ float svp_lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void svp_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void svp_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

__device__ __forceinline__ float svp_max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// One filter's image, pooled: the conv rows arrive from TMEM (16 columns per conv row, row y starts at column y*PW); the
// rows a pooled row adds are requested BEFORE the current pooled row is computed and collected after it (the TMEM round trip
// hides behind the comparisons).  Fast path (NANSAFE = false): m = max.NaN over the window — NaN iff the window holds one —
// and the offset of the first value equal to m; straight-line code.  Returns true if a NaN reached any maximum: the caller
// then repeats the image with the sequential comparisons of np.argmax (first NaN wins and sticks).  Never taken on finite data.
template <int WD, int PK, int PS, int OHp, bool NANSAFE>
__device__ __forceinline__ bool svp_pool_image(uint32_t taddr, float bk, float alpha, bool fvalid, float* __restrict__ sa,
                                               uint8_t* __restrict__ sb) {
    constexpr int PW = WD + 2, OWp = (WD - PK) / PS + 1;
    constexpr int NEW = PS < PK ? PS : PK;                        // conv rows a pooled row adds
    static_assert(PW <= 16, "a conv row must fit one 16-column TMEM load");
    float rows[PK][16];
    uint32_t nxt[NEW][16];
#pragma unroll
    for (int i = 0; i < PK; ++i) svp_tmem_ld16(taddr + (uint32_t)(i * PW), reinterpret_cast<uint32_t (&)[16]>(rows[i]));
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < PK; ++i)
#pragma unroll
        for (int c = 0; c < 16; ++c) rows[i][c] += bk;
    float chk = 0.f;
#pragma unroll
    for (int oy = 0; oy < OHp; ++oy) {
        if (oy + 1 < OHp) {
#pragma unroll
            for (int k = 0; k < NEW; ++k) svp_tmem_ld16(taddr + (uint32_t)(((oy + 1) * PS + PK - NEW + k) * PW), nxt[k]);
        }
#pragma unroll
        for (int ox = 0; ox < OWp; ++ox) {
            float v[PK * PK];
#pragma unroll
            for (int i = 0; i < PK; ++i)
#pragma unroll
                for (int j = 0; j < PK; ++j) v[i * PK + j] = rows[(oy * PS + i) % PK][ox * PS + j];
            float m = v[0];
            int bi = 0;
            if (NANSAFE) {
#pragma unroll
                for (int q = 1; q < PK * PK; ++q)
                    if (v[q] > m || (v[q] != v[q] && m == m)) { m = v[q]; bi = q; }
            } else {
#pragma unroll
                for (int q = 1; q < PK * PK; ++q) m = svp_max_nan(m, v[q]);
                bi = PK * PK - 1;
#pragma unroll
                for (int q = PK * PK - 2; q >= 0; --q) bi = (v[q] == m) ? q : bi;
                chk = svp_max_nan(chk, m);
            }
            const bool pos = m > 0.f;
            if (fvalid) {
                sa[oy * OWp + ox] = pos ? m : alpha * m;          // activations.py:17-19
                sb[oy * OWp + ox] = (uint8_t)(bi | (pos ? 16 : 0));
            }
        }
        if (oy + 1 < OHp) {
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < NEW; ++k) {
                const int y = (oy + 1) * PS + PK - NEW + k;
#pragma unroll
                for (int c = 0; c < 16; ++c) rows[y % PK][c] = __uint_as_float(nxt[k][c]) + bk;
            }
        }
    }
    return chk != chk;
}

template <int CB, int HD, int WD, int PK, int PS>
__global__ void __launch_bounds__(SVP_THREADS, 1)
tc_conv_svp_kernel(const __grid_constant__ CUtensorMap tmW, SvpParams p) {
    constexpr int PW = WD + 2;
    constexpr int NPIX = ((HD * PW + 7) / 8) * 8;             // N of every MMA: all padded pixel rows of the image
    static_assert(NPIX <= 256, "the image must fit one MMA");
    constexpr int OHp = (HD - PK) / PS + 1, OWp = (WD - PK) / PS + 1, OPp = OHp * OWp;
    constexpr int NST = (3 * NPIX <= 512) ? 3 : 2;            // accumulator stages in TMEM = epilogue groups
    constexpr uint32_t TMEM_COLS = (NST * NPIX <= 128) ? 128 : (NST * NPIX <= 256) ? 256 : 512;
    constexpr int ACC_STRIDE = NPIX;                           // columns between the accumulator stages
    constexpr int ROWB = 64, ROWU = 4, KS = 2;                // BF16: 32 channels = 64 bytes per staged row, K = 16 per MMA
    constexpr uint32_t SBO = 512;
    constexpr int W_KB = 64 * ROWB;                            // one K block of filters: 64 rows
    constexpr int NKB = 9 * CB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t region_bytes = (uint32_t)p.region_rows * ROWB;
    const uint32_t buf_bytes = ((region_bytes * CB + 1023) / 1024) * 1024;
    const uint32_t box_bytes = (uint32_t)p.Ws * p.Hs * 32 * 4;                  // one 32-channel raw block
    const uint32_t raw_bytes = ((box_bytes * CB + 127) / 128) * 128;
    const uint32_t outa_bytes = ((uint32_t)p.Cd * OPp * 4 + 127) & ~127u, outb_bytes = ((uint32_t)p.Cd * OPp + 127) & ~127u;
    uint8_t* sW = smem;                                        // [NKB][64 x 64 B]
    uint8_t* sX = sW + (size_t)NKB * W_KB;                     // 2 image buffers x CB regions
    uint8_t* sR = sX + 2 * (size_t)buf_bytes;                  // nraw raw images
    uint8_t* sOA = sR + (size_t)p.nraw * raw_bytes;            // 2 per epilogue group x activation output of an image [Cd][OPp]
    uint8_t* sOB = sOA + 2 * SVP_MAXST * (size_t)outa_bytes;   // 2 per epilogue group x byte tensor of an image
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOB + 2 * SVP_MAXST * (size_t)outb_bytes);
    uint64_t* w_full = bars;
    uint64_t* raw_full = bars + 1;     // [4]
    uint64_t* raw_empty = bars + 5;    // [4]
    uint64_t* x_full = bars + 9;       // [2]
    uint64_t* x_empty = bars + 11;     // [2]
    uint64_t* acc_full = bars + 13;    // [3]
    uint64_t* acc_empty = bars + 16;   // [3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

    // zero both image buffers once (padding ring and slack rows are never written afterwards)
    for (uint32_t e = tid; e < 2 * buf_bytes / 16; e += SVP_THREADS) reinterpret_cast<float4*>(sX)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tma_prefetch_desc(&tmW);
        mbar_init(w_full, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], SVP_TR); }
        for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], SVP_TR); mbar_init(&x_empty[s], 1); }
        for (int s = 0; s < SVP_MAXST; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == SVP_W_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == SVP_W_TMA) {
        if (elect_one()) {
            int it = 0, rs = 0, rph = 0;
            for (int u = blockIdx.x; u < p.B; u += gridDim.x, ++it) {
                if (it >= p.nraw) mbar_wait(&raw_empty[rs], rph ^ 1);
                mbar_arrive_expect_tx(&raw_full[rs], box_bytes * CB);
                bulk_load_1d(sR + (size_t)rs * raw_bytes, p.src + (size_t)u * p.Cs * p.Hs * p.Ws, box_bytes * CB, &raw_full[rs]);
                if (++rs == p.nraw) { rs = 0; rph ^= 1; }
            }
        }
    } else if (warp == SVP_W_MMA) {
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, (uint32_t)(NKB * W_KB));
            for (int kb = 0; kb < NKB; ++kb) tma_load_2d(sW + (size_t)kb * W_KB, &tmW, kb * 32, 0, w_full);   // K order: (tap, channel block)
        }
        __syncwarp();
        mbar_wait(w_full, 0);
        const uint32_t idesc = make_idesc_bf16(64, NPIX, 0, 0);
        const uint64_t w_d0 = make_smem_desc(smem_u32(sW), 16, SBO, LAYOUT_SW64);
        const uint32_t region16 = region_bytes >> 4;
        int it = 0;
        for (int u = blockIdx.x; u < p.B; u += gridDim.x, ++it) {
            const int s = it & 1, as = it % NST, an = it / NST;
            mbar_wait(&x_full[s], (it >> 1) & 1);
            if (an >= 1) mbar_wait(&acc_empty[as], (an - 1) & 1);
            tc_fence_after();
            const uint64_t x_d0 = make_smem_desc(smem_u32(sX + (size_t)s * buf_bytes), 16, SBO, LAYOUT_SW64);
            const uint32_t d_t = tmem_base + (uint32_t)(as * ACC_STRIDE);
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int cb = 0; cb < CB; ++cb) {
                            const int kb = (i * 3 + j) * CB + cb;
                            const uint64_t aw = w_d0 + (uint64_t)(kb * (W_KB >> 4));
                            const uint64_t bx = x_d0 + (uint64_t)((i * PW + j) * ROWU) + (uint64_t)((uint32_t)cb * region16);
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) {
                                if (kb == 0 && ks == 0) mma_bf16_init(d_t, aw, bx, idesc);
                                else mma_bf16_acc(d_t, aw + 2 * ks, bx + 2 * ks, idesc);
                            }
                        }
                mma_commit(&acc_full[as]);
                mma_commit(&x_empty[s]);
            }
            __syncwarp();
        }
    } else if (warp < SVP_W_EPI0) {
        // transposers: task = (pixel of the raw image, 8-channel chunk, channel block); lanes run along the pixels
        const int ttid = tid - SVP_W_TR0 * 32;
        constexpr int LT = SVP_TR * 32;
        const int npix = p.Hs * p.Ws;
        const int tasks = npix * 4 * CB;
        const uint32_t ch_stride = (uint32_t)npix * 4;
        uint32_t t_src[SVP_TB], t_dst[SVP_TB];
#pragma unroll
        for (int r = 0; r < SVP_TB; ++r) {
            const int t = ttid + r * LT;
            t_src[r] = 0; t_dst[r] = 0xffffffffu;
            if (t < tasks) {
                const int pp = t % npix, qc = t / npix;
                const int quad = qc & 3, cb = qc >> 2;
                const int rr = pp / p.Ws, x = pp - rr * p.Ws;
                const int row = (rr + p.off_y) * PW + x + p.off_x;
                t_src[r] = (uint32_t)cb * box_bytes + (uint32_t)(quad * 8) * ch_stride + (uint32_t)pp * 4;
                t_dst[r] = (uint32_t)cb * region_bytes + (uint32_t)row * ROWB + (((quad ^ (row >> 1)) & 3) << 4);
            }
        }
        int it = 0, rs = 0, rph = 0;
        for (int u = blockIdx.x; u < p.B; u += gridDim.x, ++it) {
            const int s = it & 1;
            mbar_wait(&raw_full[rs], rph);
            if (it >= 2) mbar_wait(&x_empty[s], ((it >> 1) - 1) & 1);
            const uint32_t raw_s = smem_u32(sR + (size_t)rs * raw_bytes), buf_s = smem_u32(sX + (size_t)s * buf_bytes);
#pragma unroll
            for (int r0 = 0; r0 < SVP_TB; r0 += 2) {              // 16 loads in flight, then 2 stores
                float v[2][8];
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    if (t_dst[r0 + r] != 0xffffffffu) {
                        const uint32_t a = raw_s + t_src[r0 + r];
#pragma unroll
                        for (int c = 0; c < 8; ++c) v[r][c] = svp_lds32(a + c * ch_stride);
                    }
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    if (t_dst[r0 + r] != 0xffffffffu)
                        sts128u(buf_s + t_dst[r0 + r], pack_bf16x2(v[r][0], v[r][1]), pack_bf16x2(v[r][2], v[r][3]),
                                pack_bf16x2(v[r][4], v[r][5]), pack_bf16x2(v[r][6], v[r][7]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&x_full[s]); mbar_arrive(&raw_empty[rs]); }
            if (++rs == p.nraw) { rs = 0; rph ^= 1; }
        }
    } else {
        // epilogue.  M = 64 accumulator: filter row r lives in TMEM lane (r / 16) * 32 + r % 16 — lanes 0..15 of quadrant r / 16.
        // NST groups of four warps (one per lane quadrant) take the images round robin: group g owns accumulator stage g, its
        // own pair of staging buffers, its own named barrier and its own bulk-store group — the groups never synchronise.
        const int q = warp & 3, grp = (warp - SVP_W_EPI0) >> 2;
        const int f = q * 16 + lane;
        const bool fvalid = lane < 16 && f < p.Cd;
        const float bk = (fvalid && p.bias) ? p.bias_k * __ldg(p.bias + f) : 0.f;
        const bool issuer = q == 0 && lane == 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * ACC_STRIDE);
        for (int it = grp; grp < NST; it += NST) {
            const long u = (long)blockIdx.x + (long)it * gridDim.x;
            if (u >= p.B) break;
            const int n = it / NST, ob = grp * 2 + (n & 1);
            float* sa = reinterpret_cast<float*>(sOA + (size_t)ob * outa_bytes) + (fvalid ? f : 0) * OPp;
            uint8_t* sb = sOB + (size_t)ob * outb_bytes + (fvalid ? f : 0) * OPp;
            mbar_wait(&acc_full[grp], n & 1);
            tc_fence_after();
            if (__any_sync(0xffffffffu, svp_pool_image<WD, PK, PS, OHp, false>(taddr, bk, p.alpha, fvalid, sa, sb)))
                svp_pool_image<WD, PK, PS, OHp, true>(taddr, bk, p.alpha, fvalid, sa, sb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[grp]);
            // staging buffer ob is complete once the group's four warps are here; the group's OTHER buffer (written next) was
            // handed to the store engine one of its images ago: that store must have read it before anybody passes the barrier
            fence_proxy_async_smem();
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
            if (issuer) {
                svp_bulk_s2g(p.act + (size_t)u * p.Cd * OPp, sOA + (size_t)ob * outa_bytes, (uint32_t)p.Cd * OPp * 4u);
                svp_bulk_s2g(p.idxg + (size_t)u * p.Cd * OPp, sOB + (size_t)ob * outb_bytes, (uint32_t)p.Cd * OPp);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SVP_W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t svp_smem_bytes(const SvpParams& p, int cb, int opp) {
    const size_t region = (size_t)p.region_rows * 64;
    const size_t buf = ((region * cb + 1023) / 1024) * 1024;
    const size_t raw = (((size_t)p.Ws * p.Hs * 32 * 4) * cb + 127) / 128 * 128;
    const size_t oa = ((size_t)p.Cd * opp * 4 + 127) & ~(size_t)127, ob = ((size_t)p.Cd * opp + 127) & ~(size_t)127;
    return (size_t)9 * cb * 64 * 64 + 2 * buf + (size_t)p.nraw * raw + 2 * SVP_MAXST * (oa + ob) + 256 + 1024;
}

template <int CB, int HD, int WD, int PK, int PS>
static int svp_launch(const CUtensorMap& tw, SvpParams& p, cudaStream_t st) {
    constexpr int PW = WD + 2, NPIX = ((HD * PW + 7) / 8) * 8;
    constexpr int OPp = ((HD - PK) / PS + 1) * ((WD - PK) / PS + 1);
    p.region_rows = (NPIX + 2 * PW + 2 + 7) & ~7;                     // the last tap starts 2*PW + 2 rows in
    p.nraw = 4;
    while (p.nraw > 2 && svp_smem_bytes(p, CB, OPp) > 225 * 1024) --p.nraw;
    const size_t smem = svp_smem_bytes(p, CB, OPp);
    if (smem > 225 * 1024) return set_err(DNB_ERR_UNSUPPORTED, "conv2d_pool_act_fwd: image does not fit shared memory");
    cudaFuncSetAttribute(tc_conv_svp_kernel<CB, HD, WD, PK, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_conv_svp_kernel<CB, HD, WD, PK, PS><<<std::min(p.B, kNumSMs), SVP_THREADS, smem, st>>>(tw, p);
    DNB_LAUNCH_CHECK("tc_conv2d_pool_act_fwd");
    return DNB_OK;
}

// instantiated geometries: (conv output H, W, pool window, pool stride)
#define SVP_GEOMS(X) X(12, 12, 3, 2) X(12, 12, 2, 2) X(8, 8, 3, 1) X(14, 14, 2, 2)

}  // namespace tc

using namespace tc;

bool tc_conv_svp_supported(const dnb_win_t* g, const dnb_win_t* pg) {
    if (g->kh != 3 || g->kw != 3 || g->sy != 1 || g->sx != 1 || g->pt != 1 || g->pb != 1 || g->pl != 1 || g->pr != 1) return false;
    if ((g->C != 32 && g->C != 64) || g->F > 64 || g->OH != g->H || g->OW != g->W) return false;
    if (((long)g->C * g->H * g->W) % 4) return false;                 // the 1-D bulk copy of an image
    if ((long)g->H * g->W * 4 * (g->C / 32) > (long)SVP_TB * SVP_TR * 32) return false;   // transposer task table
    if (pg->kh != pg->kw || pg->sy != pg->sx || pg->pt || pg->pb || pg->pl || pg->pr) return false;
    if (pg->C != g->F || pg->H != g->OH || pg->W != g->OW) return false;
    if (pg->OH != (g->OH - pg->kh) / pg->sy + 1 || pg->OW != (g->OW - pg->kw) / pg->sx + 1) return false;
    if (((long)g->F * pg->OH * pg->OW) % 16) return false;            // the bulk stores of an image's outputs
    bool ok = false;
#define X(HD, WD, PK, PS) ok = ok || (g->OH == HD && g->OW == WD && pg->kh == PK && pg->sy == PS);
    SVP_GEOMS(X)
#undef X
    return ok;
}

int tc_conv_svp_run(const float* x, const void* wprep, int n_pad, int Kp, const float* bias, float bias_k, float* act,
                    unsigned char* idx_gate, float alpha, const dnb_win_t* g, const dnb_win_t* pg, cudaStream_t st) {
    if (!tc_conv_svp_supported(g, pg)) return set_err(DNB_ERR_UNSUPPORTED, "conv2d_pool_act_fwd: geometry not taken by the fused kernel");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(idx_gate)) & 15)
        return set_err(DNB_ERR_UNSUPPORTED, "conv2d_pool_act_fwd: tensors must be 16-byte aligned");
    SvpParams p{};
    p.src = x; p.bias = bias; p.bias_k = bias_k; p.act = act; p.idxg = idx_gate; p.alpha = alpha;
    p.B = g->B; p.Cs = g->C; p.Hs = g->H; p.Ws = g->W; p.Cd = g->F; p.off_y = g->pt; p.off_x = g->pl;
    CUtensorMap tw;
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)n_pad}, str[1] = {(uint64_t)Kp * 2};
    uint32_t box[2] = {32, 64};
    if (!make_tmap_bf16_sw64(&tw, wprep, 2, dims, str, box)) return set_err(DNB_ERR_CUDA, "conv2d_pool_act_fwd: cuTensorMapEncodeTiled failed");
    const int cb = g->C / 32;
#define X(HD, WD, PK, PS) \
    if (g->OH == HD && g->OW == WD && pg->kh == PK && pg->sy == PS) \
        return cb == 1 ? svp_launch<1, HD, WD, PK, PS>(tw, p, st) : svp_launch<2, HD, WD, PK, PS>(tw, p, st);
    SVP_GEOMS(X)
#undef X
    return set_err(DNB_ERR_UNSUPPORTED, "conv2d_pool_act_fwd: geometry not instantiated");
}

}  // namespace dnb
