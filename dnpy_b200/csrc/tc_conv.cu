// Conv2D / PointwiseConv2D on the 5th-generation tensor cores: implicit GEMM, TF32 operands,
// fp32 accumulation in TMEM.  No im2col / col2im tensor ever exists in global memory.
//
//   forward : Y[p, f]   = sum_k col(X)[p, k] * W[f, k]          M = pixels (b,oy,ox), N = F, K = (c,i,j)
//   dgrad   : dX[p, c]  = sum_k col(dY)[p, k] * W'[c, k]        M = input pixels, N = C, K = (f,i',j')
//             (correlation of dY with the flipped, channel-transposed filter; strided convs gather
//              only the taps that land on the dY lattice)
//   wgrad   : gW^T[k, f] = sum_p col(X)[p, k] * dY[p, f]        M = (c,i,j) (+ a ones row -> gb), N = F, K = pixels
//
// One CTA = 128 rows of M.  The A operand (the im2col view) is GATHERED by four producer warps straight
// from the NCHW fp32 tensor into the canonical K-major 128B-swizzled shared-memory layout (zero padding
// by predication, fp32 kept as-is and consumed as TF32), fenced to the async proxy and handed to the
// MMA thread through an mbarrier.  The B operand (filters / dY) arrives by TMA.  One elected thread
// issues tcgen05.mma.kind::tf32; the producer warps double as epilogue warps (tcgen05.ld, lanes =
// pixels, so NCHW stores are coalesced along x).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_conv.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {
namespace tc {

constexpr int CBM = 128, CBK = 32, CSTAGES = 4;
constexpr int MAX_K = 4608;              // k lookup table capacity (e.g. C=512, 3x3)

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// byte offset of element (row, k) inside a [rows x 32] fp32 K-major tile with 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}

// ---- weight preparation: [N_pad x K_pad] fp32, zero padded, K-major (TMA friendly) ----------------
// K is ordered TAP-MAJOR, CHANNEL-MINOR: k = tap * cp + c with cp = channels rounded up to 4, so every
// 16-byte chunk of the im2col tile (4 consecutive k) belongs to ONE tap: one bounds test per chunk.
// mode 0 (forward): Wp[f][tap*cp + c]  = W[f][c][i][j]
// mode 1 (dgrad)  : Wp[c][tap*fp + f]  = W[f][c][kh-1-i][kw-1-j]
// bf: write BF16 (operands of the BF16 shifted-view kernels) instead of fp32
__global__ void prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int F, int C, int kh, int kw,
                                    int n_pad, int k_pad, int cp, int mode, int bf) {
    const int total = n_pad * k_pad;
    const int khw = kh * kw;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        int n = e / k_pad, k = e % k_pad;
        int tap = k / cp, ch = k % cp;
        float v = 0.f;
        if (tap < khw) {
            int i = tap / kw, j = tap % kw;
            if (mode == 0) {
                if (n < F && ch < C) v = w[(((size_t)n * C + ch) * kh + i) * kw + j];
            } else {
                if (n < C && ch < F) v = w[(((size_t)ch * C + n) * kh + (kh - 1 - i)) * kw + (kw - 1 - j)];
            }
        }
        if (bf) reinterpret_cast<__nv_bfloat16*>(wp)[e] = __float2bfloat16_rn(v);
        else wp[e] = v;
    }
}

struct ConvGemmParams {
    const float* src; float* dst; const float* bias;
    int B, Cs, Hs, Ws;           // gathered tensor
    int Cd, Hd, Wd;              // produced tensor (M = B*Hd*Wd pixels, N = Cd)
    int kh, kw, cp, Kp;          // cp = Cs rounded up to 4; K = kh*kw*cp, Kp = K padded to 32
    int mul_y, mul_x, div_y, div_x, off_y, off_x;
    float bias_k;
};

constexpr int GATHER_WARPS = 8;                                   // producer warps (also epilogue: warps 0-3)
constexpr int GATHER_THREADS = (GATHER_WARPS + 2) * 32;            // + TMA warp + MMA warp

// FAST: the gathered tensor has a multiple of 32 channels, so one 32-wide K block is ONE tap x 32
// channels: a single bounds test per thread per K block, then 16 strided loads and 4 vector stores.
// PLS: compile-time source plane size Hs*Ws (0 = runtime): lets the 16 channel-strided loads of a K block
// use immediate offsets from ONE base address instead of 64-bit address arithmetic per element.
template <int BN, bool STRIDED, bool FAST, int PLS>
__global__ void __launch_bounds__(GATHER_THREADS, 1)
tc_conv_gather_kernel(const __grid_constant__ CUtensorMap tmB, ConvGemmParams p) {
    constexpr int A_BYTES = CBM * CBK * 4, B_BYTES = BN * CBK * 4;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + CSTAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + CSTAGES * B_BYTES);
    uint64_t* empty = full + CSTAGES;
    uint64_t* acc_full = empty + CSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    int* ctab = reinterpret_cast<int*>(tmem_slot + 2);      // [Kp/4] per 16-byte chunk: (i << 24) | (j << 16) | c0, -1 = zero chunk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const long npix = (long)p.B * p.Hd * p.Wd;
    const long m0 = (long)blockIdx.x * CBM;
    const int n0 = blockIdx.y * BN;
    const int nkb = p.Kp / CBK;
    const int khw = p.kh * p.kw;

    for (int ch = tid; ch < p.Kp / 4; ch += GATHER_THREADS) {
        const int k = ch * 4, tap = k / p.cp, c0 = k % p.cp;
        ctab[ch] = (tap < khw) ? (((tap / p.kw) << 24) | ((tap % p.kw) << 16) | c0) : -1;
    }
    if (tid == 0) {
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < CSTAGES; ++s) { mbar_init(&full[s], GATHER_WARPS + 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == GATHER_WARPS) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < GATHER_WARPS) {
        // ===== gather producers: thread = (pixel row, half of the 8 chunks of a 128-byte K row) =====
        const int row = tid & 127, half = tid >> 7;
        const long pix = m0 + row;
        const bool valid = pix < npix;
        int b = 0, y = 0, x = 0;
        if (valid) { x = (int)(pix % p.Wd); long t = pix / p.Wd; y = (int)(t % p.Hd); b = (int)(t / p.Hd); }
        const int y0 = y * p.mul_y - p.off_y, x0 = x * p.mul_x - p.off_x;
        const long plane_s = PLS ? (long)PLS : (long)p.Hs * p.Ws;
        const float* sp = p.src + (long)b * p.Cs * plane_s;
        const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const int rx = row & 7;
        int cb = 0, ti = 0, tj = 0;                     // FAST: K block kb = (tap (ti,tj), 32-channel block cb)
        const int cblks = p.Cs >> 5;
        float tn[16];                                   // FAST: the NEXT K block's 16 values, loaded one block ahead
        auto fast_load = [&](float (&t)[16]) {
            int ty = y0 + ti, tx = x0 + tj;
            bool ok = valid && ty >= 0 && tx >= 0;
            if (STRIDED) {
                ok = ok && (ty % p.div_y) == 0 && (tx % p.div_x) == 0;
                ty /= p.div_y; tx /= p.div_x;
            }
            ok = ok && ty < p.Hs && tx < p.Ws;
            if (ok) {
                const float* g = sp + (long)(cb * 32 + half * 16) * plane_s + (long)ty * p.Ws + tx;
#pragma unroll
                for (int n = 0; n < 16; ++n) t[n] = PLS ? __ldg(g + n * PLS) : __ldg(g + n * plane_s);
            } else {
#pragma unroll
                for (int n = 0; n < 16; ++n) t[n] = 0.f;
            }
            if (++cb == cblks) { cb = 0; if (++tj == p.kw) { tj = 0; ++ti; } }
        };
        if (FAST && nkb > 0) fast_load(tn);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % CSTAGES;
            uint8_t* a = sA + s * A_BYTES + row_off;
            const uint32_t a_s = smem_u32(a);
            if (FAST) {
                float4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = make_float4(tn[4 * q], tn[4 * q + 1], tn[4 * q + 2], tn[4 * q + 3]);
                if (kb + 1 < nkb) fast_load(tn);          // next block's loads fly while this one is stored
                if (kb >= CSTAGES) mbar_wait(&empty[s], ((kb / CSTAGES) - 1) & 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) sts128(a_s + ((((half * 4 + q) ^ rx) & 7) << 4), v[q]);
            } else {
            if (kb >= CSTAGES) mbar_wait(&empty[s], ((kb / CSTAGES) - 1) & 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int kc = half * 4 + q;
                const int e = ctab[kb * 8 + kc];
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && e >= 0) {
                    int ty = y0 + (e >> 24), tx = x0 + ((e >> 16) & 0xff);
                    const int c0 = e & 0xffff;
                    bool ok = ty >= 0 && tx >= 0;
                    if (STRIDED) {
                        ok = ok && (ty % p.div_y) == 0 && (tx % p.div_x) == 0;
                        ty /= p.div_y; tx /= p.div_x;
                    }
                    ok = ok && ty < p.Hs && tx < p.Ws;
                    if (ok) {
                        const float* g = sp + (long)c0 * plane_s + (long)ty * p.Ws + tx;
                        v.x = __ldg(g);
                        if (c0 + 1 < p.Cs) v.y = __ldg(g + plane_s);
                        if (c0 + 2 < p.Cs) v.z = __ldg(g + 2 * plane_s);
                        if (c0 + 3 < p.Cs) v.w = __ldg(g + 3 * plane_s);
                    }
                }
                sts128(a_s + (((kc ^ rx) & 7) << 4), v);
            }
            }
            fence_proxy_async_smem();           // generic-proxy stores -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);   // one arrival per producer warp
        }
        // ===== epilogue: warp w reads TMEM lane quadrant (w & 3); warps 0-3 take the lower half of the
        // columns, warps 4-7 the upper half (pixel rows 0..127 map to lanes, so thread -> pixel (warp&3)*32+lane)
        {
            const int erow = (warp & 3) * 32 + lane;
            const long epix = m0 + erow;
            const bool ev = epix < npix;
            int eb = 0, ey = 0, ex = 0;
            if (ev) { ex = (int)(epix % p.Wd); long t = epix / p.Wd; ey = (int)(t % p.Hd); eb = (int)(t / p.Hd); }
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const size_t plane = (size_t)p.Hd * p.Wd;
            float* dp = p.dst + ((size_t)eb * p.Cd) * plane + (size_t)ey * p.Wd + ex;
            constexpr int NCH = (BN + 31) / 32;               // 32-column chunks
            constexpr int PER = (NCH + 1) / 2;
            const int ch_lo = (warp >> 2) * PER, ch_hi = (warp >> 2) ? NCH : PER;
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                const int c0 = ch * 32;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
                if (ev) {
                    float* o = dp + (size_t)(n0 + c0) * plane;
                    const int nvalid = min(32, min(BN - c0, p.Cd - (n0 + c0)));
                    if (p.bias) {
                        const float* bp = p.bias + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) { *o = fmaf(p.bias_k, __ldg(bp + j), __uint_as_float(v[j])); o += plane; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) { *o = __uint_as_float(v[j]); o += plane; }   // lanes = consecutive pixels -> coalesced
                    }
                }
            }
        }
    } else if (warp == GATHER_WARPS && lane == 0) {
        // ===== TMA producer for the filter tile =====
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % CSTAGES;
            if (kb >= CSTAGES) mbar_wait(&empty[s], ((kb / CSTAGES) - 1) & 1);
            mbar_arrive_expect_tx(&full[s], B_BYTES);
            tma_load_2d(sB + s * B_BYTES, &tmB, kb * CBK, n0, &full[s]);
        }
    } else if (warp == GATHER_WARPS + 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues (43-49 cycles per MMA vs 75 single-threaded) =====
        const uint32_t idesc = make_idesc_tf32(CBM, BN, 0, 0);
        const uint64_t a0 = make_smem_desc(smem_u32(sA), 16, 1024), b0 = make_smem_desc(smem_u32(sB), 16, 1024);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % CSTAGES;
            mbar_wait(&full[s], (kb / CSTAGES) & 1);
            tc_fence_after();
            const uint64_t a = a0 + (uint64_t)(s * (A_BYTES >> 4)), bq = b0 + (uint64_t)(s * (B_BYTES >> 4));
            if (elect_one()) {
                if (kb == 0) mma_tf32_init(tmem_base, a, bq, idesc); else mma_tf32_acc(tmem_base, a, bq, idesc);
                mma_tf32_acc(tmem_base, a + 2, bq + 2, idesc);
                mma_tf32_acc(tmem_base, a + 4, bq + 4, idesc);
                mma_tf32_acc(tmem_base, a + 6, bq + 6, idesc);
                mma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(acc_full);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == GATHER_WARPS) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- weight gradient ---------------------------------------------------------------------------------
// rows r = tap*C + c for r < K (TAP-MAJOR so the per-tap bounds test and pixel offset are hoisted out of
// the channel loop), row K = ones (-> gb), M padded to MT*128; columns = filters; reduction over pixels in
// blocks of 32 inside one image (dY rows are contiguous per (b,f): a 3-D TMA box, zero filled at the image
// edge).  gridDim.x CTAs split the (image, pixel-block) range and accumulate with red.add.
struct WgradParams {
    const float* x; float* gw; float* gb;
    int B, C, H, W, F, OH, OW, kh, kw, sy, sx, pt, pl;
    int K;                        // C*kh*kw
    int pb_per_img;               // ceil(OH*OW / 32)
    long kb_per_cta;
    float inv_m;
};

constexpr int WG_WARPS = 8;
constexpr int WGRAD_THREADS = (WG_WARPS + 2) * 32;

// THINK ("thin K"): at most 32 rows (K+1 <= 32, e.g. one input channel x 3x3).  The A block of a stage is
// then only 32 x 32 floats (the MMA still reads 128 rows: rows >= 32 alias later shared memory, produce
// garbage in TMEM lanes that are never read) which frees shared memory for a 16-deep ring: the kernel is a
// pure dY stream and needs ~64 KB of TMA loads in flight per SM to cover HBM latency.
template <int MT, int BN, bool THINK>
__global__ void __launch_bounds__(WGRAD_THREADS, 1)
tc_conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmD, WgradParams p) {
    constexpr int SUB = THINK ? 2 : 1;                 // 32-pixel sub-blocks per ring stage (amortises the barrier round trip)
    constexpr int ST = THINK ? 8 : ((MT == 1) ? 4 : 3);
    constexpr int A_SUB = THINK ? 32 * CBK * 4 : MT * CBM * CBK * 4, B_SUB = BN * CBK * 4;
    constexpr int A_BYTES = SUB * A_SUB, B_BYTES = SUB * B_SUB;
    constexpr uint32_t TMEM_COLS = (MT * BN <= 32) ? 32 : (MT * BN <= 64) ? 64 : (MT * BN <= 128) ? 128 : (MT * BN <= 256) ? 256 : 512;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + ST * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + ST * B_BYTES);
    uint64_t* empty = full + ST;
    uint64_t* acc_full = empty + ST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int khw = p.kh * p.kw;
    const int npx = p.OH * p.OW;
    const long total_kb = (long)p.B * p.pb_per_img;
    const long kb0 = (long)blockIdx.x * p.kb_per_cta;
    const long kb1 = min(total_kb, kb0 + p.kb_per_cta);
    const int nkb_sub = (int)max(0L, kb1 - kb0);            // 32-pixel blocks of this CTA
    const int nkb = (nkb_sub + SUB - 1) / SUB;               // ring stages
    const int n0 = blockIdx.y * BN;
    constexpr bool thin_k = THINK;
    __shared__ int s_rowc[32], s_rowy[32], s_rowx[32];
    if (thin_k && tid < 32) {
        if (tid < p.K) { const int tap = tid / p.C; s_rowc[tid] = tid % p.C; s_rowy[tid] = tap / p.kw; s_rowx[tid] = tap % p.kw; }
        else { s_rowc[tid] = -1; s_rowy[tid] = 0; s_rowx[tid] = 0; }
    }

    // rows past K+1 are never written by the producers: clear the ring once
    for (int e = tid; e < ST * A_BYTES / 16; e += WGRAD_THREADS) reinterpret_cast<float4*>(sA)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tma_prefetch_desc(&tmD);
        // thin-K mode (<= 32 rows): ONE producer warp fills a whole stage, the warps take stages round-robin
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], (thin_k ? 1 : WG_WARPS) + 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == WG_WARPS) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < WG_WARPS) {
        // ===== gather producers: lane = pixel of the 32-pixel block; warp w owns channels c = w, w+8, ... =====
        const size_t plane = (size_t)p.H * p.W;
        if (thin_k) {
            // tiny reductions (e.g. 1 input channel x 3x3 = 9 rows + the ones row): the im2col block of a stage is
            // only (K+1) x 32 floats, the kernel is a pure dY stream through TMA — keep every warp busy on its own stage
            // stage i -> producer warp i % 8: since ST % 8 == 0 a ring slot is always refilled by the same warp,
            // which therefore sees its fills in order (the parity wait can never alias an older phase)
            for (int i = warp; i < nkb; i += WG_WARPS) {
                const int s = i % ST;
                if (i >= ST) mbar_wait(&empty[s], ((i / ST) - 1) & 1);
#pragma unroll 1
                for (int q = 0; q < SUB; ++q) {
                    const long kb = kb0 + (long)i * SUB + q;
                    if (kb >= kb1) break;              // tail: dY is zero filled by TMA, stale (finite) rows contribute 0
                    const int b = (int)(kb / p.pb_per_img);
                    const int pix = (int)(kb % p.pb_per_img) * 32 + lane;
                    const bool pv = pix < npx;
                    const int oy = pv ? pix / p.OW : 0, ox = pv ? pix % p.OW : 0;
                    const int y0 = oy * p.sy - p.pt, x0 = ox * p.sx - p.pl;
                    const float* xb = p.x + (size_t)b * p.C * plane;
                    const uint32_t a_s = smem_u32(sA + s * A_BYTES + q * A_SUB);
                    for (int r0 = 0; r0 <= p.K; r0 += 8) {          // batches of 8 rows: loads first, then stores
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int r = r0 + u;
                            v[u] = (r == p.K) ? 1.0f : 0.f;
                            if (r < p.K) {
                                const int iy = y0 + s_rowy[r], ix = x0 + s_rowx[r];
                                if (pv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v[u] = __ldg(xb + ((size_t)s_rowc[r] * p.H + iy) * p.W + ix);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (r0 + u <= p.K) sts32(a_s + sw128_off(r0 + u, lane), v[u]);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        } else
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            const long kb = kb0 + i;
            const int b = (int)(kb / p.pb_per_img);
            const int pix = (int)(kb % p.pb_per_img) * 32 + lane;
            const bool pv = pix < npx;
            const int oy = pv ? pix / p.OW : 0, ox = pv ? pix % p.OW : 0;
            const int y0 = oy * p.sy - p.pt, x0 = ox * p.sx - p.pl;
            const float* xb = p.x + (size_t)b * p.C * plane;
            if (i >= ST) mbar_wait(&empty[s], ((i / ST) - 1) & 1);
            uint8_t* a = sA + s * A_BYTES;
            if ((p.C & 7) == 0) {
                // C % 8 == 0: warp w owns channels c = w (mod 8) of EVERY tap, so (row & 7) == w is constant:
                // the bounds test / pixel offset are hoisted per tap and the swizzled address just steps by 1024 B
                const uint32_t coff = (uint32_t)(warp * 128 + ((((lane >> 2) ^ warp) & 7) << 4) + ((lane & 3) << 2));
                const uint32_t a_s = smem_u32(a) + coff;
                const int nch = p.C >> 3;                 // channels of this warp per tap
                if (khw == 9 && nch <= 4) {
                    // 3x3 window, <= 32 channels: issue every load of the stage (9 taps x nch channels) before
                    // the first store — one memory latency per stage instead of nine
                    float v[9][4];
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int iy = y0 + tap / 3, ix = x0 + tap % 3;
                        const bool ok = pv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                        const float* g = xb + ((size_t)warp * p.H + iy) * p.W + ix;
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[tap][u] = (ok && u < nch) ? __ldg(g + (size_t)u * 8 * plane) : 0.f;
                    }
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t ar = a_s + (uint32_t)(((tap * p.C + warp) >> 3) * 1024);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (u < nch) sts32(ar + u * 1024, v[tap][u]);
                    }
                } else {
                int ti = 0, tj = 0;
                for (int tap = 0; tap < khw; ++tap) {
                    const int iy = y0 + ti, ix = x0 + tj;
                    const bool ok = pv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    const float* g = xb + ((size_t)warp * p.H + iy) * p.W + ix;
                    const uint32_t ar = a_s + (uint32_t)(((tap * p.C + warp) >> 3) * 1024);
                    for (int c0 = 0; c0 < nch; c0 += 4) {     // batches of 4 independent loads
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = (ok && c0 + u < nch) ? __ldg(g + (size_t)(c0 + u) * 8 * plane) : 0.f;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c0 + u < nch) sts32(ar + (c0 + u) * 1024, v[u]);
                    }
                    if (++tj == p.kw) { tj = 0; ++ti; }
                }
                }
            } else {
            // rows r = tap*C + c are dealt round-robin to the producer warps (balanced for any C)
            int tap = warp / p.C, c = warp % p.C;
#pragma unroll 4
            for (int r = warp; r < p.K; r += WG_WARPS) {
                const int iy = y0 + tap / p.kw, ix = x0 + tap % p.kw;
                const bool ok = pv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const float val = ok ? __ldg(xb + ((size_t)c * p.H + iy) * p.W + ix) : 0.f;
                *reinterpret_cast<float*>(a + (r >> 7) * (CBM * CBK * 4) + sw128_off(r & 127, lane)) = val;
                c += WG_WARPS;
                while (c >= p.C) { c -= p.C; ++tap; }
            }
            }
            if (warp == WG_WARPS - 1) {          // ones row: D[K, f] = sum_p dY[p, f] (dY is zero filled past the image)
                const int r = p.K;
                *reinterpret_cast<float*>(a + (r >> 7) * (CBM * CBK * 4) + sw128_off(r & 127, lane)) = 1.0f;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        // ===== epilogue (warps 0-3): atomically add the partial gW^T tile =====
        if (warp < 4 && nkb > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int r = mt * CBM + warp * 32 + lane;
                const int tap = r / p.C, c = r % p.C;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * BN + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int f = n0 + c0 + j;
                        if (c0 + j < BN && f < p.F) {
                            const float g = __uint_as_float(v[j]) * p.inv_m;
                            if (r < p.K) atomicAdd(p.gw + ((size_t)f * p.C + c) * khw + tap, g);
                            else if (r == p.K && p.gb) atomicAdd(p.gb + f, g);
                        }
                    }
                }
            }
        }
    } else if (warp == WG_WARPS && lane == 0) {
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            const long kb = kb0 + i;
            if (i >= ST) mbar_wait(&empty[s], ((i / ST) - 1) & 1);
            mbar_arrive_expect_tx(&full[s], B_BYTES);
#pragma unroll
            for (int q = 0; q < SUB; ++q) {
                const long kq = kb0 + (long)i * SUB + q;
                // past this CTA's range the box is aimed outside the tensor (image index B): TMA zero-fills it
                const int bi = kq < kb1 ? (int)(kq / p.pb_per_img) : p.B;
                tma_load_3d(sB + s * B_BYTES + q * B_SUB, &tmD, (int)(kq % p.pb_per_img) * 32, n0, bi, &full[s]);
            }
        }
    } else if (warp == WG_WARPS + 1) {
        // MMA issuer: warp-uniform loop, one elected lane issues
        const uint32_t idesc = make_idesc_tf32(CBM, BN, 0, 0);
        const uint64_t a0 = make_smem_desc(smem_u32(sA), 16, 1024), b0 = make_smem_desc(smem_u32(sB), 16, 1024);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % ST;
            mbar_wait(&full[s], (i / ST) & 1);
            tc_fence_after();
            const uint64_t a = a0 + (uint64_t)(s * (A_BYTES >> 4)), bq = b0 + (uint64_t)(s * (B_BYTES >> 4));
            if (elect_one()) {
#pragma unroll
                for (int q = 0; q < SUB; ++q)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int j = 0; j < CBK / 8; ++j) {
                            const uint64_t ad = a + (uint64_t)((q * A_SUB + mt * (CBM * CBK * 4)) >> 4) + 2 * j;
                            const uint64_t bd = bq + (uint64_t)((q * B_SUB) >> 4) + 2 * j;
                            if (i == 0 && q == 0 && j == 0) mma_tf32_init(tmem_base + mt * BN, ad, bd, idesc);
                            else mma_tf32_acc(tmem_base + mt * BN, ad, bd, idesc);
                        }
                mma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(acc_full);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_WARPS) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int pick_bn(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }

template <int BN, bool STRIDED, bool FAST, int PLS>
static int launch_gather(const char* name, const CUtensorMap& tb, const ConvGemmParams& p, cudaStream_t st) {
    const size_t smem = CSTAGES * (CBM * CBK * 4 + BN * CBK * 4) + 128 + (size_t)p.Kp + 1024;
    cudaFuncSetAttribute(tc_conv_gather_kernel<BN, STRIDED, FAST, PLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long npix = (long)p.B * p.Hd * p.Wd;
    dim3 grid((unsigned)((npix + CBM - 1) / CBM), (p.Cd + BN - 1) / BN);
    tc_conv_gather_kernel<BN, STRIDED, FAST, PLS><<<grid, GATHER_THREADS, smem, st>>>(tb, p);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

static int run_gather(const char* name, const float* wprep, int n_pad, ConvGemmParams p, cudaStream_t st) {
    const int bn = pick_bn(p.Cd);
    CUtensorMap tb;
    uint64_t dims[2] = {(uint64_t)p.Kp, (uint64_t)n_pad}, str[1] = {(uint64_t)p.Kp * 4};
    uint32_t box[2] = {CBK, (uint32_t)bn};
    if (!make_tmap_f32(&tb, wprep, 2, dims, str, box)) return set_err(DNB_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", name);
    const bool strided = (p.div_y != 1 || p.div_x != 1);
    const bool fast = (p.Cs % 32) == 0;
    const int pls = p.Hs * p.Ws;
    // fast path instantiated for the plane sizes of the BASELINE configs (12x12, 16x16, 32x32, 64x64); others run PLS=0
#define DNB_FAST(BN_, ST_) (pls == 144 ? launch_gather<BN_, ST_, true, 144>(name, tb, p, st) \
                          : pls == 256 ? launch_gather<BN_, ST_, true, 256>(name, tb, p, st) \
                          : pls == 1024 ? launch_gather<BN_, ST_, true, 1024>(name, tb, p, st) \
                          : pls == 4096 ? launch_gather<BN_, ST_, true, 4096>(name, tb, p, st) \
                          : launch_gather<BN_, ST_, true, 0>(name, tb, p, st))
#define DNB_GO(BN_) (strided ? (fast ? DNB_FAST(BN_, true) : launch_gather<BN_, true, false, 0>(name, tb, p, st)) \
                             : (fast ? DNB_FAST(BN_, false) : launch_gather<BN_, false, false, 0>(name, tb, p, st)))
    switch (bn) {
        case 16: return DNB_GO(16);
        case 32: return DNB_GO(32);
        case 64: return DNB_GO(64);
        default: return DNB_GO(128);
    }
#undef DNB_FAST
#undef DNB_GO
}

}  // namespace tc

using namespace tc;

// workspace: [fwd filters: pad16(F) x pad32(C*khw)] [dgrad filters: pad16(C) x pad32(F*khw)]
static size_t ws_fwd_floats(const dnb_win_t* g) { return (size_t)round_up(g->F, 128) * round_up(round_up(g->C, 4) * g->kh * g->kw, CBK); }
static size_t ws_dgrad_floats(const dnb_win_t* g) { return (size_t)round_up(g->C, 128) * round_up(round_up(g->F, 4) * g->kh * g->kw, CBK); }

// Data gradient of a STRIDED convolution through the stride-1 shifted-view kernel: dY is scattered into a zero-interleaved
// ("dilated") buffer dYu[s*oy][s*ox] = dY[oy][ox] and dX = corr(dYu, flip(W)) is the data gradient of a fictitious stride-1
// convolution with the same top/left pads.  One extra streaming pass (read dY, write s^2 x as much) instead of the gather
// kernel's predicated taps (3/4 of the MMA work of a stride-2 layer multiplies zeros there: 0.07 of HBM peak on B200).
// stride-2 specialisation (OW1 == 2*OW, OW % 4 == 0): one LDG.128 of dY feeds two STG.128 of the dilated row; odd rows are zero
__global__ void __launch_bounds__(256) dilate2_dy_kernel(const float* __restrict__ dy, float* __restrict__ du, long planes,
                                                         int OH, int OW, int OH1, int sy) {
    const int q = OW >> 2;                                   // input quads per row
    const long total = planes * OH1 * q;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int xq = (int)(i % q); long t = i / q;
        const int y = (int)(t % OH1); const long pl = t / OH1;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y % sy == 0 && y / sy < OH) d = __ldg(reinterpret_cast<const float4*>(dy + (pl * OH + y / sy) * OW) + xq);
        float4* o = reinterpret_cast<float4*>(du + (pl * OH1 + y) * (long)(2 * OW)) + 2 * xq;
        o[0] = make_float4(d.x, 0.f, d.y, 0.f);
        o[1] = make_float4(d.z, 0.f, d.w, 0.f);
    }
}

static bool dgrad_dilate_geom(const dnb_win_t* g, dnb_win_t* g1, bool bf16) {
    if ((g->sy == 1 && g->sx == 1) || getenv("DNB_NO_DILATE")) return false;
    *g1 = *g;
    g1->sy = 1; g1->sx = 1;
    g1->OH = std::max((g->OH - 1) * g->sy + 1, g->H + g->pt - g->kh + 1);
    g1->OW = (std::max((g->OW - 1) * g->sx + 1, g->W + g->pl - g->kw + 1) + 3) & ~3;     // source rows in whole 16-byte quads
    g1->pb = g1->OH - 1 + g->kh - g->H - g->pt;
    g1->pr = g1->OW - 1 + g->kw - g->W - g->pl;
    if (g1->pb < 0 || g1->pr < 0) return false;
    return tc_conv_sv_supported(g1, true) || (bf16 && tc_conv_sv_bf16_supported(g1, true));
}
static size_t ws_weights_bytes(const dnb_win_t* g) { return ((ws_fwd_floats(g) + ws_dgrad_floats(g)) * sizeof(float) + 256 + 255) & ~(size_t)255; }

size_t tc_conv_ws_bytes(const dnb_win_t* g) {
    dnb_win_t g1;
    size_t n = ws_weights_bytes(g);
    if (dgrad_dilate_geom(g, &g1, true)) n += (size_t)g->B * g->F * g1.OH * g1.OW * sizeof(float);
    return n;
}

// dYu[b,f,y,x] = (y % sy == 0 && x % sx == 0 && inside) ? dY[b,f,y/sy,x/sx] : 0; one thread per 4 consecutive x (OW1 % 4 == 0)
__global__ void __launch_bounds__(256) dilate_dy_kernel(const float* __restrict__ dy, float* __restrict__ du, long planes,
                                                        int OH, int OW, int OH1, int OW1, int sy, int sx) {
    const long total = planes * OH1 * (OW1 >> 2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int xq = (int)(i % (OW1 >> 2)); long t = i / (OW1 >> 2);
        const int y = (int)(t % OH1); const long pl = t / OH1;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (y % sy == 0 && y / sy < OH) {
            const float* row = dy + (pl * OH + y / sy) * OW;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int x = 4 * xq + e;
                if (x % sx == 0 && x / sx < OW) v[e] = __ldg(row + x / sx);
            }
        }
        reinterpret_cast<float4*>(du)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

bool tc_conv_supported(const dnb_win_t* g, int kind) {
    if (!get_encode_tiled()) return false;
    const int khw = g->kh * g->kw;
    // tiny reductions (C*kh*kw < 32, e.g. a 1-channel 3x3) are FMA/HBM work for the SIMT kernel, not a tensor-core shape
    if (kind == TC_CONV_FWD) return g->C * khw >= 32 && round_up(g->C, 4) * khw <= MAX_K && g->kh < 128 && g->kw < 256;
    if (kind == TC_CONV_DGRAD) return round_up(g->F, 4) * khw <= MAX_K && g->kh < 128 && g->kw < 256;
    // wgrad: dY rows are read by a 3-D TMA box -> 16-byte aligned plane pitch; M tiles limited by TMEM (512 columns)
    if ((g->OH * g->OW) % 4) return false;
    const int mt = (g->C * khw + 1 + CBM - 1) / CBM;
    const int bn = pick_bn(g->F);
    return mt <= 3 && mt * bn <= 512 && g->F <= 128 * 8;
}

int tc_conv_fwd(const float* x, const float* w, const float* b, float* y, const dnb_win_t* g, void* ws, bool bf16, cudaStream_t st) {
    if (!ws) return set_err(DNB_ERR_INVALID, "conv2d_fwd(tf32): workspace required");
    const int cp = round_up(g->C, 4), Kp = round_up(cp * g->kh * g->kw, CBK), n_pad = round_up(g->F, 128);
    float* wp = static_cast<float*>(ws);
    const bool sv = !getenv("DNB_NO_SV") && tc_conv_sv_supported(g, false);
    const bool bf = bf16 && sv && tc_conv_sv_bf16_supported(g, false);
    prep_weights_kernel<<<(n_pad * Kp + 255) / 256, 256, 0, st>>>(w, wp, g->F, g->C, g->kh, g->kw, n_pad, Kp, cp, 0, bf);
    DNB_LAUNCH_CHECK("conv2d_prep_weights");
    if (sv) return tc_conv_sv_run("tc_conv2d_fwd_sv", x, wp, n_pad, Kp, b, (float)(g->C * g->kh * g->kw), y, g, false, bf, st);
    ConvGemmParams p{x, y, b, g->B, g->C, g->H, g->W, g->F, g->OH, g->OW, g->kh, g->kw, cp, Kp,
                     g->sy, g->sx, 1, 1, g->pt, g->pl, (float)(g->C * g->kh * g->kw)};
    return run_gather("tc_conv2d_fwd", wp, n_pad, p, st);
}

int tc_conv_fwd_pool(const float* x, const float* w, const float* b, float* act, unsigned char* idx_gate, float alpha,
                     const dnb_win_t* g, const dnb_win_t* pool, void* ws, cudaStream_t st) {
    if (!ws) return set_err(DNB_ERR_INVALID, "conv2d_pool_act_fwd: workspace required");
    const int cp = round_up(g->C, 4), Kp = round_up(cp * g->kh * g->kw, CBK), n_pad = round_up(g->F, 128);
    float* wp = static_cast<float*>(ws);
    prep_weights_kernel<<<(n_pad * Kp + 255) / 256, 256, 0, st>>>(w, wp, g->F, g->C, g->kh, g->kw, n_pad, Kp, cp, 0, 1);
    DNB_LAUNCH_CHECK("conv2d_prep_weights");
    return tc_conv_svp_run(x, wp, n_pad, Kp, b, (float)(g->C * g->kh * g->kw), act, idx_gate, alpha, g, pool, st);
}

int tc_conv_dgrad(const float* dy, const float* w, float* dx, const dnb_win_t* g, void* ws, bool bf16, cudaStream_t st) {
    if (!ws) return set_err(DNB_ERR_INVALID, "conv2d_bwd(tf32): workspace required");
    const int cp = round_up(g->F, 4), Kp = round_up(cp * g->kh * g->kw, CBK), n_pad = round_up(g->C, 128);
    float* wp = static_cast<float*>(ws) + ws_fwd_floats(g);
    dnb_win_t g1;
    if (!getenv("DNB_NO_SV") && dgrad_dilate_geom(g, &g1, bf16)) {
        float* du = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + ws_weights_bytes(g));
        const long planes = (long)g->B * g->F;
        const long quads = planes * g1.OH * (g1.OW >> 2);
        if (g->sx == 2 && g1.OW == 2 * g->OW && g->OW % 4 == 0 && !(reinterpret_cast<uintptr_t>(dy) & 15))
            dilate2_dy_kernel<<<(unsigned)std::min<long>((quads / 2 + 255) / 256, 148L * 16), 256, 0, st>>>(dy, du, planes, g->OH, g->OW, g1.OH, g->sy);
        else
            dilate_dy_kernel<<<(unsigned)std::min<long>((quads + 255) / 256, 148L * 16), 256, 0, st>>>(dy, du, planes, g->OH, g->OW, g1.OH, g1.OW, g->sy, g->sx);
        DNB_LAUNCH_CHECK("conv2d_dilate_dy");
        const bool bf1 = bf16 && tc_conv_sv_bf16_supported(&g1, true);
        prep_weights_kernel<<<(n_pad * Kp + 255) / 256, 256, 0, st>>>(w, wp, g->F, g->C, g->kh, g->kw, n_pad, Kp, cp, 1, bf1);
        DNB_LAUNCH_CHECK("conv2d_prep_weights_t");
        return tc_conv_sv_run("tc_conv2d_bwd_data_sv", du, wp, n_pad, Kp, nullptr, 0.f, dx, &g1, true, bf1, st);
    }
    const bool sv = !getenv("DNB_NO_SV") && tc_conv_sv_supported(g, true);
    const bool bf = bf16 && sv && tc_conv_sv_bf16_supported(g, true);
    prep_weights_kernel<<<(n_pad * Kp + 255) / 256, 256, 0, st>>>(w, wp, g->F, g->C, g->kh, g->kw, n_pad, Kp, cp, 1, bf);
    DNB_LAUNCH_CHECK("conv2d_prep_weights_t");
    if (sv) return tc_conv_sv_run("tc_conv2d_bwd_data_sv", dy, wp, n_pad, Kp, nullptr, 0.f, dx, g, true, bf, st);
    ConvGemmParams p{dy, dx, nullptr, g->B, g->F, g->OH, g->OW, g->C, g->H, g->W, g->kh, g->kw, cp, Kp,
                     1, 1, g->sy, g->sx, g->kh - 1 - g->pt, g->kw - 1 - g->pl, 0.f};
    return run_gather("tc_conv2d_bwd_data", wp, n_pad, p, st);
}

template <int MT, int BN, bool THINK>
static int launch_wgrad(const CUtensorMap& td, WgradParams p, int fb, cudaStream_t st) {
    constexpr int SUB = THINK ? 2 : 1;
    constexpr int ST = THINK ? 8 : ((MT == 1) ? 4 : 3);
    // THINK: the MMA of the last stage reads 16 KB from a 4 KB A block -> keep 16 KB of slack behind the ring
    const size_t smem = ST * SUB * ((THINK ? 32 * CBK * 4 : MT * CBM * CBK * 4) + BN * CBK * 4) + (THINK ? 16384 : 0) + 512 + 1024;
    cudaFuncSetAttribute(tc_conv_wgrad_kernel<MT, BN, THINK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long total_kb = (long)p.B * p.pb_per_img;
    long ctas = std::min<long>(total_kb, std::max(1, kNumSMs / fb));
    p.kb_per_cta = (total_kb + ctas - 1) / ctas;
    ctas = (total_kb + p.kb_per_cta - 1) / p.kb_per_cta;
    dim3 grid((unsigned)ctas, fb);
    tc_conv_wgrad_kernel<MT, BN, THINK><<<grid, WGRAD_THREADS, smem, st>>>(td, p);
    DNB_LAUNCH_CHECK("tc_conv2d_bwd_weight");
    return DNB_OK;
}

int tc_conv_wgrad(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, void* ws,
                  bool bf16, cudaStream_t st) {
    (void)ws;
    if (tc_conv_wgrad_sv_supported(g)) return tc_conv_wgrad_sv(x, dy, gw, gb, g, inv_m, bf16, st);
    if (tc_pointwise_wgrad_supported(g) && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15))
        return tc_pointwise_wgrad(x, dy, gw, gb, g, inv_m, st);
    const int K = g->C * g->kh * g->kw;
    const int npx = g->OH * g->OW;
    const int mt = (K + 1 + CBM - 1) / CBM;
    const int bn = pick_bn(g->F);
    const int fb = (g->F + bn - 1) / bn;
    CUtensorMap td;
    uint64_t dims[3] = {(uint64_t)npx, (uint64_t)g->F, (uint64_t)g->B};
    uint64_t str[2] = {(uint64_t)npx * 4, (uint64_t)npx * g->F * 4};
    uint32_t box[3] = {CBK, (uint32_t)bn, 1};
    if (!make_tmap_f32(&td, dy, 3, dims, str, box)) return set_err(DNB_ERR_CUDA, "conv2d_bwd_weight(tf32): cuTensorMapEncodeTiled failed");
    WgradParams p{x, gw, gb, g->B, g->C, g->H, g->W, g->F, g->OH, g->OW, g->kh, g->kw, g->sy, g->sx, g->pt, g->pl,
                  K, (npx + 31) / 32, 0, inv_m};
#define DNB_W(MT_, BN_) launch_wgrad<MT_, BN_, false>(td, p, fb, st)
    if (mt == 1 && K + 1 <= 32) {
        switch (bn) { case 16: return launch_wgrad<1, 16, true>(td, p, fb, st); case 32: return launch_wgrad<1, 32, true>(td, p, fb, st);
                      case 64: return launch_wgrad<1, 64, true>(td, p, fb, st); default: return launch_wgrad<1, 128, true>(td, p, fb, st); }
    }
    if (mt == 1) { switch (bn) { case 16: return DNB_W(1, 16); case 32: return DNB_W(1, 32); case 64: return DNB_W(1, 64); default: return DNB_W(1, 128); } }
    if (mt == 2) { switch (bn) { case 16: return DNB_W(2, 16); case 32: return DNB_W(2, 32); case 64: return DNB_W(2, 64); default: return DNB_W(2, 128); } }
    switch (bn) { case 16: return DNB_W(3, 16); case 32: return DNB_W(3, 32); case 64: return DNB_W(3, 64); default: return DNB_W(3, 128); }
#undef DNB_W
}

}  // namespace dnb
