// Conv2D(C = 1, 3x3, stride 1, no padding) -> MaxPool2D(3x3, stride 2, no padding) -> LeakyRelu/Relu on the tensor core
// (dnpy/layers/convolution.py:193-225, pooling.py:43-99, activations.py:17-22) — conv1 + pool1 + relu1 of the MNIST CNN in
// the BF16 math mode.  The fp32 form of the block (fused_cpa.cu, cpa_fwd_kernel) is bound by FMA + compare issue: 120 us
// at B = 8192, F = 32 for 214 MB of traffic.  Here the 9-tap contraction moves to tcgen05 and the SM's issue slots are left
// to the pooling alone.
//
// One MMA covers G = 128 / F images: row m = (image p, filter f) of A holds the filter's 9 taps in K columns
// [16 p, 16 p + 9) and zeros elsewhere (block diagonal), row n of B is conv pixel n of a chunk with the im2col'ed 3x3 patch
// of image p' in K columns [16 p', 16 p' + 9):  D[(p, f)][n] = sum_tap w[f][tap] * x_p[pixel n + tap].  K columns 9..11 of
// every image slot carry 1.0 in B and the three BF16 pieces of K*bias[f] (hi + mid + lo = the fp32 value to 2^-27) in A:
// the bias (quirk Q1: it enters K = 9 times) costs no epilogue instruction.  All 128 TMEM lanes are live: an epilogue thread
// owns one (image, filter) and pools its channel in registers straight from tcgen05.ld.
//
// A chunk = 2 pooled rows = 5 conv rows x NC needed conv columns (25 of 26 for MNIST) = 125 pixel rows, N = 128; 6 chunks
// per image group; 4 accumulator stages of 128 TMEM columns.
//
//   warp 0        raw images (the G images of a group are contiguous: ONE 1-D bulk copy, ring of 4 groups, requested as
//                 soon as the builders have left a slot) + MMA issue: G k-steps (M = 128, N = 128, K = 16) per chunk
//   warps 1-3     builders: im2col of a chunk into a ring of 4 B tiles (K-major, 128B swizzle: row n = 128 B = 4 image slots
//                 x 16 BF16); the 500 (image, pixel row) pairs of a tile are dealt to the 96 threads, 9 LDS + 5 cvt.pack +
//                 2 STS.128 per pair, all loads before the first conversion
//   warps 4-11    epilogue: two groups of four warps (one per TMEM lane quadrant); a group owns every second image group and
//                 two accumulator stages (chunk c -> stage c & 1).  Per chunk a thread walks the 5 conv rows: row-wise 3-max
//                 (max.NaN) + first-equal column, then the column-wise combination with strict > (np.argmax: first maximum
//                 in row-major order), LeakyRelu, value + byte (argmax offset | gate << 4).  Outputs of 4 pooled rows are
//                 staged per warp in the layouts two tensor maps expect (values: 64B-swizzled boxes of 16 floats x 32 rows,
//                 conflict-free STS.128; bytes: 48 B rows) and leave by four TMA tensor stores per warp.
//   Measured (B = 8192, F = 32, DNB_C1P_PROF_BUILD role clocks): builders ~105 k, epilogue ~110 k, MMA warp ~77 k busy clocks
//   of ~130 k per CTA — the three roles are balanced; the pooling is ALU-pipe work (FMNMX3 / FSETP / SEL: 62 % ALU pipe).
//   Tried and measured slower: 16 column-split epilogue warps at 92 registers (96 us: per-warp fixed costs double), MMA
//   issue folded into a builder warp, blocking or polled (136 / 144 us: the builder stalls behind the accumulator stages).
//   NaN: the fast path detects a NaN maximum and repeats the chunk with the sequential comparisons of np.argmax.  Because the
//   images of a group share MMAs through zero blocks, a non-finite INPUT value reaches the other images of its group
//   (0 x Inf = NaN) — the loss of such a batch is NaN under the reference too, only the per-sample intermediates differ.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace dnb {
namespace tc {

struct C1pParams {
    const float* x; const float* w; const float* b;
    int B, F, G, H;            // images, filters, images per MMA (128 / F), image rows
    int ngroups;               // ceil(B / G)
    float alpha, bias_k;
    unsigned long long* prof;  // DNB_C1P_PROF_BUILD: per-CTA clock counters of the roles (null = off)
};

constexpr int C1P_NB = 3;                                   // builder warps
constexpr int C1P_W_MMA = 0, C1P_W_B0 = 1, C1P_W_EPI0 = 4;  // warp 0: raw bulk copies + MMA issue; warps 1-3: builders
constexpr int C1P_THREADS = (C1P_W_EPI0 + 8) * 32;
constexpr int C1P_NSLOT = 4, C1P_NRAW = 4;                  // B tiles in the ring; raw image groups in flight
static_assert(C1P_W_EPI0 % 4 == 0 && C1P_W_B0 + C1P_NB == C1P_W_EPI0, "warp roles");

__device__ __forceinline__ bool c1p_mbar_test(uint64_t* bar, uint32_t parity) {       // non-blocking phase test
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void c1p_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ float c1p_max3_nan(float a, float b, float c) {
    float d;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float c1p_lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void c1p_tensor_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t c1p_bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float c1p_bf16_val(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// One chunk = 2 pooled rows of one (image, filter): the thread walks the 5 conv rows of its TMEM lane (conv row y starts at
// column y * NC), the next row's loads in flight while the current one is pooled.  Values go to the warp's staging boxes
// (pooled rows r0, r0 + 1 of the 4-row store unit), bytes into bw[3 r0 .. 3 r0 + 5].  Returns true when a NaN reached a maximum
// of the fast path (the caller repeats the chunk with NANSAFE = true).
template <int NC, int OWP, bool NANSAFE>
__device__ __forceinline__ bool c1p_pool_chunk(uint32_t taddr, float alpha, uint32_t sv_lane, uint32_t swz, int r0, uint32_t* bw) {
    static_assert(NC > 16 && NC <= 32 && OWP % 4 == 0 && 2 * OWP + 1 == NC, "geometry");
    constexpr int HOFF = NC - 16;                       // second load: conv columns HOFF .. NC-1
    float lo[16], hi[16];
    uint32_t nlo[16], nhi[16];
    float cb[OWP];
    int ci[OWP];
    float chk = 0.f;
    c1p_tmem_ld16(taddr, reinterpret_cast<uint32_t (&)[16]>(lo));
    c1p_tmem_ld16(taddr + HOFF, reinterpret_cast<uint32_t (&)[16]>(hi));
    tmem_ld_wait();
#pragma unroll
    for (int y = 0; y < 5; ++y) {
        if (y < 4) {
            c1p_tmem_ld16(taddr + (uint32_t)((y + 1) * NC), nlo);
            c1p_tmem_ld16(taddr + (uint32_t)((y + 1) * NC + HOFF), nhi);
        }
        float outv[4];
        uint32_t outb = 0;
#pragma unroll
        for (int ox = 0; ox < OWP; ++ox) {
            auto col = [&](int c) -> float { return c < 16 ? lo[c < 16 ? c : 0] : hi[c >= 16 ? c - HOFF : 0]; };
            const float c0 = col(2 * ox), c1 = col(2 * ox + 1), c2 = col(2 * ox + 2);
            float hb; int hj;
            if (NANSAFE) {
                hb = c0; hj = 0;
                if (c1 > hb || (c1 != c1 && hb == hb)) { hb = c1; hj = 1; }
                if (c2 > hb || (c2 != c2 && hb == hb)) { hb = c2; hj = 2; }
            } else {
                hb = c1p_max3_nan(c0, c1, c2);
                hj = (c0 == hb) ? 0 : ((c1 == hb) ? 1 : 2);
            }
            if (y == 0) {
                cb[ox] = hb; ci[ox] = hj;
            } else {
                const bool take = NANSAFE ? (hb > cb[ox] || (hb != hb && cb[ox] == cb[ox])) : (hb > cb[ox]);
                if (!NANSAFE) chk = c1p_max3_nan(chk, cb[ox], hb);
                const float v = take ? hb : cb[ox];
                const int id = take ? 3 * (y & 1 ? 1 : 2) + hj : ci[ox];
                if (y & 1) {
                    cb[ox] = v; ci[ox] = id;
                } else {                                 // third row of a window: emit it; the row also opens the next window
                    const bool pos = v > 0.f;
                    outv[ox & 3] = pos ? v : alpha * v;  // activations.py:17-19
                    outb |= (uint32_t)(id | (pos ? 16 : 0)) << (8 * (ox & 3));
                    if ((ox & 3) == 3) {
                        const int r = r0 + (y >> 1) - 1;
                        const int e = r * OWP + (ox - 3);                 // float index inside the 4-row store unit
                        const uint32_t a = sv_lane + (uint32_t)((e >> 4) * 2048) + ((((uint32_t)(e & 15) >> 2) ^ swz) << 4);
                        sts128(a, make_float4(outv[0], outv[1], outv[2], outv[3]));
                        bw[r * (OWP / 4) + (ox >> 2)] = outb;
                        outb = 0;
                    }
                    cb[ox] = hb; ci[ox] = hj;
                }
            }
        }
        if (y < 4) {
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) { lo[c] = __uint_as_float(nlo[c]); hi[c] = __uint_as_float(nhi[c]); }
        }
    }
    return chk != chk;
}

#ifdef DNB_C1P_PROF_BUILD     // per-role clock counters (build with -DDNB_C1P_PROF_BUILD, run with DNB_C1P_PROF=1)
#define C1P_ACC(slot) do { if (prof_on) { const long long t1_ = clock64(); pc[slot] += (unsigned long long)(t1_ - tprev); tprev = t1_; } } while (0)
#else
#define C1P_ACC(slot) do { } while (0)
#endif

template <int OHP, int OWP, int W>
__global__ void __launch_bounds__(C1P_THREADS, 1)
tc_c1p_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmI, C1pParams p) {
    constexpr int NC = 2 * OWP + 1, NROW = 5 * NC, NPAD = (NROW + 15) & ~15;
    constexpr int NCH = OHP / 2, NPAIR = NCH / 2, UNIT = 4 * OWP;          // chunks / 4-row store units per image; floats per unit
    static_assert(OHP % 4 == 0 && NPAD <= 128 && UNIT % 16 == 0, "geometry");
    constexpr int TILE_B = NPAD * 128;
    constexpr int NBOX = UNIT / 16, SV_WARP = NBOX * 2048, SB_WARP = 32 * UNIT;
    constexpr uint32_t TMEM_COLS = 512;
    static_assert(4 * NPAD <= 512, "four accumulator stages");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int HW = p.H * W;
    const uint32_t raw_bytes = (uint32_t)(((size_t)p.G * HW * 4 + 127) & ~(size_t)127);
    uint8_t* sA = smem;                                        // 128 rows x 128 B
    uint8_t* sB = sA + 128 * 128;                              // NSLOT tiles
    uint8_t* sOV = sB + C1P_NSLOT * TILE_B;                    // 8 warps x NBOX boxes of 2 KB
    uint8_t* sOB = sOV + 8 * SV_WARP;                          // 8 warps x 32 rows x UNIT bytes
    uint8_t* sR = sOB + 8 * SB_WARP;                           // nraw raw groups
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + (size_t)C1P_NRAW * raw_bytes);
    uint64_t* raw_full = bars;            // [4]
    uint64_t* raw_empty = bars + 4;       // [4]
    uint64_t* b_full = bars + 8;          // [NSLOT]
    uint64_t* b_empty = bars + 12;        // [NSLOT]
    uint64_t* acc_full = bars + 16;       // [4]
    uint64_t* acc_empty = bars + 20;      // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int mine = (int)blockIdx.x < p.ngroups ? (p.ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

#ifdef DNB_C1P_PROF_BUILD
    const bool prof_on = p.prof != nullptr && lane == 0;
    unsigned long long pc[4] = {0, 0, 0, 0};
    long long tprev = prof_on ? clock64() : 0;
    const long long tstart = tprev;
#endif
    // A (block diagonal filters + bias pieces) and the zeroed B ring
    for (int e = tid; e < 128 * 8; e += C1P_THREADS) {
        const int m = e >> 3, j = e & 7;
        const int pi = m / p.F, f = m - pi * p.F;
        uint32_t wd[4] = {0u, 0u, 0u, 0u};
        if (j == 2 * pi) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                wd[q] = c1p_bf16_bits(__ldg(p.w + f * 9 + 2 * q)) | (c1p_bf16_bits(__ldg(p.w + f * 9 + 2 * q + 1)) << 16);
        } else if (j == 2 * pi + 1) {
            const float v = p.bias_k * __ldg(p.b + f);
            const float h = c1p_bf16_val(v), r1 = v - h;
            const float md = c1p_bf16_val(r1), r2 = r1 - md;
            wd[0] = c1p_bf16_bits(__ldg(p.w + f * 9 + 8)) | (c1p_bf16_bits(h) << 16);
            wd[1] = c1p_bf16_bits(md) | (c1p_bf16_bits(r2) << 16);
        }
        sts128u(smem_u32(sA) + (uint32_t)(m * 128 + ((j ^ (m & 7)) << 4)), wd[0], wd[1], wd[2], wd[3]);
    }
    for (uint32_t e = tid; e < (uint32_t)(C1P_NSLOT * TILE_B / 16); e += C1P_THREADS)
        reinterpret_cast<float4*>(sB)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmI);
        for (int s = 0; s < 4; ++s) {
            mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], C1P_NB);
            mbar_init(&b_full[s], C1P_NB); mbar_init(&b_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == C1P_W_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == C1P_W_MMA) {
        const uint32_t idesc = make_idesc_bf16(128, NPAD, 0, 0);
        const uint64_t a_d = make_smem_desc(smem_u32(sA), 16, 1024, LAYOUT_SW128);
        int n = 0, next_load = 0;
        for (int r = 0; 2 * r < mine; ++r)
            for (int c = 0; c < NCH; ++c)
                for (int e = 0; e < 2; ++e) {
                    const int i = 2 * r + e;
                    if (i >= mine) continue;
                    // raw images: one bulk copy per group into a ring of C1P_NRAW, requested as soon as the builders have left
                    // the slot (polled here, once per task: the ring is three groups = 18 tasks ahead of the builders)
                    if (lane == 0) {
                        while (next_load < mine && (next_load < C1P_NRAW ||
                                                    c1p_mbar_test(&raw_empty[next_load & (C1P_NRAW - 1)], ((next_load / C1P_NRAW) - 1) & 1))) {
                            const int rs = next_load & (C1P_NRAW - 1);
                            const int g = (int)blockIdx.x + next_load * (int)gridDim.x;
                            const int nimg = min(p.G, p.B - g * p.G);
                            const uint32_t bytes = (uint32_t)nimg * HW * 4u;
                            mbar_arrive_expect_tx(&raw_full[rs], bytes);
                            bulk_load_1d(sR + (size_t)rs * raw_bytes, p.x + (size_t)g * p.G * HW, bytes, &raw_full[rs]);
                            ++next_load;
                        }
                    }
                    __syncwarp();
                    const int slot = n % C1P_NSLOT, ph = (n / C1P_NSLOT) & 1;
                    ++n;
                    const int ke = r * NCH + c, st = 2 * e + (ke & 1);
                    C1P_ACC(3);
                    mbar_wait(&b_full[slot], ph);
                    C1P_ACC(0);
                    if (ke >= 2) mbar_wait(&acc_empty[st], ((ke >> 1) - 1) & 1);
                    C1P_ACC(1);
                    tc_fence_after();
                    const uint64_t b_d = make_smem_desc(smem_u32(sB + (size_t)slot * TILE_B), 16, 1024, LAYOUT_SW128);
                    const uint32_t d_t = tmem_base + (uint32_t)(st * NPAD);
                    if (elect_one()) {
                        for (int ks = 0; ks < p.G; ++ks) {
                            if (ks == 0) mma_bf16_init(d_t, a_d, b_d, idesc);
                            else mma_bf16_acc(d_t, a_d + 2 * ks, b_d + 2 * ks, idesc);
                        }
                        mma_commit(&acc_full[st]);
                        mma_commit(&b_empty[slot]);
                    }
                    __syncwarp();
                }
    } else if (warp < C1P_W_EPI0) {
        // builders: the (image slot, pixel row) pairs of a tile are dealt round robin to the 96 builder threads: pair id =
        // btid + 96 k, k < 6 (the same pairs for every task).  All loads of a thread's pairs are issued before the first
        // conversion and every packed word has its own register before the first store: a builder warp is latency bound
        // (measured with 2 warps: 1 900 clk per tile, the kernel's critical path), so the ILP is what counts.
        constexpr int NBT = C1P_NB * 32, NPR = (4 * NROW + NBT - 1) / NBT;
        const int btid = tid - C1P_W_B0 * 32;
        uint32_t src_off[NPR], dst0[NPR], dst1[NPR];
        int pimg[NPR];
#pragma unroll
        for (int k = 0; k < NPR; ++k) {
            const int id = btid + k * NBT;
            const int pi = id / NROW, nrow = id - pi * NROW;
            const int y = nrow / NC, x = nrow - y * NC;
            pimg[k] = id < 4 * NROW ? pi : 4;                                    // 4 = never valid
            src_off[k] = (uint32_t)pi * (uint32_t)HW * 4u + (uint32_t)(y * W + x) * 4u;
            dst0[k] = (uint32_t)nrow * 128u + ((((uint32_t)(2 * pi)) ^ (uint32_t)(nrow & 7)) << 4);
            dst1[k] = (uint32_t)nrow * 128u + ((((uint32_t)(2 * pi + 1)) ^ (uint32_t)(nrow & 7)) << 4);
        }
        const uint32_t one = 0x3f80u;
        int n = 0;
        for (int r = 0; 2 * r < mine; ++r)
            for (int c = 0; c < NCH; ++c)
                for (int e = 0; e < 2; ++e) {
                    const int i = 2 * r + e;
                    if (i >= mine) continue;
                    const int slot = n & (C1P_NSLOT - 1);
                    const int rs = i & (C1P_NRAW - 1);
                    C1P_ACC(3);
                    if (c == 0) mbar_wait(&raw_full[rs], (i / C1P_NRAW) & 1);
                    if (n >= C1P_NSLOT) mbar_wait(&b_empty[slot], ((n / C1P_NSLOT) - 1) & 1);
                    C1P_ACC(0);
                    ++n;
                    const int g = (int)blockIdx.x + i * (int)gridDim.x;
                    const int nimg = min(p.G, p.B - g * p.G);
                    const uint32_t raw_s = smem_u32(sR + (size_t)rs * raw_bytes) + (uint32_t)(4 * c * W) * 4u;
                    const uint32_t tile_s = smem_u32(sB + (size_t)slot * TILE_B);
                    float v[NPR][9];
#pragma unroll
                    for (int k = 0; k < NPR; ++k) {
#pragma unroll
                        for (int t = 0; t < 9; ++t) v[k][t] = 0.f;
                        if (pimg[k] < nimg) {
                            const uint32_t a = raw_s + src_off[k];
#pragma unroll
                            for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                                for (int tx = 0; tx < 3; ++tx) v[k][ty * 3 + tx] = c1p_lds32(a + (uint32_t)(ty * W + tx) * 4u);
                        }
                    }
                    uint32_t wq[NPR][5];
#pragma unroll
                    for (int k = 0; k < NPR; ++k) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) wq[k][t] = pack_bf16x2(v[k][2 * t], v[k][2 * t + 1]);
                        wq[k][4] = pack_bf16x2(v[k][8], 1.0f);
                    }
#pragma unroll
                    for (int k = 0; k < NPR; ++k)
                        if (pimg[k] < nimg) {
                            sts128u(tile_s + dst0[k], wq[k][0], wq[k][1], wq[k][2], wq[k][3]);
                            sts128u(tile_s + dst1[k], wq[k][4], one | (one << 16), 0u, 0u);
                        }
                    C1P_ACC(1);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&b_full[slot]);
                        if (c == NCH - 1) mbar_arrive(&raw_empty[rs]);
                    }
                    C1P_ACC(2);
                }
    } else {
        // epilogue: TMEM lane m = q * 32 + lane = (image m / F, filter m % F) of the group = row g * 128 + m of the outputs
        const int q = warp & 3, eg = (warp - C1P_W_EPI0) >> 2, ew = warp - C1P_W_EPI0;
        const uint32_t sv_warp = smem_u32(sOV + (size_t)ew * SV_WARP), sb_warp = smem_u32(sOB + (size_t)ew * SB_WARP);
        const uint32_t sv_lane = sv_warp + (uint32_t)lane * 64u, swz = (uint32_t)(lane >> 1) & 3u;
        const uint32_t sb_lane = sb_warp + (uint32_t)lane * UNIT;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        bool pending = false;
        for (int i = eg; i < mine; i += 2) {
            const long g = (long)blockIdx.x + (long)i * gridDim.x;
            const int row0 = (int)(g * 128 + q * 32);
#pragma unroll 1
            for (int pr = 0; pr < NPAIR; ++pr) {
                uint32_t bw[UNIT / 4];
                if (pending) {                       // the stores of the previous unit have read the staging buffers
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = 2 * pr + h;
                    const int ke = (i >> 1) * NCH + c, st = 2 * eg + (ke & 1);
                    C1P_ACC(3);
                    mbar_wait(&acc_full[st], (ke >> 1) & 1);
                    C1P_ACC(0);
                    tc_fence_after();
                    const uint32_t taddr = tlane + (uint32_t)(st * NPAD);
                    if (__any_sync(0xffffffffu, c1p_pool_chunk<NC, OWP, false>(taddr, p.alpha, sv_lane, swz, 2 * h, bw)))
                        c1p_pool_chunk<NC, OWP, true>(taddr, p.alpha, sv_lane, swz, 2 * h, bw);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[st]);
                    C1P_ACC(1);
                }
#pragma unroll
                for (int t = 0; t < UNIT / 16; ++t) sts128u(sb_lane + 16u * t, bw[4 * t], bw[4 * t + 1], bw[4 * t + 2], bw[4 * t + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int bx = 0; bx < NBOX; ++bx)
                        c1p_tensor_store_2d(&tmV, reinterpret_cast<const void*>(sOV + (size_t)ew * SV_WARP + bx * 2048), pr * UNIT + bx * 16, row0);
                    c1p_tensor_store_2d(&tmI, reinterpret_cast<const void*>(sOB + (size_t)ew * SB_WARP), pr * UNIT, row0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pending = true;
                C1P_ACC(2);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
#ifdef DNB_C1P_PROF_BUILD
    if (prof_on && (warp == C1P_W_MMA || warp == C1P_W_B0 || warp == C1P_W_EPI0 || warp == C1P_W_EPI0 + 1)) {
        const int role = warp == C1P_W_MMA ? 0 : (warp == C1P_W_B0 ? 1 : (warp == C1P_W_EPI0 ? 2 : 3));
        C1P_ACC(3);
        unsigned long long* o = p.prof + (size_t)blockIdx.x * 20 + role * 4;
        for (int k = 0; k < 4; ++k) o[k] = pc[k];
        if (role == 0) p.prof[(size_t)blockIdx.x * 20 + 16] = (unsigned long long)(clock64() - tstart);
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == C1P_W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
}

static size_t c1p_smem_bytes(int G, int HW, int nraw, int npad, int unit) {
    const size_t raw = ((size_t)G * HW * 4 + 127) & ~(size_t)127;
    return 1024 + 128 * 128 + (size_t)C1P_NSLOT * npad * 128 + 8 * (size_t)(unit / 16) * 2048 + 8 * (size_t)32 * unit + nraw * raw + 256;
}

static bool c1p_tmap(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, uint64_t cols, uint64_t rows,
                     uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows}, gstr[1] = {cols * (uint64_t)esize};
    cuuint32_t bx[2] = {box_cols, box_rows}, es[2] = {1, 1};
    return enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

using namespace tc;

static bool c1p_geometry_ok(const dnb_win_t* c, const dnb_win_t* pl) {
    if (!c || !pl || !get_encode_tiled()) return false;
    if (c->C != 1 || c->kh != 3 || c->kw != 3 || c->sy != 1 || c->sx != 1 || c->pt || c->pb || c->pl || c->pr) return false;
    if (pl->kh != 3 || pl->kw != 3 || pl->sy != 2 || pl->sx != 2 || pl->pt || pl->pb || pl->pl || pl->pr) return false;
    if (pl->B != c->B || pl->C != c->F || pl->H != c->OH || pl->W != c->OW) return false;
    if (c->OH != c->H - 2 || c->OW != c->W - 2) return false;
    if (!(pl->OH == 12 && pl->OW == 12 && c->W == 28)) return false;                 // instantiated geometry (MNIST)
    if (2 * (pl->OH - 1) + 3 > c->OH) return false;
    if (c->F != 32 && c->F != 64 && c->F != 128) return false;                       // G = 128 / F images per MMA
    return ((long)c->H * c->W * 4) % 16 == 0;
}

extern "C" {

int dnb_conv_pool_act_tc_supported(const dnb_win_t* conv, const dnb_win_t* pool, int math) {
    return (math == DNB_MATH_BF16 && c1p_geometry_ok(conv, pool)) ? 1 : 0;
}

int dnb_conv_pool_act_fwd_tc(const float* x, const float* w, const float* b, float* act_out, uint8_t* idx_gate,
                             const dnb_win_t* conv, const dnb_win_t* pool, float alpha, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(math == DNB_MATH_BF16 && c1p_geometry_ok(conv, pool), "conv_pool_act_fwd_tc: unsupported geometry / math mode");
    if (conv->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && act_out && idx_gate, "conv_pool_act_fwd_tc: null pointer");
    DNB_CHECK_ARG(!((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(act_out) | reinterpret_cast<uintptr_t>(idx_gate)) & 15),
                  "conv_pool_act_fwd_tc: buffers must be 16-byte aligned");
    constexpr int OHP = 12, OWP = 12, W = 28, NPAD = 128, UNIT = 4 * OWP;
    C1pParams p{};
    p.x = x; p.w = w; p.b = b; p.B = conv->B; p.F = conv->F; p.G = 128 / conv->F; p.H = conv->H;
    p.ngroups = (conv->B + p.G - 1) / p.G;
    p.alpha = alpha; p.bias_k = 9.f;                                 // Q1: the bias enters C*kh*kw = 9 times
    const int HW = conv->H * W;
    const size_t smem = c1p_smem_bytes(p.G, HW, C1P_NRAW, NPAD, UNIT);
    if (smem > 225 * 1024) return set_err(DNB_ERR_UNSUPPORTED, "conv_pool_act_fwd_tc: image group does not fit shared memory");
    const uint64_t OP = (uint64_t)OHP * OWP, rows = (uint64_t)conv->B * conv->F;
    CUtensorMap tv, ti;
    if (!c1p_tmap(&tv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, act_out, OP, rows, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B) ||
        !c1p_tmap(&ti, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, idx_gate, OP, rows, UNIT, 32, CU_TENSOR_MAP_SWIZZLE_NONE))
        return set_err(DNB_ERR_CUDA, "conv_pool_act_fwd_tc: cuTensorMapEncodeTiled failed");
    cudaFuncSetAttribute(tc_c1p_kernel<OHP, OWP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = std::min(p.ngroups, kNumSMs);
    if (getenv("DNB_C1P_PROF")) {            // per-role clock counters (tools/time_c1p.py): synchronous, never on the step path
        cudaMalloc(&p.prof, (size_t)grid * 20 * sizeof(unsigned long long));
        cudaMemset(p.prof, 0, (size_t)grid * 20 * sizeof(unsigned long long));
    }
    tc_c1p_kernel<OHP, OWP, W><<<grid, C1P_THREADS, smem, S(stream)>>>(tv, ti, p);
    DNB_LAUNCH_CHECK("conv_pool_act_fwd_tc");
    if (p.prof) {
        std::vector<unsigned long long> h((size_t)grid * 20);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), p.prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaFree(p.prof);
        double a[20] = {0};
        for (int b2 = 0; b2 < grid; ++b2) for (int k = 0; k < 20; ++k) a[k] += (double)h[(size_t)b2 * 20 + k] / grid;
        const char* names[4] = {"mma     (wait b_full | wait acc_empty | - | issue)", "builder (waits | loads-cvt-stores | fence + arrive | rest)",
                                "epi hx0 (wait acc_full | pool | store hand-over | rest)", "epi hx1 (wait acc_full | pool | store hand-over | rest)"};
        fprintf(stderr, "c1p prof: CTA clocks %.0f\n", a[16]);
        for (int r = 0; r < 4; ++r) fprintf(stderr, "  %s: %.0f %.0f %.0f %.0f\n", names[r], a[r * 4], a[r * 4 + 1], a[r * 4 + 2], a[r * 4 + 3]);
    }
    return DNB_OK;
}

}  // extern "C"

}  // namespace dnb
