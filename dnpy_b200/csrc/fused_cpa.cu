// Fused first block of the MNIST-shape CNN: Conv2D(C=1, 3x3, stride 1, no padding) -> MaxPool2D(3x3, stride 2, no padding)
// -> LeakyRelu/Relu  (dnpy/layers/convolution.py:193-273, pooling.py:43-99, activations.py:17-22).
//
// Why: at B=8192, F=32 the convolution output is 709 MB that the unfused step writes, re-reads in the pool, and never
// touches again; its gradient (another 709 MB) is written by the pool backward and read twice by the conv backward.
// Here neither tensor reaches HBM:
//
//   forward : x (3 KB / image) arrives by one bulk copy; a thread owns a PAIR of filters and one pooled column, walks
//             down the conv rows with a rolling 3x5 input window in registers (27 packed FFMA2 per conv row = 54 FMAs),
//             pools on the fly (first maximum wins, NaN propagates: np.argmax semantics) and emits LeakyRelu(pool) plus ONE
//             byte per pooled element: bits 0-3 = argmax offset inside the window (0..8), bit 4 = the activation gate
//             (pool > 0).  Outputs are staged in shared memory and leave by bulk stores.
//             HBM: 3.1 KB in, F*OH*OW*5 B out per image (23 KB) instead of 86 KB + 86 KB + 18 KB + 18 KB + 18 KB.
//   backward: (PER_SAMPLE routing) the conv weight gradient straight from the activation's delta: each pooled element
//             routes delta*gate to ONE conv output, so gW[f,i,j] += d * x[y*+i, x*+j] — 9 FMAs per pooled element
//             (4.7x fewer than the dense 26x26 correlation), no dY tensor at all.  delta, the gate/argmax bytes and the
//             image arrive by bulk copies; a thread keeps its filter's 9+1 sums in registers over all images of its CTA.
//
// The intermediate tensors stay observable: the host side (dnpy_b200/net.py, layers) hands out LazyDArrays that run the
// unfused kernels if somebody reads conv.output / pool.output / pool.delta / conv.delta.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {

namespace {

constexpr int CPA_MAX_THREADS = 512;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(tc::smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void group_bar(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// NaN-propagating 3-input maximum (FMNMX3.NAN): NaN if any operand is NaN
__device__ __forceinline__ float max3_nan(float a, float b, float c) {
    float d;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

struct CpaFwd {
    const float* x; const float* w; const float* b;
    float* act; uint8_t* idxg;
    int B, H, F;                    // H rows of W floats per image
    int gt, ngroups;                // threads per image group (multiple of 32), groups per CTA
    float alpha;
};

// One pooled column of one filter pair, top to bottom; geometry is compile-time so every shared-memory offset is an
// immediate and the 2*OH+1 conv rows unroll into straight-line code.
//   fast path (NANSAFE = false): value = FMNMX3.NAN of the three conv outputs of the row, position by equality (first
//   match = first maximum); returns true when a NaN reached any maximum, and the caller repeats the column with the
//   NaN-exact comparisons of np.argmax (first NaN wins and sticks).  Never taken on finite data.
template <int OH, int OW, int W, bool NANSAFE, bool PACKED = true>
__device__ __forceinline__ bool cpa_walk(const float* __restrict__ img, const float2 (&w2)[3][3], float2 bias2, float alpha,
                                         uint32_t sa, uint32_t si) {
    constexpr int OP = OH * OW, R = 2 * OH + 1;       // conv rows the pool reads: 0 .. 2*OH
    float X[3][5];
    float cb[2] = {0.f, 0.f};
    int ci[2] = {0, 0};
    float chk = 0.f;
    auto load_row = [&](float (&dst)[5], int r) {
        const float* rp = img + r * W;
        const float2 a = *reinterpret_cast<const float2*>(rp);
        const float2 b = *reinterpret_cast<const float2*>(rp + 2);
        dst[0] = a.x; dst[1] = a.y; dst[2] = b.x; dst[3] = b.y; dst[4] = rp[4];
    };
    load_row(X[0], 0);
    load_row(X[1], 1);
#pragma unroll
    for (int y = 0; y < R; ++y) {
        load_row(X[(y + 2) % 3], y + 2);
        float2 acc[3] = {bias2, bias2, bias2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float xv = X[(y + i) % 3][j + jj];
                    if (PACKED) acc[j] = __ffma2_rn(make_float2(xv, xv), w2[i][jj], acc[j]);
                    else { acc[j].x = fmaf(xv, w2[i][jj].x, acc[j].x); acc[j].y = fmaf(xv, w2[i][jj].y, acc[j].y); }
                }
        const int dy = (y & 1) ? 1 : 0;               // window row of a window that STARTED at the last even row
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float c0 = q ? acc[0].y : acc[0].x, c1 = q ? acc[1].y : acc[1].x, c2 = q ? acc[2].y : acc[2].x;
            float hb; int hj;
            if (NANSAFE) {
                hb = c0; hj = 0;
                if (c1 > hb || (c1 != c1 && hb == hb)) { hb = c1; hj = 1; }
                if (c2 > hb || (c2 != c2 && hb == hb)) { hb = c2; hj = 2; }
            } else {
                hb = max3_nan(c0, c1, c2);
                hj = (c0 == hb) ? 0 : ((c1 == hb) ? 1 : 2);
            }
            if ((y & 1) == 0) {
                if (y >= 2) {                         // third row of the window that started at y - 2: finish and emit it
                    const bool take = NANSAFE ? (hb > cb[q] || (hb != hb && cb[q] == cb[q])) : (hb > cb[q]);
                    const float v = take ? hb : cb[q];
                    const int id = take ? 6 + hj : ci[q];
                    if (!NANSAFE) chk = max3_nan(chk, v, hb);
                    const bool pos = v > 0.f;
                    constexpr int dummy = 0; (void)dummy;
                    const int o = q * OP + ((y >> 1) - 1) * OW;
                    tc::sts32(sa + 4u * o, pos ? v : alpha * v);
                    sts_u8(si + o, (uint32_t)(id | (pos ? 16 : 0)));
                }
                cb[q] = hb; ci[q] = hj;               // first row of the window starting here
            } else {
                const bool take = NANSAFE ? (hb > cb[q] || (hb != hb && cb[q] == cb[q])) : (hb > cb[q]);
                if (!NANSAFE) chk = max3_nan(chk, cb[q], hb);
                cb[q] = take ? hb : cb[q];
                ci[q] = take ? 3 * dy + hj : ci[q];
            }
        }
    }
    return chk != chk;
}

// smem: [ngroups] x { full[2] mbarriers (padded to 128 B) | img[2][H*W] | act[2][F*OP] | idx[2][F*OP bytes] }
template <int OH, int OW, int W, bool PACKED = true>
__global__ void __launch_bounds__(384, 2) cpa_fwd_kernel(CpaFwd a) {
    extern __shared__ __align__(128) uint8_t smraw[];
    constexpr int OP = OH * OW;
    const int HW = a.H * W;
    const size_t img_b = (size_t)HW * 4, act_b = (size_t)a.F * OP * 4, idx_b = (size_t)a.F * OP;
    const size_t group_b = 128 + 2 * img_b + 2 * act_b + 2 * idx_b;        // every term is a multiple of 16
    const int grp = threadIdx.x / a.gt, gtid = threadIdx.x - grp * a.gt;
    uint8_t* base = smraw + (size_t)grp * group_b;
    uint64_t* full = reinterpret_cast<uint64_t*>(base);
    float* simg = reinterpret_cast<float*>(base + 128);
    float* sact = reinterpret_cast<float*>(base + 128 + 2 * img_b);
    uint8_t* sidx = base + 128 + 2 * img_b + 2 * act_b;
    if (gtid == 0) {
        tc::mbar_init(&full[0], 1);
        tc::mbar_init(&full[1], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    // images of this group: first + n * stride
    const long first = (long)blockIdx.x * a.ngroups + grp, stride = (long)gridDim.x * a.ngroups;
    const int mine = first < a.B ? (int)((a.B - first + stride - 1) / stride) : 0;
    const int NI = (a.F >> 1) * OW;                   // items: (filter pair, pooled column)
    const bool active = gtid < NI;
    const int fp = active ? gtid / OW : 0, ox = active ? gtid - fp * OW : 0;
    const int f0 = 2 * fp;
    float2 w2[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) w2[i][j] = make_float2(__ldg(a.w + f0 * 9 + i * 3 + j), __ldg(a.w + (f0 + 1) * 9 + i * 3 + j));
    const float2 bias2 = make_float2(9.f * __ldg(a.b + f0), 9.f * __ldg(a.b + f0 + 1));      // Q1: the bias enters K = 9 times
    if (gtid == 0 && mine > 0) {
        tc::mbar_arrive_expect_tx(&full[0], (uint32_t)img_b);
        bulk_g2s(simg, a.x + first * HW, (uint32_t)img_b, &full[0]);
    }
    const int bar_id = 1 + grp;
    const uint32_t sa_item = tc::smem_u32(sact) + 4u * (uint32_t)(f0 * OP + ox);
    const uint32_t si_item = tc::smem_u32(sidx) + (uint32_t)(f0 * OP + ox);
    for (int n = 0; n < mine; ++n) {
        const int s = n & 1;
        if (gtid == 0 && n + 1 < mine) {               // stage s^1 was released by barrier A of image n-1
            tc::mbar_arrive_expect_tx(&full[s ^ 1], (uint32_t)img_b);
            bulk_g2s(simg + (size_t)(s ^ 1) * HW, a.x + (first + (long)(n + 1) * stride) * HW, (uint32_t)img_b, &full[s ^ 1]);
        }
        tc::mbar_wait(&full[s], (n >> 1) & 1);
        if (active) {
            const float* img = simg + (size_t)s * HW + 2 * ox;
            const uint32_t sa = sa_item + (uint32_t)(s * act_b), si = si_item + (uint32_t)(s * idx_b);
            if (cpa_walk<OH, OW, W, false, PACKED>(img, w2, bias2, a.alpha, sa, si))
                cpa_walk<OH, OW, W, true, PACKED>(img, w2, bias2, a.alpha, sa, si);
        }
        tc::fence_proxy_async_smem();
        group_bar(bar_id, a.gt);                        // A: staging[s] complete, image stage s no longer read
        if (gtid == 0) {
            const long img_i = first + (long)n * stride;
            bulk_s2g(a.act + img_i * a.F * OP, sact + (size_t)s * a.F * OP, (uint32_t)act_b);
            bulk_s2g(a.idxg + img_i * idx_b, sidx + (size_t)s * idx_b, (uint32_t)idx_b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the stores of image n-1 (staging[s^1]) have drained
        }
        group_bar(bar_id, a.gt);                        // B: staging[s^1] may be rewritten
    }
    if (gtid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct CpaBwd {
    const float* x; const float* dact; const uint8_t* idxg;
    float* gw; float* gb;
    int B, H, W, F, OH, OW;
    int ncons;                      // consumer threads: 32 lanes (= 32 filters) x OH pooled rows x F/32 filter groups
    int DP;                         // plane pitch of the staged delta in floats: >= OH*OW, DP/4 odd (conflict-free 128-bit loads)
    float alpha, inv_m;
};
constexpr int CPB_ST = 3;

// Weight gradient of the block.  Lane = filter (mod 32), warp = (filter group, pooled row): every lane of a warp works on
// the SAME pooled position, so the nine gathered window words differ between lanes only through the argmax offset
// (<= 9 distinct addresses, in 9 distinct banks when the image pitch is 28: conflict-free; the natural
// (filter, column) lane order measured 2.2 wavefronts per load).  The delta planes are staged with a pitch of DP floats
// (DP/4 odd) so a lane's row of OW deltas is three conflict-free 128-bit loads.
// smem: full[ST], empty[ST] (128 B) | red[F*10] (padded to 16 B) | ST x { dact[F*DP] f32 | idx[F*OP] u8 | img[H*W] f32 }
__global__ void __launch_bounds__(CPA_MAX_THREADS + 32) cpa_bwd_kernel(CpaBwd a) {
    extern __shared__ __align__(128) uint8_t smraw[];
    const int OP = a.OH * a.OW, HW = a.H * a.W;
    const size_t d_b = (size_t)a.F * a.DP * 4, i_b = (size_t)a.F * OP, x_b = (size_t)HW * 4, st_b = d_b + i_b + x_b;
    const size_t red_b = ((size_t)a.F * 10 * 4 + 15) & ~(size_t)15;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + CPB_ST;
    float* red = reinterpret_cast<float*>(smraw + 128);
    uint8_t* stages = smraw + 128 + red_b;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < CPB_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], a.ncons / 32); }
        tc::fence_barrier_init();
    }
    for (int e = tid; e < a.F * 10; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
    const int mine = (int)blockIdx.x < a.B ? (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (wid == a.ncons / 32) {                          // producer warp: F + 2 bulk copies per image, spread over the lanes
        for (int n = 0; n < mine; ++n) {
            const int s = n % CPB_ST;
            if (n >= CPB_ST) tc::mbar_wait(&empty[s], ((n / CPB_ST) - 1) & 1);
            const long b = (long)blockIdx.x + (long)n * gridDim.x;
            uint8_t* st = stages + (size_t)s * st_b;
            if (lane == 0) tc::mbar_arrive_expect_tx(&full[s], (uint32_t)((size_t)a.F * OP * 4 + i_b + x_b));
            __syncwarp();
            for (int f = lane; f < a.F; f += 32)
                bulk_g2s(st + (size_t)f * a.DP * 4, a.dact + (b * a.F + f) * OP, (uint32_t)OP * 4u, &full[s]);
            if (lane == 0) {
                bulk_g2s(st + d_b, a.idxg + b * i_b, (uint32_t)i_b, &full[s]);
                bulk_g2s(st + d_b + i_b, a.x + b * HW, (uint32_t)x_b, &full[s]);
            }
        }
        return;
    }
    const int fg = wid / a.OH, oy = wid - fg * a.OH;    // warp -> (filter group, pooled row)
    const int f = fg * 32 + lane;
    float g[9], gbias = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = 0.f;
    const int W = a.W, OW = a.OW;
    const bool vec = (OW & 3) == 0;
    for (int n = 0; n < mine; ++n) {
        const int s = n % CPB_ST;
        tc::mbar_wait(&full[s], (n / CPB_ST) & 1);
        const uint8_t* st = stages + (size_t)s * st_b;
        const float* sd = reinterpret_cast<const float*>(st) + f * a.DP + oy * OW;
        const uint8_t* si = st + d_b + f * OP + oy * OW;
        const float* sx = reinterpret_cast<const float*>(st + d_b + i_b) + 2 * oy * W;
        for (int ox0 = 0; ox0 < OW; ox0 += 4) {
            float d4[4]; unsigned i4;
            if (vec) {
                const float4 v = *reinterpret_cast<const float4*>(sd + ox0);
                d4[0] = v.x; d4[1] = v.y; d4[2] = v.z; d4[3] = v.w;
                i4 = *reinterpret_cast<const uint32_t*>(si + ox0);
            } else {
                i4 = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = ox0 + k < OW;
                    d4[k] = ok ? sd[ox0 + k] : 0.f;
                    i4 |= (ok ? (unsigned)si[ox0 + k] : 0u) << (8 * k);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!vec && ox0 + k >= OW) break;
                const unsigned ib = (i4 >> (8 * k)) & 0xffu;
                const float dd = (ib & 16u) ? d4[k] : a.alpha * d4[k];      // activations.py:21-22
                const int off = (int)(ib & 15u);
                const int wy = (off * 11) >> 5;                             // off / 3 for off in 0..8
                const int wx = off - 3 * wy;
                const float* xp = sx + wy * W + 2 * (ox0 + k) + wx;         // conv output (2oy+wy, 2ox+wx): its 3x3 input window
                gbias += dd;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) g[i * 3 + j] = fmaf(dd, xp[i * W + j], g[i * 3 + j]);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&red[f * 10 + k], g[k]);
    atomicAdd(&red[f * 10 + 9], gbias);
    group_bar(1, a.ncons);
    for (int e = tid; e < a.F * 10; e += a.ncons) {
        const int ff = e / 10, k = e - ff * 10;
        const float v = red[e] * a.inv_m;
        if (k < 9) atomicAdd(a.gw + ff * 9 + k, v); else atomicAdd(a.gb + ff, v);
    }
}

// ---- LITERAL pool routing (dnpy/layers/pooling.py:95 as written) ------------------------------------------
// `delta_view.flat[idxs[:, z, y, x]] += delta[:, z, y, x]` indexes the (B, 3, 3) window view with offsets in [0, 9): every
// sample's delta lands in SAMPLE 0's window, and NumPy's buffered fancy-index "+=" keeps, per offset, only the LAST sample
// that chose it.  So only image 0 carries a conv gradient, built from <= 9 (sample, delta) pairs per pooled position:
//   last[f, pos, off] = max { b : argmax[b, f, pos] == off }                       (cpa_last_kernel)
//   gW[f, i, j] += inv_m * sum_{pos, off} gate * delta[last, f, pos] * x0[2oy + off/3 + i, 2ox + off%3 + j]   (cpa_bwd_literal_kernel)
constexpr int CPL_CHUNK = 256;      // samples per CTA row of the scan

// grid (ceil(F*OP / 256), ceil(B / CPL_CHUNK)): a thread owns one pooled position of one filter and scans its chunk of
// samples from the newest down (coalesced: consecutive threads read consecutive bytes of a sample), stopping when all nine
// offsets have been seen; one atomicMax per offset found.
__global__ void __launch_bounds__(256) cpa_last_kernel(const uint8_t* __restrict__ idxg, int32_t* __restrict__ last,
                                                       int B, int FOP) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= FOP) return;
    const int b_hi = min(B, (int)(blockIdx.y + 1) * CPL_CHUNK) - 1, b_lo = blockIdx.y * CPL_CHUNK;
    unsigned seen = 0;
    int found[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) found[k] = -1;
    for (int b = b_hi; b >= b_lo && seen != 0x1ffu; --b) {
        const unsigned off = idxg[(size_t)b * FOP + e] & 15u;
        if (!(seen >> off & 1u)) {
            seen |= 1u << off;
#pragma unroll
            for (int k = 0; k < 9; ++k) if (k == (int)off) found[k] = b;
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (found[k] >= 0) atomicMax(last + (size_t)e * 9 + k, found[k]);
}

// grid F, block >= OP threads: CTA f reduces its filter's nine taps (+ bias) over the pooled positions; no atomics.
__global__ void __launch_bounds__(256) cpa_bwd_literal_kernel(const float* __restrict__ x0, const float* __restrict__ dact,
                                                              const uint8_t* __restrict__ idxg, const int32_t* __restrict__ last,
                                                              float* __restrict__ gw, float* __restrict__ gb,
                                                              int F, int OH, int OW, int W, float alpha, float inv_m) {
    __shared__ float red[10][8];
    const int f = blockIdx.x, OP = OH * OW, FOP = F * OP;
    float g[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) g[k] = 0.f;
    for (int pos = threadIdx.x; pos < OP; pos += blockDim.x) {
        const int oy = pos / OW, ox = pos - oy * OW;
        const int e = f * OP + pos;
#pragma unroll
        for (int off = 0; off < 9; ++off) {
            const int b = last[(size_t)e * 9 + off];
            if (b < 0) continue;
            const float d = dact[(size_t)b * FOP + e];
            const float dd = (idxg[(size_t)b * FOP + e] & 16u) ? d : alpha * d;       // the activation's backward (activations.py:21-22)
            const float* xp = x0 + (2 * oy + off / 3) * W + 2 * ox + off % 3;
            g[9] += dd;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) g[i * 3 + j] = fmaf(dd, xp[i * W + j], g[i * 3 + j]);
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const float v = warp_sum(g[k]);
        if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        if (threadIdx.x < 9) gw[f * 9 + threadIdx.x] += v * inv_m; else gb[f] += v * inv_m;
    }
}

__global__ void __launch_bounds__(256) idx_unpack_kernel(const uint8_t* __restrict__ in, int32_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)(in[i] & 15u);
}

int delta_pitch(int OP) {           // smallest pitch >= OP that is a multiple of 4 floats with an odd number of quads
    int q = (OP + 3) / 4;
    if (!(q & 1)) ++q;
    return 4 * q;
}

// forward kernels are instantiated per geometry (pooled plane, image width): MNIST (28x28 -> 12x12) and its 14x14 test variant
template <int OH, int OW, int W>
int launch_fwd(const CpaFwd& a, int grid, size_t smem, cudaStream_t st) {
    if (getenv("DNB_CPA_SCALAR")) {      // experiment: scalar FFMA instead of packed FFMA2
        cudaFuncSetAttribute(cpa_fwd_kernel<OH, OW, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cpa_fwd_kernel<OH, OW, W, false><<<grid, a.gt * a.ngroups, smem, st>>>(a);
        return 0;
    }
    cudaFuncSetAttribute(cpa_fwd_kernel<OH, OW, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cpa_fwd_kernel<OH, OW, W><<<grid, a.gt * a.ngroups, smem, st>>>(a);
    return 0;
}
bool fwd_instance(int OH, int OW, int W) { return (OH == 12 && OW == 12 && W == 28) || (OH == 5 && OW == 5 && W == 14); }

__global__ void __launch_bounds__(256) idx_set_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)((out[i] & 0xf0u) | ((unsigned)in[i] & 15u));
}

bool geometry_ok(const dnb_win_t* c, const dnb_win_t* p) {
    if (!c || !p) return false;
    if (c->C != 1 || c->kh != 3 || c->kw != 3 || c->sy != 1 || c->sx != 1) return false;
    if (c->pt || c->pb || c->pl || c->pr) return false;
    if (p->kh != 3 || p->kw != 3 || p->sy != 2 || p->sx != 2 || p->pt || p->pb || p->pl || p->pr) return false;
    if (p->B != c->B || p->C != c->F || p->H != c->OH || p->W != c->OW) return false;
    if (c->OH != c->H - 2 || c->OW != c->W - 2) return false;
    if (p->OH < 1 || p->OW < 1 || 2 * (p->OH - 1) + 3 > c->OH || 2 * (p->OW - 1) + 3 > c->OW) return false;
    if (!fwd_instance(p->OH, p->OW, c->W)) return false;
    if (c->F < 32 || (c->F & 31)) return false;                          // backward: lane = filter
    const long OP = (long)p->OH * p->OW, HW = (long)c->H * c->W;
    if ((HW * 4) % 16 || (c->F * OP) % 16 || (OP * 4) % 16) return false;    // bulk copies: 16-byte granules
    if ((c->F / 2) * p->OW > 384 || (c->F / 32) * p->OH * 32 > CPA_MAX_THREADS) return false;
    const size_t fwd_group = 128 + 2 * (size_t)HW * 4 + 2 * (size_t)c->F * OP * 5;
    const size_t bwd = 128 + (size_t)c->F * 40 + 16 + CPB_ST * ((size_t)c->F * (delta_pitch((int)OP) * 4 + OP) + (size_t)HW * 4);
    return fwd_group <= 100 * 1024 && bwd <= 100 * 1024;
}

}  // namespace

extern "C" {

int dnb_conv_pool_act_supported(const dnb_win_t* conv, const dnb_win_t* pool) { return geometry_ok(conv, pool) ? 1 : 0; }

int dnb_conv_pool_act_fwd(const float* x, const float* w, const float* b, float* act_out, uint8_t* idx_gate,
                          const dnb_win_t* conv, const dnb_win_t* pool, float alpha, dnb_stream_t stream) {
    DNB_CHECK_ARG(geometry_ok(conv, pool), "conv_pool_act_fwd: unsupported geometry");
    if (conv->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && w && b && act_out && idx_gate, "conv_pool_act_fwd: null pointer");
    DNB_CHECK_ARG(!((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(act_out) | reinterpret_cast<uintptr_t>(idx_gate)) & 15),
                  "conv_pool_act_fwd: buffers must be 16-byte aligned");
    CpaFwd a{x, w, b, act_out, idx_gate, conv->B, conv->H, conv->F, 0, 0, alpha};
    const int items = (conv->F / 2) * pool->OW;
    a.gt = (items + 31) & ~31;
    a.ngroups = std::max(1, 384 / a.gt);
    const long OP = (long)pool->OH * pool->OW;
    const size_t group_b = 128 + 2 * (size_t)conv->H * conv->W * 4 + 2 * (size_t)conv->F * OP * 5;
    const size_t smem = group_b * a.ngroups;
    const int per_sm = std::max<int>(1, std::min<size_t>(2, (220 * 1024) / smem));
    const long want = ((long)conv->B + a.ngroups - 1) / a.ngroups;
    const int grid = (int)std::min<long>(want, (long)kNumSMs * per_sm);
    if (pool->OH == 12) launch_fwd<12, 12, 28>(a, grid, smem, S(stream));
    else launch_fwd<5, 5, 14>(a, grid, smem, S(stream));
    DNB_LAUNCH_CHECK("conv_pool_act_fwd");
    return DNB_OK;
}

int dnb_conv_pool_act_bwd(const float* x, const float* act_delta, const uint8_t* idx_gate, float* gw, float* gb,
                          const dnb_win_t* conv, const dnb_win_t* pool, float alpha, float inv_m, dnb_stream_t stream) {
    DNB_CHECK_ARG(geometry_ok(conv, pool), "conv_pool_act_bwd: unsupported geometry");
    if (conv->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && act_delta && idx_gate && gw && gb, "conv_pool_act_bwd: null pointer");
    DNB_CHECK_ARG(!((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(act_delta) | reinterpret_cast<uintptr_t>(idx_gate)) & 15),
                  "conv_pool_act_bwd: buffers must be 16-byte aligned");
    const int OP = pool->OH * pool->OW;
    CpaBwd a{x, act_delta, idx_gate, gw, gb, conv->B, conv->H, conv->W, conv->F, pool->OH, pool->OW, 0, delta_pitch(OP), alpha, inv_m};
    a.ncons = (conv->F / 32) * pool->OH * 32;
    const size_t smem = 128 + (((size_t)conv->F * 40 + 15) & ~(size_t)15) +
                        CPB_ST * ((size_t)conv->F * (a.DP * 4 + OP) + (size_t)conv->H * conv->W * 4);
    cudaFuncSetAttribute(cpa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int per_sm = std::max<int>(1, std::min<size_t>(2, (220 * 1024) / smem));
    const int grid = (int)std::min<long>(conv->B, (long)kNumSMs * per_sm);
    cpa_bwd_kernel<<<grid, a.ncons + 32, smem, S(stream)>>>(a);
    DNB_LAUNCH_CHECK("conv_pool_act_bwd");
    return DNB_OK;
}

size_t dnb_conv_pool_act_bwd_literal_ws_bytes(const dnb_win_t* conv, const dnb_win_t* pool) {
    if (!conv || !pool) return 0;
    return (size_t)conv->F * pool->OH * pool->OW * 9 * sizeof(int32_t);
}

int dnb_conv_pool_act_bwd_literal(const float* x, const float* act_delta, const uint8_t* idx_gate, float* gw, float* gb,
                                  const dnb_win_t* conv, const dnb_win_t* pool, float alpha, float inv_m, void* ws,
                                  dnb_stream_t stream) {
    DNB_CHECK_ARG(geometry_ok(conv, pool), "conv_pool_act_bwd_literal: unsupported geometry");
    if (conv->B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && act_delta && idx_gate && gw && gb && ws, "conv_pool_act_bwd_literal: null pointer");
    cudaStream_t st = S(stream);
    const int OP = pool->OH * pool->OW, FOP = conv->F * OP;
    int32_t* last = static_cast<int32_t*>(ws);
    cudaError_t e = cudaMemsetAsync(last, 0xff, (size_t)FOP * 9 * sizeof(int32_t), st);      // -1: offset never chosen
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "conv_pool_act_bwd_literal: %s", cudaGetErrorString(e));
    dim3 grid((FOP + 255) / 256, (conv->B + CPL_CHUNK - 1) / CPL_CHUNK);
    cpa_last_kernel<<<grid, 256, 0, st>>>(idx_gate, last, conv->B, FOP);
    DNB_LAUNCH_CHECK("conv_pool_act_bwd_last");
    cpa_bwd_literal_kernel<<<conv->F, 256, 0, st>>>(x, act_delta, idx_gate, last, gw, gb, conv->F, pool->OH, pool->OW, conv->W,
                                                    alpha, inv_m);
    DNB_LAUNCH_CHECK("conv_pool_act_bwd_literal");
    return DNB_OK;
}

/* overwrite the argmax nibble of every byte (the gate bit stays): MaxPool2D.output_idxs assigned by the caller */
int dnb_idx_gate_set(const int32_t* idx, uint8_t* idx_gate, size_t n, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(idx_gate && idx, "idx_gate_set: null pointer");
    idx_set_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(idx, idx_gate, n);
    DNB_LAUNCH_CHECK("idx_gate_set");
    return DNB_OK;
}

int dnb_idx_gate_unpack(const uint8_t* idx_gate, int32_t* idx, size_t n, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(idx_gate && idx, "idx_gate_unpack: null pointer");
    idx_unpack_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(idx_gate, idx, n);
    DNB_LAUNCH_CHECK("idx_gate_unpack");
    return DNB_OK;
}

}  // extern "C"

}  // namespace dnb
