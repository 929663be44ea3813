// SimpleRNN backward for the tensor-core math modes (dnpy/layers/recurrent.py:190-266, literal truncated BPTT), H = 32.
//
// The reference walks t = T-1 .. 0 per sample; at every step it (a) continues the full-length chain
//     pre(t) = (delta_y[t == T-1] + pre(t+1) . Waa) * (1 - h_t^2),      dx_t = pre(t) . Wax
// and (b) runs an inner loop of up to trunc + 1 lags starting from S_0(t) = pre(t):
//     gWaa += S_k(t) (x) h_{t-k-1},  gWax += S_k(t) (x) x_{t-k},  gba += S_k(t),   S_{k+1}(t) = (S_k(t) * (1 - h_{t-k-1}^2)) . Waa.
// (a) is a serial recurrence per sample: rnn_chain_blk_kernel / rnn_chain_kernel (a warp per sample, as the forward) write pre
// for every (b, t).
// (b) has NO dependency between different (b, t): over the B*T flat rows r = b*T + t it is trunc chained
// [rows x 32] x [32 x 32] products plus one long reduction — tensor-core work.  rnn_lags_tc_kernel, per tile of 128 rows
// (thread = row = TMEM lane):
//   * S_0 = the pre tile (TMA, 128B swizzle); for each lag the thread scales its row by g = 1 - h(row-k-1)^2 (h tile by
//     TMA with an 8-row halo), writes it as the K-major A operand, ONE thread issues 4 tcgen05.mma.kind::tf32 (M = 128,
//     N = 32, K = 8) against the resident Waa^T, and the row comes back through tcgen05.ld;
//   * lag sum: both gradients pair S_k(t) with data of row t - k, so W(u) = sum_k S_k(u + k) is accumulated in the pre tile's
//     own rows (a 128-bit read-modify-write of row t - k - 1 per lag, one thread per row, lags separated by the block barrier),
//     transposed once per tile (Wt[i][u]), and the reduction over rows is ONE M = 64 MMA chain per tile with K = rows:
//     D2[m][i] += sum_u HXt[m][u] * Wt[i][u],  HXt rows = h(u-1) (32) | x(u) (D) | ones: gWaa^T, gWax^T and gba accumulate in
//     TMEM over all tiles of the CTA and leave by atomics at the end;
//   * a tile OWNS its first 128 - trunc rows (its W(u) is complete there); tiles advance by that many rows, the lag chain
//     of the halo rows is recomputed by the next tile.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {
namespace tc {

constexpr int RL_THREADS = 128, RL_HALO = 8;

struct RnnLagsParams {
    const float* x; const float* waa;
    float* gwaa; float* gwax; float* gba;
    long R;                       // B * T flat rows
    int T, D, trunc, step;        // step = 128 - trunc rows owned per tile
    long ntiles;
};

__device__ __forceinline__ float4 rl_lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float rl_lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
// element (row m, k index u) of a K-major, 128B-swizzled operand with K = 128 (4 blocks of 32 floats, `rows` rows each)
__device__ __forceinline__ uint32_t rl_kaddr(uint32_t base, int rows, int m, int u) {
    const int uu = u & 31;
    return base + (uint32_t)((u >> 5) * rows * 128 + m * 128 + ((((uu >> 2) ^ (m & 7))) << 4) + (uu & 3) * 4);
}

__global__ void __launch_bounds__(RL_THREADS)
rnn_lags_tc_kernel(const __grid_constant__ CUtensorMap tmPre, const __grid_constant__ CUtensorMap tmH, RnnLagsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sPre = smem;                          // [128][32] f32, K-major SW128 (TMA)
    uint8_t* sA = sPre + 16384;                    // [128][32] the lag's A operand
    uint8_t* sH = sA + 16384;                      // [136][32] states of rows r0 - 8 .. r0 + 127 (TMA)
    uint8_t* sWt = sH + 18432;                     // [32][128] lag sums, transposed (B operand of the reduction)
    uint8_t* sHX = sWt + 16384;                    // [64][128] h(u-1) | x(u) | ones | 0 (A operand of the reduction)
    uint8_t* sW = sHX + 32768;                     // [32][32] Waa^T: row j' = Waa[:, j']
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 4096);
    uint64_t* ld_full = bars;
    uint64_t* mma1 = bars + 1;
    uint64_t* mma2 = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sPre_s = smem_u32(sPre), sA_s = smem_u32(sA), sH_s = smem_u32(sH), sWt_s = smem_u32(sWt), sHX_s = smem_u32(sHX);

    for (int e = tid; e < 32 * 32; e += RL_THREADS) {          // Waa^T, K-major SW128: element (n = j', k = q) = Waa[q][j']
        const int n = e >> 5, q = e & 31;
        sts32(smem_u32(sW) + (uint32_t)(n * 128 + ((((q >> 2) ^ (n & 7))) << 4) + (q & 3) * 4), __ldg(p.waa + q * 32 + n));
    }
    for (uint32_t e = tid; e < 32768 / 16; e += RL_THREADS) reinterpret_cast<float4*>(sHX)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tma_prefetch_desc(&tmPre);
        tma_prefetch_desc(&tmH);
        mbar_init(ld_full, 1); mbar_init(mma1, 1); mbar_init(mma2, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d1 = tmem_base, d2 = tmem_base + 32;
    const uint32_t idesc1 = make_idesc_tf32(128, 32, 0, 0), idesc2 = make_idesc_tf32(64, 32, 0, 0);
    const uint64_t a1_d = make_smem_desc(sA_s, 16, 1024), w_d = make_smem_desc(smem_u32(sW), 16, 1024);
    const uint64_t hx_d = make_smem_desc(sHX_s, 16, 1024), wt_d = make_smem_desc(sWt_s, 16, 1024);
    const uint32_t tq = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int xs = tid & 7;                                     // swizzle term of this thread's rows in sPre / sA

    int it = 0;
    uint32_t ph1 = 0;
    if (tid == 0 && (long)blockIdx.x < p.ntiles) {
        const long r0 = (long)blockIdx.x * p.step;
        mbar_arrive_expect_tx(ld_full, 16384 + (128 + RL_HALO) * 128);
        tma_load_2d(sPre, &tmPre, 0, (int)r0, ld_full);
        tma_load_2d(sH, &tmH, 0, (int)(r0 - RL_HALO), ld_full);
    }
    for (long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const long r0 = tile * p.step;
        const long row = r0 + tid;
        const bool valid = row < p.R;
        const int t = valid ? (int)(row % p.T) : 0;
        const bool owned = valid && tid < p.step;
        // x(row) straight from global while the tiles fly
        float xv[24];
#pragma unroll
        for (int d = 0; d < 24; ++d) xv[d] = (owned && d < p.D) ? __ldg(p.x + row * p.D + d) : 0.f;
        mbar_wait(ld_full, it & 1);
        // S_0 = the thread's row of the pre tile; the tile itself becomes the lag-sum rows W(u) (rows past the end are zero)
        const uint32_t myrow = sPre_s + (uint32_t)(tid * 128);
        float S[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = rl_lds128(myrow + ((c ^ xs) << 4));
            S[4 * c] = v.x; S[4 * c + 1] = v.y; S[4 * c + 2] = v.z; S[4 * c + 3] = v.w;
        }
        for (int k = 0; k < p.trunc; ++k) {
            // A = S_k * (1 - h(row-k-1)^2); a lag that does not exist (t - k - 1 < 0) contributes zeros from here on
            const bool live = valid && t - k - 1 >= 0;
            const int hr = tid + RL_HALO - k - 1;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 h = rl_lds128(sH_s + (uint32_t)(hr * 128 + ((c ^ (hr & 7)) << 4)));
                float4 a;
                a.x = live ? S[4 * c] * (1.f - h.x * h.x) : 0.f;
                a.y = live ? S[4 * c + 1] * (1.f - h.y * h.y) : 0.f;
                a.z = live ? S[4 * c + 2] * (1.f - h.z * h.z) : 0.f;
                a.w = live ? S[4 * c + 3] * (1.f - h.w * h.w) : 0.f;
                sts128(sA_s + (uint32_t)(tid * 128 + ((c ^ xs) << 4)), a);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();                                     // also orders the lag-sum updates of lag k - 1 before those of lag k
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ks == 0) mma_tf32_init(d1, a1_d, w_d, idesc1);
                    else mma_tf32_acc(d1, a1_d + 2 * ks, w_d + 2 * ks, idesc1);
                }
                mma_commit(mma1);
            }
            mbar_wait_spin(mma1, ph1);                           // a few hundred clocks: poll, do not suspend
            ph1 ^= 1;
            tc_fence_after();
            tmem_ld_32x32(tq + (d1 - tmem_base), reinterpret_cast<uint32_t (&)[32]>(S));
            tmem_ld_wait();
            tc_fence_before();
            // lag sum: S_{k+1}(row) belongs to W(row - k - 1): a 128-bit read-modify-write of that row (one thread per row)
            const int u = tid - k - 1;
            if (u >= 0 && live) {
                const uint32_t ur = sPre_s + (uint32_t)(u * 128);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t a = ur + ((c ^ (u & 7)) << 4);
                    float4 w = rl_lds128(a);
                    w.x += S[4 * c]; w.y += S[4 * c + 1]; w.z += S[4 * c + 2]; w.w += S[4 * c + 3];
                    sts128(a, w);
                }
            }
        }
        __syncthreads();                                         // every W(u) is complete
        if (it > 0) mbar_wait(mma2, (it - 1) & 1);               // the previous reduction has read sWt / sHX
        // reduction operands (K = rows, transposed): Wt[i][u] = W(u)[i]; HXt[q][u] = h(u-1)[q] (zero before the sequence starts),
        // HXt[32 + d][u] = x(u)[d], HXt[32 + D][u] = 1; columns a tile does not own are zero in HXt
        {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 w = rl_lds128(myrow + ((c ^ xs) << 4));
                sts32(rl_kaddr(sWt_s, 32, 4 * c, tid), w.x);
                sts32(rl_kaddr(sWt_s, 32, 4 * c + 1, tid), w.y);
                sts32(rl_kaddr(sWt_s, 32, 4 * c + 2, tid), w.z);
                sts32(rl_kaddr(sWt_s, 32, 4 * c + 3, tid), w.w);
            }
            const int hr = tid + RL_HALO - 1;
            const bool hp = owned && t >= 1;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 h = rl_lds128(sH_s + (uint32_t)(hr * 128 + ((c ^ (hr & 7)) << 4)));
                sts32(rl_kaddr(sHX_s, 64, 4 * c, tid), hp ? h.x : 0.f);
                sts32(rl_kaddr(sHX_s, 64, 4 * c + 1, tid), hp ? h.y : 0.f);
                sts32(rl_kaddr(sHX_s, 64, 4 * c + 2, tid), hp ? h.z : 0.f);
                sts32(rl_kaddr(sHX_s, 64, 4 * c + 3, tid), hp ? h.w : 0.f);
            }
#pragma unroll
            for (int d = 0; d < 24; ++d)
                if (d < p.D) sts32(rl_kaddr(sHX_s, 64, 32 + d, tid), xv[d]);
            sts32(rl_kaddr(sHX_s, 64, 32 + p.D, tid), owned ? 1.f : 0.f);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();                                         // sPre / sH are free: the next tile's loads overlap the reduction
        if (tid == 0) {
            const long nt = tile + gridDim.x;
            if (nt < p.ntiles) {
                const long n0 = nt * p.step;
                mbar_arrive_expect_tx(ld_full, 16384 + (128 + RL_HALO) * 128);
                tma_load_2d(sPre, &tmPre, 0, (int)n0, ld_full);
                tma_load_2d(sH, &tmH, 0, (int)(n0 - RL_HALO), ld_full);
            }
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const uint64_t a = hx_d + (uint64_t)((ks >> 2) * (8192 >> 4) + (ks & 3) * 2);
                const uint64_t b = wt_d + (uint64_t)((ks >> 2) * (4096 >> 4) + (ks & 3) * 2);
                mma_tf32(d2, a, b, idesc2, (it > 0 || ks > 0) ? 1u : 0u);
            }
            mma_commit(mma2);
        }
    }
    if (it > 0) {
        // M = 64 accumulator: row m lives in TMEM lane (m / 16) * 32 + m % 16; D2[m][i]: m < 32 -> gWaa[i][m], 32 <= m < 32 + D ->
        // gWax[i][m - 32], m = 32 + D -> gba[i]
        mbar_wait(mma2, (it - 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tq + (d2 - tmem_base), v);
        tmem_ld_wait();
        const int m = warp * 16 + lane;
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float g = __uint_as_float(v[i]);
                if (m < 32) atomicAdd(p.gwaa + i * 32 + m, g);
                else if (m < 32 + p.D) atomicAdd(p.gwax + i * p.D + (m - 32), g);
                else if (m == 32 + p.D) atomicAdd(p.gba + i, g);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// The serial part: pre(t) for every (b, t) and dx.  A warp owns a sample (H, D <= 32): lane j keeps column j of Waa and of
// Wax in registers, the vector of the step goes through a per-warp shared scratch (broadcast LDS.128); the states of the
// next RC_PF steps wait in a register ring (their global loads are issued RC_PF steps ahead of the serial chain).
constexpr int RC_WARPS = 8, RC_PF = 8;
__global__ void __launch_bounds__(RC_WARPS * 32)
rnn_chain_kernel(const float* __restrict__ states, const float* __restrict__ dy, const float* __restrict__ waa,
                 const float* __restrict__ wax, float* __restrict__ pre_out, float* __restrict__ dx, int B, int T, int D, int H) {
    __shared__ __align__(16) float scratch[RC_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RC_WARPS + warp;
    if (b >= B) return;
    float wcol[32], xcol[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        wcol[q] = (q < H && lane < H) ? __ldg(waa + q * H + lane) : 0.f;      // Waa[q][lane]
        xcol[q] = (q < H && lane < D) ? __ldg(wax + q * D + lane) : 0.f;      // Wax[q][lane]
    }
    float* su = scratch[warp];
    auto dot32 = [&](const float (&col)[32]) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
            const float4 t = *reinterpret_cast<const float4*>(su + 4 * q4);
            c0 = fmaf(t.x, col[4 * q4], c0); c1 = fmaf(t.y, col[4 * q4 + 1], c1);
            c2 = fmaf(t.z, col[4 * q4 + 2], c2); c3 = fmaf(t.w, col[4 * q4 + 3], c3);
        }
        return (c0 + c1) + (c2 + c3);
    };
    const float* hs = states + (size_t)b * T * H;
    float hq[RC_PF];
#pragma unroll
    for (int j = 0; j < RC_PF; ++j) hq[j] = (T - 1 - j >= 0 && lane < H) ? __ldg(hs + (size_t)(T - 1 - j) * H + lane) : 0.f;
    float dh = (lane < H) ? __ldg(dy + (size_t)b * H + lane) : 0.f;          // delta_y enters at t = T - 1 only
    for (int t = T - 1; t >= 0; --t) {
        const float h = hq[0];
#pragma unroll
        for (int j = 0; j + 1 < RC_PF; ++j) hq[j] = hq[j + 1];
        hq[RC_PF - 1] = (t - RC_PF >= 0 && lane < H) ? __ldg(hs + (size_t)(t - RC_PF) * H + lane) : 0.f;
        const float pre = dh * (1.f - h * h);                                // recurrent.py:226
        if (lane < H) pre_out[((size_t)b * T + t) * H + lane] = pre;
        __syncwarp();
        su[lane] = pre;
        __syncwarp();
        dh = dot32(wcol);                                                    // recurrent.py:251
        const float dxv = dot32(xcol);                                       // recurrent.py:252
        if (lane < D) dx[((size_t)b * T + t) * D + lane] = dxv;
    }
}

// Blocked form of the chain for D <= 8 (see rnn_fwd_blk_kernel in embed_rnn.cu): lane (a, b) = (lane / 4, lane % 4) owns the
// 8 x 4 block of Waa (inputs 8b .. 8b+7, units 4a .. 4a+3) and the 8 inputs' weights of dx column a; 2 LDS.128 + 10 SHFL per
// step instead of 16 broadcast LDS.128.
// The states arrive by per-warp bulk copies (chunks of RCB_CH steps, walking backwards, double buffered): a register ring of
// global loads stalls its consumer on the long scoreboard every step (see rnn_fwd_blk_kernel).
constexpr int RCB_WARPS = 4, RCB_CH = 32;
__global__ void __launch_bounds__(RCB_WARPS * 32)
rnn_chain_blk_kernel(const float* __restrict__ states, const float* __restrict__ dy, const float* __restrict__ waa,
                     const float* __restrict__ wax, float* __restrict__ pre_out, float* __restrict__ dx, int B, int T, int D) {
    __shared__ __align__(16) float scratch[RCB_WARPS][32];
    __shared__ __align__(16) float shs[RCB_WARPS][2][RCB_CH * 32];
    __shared__ __align__(8) uint64_t full[RCB_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RCB_WARPS + warp;
    if (b >= B) return;
    const int a = lane >> 2, bq = lane & 3;
    float w[8][4], wxx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) w[i][jj] = __ldg(waa + (8 * bq + i) * 32 + 4 * a + jj);     // Waa[q][j]
        wxx[i] = a < D ? __ldg(wax + (8 * bq + i) * D + a) : 0.f;                                   // Wax[q][d = a]
    }
    float* su = scratch[warp];
    const float* hs = states + (size_t)b * T * 32;
    const int nch = (T + RCB_CH - 1) / RCB_CH;
    // chunk c (counted from the END of the sequence) = steps [lo, lo + n): lo = max(0, T - (c + 1) * RCB_CH)
    auto fetch = [&](int c) {
        const int hi = T - c * RCB_CH, lo = max(0, hi - RCB_CH);
        const uint32_t bytes = (uint32_t)(hi - lo) * 128u;
        tc::mbar_arrive_expect_tx(&full[warp][c & 1], bytes);
        tc::bulk_load_1d(shs[warp][c & 1], hs + (size_t)lo * 32, bytes, &full[warp][c & 1]);
    };
    if (lane == 0) {
        tc::mbar_init(&full[warp][0], 1); tc::mbar_init(&full[warp][1], 1);
        tc::fence_barrier_init();
        fetch(0);
    }
    __syncwarp();
    float dh = __ldg(dy + (size_t)b * 32 + lane);                            // delta_y enters at t = T - 1 only
    for (int t = T - 1; t >= 0; --t) {
        const int c = (T - 1 - t) / RCB_CH;
        const int lo = max(0, T - (c + 1) * RCB_CH);
        if (t == T - 1 - c * RCB_CH) {
            if (lane == 0 && c + 1 < nch) fetch(c + 1);
            tc::mbar_wait_spin(&full[warp][c & 1], (c >> 1) & 1);
        }
        const float h = shs[warp][c & 1][(t - lo) * 32 + lane];
        const float pre = dh * (1.f - h * h);                                // recurrent.py:226
        pre_out[((size_t)b * T + t) * 32 + lane] = pre;
        __syncwarp();
        su[lane] = pre;
        __syncwarp();
        const float4 v0 = *reinterpret_cast<const float4*>(su + 8 * bq), v1 = *reinterpret_cast<const float4*>(su + 8 * bq + 4);
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float s[5];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float p0 = v[0] * w[0][jj], p1 = v[1] * w[1][jj];
#pragma unroll
            for (int i = 2; i < 8; i += 2) { p0 = fmaf(v[i], w[i][jj], p0); p1 = fmaf(v[i + 1], w[i + 1][jj], p1); }
            s[jj] = p0 + p1;
        }
        {
            float p0 = v[0] * wxx[0], p1 = v[1] * wxx[1];
#pragma unroll
            for (int i = 2; i < 8; i += 2) { p0 = fmaf(v[i], wxx[i], p0); p1 = fmaf(v[i + 1], wxx[i + 1], p1); }
            s[4] = p0 + p1;
        }
        // transpose-reduce over the 4 lanes of the unit block (see rnn_fwd_blk_kernel); the dx column needs the full sum
        const bool o1 = bq & 1, o2 = bq & 2;
        const float r0 = (o1 ? s[1] : s[0]) + __shfl_xor_sync(0xffffffffu, o1 ? s[0] : s[1], 1);
        const float r1 = (o1 ? s[3] : s[2]) + __shfl_xor_sync(0xffffffffu, o1 ? s[2] : s[3], 1);
        dh = (o2 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, o2 ? r0 : r1, 2);  // recurrent.py:251: unit 4a + bq = lane
        s[4] += __shfl_xor_sync(0xffffffffu, s[4], 1);
        s[4] += __shfl_xor_sync(0xffffffffu, s[4], 2);
        if (bq == 0 && a < D) dx[((size_t)b * T + t) * D + a] = s[4];        // recurrent.py:252
    }
}

static bool rl_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t gdim[2] = {32, rows}, gstr[1] = {128};
    cuuint32_t bx[2] = {32, box_rows}, es[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

using namespace tc;

extern "C" {

int dnb_rnn_bwd_tc_supported(int T, int D, int H, int trunc, int math) {
    if (math < DNB_MATH_TF32 || !get_encode_tiled()) return 0;
    const int lags = std::min(T - 1, trunc);                   // recurrences per step
    return (H == 32 && D >= 1 && D <= 24 && T >= 1 && lags >= 0 && lags <= 7) ? 1 : 0;
}

size_t dnb_rnn_bwd_tc_ws_bytes(int B, int T, int H) { return ((size_t)B * T * H * sizeof(float) + 255) & ~(size_t)255; }

int dnb_rnn_bwd_tc(const float* x, const float* states, const float* dy, const float* waa, const float* wax,
                   float* dx, float* gwaa, float* gwax, float* gba,
                   int B, int T, int D, int H, int trunc, void* ws, int math, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && T > 0 && D > 0 && H > 0 && trunc >= 0, "rnn_bwd_tc: bad shape");
    DNB_CHECK_ARG(dnb_rnn_bwd_tc_supported(T, D, H, trunc, math), "rnn_bwd_tc: unsupported shape / math mode");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && states && dy && waa && wax && dx && gwaa && gwax && gba && ws, "rnn_bwd_tc: null pointer");
    DNB_CHECK_ARG(!((reinterpret_cast<uintptr_t>(states) | reinterpret_cast<uintptr_t>(ws)) & 15), "rnn_bwd_tc: buffers must be 16-byte aligned");
    cudaStream_t st = S(stream);
    float* pre = static_cast<float*>(ws);
    if (D <= 8 && !getenv("DNB_RNN_NO_BLK"))
        rnn_chain_blk_kernel<<<(B + RCB_WARPS - 1) / RCB_WARPS, RCB_WARPS * 32, 0, st>>>(states, dy, waa, wax, pre, dx, B, T, D);
    else
        rnn_chain_kernel<<<(B + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, st>>>(states, dy, waa, wax, pre, dx, B, T, D, H);
    DNB_LAUNCH_CHECK("rnn_bwd_chain");
    RnnLagsParams p{};
    p.x = x; p.waa = waa; p.gwaa = gwaa; p.gwax = gwax; p.gba = gba;
    p.R = (long)B * T; p.T = T; p.D = D; p.trunc = std::min(T - 1, trunc); p.step = 128 - p.trunc;
    p.ntiles = (p.R + p.step - 1) / p.step;
    DNB_CHECK_ARG(p.R < (1L << 31) - 256, "rnn_bwd_tc: B*T too large");
    CUtensorMap tp, th;
    if (!rl_tmap(&tp, pre, (uint64_t)p.R, 128) || !rl_tmap(&th, states, (uint64_t)p.R, 128 + RL_HALO))
        return set_err(DNB_ERR_CUDA, "rnn_bwd_tc: cuTensorMapEncodeTiled failed");
    const size_t smem = 1024 + 16384 * 2 + 18432 + 16384 + 32768 + 4096 + 64;
    cudaFuncSetAttribute(rnn_lags_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long grid = std::min<long>(p.ntiles, 2L * kNumSMs);
    rnn_lags_tc_kernel<<<(unsigned)grid, RL_THREADS, smem, st>>>(tp, th, p);
    DNB_LAUNCH_CHECK("rnn_bwd_lags_tc");
    return DNB_OK;
}

}  // extern "C"

}  // namespace dnb
