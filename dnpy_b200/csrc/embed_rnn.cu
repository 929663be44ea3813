// Embedding (dnpy/layers/core.py:54-65) and SimpleRNN (dnpy/layers/recurrent.py:12-23,153-266).
//
// Embedding.backward, literal (Q15): mean over the batch first, then NumPy's fancy-index "+="
// (gather, add, scatter), so for every vocabulary row the LAST (b,l) position in row-major order
// decides which sequence slot's mean-delta is added — reproduced with an atomicMax over positions.
//
// SimpleRNN: both directions are persistent kernels — a CTA owns a tile of samples, keeps Waa/Wax
// in shared memory and walks all T timesteps (and, in backward, the (trunc+1)-step inner BPTT
// loop) without returning to the host.  Backward is the reference's literal recurrence (Q16):
// full-length dh chain PLUS a truncated inner loop at every t, (1-h^2) applied before the Waa
// product, gradients summed (not averaged) over the batch.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace dnb {

__global__ void __launch_bounds__(256) embedding_fwd_kernel(const int32_t* __restrict__ idx, const float* __restrict__ w,
                                                            float* __restrict__ y, long n_tok, int V, int D) {
    const long total = n_tok * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long tok = i / D; int d = (int)(i % D);
        int v = idx[tok];
        y[i] = (v >= 0 && v < V) ? w[(long)v * D + d] : 0.f;
    }
}
// meand[l,d] = inv_m * sum_b dy[b,l,d] in two deterministic stages: EMB_BS batch slices (rows by, by + EMB_BS, ...) summed by
// one CTA row each with four independent accumulators, then the slices added in a fixed order (one thread per column looping
// over all B rows ran on 8 CTAs: 75 us at B=1024, L*D=2048)
constexpr int EMB_BS = 32;
__global__ void __launch_bounds__(256) embedding_mean_partial_kernel(const float* __restrict__ dy, float* __restrict__ part,
                                                                     int B, long LD) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LD) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = blockIdx.y;
    for (; b + 3 * EMB_BS < B; b += 4 * EMB_BS) {
        a0 += dy[(long)b * LD + i]; a1 += dy[(long)(b + EMB_BS) * LD + i];
        a2 += dy[(long)(b + 2 * EMB_BS) * LD + i]; a3 += dy[(long)(b + 3 * EMB_BS) * LD + i];
    }
    for (; b < B; b += EMB_BS) a0 += dy[(long)b * LD + i];
    part[(long)blockIdx.y * LD + i] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(256) embedding_mean_finish_kernel(const float* __restrict__ part, float* __restrict__ meand,
                                                                    long LD, float inv_m) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LD) return;
    float acc = 0.f;
#pragma unroll 8
    for (int s = 0; s < EMB_BS; ++s) acc += part[(long)s * LD + i];
    meand[i] = acc * inv_m;
}
// last[v] = the largest GLOBAL flat position (pos0 + i) of token v: NumPy's fancy-index "+=" keeps the last write (Q15)
// Small vocabularies (V <= EMB_LAST_V): every CTA resolves its tokens in a shared-memory table first and issues ONE global
// atomicMax per vocabulary row it has seen (262 144 tokens on 1 000 rows: the global atomics serialised on their rows, 26 us).
constexpr int EMB_LAST_V = 8192;
__global__ void __launch_bounds__(256) embedding_last_kernel(const int32_t* __restrict__ idx, int32_t* __restrict__ last,
                                                             long n_tok, int V, int pos0) {
    extern __shared__ int32_t s_last[];
    const bool local = V <= EMB_LAST_V;
    if (local) {
        for (int v = threadIdx.x; v < V; v += blockDim.x) s_last[v] = -1;
        __syncthreads();
    }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tok; i += (long)gridDim.x * blockDim.x) {
        int v = idx[i];
        if (v >= 0 && v < V) {
            if (local) atomicMax(s_last + v, pos0 + (int)i);
            else atomicMax(last + v, pos0 + (int)i);
        }
    }
    if (local) {
        __syncthreads();
        for (int v = threadIdx.x; v < V; v += blockDim.x)
            if (s_last[v] >= 0) atomicMax(last + v, s_last[v]);
    }
}
__global__ void __launch_bounds__(256) embedding_apply_kernel(const int32_t* __restrict__ last, const float* __restrict__ meand,
                                                              float* __restrict__ gw, int V, int L, int D, float scale) {
    long total = (long)V * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int v = (int)(i / D), d = (int)(i % D);
        int pos = last[v];
        if (pos >= 0) gw[i] += scale * meand[(long)(pos % L) * D + d];
    }
}

// ------------------------------------------------------------------------------------------
// SimpleRNN forward: thread (s, j) owns hidden unit j of sample s of the CTA's tile.
// ------------------------------------------------------------------------------------------
__global__ void rnn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ waa, const float* __restrict__ wax,
                               const float* __restrict__ ba, float* __restrict__ states, float* __restrict__ y,
                               int B, int T, int D, int H, int SB) {
    extern __shared__ float sm[];
    float* s_waaT = sm;                    // [k][j]  (transposed: conflict-free across j)
    float* s_waxT = s_waaT + H * H;        // [d][j]
    float* s_h = s_waxT + D * H;           // [SB][H]
    float* s_x = s_h + SB * H;             // [SB][D]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < H * H; e += nt) { int j = e / H, k = e % H; s_waaT[k * H + j] = waa[e]; }
    for (int e = tid; e < H * D; e += nt) { int j = e / D, d = e % D; s_waxT[d * H + j] = wax[e]; }
    for (int e = tid; e < SB * H; e += nt) s_h[e] = 0.f;
    const int s = tid / H, j = tid % H;
    const int b = blockIdx.x * SB + s;
    const bool act = (s < SB) && (b < B);
    const float bj = act ? ba[j] : 0.f;
    for (int t = 0; t < T; ++t) {
        __syncthreads();
        for (int e = tid; e < SB * D; e += nt) {
            int ss = e / D, d = e % D, bb = blockIdx.x * SB + ss;
            s_x[e] = (bb < B) ? x[((long)bb * T + t) * D + d] : 0.f;
        }
        __syncthreads();
        float hn = 0.f;
        if (act) {
            float a = bj;
            const float* hp = s_h + s * H;
            for (int k = 0; k < H; ++k) a = fmaf(hp[k], s_waaT[k * H + j], a);
            const float* xp = s_x + s * D;
            for (int d = 0; d < D; ++d) a = fmaf(xp[d], s_waxT[d * H + j], a);
            hn = tanhf(a);
        }
        __syncthreads();
        if (act) {
            s_h[s * H + j] = hn;
            states[((long)b * T + t) * H + j] = hn;
            if (t == T - 1) y[(long)b * H + j] = hn;
        }
    }
}

// Warp-per-sample forward for H <= 32, D <= 32: lane j keeps rows j of Waa and Wax in registers, h_{t-1} and x_t are read
// from a per-warp shared scratch with broadcast LDS.128; one __syncwarp pair per time step instead of three block barriers.
constexpr int RNN_PF = 8;                    // prefetch distance (steps) of the streamed per-step operands
constexpr int RF_WARPS = 8;
__global__ void __launch_bounds__(RF_WARPS * 32)
rnn_fwd_warp_kernel(const float* __restrict__ x, const float* __restrict__ waa, const float* __restrict__ wax,
                    const float* __restrict__ ba, float* __restrict__ states, float* __restrict__ y, int B, int T, int D, int H) {
    __shared__ __align__(16) float scratch[RF_WARPS][2][32];    // per warp: h_{t-1} | x_t
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RF_WARPS + warp;
    if (b >= B) return;
    float wrow[32], xrow[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        wrow[k] = (lane < H && k < H) ? __ldg(waa + lane * H + k) : 0.f;
        xrow[k] = (lane < H && k < D) ? __ldg(wax + lane * D + k) : 0.f;
    }
    const float bj = lane < H ? __ldg(ba + lane) : 0.f;
    float* sh = scratch[warp][0];
    float* sx = scratch[warp][1];
    sh[lane] = 0.f;
    const float* xs = x + (size_t)b * T * D;
    // the inputs of the next RNN_PF steps wait in a register ring: a global load is issued RNN_PF steps before its use (one
    // step ahead is not enough: an HBM round trip is several steps of this chain).  Measured at B=1024, T=256: 90 -> 79 us
    float xq[RNN_PF];
#pragma unroll
    for (int j = 0; j < RNN_PF; ++j) xq[j] = (j < T && lane < D) ? __ldg(xs + (size_t)j * D + lane) : 0.f;
    for (int t = 0; t < T; ++t) {
        sx[lane] = xq[0];
        __syncwarp();
#pragma unroll
        for (int j = 0; j + 1 < RNN_PF; ++j) xq[j] = xq[j + 1];
        xq[RNN_PF - 1] = (t + RNN_PF < T && lane < D) ? __ldg(xs + (size_t)(t + RNN_PF) * D + lane) : 0.f;
        float a0 = bj, a1 = 0.f, a2 = 0.f, a3 = 0.f;            // four independent chains: the sum is on the critical path
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
            const float4 h4 = *reinterpret_cast<const float4*>(sh + 4 * q4);
            a0 = fmaf(h4.x, wrow[4 * q4], a0); a1 = fmaf(h4.y, wrow[4 * q4 + 1], a1);
            a2 = fmaf(h4.z, wrow[4 * q4 + 2], a2); a3 = fmaf(h4.w, wrow[4 * q4 + 3], a3);
        }
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
            if (4 * q4 < D) {
                const float4 x4 = *reinterpret_cast<const float4*>(sx + 4 * q4);
                a0 = fmaf(x4.x, xrow[4 * q4], a0); a1 = fmaf(x4.y, xrow[4 * q4 + 1], a1);
                a2 = fmaf(x4.z, xrow[4 * q4 + 2], a2); a3 = fmaf(x4.w, xrow[4 * q4 + 3], a3);
            }
        }
        const float hn = lane < H ? tanhf((a0 + a1) + (a2 + a3)) : 0.f;
        __syncwarp();
        sh[lane] = hn;
        if (lane < H) {
            states[((size_t)b * T + t) * H + lane] = hn;
            if (t == T - 1) y[(size_t)b * H + lane] = hn;
        }
    }
}

// Blocked form for H = 32, D <= 8.  The kernel above broadcasts the whole 32-vector to every lane (8 + 2 LDS.128 of 512 B
// register write-back each per step and warp): the LSU return path (128 B/clk per SM) bounds it at ~580 clk per step with 8
// warps per SM.  Here lane (a, b) = (lane / 4, lane % 4) owns the 4 x 8 block of Waa (units 4a .. 4a+3, inputs 8b .. 8b+7):
// it reads only ITS 8 inputs (2 LDS.128), the four lanes of a unit block add their partial sums with two xor-shuffle rounds
// (8 SHFL of 128 B), and lane 4a + b keeps unit 4a + b — the natural one-unit-per-lane layout for tanh, the store and the
// next step's scratch.  Half the LSU traffic per step.
// The inputs do NOT travel through a register ring here: ncu showed the ring's consumer stalled on the long scoreboard every
// step (a scoreboard wait covers ALL loads in flight on it, the newest included — the "prefetch" exposed one full L2 / HBM
// round trip per step, ~600 clk, in every form of this kernel).  x arrives by per-warp bulk copies (chunks of RB_CH steps,
// double buffered, mbarrier) and is read from shared memory.
constexpr int RB_CH = 64;
// tanh(v) = 1 - 2 / (exp(2 v) + 1) on the SFU (ex2.approx + rcp.approx, ~2 ulp of the result; saturates cleanly at +-1):
// tanhf() is ~35 instructions with a branch — a third of this kernel's issue slots at 2 warps per scheduler
__device__ __forceinline__ float rnn_tanh(float v) { return 1.f - __fdividef(2.f, __expf(2.f * v) + 1.f); }
__global__ void __launch_bounds__(RF_WARPS * 32)
rnn_fwd_blk_kernel(const float* __restrict__ x, const float* __restrict__ waa, const float* __restrict__ wax,
                   const float* __restrict__ ba, float* __restrict__ states, float* __restrict__ y, int B, int T, int D) {
    __shared__ __align__(16) float scratch[RF_WARPS][32];
    __shared__ __align__(16) float sx[RF_WARPS][2][RB_CH * 8];
    __shared__ __align__(8) uint64_t full[RF_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RF_WARPS + warp;
    if (b >= B) return;
    const int a = lane >> 2, bq = lane & 3;
    float w[8][4], wx[2][4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i][jj] = __ldg(waa + (4 * a + jj) * 32 + 8 * bq + i);      // Waa[j][q]
#pragma unroll
        for (int e = 0; e < 2; ++e) wx[e][jj] = (2 * bq + e < D) ? __ldg(wax + (4 * a + jj) * D + 2 * bq + e) : 0.f;
    }
    const float bj = __ldg(ba + lane);
    float* sh = scratch[warp];
    sh[lane] = 0.f;
    const float* xs = x + (size_t)b * T * D;
    const int nch = (T + RB_CH - 1) / RB_CH;
    auto fetch = [&](int c) {                               // lane 0: chunk c of the sample's inputs -> buffer c & 1
        const int steps = min(RB_CH, T - c * RB_CH);
        const uint32_t bytes = (uint32_t)steps * D * 4u;
        tc::mbar_arrive_expect_tx(&full[warp][c & 1], bytes);
        tc::bulk_load_1d(sx[warp][c & 1], xs + (size_t)c * RB_CH * D, bytes, &full[warp][c & 1]);
    };
    if (lane == 0) {
        tc::mbar_init(&full[warp][0], 1); tc::mbar_init(&full[warp][1], 1);
        tc::fence_barrier_init();
        fetch(0);
    }
    __syncwarp();
    const int xo0 = min(2 * bq, D - 1), xo1 = min(2 * bq + 1, D - 1);      // (the weights of columns >= D are zero)
    for (int t = 0; t < T; ++t) {
        const int c = t / RB_CH, i = t - c * RB_CH;
        if (i == 0) {
            if (lane == 0 && c + 1 < nch) fetch(c + 1);      // buffer (c + 1) & 1 was last read in chunk c - 1
            tc::mbar_wait_spin(&full[warp][c & 1], (c >> 1) & 1);
        }
        __syncwarp();
        const float4 v0 = *reinterpret_cast<const float4*>(sh + 8 * bq), v1 = *reinterpret_cast<const float4*>(sh + 8 * bq + 4);
        const float x0 = sx[warp][c & 1][i * D + xo0], x1 = sx[warp][c & 1][i * D + xo1];
        float s[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {                     // two chains per unit: the sum is on the critical path
            float p0 = x0 * wx[0][jj], p1 = x1 * wx[1][jj];
            p0 = fmaf(v0.x, w[0][jj], p0); p1 = fmaf(v0.y, w[1][jj], p1);
            p0 = fmaf(v0.z, w[2][jj], p0); p1 = fmaf(v0.w, w[3][jj], p1);
            p0 = fmaf(v1.x, w[4][jj], p0); p1 = fmaf(v1.y, w[5][jj], p1);
            p0 = fmaf(v1.z, w[6][jj], p0); p1 = fmaf(v1.w, w[7][jj], p1);
            s[jj] = p0 + p1;
        }
        // transpose-reduce over the 4 lanes of the unit block: after round 1 a lane keeps the two units of its parity, after
        // round 2 its own unit 4a + bq = lane (3 SHFL + 6 selects instead of 8 SHFL and a 4-way pick)
        const bool o1 = bq & 1, o2 = bq & 2;
        const float r0 = (o1 ? s[1] : s[0]) + __shfl_xor_sync(0xffffffffu, o1 ? s[0] : s[1], 1);    // unit 0 / 1
        const float r1 = (o1 ? s[3] : s[2]) + __shfl_xor_sync(0xffffffffu, o1 ? s[2] : s[3], 1);    // unit 2 / 3
        const float mine = (o2 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, o2 ? r0 : r1, 2);
        const float hn = rnn_tanh(mine + bj);
        __syncwarp();
        sh[lane] = hn;
        states[((size_t)b * T + t) * 32 + lane] = hn;
        if (t == T - 1) y[(size_t)b * 32 + lane] = hn;
    }
}

// ------------------------------------------------------------------------------------------
// SimpleRNN backward (literal truncated BPTT)
// ------------------------------------------------------------------------------------------
__global__ void rnn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ states, const float* __restrict__ dy,
                               const float* __restrict__ waa, const float* __restrict__ wax,
                               float* __restrict__ dx, float* __restrict__ gwaa, float* __restrict__ gwax, float* __restrict__ gba,
                               int B, int T, int D, int H, int SB, int trunc) {
    extern __shared__ float sm[];
    float* s_waa = sm;                     // [j][j']
    float* s_wax = s_waa + H * H;          // [j][d]
    float* a_waa = s_wax + H * D;          // accumulators
    float* a_wax = a_waa + H * H;
    float* a_ba = a_wax + H * D;
    float* s_S = a_ba + H;                 // [SB][H] running inner-loop vector
    float* s_S2 = s_S + SB * H;
    float* s_pre = s_S2 + SB * H;
    float* s_dh = s_pre + SB * H;
    float* s_hp = s_dh + SB * H;           // h_{k-1}
    float* s_xk = s_hp + SB * H;           // [SB][D]
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < H * H; e += nt) { s_waa[e] = waa[e]; a_waa[e] = 0.f; }
    for (int e = tid; e < H * D; e += nt) { s_wax[e] = wax[e]; a_wax[e] = 0.f; }
    for (int e = tid; e < H; e += nt) a_ba[e] = 0.f;
    for (int e = tid; e < SB * H; e += nt) s_dh[e] = 0.f;
    const int s = tid / H, j = tid % H;
    const int b0 = blockIdx.x * SB;
    const int b = b0 + s;
    const bool act = (s < SB) && (b < B);
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        // pre_t = (dy*[t==T-1] + dh) * (1 - h_t^2)
        if (act) {
            float h = states[((long)b * T + t) * H + j];
            float d = (t == T - 1) ? dy[(long)b * H + j] : 0.f;
            float pre = (d + s_dh[s * H + j]) * (1.f - h * h);
            s_pre[s * H + j] = pre;
            s_S[s * H + j] = pre;
        } else if (s < SB) { s_pre[s * H + j] = 0.f; s_S[s * H + j] = 0.f; }
        __syncthreads();
        const int k_lo = max(0, t - trunc);
        for (int k = t; k >= k_lo; --k) {
            // stage h_{k-1} and x_k for the tile
            for (int e = tid; e < SB * H; e += nt) {
                int ss = e / H, jj = e % H, bb = b0 + ss;
                s_hp[e] = (k > 0 && bb < B) ? states[((long)bb * T + (k - 1)) * H + jj] : 0.f;
            }
            for (int e = tid; e < SB * D; e += nt) {
                int ss = e / D, d = e % D, bb = b0 + ss;
                s_xk[e] = (bb < B) ? x[((long)bb * T + k) * D + d] : 0.f;
            }
            __syncthreads();
            // gradient accumulation: every accumulator entry is owned by exactly one thread
            for (int e = tid; e < H * H; e += nt) {
                int jj = e / H, kk = e % H;
                float acc = a_waa[e];
                for (int ss = 0; ss < SB; ++ss) acc = fmaf(s_S[ss * H + jj], s_hp[ss * H + kk], acc);
                a_waa[e] = acc;
            }
            for (int e = tid; e < H * D; e += nt) {
                int jj = e / D, d = e % D;
                float acc = a_wax[e];
                for (int ss = 0; ss < SB; ++ss) acc = fmaf(s_S[ss * H + jj], s_xk[ss * D + d], acc);
                a_wax[e] = acc;
            }
            for (int e = tid; e < H; e += nt) {
                float acc = a_ba[e];
                for (int ss = 0; ss < SB; ++ss) acc += s_S[ss * H + e];
                a_ba[e] = acc;
            }
            // S <- (S * (1 - h_{k-1}^2)) . Waa      (recurrent.py:240)
            if (s < SB) {
                float acc = 0.f;
                for (int q = 0; q < H; ++q) {
                    float hp = s_hp[s * H + q];
                    acc = fmaf(s_S[s * H + q] * (1.f - hp * hp), s_waa[q * H + j], acc);
                }
                s_S2[s * H + j] = acc;
            }
            __syncthreads();
            if (s < SB) s_S[s * H + j] = s_S2[s * H + j];
            __syncthreads();
        }
        // dh <- pre . Waa ; dx_t = pre . Wax
        if (s < SB) {
            float acc = 0.f;
            for (int q = 0; q < H; ++q) acc = fmaf(s_pre[s * H + q], s_waa[q * H + j], acc);
            s_dh[s * H + j] = acc;
        }
        for (int e = tid; e < SB * D; e += nt) {
            int ss = e / D, d = e % D, bb = b0 + ss;
            if (bb < B) {
                float acc = 0.f;
                for (int q = 0; q < H; ++q) acc = fmaf(s_pre[ss * H + q], s_wax[q * D + d], acc);
                dx[((long)bb * T + t) * D + d] = acc;
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < H * H; e += nt) atomicAdd(gwaa + e, a_waa[e]);
    for (int e = tid; e < H * D; e += nt) atomicAdd(gwax + e, a_wax[e]);
    for (int e = tid; e < H; e += nt) atomicAdd(gba + e, a_ba[e]);
}

// Warp-per-sample form of the same literal algorithm for H <= 32, D <= 32 (the sentiment model: H=32, D=8).  The tile kernel
// above walks T x (trunc+1) inner steps with several block barriers and dependent global loads each: 3.3 ms at B=1024,
// T=256 on B200.  Here a sample's vectors live one element per lane; matrix-vector products read the other lanes' elements
// from a per-warp shared scratch with broadcast LDS.128 (a quarter of the MIO operations of shuffles); the lane's
// columns of Waa / Wax and its rows of the gradient accumulators stay in registers; no block barrier in the time loop.
// LAGSUM (MAXL > 0, at most MAXL = trunc + 1 lags): the weight gradients of the literal truncated BPTT are linear in the lag
// vectors S_k(t), and both pair every S_k(t) with data of step t - k (h_{t-k-1} and x_{t-k}): with
//     W(u) = sum_k S_k(u + k)        gWaa = sum_u W(u) (x) h_{u-1},   gWax = sum_u W(u) (x) x_u,   gba = sum_u W(u)
// the two outer products run ONCE per step instead of once per (step, lag).  W lives in a register ring: lag k of step t adds
// into slot k, the ring shifts down once per step, slot 0 is complete when step t is done (t runs downwards).  The inner loop
// keeps only the recurrence S <- (S (1 - h^2)) . Waa: 8 broadcast LDS.128 + 32 FMAs instead of 18 + 72.
constexpr int RW_WARPS = 8;
template <int MAXL>
__global__ void __launch_bounds__(RW_WARPS * 32, 1)
rnn_bwd_warp_kernel(const float* __restrict__ x, const float* __restrict__ states, const float* __restrict__ dy,
                    const float* __restrict__ waa, const float* __restrict__ wax,
                    float* __restrict__ dx, float* __restrict__ gwaa, float* __restrict__ gwax, float* __restrict__ gba,
                    int B, int T, int D, int H, int trunc) {
    __shared__ __align__(16) float scratch[RW_WARPS][3][32];    // per warp: u = S*(1-h^2) | h_{k-1} | x_k  (also pre)
    __shared__ float red[32 * 32 + 32 * 32 + 32];               // CTA sums of gWaa | gWax | gba
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * RW_WARPS + warp;
    const bool act = b < B;
    float wcol[32], xcol[32], aw[32], ax[32], ab = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        wcol[q] = (q < H && lane < H) ? __ldg(waa + q * H + lane) : 0.f;      // Waa[q][lane]
        xcol[q] = (q < H && lane < D) ? __ldg(wax + q * D + lane) : 0.f;      // Wax[q][lane]
        aw[q] = 0.f; ax[q] = 0.f;
    }
    float* su = scratch[warp][0];
    float* sh = scratch[warp][1];
    float* sx = scratch[warp][2];
    // sum_q v[q] * col[q] with v read from shared memory (all lanes read the same addresses: broadcast)
    auto dot32 = [&](const float* v, const float (&col)[32]) {
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;           // four independent chains: the result is on the critical path
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
            const float4 t = *reinterpret_cast<const float4*>(v + 4 * q4);
            c0 = fmaf(t.x, col[4 * q4], c0); c1 = fmaf(t.y, col[4 * q4 + 1], c1);
            c2 = fmaf(t.z, col[4 * q4 + 2], c2); c3 = fmaf(t.w, col[4 * q4 + 3], c3);
        }
        return (c0 + c1) + (c2 + c3);
    };
    float dh = 0.f;
    float wr[MAXL > 0 ? MAXL : 1];                                  // the lag-sum ring W(t), W(t-1), ...
#pragma unroll
    for (int j = 0; j < (MAXL > 0 ? MAXL : 1); ++j) wr[j] = 0.f;
    if (act) {
        const float* hs = states + (size_t)b * T * H;
        const float* xs = x + (size_t)b * T * D;
        // LAGSUM: the states h_t .. h_{t-MAXL} of a step live in a register ring (ONE new global load per step, issued a whole
        // step ahead of its first use, like x_t and the next step's x): no global latency on the serial chain
        float hr[(MAXL > 0 ? MAXL : 0) + 1];
        float x_next = 0.f;
        if constexpr (MAXL > 0) {
#pragma unroll
            for (int j = 0; j <= MAXL; ++j) hr[j] = (T - 1 - j >= 0 && lane < H) ? __ldg(hs + (size_t)(T - 1 - j) * H + lane) : 0.f;
            x_next = lane < D ? __ldg(xs + (size_t)(T - 1) * D + lane) : 0.f;
        }
        for (int t = T - 1; t >= 0; --t) {
            float h;
            float h_new = 0.f, xk_cur = 0.f;
            if constexpr (MAXL > 0) {
                h = hr[0];
                xk_cur = x_next;
                h_new = (t - MAXL - 1 >= 0 && lane < H) ? __ldg(hs + (size_t)(t - MAXL - 1) * H + lane) : 0.f;   // first used next step
                x_next = (t > 0 && lane < D) ? __ldg(xs + (size_t)(t - 1) * D + lane) : 0.f;
            } else {
                h = lane < H ? __ldg(hs + (size_t)t * H + lane) : 0.f;
            }
            const float d0 = (t == T - 1 && lane < H) ? __ldg(dy + (size_t)b * H + lane) : 0.f;
            const float pre = (d0 + dh) * (1.f - h * h);
            float S = pre;
            const int k_lo = max(0, t - trunc);
            if constexpr (MAXL > 0) {
                const int nl = t - k_lo + 1;                       // lags of this step (<= MAXL)
#pragma unroll
                for (int j = 0; j < MAXL; ++j) {
                    if (j < nl) {
                        wr[j] += S;
                        if (j + 1 < nl) {                          // the recurrence is only needed while another lag follows
                            const float hp = hr[j + 1];           // h_{t-j-1} (0 before the sequence starts)
                            __syncwarp();
                            su[lane] = S * (1.f - hp * hp);
                            __syncwarp();
                            S = dot32(su, wcol);                   // S <- (S * (1 - h_{k-1}^2)) . Waa   (recurrent.py:240)
                        }
                    }
                }
                // W(t) = wr[0] is complete: gWaa[lane][q] += W(t) h_{t-1}[q], gWax[lane][d] += W(t) x_t[d], gba += W(t)
                const float wt = wr[0];
                __syncwarp();
                sh[lane] = hr[1];
                sx[lane] = xk_cur;
                __syncwarp();
                ab += wt;
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(sh + 4 * q4);
                    aw[4 * q4] = fmaf(wt, t4.x, aw[4 * q4]); aw[4 * q4 + 1] = fmaf(wt, t4.y, aw[4 * q4 + 1]);
                    aw[4 * q4 + 2] = fmaf(wt, t4.z, aw[4 * q4 + 2]); aw[4 * q4 + 3] = fmaf(wt, t4.w, aw[4 * q4 + 3]);
                }
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    if (4 * q4 < D) {
                        const float4 t4 = *reinterpret_cast<const float4*>(sx + 4 * q4);
                        ax[4 * q4] = fmaf(wt, t4.x, ax[4 * q4]); ax[4 * q4 + 1] = fmaf(wt, t4.y, ax[4 * q4 + 1]);
                        ax[4 * q4 + 2] = fmaf(wt, t4.z, ax[4 * q4 + 2]); ax[4 * q4 + 3] = fmaf(wt, t4.w, ax[4 * q4 + 3]);
                    }
                }
#pragma unroll
                for (int j = 0; j + 1 < MAXL; ++j) wr[j] = wr[j + 1];
                wr[MAXL - 1] = 0.f;
#pragma unroll
                for (int j = 0; j < MAXL; ++j) hr[j] = hr[j + 1];
                hr[MAXL] = h_new;
            } else {
            float hp_n = (t > 0 && lane < H) ? __ldg(hs + (size_t)(t - 1) * H + lane) : 0.f;
            float xk_n = lane < D ? __ldg(xs + (size_t)t * D + lane) : 0.f;
            for (int k = t; k >= k_lo; --k) {
                const float hp = hp_n, xk = xk_n;
                if (k > k_lo) {                                    // the next lag's operands, one lag ahead of their use
                    hp_n = (k > 1 && lane < H) ? __ldg(hs + (size_t)(k - 2) * H + lane) : 0.f;
                    xk_n = lane < D ? __ldg(xs + (size_t)(k - 1) * D + lane) : 0.f;
                }
                __syncwarp();
                su[lane] = S * (1.f - hp * hp);
                sh[lane] = hp;
                sx[lane] = xk;
                __syncwarp();
                ab += S;
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {                   // gWaa[lane][q] += S * h_{k-1}[q]
                    const float4 t4 = *reinterpret_cast<const float4*>(sh + 4 * q4);
                    aw[4 * q4] = fmaf(S, t4.x, aw[4 * q4]); aw[4 * q4 + 1] = fmaf(S, t4.y, aw[4 * q4 + 1]);
                    aw[4 * q4 + 2] = fmaf(S, t4.z, aw[4 * q4 + 2]); aw[4 * q4 + 3] = fmaf(S, t4.w, aw[4 * q4 + 3]);
                }
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {                   // gWax[lane][d] += S * x_k[d]
                    if (4 * q4 < D) {
                        const float4 t4 = *reinterpret_cast<const float4*>(sx + 4 * q4);
                        ax[4 * q4] = fmaf(S, t4.x, ax[4 * q4]); ax[4 * q4 + 1] = fmaf(S, t4.y, ax[4 * q4 + 1]);
                        ax[4 * q4 + 2] = fmaf(S, t4.z, ax[4 * q4 + 2]); ax[4 * q4 + 3] = fmaf(S, t4.w, ax[4 * q4 + 3]);
                    }
                }
                S = dot32(su, wcol);                               // S <- (S * (1 - h_{k-1}^2)) . Waa   (recurrent.py:240)
            }
            }
            // dh <- pre . Waa ; dx_t = pre . Wax
            __syncwarp();
            su[lane] = pre;
            __syncwarp();
            dh = dot32(su, wcol);
            const float dxv = dot32(su, xcol);
            if (lane < D) dx[((size_t)b * T + t) * D + lane] = dxv;
        }
    }
    // CTA reduction of the per-warp accumulators, then one atomic per entry
    for (int e = threadIdx.x; e < 32 * 32 + 32 * 32 + 32; e += RW_WARPS * 32) red[e] = 0.f;
    __syncthreads();
    for (int w = 0; w < RW_WARPS; ++w) {
        if (warp == w && act) {
#pragma unroll
            for (int q = 0; q < 32; ++q) { red[lane * 32 + q] += aw[q]; red[1024 + lane * 32 + q] += ax[q]; }
            red[2048 + lane] += ab;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < H * H; e += RW_WARPS * 32) atomicAdd(gwaa + e, red[(e / H) * 32 + e % H]);
    for (int e = threadIdx.x; e < H * D; e += RW_WARPS * 32) atomicAdd(gwax + e, red[1024 + (e / D) * 32 + e % D]);
    for (int e = threadIdx.x; e < H; e += RW_WARPS * 32) atomicAdd(gba + e, red[2048 + e]);
}

static int rnn_tile(int H, int* SB, int* threads) {
    int sb = 256 / H;
    if (sb < 1) sb = 1;
    if (sb > 8) sb = 8;
    *SB = sb;
    *threads = ((sb * H + 31) / 32) * 32;
    return 0;
}

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_embedding_fwd(const int32_t* idx, const float* w, float* y, int B, int L, int V, int D, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && L > 0 && V > 0 && D > 0, "embedding_fwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(idx && w && y, "embedding_fwd: null pointer");
    long n_tok = (long)B * L;
    embedding_fwd_kernel<<<ew_grid((size_t)n_tok * D, 256), 256, 0, S(stream)>>>(idx, w, y, n_tok, V, D);
    DNB_LAUNCH_CHECK("embedding_fwd");
    return DNB_OK;
}

size_t dnb_embedding_bwd_ws_bytes(int L, int V, int D) { return ((size_t)V + (size_t)(1 + EMB_BS) * L * D) * sizeof(float); }

// Two halves around the data-parallel exchange: prepare leaves the local part of the batch mean (inv_m = 1 over the
// GLOBAL batch) in ws[V ..] and the last global position of every vocabulary row in ws[0 .. V); between the halves the ranks
// all-reduce the means (SUM) and the positions (MAX); apply adds `scale` x the mean of the winning position (scale = 1/world:
// every rank then holds 1/world of the exact gradient and the arena all-reduce restores it).
int dnb_embedding_bwd_prepare(const int32_t* idx, const float* dy, int B, int L, int V, int D, float inv_m, int pos0,
                              void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && L > 0 && V > 0 && D > 0, "embedding_bwd: bad shape");
    DNB_CHECK_ARG(ws, "embedding_bwd: null workspace");
    cudaStream_t st = S(stream);
    int32_t* last = static_cast<int32_t*>(ws);
    float* meand = reinterpret_cast<float*>(last + V);
    cudaError_t e = cudaMemsetAsync(last, 0xff, (size_t)V * sizeof(int32_t), st);
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "embedding_bwd: %s", cudaGetErrorString(e));
    long LD = (long)L * D, n_tok = (long)B * L;
    if (B == 0) {
        e = cudaMemsetAsync(meand, 0, (size_t)LD * sizeof(float), st);
        return e == cudaSuccess ? DNB_OK : set_err(DNB_ERR_CUDA, "embedding_bwd: %s", cudaGetErrorString(e));
    }
    DNB_CHECK_ARG(idx && dy, "embedding_bwd: null pointer");
    float* part = meand + LD;
    embedding_mean_partial_kernel<<<dim3((unsigned)((LD + 255) / 256), EMB_BS), 256, 0, st>>>(dy, part, B, LD);
    DNB_LAUNCH_CHECK("embedding_bwd_mean");
    embedding_mean_finish_kernel<<<(unsigned)((LD + 255) / 256), 256, 0, st>>>(part, meand, LD, inv_m);
    DNB_LAUNCH_CHECK("embedding_bwd_mean_finish");
    {
        const bool local = V <= EMB_LAST_V;
        // a CTA walks a long slice of the tokens when the table is local (fewer CTAs = fewer global atomics per row)
        const int grid = local ? (int)std::min<long>((n_tok + 2047) / 2048, kNumSMs) : ew_grid((size_t)n_tok, 256);
        embedding_last_kernel<<<grid, 256, local ? (size_t)V * sizeof(int32_t) : 0, st>>>(idx, last, n_tok, V, pos0);
    }
    DNB_LAUNCH_CHECK("embedding_bwd_last");
    return DNB_OK;
}

int dnb_embedding_bwd_apply(const void* ws, float* gw, int L, int V, int D, float scale, dnb_stream_t stream) {
    DNB_CHECK_ARG(L > 0 && V > 0 && D > 0, "embedding_bwd: bad shape");
    DNB_CHECK_ARG(ws && gw, "embedding_bwd: null pointer");
    const int32_t* last = static_cast<const int32_t*>(ws);
    const float* meand = reinterpret_cast<const float*>(last + V);
    embedding_apply_kernel<<<ew_grid((size_t)V * D, 256), 256, 0, S(stream)>>>(last, meand, gw, V, L, D, scale);
    DNB_LAUNCH_CHECK("embedding_bwd_apply");
    return DNB_OK;
}

int dnb_embedding_bwd(const int32_t* idx, const float* dy, float* gw, int B, int L, int V, int D,
                      float inv_m, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && L > 0 && V > 0 && D > 0, "embedding_bwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(idx && dy && gw && ws, "embedding_bwd: null pointer");
    int rc = dnb_embedding_bwd_prepare(idx, dy, B, L, V, D, inv_m, 0, ws, stream);
    if (rc) return rc;
    return dnb_embedding_bwd_apply(ws, gw, L, V, D, 1.0f, stream);
}

int dnb_rnn_fwd(const float* x, const float* waa, const float* wax, const float* ba,
                float* states, float* y, int B, int T, int D, int H, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && T > 0 && D > 0 && H > 0, "rnn_fwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && waa && wax && ba && states && y, "rnn_fwd: null pointer");
    if (H > 256) return set_err(DNB_ERR_UNSUPPORTED, "rnn_fwd: hidden_dim %d > 256 not supported", H);
    int SB, threads;
    if (H == 32 && D <= 8 && (T * D) % 4 == 0 && !(reinterpret_cast<uintptr_t>(x) & 15) && !getenv("DNB_RNN_NO_BLK")) {
        rnn_fwd_blk_kernel<<<(B + RF_WARPS - 1) / RF_WARPS, RF_WARPS * 32, 0, S(stream)>>>(x, waa, wax, ba, states, y, B, T, D);
        DNB_LAUNCH_CHECK("rnn_fwd");
        return DNB_OK;
    }
    if (H <= 32 && D <= 32 && !getenv("DNB_RNN_TILE_FWD")) {
        rnn_fwd_warp_kernel<<<(B + RF_WARPS - 1) / RF_WARPS, RF_WARPS * 32, 0, S(stream)>>>(x, waa, wax, ba, states, y, B, T, D, H);
        DNB_LAUNCH_CHECK("rnn_fwd");
        return DNB_OK;
    }
    rnn_tile(H, &SB, &threads);
    size_t smem = ((size_t)H * H + (size_t)D * H + (size_t)SB * H + (size_t)SB * D) * sizeof(float);
    if (smem > 200 * 1024) return set_err(DNB_ERR_UNSUPPORTED, "rnn_fwd: H=%d D=%d exceed shared memory", H, D);
    cudaFuncSetAttribute(rnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rnn_fwd_kernel<<<(B + SB - 1) / SB, threads, smem, S(stream)>>>(x, waa, wax, ba, states, y, B, T, D, H, SB);
    DNB_LAUNCH_CHECK("rnn_fwd");
    return DNB_OK;
}

int dnb_rnn_bwd(const float* x, const float* states, const float* dy, const float* waa, const float* wax,
                float* dx, float* gwaa, float* gwax, float* gba,
                int B, int T, int D, int H, int trunc, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && T > 0 && D > 0 && H > 0 && trunc >= 0, "rnn_bwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && states && dy && waa && wax && dx && gwaa && gwax && gba, "rnn_bwd: null pointer");
    if (H > 256) return set_err(DNB_ERR_UNSUPPORTED, "rnn_bwd: hidden_dim %d > 256 not supported", H);
    if (H <= 32 && D <= 32 && !getenv("DNB_RNN_TILE_BWD")) {
        const int nl = std::min(T, trunc + 1);                        // lags per step
        const dim3 grid((B + RW_WARPS - 1) / RW_WARPS);
        if (nl <= 4 && !getenv("DNB_RNN_NO_LAGSUM"))
            rnn_bwd_warp_kernel<4><<<grid, RW_WARPS * 32, 0, S(stream)>>>(x, states, dy, waa, wax, dx, gwaa, gwax, gba, B, T, D, H, trunc);
        else if (nl <= 8 && !getenv("DNB_RNN_NO_LAGSUM"))
            rnn_bwd_warp_kernel<8><<<grid, RW_WARPS * 32, 0, S(stream)>>>(x, states, dy, waa, wax, dx, gwaa, gwax, gba, B, T, D, H, trunc);
        else
            rnn_bwd_warp_kernel<0><<<grid, RW_WARPS * 32, 0, S(stream)>>>(x, states, dy, waa, wax, dx, gwaa, gwax, gba, B, T, D, H, trunc);
        DNB_LAUNCH_CHECK("rnn_bwd");
        return DNB_OK;
    }
    int SB, threads;
    rnn_tile(H, &SB, &threads);
    size_t smem = (2 * ((size_t)H * H + (size_t)H * D) + H + 5 * (size_t)SB * H + (size_t)SB * D) * sizeof(float);
    if (smem > 200 * 1024) return set_err(DNB_ERR_UNSUPPORTED, "rnn_bwd: H=%d D=%d exceed shared memory", H, D);
    cudaFuncSetAttribute(rnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rnn_bwd_kernel<<<(B + SB - 1) / SB, threads, smem, S(stream)>>>(x, states, dy, waa, wax, dx, gwaa, gwax, gba,
                                                                     B, T, D, H, SB, trunc);
    DNB_LAUNCH_CHECK("rnn_bwd");
    return DNB_OK;
}

}  // extern "C"
