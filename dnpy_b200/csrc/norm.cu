// BatchNorm (dnpy/layers/regularization.py:94-156), literal semantics:
//   * statistics over axis 0 ONLY: x is viewed as [B, n_feat], one (mu, var) per feature, so for
//     conv inputs gamma/beta/moving_* are (C,H,W)-shaped (Q8);
//   * biased variance, std = sqrt(var + eps); moving = momentum*moving + (1-momentum)*stat;
//   * backward (Q7): dgamma = sum_b dy*MU (not x_hat), dX = (1/m)*STD*(m*dy*g - sum(dy*g) - x_hat*sum(dy*g*x_hat)).
// Column reductions: a CTA owns 32 adjacent features (128-byte coalesced rows) and 8 row lanes;
// per-lane Welford partials are merged with Chan's formula.  x_hat is recomputed in backward from
// (x, mu, std) instead of being cached, saving one full tensor write + read.  Roofline: HBM.
#include "common.cuh"

namespace dnb {

constexpr int BN_COLS = 32, BN_ROWS = 8;

__global__ void __launch_bounds__(256) bn_fwd_train_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ mmu,
                                                           float* __restrict__ mvar, float* __restrict__ smu,
                                                           float* __restrict__ sstd, float* __restrict__ y,
                                                           int B, size_t N, float momentum, float eps) {
    __shared__ float s_cnt[BN_ROWS][BN_COLS], s_mean[BN_ROWS][BN_COLS], s_m2[BN_ROWS][BN_COLS];
    __shared__ float s_mu[BN_COLS], s_rstd[BN_COLS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t col = (size_t)blockIdx.x * BN_COLS + tx;
    const bool ok = col < N;
    float cnt = 0.f, mean = 0.f, m2 = 0.f;
    if (ok)
        for (int r = ty; r < B; r += BN_ROWS) {
            float v = x[(size_t)r * N + col];
            cnt += 1.f;
            float d = v - mean;
            mean += d / cnt;
            m2 = fmaf(d, v - mean, m2);
        }
    s_cnt[ty][tx] = cnt; s_mean[ty][tx] = mean; s_m2[ty][tx] = m2;
    __syncthreads();
    if (ty == 0 && ok) {
        float n = s_cnt[0][tx], mu = s_mean[0][tx], M2 = s_m2[0][tx];
#pragma unroll
        for (int r = 1; r < BN_ROWS; ++r) {
            float nb = s_cnt[r][tx];
            if (nb > 0.f) {
                float d = s_mean[r][tx] - mu, nt = n + nb;
                mu += d * (nb / nt);
                M2 += s_m2[r][tx] + d * d * (n * nb / nt);
                n = nt;
            }
        }
        float var = M2 / (float)B;
        float sd = sqrtf(var + eps);
        smu[col] = mu; sstd[col] = sd;
        mmu[col] = momentum * mmu[col] + (1.0f - momentum) * mu;
        mvar[col] = momentum * mvar[col] + (1.0f - momentum) * var;
        s_mu[tx] = mu; s_rstd[tx] = 1.0f / sd;
    }
    __syncthreads();
    if (ok) {
        const float mu = s_mu[tx], rs = s_rstd[tx], g = gamma[col], bt = beta[col];
        for (int r = ty; r < B; r += BN_ROWS) {
            size_t i = (size_t)r * N + col;
            y[i] = fmaf(g, (x[i] - mu) * rs, bt);
        }
    }
}

// 128-bit variant for n_feat % 4 == 0: a thread owns FOUR adjacent features (a CTA 128 of them: 512 contiguous bytes per row
// instead of 128 — the 32-column version ran at 47% of HBM peak on the 131 072-feature planes of the residual net), four
// independent Welford chains per thread, two rows of loads in flight.
__global__ void __launch_bounds__(256) bn_fwd_train4_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ mmu,
                                                            float* __restrict__ mvar, float* __restrict__ smu,
                                                            float* __restrict__ sstd, float* __restrict__ y,
                                                            int B, size_t N, float momentum, float eps) {
    __shared__ float s_cnt[BN_ROWS][BN_COLS], s_mean[BN_ROWS][BN_COLS][4], s_m2[BN_ROWS][BN_COLS][4];
    __shared__ float s_mu[BN_COLS][4], s_rstd[BN_COLS][4];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t N4 = N >> 2, c4 = (size_t)blockIdx.x * BN_COLS + tx;
    const bool ok = c4 < N4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float cnt = 0.f, mean[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
    auto push = [&](const float4& v) {
        cnt += 1.f;
        const float inv = 1.f / cnt;
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float d = vv[e] - mean[e];
            mean[e] += d * inv;
            m2[e] = fmaf(d, vv[e] - mean[e], m2[e]);
        }
    };
    if (ok) {
        int r = ty;
        for (; r + BN_ROWS < B; r += 2 * BN_ROWS) {
            const float4 a = __ldg(x4 + (size_t)r * N4 + c4), b = __ldg(x4 + (size_t)(r + BN_ROWS) * N4 + c4);
            push(a); push(b);
        }
        if (r < B) push(__ldg(x4 + (size_t)r * N4 + c4));
    }
    s_cnt[ty][tx] = cnt;
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_mean[ty][tx][e] = mean[e]; s_m2[ty][tx][e] = m2[e]; }
    __syncthreads();
    if (ty < 4 && ok) {                                   // warp ty merges feature e = ty of every thread's quad (Chan's formula)
        const int e = ty;
        float n = s_cnt[0][tx], mu = s_mean[0][tx][e], M2 = s_m2[0][tx][e];
#pragma unroll
        for (int r = 1; r < BN_ROWS; ++r) {
            const float nb = s_cnt[r][tx];
            if (nb > 0.f) {
                const float d = s_mean[r][tx][e] - mu, nt = n + nb;
                mu += d * (nb / nt);
                M2 += s_m2[r][tx][e] + d * d * (n * nb / nt);
                n = nt;
            }
        }
        const float var = M2 / (float)B;
        const float sd = sqrtf(var + eps);
        const size_t col = c4 * 4 + e;
        smu[col] = mu; sstd[col] = sd;
        mmu[col] = momentum * mmu[col] + (1.0f - momentum) * mu;
        mvar[col] = momentum * mvar[col] + (1.0f - momentum) * var;
        s_mu[tx][e] = mu; s_rstd[tx][e] = 1.0f / sd;
    }
    __syncthreads();
    if (ok) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        const float mu[4] = {s_mu[tx][0], s_mu[tx][1], s_mu[tx][2], s_mu[tx][3]};
        const float rs[4] = {s_rstd[tx][0], s_rstd[tx][1], s_rstd[tx][2], s_rstd[tx][3]};
        float4* y4 = reinterpret_cast<float4*>(y);
        for (int r = ty; r < B; r += BN_ROWS) {
            const size_t i = (size_t)r * N4 + c4;
            const float4 v = __ldg(x4 + i);
            y4[i] = make_float4(fmaf(g.x, (v.x - mu[0]) * rs[0], bt.x), fmaf(g.y, (v.y - mu[1]) * rs[1], bt.y),
                                fmaf(g.z, (v.z - mu[2]) * rs[2], bt.z), fmaf(g.w, (v.w - mu[3]) * rs[3], bt.w));
        }
    }
}

__global__ void __launch_bounds__(256) bn_fwd_test_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ mmu,
                                                          const float* __restrict__ mvar, float* __restrict__ smu,
                                                          float* __restrict__ sstd, float* __restrict__ y,
                                                          int B, size_t N, float eps) {
    const size_t total = (size_t)B * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t col = i % N;
        float mu = mmu[col], sd = sqrtf(mvar[col] + eps);
        if (i < N) { smu[col] = mu; sstd[col] = sd; }
        y[i] = fmaf(gamma[col], (x[i] - mu) / sd, beta[col]);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ smu, const float* __restrict__ sstd,
                                                     const float* __restrict__ dy, float* __restrict__ dx,
                                                     float* __restrict__ ggamma, float* __restrict__ gbeta, int B, size_t N) {
    __shared__ float s_a[BN_ROWS][BN_COLS], s_b[BN_ROWS][BN_COLS];
    __shared__ float s_sdy[BN_COLS], s_sdyx[BN_COLS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t col = (size_t)blockIdx.x * BN_COLS + tx;
    const bool ok = col < N;
    float mu = 0.f, sd = 1.f, g = 0.f;
    if (ok) { mu = smu[col]; sd = sstd[col]; g = gamma[col]; }
    const float rs = 1.0f / sd;
    float sdy = 0.f, sdyx = 0.f;
    if (ok)
        for (int r = ty; r < B; r += BN_ROWS) {
            size_t i = (size_t)r * N + col;
            float d = dy[i];
            sdy += d;
            sdyx = fmaf(d, (x[i] - mu) * rs, sdyx);
        }
    s_a[ty][tx] = sdy; s_b[ty][tx] = sdyx;
    __syncthreads();
    if (ty == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < BN_ROWS; ++r) { a += s_a[r][tx]; b += s_b[r][tx]; }
        s_sdy[tx] = a; s_sdyx[tx] = b;
        if (ok) {
            ggamma[col] += a * mu;      // literal: sum_b dy * mu (regularization.py:144,155)
            gbeta[col] += a;
        }
    }
    __syncthreads();
    if (ok && dx) {
        const float m = (float)B, inv_m = 1.0f / m;
        const float s1 = g * s_sdy[tx], s2 = g * s_sdyx[tx];
        for (int r = ty; r < B; r += BN_ROWS) {
            size_t i = (size_t)r * N + col;
            float xn = (x[i] - mu) * rs;
            dx[i] = inv_m * sd * (m * dy[i] * g - s1 - xn * s2);
        }
    }
}

// ---- data-parallel BatchNorm: the batch statistics of the GLOBAL minibatch ----------------------------------
// Rank-local pieces around two small exchanges, so N ranks reproduce the single-process arithmetic:
//   stats      : per-feature (mean, M2) of the local rows (Welford + Chan, as in the one-launch kernels)
//   fwd_synced : merges the world's (mean, M2) pairs (equal counts) into the global mean / biased variance, then normalises
//   bwd_sums   : local sum_b dy and sum_b dy*x_hat (+ the literal gamma / beta gradients, which are plain batch sums)
//   bwd_synced : dX from the all-reduced sums with m = the global batch
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, float* __restrict__ mean_out,
                                                       float* __restrict__ m2_out, int B, size_t N) {
    __shared__ float s_cnt[BN_ROWS][BN_COLS], s_mean[BN_ROWS][BN_COLS], s_m2[BN_ROWS][BN_COLS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t col = (size_t)blockIdx.x * BN_COLS + tx;
    const bool ok = col < N;
    float cnt = 0.f, mean = 0.f, m2 = 0.f;
    if (ok)
        for (int r = ty; r < B; r += BN_ROWS) {
            float v = x[(size_t)r * N + col];
            cnt += 1.f;
            float d = v - mean;
            mean += d / cnt;
            m2 = fmaf(d, v - mean, m2);
        }
    s_cnt[ty][tx] = cnt; s_mean[ty][tx] = mean; s_m2[ty][tx] = m2;
    __syncthreads();
    if (ty == 0 && ok) {
        float n = s_cnt[0][tx], mu = s_mean[0][tx], M2 = s_m2[0][tx];
#pragma unroll
        for (int r = 1; r < BN_ROWS; ++r) {
            float nb = s_cnt[r][tx];
            if (nb > 0.f) {
                float d = s_mean[r][tx] - mu, nt = n + nb;
                mu += d * (nb / nt);
                M2 += s_m2[r][tx] + d * d * (n * nb / nt);
                n = nt;
            }
        }
        mean_out[col] = mu; m2_out[col] = M2;
    }
}

// stats_all: [world][2][N] (mean, M2 of every rank's B rows)
__global__ void __launch_bounds__(256) bn_fwd_synced_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ mmu,
                                                            float* __restrict__ mvar, const float* __restrict__ stats_all, int world,
                                                            float* __restrict__ smu, float* __restrict__ sstd, float* __restrict__ y,
                                                            int B, size_t N, float momentum, float eps) {
    __shared__ float s_mu[BN_COLS], s_rstd[BN_COLS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t col = (size_t)blockIdx.x * BN_COLS + tx;
    const bool ok = col < N;
    if (ty == 0 && ok) {
        float n = (float)B, mu = stats_all[col], M2 = stats_all[N + col];
        for (int r = 1; r < world; ++r) {
            const float nb = (float)B, mr = stats_all[(size_t)r * 2 * N + col], m2r = stats_all[(size_t)r * 2 * N + N + col];
            const float d = mr - mu, nt = n + nb;
            mu += d * (nb / nt);
            M2 += m2r + d * d * (n * nb / nt);
            n = nt;
        }
        const float var = M2 / n;
        const float sd = sqrtf(var + eps);
        smu[col] = mu; sstd[col] = sd;
        mmu[col] = momentum * mmu[col] + (1.0f - momentum) * mu;
        mvar[col] = momentum * mvar[col] + (1.0f - momentum) * var;
        s_mu[tx] = mu; s_rstd[tx] = 1.0f / sd;
    }
    __syncthreads();
    if (ok) {
        const float mu = s_mu[tx], rs = s_rstd[tx], g = gamma[col], bt = beta[col];
        for (int r = ty; r < B; r += BN_ROWS) {
            size_t i = (size_t)r * N + col;
            y[i] = fmaf(g, (x[i] - mu) * rs, bt);
        }
    }
}

// sums: [2][N] <- (sum_b dy, sum_b dy * x_hat) of the local rows; ggamma += sum_b dy * mu, gbeta += sum_b dy (batch SUMS:
// the gradient all-reduce adds the ranks' parts)
__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ smu,
                                                          const float* __restrict__ sstd, const float* __restrict__ dy,
                                                          float* __restrict__ sums, float* __restrict__ ggamma,
                                                          float* __restrict__ gbeta, int B, size_t N) {
    __shared__ float s_a[BN_ROWS][BN_COLS], s_b[BN_ROWS][BN_COLS];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t col = (size_t)blockIdx.x * BN_COLS + tx;
    const bool ok = col < N;
    float mu = 0.f, sd = 1.f;
    if (ok) { mu = smu[col]; sd = sstd[col]; }
    const float rs = 1.0f / sd;
    float sdy = 0.f, sdyx = 0.f;
    if (ok)
        for (int r = ty; r < B; r += BN_ROWS) {
            size_t i = (size_t)r * N + col;
            float d = dy[i];
            sdy += d;
            sdyx = fmaf(d, (x[i] - mu) * rs, sdyx);
        }
    s_a[ty][tx] = sdy; s_b[ty][tx] = sdyx;
    __syncthreads();
    if (ty == 0 && ok) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < BN_ROWS; ++r) { a += s_a[r][tx]; b += s_b[r][tx]; }
        sums[col] = a; sums[N + col] = b;
        ggamma[col] += a * mu;
        gbeta[col] += a;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_synced_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ smu, const float* __restrict__ sstd,
                                                            const float* __restrict__ dy, const float* __restrict__ sums,
                                                            float* __restrict__ dx, int B, float m, size_t N) {
    const size_t total = (size_t)B * N;
    const float inv_m = 1.0f / m;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t col = i % N;
        const float g = gamma[col], mu = smu[col], sd = sstd[col];
        const float xn = (x[i] - mu) / sd;
        dx[i] = inv_m * sd * (m * dy[i] * g - g * sums[col] - xn * (g * sums[N + col]));
    }
}

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_batchnorm_fwd(const float* x, const float* gamma, const float* beta,
                      float* moving_mu, float* moving_var, float* save_mu, float* save_std,
                      float* y, int B, size_t n_feat, float momentum, float eps, int training,
                      dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0, "batchnorm_fwd: negative batch");
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && gamma && beta && moving_mu && moving_var && save_mu && save_std && y, "batchnorm_fwd: null pointer");
    if (training) {
        const bool vec = (n_feat % 4 == 0) && n_feat >= 1024 &&
                         !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                            reinterpret_cast<uintptr_t>(beta)) & 15);
        if (vec) {
            unsigned grid = (unsigned)((n_feat / 4 + BN_COLS - 1) / BN_COLS);
            bn_fwd_train4_kernel<<<grid, 256, 0, S(stream)>>>(x, gamma, beta, moving_mu, moving_var, save_mu, save_std, y,
                                                              B, n_feat, momentum, eps);
        } else {
            unsigned grid = (unsigned)((n_feat + BN_COLS - 1) / BN_COLS);
            bn_fwd_train_kernel<<<grid, 256, 0, S(stream)>>>(x, gamma, beta, moving_mu, moving_var, save_mu, save_std, y,
                                                             B, n_feat, momentum, eps);
        }
    } else {
        bn_fwd_test_kernel<<<ew_grid((size_t)B * n_feat, 256), 256, 0, S(stream)>>>(x, gamma, beta, moving_mu, moving_var,
                                                                                     save_mu, save_std, y, B, n_feat, eps);
    }
    DNB_LAUNCH_CHECK("batchnorm_fwd");
    return DNB_OK;
}

int dnb_batchnorm_bwd(const float* x, const float* gamma, const float* save_mu, const float* save_std,
                      const float* dy, float* dx, float* ggamma, float* gbeta,
                      int B, size_t n_feat, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0, "batchnorm_bwd: negative batch");
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && gamma && save_mu && save_std && dy && ggamma && gbeta, "batchnorm_bwd: null pointer");
    unsigned grid = (unsigned)((n_feat + BN_COLS - 1) / BN_COLS);
    bn_bwd_kernel<<<grid, 256, 0, S(stream)>>>(x, gamma, save_mu, save_std, dy, dx, ggamma, gbeta, B, n_feat);
    DNB_LAUNCH_CHECK("batchnorm_bwd");
    return DNB_OK;
}

int dnb_batchnorm_stats(const float* x, float* mean, float* m2, int B, size_t n_feat, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0, "batchnorm_stats: negative batch");
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && mean && m2, "batchnorm_stats: null pointer");
    bn_stats_kernel<<<(unsigned)((n_feat + BN_COLS - 1) / BN_COLS), 256, 0, S(stream)>>>(x, mean, m2, B, n_feat);
    DNB_LAUNCH_CHECK("batchnorm_stats");
    return DNB_OK;
}

int dnb_batchnorm_fwd_synced(const float* x, const float* gamma, const float* beta, float* moving_mu, float* moving_var,
                             const float* stats_all, int world, float* save_mu, float* save_std, float* y,
                             int B, size_t n_feat, float momentum, float eps, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && world >= 1, "batchnorm_fwd_synced: bad batch / world");
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && gamma && beta && moving_mu && moving_var && stats_all && save_mu && save_std && y,
                  "batchnorm_fwd_synced: null pointer");
    bn_fwd_synced_kernel<<<(unsigned)((n_feat + BN_COLS - 1) / BN_COLS), 256, 0, S(stream)>>>(
        x, gamma, beta, moving_mu, moving_var, stats_all, world, save_mu, save_std, y, B, n_feat, momentum, eps);
    DNB_LAUNCH_CHECK("batchnorm_fwd_synced");
    return DNB_OK;
}

int dnb_batchnorm_bwd_sums(const float* x, const float* save_mu, const float* save_std, const float* dy, float* sums,
                           float* ggamma, float* gbeta, int B, size_t n_feat, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0, "batchnorm_bwd_sums: negative batch");
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && save_mu && save_std && dy && sums && ggamma && gbeta, "batchnorm_bwd_sums: null pointer");
    bn_bwd_sums_kernel<<<(unsigned)((n_feat + BN_COLS - 1) / BN_COLS), 256, 0, S(stream)>>>(x, save_mu, save_std, dy, sums,
                                                                                             ggamma, gbeta, B, n_feat);
    DNB_LAUNCH_CHECK("batchnorm_bwd_sums");
    return DNB_OK;
}

int dnb_batchnorm_bwd_synced(const float* x, const float* gamma, const float* save_mu, const float* save_std,
                             const float* dy, const float* sums, float* dx, int B, int B_global, size_t n_feat,
                             dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && B_global >= B, "batchnorm_bwd_synced: bad batch");
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && gamma && save_mu && save_std && dy && sums && dx, "batchnorm_bwd_synced: null pointer");
    bn_bwd_synced_kernel<<<ew_grid((size_t)B * n_feat, 256), 256, 0, S(stream)>>>(x, gamma, save_mu, save_std, dy, sums, dx,
                                                                                   B, (float)B_global, n_feat);
    DNB_LAUNCH_CHECK("batchnorm_bwd_synced");
    return DNB_OK;
}

}  // extern "C"
