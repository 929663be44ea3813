// Row softmax (dnpy/layers/activations.py:116-135,152-160), the losses with their deltas and the
// NaN/Inf guard of Net.compute_losses (dnpy/losses.py:17-26,60-94; dnpy/net.py:342-350), and the
// accuracy metrics (dnpy/metrics.py:44-80).  One warp per row, shuffle reductions; the per-row
// losses are reduced by a single CTA in a fixed order so the scalar is run-to-run deterministic.
#include "common.cuh"

namespace dnb {

constexpr float kEps = 1e-7f;     // dnpy's epsilon = 10e-8

__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          int B, int n, int stable, int logsm) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* xr = x + (size_t)row * n;
    float* yr = y + (size_t)row * n;
    float mx = -INFINITY;
    if (stable) {
        for (int j = lane; j < n; j += 32) mx = fmaxf(mx, xr[j]);
        mx = warp_max(mx);
    } else mx = 0.f;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += expf(xr[j] - mx);
    s = warp_sum(s);
    if (!logsm) {
        for (int j = lane; j < n; j += 32) yr[j] = expf(xr[j] - mx) / s;
    } else {
        const float ls = logf(s + kEps);                       // activations.py:160
        for (int j = lane; j < n; j += 32) yr[j] = (xr[j] - mx) - ls;
    }
}

// generic Jacobian product: dx = y * (dy - sum_k dy_k y_k)   (== dy·(diag(y) - y yᵀ))
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                          float* __restrict__ dx, int B, int n) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* yr = y + (size_t)row * n;
    const float* dr = dy + (size_t)row * n;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s = fmaf(dr[j], yr[j], s);
    s = warp_sum(s);
    for (int j = lane; j < n; j += 32) dx[(size_t)row * n + j] = yr[j] * (dr[j] - s);
}

enum { LOSS_CE = 0, LOSS_CE_SM = 1, LOSS_BCE = 2, LOSS_MSE = 3, LOSS_RMSE = 4, LOSS_MAE = 5, LOSS_NLL = 6, LOSS_HINGE = 7 };

__device__ __forceinline__ int bad_bits(float v) { return (v != v ? 4 : 0) | (isinf(v) ? 8 : 0); }

// One launch per loss: a warp per row writes the delta and the row's loss term; the LAST CTA to finish (ticket counter in the
// workspace) adds the row terms in a fixed order — run-to-run deterministic — writes [loss, flags] and leaves the counter and
// the flags zeroed for the next call (the workspace is zeroed once by its owner; no memset / finish launches per step).
template <int KIND>
__global__ void __launch_bounds__(256) loss_rows_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                        float* __restrict__ delta, float* __restrict__ rows,
                                                        int* __restrict__ flags, float* __restrict__ out, int B, int n,
                                                        float scale, int root) {
    __shared__ float red[32];
    __shared__ unsigned s_ticket;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    float acc = 0.f;
    int bits = 0;
    for (int j = lane; row < B && j < n; j += 32) {
        size_t i = (size_t)row * n + j;
        float pv = p[i], tv = t[i], d, v;
        if (KIND == LOSS_CE || KIND == LOSS_CE_SM) {
            v = tv * logf(pv + kEps);
            d = (KIND == LOSS_CE_SM) ? (pv - tv) : -1.0f * (tv * 1.0f / (pv + kEps));
        } else if (KIND == LOSS_BCE) {
            v = tv * logf(pv + kEps) + (1.0f - tv) * logf(1.0f - pv + kEps);
            d = -1.0f * (tv * 1.0f / (pv + kEps) + (1.0f - tv) * 1.0f / (1.0f - pv + kEps) * -1.0f);
        } else if (KIND == LOSS_RMSE) {                       // losses.py:34-43: sqrt of the mean; delta = e / sqrt(e^2 + eps)
            float e = pv - tv;
            v = e * e;
            d = e / sqrtf(e * e + kEps);
        } else if (KIND == LOSS_MAE) {                        // losses.py:50-58
            float e = pv - tv;
            v = fabsf(e);
            d = (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : (e != e ? e : 0.f));
        } else if (KIND == LOSS_NLL) {                        // losses.py:105-114: inputs are log-probabilities
            v = tv * pv;
            d = expf(pv) - tv;
        } else if (KIND == LOSS_HINGE) {                      // losses.py:121-133
            v = fmaxf(0.f, 1.0f - tv * pv);
            d = (tv * pv > 1.0f) ? 0.f * (-1.0f * tv) : -1.0f * tv;
        } else {
            float e = pv - tv;
            v = e * e;
            d = 2.0f * e;
        }
        delta[i] = d;
        bits |= bad_bits(d);
        acc += v;
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
    if (lane == 0 && row < B) {
        rows[row] = acc;
        if (bits) atomicOr(flags, bits);
    }
    unsigned* ticket = reinterpret_cast<unsigned*>(flags + 1);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    float sum = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) sum += __ldcg(rows + i);
    const float tot = block_sum<256>(sum, red);
    if (threadIdx.x == 0) {
        float loss = tot * scale;
        if (root) loss = sqrtf(loss);                       // RMSE
        const int f = __ldcg(flags) | (loss != loss ? 1 : 0) | (isinf(loss) ? 2 : 0);
        out[0] = loss;
        out[1] = (float)f;
        *flags = 0;
        *ticket = 0u;
    }
}

template <int KIND>
static int run_loss(const char* name, const float* p, const float* t, float* delta, float* out, int B, int n,
                    float scale, void* ws, cudaStream_t st) {
    DNB_CHECK_ARG(B > 0 && n > 0, "%s: empty batch", name);
    DNB_CHECK_ARG(p && t && delta && out && ws, "%s: null pointer", name);
    float* rows = static_cast<float*>(ws);
    int* flags = reinterpret_cast<int*>(rows + B);          // [flags, ticket]: zero on entry, left zero on exit
    loss_rows_kernel<KIND><<<(B + 7) / 8, 256, 0, st>>>(p, t, delta, rows, flags, out, B, n, scale, KIND == LOSS_RMSE ? 1 : 0);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

// metrics: count rows whose first-argmax agrees / whose thresholded class agrees
__global__ void __launch_bounds__(256) metric_cat_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                         float* __restrict__ out, int B, int n) {
    __shared__ float red[32];
    float cnt = 0.f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
        const float* pr = p + (size_t)r * n;
        const float* tr = t + (size_t)r * n;
        int ip = 0, it = 0;
        float bp = pr[0], bt = tr[0];
        for (int j = 1; j < n; ++j) {
            if (pr[j] > bp) { bp = pr[j]; ip = j; }
            if (tr[j] > bt) { bt = tr[j]; it = j; }
        }
        cnt += (ip == it) ? 1.f : 0.f;
    }
    float s = block_sum<256>(cnt, red);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(out, s);
}
__global__ void __launch_bounds__(256) metric_bin_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                         float* __restrict__ out, int n, float thr) {
    __shared__ float red[32];
    float cnt = 0.f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        cnt += ((p[r] > thr) == (t[r] > thr)) ? 1.f : 0.f;
    float s = block_sum<256>(cnt, red);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(out, s);
}

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_softmax_fwd(const float* x, float* y, int B, int n, int stable, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && n > 0, "softmax_fwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && y, "softmax_fwd: null pointer");
    softmax_fwd_kernel<<<(B + 7) / 8, 256, 0, S(stream)>>>(x, y, B, n, stable, 0);
    DNB_LAUNCH_CHECK("softmax_fwd");
    return DNB_OK;
}
int dnb_logsoftmax_fwd(const float* x, float* y, int B, int n, int stable, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && n > 0, "logsoftmax_fwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(x && y, "logsoftmax_fwd: null pointer");
    softmax_fwd_kernel<<<(B + 7) / 8, 256, 0, S(stream)>>>(x, y, B, n, stable, 1);
    DNB_LAUNCH_CHECK("logsoftmax_fwd");
    return DNB_OK;
}
int dnb_softmax_bwd(const float* y, const float* dy, float* dx, int B, int n, dnb_stream_t stream) {
    DNB_CHECK_ARG(B >= 0 && n > 0, "softmax_bwd: bad shape");
    if (B == 0) return DNB_OK;
    DNB_CHECK_ARG(y && dy && dx, "softmax_bwd: null pointer");
    softmax_bwd_kernel<<<(B + 7) / 8, 256, 0, S(stream)>>>(y, dy, dx, B, n);
    DNB_LAUNCH_CHECK("softmax_bwd");
    return DNB_OK;
}

size_t dnb_loss_ws_bytes(int B) { return ((size_t)(B > 0 ? B : 0) + 4) * sizeof(float); }

int dnb_loss_ce(const float* p, const float* t, float* delta, float* out, int B, int n,
                int softmax_output, float loss_scale, void* ws, dnb_stream_t stream) {
    // loss = -mean_b sum_k t*log(p+eps)
    if (softmax_output) return run_loss<LOSS_CE_SM>("loss_ce", p, t, delta, out, B, n, -loss_scale, ws, S(stream));
    return run_loss<LOSS_CE>("loss_ce", p, t, delta, out, B, n, -loss_scale, ws, S(stream));
}
int dnb_loss_bce(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 1, "loss_bce: the reference's float(mean(axis=0)) only accepts one output column (got %d)", n);
    return run_loss<LOSS_BCE>("loss_bce", p, t, delta, out, B, n, -loss_scale, ws, S(stream));
}
int dnb_loss_mse(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 1, "loss_mse: the reference's float(mean(axis=0)) only accepts one output column (got %d)", n);
    return run_loss<LOSS_MSE>("loss_mse", p, t, delta, out, B, n, loss_scale, ws, S(stream));
}

int dnb_loss_rmse(const float* p, const float* t, float* delta, float* out, int B, int n,
                  float loss_scale, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 1, "loss_rmse: the reference's float(sqrt(mean(axis=0))) only accepts one output column (got %d)", n);
    return run_loss<LOSS_RMSE>("loss_rmse", p, t, delta, out, B, n, loss_scale, ws, S(stream));
}
int dnb_loss_mae(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 1, "loss_mae: the reference's float(mean(axis=0)) only accepts one output column (got %d)", n);
    return run_loss<LOSS_MAE>("loss_mae", p, t, delta, out, B, n, loss_scale, ws, S(stream));
}
int dnb_loss_nll(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream) {
    // loss = -mean_b sum_k t * p   (p = log-probabilities)
    return run_loss<LOSS_NLL>("loss_nll", p, t, delta, out, B, n, -loss_scale, ws, S(stream));
}
int dnb_loss_hinge(const float* p, const float* t, float* delta, float* out, int B, int n,
                   float loss_scale, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 1, "loss_hinge: the reference's float(mean(axis=0)) only accepts one output column (got %d)", n);
    return run_loss<LOSS_HINGE>("loss_hinge", p, t, delta, out, B, n, loss_scale, ws, S(stream));
}
int dnb_metric_categorical(const float* p, const float* t, float* out, int B, int n, dnb_stream_t stream) {
    DNB_CHECK_ARG(B > 0 && n > 0 && p && t && out, "metric_categorical: bad arguments");
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), S(stream));
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "metric_categorical: %s", cudaGetErrorString(e));
    int grid = min((B + 255) / 256, kNumSMs * 4);
    metric_cat_kernel<<<grid, 256, 0, S(stream)>>>(p, t, out, B, n);
    DNB_LAUNCH_CHECK("metric_categorical");
    return DNB_OK;
}
int dnb_metric_binary(const float* p, const float* t, float* out, int B, int n, float thr, dnb_stream_t stream) {
    DNB_CHECK_ARG(B > 0 && n > 0 && p && t && out, "metric_binary: bad arguments");
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), S(stream));
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "metric_binary: %s", cudaGetErrorString(e));
    int total = B * n;
    int grid = min((total + 255) / 256, kNumSMs * 4);
    metric_bin_kernel<<<grid, 256, 0, S(stream)>>>(p, t, out, total, thr);
    DNB_LAUNCH_CHECK("metric_binary");
    return DNB_OK;
}

}  // extern "C"
