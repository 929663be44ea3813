// Bandwidth-bound elementwise kernels: activations, Dropout/Add plumbing, fills and the
// fused single-pass optimizers over the flat parameter arena.
//
// All are grid-stride, 128-bit vectorised (ld/st .v4) when every pointer is 16-byte aligned,
// with a scalar tail; grids are a multiple of the 148 SMs.  Roofline: HBM, algorithmic bytes
// are listed per entry point in DESIGN.md §4.
#include "common.cuh"
#include <vector>
#include <map>
#include <string>

namespace dnb {

static thread_local char g_err[512] = {0};
char* err_buf() { return g_err; }
int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

thread_local cudaStream_t g_cur_stream = nullptr;

int num_sms() {
    static int cached = 0;          // one process drives one GPU (one rank per device)
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
        cached = n;
    return cached > 0 ? cached : 148;
}

// ---- launch bookkeeping ---------------------------------------------------------------------
static long long g_launches = 0;
struct ProfState {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;      // names[i] labels the interval ev[i] -> ev[i+1]
};
static ProfState g_prof;

void note_launch(const char* name) {
    ++g_launches;
    if (!g_prof.on || g_prof.ev.size() >= (1u << 20)) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, g_cur_stream);
    g_prof.ev.push_back(e);
    g_prof.names.push_back(name);
}

constexpr int EW_BLOCK = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- generic 1/2/3-input maps -----------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(EW_BLOCK) map1_kernel(const float* __restrict__ a, float* __restrict__ y,
                                                        size_t n, bool vec, F f) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t n4 = vec ? n / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 v = ld4(a + 4 * i);
        v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
        st4(y + 4 * i, v);
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) y[i] = f(a[i]);
}

template <typename F>
__global__ void __launch_bounds__(EW_BLOCK) map2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        float* __restrict__ y, size_t n, bool vec, F f) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t n4 = vec ? n / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 u = ld4(a + 4 * i), v = ld4(b + 4 * i), r;
        r.x = f(u.x, v.x); r.y = f(u.y, v.y); r.z = f(u.z, v.z); r.w = f(u.w, v.w);
        st4(y + 4 * i, r);
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) y[i] = f(a[i], b[i]);
}

template <typename F>
static int launch_map1(const char* name, const float* a, float* y, size_t n, cudaStream_t st, F f) {
    if (n == 0) return DNB_OK;
    bool vec = aligned16(a) && aligned16(y);
    map1_kernel<<<ew_grid(vec ? n / 4 + 1 : n, EW_BLOCK), EW_BLOCK, 0, st>>>(a, y, n, vec, f);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}
template <typename F>
static int launch_map2(const char* name, const float* a, const float* b, float* y, size_t n, cudaStream_t st, F f) {
    if (n == 0) return DNB_OK;
    bool vec = aligned16(a) && aligned16(b) && aligned16(y);
    map2_kernel<<<ew_grid(vec ? n / 4 + 1 : n, EW_BLOCK), EW_BLOCK, 0, st>>>(a, b, y, n, vec, f);
    DNB_LAUNCH_CHECK(name);
    return DNB_OK;
}

// ---- functors ---------------------------------------------------------------------------
struct LeakyFwd { float alpha; __device__ float operator()(float x) const { return x > 0.f ? x : x * alpha; } };
struct LeakyBwd { float alpha; __device__ float operator()(float x, float d) const { return x > 0.f ? d : d * alpha; } };
struct SigmoidFwd {   // 1/(1+exp(-x+eps)+eps), activations.py:78
    __device__ float operator()(float x) const { return 1.0f / (1.0f + expf(-x + 1e-7f) + 1e-7f); }
};
struct SigmoidBwd { __device__ float operator()(float y, float d) const { return d * (y * (1.f - y)); } };
struct TanhFwd {      // (e^x - e^-x)/(e^x + e^-x) evaluated in float64 overflows to NaN beyond ~709.78
    __device__ float operator()(float x) const {
        return (fabsf(x) > 709.782712893384f) ? __int_as_float(0x7fc00000) : tanhf(x);
    }
};
struct TanhBwd { __device__ float operator()(float y, float d) const { return d * (1.f - y * y); } };
struct MulOp { __device__ float operator()(float a, float b) const { return a * b; } };
struct ScaleOp { float s; __device__ float operator()(float a) const { return a * s; } };

// ---- fills --------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_BLOCK) fill_kernel(T* __restrict__ p, size_t n, T v) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += stride) p[i] = v;
}

// ---- PRelu (activations.py:59-66): alpha broadcast over the batch, galpha column sums -------
__global__ void __launch_bounds__(EW_BLOCK) prelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                                             float* __restrict__ y, size_t total, size_t nf) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < total; i += stride) {
        float v = x[i];
        y[i] = v > 0.f ? v : v * alpha[i % nf];
    }
}
// one thread per feature column walks the batch (coalesced across the warp); writes dx, accumulates galpha
__global__ void __launch_bounds__(EW_BLOCK) prelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                                             const float* __restrict__ dy, float* __restrict__ dx,
                                                             float* __restrict__ galpha, int B, size_t nf, float inv_m) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nf) return;
    float a = alpha[j], acc = 0.f;
    for (int b = 0; b < B; ++b) {
        size_t i = (size_t)b * nf + j;
        float xv = x[i], d = dy[i];
        bool gate = xv > 0.f;
        dx[i] = gate ? d : d * a;
        acc += gate ? 0.f : d;
    }
    galpha[j] += acc * inv_m;
}

// ---- Add over n inputs (merge.py:13-16) --------------------------------------------------
struct PtrPack { const float* p[8]; };
__global__ void __launch_bounds__(EW_BLOCK) add_n_kernel(PtrPack xs, int n_in, float* __restrict__ y, size_t n, bool vec) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = vec ? n / 4 : 0;                       // 128-bit body, scalar tail
    for (size_t i = tid; i < n4; i += stride) {
        float4 acc = __ldg(reinterpret_cast<const float4*>(xs.p[0]) + i);
        for (int k = 1; k < n_in; ++k) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(xs.p[k]) + i);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        reinterpret_cast<float4*>(y)[i] = acc;
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) {
        float acc = xs.p[0][i];
        for (int k = 1; k < n_in; ++k) acc += xs.p[k][i];
        y[i] = acc;
    }
}

// ---- point-wise merges over n inputs (merge.py:13-16,29-62): out = x0 op x1 op x2 ... (left fold, the reference's order) ----
__device__ __forceinline__ float merge_op(int op, float a, float b) {
    switch (op) {
        case DNB_MERGE_SUB: return a - b;
        case DNB_MERGE_MUL: return a * b;
        case DNB_MERGE_DIV: return a / b;
        case DNB_MERGE_POW: return powf(a, b);
        case DNB_MERGE_MAX: return (a != a || b != b) ? (a + b) : fmaxf(a, b);      // np.maximum propagates NaN
        case DNB_MERGE_MIN: return (a != a || b != b) ? (a + b) : fminf(a, b);
        default: return a + b;
    }
}
__global__ void __launch_bounds__(EW_BLOCK) merge_n_kernel(PtrPack xs, int n_in, float* __restrict__ y, size_t n, int op, float post) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += stride) {
        float acc = xs.p[0][i];
        for (int k = 1; k < n_in; ++k) acc = merge_op(op, acc, xs.p[k][i]);
        y[i] = acc * post;
    }
}

// ---- optimizers (dnpy/optimizers.py) -----------------------------------------------------
// Adam, literal: no bias correction, epsilon under the sqrt (Q13).  28 B/param of HBM traffic.
__global__ void __launch_bounds__(EW_BLOCK) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ v, float* __restrict__ s, size_t n, bool vec,
                                                        float lr, float b1, float b2, float eps) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const float c1 = 1.0f - b1, c2 = 1.0f - b2;
    size_t n4 = vec ? n / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        float4 P = ld4(p + 4 * i), G = ld4(g + 4 * i), V = ld4(v + 4 * i), Sv = ld4(s + 4 * i);
#define ADAM1(c) V.c = b1 * V.c + c1 * G.c; Sv.c = b2 * Sv.c + c2 * G.c * G.c; P.c -= lr * (V.c / sqrtf(Sv.c + eps));
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        st4(p + 4 * i, P); st4(v + 4 * i, V); st4(s + 4 * i, Sv);
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) {
        float G = g[i];
        float V = b1 * v[i] + c1 * G, Sv = b2 * s[i] + c2 * G * G;
        v[i] = V; s[i] = Sv;
        p[i] -= lr * (V / sqrtf(Sv + eps));
    }
}
__global__ void __launch_bounds__(EW_BLOCK) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ s, size_t n, float lr, float rho, float eps) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += stride) {
        float G = g[i];
        float Sv = rho * s[i] + (1.0f - rho) * G * G;
        s[i] = Sv;
        p[i] -= lr * (G / sqrtf(Sv + eps));
    }
}
__global__ void __launch_bounds__(EW_BLOCK) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ v, size_t n, float lr, float mom, int nesterov) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += stride) {
        float vp = v[i];
        float V = mom * vp + (1.0f - mom) * g[i];       // dampened EMA (optimizers.py:39)
        v[i] = V;
        float nv = nesterov ? (-mom * vp + (1.0f + mom) * V) : V;
        p[i] -= lr * nv;
    }
}

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_version(void) { return 100; }
const char* dnb_last_error(void) { return err_buf(); }

long long dnb_launch_count(void) { return g_launches; }

int dnb_prof_begin(dnb_stream_t stream) {
    for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
    g_prof.ev.clear();
    g_prof.names.clear();
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreate(&e);
    if (rc != cudaSuccess) return set_err(DNB_ERR_CUDA, "prof_begin: %s", cudaGetErrorString(rc));
    cudaEventRecord(e, S(stream));
    g_prof.ev.push_back(e);
    g_prof.on = true;
    return DNB_OK;
}

int dnb_prof_end(char* json_out, size_t cap) {
    g_prof.on = false;
    if (g_prof.ev.empty()) return set_err(DNB_ERR_INVALID, "prof_end without prof_begin");
    cudaError_t rc = cudaEventSynchronize(g_prof.ev.back());
    if (rc != cudaSuccess) return set_err(DNB_ERR_CUDA, "prof_end: %s", cudaGetErrorString(rc));
    std::map<std::string, std::pair<long long, double>> agg;
    for (size_t i = 0; i + 1 < g_prof.ev.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]);
        auto& a = agg[g_prof.names[i]];
        a.first += 1;
        a.second += ms;
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s\"%s\": [%lld, %.6f]", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        out += buf;
        first = false;
    }
    out += "}";
    for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
    g_prof.ev.clear();
    g_prof.names.clear();
    if (!json_out || out.size() + 1 > cap) return set_err(DNB_ERR_INVALID, "prof_end: buffer too small (%zu needed)", out.size() + 1);
    memcpy(json_out, out.c_str(), out.size() + 1);
    return DNB_OK;
}

int dnb_device_query(int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (has_tcgen05) *has_tcgen05 = (p.major == 10);
    return DNB_OK;
}

int dnb_fill_f32(float* p, size_t n, float value, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(p, "fill_f32: null pointer");
    if (value == 0.0f) {
        cudaError_t e = cudaMemsetAsync(p, 0, n * sizeof(float), S(stream));
        if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "fill_f32: %s", cudaGetErrorString(e));
        return DNB_OK;
    }
    fill_kernel<float><<<ew_grid(n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(p, n, value);
    DNB_LAUNCH_CHECK("fill_f32");
    return DNB_OK;
}
int dnb_fill_i32(int32_t* p, size_t n, int32_t value, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(p, "fill_i32: null pointer");
    fill_kernel<int32_t><<<ew_grid(n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(p, n, value);
    DNB_LAUNCH_CHECK("fill_i32");
    return DNB_OK;
}

int dnb_leakyrelu_fwd(const float* x, float* y, size_t n, float alpha, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (x && y), "leakyrelu_fwd: null pointer");
    return launch_map1("leakyrelu_fwd", x, y, n, S(stream), LeakyFwd{alpha});
}
int dnb_leakyrelu_bwd(const float* x, const float* dy, float* dx, size_t n, float alpha, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (x && dy && dx), "leakyrelu_bwd: null pointer");
    return launch_map2("leakyrelu_bwd", x, dy, dx, n, S(stream), LeakyBwd{alpha});
}
int dnb_sigmoid_fwd(const float* x, float* y, size_t n, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (x && y), "sigmoid_fwd: null pointer");
    return launch_map1("sigmoid_fwd", x, y, n, S(stream), SigmoidFwd{});
}
int dnb_sigmoid_bwd(const float* y, const float* dy, float* dx, size_t n, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (y && dy && dx), "sigmoid_bwd: null pointer");
    return launch_map2("sigmoid_bwd", y, dy, dx, n, S(stream), SigmoidBwd{});
}
int dnb_tanh_fwd(const float* x, float* y, size_t n, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (x && y), "tanh_fwd: null pointer");
    return launch_map1("tanh_fwd", x, y, n, S(stream), TanhFwd{});
}
int dnb_tanh_bwd(const float* y, const float* dy, float* dx, size_t n, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (y && dy && dx), "tanh_bwd: null pointer");
    return launch_map2("tanh_bwd", y, dy, dx, n, S(stream), TanhBwd{});
}
int dnb_mul(const float* a, const float* b, float* y, size_t n, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (a && b && y), "mul: null pointer");
    return launch_map2("mul", a, b, y, n, S(stream), MulOp{});
}
int dnb_scale(const float* x, float* y, size_t n, float s, dnb_stream_t stream) {
    DNB_CHECK_ARG(n == 0 || (x && y), "scale: null pointer");
    return launch_map1("scale", x, y, n, S(stream), ScaleOp{s});
}

int dnb_prelu_fwd(const float* x, const float* alpha, float* y, int B, size_t n_feat, dnb_stream_t stream) {
    size_t total = (size_t)B * n_feat;
    if (total == 0) return DNB_OK;
    DNB_CHECK_ARG(x && alpha && y, "prelu_fwd: null pointer");
    prelu_fwd_kernel<<<ew_grid(total, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(x, alpha, y, total, n_feat);
    DNB_LAUNCH_CHECK("prelu_fwd");
    return DNB_OK;
}
int dnb_prelu_bwd(const float* x, const float* alpha, const float* dy, float* dx, float* galpha,
                  int B, size_t n_feat, float inv_m, dnb_stream_t stream) {
    if ((size_t)B * n_feat == 0) return DNB_OK;
    DNB_CHECK_ARG(x && alpha && dy && dx && galpha, "prelu_bwd: null pointer");
    int grid = (int)((n_feat + EW_BLOCK - 1) / EW_BLOCK);
    prelu_bwd_kernel<<<grid, EW_BLOCK, 0, S(stream)>>>(x, alpha, dy, dx, galpha, B, n_feat, inv_m);
    DNB_LAUNCH_CHECK("prelu_bwd");
    return DNB_OK;
}

int dnb_add_n(const float* const* xs, int n_in, float* y, size_t n, dnb_stream_t stream) {
    DNB_CHECK_ARG(n_in >= 1 && n_in <= 8, "add_n: between 1 and 8 inputs supported, got %d", n_in);
    if (n == 0) return DNB_OK;
    PtrPack pk;
    for (int i = 0; i < 8; ++i) pk.p[i] = (i < n_in) ? xs[i] : nullptr;
    for (int i = 0; i < n_in; ++i) DNB_CHECK_ARG(xs[i], "add_n: null input %d", i);
    bool vec = aligned16(y);
    for (int i = 0; i < n_in; ++i) vec = vec && aligned16(xs[i]);
    add_n_kernel<<<ew_grid(vec ? n / 4 + 1 : n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(pk, n_in, y, n, vec);
    DNB_LAUNCH_CHECK("add_n");
    return DNB_OK;
}

int dnb_merge_n(const float* const* xs, int n_in, float* y, size_t n, int op, float post_scale, dnb_stream_t stream) {
    DNB_CHECK_ARG(n_in >= 1 && n_in <= 8, "merge_n: between 1 and 8 inputs supported, got %d", n_in);
    DNB_CHECK_ARG(op >= DNB_MERGE_ADD && op <= DNB_MERGE_MIN, "merge_n: unknown operator %d", op);
    if (n == 0) return DNB_OK;
    PtrPack pk;
    for (int i = 0; i < 8; ++i) pk.p[i] = (i < n_in) ? xs[i] : nullptr;
    for (int i = 0; i < n_in; ++i) DNB_CHECK_ARG(xs[i], "merge_n: null input %d", i);
    merge_n_kernel<<<ew_grid(n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(pk, n_in, y, n, op, post_scale);
    DNB_LAUNCH_CHECK("merge_n");
    return DNB_OK;
}

int dnb_opt_adam(float* p, const float* g, float* v, float* s, size_t n, float lr,
                 float beta1, float beta2, float eps, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(p && g && v && s, "opt_adam: null pointer");
    bool vec = aligned16(p) && aligned16(g) && aligned16(v) && aligned16(s);
    adam_kernel<<<ew_grid(vec ? n / 4 + 1 : n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(p, g, v, s, n, vec, lr, beta1, beta2, eps);
    DNB_LAUNCH_CHECK("opt_adam");
    return DNB_OK;
}
int dnb_opt_rmsprop(float* p, const float* g, float* s, size_t n, float lr, float rho, float eps, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(p && g && s, "opt_rmsprop: null pointer");
    rmsprop_kernel<<<ew_grid(n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(p, g, s, n, lr, rho, eps);
    DNB_LAUNCH_CHECK("opt_rmsprop");
    return DNB_OK;
}
int dnb_opt_sgd(float* p, const float* g, float* v, size_t n, float lr, float momentum, int nesterov, dnb_stream_t stream) {
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(p && g && v, "opt_sgd: null pointer");
    sgd_kernel<<<ew_grid(n, EW_BLOCK), EW_BLOCK, 0, S(stream)>>>(p, g, v, n, lr, momentum, nesterov);
    DNB_LAUNCH_CHECK("opt_sgd");
    return DNB_OK;
}

}  // extern "C"
