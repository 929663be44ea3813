// MaxPool2D / AvgPool2D (dnpy/layers/pooling.py:43-99, 118-174), literal semantics:
//   * the input is ZERO padded and the pads take part in max / mean (Q6);
//   * argmax = first maximum of the row-major flattened window (np.argmax), stored as the
//     window offset in [0, py*px) — bit-exact integer work;
//   * MaxPool2D.backward LITERAL route (Q5): the flat index addresses the (B,py,px) window view,
//     so every sample's delta lands in sample 0's window and the LAST sample that selected an
//     offset wins; PER_SAMPLE is the intended scatter-add (identical when B == 1);
//   * AvgPool2D.backward ASSIGNS delta/(py*px) to the window, later windows overwrite.
// All kernels are gather-style (no atomics on the data path), coalesced along x.  Roofline: HBM.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace dnb {

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          int32_t* __restrict__ idx, dnb_win_t g) {
    const long n = (long)g.B * g.C * g.OH * g.OW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ox = (int)(i % g.OW); long t = i / g.OW; int oy = (int)(t % g.OH); long bc = t / g.OH;
        const float* xp = x + bc * g.H * g.W;
        float best = 0.f; int bi = 0; bool first = true;
        for (int ii = 0; ii < g.kh; ++ii) {
            int iy = oy * g.sy + ii - g.pt;
            bool rowin = (iy >= 0 && iy < g.H);
            for (int jj = 0; jj < g.kw; ++jj) {
                int ix = ox * g.sx + jj - g.pl;
                float v = (rowin && ix >= 0 && ix < g.W) ? __ldg(xp + (long)iy * g.W + ix) : 0.f;   // zero pad
                // np.argmax semantics: first maximum wins; NaN propagates as the maximum
                if (first || v > best || (v != v && best == best)) { best = v; bi = ii * g.kw + jj; first = false; }
            }
        }
        y[i] = best;
        idx[i] = bi;
    }
}

// Fast path for unpadded pools with compile-time window/stride and an even input row pitch: a thread
// produces FOUR adjacent outputs of one row, reading each input row once with 8-byte loads
// ((3*S + K) columns instead of 4*K scalar loads) and 32-bit index arithmetic.
template <int K, int S>
__global__ void __launch_bounds__(256) maxpool_fwd_strip_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                int32_t* __restrict__ idx, int planes, int H, int W, int OH, int OW) {
    constexpr int NCOL = 3 * S + K;                    // input columns feeding 4 outputs
    constexpr int NV = (NCOL + 1) / 2;                 // float2 loads per row
    const int spr = (OW + 3) >> 2;
    const long total = (long)planes * OH * spr;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int sx = (int)(t % spr);
        const long r = t / spr;
        const int oy = (int)(r % OH);
        const long pl = r / OH;
        const int ox0 = sx * 4;
        const float* xp = x + pl * (long)H * W + (long)(oy * S) * W + ox0 * S;
        float best[4];
        int bi[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) { best[p] = 0.f; bi[p] = -1; }
        const int ncol_ok = W - ox0 * S;               // columns available in this row from ox0*S on
#pragma unroll
        for (int i = 0; i < K; ++i) {
            float v[2 * NV];
            const float2* rp = reinterpret_cast<const float2*>(xp + (long)i * W);
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                if (2 * q + 1 < ncol_ok) { float2 f = __ldg(rp + q); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
                else { v[2 * q] = (2 * q < ncol_ok) ? __ldg(xp + (long)i * W + 2 * q) : 0.f; v[2 * q + 1] = 0.f; }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const float c = v[p * S + j];
                    if (bi[p] < 0 || c > best[p] || (c != c && best[p] == best[p])) { best[p] = c; bi[p] = i * K + j; }
                }
        }
        const long o = (pl * OH + oy) * (long)OW + ox0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (ox0 + p < OW) { y[o + p] = best[p]; idx[o + p] = bi[p]; }
    }
}

// Ring path for unpadded pools whose planes are whole 16-byte quads: ALL global traffic is bulk-asynchronous (TMA 1-D
// copies).  The (b,c) planes are contiguous, so a chunk of G planes is one cp.async.bulk into a 4-stage shared ring
// (a producer lane keeps it full); 256 consumer threads take the windows out of shared memory (geometry per thread is
// chunk independent and lives in registers), put values and argmax offsets into a double-buffered shared output
// chunk, and one thread writes each finished chunk back with two bulk stores.  No L1/LSU pressure from the
// overlapping, row-strided window reads (the strip kernel above is l1tex bound at ~54% of HBM peak).
constexpr int MPR_CONS = 256, MPR_THREADS = MPR_CONS + 32, MPR_ST = 4, MPR_MAXR = 6;

__device__ __forceinline__ void mp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mp_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(tc::smem_u32(src)), "r"(bytes) : "memory");
}

template <int K, int S, bool VEC2>
__global__ void __launch_bounds__(MPR_THREADS, 2) maxpool_fwd_ring_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                           int32_t* __restrict__ idx, int nchunks, int G,
                                                                           int H, int W, int OH, int OW) {
    extern __shared__ __align__(128) uint8_t smraw[];
    const int plane = H * W, oplane = OH * OW;
    const int n_in = G * plane, n_out = G * oplane;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + MPR_ST;
    float* sin = reinterpret_cast<float*>(smraw + 128);                 // [ST][n_in]
    float* sy = sin + (size_t)MPR_ST * n_in;                            // [2][n_out]
    int32_t* si = reinterpret_cast<int32_t*>(sy + 2 * n_out);           // [2][n_out]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < MPR_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], MPR_CONS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int mine = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // chunks blockIdx.x + n*gridDim.x

    if (wid == MPR_CONS / 32) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)n_in * 4u;
            for (int n = 0; n < mine; ++n) {
                const int s = n % MPR_ST;
                if (n >= MPR_ST) tc::mbar_wait(&empty[s], ((n / MPR_ST) - 1) & 1);
                const long c = (long)blockIdx.x + (long)n * gridDim.x;
                tc::mbar_arrive_expect_tx(&full[s], bytes);
                mp_bulk_g2s(sin + (size_t)s * n_in, x + c * n_in, bytes, &full[s]);
            }
        }
        return;
    }
    // chunk-independent geometry: output e = tid + r*256 of the chunk reads its window at in_off[r]
    int in_off[MPR_MAXR];
#pragma unroll
    for (int r = 0; r < MPR_MAXR; ++r) {
        const int e = min(tid + r * MPR_CONS, n_out - 1);
        const int p = e / oplane, o = e - p * oplane;
        const int oy = o / OW, ox = o - oy * OW;
        in_off[r] = p * plane + oy * S * W + ox * S;
    }
    for (int n = 0; n < mine; ++n) {
        const int s = n % MPR_ST, ob = n & 1;
        tc::mbar_wait(&full[s], (n / MPR_ST) & 1);
        const float* in = sin + (size_t)s * n_in;
        float* oy_ = sy + ob * n_out;
        int32_t* oi_ = si + ob * n_out;
#pragma unroll
        for (int r = 0; r < MPR_MAXR; ++r) {
            const int e = tid + r * MPR_CONS;
            if (e < n_out) {
                const float* wp = in + in_off[r];
                float best = 0.f; int bi = -1;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    float v[K];
                    if (VEC2) {          // S == 2, W even, plane even: the window row starts on an 8-byte boundary
#pragma unroll
                        for (int j = 0; j + 1 < K; j += 2) {
                            const float2 f = *reinterpret_cast<const float2*>(wp + i * W + j);
                            v[j] = f.x; v[j + 1] = f.y;
                        }
                        if (K & 1) v[K - 1] = wp[i * W + K - 1];
                    } else {
#pragma unroll
                        for (int j = 0; j < K; ++j) v[j] = wp[i * W + j];
                    }
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const float c = v[j];
                        // np.argmax: first maximum wins; NaN propagates as the maximum
                        if (bi < 0 || c > best || (c != c && best == best)) { best = c; bi = i * K + j; }
                    }
                }
                oy_[e] = best;
                oi_[e] = bi;
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);
        // hand the finished output chunk to the bulk-store engine; the OTHER output buffer must be drained before the
        // next chunk overwrites it (its store was committed one chunk ago)
        tc::fence_proxy_async_smem();
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(MPR_CONS) : "memory");
        if (tid == 0) {
            const long c = (long)blockIdx.x + (long)n * gridDim.x;
            mp_bulk_s2g(y + c * n_out, oy_, (uint32_t)n_out * 4u);
            mp_bulk_s2g(idx + c * n_out, oi_, (uint32_t)n_out * 4u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// planes per chunk for the ring kernel (0: shape not eligible)
static int mpr_chunk_planes(long planes, int H, int W, int OH, int OW, long stage_budget = 24 * 1024, int omult = 4) {
    const long plane = (long)H * W, oplane = (long)OH * OW;
    if (plane % 4) return 0;
    const long gmax = std::min<long>({planes, stage_budget / (plane * 4), (long)(MPR_MAXR * MPR_CONS) / oplane});
    for (long G = gmax; G >= 1; --G)
        if (planes % G == 0 && (G * oplane) % omult == 0) return (int)G;
    return 0;
}
template <int K, int S>
static bool maxpool_fwd_ring(const float* x, float* y, int32_t* idx, const dnb_win_t* g, cudaStream_t st) {
    const long planes = (long)g->B * g->C;
    const int G = mpr_chunk_planes(planes, g->H, g->W, g->OH, g->OW);
    if (!G || getenv("DNB_NO_POOL_RING")) return false;
    const long nchunks = planes / G;
    if (nchunks > 0x7fffffffL) return false;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(idx)) & 15) return false;
    const size_t smem = 128 + ((size_t)MPR_ST * G * g->H * g->W + 4ull * G * g->OH * g->OW) * 4;
    const int grid = (int)std::min<long>(nchunks, 2L * kNumSMs);
    const bool v2 = (S == 2) && (g->W % 2 == 0);
    if (v2) {
        cudaFuncSetAttribute(maxpool_fwd_ring_kernel<K, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        maxpool_fwd_ring_kernel<K, S, true><<<grid, MPR_THREADS, smem, st>>>(x, y, idx, (int)nchunks, G, g->H, g->W, g->OH, g->OW);
    } else {
        cudaFuncSetAttribute(maxpool_fwd_ring_kernel<K, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        maxpool_fwd_ring_kernel<K, S, false><<<grid, MPR_THREADS, smem, st>>>(x, y, idx, (int)nchunks, G, g->H, g->W, g->OH, g->OW);
    }
    return true;
}

// PER_SAMPLE backward, ring form (unpadded pools, same eligibility as the forward ring): dY and the argmax offsets of a
// chunk of G planes arrive by two bulk copies, the dX chunk is built in shared memory and leaves by one bulk store.
// Windows of the same class (oy mod CL, ox mod CL), CL = ceil(K/S), never overlap: each class is a pass of plain
// read-modify-writes (one thread per window: 3 LDS + 1 STS), passes separated by a barrier — deterministic, no atomics,
// ~9x fewer shared-memory operations than gathering the covering windows of every input pixel (the gather kernels are
// l1tex bound at 24% of HBM peak on B200).
template <int K, int S>
__global__ void __launch_bounds__(MPR_THREADS, 2) maxpool_bwd_ring_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                           float* __restrict__ dx, int nchunks, int G,
                                                                           int H, int W, int OH, int OW) {
    constexpr int CL = (K + S - 1) / S;
    extern __shared__ __align__(128) uint8_t smraw[];
    const int plane = H * W, oplane = OH * OW;
    const int n_in = G * plane, n_out = G * oplane;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + MPR_ST;
    float* sin = reinterpret_cast<float*>(smraw + 128);                 // [ST][2 * n_out]: dY chunk, then argmax chunk
    float* sdx = sin + (size_t)MPR_ST * 2 * n_out;                      // [2][n_in]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < MPR_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], MPR_CONS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int mine = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // chunks blockIdx.x + n*gridDim.x

    if (wid == MPR_CONS / 32) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)n_out * 4u;
            for (int n = 0; n < mine; ++n) {
                const int s = n % MPR_ST;
                if (n >= MPR_ST) tc::mbar_wait(&empty[s], ((n / MPR_ST) - 1) & 1);
                const long c = (long)blockIdx.x + (long)n * gridDim.x;
                tc::mbar_arrive_expect_tx(&full[s], 2 * bytes);
                mp_bulk_g2s(sin + (size_t)s * 2 * n_out, dy + c * n_out, bytes, &full[s]);
                mp_bulk_g2s(sin + (size_t)s * 2 * n_out + n_out, idx + c * n_out, bytes, &full[s]);
            }
        }
        return;
    }
    // chunk-independent geometry: window e = tid + r*256 of the chunk starts at base[r] in the dX chunk and is of class cls[r]
    int base[MPR_MAXR], cls[MPR_MAXR];
#pragma unroll
    for (int r = 0; r < MPR_MAXR; ++r) {
        const int e = tid + r * MPR_CONS;
        const int ec = min(e, n_out - 1);
        const int p = ec / oplane, o = ec - p * oplane;
        const int oy = o / OW, ox = o - oy * OW;
        base[r] = p * plane + oy * S * W + ox * S;
        cls[r] = e < n_out ? (oy % CL) * CL + (ox % CL) : -1;
    }
    const int n_in4 = n_in >> 2;                                        // n_in is a multiple of 4 (eligibility)
    for (int n = 0; n < mine; ++n) {
        const int s = n % MPR_ST, ob = n & 1;
        float* out = sdx + (size_t)ob * n_in;
        // buffer ob was handed to the store engine two chunks ago; that store was drained before chunk n-1 was committed
        for (int e = tid; e < n_in4; e += MPR_CONS) reinterpret_cast<float4*>(out)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        tc::mbar_wait(&full[s], (n / MPR_ST) & 1);
        const float* sd = sin + (size_t)s * 2 * n_out;
        const int32_t* si = reinterpret_cast<const int32_t*>(sd + n_out);
        float d[MPR_MAXR]; int tgt[MPR_MAXR];
#pragma unroll
        for (int r = 0; r < MPR_MAXR; ++r) {
            if (cls[r] >= 0) {
                const int e = tid + r * MPR_CONS;
                const int off = si[e];
                d[r] = sd[e];
                tgt[r] = base[r] + (off / K) * W + (off % K);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);                      // the stage has been read into registers
        asm volatile("bar.sync 1, %0;" ::"n"(MPR_CONS) : "memory");     // zero fill complete
#pragma unroll
        for (int c = 0; c < CL * CL; ++c) {
#pragma unroll
            for (int r = 0; r < MPR_MAXR; ++r)
                if (cls[r] == c) out[tgt[r]] += d[r];
            if (c + 1 < CL * CL) asm volatile("bar.sync 1, %0;" ::"n"(MPR_CONS) : "memory");
        }
        tc::fence_proxy_async_smem();
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(MPR_CONS) : "memory");
        if (tid == 0) {
            const long c = (long)blockIdx.x + (long)n * gridDim.x;
            mp_bulk_s2g(dx + c * n_in, out, (uint32_t)n_in * 4u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int K, int S>
static bool maxpool_bwd_ring(const float* dy, const int32_t* idx, float* dx, const dnb_win_t* g, cudaStream_t st) {
    const long planes = (long)g->B * g->C;
    const int G = mpr_chunk_planes(planes, g->H, g->W, g->OH, g->OW);
    if (!G || getenv("DNB_NO_POOL_RING")) return false;
    const long nchunks = planes / G;
    if (nchunks > 0x7fffffffL) return false;
    if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(idx)) & 15) return false;
    const size_t smem = 128 + ((size_t)MPR_ST * 2 * G * g->OH * g->OW + 2ull * G * g->H * g->W) * 4;
    const int grid = (int)std::min<long>(nchunks, 2L * kNumSMs);
    cudaFuncSetAttribute(maxpool_bwd_ring_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    maxpool_bwd_ring_kernel<K, S><<<grid, MPR_THREADS, smem, st>>>(dy, idx, dx, (int)nchunks, G, g->H, g->W, g->OH, g->OW);
    return true;
}

// PER_SAMPLE: dx[b,c,iy,ix] = sum over windows covering it whose argmax points here.
__global__ void __launch_bounds__(256) maxpool_bwd_per_sample_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                     float* __restrict__ dx, dnb_win_t g) {
    const long n = (long)g.B * g.C * g.H * g.W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ix = (int)(i % g.W); long t = i / g.W; int iy = (int)(t % g.H); long bc = t / g.H;
        const float* dp = dy + bc * g.OH * g.OW;
        const int32_t* ip = idx + bc * g.OH * g.OW;
        const int py = iy + g.pt, px = ix + g.pl;          // padded coordinates
        float acc = 0.f;
        // windows (oy,ox) with oy*sy <= py < oy*sy+kh, in the reference's (y,x) accumulation order
        int oy_lo = max(0, (py - g.kh + g.sy) / g.sy), oy_hi = min(g.OH - 1, py / g.sy);
        int ox_lo = max(0, (px - g.kw + g.sx) / g.sx), ox_hi = min(g.OW - 1, px / g.sx);
        for (int oy = oy_lo; oy <= oy_hi; ++oy)
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                int off = (py - oy * g.sy) * g.kw + (px - ox * g.sx);
                if (ip[oy * g.OW + ox] == off) acc += dp[oy * g.OW + ox];
            }
        dx[i] = acc;
    }
}

// PER_SAMPLE fast path (no padding, stride 2, K in {2,3}): a thread owns 4 adjacent input pixels of one row and
// visits the <= 2 x 3 windows that can reach them, in the reference's (oy, ox) accumulation order.
template <int K>
__global__ void __launch_bounds__(256) maxpool_bwd_ps_strip_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                   float* __restrict__ dx, int planes, int H, int W, int OH, int OW) {
    const int spr = (W + 3) >> 2;
    const long total = (long)planes * H * spr;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int sx = (int)(t % spr);
        const long r = t / spr;
        const int iy = (int)(r % H);
        const long pl = r / H;
        const int ix0 = sx * 4;
        const int oy_lo = iy >= K - 1 ? (iy - K + 2) >> 1 : 0, oy_hi = min(OH - 1, iy >> 1);
        const int ox_lo = ix0 >= K - 1 ? (ix0 - K + 2) >> 1 : 0, ox_hi = min(OW - 1, (ix0 + 3) >> 1);
        const float* dp = dy + pl * (long)OH * OW;
        const int32_t* ip = idx + pl * (long)OH * OW;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            const int wi = iy - 2 * oy;                    // window row that lands on iy
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                const int off = __ldg(ip + oy * OW + ox);
                const int oi = off / K, oj = off - oi * K;
                const int col = 2 * ox + oj - ix0;
                if (oi == wi && col >= 0 && col < 4) {
                    const float d = __ldg(dp + oy * OW + ox);
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (q == col) acc[q] += d;
                }
            }
        }
        float* xp = dx + (pl * H + iy) * (long)W + ix0;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (ix0 + q < W) xp[q] = acc[q];
    }
}

// LITERAL step 1: last[c,oy,ox,p] = max b with idx[b,c,oy,ox] == p  (int32, pre-filled with -1)
// step 1b: samples [0, b_end) in chunks, only for positions step 1a could not finish (windows with an offset that almost
// never wins, e.g. the zero-padded border column of a "same" convolution's output).  16 loads in flight per thread: a
// dependent chain of single loads made this kernel pure DRAM latency (40 us for 48 MB on B200).
__global__ void __launch_bounds__(256) maxpool_last_kernel(const int32_t* __restrict__ idx, int32_t* __restrict__ last,
                                                           const int32_t* __restrict__ done, dnb_win_t g, int bchunk, int b_end) {
    const long npos = (long)g.C * g.OH * g.OW;
    const long pos = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos || done[pos]) return;
    const int P = g.kh * g.kw;
    const int b_hi = min(b_end, (int)(blockIdx.y + 1) * bchunk) - 1, b_lo = blockIdx.y * bchunk;
    // walk this chunk from its last sample down; the first hit per offset is the chunk's last
    unsigned long long seen_lo = 0ull;      // offsets 0..63 tracked in a bitmask, the rest go straight to atomicMax
    for (int b0 = b_hi; b0 >= b_lo; b0 -= 16) {
        int pv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) pv[u] = (b0 - u >= b_lo) ? __ldg(idx + (long)(b0 - u) * npos + pos) : -1;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int p = pv[u];
            if (p < 0) continue;
            if (p < 64) {
                if (seen_lo >> p & 1ull) continue;
                seen_lo |= 1ull << p;
            }
            atomicMax(last + pos * P + p, b0 - u);
        }
    }
}
// LITERAL step 1a: the newest TOP samples decide almost every (window, offset) pair.  A CTA takes 32 consecutive window
// positions (coalesced along the position axis) and its LT_WARPS warps split the TOP newest samples between them: ONE batch
// of TOP / LT_WARPS loads per thread instead of a serial walk, merged with shared-memory atomicMax.  done[pos] tells step 1b
// (maxpool_last_kernel over the older samples) to skip a position whose kh*kw offsets have all been seen.
constexpr int LT_WARPS = 16, LT_TOP = 128, LT_PER = LT_TOP / LT_WARPS;
__global__ void __launch_bounds__(LT_WARPS * 32) maxpool_last_top_kernel(const int32_t* __restrict__ idx, int32_t* __restrict__ last,
                                                                         int32_t* __restrict__ done, dnb_win_t g) {
    extern __shared__ int32_t slast[];                     // [32 positions][P]
    const long npos = (long)g.C * g.OH * g.OW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long pos = (long)blockIdx.x * 32 + lane;
    const int P = g.kh * g.kw;
    for (int e = threadIdx.x; e < 32 * P; e += LT_WARPS * 32) slast[e] = -1;
    __syncthreads();
    const int b_lo = max(0, g.B - LT_TOP);
    if (pos < npos) {
        int pv[LT_PER];
#pragma unroll
        for (int u = 0; u < LT_PER; ++u) {
            const int b = g.B - 1 - warp - u * LT_WARPS;
            pv[u] = b >= b_lo ? __ldg(idx + (long)b * npos + pos) : -1;
        }
#pragma unroll
        for (int u = 0; u < LT_PER; ++u)
            if (pv[u] >= 0) atomicMax(&slast[lane * P + pv[u]], g.B - 1 - warp - u * LT_WARPS);
    }
    __syncthreads();
    const long pos0 = (long)blockIdx.x * 32;
    for (int e = threadIdx.x; e < 32 * P; e += LT_WARPS * 32)
        if (pos0 + e / P < npos && slast[e] >= 0) last[pos0 * P + e] = slast[e];
    if (threadIdx.x < 32 && pos < npos) {
        bool all = true;
        for (int p = 0; p < P; ++p) all = all && slast[lane * P + p] >= 0;
        done[pos] = all || b_lo == 0;
    }
}
// LITERAL step 2: dx[0,c,iy,ix] = sum over covering windows of dy[last, c, oy, ox]; dx[b>0] = 0
__global__ void __launch_bounds__(256) maxpool_bwd_literal_kernel(const float* __restrict__ dy, const int32_t* __restrict__ last,
                                                                  float* __restrict__ dx, dnb_win_t g) {
    const long n0 = (long)g.C * g.H * g.W;
    const long npos = (long)g.C * g.OH * g.OW;
    const int P = g.kh * g.kw;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += (long)gridDim.x * blockDim.x) {
        int ix = (int)(i % g.W); long t = i / g.W; int iy = (int)(t % g.H); int c = (int)(t / g.H);
        const int py = iy + g.pt, px = ix + g.pl;
        float acc = 0.f;
        int oy_lo = max(0, (py - g.kh + g.sy) / g.sy), oy_hi = min(g.OH - 1, py / g.sy);
        int ox_lo = max(0, (px - g.kw + g.sx) / g.sx), ox_hi = min(g.OW - 1, px / g.sx);
        for (int oy = oy_lo; oy <= oy_hi; ++oy)
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                int off = (py - oy * g.sy) * g.kw + (px - ox * g.sx);
                long pos = ((long)c * g.OH + oy) * g.OW + ox;
                int b = last[pos * P + off];
                if (b >= 0) acc += dy[(long)b * npos + pos];
            }
        dx[i] = acc;
    }
}

__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, dnb_win_t g) {
    const long n = (long)g.B * g.C * g.OH * g.OW;
    const float inv = 1.0f / (float)(g.kh * g.kw);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ox = (int)(i % g.OW); long t = i / g.OW; int oy = (int)(t % g.OH); long bc = t / g.OH;
        const float* xp = x + bc * g.H * g.W;
        float acc = 0.f;
        for (int ii = 0; ii < g.kh; ++ii) {
            int iy = oy * g.sy + ii - g.pt;
            if (iy < 0 || iy >= g.H) continue;
            for (int jj = 0; jj < g.kw; ++jj) {
                int ix = ox * g.sx + jj - g.pl;
                if (ix >= 0 && ix < g.W) acc += __ldg(xp + (long)iy * g.W + ix);
            }
        }
        y[i] = acc * inv;
    }
}

// assignment semantics: the LAST window in (oy,ox) loop order that covers the pixel wins
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, dnb_win_t g) {
    const long n = (long)g.B * g.C * g.H * g.W;
    const float inv = 1.0f / (float)(g.kh * g.kw);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ix = (int)(i % g.W); long t = i / g.W; int iy = (int)(t % g.H); long bc = t / g.H;
        const int py = iy + g.pt, px = ix + g.pl;
        int oy_lo = max(0, (py - g.kh + g.sy) / g.sy), oy_hi = min(g.OH - 1, py / g.sy);
        int ox_lo = max(0, (px - g.kw + g.sx) / g.sx), ox_hi = min(g.OW - 1, px / g.sx);
        float v = 0.f;
        if (oy_lo <= oy_hi && ox_lo <= ox_hi) v = dy[bc * g.OH * g.OW + (long)oy_hi * g.OW + ox_hi] * 1.0f * inv;
        dx[i] = v;
    }
}

static int check_pool(const char* name, const dnb_win_t* g) {
    DNB_CHECK_ARG(g, "%s: null geometry", name);
    DNB_CHECK_ARG(g->B >= 0 && g->C > 0 && g->H > 0 && g->W > 0 && g->OH > 0 && g->OW > 0 && g->kh > 0 && g->kw > 0 &&
                  g->sy > 0 && g->sx > 0, "%s: non-positive dimension", name);
    DNB_CHECK_ARG(g->pt >= 0 && g->pb >= 0 && g->pl >= 0 && g->pr >= 0, "%s: negative padding", name);
    DNB_CHECK_ARG(g->F == g->C, "%s: pooling keeps the channel count (F == C)", name);
    return DNB_OK;
}


// ---- MaxPool2D -> LeakyRelu/Relu pair, forward (dnb_maxpool2d_act_fwd) ------------------------------------------------------
// The (b,c) planes stream through the same bulk-copy ring as maxpool_fwd_ring_kernel; the outputs are 5 bytes per pooled
// element (activation value + argmax offset | gate << 4), so they leave straight from registers as coalesced stores and the
// consumer warps never meet at a block barrier: each warp waits for a stage, takes its 32-element slices of the chunk, stores,
// releases the stage.  Branch-free argmax: m = max.NaN tree (NaN iff the window holds one), offset = first position equal to m
// (first NaN when m is NaN) — np.argmax order.
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <int K, int S, bool VEC2>
__global__ void __launch_bounds__(MPR_THREADS, 2) maxpool_act_fwd_ring_kernel(const float* __restrict__ x, float* __restrict__ act,
                                                                               uint8_t* __restrict__ ig, int nchunks, int G,
                                                                               int H, int W, int OH, int OW, float alpha) {
    extern __shared__ __align__(128) uint8_t smraw[];
    const int plane = H * W, oplane = OH * OW;
    const int n_in = G * plane, n_out = G * oplane;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + MPR_ST;
    float* sin = reinterpret_cast<float*>(smraw + 128);                 // [ST][n_in]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < MPR_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], MPR_CONS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int mine = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (wid == MPR_CONS / 32) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)n_in * 4u;
            for (int n = 0; n < mine; ++n) {
                const int s = n % MPR_ST;
                if (n >= MPR_ST) tc::mbar_wait(&empty[s], ((n / MPR_ST) - 1) & 1);
                const long c = (long)blockIdx.x + (long)n * gridDim.x;
                tc::mbar_arrive_expect_tx(&full[s], bytes);
                mp_bulk_g2s(sin + (size_t)s * n_in, x + c * n_in, bytes, &full[s]);
            }
        }
        return;
    }
    int in_off[MPR_MAXR];
#pragma unroll
    for (int r = 0; r < MPR_MAXR; ++r) {
        const int e = min(tid + r * MPR_CONS, n_out - 1);
        const int p = e / oplane, o = e - p * oplane;
        const int oy = o / OW, ox = o - oy * OW;
        in_off[r] = p * plane + oy * S * W + ox * S;
    }
    for (int n = 0; n < mine; ++n) {
        const int s = n % MPR_ST;
        tc::mbar_wait(&full[s], (n / MPR_ST) & 1);
        const float* in = sin + (size_t)s * n_in;
        const long c = (long)blockIdx.x + (long)n * gridDim.x;
        float* ga = act + c * n_out;
        uint8_t* gi = ig + c * n_out;
#pragma unroll
        for (int r = 0; r < MPR_MAXR; ++r) {
            const int e = tid + r * MPR_CONS;
            if (e < n_out) {
                const float* wp = in + in_off[r];
                float v[K * K];
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (VEC2) {          // S == 2, W even, plane even: the window row starts on an 8-byte boundary
#pragma unroll
                        for (int j = 0; j + 1 < K; j += 2) {
                            const float2 f = *reinterpret_cast<const float2*>(wp + i * W + j);
                            v[i * K + j] = f.x; v[i * K + j + 1] = f.y;
                        }
                        if (K & 1) v[i * K + K - 1] = wp[i * W + K - 1];
                    } else {
#pragma unroll
                        for (int j = 0; j < K; ++j) v[i * K + j] = wp[i * W + j];
                    }
                }
                float m = v[0];
#pragma unroll
                for (int q = 1; q < K * K; ++q) m = fmax_nan(m, v[q]);
                int bi = K * K - 1;
                if (m == m) {
#pragma unroll
                    for (int q = K * K - 2; q >= 0; --q) bi = (v[q] == m) ? q : bi;
                } else {                 // np.argmax: the first NaN is the maximum
#pragma unroll
                    for (int q = K * K - 2; q >= 0; --q) bi = (v[q] != v[q]) ? q : bi;
                }
                const bool pos = m > 0.f;
                ga[e] = pos ? m : alpha * m;                                       // activations.py:17-19
                gi[e] = (uint8_t)(bi | (pos ? 16 : 0));
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);
    }
}

template <int K, int S>
static bool maxpool_act_fwd_ring(const float* x, float* act, uint8_t* ig, const dnb_win_t* g, float alpha, cudaStream_t st) {
    const long planes = (long)g->B * g->C;
    const int G = mpr_chunk_planes(planes, g->H, g->W, g->OH, g->OW, 24 * 1024, 1);
    if (!G || getenv("DNB_NO_POOL_RING")) return false;
    const long nchunks = planes / G;
    if (nchunks > 0x7fffffffL) return false;
    if (reinterpret_cast<uintptr_t>(x) & 15) return false;
    const size_t smem = 128 + (size_t)MPR_ST * G * g->H * g->W * 4;
    const int grid = (int)std::min<long>(nchunks, 2L * kNumSMs);
    const bool v2 = (S == 2) && (g->W % 2 == 0);
    if (v2) {
        cudaFuncSetAttribute(maxpool_act_fwd_ring_kernel<K, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        maxpool_act_fwd_ring_kernel<K, S, true><<<grid, MPR_THREADS, smem, st>>>(x, act, ig, (int)nchunks, G, g->H, g->W, g->OH, g->OW, alpha);
    } else {
        cudaFuncSetAttribute(maxpool_act_fwd_ring_kernel<K, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        maxpool_act_fwd_ring_kernel<K, S, false><<<grid, MPR_THREADS, smem, st>>>(x, act, ig, (int)nchunks, G, g->H, g->W, g->OH, g->OW, alpha);
    }
    return true;
}

// ---- MaxPool2D -> LeakyRelu/Relu pair, backward (dnb_pool_act_bwd, per-sample routing) -----------------------------------------
// Scatter form of maxpool_bwd_ring_kernel without block barriers: a consumer warp OWNS G/8 whole planes of every chunk — it
// zero-fills them in the shared dX chunk, adds its windows class by class ((oy mod CL, ox mod CL): windows of a class never
// overlap, so a pass is plain read-modify-writes; passes separated by __syncwarp — deterministic, no atomics) and hands its
// planes to the bulk-store engine itself.  The activation's backward (activations.py:21-22: delta where the gate bit is set,
// alpha*delta elsewhere) happens on the way in; 5 bytes per pooled element arrive, 4 bytes per input pixel leave.
constexpr int PAB_MAXR = 8;
template <int K, int S>
__global__ void __launch_bounds__(MPR_THREADS, 3) pool_act_bwd_warp_kernel(const float* __restrict__ dact, const uint8_t* __restrict__ ig,
                                                                            float* __restrict__ dx, int nchunks, int G, int ppw,
                                                                            int H, int W, int OH, int OW, float alpha) {
    constexpr int CL = (K + S - 1) / S;
    extern __shared__ __align__(128) uint8_t smraw[];
    const int plane = H * W, oplane = OH * OW;
    const int n_in = G * plane, n_out = G * oplane;
    const int stage_f = n_out + ((n_out + 3) >> 2);                     // floats per stage: deltas, then the bytes
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw);
    uint64_t* empty = full + MPR_ST;
    float* sin = reinterpret_cast<float*>(smraw + 128);                 // [ST][stage_f]
    float* sdx = sin + (size_t)MPR_ST * stage_f;                        // [2][n_in]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < MPR_ST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], MPR_CONS / 32); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int mine = (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (wid == MPR_CONS / 32) {
        if (lane == 0) {
            for (int n = 0; n < mine; ++n) {
                const int s = n % MPR_ST;
                if (n >= MPR_ST) tc::mbar_wait(&empty[s], ((n / MPR_ST) - 1) & 1);
                const long c = (long)blockIdx.x + (long)n * gridDim.x;
                tc::mbar_arrive_expect_tx(&full[s], (uint32_t)n_out * 5u);
                mp_bulk_g2s(sin + (size_t)s * stage_f, dact + c * n_out, (uint32_t)n_out * 4u, &full[s]);
                mp_bulk_g2s(sin + (size_t)s * stage_f + n_out, ig + c * n_out, (uint32_t)n_out, &full[s]);
            }
        }
        return;
    }
    // this warp's planes [p0, p0 + np) of a chunk and its windows e = lane + 32 r (chunk independent)
    const int p0 = wid * ppw, np = max(0, min(ppw, G - p0));
    const int nwin = np * oplane, npix4 = (np * plane) >> 2;
    int base[PAB_MAXR], cls[PAB_MAXR];
#pragma unroll
    for (int r = 0; r < PAB_MAXR; ++r) {
        const int e = lane + 32 * r;
        const int ec = min(e, max(nwin - 1, 0));
        const int p = ec / oplane, o = ec - p * oplane;
        const int oy = o / OW, ox = o - oy * OW;
        base[r] = (p0 + p) * plane + oy * S * W + ox * S;
        cls[r] = e < nwin ? (oy % CL) * CL + (ox % CL) : -1;
    }
    for (int n = 0; n < mine; ++n) {
        const int s = n % MPR_ST, ob = n & 1;
        float* out = sdx + (size_t)ob * n_in;
        // this warp's part of buffer ob was handed to the store engine two chunks ago: at most ONE newer store may be unread
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        float4* o4 = reinterpret_cast<float4*>(out + p0 * plane);
        for (int e = lane; e < npix4; e += 32) o4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        tc::mbar_wait(&full[s], (n / MPR_ST) & 1);
        const float* sd = sin + (size_t)s * stage_f;
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(sd + n_out);
        float d[PAB_MAXR]; int tgt[PAB_MAXR];
#pragma unroll
        for (int r = 0; r < PAB_MAXR; ++r) {
            if (cls[r] >= 0) {
                const int e = p0 * oplane + lane + 32 * r;
                const unsigned ib = sb[e];
                const int off = (int)(ib & 15u);
                const float dv = sd[e];
                d[r] = (ib & 16u) ? dv : alpha * dv;
                tgt[r] = base[r] + (off / K) * W + (off % K);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[s]);                      // the stage has been read into registers
#pragma unroll
        for (int c = 0; c < CL * CL; ++c) {
#pragma unroll
            for (int r = 0; r < PAB_MAXR; ++r)
                if (cls[r] == c) out[tgt[r]] += d[r];
            __syncwarp();
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && np > 0) {
            const long c = (long)blockIdx.x + (long)n * gridDim.x;
            mp_bulk_s2g(dx + c * n_in + (long)p0 * plane, out + p0 * plane, (uint32_t)(np * plane) * 4u);
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int K, int S>
static bool pool_act_bwd_warp(const float* dact, const uint8_t* ig, float* dx, const dnb_win_t* g, float alpha, cudaStream_t st) {
    const long planes = (long)g->B * g->C;
    const long plane = (long)g->H * g->W, oplane = (long)g->OH * g->OW;
    if (plane % 4 || getenv("DNB_NO_POOL_RING")) return false;
    // planes per chunk: a multiple of the 8 consumer warps, <= 18 KB of dX per buffer, <= 256 windows per warp, the chunk's
    // deltas and bytes whole 16-byte groups
    int G = 0;
    for (long c = std::min<long>({planes, (18 * 1024) / (plane * 4), 8 * (32L * PAB_MAXR / oplane)}) / 8 * 8; c >= 8; c -= 8)
        if (planes % c == 0 && (c * oplane) % 16 == 0) { G = (int)c; break; }
    if (!G) return false;
    const long nchunks = planes / G;
    if (nchunks > 0x7fffffffL) return false;
    if ((reinterpret_cast<uintptr_t>(dact) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(ig)) & 15) return false;
    const long n_out = G * oplane;
    const size_t smem = 128 + ((size_t)MPR_ST * (n_out + ((n_out + 3) >> 2)) + 2ull * G * plane) * 4;
    const int grid = (int)std::min<long>(nchunks, 3L * kNumSMs);
    cudaFuncSetAttribute(pool_act_bwd_warp_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pool_act_bwd_warp_kernel<K, S><<<grid, MPR_THREADS, smem, st>>>(dact, ig, dx, (int)nchunks, G, G / 8, g->H, g->W, g->OH, g->OW, alpha);
    return true;
}

// MaxPool2D.forward + LeakyRelu.forward, any geometry with <= 16 window positions: act and the byte tensor only
__global__ void __launch_bounds__(256) maxpool_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ act,
                                                              uint8_t* __restrict__ ig, dnb_win_t g, float alpha) {
    const long n = (long)g.B * g.C * g.OH * g.OW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int ox = (int)(i % g.OW); long t = i / g.OW; int oy = (int)(t % g.OH); long bc = t / g.OH;
        const float* xp = x + bc * g.H * g.W;
        float best = 0.f; int bi = 0; bool first = true;
        for (int ii = 0; ii < g.kh; ++ii) {
            int iy = oy * g.sy + ii - g.pt;
            bool rowin = (iy >= 0 && iy < g.H);
            for (int jj = 0; jj < g.kw; ++jj) {
                int ix = ox * g.sx + jj - g.pl;
                float v = (rowin && ix >= 0 && ix < g.W) ? __ldg(xp + (long)iy * g.W + ix) : 0.f;   // zero pad
                if (first || v > best || (v != v && best == best)) { best = v; bi = ii * g.kw + jj; first = false; }
            }
        }
        const bool pos = best > 0.f;
        act[i] = pos ? best : alpha * best;
        ig[i] = (uint8_t)(bi | (pos ? 16 : 0));
    }
}

// ---- LeakyRelu.backward + MaxPool2D.backward (PER_SAMPLE routing) from the byte tensor -------------------------------
// dx[plane][(oy*S + off/K)*W + ox*S + off%K] += gate ? delta : alpha*delta.  A CTA builds PB_G planes at a time in shared memory
// (zero fill, one shared atomicAdd per pooled element — overlapping windows may hit the same pixel) and writes them out with
// 128-bit stores: the only HBM traffic is delta + bytes in, dx out.
constexpr int PB_G = 32;
__global__ void __launch_bounds__(256) pool_act_bwd_kernel(const float* __restrict__ dact, const uint8_t* __restrict__ idxg,
                                                           float* __restrict__ dx, long planes, int H, int W, int OH, int OW,
                                                           int K, int S, float alpha) {
    extern __shared__ __align__(16) float sp[];           // [PB_G][H*W]
    const int plane = H * W, OP = OH * OW;
    const long groups = (planes + PB_G - 1) / PB_G;
    for (long gidx = blockIdx.x; gidx < groups; gidx += gridDim.x) {
        const long p0 = gidx * PB_G;
        const int np = (int)min((long)PB_G, planes - p0);
        const int n_in = np * plane, n_out = np * OP;
        for (int e = threadIdx.x; e < (n_in >> 2); e += 256) reinterpret_cast<float4*>(sp)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = (n_in & ~3) + threadIdx.x; e < n_in; e += 256) sp[e] = 0.f;
        __syncthreads();
        for (int e = threadIdx.x; e < n_out; e += 256) {
            const int pl = e / OP, o = e - pl * OP;
            const int oy = o / OW, ox = o - oy * OW;
            const float d = __ldg(dact + p0 * OP + e);
            const unsigned ib = __ldg(idxg + p0 * OP + e);
            const int off = (int)(ib & 15u);
            const float dd = (ib & 16u) ? d : alpha * d;
            atomicAdd(sp + pl * plane + (oy * S + off / K) * W + ox * S + off % K, dd);
        }
        __syncthreads();
        float* out = dx + p0 * plane;
        if (((plane * PB_G) & 3) == 0 && !(reinterpret_cast<uintptr_t>(dx) & 15)) {
            for (int e = threadIdx.x; e < (n_in >> 2); e += 256) reinterpret_cast<float4*>(out)[e] = reinterpret_cast<const float4*>(sp)[e];
            for (int e = (n_in & ~3) + threadIdx.x; e < n_in; e += 256) out[e] = sp[e];
        } else {
            for (int e = threadIdx.x; e < n_in; e += 256) out[e] = sp[e];
        }
        __syncthreads();
    }
}

}  // namespace dnb

using namespace dnb;

extern "C" {

int dnb_maxpool2d_fwd(const float* x, float* y, int32_t* idx, const dnb_win_t* g, dnb_stream_t stream) {
    int rc = check_pool("maxpool2d_fwd", g);
    if (rc) return rc;
    long n = (long)g->B * g->C * g->OH * g->OW;
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(x && y && idx, "maxpool2d_fwd: null pointer");
    const bool nopad = !(g->pt | g->pb | g->pl | g->pr);
    if (nopad && g->kh == g->kw && g->sy == g->sx && (g->OH - 1) * g->sy + g->kh <= g->H && (g->OW - 1) * g->sx + g->kw <= g->W) {
        bool done = false;
        if (g->kh == 3 && g->sy == 2) done = maxpool_fwd_ring<3, 2>(x, y, idx, g, S(stream));
        else if (g->kh == 2 && g->sy == 2) done = maxpool_fwd_ring<2, 2>(x, y, idx, g, S(stream));
        else if (g->kh == 3 && g->sy == 1) done = maxpool_fwd_ring<3, 1>(x, y, idx, g, S(stream));
        else if (g->kh == 2 && g->sy == 1) done = maxpool_fwd_ring<2, 1>(x, y, idx, g, S(stream));
        if (done) { DNB_LAUNCH_CHECK("maxpool2d_fwd"); return DNB_OK; }
    }
    const bool even = (g->W % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    if (nopad && even && g->kh == g->kw && g->sy == g->sx && g->sy == 2 && (g->kh == 2 || g->kh == 3) &&
        (g->OH - 1) * g->sy + g->kh <= g->H) {
        const long strips = (long)g->B * g->C * g->OH * ((g->OW + 3) / 4);
        const int planes = g->B * g->C;
        if (g->kh == 3) maxpool_fwd_strip_kernel<3, 2><<<ew_grid(strips, 256), 256, 0, S(stream)>>>(x, y, idx, planes, g->H, g->W, g->OH, g->OW);
        else maxpool_fwd_strip_kernel<2, 2><<<ew_grid(strips, 256), 256, 0, S(stream)>>>(x, y, idx, planes, g->H, g->W, g->OH, g->OW);
        DNB_LAUNCH_CHECK("maxpool2d_fwd");
        return DNB_OK;
    }
    maxpool_fwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, y, idx, *g);
    DNB_LAUNCH_CHECK("maxpool2d_fwd");
    return DNB_OK;
}

size_t dnb_maxpool2d_bwd_ws_bytes(const dnb_win_t* g, int route) {
    if (!g || route != DNB_ROUTE_LITERAL) return 0;
    return (size_t)g->C * g->OH * g->OW * (g->kh * g->kw + 1) * sizeof(int32_t);      // last[npos][P] + done[npos]
}

int dnb_pool_act_supported(const dnb_win_t* pool) {
    if (!pool || check_pool("pool_act_supported", pool)) return 0;
    if (pool->kh != pool->kw || pool->sy != pool->sx || (pool->pt | pool->pb | pool->pl | pool->pr) || pool->kh * pool->kw > 16) return 0;
    if ((pool->OH - 1) * pool->sy + pool->kh > pool->H || (pool->OW - 1) * pool->sx + pool->kw > pool->W) return 0;
    return (size_t)PB_G * pool->H * pool->W * sizeof(float) <= 200 * 1024 ? 1 : 0;
}

int dnb_maxpool2d_act_fwd(const float* x, float* act_out, uint8_t* idx_gate, const dnb_win_t* g, float alpha, dnb_stream_t stream) {
    int rc = check_pool("maxpool2d_act_fwd", g);
    if (rc) return rc;
    if (g->kh * g->kw > 16) return set_err(DNB_ERR_UNSUPPORTED, "maxpool2d_act_fwd: windows of more than 16 positions do not fit the byte format");
    const long n = (long)g->B * g->C * g->OH * g->OW;
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(x && act_out && idx_gate, "maxpool2d_act_fwd: null pointer");
    if (!(g->pt | g->pb | g->pl | g->pr) && g->kh == g->kw && g->sy == g->sx &&
        (g->OH - 1) * g->sy + g->kh <= g->H && (g->OW - 1) * g->sx + g->kw <= g->W) {
        bool done = false;
        if (g->kh == 3 && g->sy == 2) done = maxpool_act_fwd_ring<3, 2>(x, act_out, idx_gate, g, alpha, S(stream));
        else if (g->kh == 2 && g->sy == 2) done = maxpool_act_fwd_ring<2, 2>(x, act_out, idx_gate, g, alpha, S(stream));
        else if (g->kh == 3 && g->sy == 1) done = maxpool_act_fwd_ring<3, 1>(x, act_out, idx_gate, g, alpha, S(stream));
        else if (g->kh == 2 && g->sy == 1) done = maxpool_act_fwd_ring<2, 1>(x, act_out, idx_gate, g, alpha, S(stream));
        if (done) { DNB_LAUNCH_CHECK("maxpool2d_act_fwd"); return DNB_OK; }
    }
    maxpool_act_fwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, act_out, idx_gate, *g, alpha);
    DNB_LAUNCH_CHECK("maxpool2d_act_fwd");
    return DNB_OK;
}

/* LeakyRelu.backward (activations.py:21-22) + MaxPool2D.backward with the per-sample routing, straight from the activation's
 * incoming delta and the byte tensor: dx is the pool's parent delta [B][C][H][W] */
int dnb_pool_act_bwd(const float* act_delta, const uint8_t* idx_gate, float* dx, const dnb_win_t* pool, float alpha,
                     dnb_stream_t stream) {
    DNB_CHECK_ARG(pool && pool->kh == pool->kw && pool->sy == pool->sx && !(pool->pt | pool->pb | pool->pl | pool->pr) &&
                  (pool->OH - 1) * pool->sy + pool->kh <= pool->H && (pool->OW - 1) * pool->sx + pool->kw <= pool->W,
                  "pool_act_bwd: square unpadded pools only");
    const long planes = (long)pool->B * pool->C;
    if (planes == 0) return DNB_OK;
    DNB_CHECK_ARG(act_delta && idx_gate && dx, "pool_act_bwd: null pointer");
    {
        bool done = false;
        if (pool->kh == 3 && pool->sy == 2) done = pool_act_bwd_warp<3, 2>(act_delta, idx_gate, dx, pool, alpha, S(stream));
        else if (pool->kh == 2 && pool->sy == 2) done = pool_act_bwd_warp<2, 2>(act_delta, idx_gate, dx, pool, alpha, S(stream));
        else if (pool->kh == 3 && pool->sy == 1) done = pool_act_bwd_warp<3, 1>(act_delta, idx_gate, dx, pool, alpha, S(stream));
        else if (pool->kh == 2 && pool->sy == 1) done = pool_act_bwd_warp<2, 1>(act_delta, idx_gate, dx, pool, alpha, S(stream));
        if (done) { DNB_LAUNCH_CHECK("pool_act_bwd"); return DNB_OK; }
    }
    const size_t smem = (size_t)PB_G * pool->H * pool->W * sizeof(float);
    DNB_CHECK_ARG(smem <= 200 * 1024, "pool_act_bwd: plane of %d x %d is too large", pool->H, pool->W);
    cudaFuncSetAttribute(pool_act_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long groups = (planes + PB_G - 1) / PB_G;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
    pool_act_bwd_kernel<<<(unsigned)std::min<long>(groups, (long)kNumSMs * per_sm), 256, smem, S(stream)>>>(
        act_delta, idx_gate, dx, planes, pool->H, pool->W, pool->OH, pool->OW, pool->kh, pool->sy, alpha);
    DNB_LAUNCH_CHECK("pool_act_bwd");
    return DNB_OK;
}


int dnb_maxpool2d_bwd(const float* dy, const int32_t* idx, float* dx, const dnb_win_t* g,
                      int route, void* ws, dnb_stream_t stream) {
    DNB_CHECK_ARG(route == DNB_ROUTE_PER_SAMPLE || (g && g->B == 1) || ws, "maxpool2d_bwd: the literal route needs a workspace");
    int rc = check_pool("maxpool2d_bwd", g);
    if (rc) return rc;
    long n = (long)g->B * g->C * g->H * g->W;
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(dy && idx && dx, "maxpool2d_bwd: null pointer");
    cudaStream_t st = S(stream);
    if (route == DNB_ROUTE_PER_SAMPLE || g->B == 1) {
        const bool nopad = !(g->pt | g->pb | g->pl | g->pr);
        if (nopad && g->kh == g->kw && g->sy == g->sx) {
            bool done = false;
            if (g->kh == 3 && g->sy == 2) done = maxpool_bwd_ring<3, 2>(dy, idx, dx, g, st);
            else if (g->kh == 2 && g->sy == 2) done = maxpool_bwd_ring<2, 2>(dy, idx, dx, g, st);
            else if (g->kh == 3 && g->sy == 1) done = maxpool_bwd_ring<3, 1>(dy, idx, dx, g, st);
            else if (g->kh == 2 && g->sy == 1) done = maxpool_bwd_ring<2, 1>(dy, idx, dx, g, st);
            if (done) { DNB_LAUNCH_CHECK("maxpool2d_bwd"); return DNB_OK; }
        }
        if (nopad && g->kh == g->kw && g->sy == 2 && g->sx == 2 && (g->kh == 2 || g->kh == 3)) {
            const long strips = (long)g->B * g->C * g->H * ((g->W + 3) / 4);
            if (g->kh == 3) maxpool_bwd_ps_strip_kernel<3><<<ew_grid(strips, 256), 256, 0, st>>>(dy, idx, dx, g->B * g->C, g->H, g->W, g->OH, g->OW);
            else maxpool_bwd_ps_strip_kernel<2><<<ew_grid(strips, 256), 256, 0, st>>>(dy, idx, dx, g->B * g->C, g->H, g->W, g->OH, g->OW);
            DNB_LAUNCH_CHECK("maxpool2d_bwd");
            return DNB_OK;
        }
        maxpool_bwd_per_sample_kernel<<<ew_grid(n, 256), 256, 0, st>>>(dy, idx, dx, *g);
        DNB_LAUNCH_CHECK("maxpool2d_bwd");
        return DNB_OK;
    }
    DNB_CHECK_ARG(route == DNB_ROUTE_LITERAL, "maxpool2d_bwd: unknown route %d", route);
    rc = dnb_maxpool2d_bwd_prepare(idx, g, ws, stream);
    if (rc) return rc;
    return dnb_maxpool2d_bwd_apply(dy, dx, g, ws, stream);
}

// The LITERAL route in two halves, so a host can run the first (it depends on the forward's argmax offsets only) on a side
// stream as soon as the forward has finished — latency-bound kernels hidden behind the rest of the step:
//   prepare: ws <- last[c,oy,ox,p] = the newest sample whose window (c,oy,ox) has its maximum at offset p
//   apply  : dx[0] <- gather of dY through `last`, dx[1..B-1] <- 0
int dnb_maxpool2d_bwd_prepare(const int32_t* idx, const dnb_win_t* g, void* ws, dnb_stream_t stream) {
    int rc = check_pool("maxpool2d_bwd", g);
    if (rc) return rc;
    if ((long)g->B * g->C * g->H * g->W == 0) return DNB_OK;
    DNB_CHECK_ARG(idx && ws, "maxpool2d_bwd: the literal route needs the argmax offsets and a workspace");
    cudaStream_t st = S(stream);
    const long npos = (long)g->C * g->OH * g->OW;
    const int P = g->kh * g->kw;
    int32_t* last = static_cast<int32_t*>(ws);
    cudaError_t e = cudaMemsetAsync(last, 0xff, (size_t)npos * P * sizeof(int32_t), st);     // -1
    if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "maxpool2d_bwd: %s", cudaGetErrorString(e));
    int32_t* done = last + npos * P;
    const int top = LT_TOP, bchunk = 64;
    DNB_CHECK_ARG((size_t)32 * P * sizeof(int32_t) <= 48 * 1024, "maxpool2d_bwd: window of %d elements is too large for the literal route", P);
    maxpool_last_top_kernel<<<(unsigned)((npos + 31) / 32), LT_WARPS * 32, (size_t)32 * P * sizeof(int32_t), st>>>(idx, last, done, *g);
    DNB_LAUNCH_CHECK("maxpool2d_bwd_last_top");
    if (g->B > top) {
        dim3 grid((unsigned)((npos + 255) / 256), (g->B - top + bchunk - 1) / bchunk);
        maxpool_last_kernel<<<grid, 256, 0, st>>>(idx, last, done, *g, bchunk, g->B - top);
        DNB_LAUNCH_CHECK("maxpool2d_bwd_last");
    }
    return DNB_OK;
}

int dnb_maxpool2d_bwd_apply(const float* dy, float* dx, const dnb_win_t* g, const void* ws, dnb_stream_t stream) {
    int rc = check_pool("maxpool2d_bwd", g);
    if (rc) return rc;
    const long n = (long)g->B * g->C * g->H * g->W;
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(dy && dx && ws, "maxpool2d_bwd: null pointer");
    cudaStream_t st = S(stream);
    const int32_t* last = static_cast<const int32_t*>(ws);
    const long n0 = (long)g->C * g->H * g->W;
    if (g->B > 1) {
        cudaError_t e = cudaMemsetAsync(dx + n0, 0, (size_t)(n - n0) * sizeof(float), st);   // samples 1..B-1 get no delta
        if (e != cudaSuccess) return set_err(DNB_ERR_CUDA, "maxpool2d_bwd: %s", cudaGetErrorString(e));
    }
    maxpool_bwd_literal_kernel<<<ew_grid(n0, 256), 256, 0, st>>>(dy, last, dx, *g);
    DNB_LAUNCH_CHECK("maxpool2d_bwd_literal");
    return DNB_OK;
}

int dnb_avgpool2d_fwd(const float* x, float* y, const dnb_win_t* g, dnb_stream_t stream) {
    int rc = check_pool("avgpool2d_fwd", g);
    if (rc) return rc;
    long n = (long)g->B * g->C * g->OH * g->OW;
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(x && y, "avgpool2d_fwd: null pointer");
    avgpool_fwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(x, y, *g);
    DNB_LAUNCH_CHECK("avgpool2d_fwd");
    return DNB_OK;
}

int dnb_avgpool2d_bwd(const float* dy, float* dx, const dnb_win_t* g, dnb_stream_t stream) {
    int rc = check_pool("avgpool2d_bwd", g);
    if (rc) return rc;
    // pooling.py:169 tiles the delta exactly 4 times: any other window size raises in the reference
    DNB_CHECK_ARG(g->kh * g->kw == 4, "avgpool2d_bwd: the reference backward only supports 4-element pools (got %dx%d)", g->kh, g->kw);
    long n = (long)g->B * g->C * g->H * g->W;
    if (n == 0) return DNB_OK;
    DNB_CHECK_ARG(dy && dx, "avgpool2d_bwd: null pointer");
    avgpool_bwd_kernel<<<ew_grid(n, 256), 256, 0, S(stream)>>>(dy, dx, *g);
    DNB_LAUNCH_CHECK("avgpool2d_bwd");
    return DNB_OK;
}

}  // extern "C"
