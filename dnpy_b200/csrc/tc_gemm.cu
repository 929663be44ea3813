// TF32 tensor-core GEMM for the dense contractions of the hot path (Dense fwd / dX / dW and the 1x1
// convolutions): C[M,N] = A[M,K] · B[K,N] with fp32 storage, TF32 operands and fp32 accumulation.
//
// Blackwell-native structure (no mma.sync, no wgmma):
//   * operands are fetched by TMA (cp.async.bulk.tensor, 128B swizzle, zero fill past the edges)
//     into a 4-stage shared-memory ring, signalled through mbarrier transaction counts;
//   * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8 per instruction) with the
//     accumulator tile living in TMEM; tcgen05.commit releases ring slots / publishes the tile;
//   * four epilogue warps read the accumulator with tcgen05.ld (32 lanes x 32 columns) and apply the
//     fused epilogue: bias add (Dense.forward) or "+= (acc + reg'(W)) / m" (Dense.backward dW);
//   * both operand orientations are supported without a transpose pass: K-major tiles use the
//     canonical K-major layout, M/N-contiguous matrices (X^T, W, dY as B operand) are loaded as
//     32x32 boxes into the canonical MN-major layout and flagged in the instruction descriptor;
//   * split-K over gridDim.z (atomic accumulate epilogue) for the skinny dW shapes.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include "tc_gemm.cuh"

namespace dnb {
namespace tc {

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swz) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (swz == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool make_tmap_bf16_sw64(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

constexpr int BM = 128, BK = 32;
// ring depth: a TMA round trip is ~0.8 us on B200, so a k block costs latency / stages while the ring is latency bound —
// as many stages as fit next to the epilogue staging (BN = 128: 6 x 32 KB, BN = 64: 8 x 24 KB, BN = 32: 10 x 20 KB)
template <int BN> struct GemmStages { static constexpr int value = BN == 128 ? 6 : (BN == 64 ? 8 : 10); };
constexpr int GEMM_THREADS = 192;     // warp 0: TMA, warp 1: MMA issue, warps 2-5: epilogue

struct GemmParams {
    float* C; const float* bias; const float* W;
    int M, N, K;
    int a_mn, b_mn;             // operand stored MN-contiguous?
    int grad;                   // epilogue: 0 store(+bias), 1 accumulate gradient
    float inv_m, l1, l2;
    int kb_per_split;           // k-blocks handled by one z-slice
    unsigned long long* prof;   // DNB_GEMM_PROF: globaltimer stamps of CTA (0,0,0)
};
__device__ __forceinline__ void gemm_stamp(const GemmParams& p, int slot) {
    if (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.prof[slot] = g;
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGES = GemmStages<BN>::value;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) gemm_stamp(p, 0);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int num_kb_total = (p.K + BK - 1) / BK;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) gemm_stamp(p, 1);

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES, k = (kb_begin + i) * BK;
            if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
            mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
            uint8_t* a = sA + s * A_BYTES;
            uint8_t* b = sB + s * B_BYTES;
            if (!p.a_mn) tma_load_2d(a, &tmA, k, m0, &full[s]);                       // box {32 k, 128 m}
            else for (int j = 0; j < BM / 32; ++j) tma_load_2d(a + j * 4096, &tmA, m0 + j * 32, k, &full[s]);   // box {32 m, 32 k}
            if (!p.b_mn) tma_load_2d(b, &tmB, k, n0, &full[s]);                       // box {32 k, BN n}
            else for (int j = 0; j < BN / 32; ++j) tma_load_2d(b + j * 4096, &tmB, n0 + j * 32, k, &full[s]);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
        const uint32_t idesc = make_idesc_tf32(BM, BN, p.a_mn, p.b_mn);
        // K-major: step 32 bytes inside the swizzled 128-byte row; MN-major (32B-atom layout): next 8-row K group (1024 B)
        const uint64_t a0 = p.a_mn ? make_smem_desc(smem_u32(sA), 4096, 512, LAYOUT_SW128_BASE32B) : make_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t b0 = p.b_mn ? make_smem_desc(smem_u32(sB), 4096, 512, LAYOUT_SW128_BASE32B) : make_smem_desc(smem_u32(sB), 16, 1024);
        const uint32_t a_step = p.a_mn ? (1024 >> 4) : (32 >> 4), b_step = p.b_mn ? (1024 >> 4) : (32 >> 4);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            mbar_wait(&full[s], (i / STAGES) & 1);
            tc_fence_after();
            if (i == 0 && lane == 0) gemm_stamp(p, 2);
            const uint64_t a = a0 + (uint64_t)(s * (A_BYTES >> 4)), b = b0 + (uint64_t)(s * (B_BYTES >> 4));
            if (elect_one()) {
                if (i == 0) mma_tf32_init(tmem_base, a, b, idesc); else mma_tf32_acc(tmem_base, a, b, idesc);
                mma_tf32_acc(tmem_base, a + a_step, b + b_step, idesc);
                mma_tf32_acc(tmem_base, a + 2 * a_step, b + 2 * b_step, idesc);
                mma_tf32_acc(tmem_base, a + 3 * a_step, b + 3 * b_step, idesc);
                mma_commit(&empty[s]);          // slot reusable once these MMAs have read it
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(acc_full);  // accumulator tile complete
    } else if (warp >= 2) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;             // TMEM lane quadrant this warp may access
        const int row = m0 + q * 32 + lane;
        if (nkb > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        if (warp == 2 && lane == 0) gemm_stamp(p, 3);
        // The accumulator row of a thread is one row of C: storing it directly scatters every warp store over 32 rows
        // (32 sectors per instruction — 9 us for a 128 x 128 tile on B200).  Stage the tile in the (now idle) operand ring,
        // then every warp writes whole rows: 16 bytes per lane, one contiguous segment per instruction.
        constexpr int PITCH = BN + 4;                       // floats; rows 16-byte aligned, conflict-free for 128-bit accesses
        float* stg = reinterpret_cast<float*>(sA);
        const int rl = q * 32 + lane;                       // tile row of this thread
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            if (nkb > 0) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(stg + rl * PITCH + c0 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        if (warp == 2 && lane == 0) gemm_stamp(p, 5);
        asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps
        if (warp == 2 && lane == 0) gemm_stamp(p, 6);
        const bool vec = (p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
        // a lane owns ONE 4-column group of the tile (BN <= 128): everything that depends on the column only — the bias,
        // the regulariser term of a weight-gradient tile — stays out of the row loop, and the rows go RB at a time (loads of
        // all RB rows in flight before the first store: the loop is latency bound otherwise, ~5 us per tile on B200)
        static_assert(BN / 4 <= 32, "one column group per lane");
        const int c4 = lane, col = n0 + c4 * 4;
        if (c4 < BN / 4 && col < p.N) {
            const int nv = min(4, p.N - col);
            float bias4[4] = {0.f, 0.f, 0.f, 0.f};
            if (!p.grad && p.bias)
                for (int e = 0; e < nv; ++e) bias4[e] = __ldg(p.bias + col + e);
            const bool reg = p.grad && blockIdx.z == 0 && (p.l1 != 0.f || p.l2 != 0.f);
            constexpr int RB = 8;
            for (int r0 = warp - 2; r0 < BM; r0 += 4 * RB) {
                float a[RB][4], c[RB][4];
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int r = r0 + 4 * u, grow = m0 + r;
                    if (r < BM && grow < p.M) {
                        const float4 a4 = *reinterpret_cast<const float4*>(stg + r * PITCH + c4 * 4);
                        a[u][0] = a4.x; a[u][1] = a4.y; a[u][2] = a4.z; a[u][3] = a4.w;
                        const size_t o = (size_t)grow * p.N + col;
                        if (p.grad && gridDim.z == 1) {
                            if (vec) {
                                const float4 c4v = *reinterpret_cast<const float4*>(p.C + o);
                                c[u][0] = c4v.x; c[u][1] = c4v.y; c[u][2] = c4v.z; c[u][3] = c4v.w;
                            } else {
                                for (int e = 0; e < nv; ++e) c[u][e] = p.C[o + e];
                            }
                        }
                        if (reg)
                            for (int e = 0; e < nv; ++e) c[u][e] = (p.grad && gridDim.z == 1 ? c[u][e] : 0.f) + reg_term(p.W[o + e], p.l1, p.l2) * p.inv_m;
                    }
                }
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int r = r0 + 4 * u, grow = m0 + r;
                    if (r < BM && grow < p.M) {
                        const size_t o = (size_t)grow * p.N + col;
                        if (!p.grad) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) a[u][e] += bias4[e];
                            if (vec) *reinterpret_cast<float4*>(p.C + o) = make_float4(a[u][0], a[u][1], a[u][2], a[u][3]);
                            else for (int e = 0; e < nv; ++e) p.C[o + e] = a[u][e];
                        } else if (gridDim.z == 1) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) a[u][e] = a[u][e] * p.inv_m + c[u][e];
                            if (vec) *reinterpret_cast<float4*>(p.C + o) = make_float4(a[u][0], a[u][1], a[u][2], a[u][3]);
                            else for (int e = 0; e < nv; ++e) p.C[o + e] = a[u][e];
                        } else if (vec) {           // split-K partial: one 16-byte vector reduction per lane
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.C + o),
                                         "f"(a[u][0] * p.inv_m + (reg ? c[u][0] : 0.f)), "f"(a[u][1] * p.inv_m + (reg ? c[u][1] : 0.f)),
                                         "f"(a[u][2] * p.inv_m + (reg ? c[u][2] : 0.f)), "f"(a[u][3] * p.inv_m + (reg ? c[u][3] : 0.f))
                                         : "memory");
                        } else {
                            for (int e = 0; e < nv; ++e) atomicAdd(p.C + o + e, a[u][e] * p.inv_m + (reg ? c[u][e] : 0.f));
                        }
                    }
                }
            }
        }
    }
    if (warp == 2 && lane == 0) gemm_stamp(p, 7);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) gemm_stamp(p, 4);
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int splits, cudaStream_t st) {
    constexpr size_t smem = GemmStages<BN>::value * (BM * BK * 4 + BN * BK * 4) + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
    tc_gemm_kernel<BN><<<grid, GEMM_THREADS, smem, st>>>(ta, tb, p);
    DNB_LAUNCH_CHECK("tc_gemm_tf32");
    return DNB_OK;
}

}  // namespace tc

bool tc_gemm_supported(int M, int N, int K, bool a_kmajor, bool b_kmajor) {
    if (!tc::get_encode_tiled()) return false;
    if (M < 1 || N < 32 || K < 8) return false;
    // TMA: 16-byte aligned row pitch of the stored matrices
    if (a_kmajor ? (K % 4) : (M % 4)) return false;
    if (b_kmajor ? (K % 4) : (N % 4)) return false;
    return true;
}

int tc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int flags,
            const float* bias, const float* W, float inv_m, float l1, float l2, cudaStream_t st) {
    using namespace tc;
    const int a_mn = (flags & TC_A_MMAJOR) ? 1 : 0, b_mn = (flags & TC_B_NMAJOR) ? 1 : 0, grad = (flags & TC_ACCUM_GRAD) ? 1 : 0;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return set_err(DNB_ERR_INVALID, "tc_gemm: operands must be 16-byte aligned");
    int BN = N >= 96 ? 128 : (N >= 48 ? 64 : 32);
    // small problems: narrower N tiles until the grid covers most of the SMs — a CTA pulls its operands from L2 at
    // ~64 B/clk, so a 32-CTA grid leaves the k loop (270 ns per 32-wide k block at BN = 128) bound by per-SM L2 bandwidth
    // (weight gradients fill the machine by splitting K instead)
    if (!grad)
        while (BN > 32 && ((M + BM - 1) / BM) * ((N + BN - 1) / BN) < (kNumSMs * 3) / 4) BN >>= 1;
    CUtensorMap ta, tb;
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M; str[0] = (uint64_t)K * 4; box[0] = BK; box[1] = BM; }
    else       { dims[0] = M; dims[1] = K; str[0] = (uint64_t)M * 4; box[0] = 32; box[1] = BK; }
    if (!make_tmap_f32(&ta, A, 2, dims, str, box, a_mn)) return set_err(DNB_ERR_CUDA, "tc_gemm: cuTensorMapEncodeTiled(A) failed");
    if (!b_mn) { dims[0] = K; dims[1] = N; str[0] = (uint64_t)K * 4; box[0] = BK; box[1] = (uint32_t)BN; }
    else       { dims[0] = N; dims[1] = K; str[0] = (uint64_t)N * 4; box[0] = 32; box[1] = BK; }
    if (!make_tmap_f32(&tb, B, 2, dims, str, box, b_mn)) return set_err(DNB_ERR_CUDA, "tc_gemm: cuTensorMapEncodeTiled(B) failed");

    const int nkb = (K + BK - 1) / BK;
    int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int splits = 1;
    if (grad) {                       // skinny weight-gradient shapes: fill the 148 SMs by splitting K
        splits = max(1, min(nkb / 4, (kNumSMs + tiles - 1) / tiles));
    }
    int per = (nkb + splits - 1) / splits;
    splits = (nkb + per - 1) / per;
    GemmParams p{C, bias, W, M, N, K, a_mn, b_mn, grad, inv_m, l1, l2, per, nullptr};
    static unsigned long long* dprof = nullptr;
    if (getenv("DNB_GEMM_PROF")) {
        if (!dprof) cudaMalloc(&dprof, 8 * sizeof(unsigned long long));
        p.prof = dprof;
    }
    const int rc = BN == 128 ? launch<128>(ta, tb, p, splits, st) : (BN == 64 ? launch<64>(ta, tb, p, splits, st) : launch<32>(ta, tb, p, splits, st));
    if (p.prof && rc == DNB_OK) {
        unsigned long long h[8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.prof, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[gemm prof] M=%d N=%d K=%d BN=%d splits=%d: setup %llu ns, first stage landed +%llu, accumulator complete +%llu (staged +%llu, barrier +%llu, stored +%llu), epilogue done +%llu\n",
                M, N, K, BN, splits, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[4] - h[0]);
    }
    return rc;
}

}  // namespace dnb

extern "C" int dnb_gemm_tf32(const float* A, const float* B, float* C, int M, int N, int K,
                             int a_mmajor, int b_nmajor, dnb_stream_t stream) {
    using namespace dnb;
    DNB_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm_tf32: bad arguments");
    if (!tc_gemm_supported(M, N, K, !a_mmajor, !b_nmajor))
        return set_err(DNB_ERR_UNSUPPORTED, "gemm_tf32: shape M=%d N=%d K=%d not supported by the tensor-core path", M, N, K);
    return tc_gemm(A, B, C, M, N, K, (a_mmajor ? TC_A_MMAJOR : 0) | (b_nmajor ? TC_B_NMAJOR : 0), nullptr, nullptr,
                   0.f, 0.f, 0.f, S(stream));
}
