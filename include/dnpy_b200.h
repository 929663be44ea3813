/*
 * dnpy_b200 — C ABI of the B200 (sm_100a) execution backend for the dnpy hot path.
 *
 * The reference (salvacarrion/dnpy) is pure Python/NumPy and has no FFI of its
 * own; the boundary it exposes is the duck-typed Layer / Loss / Optimizer
 * protocol (dnpy/layers/base.py:6-41, dnpy/losses.py:4-14,
 * dnpy/optimizers.py:4-12) driven by Net.fit (dnpy/net.py:174-195).  Each entry
 * point below is what a ctypes binding for one of those methods calls; the
 * reference method it replaces is cited next to it (paths relative to the
 * reference root).  INTEGRATION.md shows the reference-side stub for each.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to dense row-major fp32 (int32 where
 *     stated), NCHW for images, exactly the reference's array layouts;
 *   - no hidden allocation, no host synchronisation: work is enqueued on
 *     `stream` (a cudaStream_t passed as void*; NULL = legacy default stream),
 *     so a whole training step can be captured in a CUDA graph;
 *   - workspaces are caller-provided; `*_ws_bytes` functions return the size;
 *   - return value: 0 on success, <0 on error (DNB_ERR_*).  The message is
 *     available from dnb_last_error() (thread-local).  The Python host maps
 *     DNB_ERR_INVALID -> ValueError, DNB_ERR_UNSUPPORTED -> NotImplementedError,
 *     DNB_ERR_CUDA -> RuntimeError — the reference's own exception types;
 *   - `math`: DNB_MATH_FP32 = SIMT fp32 FMA path (parity rel 1e-5 vs the float64
 *     reference); DNB_MATH_TF32 = tcgen05 tensor-core path, fp32 storage, TF32
 *     operands, fp32 accumulate in TMEM (parity rel 2e-2); DNB_MATH_BF16 = the same
 *     path with BF16 operands (still fp32 storage and fp32 accumulation, parity rel
 *     2e-2) in the shifted-view convolution kernels, TF32 elsewhere.  Chosen once per net.
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef DNPY_B200_H
#define DNPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNB_OK               0
#define DNB_ERR_INVALID     (-1)
#define DNB_ERR_CUDA        (-2)
#define DNB_ERR_UNSUPPORTED (-3)

#define DNB_MATH_FP32 0
#define DNB_MATH_TF32 1
#define DNB_MATH_BF16 2

/* MaxPool2D.backward batch routing (dnpy/layers/pooling.py:95) */
#define DNB_ROUTE_LITERAL    0   /* every sample's delta lands in sample 0, last sample wins */
#define DNB_ROUTE_PER_SAMPLE 1   /* intended per-sample scatter-add (== literal when B == 1) */

typedef void* dnb_stream_t;

/* Regulariser term lambda_l1*sign(p) + lambda_l2*2p (dnpy/regularizers.py:25-53); zeros = none. */
typedef struct { float l1; float l2; } dnb_reg_t;

/* 2-D window geometry shared by conv and pool (dnpy/utils.py:52-109 computes these on the host). */
typedef struct {
    int B, C, H, W;        /* input  (batch, channels, height, width)            */
    int F, OH, OW;         /* output (filters/channels, height, width)           */
    int kh, kw;            /* window                                             */
    int sy, sx;            /* strides                                            */
    int pt, pb, pl, pr;    /* zero padding: top, bottom, left, right             */
} dnb_win_t;

/* ---- library ---------------------------------------------------------------- */
int         dnb_version(void);
const char* dnb_last_error(void);
/* sm count, compute capability, whether the tcgen05 path is usable on this device */
int dnb_device_query(int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05);

/* ---- launch accounting (bench.py's gpu_launches / per-kernel roofline timing) ---------- */
/* kernels launched by this library since it was loaded */
long long dnb_launch_count(void);
/* between begin and end every kernel launch records a CUDA event on its launching stream; end
 * synchronises and writes {"kernel name": [launches, total_ms], ...} (JSON) into json_out. */
int dnb_prof_begin(dnb_stream_t stream);
int dnb_prof_end(char* json_out, size_t cap);

/* ---- Net.reset_grads (net.py:299-302) / generic fills ----------------------- */
int dnb_fill_f32(float* p, size_t n, float value, dnb_stream_t stream);
int dnb_fill_i32(int32_t* p, size_t n, int32_t value, dnb_stream_t stream);

/* ---- Dense (layers/core.py:116-135) ----------------------------------------- */
/* y[B,out] = x[B,in] @ w[in,out] + b[out] */
int dnb_dense_fwd(const float* x, const float* w, const float* b, float* y,
                  int B, int in, int out, int math, dnb_stream_t stream);
/* dx[B,in] = dy @ w^T (dx may be NULL);  gw += (x^T @ dy + reg_w'(w)) * inv_m;
 * gb += (sum_b dy + reg_b'(b)) * inv_m.  inv_m = 1/B single-GPU, 1/B_global when
 * the batch is sharded (regulariser then scaled by reg_scale = 1/world so the
 * all-reduced sum matches the single-process value). */
int dnb_dense_bwd(const float* x, const float* w, const float* b, const float* dy,
                  float* dx, float* gw, float* gb, int B, int in, int out, float inv_m,
                  dnb_reg_t reg_w, dnb_reg_t reg_b, float reg_scale, int math, dnb_stream_t stream);
/* dx[B,in] = dy @ w^T alone (core.py:121 without 129-135): the delta of an Input layer is dead on the training
 * path, so Net skips it in dnb_dense_bwd (dx = NULL) and runs this only when somebody reads `l_in.delta`. */
int dnb_dense_bwd_data(const float* w, const float* dy, float* dx, int B, int in, int out, int math,
                       dnb_stream_t stream);

/* ---- Conv2D / PointwiseConv2D (layers/convolution.py:193-273, 276-284) ------ */
size_t dnb_conv2d_ws_bytes(const dnb_win_t* g, int math);
/* y[B,F,OH,OW] = sum_{c,i,j} xpad * w[F,C,kh,kw] + (C*kh*kw) * b[F]   (bias quirk Q1) */
int dnb_conv2d_fwd(const float* x, const float* w, const float* b, float* y,
                   const dnb_win_t* g, void* ws, int math, dnb_stream_t stream);
/* dx[B,C,H,W] (cropped) ; gw += mean_b-sum_yx(dy * xpad) + OH*OW*reg_w'(w)*reg_scale ;
 * gb += inv_m * sum dy + OH*OW*reg_b'(b)*reg_scale  (Q2).  dx may be NULL. */
int dnb_conv2d_bwd(const float* x, const float* w, const float* b, const float* dy,
                   float* dx, float* gw, float* gb, const dnb_win_t* g, float inv_m,
                   dnb_reg_t reg_w, dnb_reg_t reg_b, float reg_scale,
                   void* ws, int math, dnb_stream_t stream);
/* dx alone (convolution.py:244-250,271-273 without 258-268): lazily materialised delta of an Input layer. */
int dnb_conv2d_bwd_data(const float* w, const float* dy, float* dx, const dnb_win_t* g, void* ws, int math,
                        dnb_stream_t stream);

/* ---- fused block: Conv2D(C=1, 3x3, stride 1, "none") -> MaxPool2D(3x3, stride 2, "none") -> LeakyRelu/Relu ------
 * (convolution.py:193-273 + pooling.py:43-99 + activations.py:17-22 in one pass: the F-channel convolution output and
 * its gradient never reach HBM).  `idx_gate` holds one byte per pooled element: bits 0-3 = MaxPool2D.output_idxs
 * (row-major offset of the first maximum in the window), bit 4 = the activation's gate (pool output > 0).
 * _supported: 1 when the two geometries are eligible.  _fwd writes act_out = LeakyRelu(pool(conv(x))) and idx_gate.
 * _bwd (PER_SAMPLE pool routing): gw += inv_m * sum, gb += inv_m * sum of the conv gradients, computed straight from the
 * activation's incoming delta (act_delta); no regulariser term.  dnb_idx_gate_unpack recovers the int32 argmax tensor. */
int dnb_conv_pool_act_supported(const dnb_win_t* conv, const dnb_win_t* pool);
int dnb_conv_pool_act_fwd(const float* x, const float* w, const float* b, float* act_out, uint8_t* idx_gate,
                          const dnb_win_t* conv, const dnb_win_t* pool, float alpha, dnb_stream_t stream);
int dnb_conv_pool_act_bwd(const float* x, const float* act_delta, const uint8_t* idx_gate, float* gw, float* gb,
                          const dnb_win_t* conv, const dnb_win_t* pool, float alpha, float inv_m, dnb_stream_t stream);
/* the same gradients under the reference's LITERAL pool routing (pooling.py:95: every sample's delta lands in sample 0's
 * window, the last sample per argmax offset wins): only image 0 carries a conv gradient.  ws: _ws_bytes (int32 table
 * last[F][OH*OW][9]). */
size_t dnb_conv_pool_act_bwd_literal_ws_bytes(const dnb_win_t* conv, const dnb_win_t* pool);
int dnb_conv_pool_act_bwd_literal(const float* x, const float* act_delta, const uint8_t* idx_gate, float* gw, float* gb,
                                  const dnb_win_t* conv, const dnb_win_t* pool, float alpha, float inv_m, void* ws,
                                  dnb_stream_t stream);
/* The forward of the same block on the TENSOR CORE (csrc/tc_conv_c1p.cu; math = DNB_MATH_BF16, MNIST geometry 28x28 ->
 * 12x12, F = 32 / 64 / 128): the 9-tap contraction runs as block-diagonal tcgen05 MMAs over 128 / F images at a time (BF16
 * operands, fp32 accumulation, the bias carried exactly as three BF16 pieces), the epilogue pools a filter's channel in
 * registers.  Same outputs and byte format as dnb_conv_pool_act_fwd; the backward calls above pair with either forward. */
int dnb_conv_pool_act_tc_supported(const dnb_win_t* conv, const dnb_win_t* pool, int math);
int dnb_conv_pool_act_fwd_tc(const float* x, const float* w, const float* b, float* act_out, uint8_t* idx_gate,
                             const dnb_win_t* conv, const dnb_win_t* pool, float alpha, int math, dnb_stream_t stream);
/* Conv2D -> MaxPool2D -> LeakyRelu/Relu behind a TENSOR-CORE convolution (the shifted-view tcgen05 kernels: stride 1, 32*k
 * input channels, one band per image — conv2 of the CNN; tf32 / bf16 math modes): the epilogue stages the output image in
 * shared memory and pools it there, so only act_out and idx_gate (one byte per pooled element: argmax offset | gate << 4) are
 * written — convolution.py:193-225 + pooling.py:43-99 + activations.py:17-19.  The backward pair is dnb_pool_act_bwd (it
 * rebuilds the convolution's dY) followed by dnb_conv2d_bwd. */
int dnb_conv2d_pool_act_supported(const dnb_win_t* conv, const dnb_win_t* pool, int math);
int dnb_conv2d_pool_act_fwd(const float* x, const float* w, const float* b, float* act_out, uint8_t* idx_gate,
                            const dnb_win_t* conv, const dnb_win_t* pool, float alpha, void* ws, int math,
                            dnb_stream_t stream);
/* MaxPool2D -> LeakyRelu/Relu where the activation is the pool's only consumer (pooling.py:43-99 + activations.py:17-22).
 * Forward: ONLY the activation's output and one byte per pooled element (argmax offset | gate << 4) are written, 5 bytes per
 * pooled element instead of 16 (pool output, int32 argmax, activation output).  Backward (per-sample routing):
 * LeakyRelu.backward + MaxPool2D.backward in one pass from the activation's delta and the byte tensor into the pool's
 * parent delta dx[B][C][H][W].  dnb_pool_act_supported: 1 when both calls take the geometry (<= 16 window positions,
 * square unpadded windows inside the plane). */
int dnb_pool_act_supported(const dnb_win_t* pool);
int dnb_maxpool2d_act_fwd(const float* x, float* act_out, uint8_t* idx_gate, const dnb_win_t* pool, float alpha,
                          dnb_stream_t stream);
int dnb_pool_act_bwd(const float* act_delta, const uint8_t* idx_gate, float* dx, const dnb_win_t* pool, float alpha,
                     dnb_stream_t stream);
int dnb_idx_gate_unpack(const uint8_t* idx_gate, int32_t* idx, size_t n, dnb_stream_t stream);
/* the inverse for a caller that ASSIGNS MaxPool2D.output_idxs (pooling.py:74-76 stores it as a plain attribute): the
 * argmax nibble of every byte is replaced, the gate bit kept */
int dnb_idx_gate_set(const int32_t* idx, uint8_t* idx_gate, size_t n, dnb_stream_t stream);

/* ---- DepthwiseConv2D (layers/convolution.py:323-388); w is [C,1,kh,kw] ------- */
int dnb_dwconv2d_fwd(const float* x, const float* w, const float* b, float* y,
                     const dnb_win_t* g, dnb_stream_t stream);
int dnb_dwconv2d_bwd(const float* x, const float* w, const float* b, const float* dy,
                     float* dx, float* gw, float* gb, const dnb_win_t* g, float inv_m,
                     dnb_reg_t reg_b, float reg_scale, dnb_stream_t stream);

/* ---- MaxPool2D / AvgPool2D (layers/pooling.py:43-99, 118-174) --------------- */
/* idx[B,C,OH,OW] int32: first max of the zero-padded row-major window (bit-exact vs np.argmax) */
int dnb_maxpool2d_fwd(const float* x, float* y, int32_t* idx, const dnb_win_t* g, dnb_stream_t stream);
size_t dnb_maxpool2d_bwd_ws_bytes(const dnb_win_t* g, int route);
int dnb_maxpool2d_bwd(const float* dy, const int32_t* idx, float* dx, const dnb_win_t* g,
                      int route, void* ws, dnb_stream_t stream);
/* The LITERAL route in two halves (same workspace): `prepare` needs only the forward's argmax offsets; `apply` gathers dY and
 * zero-fills samples 1..B-1.  (B == 1: use dnb_maxpool2d_bwd.)  Running `prepare` on a side stream right after the forward
 * pass was measured on B200: -10 us on the resident step, +55 us end to end (it disturbs the H2D copy overlap) — the host
 * mirror keeps the single call. */
int dnb_maxpool2d_bwd_prepare(const int32_t* idx, const dnb_win_t* g, void* ws, dnb_stream_t stream);
int dnb_maxpool2d_bwd_apply(const float* dy, float* dx, const dnb_win_t* g, const void* ws, dnb_stream_t stream);
int dnb_avgpool2d_fwd(const float* x, float* y, const dnb_win_t* g, dnb_stream_t stream);
int dnb_avgpool2d_bwd(const float* dy, float* dx, const dnb_win_t* g, dnb_stream_t stream);

/* ---- Activations (layers/activations.py) ------------------------------------ */
int dnb_leakyrelu_fwd(const float* x, float* y, size_t n, float alpha, dnb_stream_t stream);      /* :17-19 */
int dnb_leakyrelu_bwd(const float* x, const float* dy, float* dx, size_t n, float alpha, dnb_stream_t stream); /* :21-22 */
/* alpha[n_feat] broadcast over the batch; galpha += inv_m * sum_b dy*(x<=0)        :59-66 */
int dnb_prelu_fwd(const float* x, const float* alpha, float* y, int B, size_t n_feat, dnb_stream_t stream);
int dnb_prelu_bwd(const float* x, const float* alpha, const float* dy, float* dx, float* galpha,
                  int B, size_t n_feat, float inv_m, dnb_stream_t stream);
int dnb_sigmoid_fwd(const float* x, float* y, size_t n, dnb_stream_t stream);                      /* :77-78 */
int dnb_sigmoid_bwd(const float* y, const float* dy, float* dx, size_t n, dnb_stream_t stream);    /* :80-81 */
int dnb_tanh_fwd(const float* x, float* y, size_t n, dnb_stream_t stream);                         /* :92-95 */
int dnb_tanh_bwd(const float* y, const float* dy, float* dx, size_t n, dnb_stream_t stream);       /* :97-98 */
int dnb_softmax_fwd(const float* x, float* y, int B, int n, int stable, dnb_stream_t stream);      /* :116-124 */
int dnb_softmax_bwd(const float* y, const float* dy, float* dx, int B, int n, dnb_stream_t stream);/* :130-135 */
int dnb_logsoftmax_fwd(const float* x, float* y, int B, int n, int stable, dnb_stream_t stream);   /* :152-160 */

/* ---- Dropout / BatchNorm / Add (layers/regularization.py, merge.py) --------- */
/* y = x * gate (training; gate drawn on the host from NumPy's global stream for bit-exact
 * masks) — the same kernel serves Dropout.backward (regularization.py:17-25) */
int dnb_mul(const float* a, const float* b, float* y, size_t n, dnb_stream_t stream);
int dnb_scale(const float* x, float* y, size_t n, float s, dnb_stream_t stream);
/* y = sum of n_in inputs (merge.py:13-16 with np.add) */
int dnb_add_n(const float* const* xs, int n_in, float* y, size_t n, dnb_stream_t stream);
/* MPWLayer.forward for the other point-wise merges (merge.py:13-16, 29-62): y = ((x0 op x1) op x2) ... * post_scale.
 * Average.forward (merge.py:71-75) = DNB_MERGE_ADD with post_scale = 1/n_in.  Backward is a delta hand-over
 * (merge.py:18-20; Average: dnb_scale by 1/n_in per parent, merge.py:77-79). */
#define DNB_MERGE_ADD 0
#define DNB_MERGE_SUB 1
#define DNB_MERGE_MUL 2
#define DNB_MERGE_DIV 3
#define DNB_MERGE_POW 4
#define DNB_MERGE_MAX 5
#define DNB_MERGE_MIN 6
int dnb_merge_n(const float* const* xs, int n_in, float* y, size_t n, int op, float post_scale, dnb_stream_t stream);
/* BatchNorm over axis 0 of x[B, n_feat] (regularization.py:94-136).  training: computes
 * mu/var (biased) into save_mu/save_std (std = sqrt(var+eps)) and updates moving_* in place
 * with `momentum`; test: uses moving_*.  (Sharded batches: see dnb_batchnorm_stats / _fwd_synced below.) */
int dnb_batchnorm_fwd(const float* x, const float* gamma, const float* beta,
                      float* moving_mu, float* moving_var, float* save_mu, float* save_std,
                      float* y, int B, size_t n_feat, float momentum, float eps, int training,
                      dnb_stream_t stream);
/* literal backward (Q7): ggamma += sum_b dy*mu ; gbeta += sum_b dy ;
 * dx = (1/m) * std * (m*dy*gamma - sum_b(dy*gamma) - xn * sum_b(dy*gamma*xn)) */
int dnb_batchnorm_bwd(const float* x, const float* gamma, const float* save_mu, const float* save_std,
                      const float* dy, float* dx, float* ggamma, float* gbeta,
                      int B, size_t n_feat, dnb_stream_t stream);
/* The same layer with the batch sharded over ranks (one process per GPU): the rank-local halves around the two exchanges
 * that make N ranks compute the single-process statistics (regularization.py:98-99 and 148-152 over the GLOBAL batch).
 *   forward : dnb_batchnorm_stats -> every rank's (mean, M2) gathered into stats_all[world][2][n_feat] ->
 *             dnb_batchnorm_fwd_synced (Chan merge of equal-count partitions, normalise, moving statistics)
 *   backward: dnb_batchnorm_bwd_sums (local sum_b dy, sum_b dy*xn into sums[2][n_feat]; ggamma / gbeta are batch SUMS and
 *             are accumulated here, the gradient all-reduce adds the ranks' parts) -> all-reduce(sums) ->
 *             dnb_batchnorm_bwd_synced (dx with m = B_global) */
int dnb_batchnorm_stats(const float* x, float* mean, float* m2, int B, size_t n_feat, dnb_stream_t stream);
int dnb_batchnorm_fwd_synced(const float* x, const float* gamma, const float* beta, float* moving_mu, float* moving_var,
                             const float* stats_all, int world, float* save_mu, float* save_std, float* y,
                             int B, size_t n_feat, float momentum, float eps, dnb_stream_t stream);
int dnb_batchnorm_bwd_sums(const float* x, const float* save_mu, const float* save_std, const float* dy, float* sums,
                           float* ggamma, float* gbeta, int B, size_t n_feat, dnb_stream_t stream);
int dnb_batchnorm_bwd_synced(const float* x, const float* gamma, const float* save_mu, const float* save_std,
                             const float* dy, const float* sums, float* dx, int B, int B_global, size_t n_feat,
                             dnb_stream_t stream);

/* ---- Embedding (layers/core.py:54-65) --------------------------------------- */
int dnb_embedding_fwd(const int32_t* idx, const float* w, float* y, int B, int L, int V, int D, dnb_stream_t stream);
/* gw[v] += mean_b(dy)[l*] where (b*,l*) is the LAST row-major position with idx == v (Q15).
 * ws: (V + L*D) * 4 bytes. */
size_t dnb_embedding_bwd_ws_bytes(int L, int V, int D);
int dnb_embedding_bwd(const int32_t* idx, const float* dy, float* gw, int B, int L, int V, int D,
                      float inv_m, void* ws, dnb_stream_t stream);
/* The two halves of dnb_embedding_bwd, for a sharded batch: prepare leaves last[V] (int32: largest GLOBAL flat position
 * pos0 + b*L + l of every vocabulary row, -1 if unused) followed by the local part of mean_b(dy) (inv_m = 1/B_global) in
 * ws; the ranks all-reduce the positions (MAX) and the means (SUM); apply adds scale * mean[l*] to gw (scale = 1/world:
 * the gradient all-reduce then restores the exact single-process value). */
int dnb_embedding_bwd_prepare(const int32_t* idx, const float* dy, int B, int L, int V, int D, float inv_m, int pos0,
                              void* ws, dnb_stream_t stream);
int dnb_embedding_bwd_apply(const void* ws, float* gw, int L, int V, int D, float scale, dnb_stream_t stream);

/* ---- SimpleRNN (layers/recurrent.py:12-23, 153-266) ------------------------- */
/* states[B,T,H]: h_t = tanh(h_{t-1} waa^T + x_t wax^T + ba), h_0 = 0; y[B,H] = states[:, T-1] */
int dnb_rnn_fwd(const float* x, const float* waa, const float* wax, const float* ba,
                float* states, float* y, int B, int T, int D, int H, dnb_stream_t stream);
/* literal truncated BPTT (Q16): gradients are batch SUMS; dx[B,T,D] */
int dnb_rnn_bwd(const float* x, const float* states, const float* dy, const float* waa, const float* wax,
                float* dx, float* gwaa, float* gwax, float* gba,
                int B, int T, int D, int H, int trunc, dnb_stream_t stream);

/* The same gradients for the tensor-core math modes (csrc/rnn_tc.cu; H = 32, D <= 24, at most 7 recurrences per step): the
 * serial chain pre(t), dx_t stays a warp per sample (rnn_chain_kernel), the (trunc + 1) lag recurrences of ALL (sample, step)
 * rows run as [128 x 32] x [32 x 32] tcgen05.mma.kind::tf32 tiles and the weight gradients as one long reduction through the
 * lag sum W(u) = sum_k S_k(u + k)  (recurrent.py:228-247).  ws: dnb_rnn_bwd_tc_ws_bytes (pre for every row). */
int dnb_rnn_bwd_tc_supported(int T, int D, int H, int trunc, int math);
size_t dnb_rnn_bwd_tc_ws_bytes(int B, int T, int H);
int dnb_rnn_bwd_tc(const float* x, const float* states, const float* dy, const float* waa, const float* wax,
                   float* dx, float* gwaa, float* gwax, float* gba,
                   int B, int T, int D, int H, int trunc, void* ws, int math, dnb_stream_t stream);

/* ---- Losses (dnpy/losses.py) + the NaN/Inf guard of net.py:342-350 ---------- */
/* out[0] = loss, out[1] = flag bits: 1 NaN loss, 2 Inf loss, 4 NaN delta, 8 Inf delta.
 * Deltas are never divided by the batch (Q17).  loss_scale = 1/B (1/B_global when sharded).
 * ws: dnb_loss_ws_bytes(B) bytes, ZERO-FILLED by its owner before the first call; every call leaves it ready for the next
 * (it carries the ticket counter through which the last CTA of the single launch reduces the row terms in a fixed order). */
size_t dnb_loss_ws_bytes(int B);
int dnb_loss_ce(const float* p, const float* t, float* delta, float* out, int B, int n,
                int softmax_output, float loss_scale, void* ws, dnb_stream_t stream);   /* :82-94 */
int dnb_loss_bce(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream);                        /* :65-73 */
int dnb_loss_mse(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream);
/* RMSE / MAE / NLL / Hinge (losses.py:29-57, 97-133), same protocol: out[0] = loss, out[1] = NaN/Inf flags, delta filled.
 * RMSE, MAE and Hinge accept one output column (the reference takes float() of a per-column mean). */
int dnb_loss_rmse(const float* p, const float* t, float* delta, float* out, int B, int n,
                  float loss_scale, void* ws, dnb_stream_t stream);
int dnb_loss_mae(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream);
int dnb_loss_nll(const float* p, const float* t, float* delta, float* out, int B, int n,
                 float loss_scale, void* ws, dnb_stream_t stream);
int dnb_loss_hinge(const float* p, const float* t, float* delta, float* out, int B, int n,
                   float loss_scale, void* ws, dnb_stream_t stream);                        /* :22-26 */
/* metrics.py:44-80: out[0] = number of correct rows (host divides) */
int dnb_metric_categorical(const float* p, const float* t, float* out, int B, int n, dnb_stream_t stream);
int dnb_metric_binary(const float* p, const float* t, float* out, int B, int n, float thr, dnb_stream_t stream);

/* ---- Optimizers over the flat parameter arena (dnpy/optimizers.py) ----------- */
/* grad_scale multiplies g first (1 single-GPU; folds the data-parallel 1/N where needed). */
int dnb_opt_adam(float* p, const float* g, float* v, float* s, size_t n, float lr,
                 float beta1, float beta2, float eps, dnb_stream_t stream);               /* :106-126 */
int dnb_opt_rmsprop(float* p, const float* g, float* s, size_t n, float lr, float rho, float eps,
                    dnb_stream_t stream);                                                  /* :73-84 */
int dnb_opt_sgd(float* p, const float* g, float* v, size_t n, float lr, float momentum, int nesterov,
                dnb_stream_t stream);                                                      /* :31-55 */

/* ---- Tensor-core GEMM exposed for testing: C[M,N] = A · B, TF32 tcgen05, fp32 accumulate.
 * a_mmajor = 0: A stored [M,K] row-major, 1: stored [K,M]; b_nmajor = 0: B stored [N,K], 1: stored [K,N]. */
int dnb_gemm_tf32(const float* A, const float* Bm, float* C, int M, int N, int K,
                  int a_mmajor, int b_nmajor, dnb_stream_t stream);

/* ---- data-parallel gradient exchange over NVLink / NVSwitch peer memory (csrc/dp_peer.cu) ----------------------------
 * The shard of Net.fit's minibatch loop (net.py:176-195) across the GPUs of one box needs ONE exchange per step: the sum of
 * the flat gradient arenas.  dnb_ipc_alloc gives zero-filled device memory plus a 64-byte handle another process of the node
 * maps with dnb_ipc_open; dnb_dp_allreduce is one kernel per rank and step (flags and 16-byte loads through peer memory, ranks
 * summed in order 0..N-1 so the replicas stay bit-identical): bufs[r] / flags[r] are rank r's arena and flag block
 * (dnb_dp_flag_bytes, one block per arena) as mapped in the calling process, n a multiple of 4, scratch n local floats. */
size_t dnb_dp_flag_bytes(void);
int dnb_ipc_alloc(void** ptr, size_t bytes, void* handle_out64);
int dnb_ipc_open(const void* handle64, void** ptr);
int dnb_ipc_close(void* ptr);
int dnb_ipc_free(void* ptr);
int dnb_dp_allreduce(void* const* bufs, void* const* flags, float* scratch, size_t n, int rank, int world, dnb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DNPY_B200_H */
