// Interface of the tcgen05/TMEM implicit-GEMM convolution kernels (tc_conv.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/dnpy_b200.h"

namespace dnb {

enum { TC_CONV_FWD = 0, TC_CONV_DGRAD = 1, TC_CONV_WGRAD = 2 };

size_t tc_conv_ws_bytes(const dnb_win_t* g);
bool tc_conv_supported(const dnb_win_t* g, int kind);
int tc_conv_fwd(const float* x, const float* w, const float* b, float* y, const dnb_win_t* g, void* ws, bool bf16, cudaStream_t st);
int tc_conv_dgrad(const float* dy, const float* w, float* dx, const dnb_win_t* g, void* ws, bool bf16, cudaStream_t st);
int tc_conv_wgrad(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, void* ws,
                  bool bf16, cudaStream_t st);

// tc_conv_sv.cu: shifted-view kernel (stride 1, 32-channel multiples) used by tc_conv_fwd / tc_conv_dgrad
// bf16: operands staged as BF16 (the filters `wprep` are then a [n_pad x Kp] BF16 matrix); only the column-folded 3x3 variant
bool tc_conv_sv_supported(const dnb_win_t* g, bool dgrad);
bool tc_conv_sv_bf16_supported(const dnb_win_t* g, bool dgrad);
int tc_conv_sv_run(const char* name, const float* src, const void* wprep, int n_pad, int Kp, const float* bias, float bias_k,
                   float* dst, const dnb_win_t* g, bool dgrad, bool bf16, cudaStream_t st);

// tc_conv_svp.cu: 3x3 'same' stride-1 convolution (32 / 64 input channels, <= 64 filters, BF16 operands) with the
// MaxPool2D -> LeakyRelu pair behind it in the epilogue.  Operand roles are swapped against tc_conv_sv.cu (filters = M,
// the pixels of the staged image = N): a TMEM lane holds one filter's whole image, so pooling is register work.
bool tc_conv_svp_supported(const dnb_win_t* g, const dnb_win_t* pool);
int tc_conv_svp_run(const float* x, const void* wprep, int n_pad, int Kp, const float* bias, float bias_k, float* act,
                    unsigned char* idx_gate, float alpha, const dnb_win_t* g, const dnb_win_t* pool, cudaStream_t st);
int tc_conv_fwd_pool(const float* x, const float* w, const float* b, float* act, unsigned char* idx_gate, float alpha,
                     const dnb_win_t* g, const dnb_win_t* pool, void* ws, cudaStream_t st);

// tc_conv_wg.cu: shifted-view weight gradient (3x3 stride 1, 32 input channels, small images)
bool tc_conv_wgrad_sv_supported(const dnb_win_t* g);
int tc_conv_wgrad_sv(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, bool bf16, cudaStream_t st);

// tc_conv_thin.cu: tensor-core weight gradient of the single-channel 3x3 stride-1 convolution (dY straight from TMA)
bool tc_pointwise_wgrad_supported(const dnb_win_t* g);
int tc_pointwise_wgrad(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, cudaStream_t st);
bool tc_thin_wgrad_supported(const dnb_win_t* g);
int tc_thin_wgrad(const float* x, const float* dy, float* gw, float* gb, const dnb_win_t* g, float inv_m, cudaStream_t st);

}  // namespace dnb
