// placeholder — replaced by the tcgen05 implementation
#include "common.cuh"
#include "tc_conv.cuh"
namespace dnb {
size_t tc_conv_ws_bytes(const dnb_win_t*) { return 0; }
bool tc_conv_supported(const dnb_win_t*, int) { return false; }
int tc_conv_fwd(const float*, const float*, const float*, float*, const dnb_win_t*, void*, cudaStream_t) { return DNB_ERR_UNSUPPORTED; }
int tc_conv_dgrad(const float*, const float*, float*, const dnb_win_t*, void*, cudaStream_t) { return DNB_ERR_UNSUPPORTED; }
int tc_conv_wgrad(const float*, const float*, float*, float*, const dnb_win_t*, float, void*, cudaStream_t) { return DNB_ERR_UNSUPPORTED; }
}
