// placeholder — replaced by the tcgen05 implementation
#include "common.cuh"
#include "tc_gemm.cuh"
namespace dnb {
bool tc_gemm_supported(int, int, int, bool, bool) { return false; }
int tc_gemm(const float*, const float*, float*, int, int, int, int, const float*, const float*, float, float, float, cudaStream_t) {
    return set_err(DNB_ERR_UNSUPPORTED, "tc_gemm: not built");
}
}
extern "C" int dnb_gemm_tf32_nt(const float*, const float*, float*, int, int, int, dnb_stream_t) {
    return dnb::set_err(DNB_ERR_UNSUPPORTED, "gemm_tf32_nt: not built");
}
