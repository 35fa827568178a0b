#include "gemm.cuh"
namespace tnb {
int gemm_dmma_f64(const GemmProblem&, cudaStream_t) { return TNB200_ERR_UNSUPPORTED; }
}
extern "C" {
int32_t tnb200_svd(const tnb200_tensor_t*, const tnb200_tensor_t*, const tnb200_tensor_t*, const tnb200_tensor_t*, int32_t*, void*) { tnb::set_error("svd: not built yet"); return TNB200_ERR_UNSUPPORTED; }
int32_t tnb200_svd_truncation_count(const tnb200_tensor_t*, int64_t, int32_t, double, int32_t, int64_t*, void*) { tnb::set_error("svd: not built yet"); return TNB200_ERR_UNSUPPORTED; }
int32_t tnb200_qr(const tnb200_tensor_t*, const tnb200_tensor_t*, const tnb200_tensor_t*, int32_t, void*) { tnb::set_error("qr: not built yet"); return TNB200_ERR_UNSUPPORTED; }
int32_t tnb200_blocksparse_tensordot(const void*, const void*, void*, int32_t, int32_t, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, int64_t, int64_t, int32_t, void*) { tnb::set_error("blocksparse: not built yet"); return TNB200_ERR_UNSUPPORTED; }
}
