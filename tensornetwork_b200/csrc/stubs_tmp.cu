#include "gemm.cuh"
namespace tnb {
}
extern "C" {
int32_t tnb200_blocksparse_tensordot(const void*, const void*, void*, int32_t, int32_t, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, int64_t, int64_t, int32_t, void*) { tnb::set_error("blocksparse: not built yet"); return TNB200_ERR_UNSUPPORTED; }
}
