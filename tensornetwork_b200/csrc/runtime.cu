// runtime.cu — error state, device info, stream-ordered scratch allocator.
#include "common.cuh"
#include <mutex>

namespace tnb {
static thread_local char t_err[1024] = "";
static thread_local char t_kernel[64] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
void set_kernel_name(const char* name) {
  strncpy(t_kernel, name, sizeof(t_kernel) - 1);
  t_kernel[sizeof(t_kernel) - 1] = 0;
}

static std::once_flag g_pool_once;
static void init_pool() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
}
int ws_alloc(void** p, size_t bytes, cudaStream_t st) {
  std::call_once(g_pool_once, init_pool);
  TNB_CHECK_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, st));
  return 0;
}
int ws_free(void* p, cudaStream_t st) {
  if (p) TNB_CHECK_CUDA(cudaFreeAsync(p, st));
  return 0;
}
}  // namespace tnb

extern "C" {
const char* tnb200_last_error(void) { return tnb::t_err; }
const char* tnb200_last_kernel(void) { return tnb::t_kernel; }
int32_t tnb200_abi_version(void) { return TNB200_ABI_VERSION; }
int64_t tnb200_launch_count(void) { return tnb::g_launches.load(); }
int32_t tnb200_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                           int64_t* total_mem) {
  int dev = 0;
  TNB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  TNB_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
  return 0;
}
}
