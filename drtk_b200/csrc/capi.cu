// capi.cu -- ABI version and error strings of libdrtk_b200.so (see include/drtk_b200.h).
#include "common.cuh"

#include <atomic>

namespace drtk {
int num_sms() {
  static std::atomic<int> cache[64];  // 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
}  // namespace drtk

extern "C" int drtk_b200_abi_version(void) { return DRTK_B200_ABI_VERSION; }

extern "C" const char* drtk_b200_error_string(int code) {
  if (code == 0) return "success";
  if (code == DRTK_B200_EINVAL) return "drtk_b200: invalid argument";
  if (code == DRTK_B200_EWORKSPACE) return "drtk_b200: workspace missing or too small";
  if (code == DRTK_B200_EUNSUPPORTED) return "drtk_b200: unsupported configuration";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "drtk_b200: unknown error";
}
