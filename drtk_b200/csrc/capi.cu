// capi.cu -- ABI version and error strings of libdrtk_b200.so (see include/drtk_b200.h).
#include "common.cuh"

extern "C" int drtk_b200_abi_version(void) { return DRTK_B200_ABI_VERSION; }

extern "C" const char* drtk_b200_error_string(int code) {
  if (code == 0) return "success";
  if (code == DRTK_B200_EINVAL) return "drtk_b200: invalid argument";
  if (code == DRTK_B200_EWORKSPACE) return "drtk_b200: workspace missing or too small";
  if (code == DRTK_B200_EUNSUPPORTED) return "drtk_b200: unsupported configuration";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "drtk_b200: unknown error";
}
