// capi.cu - version / error / device entry points of the C ABI (include/mal_b200.h).
#include "mal_common.cuh"

extern "C" int mal_abi_version(void) { return MAL_ABI_VERSION; }

extern "C" const char* mal_last_error(void) { return mal::err_buf(); }

extern "C" int mal_check_device(int device) {
#ifdef MAL_EMU
  (void)device;
  return MAL_OK;
#else
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return mal::fail(MAL_ERR_LAUNCH, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
  if (prop.major != 10)
    return mal::fail(MAL_ERR_ARCH, "device %d is sm_%d%d; libmal_b200 is built for sm_100a only", device, prop.major,
                     prop.minor);
  return MAL_OK;
#endif
}
