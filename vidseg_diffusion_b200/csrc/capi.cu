// Library-wide state of libvidseg_b200.so: error string, launch counter, version probes.
#include "common.cuh"

namespace vidseg {
thread_local char g_last_error[512] = "";
std::atomic<long long> g_launch_count{0};
}  // namespace vidseg

VS_API const char* vidseg_last_error(void) { return vidseg::g_last_error; }
VS_API int vidseg_abi_version(void) { return 1; }
VS_API long long vidseg_launch_count(void) { return vidseg::g_launch_count.load(); }
VS_API int vidseg_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
