// Library-level entry points: ABI version, thread-local error string, device capability check.
#include <atomic>

#include "common.cuh"

namespace b200 {

unsigned long long launch_count();

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(); }

int check_launch(const char* what) {
  g_launches.fetch_add(1);  // every kernel launch of this library is followed by exactly one check_launch
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace b200

extern "C" {

int b200unet_abi_version(void) { return B200UNET_ABI_VERSION; }

const char* b200unet_last_error(void) { return b200::err_buf(); }

int b200unet_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

unsigned long long b200unet_launch_count(void) { return b200::launch_count(); }

int b200unet_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
}
