// Error plumbing + version of liblafs_b200.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {
static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LAFS_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return LAFS_ERR_CUDA;
  }
  return LAFS_OK;
}
int bind_device_of(const void* device_ptr) {
  if (device_ptr == nullptr) return LAFS_OK;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, device_ptr) != cudaSuccess) {
    cudaGetLastError();
    return LAFS_OK;   // not a CUDA pointer we can classify: keep the thread's current device
  }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return LAFS_OK;
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != a.device) {
    if (cudaSetDevice(a.device) != cudaSuccess) {
      set_last_error("cudaSetDevice(%d): %s", a.device, cudaGetErrorString(cudaGetLastError()));
      return LAFS_ERR_CUDA;
    }
  }
  // force the primary context current on this thread once (driver entry points such as the TMA
  // descriptor encoder need it); never again, so that entry points stay legal inside a CUDA-graph
  // stream capture (cudaFree invalidates a capture)
  static thread_local int bound_device = -1;
  if (bound_device != a.device) {
    cudaFree(0);
    bound_device = a.device;
  }
  return LAFS_OK;
}
}  // namespace lafs

extern "C" int lafs_version(void) { return 100; }

extern "C" const char* lafs_last_error_string(void) { return lafs::g_err; }

extern "C" int lafs_device_ok(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    lafs::set_last_error("lafs_device_ok: %s", cudaGetErrorString(cudaGetLastError()));
    return LAFS_ERR_CUDA;
  }
  return p.major == 10 ? 1 : 0;
}
