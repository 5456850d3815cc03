// capi.cu -- status / error plumbing of the C ABI (include/gccb200.h).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include <mutex>

namespace gccb {
#ifndef GCCB_EMU
unsigned long long g_launch_count = 0;   // kernels enqueued by this library (bench.py: gpu_launches)
#endif
static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return GCCB_ERR_CUDA;
  }
  return GCCB_OK;
}

#ifndef GCCB_EMU
// a library side stream with the caller stream's scheduling priority (a default-priority weight-gradient
// stream starves behind queued eigensolver CTAs), or the lowest priority on request
static cudaStream_t create_stream_like(cudaStream_t like, bool lowest_priority) {
  int prio = 0;
  if (lowest_priority || cudaStreamGetPriority(like, &prio) != cudaSuccess) { cudaGetLastError(); prio = 0; }
  cudaStream_t s = nullptr;
  cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio);
  return s;
}

StreamKit* stream_kit(cudaStream_t caller, int family, bool lowest_priority) {
  static StreamKit kits[16];
  static int nkits = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);                                    // streams belong to a device: the legacy stream (0) of two
  for (int i = 0; i < nkits; ++i)                         // devices must not share side streams
    if (kits[i].key == caller && kits[i].family == family && kits[i].dev == dev) return &kits[i];
  StreamKit* k;
  if (nkits < 16) {
    k = &kits[nkits++];
    for (int i = 0; i < 5; ++i) k->side[i] = create_stream_like(caller, lowest_priority);
    for (int i = 0; i < 24; ++i) cudaEventCreateWithFlags(&k->ev[i], cudaEventDisableTiming);
  } else {
    k = &kits[15];                                        // more caller streams than kits: share the last one
  }
  k->key = caller;
  k->family = family;
  k->dev = dev;
  return k;
}
#endif
}  // namespace gccb

extern "C" int gccb_version(void) { return GCCB_VERSION; }

extern "C" unsigned long long gccb_launch_count(void) {
#ifndef GCCB_EMU
  return gccb::g_launch_count;
#else
  return 0;
#endif
}

extern "C" const char* gccb_last_error(void) { return gccb::g_err; }

extern "C" int gccb_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    gccb::set_last_error("gccb_arch: no CUDA device");
    cudaGetLastError();
    return GCCB_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}
