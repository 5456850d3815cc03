// partition.cu -- SM partitioning with CUDA green contexts (driver API, CUDA >= 12.4).
//
// The pretraining step has two very different halves running concurrently (engine.py): a few
// thousand long-lived, shared-memory-heavy sampler / eigensolver CTAs and ~100 short dependent
// training kernels.  Sharing SMs, the short kernels queue behind the long CTAs' shared memory and
// thread slots.  A green context gives each half its own SMs: every stream created here launches
// only on its group.  Driver entry points are resolved through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda and still loads on a CPU-only box.
#include "common.cuh"

#ifndef GCCB_EMU
#include <cuda.h>
#include <mutex>

namespace gccb {
void set_last_error(const char* fmt, ...);

namespace {
struct Drv {
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*,
                                        unsigned int, unsigned int);
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);
  CUresult (*GreenCtxGetDevResource)(CUgreenCtx, CUdevResource*, CUdevResourceType);
  CUresult (*StreamGetGreenCtx)(CUstream, CUgreenCtx*);
  bool ok;
};

template <class F>
bool load(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

Drv* drv() {
  static Drv d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.ok = load("cuDeviceGet", &d.DeviceGet) && load("cuDeviceGetDevResource", &d.DeviceGetDevResource) &&
           load("cuDevSmResourceSplitByCount", &d.DevSmResourceSplitByCount) &&
           load("cuDevResourceGenerateDesc", &d.DevResourceGenerateDesc) &&
           load("cuGreenCtxCreate", &d.GreenCtxCreate) && load("cuGreenCtxStreamCreate", &d.GreenCtxStreamCreate) &&
           load("cuGreenCtxGetDevResource", &d.GreenCtxGetDevResource) &&
           load("cuStreamGetGreenCtx", &d.StreamGetGreenCtx);
  });
  return &d;
}
}  // namespace

// A side stream that launches where and how `like` launches: same priority, and inside the same
// green context if `like` belongs to one.
cudaStream_t create_stream_like(cudaStream_t like, bool lowest_priority) {
  int prio = 0;
  if (lowest_priority || cudaStreamGetPriority(like, &prio) != cudaSuccess) { cudaGetLastError(); prio = 0; }
  Drv* d = drv();
  if (d->ok && like) {
    CUgreenCtx g = nullptr;
    if (d->StreamGetGreenCtx((CUstream)like, &g) == CUDA_SUCCESS && g) {
      CUstream s = nullptr;
      if (d->GreenCtxStreamCreate(&s, g, CU_STREAM_NON_BLOCKING, prio) == CUDA_SUCCESS) return (cudaStream_t)s;
    }
  }
  cudaStream_t s = nullptr;
  cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio);   // same scheduling priority as the caller's stream
  return s;
}
}  // namespace gccb

struct gccb_partition {
  CUgreenCtx ctx[2];
  int sms[2];
};

extern "C" int gccb_partition_create(int32_t device, int32_t first_sms, gccb_partition_t** out) {
  using namespace gccb;
  if (!out || first_sms <= 0) {
    set_last_error("gccb_partition_create: bad argument");
    return GCCB_ERR_BADARG;
  }
  Drv* d = drv();
  if (!d->ok) {
    set_last_error("gccb_partition_create: green contexts need a CUDA 12.4+ driver");
    return GCCB_ERR_CUDA;
  }
  cudaSetDevice(device);
  cudaFree(nullptr);                                     // make sure the primary context exists
  CUdevice dev;
  CUdevResource all, grp, rem;
  unsigned int ngroups = 1;
  CUresult r = d->DeviceGet(&dev, device);
  if (r == CUDA_SUCCESS) r = d->DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
  if (r == CUDA_SUCCESS && (unsigned)first_sms >= all.sm.smCount) {
    set_last_error("gccb_partition_create: %d of %u SMs leaves nothing for the second group", first_sms, all.sm.smCount);
    return GCCB_ERR_BADARG;
  }
  if (r == CUDA_SUCCESS) r = d->DevSmResourceSplitByCount(&grp, &ngroups, &all, &rem, 0, (unsigned)first_sms);
  if (r != CUDA_SUCCESS || ngroups != 1) {
    set_last_error("gccb_partition_create: SM split failed (driver error %d)", (int)r);
    return GCCB_ERR_CUDA;
  }
  gccb_partition* p = new gccb_partition();
  CUdevResource parts[2] = {grp, rem};
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc;
    r = d->DevResourceGenerateDesc(&desc, &parts[i], 1);
    if (r == CUDA_SUCCESS) r = d->GreenCtxCreate(&p->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM);
    if (r != CUDA_SUCCESS) {
      set_last_error("gccb_partition_create: green context %d failed (driver error %d)", i, (int)r);
      delete p;
      return GCCB_ERR_CUDA;
    }
    p->sms[i] = (int)parts[i].sm.smCount;
  }
  *out = p;
  return GCCB_OK;
}

extern "C" int32_t gccb_partition_sm_count(const gccb_partition_t* p, int32_t which) {
  return p && (which == 0 || which == 1) ? p->sms[which] : 0;
}

extern "C" int gccb_partition_stream(gccb_partition_t* p, int32_t which, int32_t priority, gccb_stream_t* out) {
  using namespace gccb;
  if (!p || !out || which < 0 || which > 1) {
    set_last_error("gccb_partition_stream: bad argument");
    return GCCB_ERR_BADARG;
  }
  CUstream s = nullptr;
  CUresult r = drv()->GreenCtxStreamCreate(&s, p->ctx[which], CU_STREAM_NON_BLOCKING, priority);
  if (r != CUDA_SUCCESS) {
    set_last_error("gccb_partition_stream: driver error %d", (int)r);
    return GCCB_ERR_CUDA;
  }
  *out = (gccb_stream_t)s;
  return GCCB_OK;
}

#else   // GCCB_EMU: no streams under the CPU emulator

struct gccb_partition { int unused; };
extern "C" int gccb_partition_create(int32_t, int32_t, gccb_partition_t**) { return GCCB_ERR_CUDA; }
extern "C" int32_t gccb_partition_sm_count(const gccb_partition_t*, int32_t) { return 0; }
extern "C" int gccb_partition_stream(gccb_partition_t*, int32_t, int32_t, gccb_stream_t*) { return GCCB_ERR_CUDA; }
#endif
