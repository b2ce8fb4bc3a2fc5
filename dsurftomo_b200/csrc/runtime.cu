// Device selection, error reporting and the gfortran-convention drop-in wrappers that need
// no kernel code of their own (they forward to the dsurf_* entry points).
#include <cstring>
#include <mutex>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "build_info.h"

namespace dsurf {
thread_local std::string g_last_error;
static int g_device = -1;
static int g_requested_device = -1;  // set by dsurf_set_device (guarded by g_mu through ensure_device)
static int g_sm_count = 148;
static std::mutex g_mu;

void set_error(const char *file, int line, const char *what) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s:%d: %s", file, line, what);
  g_last_error = buf;
}

int ensure_device() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device >= 0) {
    cudaSetDevice(g_device);
    return DSURF_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_last_error = "no CUDA device: libdsurf_b200 has no CPU fallback (needs an sm_100 GPU)";
    return DSURF_ERR_NO_CUDA;
  }
  int dev = 0;
  if (g_requested_device >= 0)
    dev = g_requested_device % n;  // dsurf_set_device
  else if (const char *lr = getenv("LOCAL_RANK"))
    dev = atoi(lr) % n;  // torchrun's default; only read, never written
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess || p.major < 10) {
    g_last_error = "device is not sm_100-class: libdsurf_b200 is built for sm_100a only";
    return DSURF_ERR_NO_CUDA;
  }
  cudaSetDevice(dev);
  g_device = dev;
  g_sm_count = p.multiProcessorCount;
  return DSURF_OK;
}
int sm_count() { return g_sm_count; }
}  // namespace dsurf

using namespace dsurf;

extern "C" const char *dsurf_last_error(void) { return g_last_error.c_str(); }
extern "C" int dsurf_set_device(int device) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_device = -1;
  }
  g_requested_device = device;
  return ensure_device();
}
extern "C" const char *dsurf_build_info(void) {
  return "libdsurf_b200: sm_100a, --fmad=false, fp32 eikonal/rays/LSMR, fp64 dispersion search; src " DSURF_SRC_HASH;
}

static void fatal(int rc, const char *who) {
  // the reference prints to unit 6 and STOPs
  if (rc == DSURF_ERR_SOURCE_OUTSIDE) {
    printf(" Source lies outside bounds of model (lat,long)\n TERMINATING PROGRAM!!!\n");
  } else if (rc == DSURF_ERR_RECEIVER_OUTSIDE) {
    printf(" Receiver lies outside model (lat,long)\n TERMINATING PROGRAM!!!!\n");
  } else {
    fprintf(stderr, "%s: libdsurf_b200 error %d: %s\n", who, rc, dsurf_last_error());
  }
  fflush(stdout);
  exit(1);
}

extern "C" void __lsmrmodule_MOD_lsmr(const int *m, const int *n, const int *leniw, const int *lenrw,
                                      const int *iw, const float *rw, const float *b,
                                      const float *damp, const float *atol, const float *btol,
                                      const float *conlim, const int *itnlim, const int *localSize,
                                      const int *nout, float *x, int *istop, int *itn, float *normA,
                                      float *condA, float *normr, float *normAr, float *normx) {
  (void)nout;  // undefined in the reference caller (main.f90:107): treated as "silent"
  int rc = dsurf_lsmr(*m, *n, *leniw, *lenrw, iw, rw, b, *damp, *atol, *btol, *conlim, *itnlim,
                      *localSize, x, istop, itn, normA, condA, normr, normAr, normx);
  if (rc != DSURF_OK) fatal(rc, "LSMR");
}

extern "C" void aprod_(const int *mode, const int *m, const int *n, float *x, float *y,
                       const int *leniw, const int *lenrw, const int *iw, const float *rw) {
  int rc = dsurf_aprod(*mode, *m, *n, x, y, *leniw, *lenrw, iw, rw);
  if (rc != DSURF_OK) fatal(rc, "aprod");
}

extern "C" void dsurf_fatal_(const int *rc) { fatal(*rc, "dsurf"); }
