// Shared helpers for the sm_100a kernels of libdsurf_b200.so.
// The whole library is compiled with --fmad=false: parity with the reference (built with
// -O -ffloat-store, no FMA, src/Makefile:3-4) requires that a*b+c is never contracted.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace dsurf {

extern thread_local std::string g_last_error;
void set_error(const char *file, int line, const char *what);

#define DS_CUDA(call)                                                  \
  do {                                                                 \
    cudaError_t _e = (call);                                           \
    if (_e != cudaSuccess) {                                           \
      ::dsurf::set_error(__FILE__, __LINE__, cudaGetErrorString(_e));  \
      return DSURF_ERR_CUDA;                                           \
    }                                                                  \
  } while (0)

#define DS_CHECK(expr)                 \
  do {                                 \
    int _s = (expr);                   \
    if (_s != DSURF_OK) return _s;     \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

template <typename T>
struct DevBuf {  // minimal owning device buffer (grow-only)
  T *p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  cudaError_t reserve(size_t n, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return cudaSuccess;
    T *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(st);
    }
    if (p) cudaFree(p);
    p = q;
    cap = n;
    return cudaSuccess;
  }
};

int ensure_device();  // selects the device once; DSURF_ERR_NO_CUDA if none
int sm_count();

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace dsurf
