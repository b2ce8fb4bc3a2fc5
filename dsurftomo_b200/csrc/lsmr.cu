// K7 -- LSMR (Fong & Saunders) with local reorthogonalisation on one B200, replacing
// LSMRmodule::LSMR (src/lsmrModule.f90:36-750), aprod (src/aprod.f90:7-60) and the REAL*4
// dnrm2/dscal of src/lsmrblas.f90:247-359.
//
// Data layout in HBM: the COO triplets the reference passes in iw/rw are turned once per call
// into CSR (for u += A v) and CSC (for v += A' u) so that both products stream (col,val) /
// (row,val) pairs with fully coalesced 128-byte warp loads and gather the (L2-resident) dense
// vector.  Storage is fp32 like the reference (real(dp) is REAL*4, lsmrDataModule.f90:21);
// every reduction (row sums, norms, dot products) accumulates in fp64 in a fixed order, so the
// result is deterministic and at least as accurate as the reference's serial fp32 sums.
// Scalar recurrences (plane rotations, norm/cond estimates, stopping rules) are evaluated in
// fp32 in the reference's operation order by single-thread kernels, so no host round trip is
// needed inside an iteration except reading the stop flag.
//
// Bound: HBM bandwidth.  Algorithmic bytes per iteration:
//   16*nnz + 8*(m+1) + 12*m + 80*n.
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>
#include "../../include/dsurftomo_b200.h"
#include <cooperative_groups.h>
#include "common.cuh"
#include "lsmr.cuh"

namespace dsurf {

// ---------------------------------------------------------------------------------------------
// sparse products: one warp per row of the compressed structure (CSR row or CSC column)
//   y[r] = float( double(fl(s * y[r])) + sum_k double(val[k]) * double(x[idx[k]]) )
// (the reference first scales y by s with dscal, then accumulates, lsmrModule.f90:484-486,495-497)
// Optionally accumulates sum(y^2) per block into partial[blockIdx.x] (fixed order).
// ---------------------------------------------------------------------------------------------
constexpr int kSpmvWarps = 8;
constexpr int kShortRow = 32;  // rows/columns with fewer entries are handled one per thread

// Long rows: one warp per row, 16-byte vector loads of (idx,val) after peeling to a 16-byte
// boundary, two vectors (8 entries) per lane in flight.  rowlist == nullptr -> rows 0..nrows-1.
__global__ void __launch_bounds__(kSpmvWarps * 32)
k_spmv_warp(const long long *__restrict__ ptr, const int *__restrict__ idx,
            const float *__restrict__ val, const float *__restrict__ x, float *__restrict__ y,
            const float *__restrict__ scale_ptr, float scale_sign, const int *__restrict__ rowlist,
            int nrows, double *__restrict__ partial, const int *__restrict__ stop) {
  __shared__ double sh[kSpmvWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wi = blockIdx.x * kSpmvWarps + w;
  double sq = 0.0;
  if (stop == nullptr || *stop == 0) {
    if (wi < nrows) {
      const int r = rowlist ? rowlist[wi] : wi;
      const long long b = ptr[r], e = ptr[r + 1];
      long long k0 = (b + 3) & ~3LL;
      if (k0 > e) k0 = e;
      double acc0 = 0.0, acc1 = 0.0;
      if (b + lane < k0) acc0 += (double)val[b + lane] * (double)__ldg(x + idx[b + lane]);
      const long long nv = (e - k0) >> 2;
      const int4 *idx4 = reinterpret_cast<const int4 *>(idx + k0);
      const float4 *val4 = reinterpret_cast<const float4 *>(val + k0);
      long long v = lane;
      for (; v + 32 < nv; v += 64) {
        const int4 ia = idx4[v], ib = idx4[v + 32];
        const float4 va = val4[v], vb = val4[v + 32];
        acc0 += (double)va.x * (double)__ldg(x + ia.x);
        acc1 += (double)va.y * (double)__ldg(x + ia.y);
        acc0 += (double)va.z * (double)__ldg(x + ia.z);
        acc1 += (double)va.w * (double)__ldg(x + ia.w);
        acc0 += (double)vb.x * (double)__ldg(x + ib.x);
        acc1 += (double)vb.y * (double)__ldg(x + ib.y);
        acc0 += (double)vb.z * (double)__ldg(x + ib.z);
        acc1 += (double)vb.w * (double)__ldg(x + ib.w);
      }
      for (; v < nv; v += 32) {
        const int4 ia = idx4[v];
        const float4 va = val4[v];
        acc0 += (double)va.x * (double)__ldg(x + ia.x);
        acc1 += (double)va.y * (double)__ldg(x + ia.y);
        acc0 += (double)va.z * (double)__ldg(x + ia.z);
        acc1 += (double)va.w * (double)__ldg(x + ia.w);
      }
      const long long kt = k0 + (nv << 2) + lane;
      if (kt < e) acc1 += (double)val[kt] * (double)__ldg(x + idx[kt]);
      const double acc = warp_sum(acc0 + acc1);
      if (lane == 0) {
        const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
        const float y0 = scale_ptr ? s * y[r] : 0.0f;
        const float yn = (float)((double)y0 + acc);
        y[r] = yn;
        sq = (double)yn * (double)yn;
      }
    }
  }
  if (partial) {
    if (lane == 0) sh[w] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < kSpmvWarps; i++) t += sh[i];
      partial[blockIdx.x] = t;
    }
  }
}

// Short rows (e.g. the 1- and 7-entry smoothing rows): one thread per row.
__global__ void __launch_bounds__(256)
k_spmv_short(const long long *__restrict__ ptr, const int *__restrict__ idx, const float *__restrict__ val,
             const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ scale_ptr,
             float scale_sign, const int *__restrict__ rowlist, int nrows, double *__restrict__ partial,
             const int *__restrict__ stop) {
  __shared__ double sh[256];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double sq = 0.0;
  if ((stop == nullptr || *stop == 0) && t < nrows) {
    const int r = rowlist[t];
    const long long b = ptr[r], e = ptr[r + 1];
    double acc = 0.0;
    for (long long k = b; k < e; k++) acc += (double)val[k] * (double)__ldg(x + idx[k]);
    const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
    const float y0 = scale_ptr ? s * y[r] : 0.0f;
    const float yn = (float)((double)y0 + acc);
    y[r] = yn;
    sq = (double)yn * (double)yn;
  }
  if (partial) {
    sh[threadIdx.x] = sq;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
  }
}

// ---------------------------------------------------------------------------------------------
// Depth-blocked products.  A sensitivity row touches the same B-spline vertices at every depth
// (row = S(k,vertex) * fdm(vertex), CalSurfG.f90:1396-1399), so the nz-1 depth columns of one
// vertex are stored as one block: 1 index + 8 values (36 B for up to 8 non-zeros instead of
// 8 B each) and the n-vectors are kept internally in the permuted order [vertex][depth] so that a
// block reads one aligned 32-byte sector of x.  Blocks are built from the caller's COO on the
// device (build_blocked); entries the reference dropped by its |row| > 1e-4 test are zeros.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dot8(const float4 a0, const float4 a1, const float4 b0, const float4 b1) {
  double t = (double)a0.x * (double)b0.x;
  t += (double)a0.y * (double)b0.y;
  t += (double)a0.z * (double)b0.z;
  t += (double)a0.w * (double)b0.w;
  t += (double)a1.x * (double)b1.x;
  t += (double)a1.y * (double)b1.y;
  t += (double)a1.z * (double)b1.z;
  t += (double)a1.w * (double)b1.w;
  return t;
}

// y[r] = s*y[r] + sum_blocks val[b][0..7] . x[pos[b]*8 .. +7]      (rows of A; warp per row)
__global__ void __launch_bounds__(kSpmvWarps * 32)
k_bspmv_rows(const long long *__restrict__ ptr, const int *__restrict__ pos, const float4 *__restrict__ val,
             const float4 *__restrict__ x4, float *__restrict__ y, const float *__restrict__ scale_ptr,
             float scale_sign, const int *__restrict__ rowlist, int nrows, int per_thread,
             double *__restrict__ partial, const int *__restrict__ stop) {
  __shared__ double sh[kSpmvWarps * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double sq = 0.0;
  if (stop == nullptr || *stop == 0) {
    if (!per_thread) {
      const int wi = blockIdx.x * kSpmvWarps + w;
      if (wi < nrows) {
        const int r = rowlist[wi];
        const long long b0 = ptr[r], b1 = ptr[r + 1];
        double acc0 = 0.0, acc1 = 0.0;
        long long b = b0 + lane;
        for (; b + 32 < b1; b += 64) {
          const int pa = pos[b], pb = pos[b + 32];
          const float4 va0 = val[2 * b], va1 = val[2 * b + 1];
          const float4 vb0 = val[2 * (b + 32)], vb1 = val[2 * (b + 32) + 1];
          acc0 += dot8(va0, va1, __ldg(x4 + 2 * (size_t)pa), __ldg(x4 + 2 * (size_t)pa + 1));
          acc1 += dot8(vb0, vb1, __ldg(x4 + 2 * (size_t)pb), __ldg(x4 + 2 * (size_t)pb + 1));
        }
        for (; b < b1; b += 32) {
          const int pa = pos[b];
          acc0 += dot8(val[2 * b], val[2 * b + 1], __ldg(x4 + 2 * (size_t)pa), __ldg(x4 + 2 * (size_t)pa + 1));
        }
        const double acc = warp_sum(acc0 + acc1);
        if (lane == 0) {
          const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
          const float y0 = scale_ptr ? s * y[r] : 0.0f;
          const float yn = (float)((double)y0 + acc);
          y[r] = yn;
          sq = (double)yn * (double)yn;
        }
      }
    } else {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t < nrows) {
        const int r = rowlist[t];
        double acc = 0.0;
        for (long long b = ptr[r]; b < ptr[r + 1]; b++) {
          const int pa = pos[b];
          acc += dot8(val[2 * b], val[2 * b + 1], __ldg(x4 + 2 * (size_t)pa), __ldg(x4 + 2 * (size_t)pa + 1));
        }
        const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
        const float y0 = scale_ptr ? s * y[r] : 0.0f;
        const float yn = (float)((double)y0 + acc);
        y[r] = yn;
        sq = (double)yn * (double)yn;
      }
    }
  }
  if (partial) {
    sh[threadIdx.x] = sq;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
  }
}

// y[c*8+k] = s*y[c*8+k] + sum_blocks val[b][k] * u[row[b]]   (vertex columns of A; warp per vertex)
__global__ void __launch_bounds__(kSpmvWarps * 32)
k_bspmv_cols(const long long *__restrict__ ptr, const int *__restrict__ row, const float4 *__restrict__ val,
             const float *__restrict__ u, float *__restrict__ y, const float *__restrict__ scale_ptr,
             float scale_sign, const int *__restrict__ collist, int ncols, int per_thread,
             double *__restrict__ partial, const int *__restrict__ stop) {
  __shared__ double sh[kSpmvWarps * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double sq = 0.0;
  if (stop == nullptr || *stop == 0) {
    int c = -1;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool writer = false;
    if (!per_thread) {
      const int wi = blockIdx.x * kSpmvWarps + w;
      if (wi < ncols) {
        c = collist[wi];
        const long long b0 = ptr[c], b1 = ptr[c + 1];
        for (long long b = b0 + lane; b < b1; b += 32) {
          const double uu = (double)__ldg(u + row[b]);
          const float4 v0 = val[2 * b], v1 = val[2 * b + 1];
          a[0] += (double)v0.x * uu; a[1] += (double)v0.y * uu; a[2] += (double)v0.z * uu; a[3] += (double)v0.w * uu;
          a[4] += (double)v1.x * uu; a[5] += (double)v1.y * uu; a[6] += (double)v1.z * uu; a[7] += (double)v1.w * uu;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = warp_sum(a[k]);
        writer = lane == 0;
      }
    } else {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t < ncols) {
        c = collist[t];
        for (long long b = ptr[c]; b < ptr[c + 1]; b++) {
          const double uu = (double)__ldg(u + row[b]);
          const float4 v0 = val[2 * b], v1 = val[2 * b + 1];
          a[0] += (double)v0.x * uu; a[1] += (double)v0.y * uu; a[2] += (double)v0.z * uu; a[3] += (double)v0.w * uu;
          a[4] += (double)v1.x * uu; a[5] += (double)v1.y * uu; a[6] += (double)v1.z * uu; a[7] += (double)v1.w * uu;
        }
        writer = true;
      }
    }
    if (writer) {
      const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float y0 = scale_ptr ? s * y[(size_t)c * 8 + k] : 0.0f;
        const float yn = (float)((double)y0 + a[k]);
        y[(size_t)c * 8 + k] = yn;
        sq += (double)yn * (double)yn;
      }
    }
  }
  if (partial) {
    sh[threadIdx.x] = sq;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent long-row / long-column products (depth-blocked layout).  ncu of the one-warp-per-row
// kernels above (profiles/r01_lsmr_iter_kernels.md): 3.1 TB/s (rows) and 2.35 TB/s (columns) of DRAM
// reads, long-scoreboard stalls 20-28 per issue -- every warp lived for one row only, i.e. spent its
// life in the dependent chain rowlist -> ptr -> pos -> x and in the block-wide partial-sum barrier.
// Here a fixed grid of warps strides over the row list: the next row's descriptor is fetched while
// the current row streams, four 36-byte blocks per lane are in flight at once, and the block-level
// sum of squares is reduced once per CTA.  The column list is sorted by length (longest first), so
// the grid-stride assignment is balanced.  Row r is always handled by warp (r mod #warps): the
// summation order, and therefore the result, is fixed for a given device.
// ---------------------------------------------------------------------------------------------
constexpr int kPersistBlocksPerSM = 4;

// one 32-byte block (8 floats) per lane in ONE request: sm_100a's 256-bit read-only load
// (SASS LDG.E.ENL2.256.CONSTANT).  With two 128-bit loads the scattered x gather alone costs 64 L1
// wavefronts per 32 blocks and caps the kernel near 4 TB/s; one 256-bit load halves that.
struct __align__(32) Float8 {
  float4 lo, hi;
};
__device__ __forceinline__ Float8 ldg256(const void *p) {
  Float8 r;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
      : "l"(p));
  return r;
}

template <bool W256>
__global__ void __launch_bounds__(kSpmvWarps * 32, kPersistBlocksPerSM)
k_bspmv_rows_p(const long long *__restrict__ ptr, const int *__restrict__ pos, const float4 *__restrict__ val,
               const float4 *__restrict__ x4, float *__restrict__ y, const float *__restrict__ scale_ptr,
               float scale_sign, const int *__restrict__ rowlist, int nrows, double *__restrict__ partial,
               const int *__restrict__ stop) {
  __shared__ double sh[kSpmvWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = gridDim.x * kSpmvWarps;
  double sq = 0.0;
  if (stop == nullptr || *stop == 0) {
    const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
    int wi = blockIdx.x * kSpmvWarps + w;
    int r = wi < nrows ? rowlist[wi] : 0;
    long long b0 = 0, b1 = 0;
    if (wi < nrows) {
      b0 = ptr[r];
      b1 = ptr[r + 1];
    }
    while (wi < nrows) {
      const int win = wi + nw;
      const int rn = win < nrows ? rowlist[win] : 0;  // next descriptor: in flight while this row streams
      long long nb0 = 0, nb1 = 0;
      if (win < nrows) {
        nb0 = ptr[rn];
        nb1 = ptr[rn + 1];
      }
      const float yold = (lane == 0 && scale_ptr) ? y[r] : 0.0f;
      double acc = 0.0;
      for (long long b = b0 + lane; b < b1; b += 128) {
        int p[4];
        Float8 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const long long bb = b + 32 * u;
          const bool ok = bb < b1;
          p[u] = ok ? pos[bb] : 0;
          v[u].lo = v[u].hi = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            if (W256) {
              v[u] = ldg256(val + 2 * bb);
            } else {
              v[u].lo = val[2 * bb];
              v[u].hi = val[2 * bb + 1];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          Float8 xx;
          if (W256) {
            xx = ldg256(x4 + 2 * (size_t)p[u]);
          } else {
            xx.lo = __ldg(x4 + 2 * (size_t)p[u]);
            xx.hi = __ldg(x4 + 2 * (size_t)p[u] + 1);
          }
          acc += dot8(v[u].lo, v[u].hi, xx.lo, xx.hi);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const float y0 = scale_ptr ? s * yold : 0.0f;
        const float yn = (float)((double)y0 + acc);
        y[r] = yn;
        sq += (double)yn * (double)yn;
      }
      wi = win;
      r = rn;
      b0 = nb0;
      b1 = nb1;
    }
  }
  if (partial) {
    if (lane == 0) sh[w] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < kSpmvWarps; k++) t += sh[k];
      partial[blockIdx.x] = t;
    }
  }
}

template <bool W256>
__global__ void __launch_bounds__(kSpmvWarps * 32, kPersistBlocksPerSM)
k_bspmv_cols_p(const long long *__restrict__ ptr, const int *__restrict__ row, const float4 *__restrict__ val,
               const float *__restrict__ u, float *__restrict__ y, const float *__restrict__ scale_ptr,
               float scale_sign, const int *__restrict__ collist, int ncols, double *__restrict__ partial,
               const int *__restrict__ stop) {
  __shared__ double sh[kSpmvWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = gridDim.x * kSpmvWarps;
  double sq = 0.0;
  if (stop == nullptr || *stop == 0) {
    const float s = scale_ptr ? scale_sign * (*scale_ptr) : 0.0f;
    int wi = blockIdx.x * kSpmvWarps + w;
    int c = wi < ncols ? collist[wi] : 0;
    long long b0 = 0, b1 = 0;
    if (wi < ncols) {
      b0 = ptr[c];
      b1 = ptr[c + 1];
    }
    while (wi < ncols) {
      const int win = wi + nw;
      const int cn = win < ncols ? collist[win] : 0;
      long long nb0 = 0, nb1 = 0;
      if (win < ncols) {
        nb0 = ptr[cn];
        nb1 = ptr[cn + 1];
      }
      float4 yo0 = make_float4(0.f, 0.f, 0.f, 0.f), yo1 = yo0;
      if (lane == 0 && scale_ptr) {
        yo0 = reinterpret_cast<const float4 *>(y)[2 * (size_t)c];
        yo1 = reinterpret_cast<const float4 *>(y)[2 * (size_t)c + 1];
      }
      double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (long long b = b0 + lane; b < b1; b += 128) {
        int rr[4];
        Float8 vv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const long long bb = b + 32 * q;
          const bool ok = bb < b1;
          rr[q] = ok ? row[bb] : 0;
          vv[q].lo = vv[q].hi = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            if (W256) {
              vv[q] = ldg256(val + 2 * bb);
            } else {
              vv[q].lo = val[2 * bb];
              vv[q].hi = val[2 * bb + 1];
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double uu = (double)__ldg(u + rr[q]);
          const float4 v0 = vv[q].lo, v1 = vv[q].hi;
          a[0] += (double)v0.x * uu; a[1] += (double)v0.y * uu; a[2] += (double)v0.z * uu; a[3] += (double)v0.w * uu;
          a[4] += (double)v1.x * uu; a[5] += (double)v1.y * uu; a[6] += (double)v1.z * uu; a[7] += (double)v1.w * uu;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = warp_sum(a[k]);
      if (lane == 0) {
        const float yo[8] = {yo0.x, yo0.y, yo0.z, yo0.w, yo1.x, yo1.y, yo1.z, yo1.w};
        float yn[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const float y0 = scale_ptr ? s * yo[k] : 0.0f;
          yn[k] = (float)((double)y0 + a[k]);
          sq += (double)yn[k] * (double)yn[k];
        }
        reinterpret_cast<float4 *>(y)[2 * (size_t)c] = make_float4(yn[0], yn[1], yn[2], yn[3]);
        reinterpret_cast<float4 *>(y)[2 * (size_t)c + 1] = make_float4(yn[4], yn[5], yn[6], yn[7]);
      }
      wi = win;
      c = cn;
      b0 = nb0;
      b1 = nb1;
    }
  }
  if (partial) {
    if (lane == 0) sh[w] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < kSpmvWarps; k++) t += sh[k];
      partial[blockIdx.x] = t;
    }
  }
}

// build helpers
__global__ void k_blk_keys(const int *__restrict__ rows1, const int *__restrict__ cols1, long long nnz, int P,
                           long long M, int by_rows, unsigned long long *__restrict__ keys, int *__restrict__ perm) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += tot) {
    const long long r = rows1[i] - 1, c = cols1[i] - 1;
    const long long pos0 = c % P;
    keys[i] = by_rows ? (unsigned long long)(r * (long long)P + pos0) : (unsigned long long)(pos0 * M + r);
    perm[i] = (int)i;
  }
}
__global__ void k_blk_heads(const unsigned long long *__restrict__ keys, long long nnz, int *__restrict__ head) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += tot)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
__global__ void k_blk_fill(const unsigned long long *__restrict__ keys, const int *__restrict__ perm,
                           const long long *__restrict__ bid_incl, const int *__restrict__ cols1,
                           const float *__restrict__ vals, long long nnz, int P, long long M, int by_rows,
                           int *__restrict__ bidx, float *__restrict__ bval, int *__restrict__ cnt) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += tot) {
    const long long b = bid_incl[i] - 1;
    const unsigned long long key = keys[i];
    const int src = perm[i];
    const int k = (cols1[src] - 1) / P;
    // duplicates (same row, same column) add up like in the reference's aprod
    atomicAdd(bval + b * 8 + k, vals[src]);
    if (i == 0 || keys[i - 1] != key) {
      long long outer, inner;
      if (by_rows) {
        outer = (long long)(key / (unsigned long long)P);
        inner = (long long)(key % (unsigned long long)P);
      } else {
        outer = (long long)(key / (unsigned long long)M);
        inner = (long long)(key % (unsigned long long)M);
      }
      bidx[b] = (int)inner;
      atomicAdd(cnt + outer, 1);
    }
  }
}
// x (reference order k*P+pos) <-> internal order pos*8+k
__global__ void k_unpermute(const float *__restrict__ xin, float *__restrict__ xout, int P, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P * K) {
    const int k = i / P, pos = i % P;
    xout[i] = xin[(size_t)pos * 8 + k];
  }
}

// one-block deterministic reduction of `np` partials (fixed strided order + tree)
__device__ double block_reduce_partials(const double *partial, int np) {
  __shared__ double sh[1024];
  double t = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) t += partial[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  return sh[0];
}

// ---------------------------------------------------------------------------------------------
// scalar state (all REAL*4 like the reference's locals, lsmrModule.f90:336-351)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float d2norm_f(float a, float b) {  // lsmrModule.f90:686-709
  const float scale = fabsf(a) + fabsf(b);
  if (scale == 0.0f) return 0.0f;
  const float ra = a / scale, rb = b / scale;
  return scale * sqrtf(ra * ra + rb * rb);
}

// after u = A v - alpha u : beta = ||u||  (plus its all-reduced partner in multi-GPU runs); also
// advances the circular-buffer pointer of the local reorthogonalisation (lsmrModule.f90:717-724)
__global__ void k_beta(LsmrScalars *S, const double *partial, int np, const double *extra, int localVecs) {
  if (S->stop) return;
  double s = block_reduce_partials(partial, np);
  if (threadIdx.x == 0) {
    if (extra) s = *extra;  // multi-GPU: globally reduced sum of squares
    S->sum_u2 = s;
    const float beta = (float)sqrt(s);
    S->beta = beta;
    S->beta_pos = beta > 0.0f;
    S->inv_beta = beta > 0.0f ? 1.0f / beta : 0.0f;
    S->neg_beta = -beta;
    if (localVecs > 0 && beta > 0.0f) {
      if (S->localPointer < localVecs) {
        S->localPointer = S->localPointer + 1;
      } else {
        S->localPointer = 1;
        S->queueFull = 1;
      }
      S->enq_slot = S->localPointer - 1;
      S->orthoLimit = S->queueFull ? localVecs : S->localPointer;
    }
  }
}

// u *= 1/beta (m) ; localV enqueue (lsmrModule.f90:715-726) ; v *= -beta is folded into the
// spmtv kernel's scale argument.
__global__ void k_scale_u_enqueue(const LsmrScalars *S, float *u, int m, const float *v,
                                  float *localV, int n, int do_enqueue) {
  if (S->stop || !S->beta_pos) return;
  const float ib = S->inv_beta;
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += tot) u[i] = ib * u[i];
  if (do_enqueue) {
    float *q = localV + (size_t)S->enq_slot * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) q[i] = v[i];
  }
}

// reorthogonalisation step c (lsmrModule.f90:741-746): every block first finishes step c-1
// (d_{c-1} = REAL*4 dot product, reduced from the previous step's partials in a fixed order),
// then v -= d_{c-1} q_{c-1}, then partial dot(v, q_c) (c < limit) or partial sum(v^2) (c == limit).
// Partials alternate between two buffers (pin = buffer of step c-1, pout = buffer of step c);
// steps beyond the current limit do nothing, so the alpha partials stay in buffer (limit & 1).
__global__ void k_reorth(const LsmrScalars *S, float *v, const float *localV, int n, int c, const double *pin,
                         double *pout, int np) {
  __shared__ double sh[256];
  if (S->stop || !S->beta_pos) return;
  const int lim = S->orthoLimit;
  if (c > lim) return;
  float dprev = 0.0f;
  if (c > 0) dprev = (float)block_reduce_partials(pin, np);
  __syncthreads();
  double acc = 0.0;
  const float *qp = (c > 0) ? localV + (size_t)(c - 1) * n : nullptr;
  const float *qc = (c < lim) ? localV + (size_t)c * n : nullptr;
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) {
    float vi = v[i];
    if (qp) {
      vi = vi - dprev * qp[i];
      v[i] = vi;
    }
    acc += qc ? (double)vi * (double)qc[i] : (double)vi * (double)vi;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) pout[blockIdx.x] = sh[0];
}

__device__ void rotations_scalar(LsmrScalars *S, float damp);
__device__ void tests_scalar(LsmrScalars *S, double sumx2, float atol, float btol, float ctol, int itnlim,
                             int force_iters);

// plane rotations + estimates up to the vector updates (lsmrModule.f90:508-537)
__global__ void k_rotations(LsmrScalars *S, float damp, const double *part0, const double *part1, int np) {
  if (S->stop) return;
  if (S->beta_pos) {  // alpha = ||v|| from the partials of the last reorthogonalisation step
    const double s = block_reduce_partials((S->orthoLimit & 1) ? part1 : part0, np);
    if (threadIdx.x == 0) {
      const float a = (float)sqrt(s);
      S->alpha = a;
      S->inv_alpha = a > 0.0f ? 1.0f / a : 1.0f;
      S->alpha_pos = a > 0.0f;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  rotations_scalar(S, damp);
}

// scalar recurrences of one iteration after alpha is known (lsmrModule.f90:508-584); one thread
__device__ void rotations_scalar(LsmrScalars *S, float damp) {
  const float alpha = S->alpha, beta = S->beta;
  const float alphahat = d2norm_f(S->alphabar, damp);
  const float chat = S->alphabar / alphahat;
  const float shat = damp / alphahat;
  const float rhoold = S->rho;
  const float rho = d2norm_f(alphahat, beta);
  const float c = alphahat / rho;
  const float s = beta / rho;
  const float thetanew = s * alpha;
  S->alphabar = c * alpha;
  const float rhobarold = S->rhobar;
  const float zetaold = S->zeta;
  const float thetabar = S->sbar * rho;
  const float rhotemp = S->cbar * rho;
  const float rhobar = d2norm_f(S->cbar * rho, thetanew);
  const float cbar = S->cbar * rho / rhobar;
  const float sbar = thetanew / rhobar;
  const float zeta = cbar * S->zetabar;
  const float zetabar = -sbar * S->zetabar;
  S->rho = rho;
  S->rhobar = rhobar;
  S->cbar = cbar;
  S->sbar = sbar;
  S->zeta = zeta;
  S->zetabar = zetabar;
  S->f1 = thetabar * rho / (rhoold * rhobarold);
  S->f2 = zeta / (rho * rhobar);
  S->f3 = thetanew / rho;
  // ||r|| estimate (lsmrModule.f90:546-572)
  const float betaacute = chat * S->betadd;
  const float betacheck = -shat * S->betadd;
  const float betahat = c * betaacute;
  S->betadd = -s * betaacute;
  const float thetatildeold = S->thetatilde;
  const float rhotildeold = d2norm_f(S->rhodold, thetabar);
  const float ctildeold = S->rhodold / rhotildeold;
  const float stildeold = thetabar / rhotildeold;
  S->thetatilde = stildeold * rhobar;
  S->rhodold = ctildeold * rhobar;
  S->betad = -stildeold * S->betad + ctildeold * betahat;
  S->tautildeold = (zetaold - thetatildeold * S->tautildeold) / rhotildeold;
  const float taud = (zeta - S->thetatilde * S->tautildeold) / S->rhodold;
  S->d = S->d + betacheck * betacheck;
  {
    const float e = S->betad - taud;
    S->normr = sqrtf(S->d + e * e + S->betadd * S->betadd);
  }
  // ||A||, cond(A) (lsmrModule.f90:574-584)
  S->normA2 = S->normA2 + beta * beta;
  S->normA = sqrtf(S->normA2);
  S->normA2 = S->normA2 + alpha * alpha;
  S->maxrbar = fmaxf(S->maxrbar, rhobarold);
  if (S->itn + 1 > 1) S->minrbar = fminf(S->minrbar, rhobarold);
  S->condA = fmaxf(S->maxrbar, rhotemp) / fminf(S->minrbar, rhotemp);
  S->normAr = fabsf(zetabar);
}

// v *= 1/alpha ; hbar = h - f1 hbar ; x = x + f2 hbar ; h = v - f3 h ; partial sum(x^2)
// (lsmrModule.f90:503-505, 539-541)
__global__ void k_update(const LsmrScalars *S, float *v, float *h, float *hbar, float *x, int n,
                         double *partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  if (!S->stop) {
    const bool scale_v = S->beta_pos && S->alpha_pos;
    const float ia = S->inv_alpha, f1 = S->f1, f2 = S->f2, f3 = S->f3;
    const long long tot = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) {
      float vi = v[i];
      if (scale_v) {
        vi = ia * vi;
        v[i] = vi;
      }
      const float hb = h[i] - f1 * hbar[i];
      hbar[i] = hb;
      const float xi = x[i] + f2 * hb;
      x[i] = xi;
      h[i] = vi - f3 * h[i];
      acc += (double)xi * (double)xi;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// normx + stopping rules (lsmrModule.f90:586-616)
__global__ void k_tests(LsmrScalars *S, const double *partial, int np, float atol, float btol,
                        float ctol, int itnlim, int force_iters) {
  if (S->stop) return;
  const double s = block_reduce_partials(partial, np);
  if (threadIdx.x != 0) return;
  tests_scalar(S, s, atol, btol, ctol, itnlim, force_iters);
}

// normx + stopping rules (lsmrModule.f90:586-616); one thread
__device__ void tests_scalar(LsmrScalars *S, double s, float atol, float btol, float ctol, int itnlim,
                             int force_iters) {
  const float one = 1.0f;
  S->itn = S->itn + 1;
  const float normx = (float)sqrt(s);
  S->normx = normx;
  const float test1 = S->normr / S->normb;
  const float test2 = S->normAr / (S->normA * S->normr);
  const float test3 = one / S->condA;
  const float t1 = test1 / (one + S->normA * normx / S->normb);
  const float rtol = btol + atol * S->normA * normx / S->normb;
  int istop = 0;
  if (S->itn >= itnlim) istop = 7;
  if (!force_iters) {
    if (one + test3 <= one) istop = 6;
    if (one + test2 <= one) istop = 5;
    if (one + t1 <= one) istop = 4;
    if (test3 <= ctol) istop = 3;
    if (test2 <= atol) istop = 2;
    if (test1 <= rtol) istop = 1;
  }
  S->istop = istop;
  if (istop != 0) S->stop = 1;
}

// ---------------------------------------------------------------------------------------------
// Fused small-vector phases of one iteration on ONE thread-block cluster.
// The n- and m-vectors of LSMR are tiny next to the matrix (n = 1.3e5 floats at cfg 3), so the 16
// separate launches of the unfused path (beta, scale, 11 reorthogonalisation steps, rotations,
// update, tests) are pure launch/drain latency: ~100 us of a 460 us iteration.  Here a cluster of
// kCl CTAs x 1024 threads keeps its slice of v in shared memory across all modified Gram-Schmidt
// steps, reduces dot products CTA -> cluster through distributed shared memory in a fixed order
// (every CTA ends up with the same bits), and separates dependent steps with the hardware
// cluster barrier instead of a kernel boundary.  Arithmetic per element and the order of the
// reference's statements (lsmrModule.f90:484-616, 715-748) are those of the unfused kernels.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kFusedThreads = 1024;
constexpr int kClMax = 16;

// deterministic CTA sum (1024 threads): shuffle tree per warp, then a shuffle tree over the 32 warp sums
__device__ __forceinline__ double cta_sum(double v, double *wsum) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
  if (lane == 0) wsum[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = wsum[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(kFull, t, o);
    if (lane == 0) wsum[32] = t;
  }
  __syncthreads();
  const double r = wsum[32];
  __syncthreads();
  return r;
}

// cluster all-reduce of one double per CTA: every CTA deposits its sum into slot [rank] of every
// CTA's `red` array (DSMEM stores), hardware cluster barrier, then sums the slots in rank order.
__device__ __forceinline__ double cluster_sum(cg::cluster_group &cl, double cta_val, double *red /*[kClMax]*/) {
  const unsigned nb = cl.num_blocks(), me = cl.block_rank();
  if (threadIdx.x < nb) *cl.map_shared_rank(red + me, threadIdx.x) = cta_val;
  cl.sync();
  double t = 0.0;
  for (unsigned r = 0; r < nb; r++) t += red[r];
  return t;
}

// beta = ||u|| (+ circular-buffer bookkeeping), u *= 1/beta, localV(:,slot) = v
// (lsmrModule.f90:486-487, 715-726): k_beta + k_scale_u_enqueue in one launch.
__global__ void __launch_bounds__(kFusedThreads, 1)
k_fused_beta(LsmrScalars *S, const double *partial, int np, const double *extra, int localVecs, float *u, int m,
             const float *v, float *localV, int n) {
  cg::cluster_group cl = cg::this_cluster();
  if (S->stop) return;  // uniform over the cluster
  double s = block_reduce_partials(partial, np);  // same fixed order in every CTA -> same bits
  if (extra) s = *extra;
  const float beta = (float)sqrt(s);
  const bool bp = beta > 0.0f;
  const float ib = bp ? 1.0f / beta : 0.0f;
  int lp = S->localPointer, qf = S->queueFull, slot = S->enq_slot, lim = S->orthoLimit;
  if (localVecs > 0 && bp) {
    if (lp < localVecs) {
      lp = lp + 1;
    } else {
      lp = 1;
      qf = 1;
    }
    slot = lp - 1;
    lim = qf ? localVecs : lp;
  }
  cl.sync();  // every CTA has read the old bookkeeping before rank 0 overwrites it
  if (cl.block_rank() == 0 && threadIdx.x == 0) {
    S->sum_u2 = s;
    S->beta = beta;
    S->beta_pos = bp;
    S->inv_beta = ib;
    S->neg_beta = -beta;
    S->localPointer = lp;
    S->queueFull = qf;
    S->enq_slot = slot;
    S->orthoLimit = lim;
  }
  if (!bp) return;
  const long long tot = (long long)cl.num_blocks() * blockDim.x;
  const long long t0 = (long long)cl.block_rank() * blockDim.x + threadIdx.x;
  for (long long i = t0; i < m; i += tot) u[i] = ib * u[i];
  if (localVecs > 0) {
    float *q = localV + (size_t)slot * n;
    for (long long i = t0; i < n; i += tot) q[i] = v[i];
  }
}

// local reorthogonalisation (modified Gram-Schmidt against the stored window), alpha = ||v||,
// plane rotations, v/h/hbar/x updates, ||x|| and the stopping tests (lsmrModule.f90:498-616, 731-748):
// k_reorth x (localVecs+1) + k_rotations + k_update + k_tests in one launch.  CTA `rank` owns elements
// [rank*chunk, min(n, (rank+1)*chunk)).  The Gram-Schmidt steps are serial (each dot product needs the
// previous subtraction), so what matters is the latency of one step: with <= kEPT elements per thread
// the slice of v and two window vectors live in REGISTERS and the next window vector is prefetched
// while the current dot product is being reduced; larger slices keep v in shared memory.
constexpr int kEPT = 10;

struct TailShared {
  double wsum[33];
  double red[2][kClMax];
  float bc[4];
  int bflag;
};

// rank 0 / thread 0: alpha, rotations, broadcast block for the update phase
__device__ __forceinline__ void tail_scalars(LsmrScalars *S, bool bp, double sumv2, float damp, TailShared &sh) {
  if (bp) {  // alpha = ||v|| after the last subtraction (lsmrModule.f90:503)
    const float a = (float)sqrt(sumv2);
    S->alpha = a;
    S->inv_alpha = a > 0.0f ? 1.0f / a : 1.0f;
    S->alpha_pos = a > 0.0f;
  }
  rotations_scalar(S, damp);
  sh.bc[0] = S->inv_alpha;
  sh.bc[1] = S->f1;
  sh.bc[2] = S->f2;
  sh.bc[3] = S->f3;
  sh.bflag = (S->beta_pos && S->alpha_pos) ? 1 : 0;
}

__global__ void __launch_bounds__(kFusedThreads, 1)
k_fused_tail(LsmrScalars *S, float *__restrict__ v, const float *__restrict__ localV, float *__restrict__ h,
             float *__restrict__ hbar, float *__restrict__ x, int n, int chunk, float damp, float atol, float btol,
             float ctol, int itnlim, int force_iters) {
  extern __shared__ float vs[];
  __shared__ TailShared sh;
  cg::cluster_group cl = cg::this_cluster();
  if (S->stop) return;  // uniform over the cluster
  const int rank = (int)cl.block_rank(), tid = threadIdx.x;
  const int i0 = min(n, rank * chunk), cnt = min(n, i0 + chunk) - i0;
  const bool bp = S->beta_pos != 0;
  const int lim = S->orthoLimit;
  int buf = 0;
  double tot = 0.0;
  if (chunk <= kEPT * kFusedThreads) {
    // ------------------------------------------------ register path
    float vr[kEPT], qa[kEPT], qb[kEPT];
    bool ok[kEPT];
#pragma unroll
    for (int u = 0; u < kEPT; u++) {
      const int j = tid + u * kFusedThreads;
      ok[u] = j < cnt;
      vr[u] = ok[u] ? v[i0 + j] : 0.0f;
      qa[u] = 0.0f;
      qb[u] = (bp && lim > 0 && ok[u]) ? localV[i0 + j] : 0.0f;  // q_0
    }
    if (bp) {
      float dprev = 0.0f;
      for (int c = 0; c <= lim; c++) {
        // here: qa = q_{c-1} (c > 0), qb = q_c (c < lim)
        double acc = 0.0;
        const bool more = c + 1 < lim;
        const float *qn = localV + (size_t)(c + 1) * n + i0;
#pragma unroll
        for (int u = 0; u < kEPT; u++) {
          if (c > 0) vr[u] = vr[u] - dprev * qa[u];
          qa[u] = (more && ok[u]) ? qn[tid + u * kFusedThreads] : 0.0f;  // prefetch q_{c+1}; lands during the reduction
        }
#pragma unroll
        for (int u = 0; u < kEPT; u++) {
          acc += (c < lim) ? (double)vr[u] * (double)qb[u] : (double)vr[u] * (double)vr[u];
        }
        tot = cluster_sum(cl, cta_sum(acc, sh.wsum), sh.red[buf]);
        buf ^= 1;
        dprev = (float)tot;
#pragma unroll
        for (int u = 0; u < kEPT; u++) {  // rotate: q_c becomes "previous", the prefetched q_{c+1} "current"
          const float t = qa[u];
          qa[u] = qb[u];
          qb[u] = t;
        }
      }
    }
    if (rank == 0 && tid == 0) tail_scalars(S, bp, tot, damp, sh);
    // h, hbar, x of this thread: issued before the barrier, consumed after it
    float hr[kEPT], hbr[kEPT], xr[kEPT];
#pragma unroll
    for (int u = 0; u < kEPT; u++) {
      const int i = i0 + tid + u * kFusedThreads;
      hr[u] = ok[u] ? h[i] : 0.0f;
      hbr[u] = ok[u] ? hbar[i] : 0.0f;
      xr[u] = ok[u] ? x[i] : 0.0f;
    }
    cl.sync();
    const float *bc0 = cl.map_shared_rank(sh.bc, 0);
    const float ia = bc0[0], f1 = bc0[1], f2 = bc0[2], f3 = bc0[3];
    const bool scale_v = *cl.map_shared_rank(&sh.bflag, 0) != 0;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < kEPT; u++) {
      if (ok[u]) {
        const int i = i0 + tid + u * kFusedThreads;
        float vi = vr[u];
        if (scale_v) vi = ia * vi;
        v[i] = vi;
        const float hb = hr[u] - f1 * hbr[u];
        hbar[i] = hb;
        const float xi = xr[u] + f2 * hb;
        x[i] = xi;
        h[i] = vi - f3 * hr[u];
        acc += (double)xi * (double)xi;
      }
    }
    tot = cluster_sum(cl, cta_sum(acc, sh.wsum), sh.red[buf]);  // also keeps rank 0's smem alive until all have read it
    if (rank == 0 && tid == 0) tests_scalar(S, tot, atol, btol, ctol, itnlim, force_iters);
    return;
  }
  // -------------------------------------------------- shared-memory path (large n)
  for (int j = tid; j < cnt; j += kFusedThreads) vs[j] = v[i0 + j];
  __syncthreads();
  if (bp) {
    float dprev = 0.0f;
    for (int c = 0; c <= lim; c++) {
      const float *qp = (c > 0) ? localV + (size_t)(c - 1) * n + i0 : nullptr;
      const float *qc = (c < lim) ? localV + (size_t)c * n + i0 : nullptr;
      double acc = 0.0;
      for (int j0 = tid; j0 < cnt; j0 += 4 * kFusedThreads) {
        float a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {  // all loads of the unrolled group in flight together
          const int j = j0 + u * kFusedThreads;
          a[u] = (qp && j < cnt) ? qp[j] : 0.0f;
          b[u] = (qc && j < cnt) ? qc[j] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u * kFusedThreads;
          if (j < cnt) {
            float vi = vs[j];
            if (qp) {
              vi = vi - dprev * a[u];
              vs[j] = vi;
            }
            acc += qc ? (double)vi * (double)b[u] : (double)vi * (double)vi;
          }
        }
      }
      tot = cluster_sum(cl, cta_sum(acc, sh.wsum), sh.red[buf]);
      buf ^= 1;
      dprev = (float)tot;
    }
  }
  if (rank == 0 && tid == 0) tail_scalars(S, bp, tot, damp, sh);
  cl.sync();
  const float *bc0 = cl.map_shared_rank(sh.bc, 0);
  const float ia = bc0[0], f1 = bc0[1], f2 = bc0[2], f3 = bc0[3];
  const bool scale_v = *cl.map_shared_rank(&sh.bflag, 0) != 0;
  double acc = 0.0;
  for (int j = tid; j < cnt; j += kFusedThreads) {
    const int i = i0 + j;
    float vi = vs[j];
    if (scale_v) vi = ia * vi;
    v[i] = vi;
    const float hi = h[i];
    const float hb = hi - f1 * hbar[i];
    hbar[i] = hb;
    const float xi = x[i] + f2 * hb;
    x[i] = xi;
    h[i] = vi - f3 * hi;
    acc += (double)xi * (double)xi;
  }
  tot = cluster_sum(cl, cta_sum(acc, sh.wsum), sh.red[buf]);
  if (rank == 0 && tid == 0) tests_scalar(S, tot, atol, btol, ctol, itnlim, force_iters);
}

// initial phase: after u=b, beta=||u||: scale; after v=A'u: alpha, scale, h=v, localV(:,1)=v
__global__ void k_init_scale_copy(float *dst, const float *src, float *dst2, const float *scale,
                                  long long n) {
  const float s = *scale;
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) {
    const float t = s * src[i];
    dst[i] = t;
    if (dst2) dst2[i] = t;
  }
}
__global__ void k_sumsq(const float *x, long long n, double *partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot)
    acc += (double)x[i] * (double)x[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// dist: reduce partials into one device double (then all-reduced by NCCL)
// ---------------------------------------------------------------------------------------------
// One-shot exchange of the distributed LSMR over peer memory (NVLink / NVSwitch), replacing the
// per-iteration ncclAllReduce: rows of A are partitioned over ranks, so every iteration needs the sum
// over ranks of the partial A'u (n floats) and of the partial ||u||^2 (one double) --
// lsmrModule.f90:486-497 with aprod.f90:40-55 split by rows.  Every rank owns an exchange buffer that all
// peers have mapped (CUDA IPC):  [slot 0 | slot 1 | arrival flags | epoch],  slot = n floats + 1 double.
//   k_xchg_signal   after the local A'u has been written into this rank's slot (epoch parity): bump the local
//                   epoch and store it into flag[rank] of EVERY rank's buffer (system-scope release);
//   k_xchg_reduce   wait until all flags of the local buffer reached the epoch (acquire), then every rank adds
//                   the same n-vectors in the same order rank 0..N-1 -- reading each peer's slot straight over
//                   NVLink -- so the replicated vectors stay bit-identical on all ranks; fused with the update
//                   v = -beta v + (1/beta) A'u of :495-497 needs beta, hence the separate k_fused_beta between.
// Two slots: a rank can run at most one iteration ahead of its slowest peer (it needs that peer's flag of the
// intermediate iteration), so slot parity never collides.  No NCCL call, no host interaction: the whole
// iteration (two of them, one per parity) is a CUDA graph.  A watchdog turns a missing peer into an error
// instead of a hang.
struct XchgHdr {
  int flag[64];
  int epoch;
  int error;
};
__host__ __device__ __forceinline__ XchgHdr *xchg_hdr(char *buf, size_t stride) { return reinterpret_cast<XchgHdr *>(buf + 2 * stride); }

__global__ void k_xchg_signal(const LsmrScalars *S, char *const *peers, int rank, int nranks, size_t stride,
                              const double *part, int np) {
  if (S->stop) return;
  char *mine = peers[rank];
  XchgHdr *h = xchg_hdr(mine, stride);
  const double usum = block_reduce_partials(part, np);  // local ||u||^2, same fixed order as everywhere
  __shared__ int ep;
  if (threadIdx.x == 0) {
    ep = h->epoch + 1;
    h->epoch = ep;
    *reinterpret_cast<double *>(mine + (size_t)(ep & 1) * stride + stride - sizeof(double)) = usum;
    __threadfence_system();  // the slot (written by the product kernel before this one, and usum) before the flags
  }
  __syncthreads();
  if ((int)threadIdx.x < nranks) {
    volatile int *f = &xchg_hdr(peers[threadIdx.x], stride)->flag[rank];
    *f = ep;
  }
}

__global__ void k_xchg_reduce(LsmrScalars *S, char *const *peers, int rank, int nranks, size_t stride, int n,
                              float *vsum, double *usum_out) {
  if (S->stop) return;
  XchgHdr *h = xchg_hdr(peers[rank], stride);
  __shared__ int ep;
  if (threadIdx.x == 0) {
    const int e = h->epoch;
    const long long t0 = clock64();
    bool ok = true;
    for (int r = 0; r < nranks && ok; r++) {
      volatile int *f = &h->flag[r];
      while (*f < e) {
        if (clock64() - t0 > 8000000000ll) {  // ~4 s: a peer is gone
          ok = false;
          break;
        }
      }
    }
    if (!ok) h->error = 1;
    __threadfence_system();
    ep = ok ? e : -1;
  }
  __syncthreads();
  if (ep < 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) S->stop = 1;  // the host reports DSURF_ERR_NCCL (exchange timed out)
    return;
  }
  const size_t off = (size_t)(ep & 1) * stride;
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) {
    float acc = 0.0f;
    for (int r = 0; r < nranks; r++) acc += __ldcv(reinterpret_cast<const float *>(peers[r] + off) + i);
    vsum[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double u = 0.0;
    for (int r = 0; r < nranks; r++) u += __ldcv(reinterpret_cast<const double *>(peers[r] + off + stride - sizeof(double)));
    *usum_out = u;
  }
}

__global__ void k_reduce_to(double *out, const double *partial, int np) {
  const double s = block_reduce_partials(partial, np);
  if (threadIdx.x == 0) *out = s;
}
// dist: v = (-beta) v + (1/beta) * vpart   (vpart = all-reduced A'u_raw)
__global__ void k_combine_v(const LsmrScalars *S, float *v, const float *vpart, int n) {
  if (S->stop || !S->beta_pos) return;
  const float nb = S->neg_beta, ib = S->inv_beta;
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot)
    v[i] = nb * v[i] + ib * vpart[i];
}
__global__ void k_init_beta(LsmrScalars *S, const double *partial, int np, const double *extra) {
  double s = block_reduce_partials(partial, np);
  if (threadIdx.x == 0) {
    if (extra) s = *extra;
    const float beta = (float)sqrt(s);
    S->beta = beta;
    S->inv_beta = beta > 0.0f ? 1.0f / beta : 0.0f;
    S->beta_pos = beta > 0.0f;
  }
}
__global__ void k_init_alpha(LsmrScalars *S, const double *partial, int np, int localVecs) {
  const double s = block_reduce_partials(partial, np);
  if (threadIdx.x == 0) {
    const float alpha = S->beta_pos ? (float)sqrt(s) : 0.0f;
    const float beta = S->beta;
    S->alpha = alpha;
    S->inv_alpha = alpha > 0.0f ? 1.0f / alpha : 1.0f;
    S->alpha_pos = alpha > 0.0f;
    S->normAr = alpha * beta;
    S->itn = 0;
    S->zetabar = alpha * beta;
    S->alphabar = alpha;
    S->rho = 1;
    S->rhobar = 1;
    S->cbar = 1;
    S->sbar = 0;
    S->betadd = beta;
    S->betad = 0;
    S->rhodold = 1;
    S->tautildeold = 0;
    S->thetatilde = 0;
    S->zeta = 0;
    S->d = 0;
    S->normA2 = alpha * alpha;
    S->maxrbar = 0.0f;
    S->minrbar = 1e+30f;
    S->normb = beta;
    S->istop = 0;
    S->normr = beta;
    S->localPointer = 1;
    S->queueFull = 0;
    S->enq_slot = 0;
    S->orthoLimit = localVecs > 0 ? 1 : 0;
    S->dot_d = 0.0f;
    S->stop = (alpha * beta == 0.0f) ? 1 : 0;  // lsmrModule.f90:399-400: exit if A'b = 0
    S->normA = 0;
    S->condA = 0;
    S->normx = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// COO -> CSR / CSC
// ---------------------------------------------------------------------------------------------
__global__ void k_count(const int *keys1, long long nnz, int *cnt) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += tot)
    atomicAdd(cnt + (keys1[i] - 1), 1);
}
__global__ void k_check_sorted(const int *keys1, long long nnz, int *flag) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i + 1 < nnz; i += tot)
    if (keys1[i] > keys1[i + 1]) *flag = 1;
}
__global__ void k_iota(int *a, long long n) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) a[i] = (int)i;
}
__global__ void k_gather_pairs(const int *perm, const int *other1, const float *vals, long long nnz,
                               int *out_idx0, float *out_val) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz; i += tot) {
    const int p = perm[i];
    out_idx0[i] = other1[p] - 1;
    out_val[i] = vals[p];
  }
}
__global__ void k_to_ptr(const int *cnt, const long long *scan_excl, int nrows, long long nnz,
                         long long *ptr) {
  const int tot = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nrows; i += tot)
    ptr[i] = (i < nrows) ? scan_excl[i] : nnz;
}
__global__ void k_widen(const int *cnt, long long *out, int n) {
  const int tot = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += tot) out[i] = cnt[i];
}

static int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Builds ptr/idx/val compressed by `keys1` (1-based, nkeys distinct) from COO; stable in k.
static int build_compressed(cudaStream_t st, const int *keys1, const int *other1, const float *vals,
                            long long nnz, int nkeys, DevBuf<long long> &ptr, DevBuf<int> &idx,
                            DevBuf<float> &val) {
  DevBuf<int> cnt, flag;
  DevBuf<long long> wide, scan;
  if (cnt.reserve(nkeys + 1) || flag.reserve(1) || wide.reserve(nkeys + 1) || scan.reserve(nkeys + 1) ||
      ptr.reserve(nkeys + 1) || idx.reserve(nnz > 0 ? nnz : 1) || val.reserve(nnz > 0 ? nnz : 1)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed in build_compressed");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cudaMemsetAsync(cnt.p, 0, (nkeys + 1) * sizeof(int), st));
  DS_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
  if (nnz > 0) {
    k_count<<<grid_for(nnz), 256, 0, st>>>(keys1, nnz, cnt.p);
    k_check_sorted<<<grid_for(nnz), 256, 0, st>>>(keys1, nnz, flag.p);
  }
  k_widen<<<grid_for(nkeys + 1), 256, 0, st>>>(cnt.p, wide.p, nkeys + 1);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, wide.p, scan.p, nkeys + 1, st);
  DevBuf<char> tmp;
  if (tmp.reserve(tb + 16)) return DSURF_ERR_CUDA;
  cub::DeviceScan::ExclusiveSum(tmp.p, tb, wide.p, scan.p, nkeys + 1, st);
  k_to_ptr<<<grid_for(nkeys + 1), 256, 0, st>>>(cnt.p, scan.p, nkeys, nnz, ptr.p);
  int hflag = 0;
  DS_CUDA(cudaMemcpyAsync(&hflag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  if (nnz == 0) return DSURF_OK;
  if (!hflag) {  // already grouped by key in ascending order: identity permutation
    DevBuf<int> perm;
    if (perm.reserve(nnz)) return DSURF_ERR_CUDA;
    k_iota<<<grid_for(nnz), 256, 0, st>>>(perm.p, nnz);
    k_gather_pairs<<<grid_for(nnz), 256, 0, st>>>(perm.p, other1, vals, nnz, idx.p, val.p);
    DS_CUDA(cudaStreamSynchronize(st));
    return DSURF_OK;
  }
  // stable radix sort of (key, k) pairs
  DevBuf<int> k_in, k_out, p_in, p_out;
  if (k_out.reserve(nnz) || p_in.reserve(nnz) || p_out.reserve(nnz)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (sort buffers)");
    return DSURF_ERR_CUDA;
  }
  k_iota<<<grid_for(nnz), 256, 0, st>>>(p_in.p, nnz);
  int end_bit = 1;
  while ((1ll << end_bit) <= (long long)nkeys + 1 && end_bit < 31) end_bit++;
  size_t sb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sb, keys1, k_out.p, p_in.p, p_out.p, (long long)nnz, 0, end_bit, st);
  DevBuf<char> stmp;
  if (stmp.reserve(sb + 16)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (sort temp)");
    return DSURF_ERR_CUDA;
  }
  cub::DeviceRadixSort::SortPairs(stmp.p, sb, keys1, k_out.p, p_in.p, p_out.p, (long long)nnz, 0, end_bit, st);
  k_gather_pairs<<<grid_for(nnz), 256, 0, st>>>(p_out.p, other1, vals, nnz, idx.p, val.p);
  DS_CUDA(cudaStreamSynchronize(st));
  DS_CUDA(cudaGetLastError());
  return DSURF_OK;
}

}  // namespace dsurf

using namespace dsurf;

// one compressed structure (CSR of A, or CSC of A) with its long/short row classification
struct Compressed {
  DevBuf<long long> ptr;
  DevBuf<int> idx;
  DevBuf<float> val;
  DevBuf<int> longl, shortl;
  int nrows = 0, nlong = 0, nshort = 0;
  int kind = 0;  // 0 scalar CSR/CSC, 1 depth-blocked rows, 2 depth-blocked vertex columns
  long long nblk = 0;
  // grid of the long-row kernel: one warp per row (generic layout) or the persistent grid (blocked layouts)
  // Blocked rows (kind 1): persistent grid, kPersistBlocksPerSM CTAs per SM.  Blocked vertex columns
  // (kind 2): one warp per column, longest first -- columns are ~4x longer than rows and few (P), so the
  // hardware block scheduler balances them better than a fixed stride (measured: 116 vs 159 us).
  // DSURF_LSMR_GRID_ROWS / DSURF_LSMR_GRID_COLS = CTAs per SM (0 = one warp per row) override for A/B runs.
  int long_grid() const {
    const int per_row = (nlong + kSpmvWarps - 1) / kSpmvWarps;
    if (kind == 0) return per_row;
    int per_sm = kind == 1 ? kPersistBlocksPerSM : 0;
    if (const char *e = getenv(kind == 1 ? "DSURF_LSMR_GRID_ROWS" : "DSURF_LSMR_GRID_COLS")) per_sm = atoi(e);
    if (kind == 1 ? getenv("DSURF_LSMR_OLD_ROWS") != nullptr : getenv("DSURF_LSMR_NEW_COLS") == nullptr) per_sm = 0;
    if (per_sm <= 0) return per_row;
    return std::min(per_row, sm_count() * per_sm);
  }
  int blocks() const { return long_grid() + (nshort + 255) / 256; }
};

struct dsurf_lsmr_sys {
  int m = 0, n = 0;
  long long nnz = 0;
  cudaStream_t st = nullptr;
  bool own_stream = false;
  Compressed A, At;  // A: rows of A (u += A v); At: columns of A (v += A'u)
  int P = 0, K = 0;  // depth-blocked layout (P vertices x K depths) when K > 0
  int n_int = 0;     // length of the n-vectors in the internal layout (P*8 when blocked)
  DevBuf<float> xout;
  DevBuf<float> b, u, v, h, hbar, x, localV;
  DevBuf<double> partial, partial2;
  DevBuf<LsmrScalars> S;
  int np_cap = 0;
  // multi-GPU (rows partitioned over ranks): NCCL communicator owned by the host side (dist.cu)
  void *comm = nullptr;
  int rank = 0, nranks = 1;
  DevBuf<float> vpart;
  DevBuf<double> red;
  // one-shot exchange over peer memory (NVLink): see k_xchg_signal / k_xchg_reduce
  char *xbuf = nullptr;             // this rank's exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle)
  size_t xstride = 0, xbytes = 0;   // bytes of one parity slot / of the whole buffer
  DevBuf<char *> xpeers;            // device table: exchange buffers of every rank (this rank's own at [rank])
  bool xready = false;
  std::vector<char *> xmapped;      // peers' buffers opened with cudaIpcOpenMemHandle (closed with the system)
  bool solved = false;
  // fused small-vector phases (k_fused_beta / k_fused_tail): cluster size (0 = unfused path),
  // elements of the n-vectors per CTA, dynamic shared memory of the tail kernel
  int fused_cl = -1, fused_chunk = 0;
  size_t fused_smem = 0;
  cudaStream_t st2 = nullptr;  // side stream: short-row products run beside the long-row kernel
  cudaEvent_t evf = nullptr, evj = nullptr;
  ~dsurf_lsmr_sys() {
    if (evf) cudaEventDestroy(evf);
    if (evj) cudaEventDestroy(evj);
    if (st2) cudaStreamDestroy(st2);
    if (own_stream && st) cudaStreamDestroy(st);
    for (char *p : xmapped) cudaIpcCloseMemHandle(p);
    if (xbuf) cudaFree(xbuf);
  }
};

namespace dsurf {
int lsmr_allreduce(void *comm, float *buf, size_t n, double *dbuf, size_t nd, cudaStream_t st);  // dist.cu

// fork/join of the side stream (works inside stream capture: becomes two parallel graph branches)
static cudaStream_t fork_side(dsurf_lsmr_sys *s, cudaStream_t st) {
  if (!s->st2 || getenv("DSURF_LSMR_NO_FORK") != nullptr) return st;
  if (cudaEventRecord(s->evf, st) != cudaSuccess || cudaStreamWaitEvent(s->st2, s->evf, 0) != cudaSuccess) {
    cudaGetLastError();
    return st;
  }
  return s->st2;
}
static void join_side(dsurf_lsmr_sys *s, cudaStream_t st, cudaStream_t side) {
  cudaEventRecord(s->evj, side);
  cudaStreamWaitEvent(st, s->evj, 0);
}

static int classify_rows(cudaStream_t st, Compressed &C, int nrows) {
  std::vector<long long> hp((size_t)nrows + 1);
  DS_CUDA(cudaMemcpyAsync(hp.data(), C.ptr.p, ((size_t)nrows + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  std::vector<int> lo, sh;
  for (int r = 0; r < nrows; r++) {
    const long long len = hp[r + 1] - hp[r];
    if (len >= kShortRow)
      lo.push_back(r);
    else
      sh.push_back(r);  // includes empty rows (they still need y = s*y)
  }
  // longest first: the persistent kernels stride over this list, so every warp gets the same mix of lengths
  std::stable_sort(lo.begin(), lo.end(), [&](int a, int b) { return hp[a + 1] - hp[a] > hp[b + 1] - hp[b]; });
  C.nrows = nrows;
  C.nlong = (int)lo.size();
  C.nshort = (int)sh.size();
  if (C.longl.reserve(std::max<size_t>(1, lo.size())) || C.shortl.reserve(std::max<size_t>(1, sh.size()))) return DSURF_ERR_CUDA;
  if (!lo.empty()) DS_CUDA(cudaMemcpyAsync(C.longl.p, lo.data(), lo.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!sh.empty()) DS_CUDA(cudaMemcpyAsync(C.shortl.p, sh.data(), sh.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  DS_CUDA(cudaStreamSynchronize(st));
  return DSURF_OK;
}

// geometry hint for the depth-blocked layout: set by dsurf_plan_create / dsurf_lsmr_hint_geometry
static int g_hint_P = 0, g_hint_K = 0;

// Builds the depth-blocked structure of A (by_rows) or A' (vertex columns) from device COO.
static int build_blocked(cudaStream_t st, const int *rows1, const int *cols1, const float *vals, long long nnz,
                         int m, int P, bool by_rows, Compressed &C) {
  const int nouter = by_rows ? m : P;
  DevBuf<unsigned long long> k_in, k_out;
  DevBuf<int> p_in, p_out, head, cnt;
  DevBuf<long long> bid, wide, scan;
  DevBuf<char> tmp;
  const size_t nn = nnz > 0 ? (size_t)nnz : 1;
  if (k_in.reserve(nn) || k_out.reserve(nn) || p_in.reserve(nn) || p_out.reserve(nn) || head.reserve(nn) ||
      bid.reserve(nn) || cnt.reserve(nouter + 1) || wide.reserve(nouter + 1) || scan.reserve(nouter + 1) ||
      C.ptr.reserve(nouter + 1)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (blocked build)");
    return DSURF_ERR_CUDA;
  }
  k_blk_keys<<<grid_for(nnz), 256, 0, st>>>(rows1, cols1, nnz, P, (long long)m, by_rows ? 1 : 0, k_in.p, p_in.p);
  int end_bit = 1;
  const unsigned long long maxkey = (unsigned long long)(by_rows ? (long long)m * P : (long long)P * m);
  while ((1ull << end_bit) <= maxkey && end_bit < 63) end_bit++;
  size_t sb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sb, k_in.p, k_out.p, p_in.p, p_out.p, nnz, 0, end_bit, st);
  if (tmp.reserve(sb + 16)) return DSURF_ERR_CUDA;
  cub::DeviceRadixSort::SortPairs(tmp.p, sb, k_in.p, k_out.p, p_in.p, p_out.p, nnz, 0, end_bit, st);
  k_blk_heads<<<grid_for(nnz), 256, 0, st>>>(k_out.p, nnz, head.p);
  size_t tb = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tb, head.p, bid.p, nnz, st);
  if (tmp.reserve(tb + 16)) return DSURF_ERR_CUDA;
  cub::DeviceScan::InclusiveSum(tmp.p, tb, head.p, bid.p, nnz, st);
  long long nblk = 0;
  if (nnz > 0) DS_CUDA(cudaMemcpyAsync(&nblk, bid.p + (nnz - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  C.nblk = nblk;
  const size_t nb = nblk > 0 ? (size_t)nblk : 1;
  if (C.idx.reserve(nb) || C.val.reserve(nb * 8)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (blocks)");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cudaMemsetAsync(C.val.p, 0, nb * 8 * sizeof(float), st));
  DS_CUDA(cudaMemsetAsync(cnt.p, 0, (nouter + 1) * sizeof(int), st));
  if (nnz > 0)
    k_blk_fill<<<grid_for(nnz), 256, 0, st>>>(k_out.p, p_out.p, bid.p, cols1, vals, nnz, P, (long long)m,
                                             by_rows ? 1 : 0, C.idx.p, C.val.p, cnt.p);
  k_widen<<<grid_for(nouter + 1), 256, 0, st>>>(cnt.p, wide.p, nouter + 1);
  cub::DeviceScan::ExclusiveSum(nullptr, tb, wide.p, scan.p, nouter + 1, st);
  if (tmp.reserve(tb + 16)) return DSURF_ERR_CUDA;
  cub::DeviceScan::ExclusiveSum(tmp.p, tb, wide.p, scan.p, nouter + 1, st);
  k_to_ptr<<<grid_for(nouter + 1), 256, 0, st>>>(cnt.p, scan.p, nouter, nblk, C.ptr.p);
  DS_CUDA(cudaStreamSynchronize(st));
  DS_CUDA(cudaGetLastError());
  C.kind = by_rows ? 1 : 2;
  return DSURF_OK;
}

// y = s*y + C x (all rows); partial (may be null) receives C.blocks() block sums of y^2
// `fork` (may be null): system whose side stream runs the short-row kernel beside the long-row one
static void launch_product(cudaStream_t st, const Compressed &C, const float *x, float *y, const float *scale_ptr,
                           float sign, double *partial, const int *stop, dsurf_lsmr_sys *fork = nullptr) {
  const int ga = C.long_grid();
  // Measured on B200 (gpurun_out/s13_*.json, probe system of cfg-3 shape, us per product):
  //   rows: first-generation warp-per-row 118.5 | persistent + 4 blocks in flight, 128-bit loads 116.2 | 256-bit 112.2
  //   cols: first-generation warp-per-column (longest first) 115.5 | 4 blocks in flight, 128-bit 126.3 | 256-bit 167.3
  // so rows use k_bspmv_rows_p<256-bit> and columns keep k_bspmv_cols.  A/B knobs: DSURF_LSMR_OLD_ROWS=1,
  // DSURF_LSMR_NEW_COLS=1, DSURF_LSMR_W128=1.
  static const bool old_rows = getenv("DSURF_LSMR_OLD_ROWS") != nullptr, old_cols = getenv("DSURF_LSMR_NEW_COLS") == nullptr;
  static const bool w256 = getenv("DSURF_LSMR_W128") == nullptr;
  const bool persist = (C.kind == 1 && !old_rows) || (C.kind == 2 && !old_cols);
  cudaStream_t ss = (fork && C.nlong > 0 && C.nshort > 0) ? fork_side(fork, st) : st;
  struct Join {
    dsurf_lsmr_sys *f;
    cudaStream_t st, ss;
    ~Join() {
      if (ss != st) join_side(f, st, ss);
    }
  } join{fork, st, ss};
  if (C.kind == 1 || C.kind == 2) {
    const float4 *v4 = reinterpret_cast<const float4 *>(C.val.p);
    const int gb = (C.nshort + 255) / 256;
    if (C.kind == 1) {
      const float4 *x4 = reinterpret_cast<const float4 *>(x);
      if (C.nlong > 0 && persist)
        (w256 ? k_bspmv_rows_p<true> : k_bspmv_rows_p<false>)<<<ga, kSpmvWarps * 32, 0, st>>>(
            C.ptr.p, C.idx.p, v4, x4, y, scale_ptr, sign, C.longl.p, C.nlong, partial, stop);
      else if (C.nlong > 0)
        k_bspmv_rows<<<ga, kSpmvWarps * 32, 0, st>>>(C.ptr.p, C.idx.p, v4, x4, y, scale_ptr, sign, C.longl.p, C.nlong, 0,
                                                     partial, stop);
      if (C.nshort > 0)
        k_bspmv_rows<<<gb, 256, 0, ss>>>(C.ptr.p, C.idx.p, v4, x4, y, scale_ptr, sign, C.shortl.p, C.nshort, 1,
                                         partial ? partial + ga : nullptr, stop);
    } else {
      if (C.nlong > 0 && persist)
        (w256 ? k_bspmv_cols_p<true> : k_bspmv_cols_p<false>)<<<ga, kSpmvWarps * 32, 0, st>>>(
            C.ptr.p, C.idx.p, v4, x, y, scale_ptr, sign, C.longl.p, C.nlong, partial, stop);
      else if (C.nlong > 0)
        k_bspmv_cols<<<ga, kSpmvWarps * 32, 0, st>>>(C.ptr.p, C.idx.p, v4, x, y, scale_ptr, sign, C.longl.p, C.nlong, 0,
                                                     partial, stop);
      if (C.nshort > 0)
        k_bspmv_cols<<<gb, 256, 0, ss>>>(C.ptr.p, C.idx.p, v4, x, y, scale_ptr, sign, C.shortl.p, C.nshort, 1,
                                         partial ? partial + ga : nullptr, stop);
    }
    return;
  }
  if (C.nlong > 0)
    k_spmv_warp<<<ga, kSpmvWarps * 32, 0, st>>>(C.ptr.p, C.idx.p, C.val.p, x, y, scale_ptr, sign, C.longl.p, C.nlong,
                                                partial, stop);
  if (C.nshort > 0)
    k_spmv_short<<<(C.nshort + 255) / 256, 256, 0, ss>>>(C.ptr.p, C.idx.p, C.val.p, x, y, scale_ptr, sign, C.shortl.p,
                                                         C.nshort, partial ? partial + ga : nullptr, stop);
}

int lsmr_sys_create_dev(dsurf_lsmr_sys **out, int m, int n, long long nnz, const int *d_rows1,
                        const int *d_cols1, const float *d_vals, const float *d_b) {
  auto *s = new dsurf_lsmr_sys();
  s->m = m;
  s->n = n;
  s->nnz = nnz;
  if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) {
    set_error(__FILE__, __LINE__, "cudaStreamCreate failed");
    delete s;
    return DSURF_ERR_CUDA;
  }
  s->own_stream = true;
  if (cudaStreamCreateWithFlags(&s->st2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->evf, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->evj, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    s->st2 = nullptr;  // no side stream: short-row products stay on the main stream
  }
  cudaStream_t st = s->st;
  cudaDeviceSynchronize();  // inputs were uploaded on the legacy default stream
  int rc = DSURF_OK;
  const bool blocked = g_hint_K >= 2 && g_hint_K <= 8 && (long long)g_hint_P * g_hint_K == n &&
                       getenv("DSURF_LSMR_NO_BLOCK") == nullptr;
  s->n_int = n;
  if (blocked) {
    s->P = g_hint_P;
    s->K = g_hint_K;
    s->n_int = g_hint_P * 8;
    rc = build_blocked(st, d_rows1, d_cols1, d_vals, nnz, m, s->P, true, s->A);
    if (rc == DSURF_OK) rc = build_blocked(st, d_rows1, d_cols1, d_vals, nnz, m, s->P, false, s->At);
    if (rc == DSURF_OK) rc = classify_rows(st, s->A, m);
    if (rc == DSURF_OK) rc = classify_rows(st, s->At, s->P);
  } else {
    rc = build_compressed(st, d_rows1, d_cols1, d_vals, nnz, m, s->A.ptr, s->A.idx, s->A.val);
    if (rc == DSURF_OK) rc = build_compressed(st, d_cols1, d_rows1, d_vals, nnz, n, s->At.ptr, s->At.idx, s->At.val);
    if (rc == DSURF_OK) rc = classify_rows(st, s->A, m);
    if (rc == DSURF_OK) rc = classify_rows(st, s->At, n);
  }
  if (rc != DSURF_OK) {
    delete s;
    return rc;
  }
  const int np = std::max(std::max(s->A.blocks(), s->At.blocks()), 4096) + 16;
  s->np_cap = np;
  const size_t ni = (size_t)s->n_int;
  if (s->b.reserve(m) || s->u.reserve(m) || s->v.reserve(ni) || s->h.reserve(ni) || s->hbar.reserve(ni) ||
      s->x.reserve(ni) || s->xout.reserve(ni) || s->partial.reserve(np) || s->partial2.reserve(np) ||
      s->S.reserve(1) || s->vpart.reserve(ni + 8) || s->red.reserve(8)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (lsmr vectors)");
    delete s;
    return DSURF_ERR_CUDA;
  }
  cudaMemcpyAsync(s->b.p, d_b, (size_t)m * sizeof(float), cudaMemcpyDeviceToDevice, st);
  cudaStreamSynchronize(st);
  *out = s;
  return DSURF_OK;
}
}  // namespace dsurf

namespace dsurf {
float *lsmr_x_dev(dsurf_lsmr_sys *s) { return (s && s->solved) ? s->xout.p : nullptr; }
}  // namespace dsurf

extern "C" int dsurf_lsmr_create(dsurf_lsmr_sys **sys, int m, int n, int64_t nar, const int *rows1,
                                 const int *cols1, const float *vals, const float *b) {
  DS_CHECK(ensure_device());
  if (!sys || m <= 0 || n <= 0 || nar < 0) return DSURF_ERR_BAD_ARG;
  DevBuf<int> dr, dc;
  DevBuf<float> dv, db;
  const size_t nn = nar > 0 ? (size_t)nar : 1;
  if (dr.reserve(nn) || dc.reserve(nn) || dv.reserve(nn) || db.reserve(m)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (COO upload)");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cudaMemcpy(dr.p, rows1, nar * sizeof(int), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dc.p, cols1, nar * sizeof(int), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dv.p, vals, nar * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(db.p, b, (size_t)m * sizeof(float), cudaMemcpyHostToDevice));
  return lsmr_sys_create_dev(sys, m, n, nar, dr.p, dc.p, dv.p, db.p);
}

// Tells the solver that columns are ordered k*P + vertex with P = (nx-2)(ny-2) vertices and
// K = nz-1 depths (main.f90:289, CalSurfG.f90:1400), which enables the depth-blocked layout for
// systems with n == P*K.  Called by dsurf_plan_create (i.e. by every CalSurfG call).
extern "C" int dsurf_lsmr_hint_geometry(int nx, int ny, int nz) {
  g_hint_P = (nx - 2) * (ny - 2);
  g_hint_K = nz - 1;
  return DSURF_OK;
}

extern "C" int dsurf_lsmr_destroy(dsurf_lsmr_sys *sys) {
  delete sys;
  return DSURF_OK;
}
extern "C" int64_t dsurf_lsmr_nnz(const dsurf_lsmr_sys *sys) { return sys ? sys->nnz : 0; }
extern "C" int dsurf_lsmr_fused_cluster(const dsurf_lsmr_sys *sys) { return sys ? sys->fused_cl : 0; }
extern "C" int dsurf_lsmr_set_comm(dsurf_lsmr_sys *sys, void *comm, int rank, int nranks) {
  if (!sys) return DSURF_ERR_BAD_ARG;
  sys->comm = comm;
  sys->rank = rank;
  sys->nranks = nranks;
  return DSURF_OK;
}

// ---- peer-memory exchange set-up (see k_xchg_signal): export this rank's buffer, then map everybody's
extern "C" int dsurf_lsmr_xchg_export(dsurf_lsmr_sys *s, void *handle64) {
  if (!s || !handle64) return DSURF_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!s->xbuf) {
    s->xstride = (((size_t)s->n_int * sizeof(float) + sizeof(double)) + 255) & ~(size_t)255;
    s->xbytes = 2 * s->xstride + sizeof(XchgHdr);
    DS_CUDA(cudaMalloc(&s->xbuf, s->xbytes));
    DS_CUDA(cudaMemset(s->xbuf, 0, s->xbytes));
  }
  cudaIpcMemHandle_t h;
  DS_CUDA(cudaIpcGetMemHandle(&h, s->xbuf));
  memcpy(handle64, &h, 64);
  return DSURF_OK;
}
extern "C" int dsurf_lsmr_xchg_attach(dsurf_lsmr_sys *s, const void *handles, int rank, int nranks) {
  if (!s || !handles || !s->xbuf || rank < 0 || rank >= nranks || nranks > 64) return DSURF_ERR_BAD_ARG;
  for (char *p : s->xmapped) cudaIpcCloseMemHandle(p);  // attached before: drop the old mappings
  s->xmapped.clear();
  s->xready = false;
  std::vector<char *> tab(nranks, nullptr);
  for (int r = 0; r < nranks; r++) {
    if (r == rank) {
      tab[r] = s->xbuf;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * (size_t)r, 64);
    void *ptr = nullptr;
    DS_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    tab[r] = (char *)ptr;
    s->xmapped.push_back((char *)ptr);
  }
  if (s->xpeers.reserve(nranks)) return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemcpy(s->xpeers.p, tab.data(), nranks * sizeof(char *), cudaMemcpyHostToDevice));
  s->rank = rank;
  s->nranks = nranks;
  s->xready = getenv("DSURF_LSMR_NCCL_ONLY") == nullptr;
  return DSURF_OK;
}

// Chooses the cluster size of the fused small-vector kernels: the largest of 16 (non-portable) / 8
// whose per-CTA slice of v fits in shared memory and that the device can co-schedule; 0 = unfused.
static void choose_fused(dsurf_lsmr_sys *s) {
  if (s->fused_cl >= 0) return;
  s->fused_cl = 0;
  if (getenv("DSURF_LSMR_NO_FUSE") != nullptr) return;
  for (int cl : {16, 8}) {
    const int chunk = (s->n_int + cl - 1) / cl;
    const size_t smem = chunk <= kEPT * kFusedThreads ? 0 : (size_t)chunk * sizeof(float);  // register path needs none
    if (smem > 200 * 1024) continue;
    if (cl > 8 && (cudaFuncSetAttribute(k_fused_tail, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                   cudaFuncSetAttribute(k_fused_beta, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
      cudaGetLastError();
      continue;
    }
    if (cudaFuncSetAttribute(k_fused_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl);
    cfg.blockDim = dim3(kFusedThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, k_fused_tail, &cfg) != cudaSuccess || ncl < 1) {
      cudaGetLastError();
      continue;
    }
    s->fused_cl = cl;
    s->fused_chunk = chunk;
    s->fused_smem = smem;
    return;
  }
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kern)(KArgs...), int cl, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// one LSMR iteration's kernel sequence (lsmrModule.f90:475-616) on stream st
static int enqueue_iteration(dsurf_lsmr_sys *s, float damp, float atol, float btol, float ctol, int itnlim,
                             int force_iters, int localVecs, bool dist, int parity = 0) {
  cudaStream_t st = s->st;
  const int m = s->m, n = s->n_int;  // n-vectors in the internal layout
  LsmrScalars *S = s->S.p;
  double *part = s->partial.p, *part2 = s->partial2.p;
  const int gvec = std::min(std::max((n + 255) / 256, 1), sm_count() * 4);
  const int gvecm = std::min(std::max((std::max(m, n) + 255) / 256, 1), sm_count() * 4);
  const int cl = s->fused_cl;
  // u = A v - alpha u ; beta (:484-487)
  launch_product(st, s->A, s->v.p, s->u.p, &S->alpha, -1.0f, part, &S->stop, s);
  if (!dist) {
    if (cl > 0) {
      DS_CUDA(launch_cluster(k_fused_beta, cl, 0, st, S, (const double *)part, s->A.blocks(), (const double *)nullptr,
                             localVecs, s->u.p, m, (const float *)s->v.p, s->localV.p, n));
    } else {
      k_beta<<<1, 1024, 0, st>>>(S, part, s->A.blocks(), nullptr, localVecs);
      k_scale_u_enqueue<<<gvecm, 256, 0, st>>>(S, s->u.p, m, s->v.p, s->localV.p, n, localVecs > 0 ? 1 : 0);
    }
    // v = A'u - beta v (:495-497)
    launch_product(st, s->At, s->u.p, s->v.p, &S->neg_beta, 1.0f, nullptr, &S->stop, s);
  } else if (s->xready) {
    // peer-memory exchange: local A'u straight into this rank's slot, signal, wait + sum over ranks
    float *slot = reinterpret_cast<float *>(s->xbuf + (size_t)parity * s->xstride);
    launch_product(st, s->At, s->u.p, slot, nullptr, 0.0f, nullptr, &S->stop, nullptr);
    k_xchg_signal<<<1, 1024, 0, st>>>(S, s->xpeers.p, s->rank, s->nranks, s->xstride, part, s->A.blocks());
    k_xchg_reduce<<<gvec, 256, 0, st>>>(S, s->xpeers.p, s->rank, s->nranks, s->xstride, n, s->vpart.p, s->red.p);
    if (cl > 0) {
      DS_CUDA(launch_cluster(k_fused_beta, cl, 0, st, S, (const double *)part, s->A.blocks(), (const double *)s->red.p,
                             localVecs, s->u.p, m, (const float *)s->v.p, s->localV.p, n));
    } else {
      k_beta<<<1, 1024, 0, st>>>(S, part, s->A.blocks(), s->red.p, localVecs);
      k_scale_u_enqueue<<<gvecm, 256, 0, st>>>(S, s->u.p, m, s->v.p, s->localV.p, n, localVecs > 0 ? 1 : 0);
    }
    k_combine_v<<<gvec, 256, 0, st>>>(S, s->v.p, s->vpart.p, n);
  } else {
    // one fused exchange per iteration: partial A'u_raw (n floats) + partial ||u||^2 (1 double)
    k_reduce_to<<<1, 1024, 0, st>>>(s->red.p, part, s->A.blocks());
    launch_product(st, s->At, s->u.p, s->vpart.p, nullptr, 0.0f, nullptr, &S->stop, nullptr);
    DS_CHECK(lsmr_allreduce(s->comm, s->vpart.p, n, s->red.p, 1, st));
    if (cl > 0) {
      DS_CUDA(launch_cluster(k_fused_beta, cl, 0, st, S, (const double *)part, s->A.blocks(), (const double *)s->red.p,
                             localVecs, s->u.p, m, (const float *)s->v.p, s->localV.p, n));
    } else {
      k_beta<<<1, 1024, 0, st>>>(S, part, s->A.blocks(), s->red.p, localVecs);
      k_scale_u_enqueue<<<gvecm, 256, 0, st>>>(S, s->u.p, m, s->v.p, s->localV.p, n, localVecs > 0 ? 1 : 0);
    }
    k_combine_v<<<gvec, 256, 0, st>>>(S, s->v.p, s->vpart.p, n);
  }
  if (cl > 0) {
    // local reorthogonalisation + alpha + rotations + updates + tests in one cluster launch
    DS_CUDA(launch_cluster(k_fused_tail, cl, s->fused_smem, st, S, s->v.p, (const float *)s->localV.p, s->h.p,
                           s->hbar.p, s->x.p, n, s->fused_chunk, damp, atol, btol, ctol, itnlim, force_iters));
    return DSURF_OK;
  }
  // local reorthogonalisation + alpha (:498-504, 731-748); steps beyond the current limit no-op
  for (int c = 0; c <= localVecs; c++)
    k_reorth<<<gvec, 256, 0, st>>>(S, s->v.p, s->localV.p, n, c, (c & 1) ? part : part2, (c & 1) ? part2 : part, gvec);
  k_rotations<<<1, 256, 0, st>>>(S, damp, part, part2, gvec);
  k_update<<<gvec, 256, 0, st>>>(S, s->v.p, s->h.p, s->hbar.p, s->x.p, n, part);
  k_tests<<<1, 1024, 0, st>>>(S, part, gvec, atol, btol, ctol, itnlim, force_iters);
  return DSURF_OK;
}

extern "C" int dsurf_lsmr_solve(dsurf_lsmr_sys *s, float damp, float atol, float btol, float conlim,
                                int itnlim, int localSize, int force_iters, float *x_host, int *istop,
                                int *itn, float *normA, float *condA, float *normr, float *normAr,
                                float *normx, double *ms_total, double *ms_spmv, double *ms_spmtv) {
  if (!s) return DSURF_ERR_BAD_ARG;
  DS_CHECK(ensure_device());
  cudaStream_t st = s->st;
  const int m = s->m, n = s->n_int;  // n-vectors in the internal layout
  const int localVecs = std::min(localSize, std::min(m, s->n));
  if (localVecs > 0 && s->localV.reserve((size_t)n * localVecs)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (localV)");
    return DSURF_ERR_CUDA;
  }
  LsmrScalars *S = s->S.p;
  double *part = s->partial.p;
  const int gvec = std::min(std::max((n + 255) / 256, 1), sm_count() * 4);
  const int gvecm = std::min(std::max((std::max(m, n) + 255) / 256, 1), sm_count() * 4);
  const float ctol = conlim > 0.0f ? 1.0f / conlim : 0.0f;
  const bool dist = s->comm != nullptr && s->nranks > 1;
  choose_fused(s);
  // events and the iteration graph are released on every return path, error paths included
  struct Guard {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    ~Guard() {
      if (gexec) cudaGraphExecDestroy(gexec);
      if (graph) cudaGraphDestroy(graph);
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    }
  } guard;
  cudaEventCreate(&guard.e0);
  cudaEventCreate(&guard.e1);
  cudaEvent_t e0 = guard.e0, e1 = guard.e1;
  DS_CUDA(cudaMemsetAsync(S, 0, sizeof(LsmrScalars), st));
  DS_CUDA(cudaMemsetAsync(s->x.p, 0, (size_t)n * sizeof(float), st));
  DS_CUDA(cudaMemsetAsync(s->hbar.p, 0, (size_t)n * sizeof(float), st));
  DS_CUDA(cudaMemsetAsync(s->v.p, 0, (size_t)n * sizeof(float), st));
  cudaEventRecord(e0, st);
  // ---- u = b ; beta = ||u|| ; u /= beta ; v = A'u ; alpha = ||v|| ; v /= alpha (:380-397)
  k_sumsq<<<gvecm, 256, 0, st>>>(s->b.p, m, part);
  if (dist) {
    k_reduce_to<<<1, 1024, 0, st>>>(s->red.p, part, gvecm);
    DS_CHECK(lsmr_allreduce(s->comm, nullptr, 0, s->red.p, 1, st));
  }
  k_init_beta<<<1, 1024, 0, st>>>(S, part, gvecm, dist ? s->red.p : nullptr);
  k_init_scale_copy<<<gvecm, 256, 0, st>>>(s->u.p, s->b.p, nullptr, &S->inv_beta, m);
  if (!dist) {
    launch_product(st, s->At, s->u.p, s->v.p, nullptr, 0.0f, part, nullptr);
    k_init_alpha<<<1, 1024, 0, st>>>(S, part, s->At.blocks(), localVecs);
  } else {
    launch_product(st, s->At, s->u.p, s->v.p, nullptr, 0.0f, nullptr, nullptr);
    DS_CHECK(lsmr_allreduce(s->comm, s->v.p, n, nullptr, 0, st));
    k_sumsq<<<gvec, 256, 0, st>>>(s->v.p, n, part);
    k_init_alpha<<<1, 1024, 0, st>>>(S, part, gvec, localVecs);
  }
  k_init_scale_copy<<<gvec, 256, 0, st>>>(s->v.p, s->v.p, s->h.p, &S->inv_alpha, n);
  if (localVecs > 0)
    DS_CUDA(cudaMemcpyAsync(s->localV.p, s->v.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  int h_stop = 0;
  DS_CUDA(cudaMemcpyAsync(&h_stop, &S->stop, sizeof(int), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  DS_CUDA(cudaGetLastError());
  // ---- main loop.  Single GPU: the iteration is captured once into a CUDA graph and replayed;
  // every kernel no-ops once the device-side stop flag is set, so the host only polls the flag
  // every kPoll iterations (the solution is frozen at the exact stopping iteration).
  cudaGraph_t &graph = guard.graph;
  cudaGraphExec_t &gexec = guard.gexec;
  const bool xchg = dist && s->xready;
  int par0 = 0;  // slot parity of the next iteration = parity of (epoch + 1), identical on every rank
  if (xchg) {
    int e0 = 0;
    DS_CUDA(cudaMemcpy(&e0, &xchg_hdr(s->xbuf, s->xstride)->epoch, sizeof(int), cudaMemcpyDeviceToHost));
    par0 = (e0 + 1) & 1;
  }
  const bool use_graph = (!dist || xchg) && getenv("DSURF_LSMR_NO_GRAPH") == nullptr;
  if (use_graph && !h_stop) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rc = enqueue_iteration(s, damp, atol, btol, ctol, itnlim, force_iters, localVecs, dist, par0);
      if (xchg && rc == DSURF_OK)  // the exchange alternates between two slots: one graph = two iterations
        rc = enqueue_iteration(s, damp, atol, btol, ctol, itnlim, force_iters, localVecs, dist, par0 ^ 1);
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc != DSURF_OK || ce != cudaSuccess || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        graph = nullptr;
        gexec = nullptr;
        cudaGetLastError();
      }
    } else {
      cudaGetLastError();
    }
  }
  const int kPoll = gexec ? 4 : 1;
  int launched = 0;
  while (!h_stop) {
    for (int k = 0; k < kPoll; k++) {
      if (gexec) {
        DS_CUDA(cudaGraphLaunch(gexec, st));
      } else {
        DS_CHECK(enqueue_iteration(s, damp, atol, btol, ctol, itnlim, force_iters, localVecs, dist, (par0 + launched) & 1));
      }
      launched++;
    }
    DS_CUDA(cudaMemcpyAsync(&h_stop, &S->stop, sizeof(int), cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    if (launched > itnlim + 8) break;
  }
  cudaEventRecord(e1, st);
  if (xchg) {
    int xerr = 0;
    DS_CUDA(cudaMemcpyAsync(&xerr, &xchg_hdr(s->xbuf, s->xstride)->error, sizeof(int), cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    if (xerr) {
      set_error(__FILE__, __LINE__, "distributed LSMR: a peer never signalled its partial A'u (peer-memory exchange timed out)");
      return DSURF_ERR_NCCL;
    }
  }
  LsmrScalars hs;
  DS_CUDA(cudaMemcpyAsync(&hs, S, sizeof(hs), cudaMemcpyDeviceToHost, st));
  // x in the reference's column order k*P + pos stays in xout (read by the device-side model update)
  if (s->K > 0)
    k_unpermute<<<(s->n + 255) / 256, 256, 0, st>>>(s->x.p, s->xout.p, s->P, s->K);
  else
    DS_CUDA(cudaMemcpyAsync(s->xout.p, s->x.p, (size_t)s->n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  s->solved = true;
  if (x_host) DS_CUDA(cudaMemcpyAsync(x_host, s->xout.p, (size_t)s->n * sizeof(float), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  DS_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_total) *ms_total = ms;
  // per-product timings (bench / roofline): 10 stand-alone launches each, after the solve
  if (ms_spmv || ms_spmtv) {
    const int reps = 10;
    for (int which = 0; which < 2; which++) {
      double *dst = which == 0 ? ms_spmv : ms_spmtv;
      if (!dst) continue;
      float *ybuf = which == 0 ? s->u.p : s->vpart.p;  // scratch outputs (solve is finished)
      cudaEventRecord(e0, st);
      for (int r = 0; r < reps; r++) {
        if (which == 0)
          launch_product(st, s->A, s->v.p, ybuf, nullptr, 0.0f, nullptr, nullptr);
        else
          launch_product(st, s->At, s->u.p, ybuf, nullptr, 0.0f, nullptr, nullptr);
      }
      cudaEventRecord(e1, st);
      DS_CUDA(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&ms, e0, e1);
      *dst = (double)ms / reps * (double)hs.itn;  // scaled to the iteration count like before
    }
  }
  int is = hs.istop;
  if (damp > 0.0f && is == 2) is = 3;  // lsmrModule.f90:654
  if (istop) *istop = is;
  if (itn) *itn = hs.itn;
  if (normA) *normA = hs.normA;
  if (condA) *condA = hs.condA;
  if (normr) *normr = hs.normr;
  if (normAr) *normAr = hs.normAr;
  if (normx) *normx = hs.normx;
  return DSURF_OK;
}

// aprod drop-in: one product on the device (mode 1: y += A x, mode 2: x += A'y); the COO is
// compressed on the fly, so this entry point exists for interface completeness and tests --
// LSMR itself never calls it.
extern "C" int dsurf_aprod(int mode, int m, int n, float *x, float *y, int leniw, int lenrw,
                           const int *iw, const float *rw) {
  DS_CHECK(ensure_device());
  if (!iw || !rw || leniw < 1) return DSURF_ERR_BAD_ARG;
  const long long nar = iw[0];
  if (leniw < 2 * nar + 1 || lenrw < nar) return DSURF_ERR_BAD_ARG;
  DevBuf<int> dr, dc;
  DevBuf<float> dv, dx, dy;
  Compressed C;
  const size_t nn = nar > 0 ? nar : 1;
  if (dr.reserve(nn) || dc.reserve(nn) || dv.reserve(nn) || dx.reserve(n) || dy.reserve(m)) return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemcpy(dr.p, iw + 1, nar * sizeof(int), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dc.p, iw + 1 + nar, nar * sizeof(int), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dv.p, rw, nar * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dx.p, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dy.p, y, (size_t)m * sizeof(float), cudaMemcpyHostToDevice));
  DevBuf<float> one;
  if (one.reserve(1)) return DSURF_ERR_CUDA;
  const float h1 = 1.0f;
  DS_CUDA(cudaMemcpy(one.p, &h1, sizeof(float), cudaMemcpyHostToDevice));
  if (mode == 1) {
    DS_CHECK(build_compressed(0, dr.p, dc.p, dv.p, nar, m, C.ptr, C.idx, C.val));
    DS_CHECK(classify_rows(0, C, m));
    launch_product(0, C, dx.p, dy.p, one.p, 1.0f, nullptr, nullptr);
    DS_CUDA(cudaMemcpy(y, dy.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost));
  } else {
    DS_CHECK(build_compressed(0, dc.p, dr.p, dv.p, nar, n, C.ptr, C.idx, C.val));
    DS_CHECK(classify_rows(0, C, n));
    launch_product(0, C, dy.p, dx.p, one.p, 1.0f, nullptr, nullptr);
    DS_CUDA(cudaMemcpy(x, dx.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  DS_CUDA(cudaGetLastError());
  return DSURF_OK;
}

extern "C" int dsurf_lsmr(int m, int n, int leniw, int lenrw, const int *iw, const float *rw,
                          const float *b, float damp, float atol, float btol, float conlim,
                          int itnlim, int localSize, float *x, int *istop, int *itn, float *normA,
                          float *condA, float *normr, float *normAr, float *normx) {
  if (!iw || !rw || !b || !x || leniw < 1) return DSURF_ERR_BAD_ARG;
  const long long nar = iw[0];
  if (leniw < 2 * nar + 1 || lenrw < nar) return DSURF_ERR_BAD_ARG;
  dsurf_lsmr_sys *sys = nullptr;
  DS_CHECK(dsurf_lsmr_create(&sys, m, n, nar, iw + 1, iw + 1 + nar, rw, b));
  int rc = dsurf_lsmr_solve(sys, damp, atol, btol, conlim, itnlim, localSize, 0, x, istop, itn, normA,
                            condA, normr, normAr, normx, nullptr, nullptr, nullptr);
  dsurf_lsmr_destroy(sys);
  return rc;
}
