// Device-resident host glue between CalSurfG and LSMR, replacing the host loops of the
// reference's main program (src/main.f90:361-466 residuals, getpercentile outlier weights, row
// scaling, DWS statistics, smoothing rows, iw packing; :518-532 model update) and
// src/getpercentile.f90.  SURVEY.md section 8(f) row 1: with this stage the sparse matrix never
// leaves HBM between the ray kernels and the LSMR products.
//
// Arithmetic is REAL*4 as in the reference except the DWS column sums (log-file statistics only),
// which are accumulated in fp64 so that the atomics' order cannot change the printed value.
#include <cub/cub.cuh>
#include <vector>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "glue.cuh"

namespace dsurf {

static int gridn(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// cbst(i) = obst(i) - dsyn(i)   (main.f90:361-363)
__global__ void k_residual(const float *__restrict__ obst, const float *__restrict__ dsyn, float *__restrict__ cbst,
                           int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cbst[i] = obst[i] - dsyn[i];
}

// datweight / outlier rejection (main.f90:365-372); lo = q25*threshold0, hi = q75*threshold0
__global__ void k_outliers(float *__restrict__ cbst, float *__restrict__ datw, int n, float lo, float hi) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float c = cbst[i];
    const bool out = (c < lo) || (c > hi);
    datw[i] = out ? 0.0f : 1.0f;
    if (out) cbst[i] = 0.0f;
  }
}

// rw(i) = rw(i)*datweight(iw(1+i)) and norm(col(i)) += abs(rw(i))   (main.f90:378-385)
__global__ void k_scale_dws(float *__restrict__ rw, const int *__restrict__ rows1, const int *__restrict__ cols1,
                            const float *__restrict__ datw, long long nar, double *__restrict__ norm) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nar; i += (long long)gridDim.x * blockDim.x) {
    const float v = rw[i] * datw[rows1[i] - 1];
    rw[i] = v;
    if (v != 0.0f) atomicAdd(&norm[cols1[i] - 1], (double)fabsf(v));
  }
}

// maxnorm / averdws (main.f90:386-393): out[0] = max, out[1] = sum
__global__ void k_dws_stats(const double *__restrict__ norm, int n, double *__restrict__ out) {
  __shared__ double smax[32], ssum[32];
  double mx = 0.0, sm = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = norm[i];
    mx = v > mx ? v : mx;
    sm += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double m2 = __shfl_xor_sync(kFull, mx, o);
    mx = m2 > mx ? m2 : mx;
    sm += __shfl_xor_sync(kFull, sm, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = mx;
    ssum[threadIdx.x >> 5] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
      mx = smax[w] > mx ? smax[w] : mx;
      sm += ssum[w];
    }
    out[0] = mx;
    out[1] = sm;
  }
}

// main.f90:518-532: clip dv to +-0.5 (written back), add to the interior nodes, clamp the model
__global__ void k_model_update(float *__restrict__ vels, float *__restrict__ dv, int nx, int ny, int nz, float minvel,
                               float maxvel) {
  const int nvx = nx - 2, nvz = ny - 2;
  const int n = nvx * nvz * (nz - 1);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = t % nvx, j = (t / nvx) % nvz, k = t / (nvx * nvz);
    float d = dv[t];
    if (d >= 0.500f) d = 0.500f;
    if (d <= -0.500f) d = -0.500f;
    dv[t] = d;
    const size_t o = (size_t)k * nx * ny + (size_t)(j + 1) * nx + (i + 1);
    float v = vels[o] + d;
    if (v < minvel) v = minvel;
    if (v > maxvel) v = maxvel;
    vels[o] = v;
  }
}

// getpercentile.f90: RA(int(0.25*N)), RA(int(0.75*N)) of the ascending sort (1-based); the
// reference heap-sorts a copy -- any ascending sort yields the same two values.
int glue_percentiles(cudaStream_t st, const float *d_x, int n, DevBuf<float> &sorted, DevBuf<char> &tmp, float *q25,
                     float *q75) {
  if (n < 2) return DSURF_ERR_BAD_ARG;
  if (sorted.reserve(n)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (percentile sort)");
    return DSURF_ERR_CUDA;
  }
  size_t bytes = 0;
  DS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_x, sorted.p, n, 0, 32, st));
  if (tmp.reserve(bytes)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (percentile sort scratch)");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, d_x, sorted.p, n, 0, 32, st));
  int i25 = (int)(0.25f * (float)n), i75 = (int)(0.75f * (float)n);
  // N < 4 makes the reference read RA(0) (out of bounds); clamp instead of reproducing that
  i25 = i25 < 1 ? 1 : i25;
  i75 = i75 < 1 ? 1 : i75;
  DS_CUDA(cudaMemcpyAsync(q25, sorted.p + (i25 - 1), sizeof(float), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaMemcpyAsync(q75, sorted.p + (i75 - 1), sizeof(float), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  return DSURF_OK;
}

// Smoothing (first-order Laplacian) rows of main.f90:418-459 as COO triplets with 1-based rows
// dall+count3 and the reference's triplet order.  Host-side: static per geometry and weight.
void glue_smoothing_rows(int nx, int ny, int nz, int dall, float weight, std::vector<int> &rows1,
                         std::vector<int> &cols1, std::vector<float> &vals, int *count3_out) {
  const int nvx = nx - 2, nvz = ny - 2, nk = nz - 1;
  const int plane = nvz * nvx;
  rows1.clear();
  cols1.clear();
  vals.clear();
  int count3 = 0;
  const float wc = 6.0f * weight, wo = -1.0f * weight, we = 2.0f * weight;
  for (int k = 1; k <= nk; k++)
    for (int j = 1; j <= nvz; j++)
      for (int i = 1; i <= nvx; i++) {
        count3++;
        const int c0 = (k - 1) * plane + (j - 1) * nvx + i;
        const int r = dall + count3;
        const bool edge = i == 1 || i == nvx || j == 1 || j == nvz || k == 1 || k == nk;
        if (edge) {
          rows1.push_back(r);
          cols1.push_back(c0);
          vals.push_back(we);
        } else {
          const int off[7] = {0, -1, 1, -nvx, nvx, -plane, plane};
          for (int q = 0; q < 7; q++) {
            rows1.push_back(r);
            cols1.push_back(c0 + off[q]);
            vals.push_back(q == 0 ? wc : wo);
          }
        }
      }
  *count3_out = count3;
}

int glue_apply(cudaStream_t st, int dall, int maxvp, long long nar, const float *d_obst, const float *d_dsyn,
               float threshold0, const int *d_rows1, const int *d_cols1, float *d_rw, float *d_cbst, float *d_datw,
               double *d_norm, DevBuf<float> &sorted, DevBuf<char> &tmp, GlueStats *stats) {
  k_residual<<<gridn(dall), 256, 0, st>>>(d_obst, d_dsyn, d_cbst, dall);
  float q25 = 0, q75 = 0;
  DS_CHECK(glue_percentiles(st, d_cbst, dall, sorted, tmp, &q25, &q75));
  const float lo = q25 * threshold0, hi = q75 * threshold0;
  k_outliers<<<gridn(dall), 256, 0, st>>>(d_cbst, d_datw, dall, lo, hi);
  DS_CUDA(cudaMemsetAsync(d_norm, 0, ((size_t)maxvp + 2) * sizeof(double), st));
  if (nar > 0) k_scale_dws<<<gridn(nar), 256, 0, st>>>(d_rw, d_rows1, d_cols1, d_datw, nar, d_norm);
  k_dws_stats<<<1, 1024, 0, st>>>(d_norm, maxvp, d_norm + maxvp);
  double h[2] = {0, 0};
  DS_CUDA(cudaMemcpyAsync(h, d_norm + maxvp, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  DS_CUDA(cudaGetLastError());
  if (stats) {
    stats->q25 = q25;
    stats->q75 = q75;
    stats->maxnorm = (float)h[0];
    stats->averdws = (float)(h[1] / (double)maxvp);
  }
  return DSURF_OK;
}

int glue_model_update(cudaStream_t st, float *d_vels, float *d_dv, int nx, int ny, int nz, float minvel, float maxvel) {
  const int n = (nx - 2) * (ny - 2) * (nz - 1);
  k_model_update<<<gridn(n), 256, 0, st>>>(d_vels, d_dv, nx, ny, nz, minvel, maxvel);
  DS_CUDA(cudaGetLastError());
  return DSURF_OK;
}

}  // namespace dsurf
