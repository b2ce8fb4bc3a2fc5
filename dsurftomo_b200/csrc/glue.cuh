// Device-resident host glue (glue.cu): main.f90:361-466, :518-532, getpercentile.f90.
#pragma once
#include <vector>
#include "common.cuh"

namespace dsurf {
struct GlueStats {
  float q25, q75, maxnorm, averdws;
};
int glue_percentiles(cudaStream_t st, const float *d_x, int n, DevBuf<float> &sorted, DevBuf<char> &tmp, float *q25,
                     float *q75);
void glue_smoothing_rows(int nx, int ny, int nz, int dall, float weight, std::vector<int> &rows1,
                         std::vector<int> &cols1, std::vector<float> &vals, int *count3_out);
int glue_apply(cudaStream_t st, int dall, int maxvp, long long nar, const float *d_obst, const float *d_dsyn,
               float threshold0, const int *d_rows1, const int *d_cols1, float *d_rw, float *d_cbst, float *d_datw,
               double *d_norm, DevBuf<float> &sorted, DevBuf<char> &tmp, GlueStats *stats);
int glue_model_update(cudaStream_t st, float *d_vels, float *d_dv, int nx, int ny, int nz, float minvel, float maxvel);
}  // namespace dsurf

struct dsurf_lsmr_sys;
namespace dsurf {
int lsmr_sys_create_dev(dsurf_lsmr_sys **out, int m, int n, long long nnz, const int *d_rows1, const int *d_cols1,
                        const float *d_vals, const float *d_b);
float *lsmr_x_dev(dsurf_lsmr_sys *s);  // solution of the last solve in the reference's column order (device)
}  // namespace dsurf
