// ORACLE (test infrastructure, see oracle.h) -- CPU restatement of
// depthkernel (src/CalSurfG.f90:1-169), refineGrid2LayerMdl (:2352-2411) and
// caldespersion (:2866-2927).  The OpenMP loop over jj mirrors the reference's only
// parallel region (CalSurfG.f90:39-44); each call of surfdisp96 has serial semantics.
#include <cmath>
#include <vector>
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
constexpr int NL = 200;
constexpr int NP = 60;

inline float pow3(float x) { return x * (x * x); }
inline float pow4(float x) { float x2 = x * x; return x2 * x2; }
inline float pow5(float x) { float x2 = x * x; return x2 * (x * x2); }

// Brocher relations, CalSurfG.f90:49-53 (REAL*4)
inline void brocher(float vs, float &vp, float &rho) {
  vp = 0.9409f + 2.0947f * vs - 0.8206f * (vs * vs) + 0.2683f * pow3(vs) - 0.0251f * pow4(vs);
  rho = 1.6612f * vp - 0.4721f * (vp * vp) + 0.0671f * pow3(vp) - 0.0043f * pow4(vp) +
        0.000106f * pow5(vp);
}
}  // namespace

// CalSurfG.f90:2352-2411
extern "C" void oracle_refine_grid2layer(float minthk0, int mmax, const float *dep, const float *vp,
                                         const float *vs, const float *rho, int *rmax, float *rdep,
                                         float *rvp, float *rvs, float *rrho, float *rthk) {
  int k = 0;
  float initdep = 0.0f;
  for (int i = 1; i <= mmax - 1; i++) {
    float thk = dep[i] - dep[i - 1];
    float minthk = thk / minthk0;
    int nsublay = (int)((thk + 1.0e-4f) / minthk) + 1;
    float newthk = thk / (float)nsublay;
    for (int j = 1; j <= nsublay; j++) {
      k = k + 1;
      rthk[k - 1] = newthk;
      rdep[k - 1] = initdep + rthk[k - 1];
      initdep = rdep[k - 1];
      rvp[k - 1] = vp[i - 1] + (float)(2 * j - 1) * (vp[i] - vp[i - 1]) / (float)(2 * nsublay);
      rvs[k - 1] = vs[i - 1] + (float)(2 * j - 1) * (vs[i] - vs[i - 1]) / (float)(2 * nsublay);
      rrho[k - 1] = rho[i - 1] + (float)(2 * j - 1) * (rho[i] - rho[i - 1]) / (float)(2 * nsublay);
    }
  }
  k = k + 1;
  rthk[k - 1] = 0.0f;
  rvp[k - 1] = vp[mmax - 1];
  rvs[k - 1] = vs[mmax - 1];
  rrho[k - 1] = rho[mmax - 1];
  rdep[k - 1] = dep[mmax - 1];
  *rmax = k;
}

namespace {
void column_curve(float minthk, int mmax, const float *depz, const float *vp, const float *vs,
                  const float *rho, int iwave, int igr, int kmax, const double *t, double *cg) {
  float rdep[NL], rvp[NL], rvs[NL], rrho[NL], rthk[NL];
  int rmax;
  oracle_refine_grid2layer(minthk, mmax, depz, vp, vs, rho, &rmax, rdep, rvp, rvs, rrho, rthk);
  oracle_surfdisp96(rthk, rvp, rvs, rrho, rmax, /*iflsph*/ 1, iwave, /*mode*/ 1, igr, kmax, t, cg);
}
}  // namespace

// CalSurfG.f90:1-169
extern "C" void oracle_depthkernel(int nx, int ny, int nz, const float *vel, double *pv,
                                   double *sen_vs, double *sen_vp, double *sen_rho, int iwave,
                                   int igr, int kmax, const double *t, const float *depz,
                                   float minthk, int nthreads) {
  const int mmax = nz;
  const float dlnVs = 0.01f, dlnVp = 0.01f, dlnrho = 0.01f;
  const size_t ncol = (size_t)nx * ny;
#ifdef _OPENMP
  int nt = nthreads > 0 ? nthreads : 1;
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
  for (int jj = 1; jj <= ny; jj++) {
    std::vector<float> vsz(nz), vpz(nz), rhoz(nz), vsm(nz), vpm(nz), rhom(nz);
    double cg1[NP + 20], cg2[NP + 20], cgRc[NP + 20];
    for (int ii = 1; ii <= nx; ii++) {
      const size_t colid = (size_t)(jj - 1) * nx + (ii - 1);
      for (int k = 0; k < nz; k++) vsz[k] = vel[(size_t)k * nx * ny + (size_t)(jj - 1) * nx + (ii - 1)];
      for (int k = 0; k < nz; k++) brocher(vsz[k], vpz[k], rhoz[k]);
      column_curve(minthk, mmax, depz, vpz.data(), vsz.data(), rhoz.data(), iwave, igr, kmax, t, cgRc);
      for (int n = 0; n < kmax; n++) pv[(size_t)n * ncol + colid] = cgRc[n];
      for (int k = 0; k < nz; k++) {
        vsm[k] = vsz[k];
        vpm[k] = vpz[k];
        rhom[k] = rhoz[k];
      }
      for (int i = 0; i < mmax; i++) {
        // dc/dVs, CalSurfG.f90:77-93
        vsm[i] = vsz[i] - 0.5f * dlnVs * vsz[i];
        column_curve(minthk, mmax, depz, vpm.data(), vsm.data(), rhom.data(), iwave, igr, kmax, t, cg1);
        vsm[i] = vsz[i] + 0.5f * dlnVs * vsz[i];
        column_curve(minthk, mmax, depz, vpm.data(), vsm.data(), rhom.data(), iwave, igr, kmax, t, cg2);
        vsm[i] = vsz[i];
        for (int n = 0; n < kmax; n++)
          sen_vs[((size_t)i * kmax + n) * ncol + colid] = (cg2[n] - cg1[n]) / (double)(dlnVs * vsz[i]);
        // dc/dVp, CalSurfG.f90:107-123
        vpm[i] = vpz[i] - 0.5f * dlnVp * vpz[i];
        column_curve(minthk, mmax, depz, vpm.data(), vsm.data(), rhom.data(), iwave, igr, kmax, t, cg1);
        vpm[i] = vpz[i] + 0.5f * dlnVp * vpz[i];
        column_curve(minthk, mmax, depz, vpm.data(), vsm.data(), rhom.data(), iwave, igr, kmax, t, cg2);
        vpm[i] = vpz[i];
        for (int n = 0; n < kmax; n++)
          sen_vp[((size_t)i * kmax + n) * ncol + colid] = (cg2[n] - cg1[n]) / (double)(dlnVp * vpz[i]);
        // dc/drho, CalSurfG.f90:134-150
        rhom[i] = rhoz[i] - 0.5f * dlnrho * rhoz[i];
        column_curve(minthk, mmax, depz, vpm.data(), vsm.data(), rhom.data(), iwave, igr, kmax, t, cg1);
        rhom[i] = rhoz[i] + 0.5f * dlnrho * rhoz[i];
        column_curve(minthk, mmax, depz, vpm.data(), vsm.data(), rhom.data(), iwave, igr, kmax, t, cg2);
        rhom[i] = rhoz[i];
        for (int n = 0; n < kmax; n++)
          sen_rho[((size_t)i * kmax + n) * ncol + colid] = (cg2[n] - cg1[n]) / (double)(dlnrho * rhoz[i]);
      }
    }
  }
}

// CalSurfG.f90:2866-2927
extern "C" void oracle_caldespersion(int nx, int ny, int nz, const float *vel, double *pv, int iwave,
                                     int igr, int kmax, const double *t, const float *depz,
                                     float minthk, int nthreads) {
  const size_t ncol = (size_t)nx * ny;
#ifdef _OPENMP
  int nt = nthreads > 0 ? nthreads : 1;
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
  for (int jj = 1; jj <= ny; jj++) {
    std::vector<float> vsz(nz), vpz(nz), rhoz(nz);
    double cgRc[NP + 20];
    for (int ii = 1; ii <= nx; ii++) {
      const size_t colid = (size_t)(jj - 1) * nx + (ii - 1);
      for (int k = 0; k < nz; k++) vsz[k] = vel[(size_t)k * nx * ny + (size_t)(jj - 1) * nx + (ii - 1)];
      for (int k = 0; k < nz; k++) brocher(vsz[k], vpz[k], rhoz[k]);
      column_curve(minthk, nz, depz, vpz.data(), vsz.data(), rhoz.data(), iwave, igr, kmax, t, cgRc);
      for (int n = 0; n < kmax; n++) pv[(size_t)n * ncol + colid] = cgRc[n];
    }
  }
}
