// ORACLE (test infrastructure, see oracle.h) -- CPU restatement of subroutine CalSurfG
// (src/CalSurfG.f90:939-1459): dispersion maps + depth kernels per data type, then for every
// (period-type knumi, gather srcnum, ig): gridder + refined/coarse FMM + receiver times +
// rays + Frechet row assembly into COO triplets.
//
// Differences in *mechanism* only (results identical):
//  * the reference copies whole sen_* slabs into combined arrays per gather (:1146-1168) and
//    indexes them with knumi; here the per-type arrays are indexed with knumi - type offset.
//    (Requires wavetype/igrt of a gather to match the block knumi lies in, which the
//    reference's reader guarantees, main.f90:247-251.)
//  * the dense row(nparpi) scan (:1383,1425-1432) is replaced by an ordered walk over the
//    non-zero fdm entries that emits the same (nn ascending) triplets.
//  * gathers may be evaluated on several threads (mode 1) and concatenated in order.
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>
#include "fmm.h"
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
using oracle::Fmm;
using clk = std::chrono::steady_clock;
inline double secs(clk::time_point a, clk::time_point b) {
  return std::chrono::duration<double>(b - a).count();
}
inline float pow3(float x) { return x * (x * x); }
inline float pow4(float x) { float x2 = x * x; return x2 * x2; }

struct RowOut {
  std::vector<float> val;
  std::vector<int> col;
  std::vector<int> rowlen;  // per ray
  std::vector<float> tt;    // per ray (ig==1 pass)
};
}  // namespace

extern "C" int oracle_fmm_sweep(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                                const double *pv, float scx, float scz, float *veln, float *ttn,
                                float *ttnr, int *nstsr, float *rgeom) {
  Fmm f;
  f.setup(nx, ny, goxd, gozd, dvxd, dvzd);
  f.solve_source(pv, scx, scz);
  if (f.error) return f.error;
  for (int ix = 1; ix <= f.nnx; ix++)
    for (int iz = 1; iz <= f.nnz; iz++) {
      size_t o = (size_t)(ix - 1) * f.nnz + (iz - 1);
      if (veln) veln[o] = f.V(iz, ix);
      if (ttn) ttn[o] = f.T(iz, ix);
    }
  for (int ix = 1; ix <= f.nnxr; ix++)
    for (int iz = 1; iz <= f.nnzr; iz++) {
      size_t o = (size_t)(ix - 1) * f.nnzr + (iz - 1);
      if (ttnr) ttnr[o] = f.TR(iz, ix);
      if (nstsr) nstsr[o] = f.SR(iz, ix);
    }
  if (rgeom) {
    rgeom[0] = f.goxr;
    rgeom[1] = f.gozr;
    rgeom[2] = f.dnxr;
    rgeom[3] = f.dnzr;
    rgeom[4] = (float)f.nnxr;
    rgeom[5] = (float)f.nnzr;
  }
  return 0;
}

extern "C" int oracle_sweep_rays(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                                 const double *pv, float scx, float scz, int nrc, const float *rcx,
                                 const float *rcz, float *tt, float *fdm) {
  Fmm f;
  f.setup(nx, ny, goxd, gozd, dvxd, dvzd);
  f.solve_source(pv, scx, scz);
  if (f.error) return f.error;
  size_t fsz = (size_t)(f.nvz + 2) * (f.nvx + 2);
  for (int r = 0; r < nrc; r++) {
    if (tt) tt[r] = f.srtimes(scx, scz, rcx[r], rcz[r]);
    if (fdm) f.rpaths(scx, scz, fdm + fsz * r, rcx[r], rcz[r]);
    if (f.error) return f.error;
  }
  return 0;
}

// ray geometry of every receiver of one sweep: npts[nrc]; px/pz[nrc][cap] = rgx/rgz (radians)
extern "C" int oracle_sweep_paths(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                                  const double *pv, float scx, float scz, int nrc, const float *rcx,
                                  const float *rcz, int cap, int *npts, float *px, float *pz) {
  Fmm f;
  f.setup(nx, ny, goxd, gozd, dvxd, dvzd);
  f.solve_source(pv, scx, scz);
  if (f.error) return f.error;
  std::vector<float> fdm((size_t)(f.nvz + 2) * (f.nvx + 2)), x, z;
  f.path_x = &x;
  f.path_z = &z;
  for (int r = 0; r < nrc; r++) {
    f.rpaths(scx, scz, fdm.data(), rcx[r], rcz[r]);
    if (f.error) return f.error;
    npts[r] = (int)x.size();
    for (int j = 0; j < (int)x.size() && j < cap; j++) {
      px[(size_t)r * cap + j] = x[j];
      pz[(size_t)r * cap + j] = z[j];
    }
  }
  return 0;
}

namespace {
// gather loop of CalSurfG (:1135-1456) on given dispersion results.  pv[t] / sen[t][q] per type
// t = Rc,Rg,Lc,Lg (pv of Rc/Lc have kmax columns, see :1003-1004), q = vs,vp,rho.
// Only gathers [g_lo, g_hi) of the flattened (knumi, srcnum) nest are evaluated (g_hi < 0: all);
// rows are numbered globally as in a full run.
int calsurfg_core(int nx, int ny, int nz, const float *vels, int *iw, float *rw, int *col, float *dsurf,
                  float goxdf, float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                  int kmaxLg, const double *const pvp[4], const double *const senp[4][3], const int *wavetype,
                  const int *igrt, const int *periods, const float *depz, const float *scxf, const float *sczf,
                  const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1, int kmax,
                  int nsrcsurf, int nrcf, int *nar_out, int nthreads, int mode, int *rbint_out,
                  double *stage_seconds, int g_lo, int g_hi, double t_disp) {
  const float ftol = 1e-4f;  // CalSurfG.f90:1029
  const size_t ncol = (size_t)nx * ny;
  const int nvx = nx - 2, nvz = ny - 2;
  if (nthreads < 1) nthreads = 1;
  const double *pvRc = pvp[0], *pvRg = pvp[1], *pvLc = pvp[2], *pvLg = pvp[3];
  const double *const *sRc = senp[0], *const *sRg = senp[1], *const *sLc = senp[2], *const *sLg = senp[3];
  auto t1 = clk::now();
  const int kmax1 = kmaxRc, kmax2 = kmaxRc + kmaxRg, kmax3 = kmaxRc + kmaxRg + kmaxLc;

  // ---- flatten the (knumi, srcnum) loop nest of :1144-1145
  struct Gather { int knumi, srcnum; };
  std::vector<Gather> gathers;
  std::vector<int> grow;  // first global row of each gather
  {
    int row = 0;
    for (int knumi = 1; knumi <= kmax; knumi++)
      for (int srcnum = 1; srcnum <= nsrcsurf1[knumi - 1]; srcnum++) {
        gathers.push_back({knumi, srcnum});
        grow.push_back(row);
        row += nrc1[(size_t)(knumi - 1) * nsrcsurf + (srcnum - 1)];
      }
  }
  const long gl = g_lo < 0 ? 0 : g_lo, gh = (g_hi < 0 || g_hi > (long)gathers.size()) ? (long)gathers.size() : g_hi;
  std::vector<RowOut> outs(gathers.size());
  int err = 0, rbint = 0;
  double fmm_s = 0, ray_s = 0;
  long nsweeps = 0, nrays = 0;
  const int gthreads = (mode == 1) ? nthreads : 1;
  auto I2 = [&](int srcnum, int knumi) { return (size_t)(knumi - 1) * nsrcsurf + (srcnum - 1); };
  auto I3 = [&](int istep, int srcnum, int knumi) {
    return ((size_t)(knumi - 1) * nsrcsurf + (srcnum - 1)) * nrcf + (istep - 1);
  };
#ifdef _OPENMP
#pragma omp parallel num_threads(gthreads) reduction(+ : fmm_s, ray_s, nsweeps, nrays)
#endif
  {
    Fmm f;
    f.setup(nx, ny, goxdf, gozdf, dvxdf, dvzdf);
    std::vector<float> fdm((size_t)(nvz + 2) * (nvx + 2));
    std::vector<float> coe_a(nz), coe_rho(nz), vpft(nz);
    std::vector<int> nzj, nzk;
    std::vector<float> vals;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (long g = gl; g < gh; g++) {
      const int knumi = gathers[g].knumi, srcnum = gathers[g].srcnum;
      const int wt = wavetype[I2(srcnum, knumi)], gr = igrt[I2(srcnum, knumi)];
      const int per = periods[I2(srcnum, knumi)];
      const double *velf = nullptr;
      const double *const *sen = nullptr;  // [0]=vs [1]=vp [2]=rho of this type
      int koff = 0, ktype = 1;
      if (wt == 2 && gr == 0) { velf = &pvRc[(size_t)(per - 1) * ncol]; sen = sRc; koff = 0; ktype = kmaxRc; }
      if (wt == 2 && gr == 1) { velf = &pvRg[(size_t)(per - 1) * ncol]; sen = sRg; koff = kmax1; ktype = kmaxRg; }
      if (wt == 1 && gr == 0) { velf = &pvLc[(size_t)(per - 1) * ncol]; sen = sLc; koff = kmax2; ktype = kmaxLc; }
      if (wt == 1 && gr == 1) { velf = &pvLg[(size_t)(per - 1) * ncol]; sen = sLg; koff = kmax3; ktype = kmaxLg; }
      if (!velf) continue;
      const int ksen = knumi - koff;  // column of sen_*(:,knumi,:) inside this type's slab
      const int igroup = (gr == 1) ? 2 : 1;
      const float x = scxf[I2(srcnum, knumi)], z = sczf[I2(srcnum, knumi)];
      const int nrc = nrc1[I2(srcnum, knumi)];
      RowOut &o = outs[g];
      for (int ig = 1; ig <= igroup; ig++) {
        if (ig == 2 && wt == 2) velf = &pvRc[(size_t)(per - 1) * ncol];
        if (ig == 2 && wt == 1) velf = &pvLc[(size_t)(per - 1) * ncol];
        auto ta = clk::now();
        f.solve_source(velf, x, z);
        auto tb = clk::now();
        fmm_s += secs(ta, tb);
        nsweeps++;
        if (f.error) {
#ifdef _OPENMP
#pragma omp critical
#endif
          err = f.error;
          f.error = 0;
          break;
        }
        for (int istep = 1; istep <= nrc; istep++) {
          const float rx = rcxf[I3(istep, srcnum, knumi)], rz = rczf[I3(istep, srcnum, knumi)];
          if (ig == 1) o.tt.push_back(f.srtimes(x, z, rx, rz));
          if (gr == 0 || (ig == 2 && gr == 1)) {
            f.rpaths(x, z, fdm.data(), rx, rz);
            nrays++;
            // ---- row assembly, :1383-1432
            nzj.clear();
            nzk.clear();
            vals.clear();
            const bool brocher = depz[nz - 2] < 35.0f;
            for (int jj = 1; jj <= nvz; jj++)
              for (int kk = 1; kk <= nvx; kk++) {
                const float fd = fdm[(size_t)kk * (nvz + 2) + jj];
                if (std::fabs(fd) >= ftol) {
                  for (int k = 0; k < nz - 1; k++) {
                    const float v = vels[(size_t)k * nx * ny + (size_t)jj * nx + kk];  // vels(kk+1,jj+1,k+1)
                    if (brocher) {
                      coe_a[k] = (2.0947f - 0.8206f * 2 * v + 0.2683f * 3 * (v * v) - 0.0251f * 4 * pow3(v));
                      vpft[k] = 0.9409f + 2.0947f * v - 0.8206f * (v * v) + 0.2683f * pow3(v) - 0.0251f * pow4(v);
                    } else {
                      coe_a[k] = (2.2110f - 0.8984f * 2 * v + 0.2786f * 3 * (v * v) - 0.02412f * 4 * pow3(v));
                      vpft[k] = 0.9098f + 2.2110f * v - 0.8984f * (v * v) + 0.2786f * pow3(v) - 0.02412f * pow4(v);
                    }
                    const float p = vpft[k];
                    coe_rho[k] = coe_a[k] * (1.6612f - 0.4721f * 2 * p + 0.0671f * 3 * (p * p) -
                                             0.0043f * 4 * pow3(p) + 0.000106f * 5 * pow4(p));
                    const size_t cidx = (size_t)jj * (nvx + 2) + kk;  // jj*(nvx+2)+kk+1, 0-based
                    const size_t si = ((size_t)k * ktype + (ksen - 1)) * ncol + cidx;
                    const double r = (sen[1][si] * (double)coe_a[k] + sen[2][si] * (double)coe_rho[k] + sen[0][si]) *
                                     (double)fd;
                    vals.push_back((float)r);
                  }
                  nzj.push_back(jj);
                  nzk.push_back(kk);
                }
              }
            int cnt = 0;
            const size_t nnzv = nzj.size();
            for (int k = 0; k < nz - 1; k++)
              for (size_t q = 0; q < nnzv; q++) {
                const float r = vals[q * (nz - 1) + k];
                if (std::fabs(r) > ftol) {
                  o.val.push_back(r);
                  o.col.push_back(k * nvz * nvx + (nzj[q] - 1) * nvx + nzk[q]);
                  cnt++;
                }
              }
            o.rowlen.push_back(cnt);
          }
          if (f.error) break;
        }
        ray_s += secs(tb, clk::now());
        if (f.error) {
#ifdef _OPENMP
#pragma omp critical
#endif
          err = f.error;
          f.error = 0;
          break;
        }
      }
#ifdef _OPENMP
#pragma omp critical
#endif
      if (f.rbint) rbint = 1;
    }
  }
  auto t2 = clk::now();
  // ---- ordered concatenation: count1 / count11 / nar bookkeeping of :1135-1136,1369-1431
  int nar = 0;
  for (long g = gl; g < gh; g++) {
    RowOut &o = outs[g];
    int count1 = grow[g];
    int count11 = count1;
    for (float t : o.tt) dsurf[count1++] = t;
    size_t p = 0;
    for (int len : o.rowlen) {
      count11++;
      for (int q = 0; q < len; q++, p++) {
        rw[nar] = o.val[p];
        iw[nar + 1] = count11;  // iw(nar+1) with 1-based nar -> C index nar+1 after increment
        col[nar] = o.col[p];
        nar++;
      }
    }
  }
  *nar_out = nar;
  if (rbint_out) *rbint_out = rbint;
  if (stage_seconds) {
    stage_seconds[0] = t_disp;
    stage_seconds[1] = secs(t1, t2);
    stage_seconds[2] = fmm_s;
    stage_seconds[3] = ray_s;
    stage_seconds[4] = (double)nsweeps;
    stage_seconds[5] = (double)nrays;
  }
  return err;
}
}  // namespace

extern "C" int oracle_calsurfg(int nx, int ny, int nz, int nparpi, const float *vels, int *iw,
                               float *rw, int *col, float *dsurf, float goxdf, float gozdf,
                               float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                               int kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                               const double *tLg, const int *wavetype, const int *igrt,
                               const int *periods, const float *depz, float minthk,
                               const float *scxf, const float *sczf, const float *rcxf,
                               const float *rczf, const int *nrc1, const int *nsrcsurf1, int kmax,
                               int nsrcsurf, int nrcf, int *nar_out, int nthreads, int mode,
                               int *rbint_out, double *stage_seconds) {
  (void)nparpi;
  const size_t ncol = (size_t)nx * ny;
  if (nthreads < 1) nthreads = 1;
  auto t0 = clk::now();
  // ---- dispersion maps and depth kernels, :1098-1133.  pvRc/pvLc are dimensioned with kmax
  // columns because caldespersion overwrites their leading kmaxRg/kmaxLg columns (:1110,1128).
  std::vector<double> pvRc(ncol * std::max(kmax, 1), 0.0), pvRg(ncol * std::max(kmaxRg, 1), 0.0),
      pvLc(ncol * std::max(kmax, 1), 0.0), pvLg(ncol * std::max(kmaxLg, 1), 0.0);
  std::vector<double> sRc[3], sRg[3], sLc[3], sLg[3];
  for (int q = 0; q < 3; q++) {
    sRc[q].assign(ncol * std::max(kmaxRc, 1) * nz, 0.0);
    sRg[q].assign(ncol * std::max(kmaxRg, 1) * nz, 0.0);
    sLc[q].assign(ncol * std::max(kmaxLc, 1) * nz, 0.0);
    sLg[q].assign(ncol * std::max(kmaxLg, 1) * nz, 0.0);
  }
  if (kmaxRc > 0)
    oracle_depthkernel(nx, ny, nz, vels, pvRc.data(), sRc[0].data(), sRc[1].data(), sRc[2].data(), 2, 0,
                       kmaxRc, tRc, depz, minthk, nthreads);
  if (kmaxRg > 0) {
    oracle_caldespersion(nx, ny, nz, vels, pvRc.data(), 2, 0, kmaxRg, tRg, depz, minthk, nthreads);
    oracle_depthkernel(nx, ny, nz, vels, pvRg.data(), sRg[0].data(), sRg[1].data(), sRg[2].data(), 2, 1,
                       kmaxRg, tRg, depz, minthk, nthreads);
  }
  if (kmaxLc > 0)
    oracle_depthkernel(nx, ny, nz, vels, pvLc.data(), sLc[0].data(), sLc[1].data(), sLc[2].data(), 1, 0,
                       kmaxLc, tLc, depz, minthk, nthreads);
  if (kmaxLg > 0) {
    oracle_caldespersion(nx, ny, nz, vels, pvLc.data(), 1, 0, kmaxLg, tLg, depz, minthk, nthreads);
    oracle_depthkernel(nx, ny, nz, vels, pvLg.data(), sLg[0].data(), sLg[1].data(), sLg[2].data(), 1, 1,
                       kmaxLg, tLg, depz, minthk, nthreads);
  }
  const double *pvp[4] = {pvRc.data(), pvRg.data(), pvLc.data(), pvLg.data()};
  const double *senp[4][3];
  for (int q = 0; q < 3; q++) {
    senp[0][q] = sRc[q].data();
    senp[1][q] = sRg[q].data();
    senp[2][q] = sLc[q].data();
    senp[3][q] = sLg[q].data();
  }
  return calsurfg_core(nx, ny, nz, vels, iw, rw, col, dsurf, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, kmaxRg, kmaxLc,
                       kmaxLg, pvp, senp, wavetype, igrt, periods, depz, scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1,
                       kmax, nsrcsurf, nrcf, nar_out, nthreads, mode, rbint_out, stage_seconds, -1, -1,
                       secs(t0, clk::now()));
}

// Same gather loop on caller-provided dispersion results (bench.py's CPU arm: the sweep stage
// without re-running the dispersion stage).  pv/sen layouts as oracle_depthkernel.
extern "C" int oracle_calsurfg_pre(int nx, int ny, int nz, const float *vels, int *iw, float *rw, int *col,
                                   float *dsurf, float goxdf, float gozdf, float dvxdf, float dvzdf, int kmaxRc,
                                   int kmaxRg, int kmaxLc, int kmaxLg, const double *const *pv4,
                                   const double *const *sen12, const int *wavetype, const int *igrt,
                                   const int *periods, const float *depz, const float *scxf, const float *sczf,
                                   const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                                   int kmax, int nsrcsurf, int nrcf, int *nar, int nthreads, int mode,
                                   int *rbint_out, double *stage_seconds, int g_lo, int g_hi) {
  const double *pvp[4] = {pv4[0], pv4[1], pv4[2], pv4[3]};
  const double *senp[4][3];
  for (int t = 0; t < 4; t++)
    for (int q = 0; q < 3; q++) senp[t][q] = sen12[t * 3 + q];
  return calsurfg_core(nx, ny, nz, vels, iw, rw, col, dsurf, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, kmaxRg, kmaxLc,
                       kmaxLg, pvp, senp, wavetype, igrt, periods, depz, scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1,
                       kmax, nsrcsurf, nrcf, nar, nthreads, mode, rbint_out, stage_seconds, g_lo, g_hi, 0.0);
}

// subroutine synthetic (CalSurfG.f90:2412-2865), noise-free part: caldespersion per data type
// (group maps for the group types, :2552-2613), then for every gather one eikonal solve on the
// gdx = gdz = 5 propagation grid and srtimes for every receiver (:2629-2839).  obst(i) = t; the
// caller adds the reference's t*gaussian()*noiselevel term (random_number stream, not restated).
// pv_out (may be NULL): the four maps back to back, each [kmaxX][nx*ny].
extern "C" int oracle_synthetic(int nx, int ny, int nz, const float *vels, float *obst, float goxdf, float gozdf,
                                float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc, int kmaxLg,
                                const double *tRc, const double *tRg, const double *tLc, const double *tLg,
                                const int *wavetype, const int *igrt, const int *periods, const float *depz,
                                float minthk, const float *scxf, const float *sczf, const float *rcxf,
                                const float *rczf, const int *nrc1, const int *nsrcsurf1, int kmax, int nsrcsurf,
                                int nrcf, int nthreads, double *pv_out) {
  const size_t ncol = (size_t)nx * ny;
  const int kt[4] = {kmaxRc, kmaxRg, kmaxLc, kmaxLg};
  const double *tp[4] = {tRc, tRg, tLc, tLg};
  const int iwave[4] = {2, 2, 1, 1}, igr[4] = {0, 1, 0, 1};
  std::vector<double> pv[4];
  for (int t = 0; t < 4; t++) {
    if (kt[t] <= 0) continue;
    pv[t].assign(ncol * kt[t], 0.0);
    oracle_caldespersion(nx, ny, nz, vels, pv[t].data(), iwave[t], igr[t], kt[t], tp[t], depz, minthk, nthreads);
  }
  if (pv_out) {
    size_t o = 0;
    for (int t = 0; t < 4; t++) {
      for (size_t i = 0; i < pv[t].size(); i++) pv_out[o + i] = pv[t][i];
      o += pv[t].size();
    }
  }
  struct G { int knumi, srcnum, row; };
  std::vector<G> gs;
  int row = 0;
  for (int knumi = 1; knumi <= kmax; knumi++)
    for (int srcnum = 1; srcnum <= nsrcsurf1[knumi - 1]; srcnum++) {
      gs.push_back({knumi, srcnum, row});
      row += nrc1[(size_t)(knumi - 1) * nsrcsurf + (srcnum - 1)];
    }
  int err = 0;
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    Fmm f;
    f.setup(nx, ny, goxdf, gozdf, dvxdf, dvzdf, 5);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (long g = 0; g < (long)gs.size(); g++) {
      const size_t i2 = (size_t)(gs[g].knumi - 1) * nsrcsurf + (gs[g].srcnum - 1);
      const int wt = wavetype[i2], gr = igrt[i2], per = periods[i2];
      int t = -1;
      if (wt == 2 && gr == 0) t = 0;
      if (wt == 2 && gr == 1) t = 1;
      if (wt == 1 && gr == 0) t = 2;
      if (wt == 1 && gr == 1) t = 3;
      if (t < 0 || kt[t] <= 0) continue;
      const float x = scxf[i2], z = sczf[i2];
      f.solve_source(&pv[t][(size_t)(per - 1) * ncol], x, z);
      if (!f.error)
        for (int istep = 1; istep <= nrc1[i2]; istep++) {
          const size_t i3 = i2 * nrcf + (istep - 1);
          obst[gs[g].row + istep - 1] = f.srtimes(x, z, rcxf[i3], rczf[i3]);
          if (f.error) break;
        }
      if (f.error) {
#ifdef _OPENMP
#pragma omp critical
#endif
        err = f.error;
        f.error = 0;
      }
    }
  }
  return err;
}
