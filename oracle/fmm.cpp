// ORACLE (test infrastructure, see oracle.h) -- CPU restatement of the reference's 2-D
// fast-marching eikonal solver, B-spline gridding, receiver times and ray/Frechet tracer:
// src/CalSurfG.f90 modules globalp/traveltime (:181-923), gridder (:1460), bsplrefine (:1562),
// srtimes (:1636), rpaths (:1771), bilinear (:2328) and the per-source refinement
// orchestration of CalSurfG (:1186-1355).  All arithmetic is REAL*4, evaluated in the
// reference's operand order (build with -ffp-contract=off).
#include "fmm.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace oracle {

static inline float cube(float x) { return x * (x * x); }  // gfortran powi(3)

static inline void bspline4(float u, float o[5]) {  // CalSurfG.f90:1510-1513 (and :2180-2187)
  o[1] = cube(1.0f - u) / 6.0f;
  o[2] = (4.0f - 6.0f * (u * u) + 3.0f * cube(u)) / 6.0f;
  o[3] = (1.0f + 3.0f * u + 3.0f * (u * u) - 3.0f * cube(u)) / 6.0f;
  o[4] = cube(u) / 6.0f;
}

// CalSurfG.f90:1032-1094
void Fmm::setup(int nx, int ny, float goxdf, float gozdf, float dvxdf, float dvzdf, int gd) {
  gdx = gd; gdz = gd; asgr = 1; sgdl = 8; sgs = 8; earth = 6371.0f; fom = 1; snb = 0.5f;
  goxd = goxdf; gozd = gozdf; dvxd = dvxdf; dvzd = dvzdf;
  nvx = nx - 2;
  nvz = ny - 2;
  velv.assign((size_t)(nvz + 2) * (nvx + 2), 0.0f);
  dvx = dvxd * pi / 180.0f;
  dvz = dvzd * pi / 180.0f;
  gox = (90.0f - goxd) * pi / 180.0f;
  goz = gozd * pi / 180.0f;
  nnx = (nvx - 1) * gdx + 1;
  nnz = (nvz - 1) * gdz + 1;
  dnx = dvx / (float)gdx;
  dnz = dvz / (float)gdz;
  dnxd = dvxd / (float)gdx;
  dnzd = dvzd / (float)gdz;
  const int rmaxn = 2 * sgs * sgdl + 1;  // largest refined extent (:1233-1234)
  ld = std::max(nnz, rmaxn);
  ncols = std::max(nnx, rmaxn);
  size_t tot = (size_t)ld * ncols;
  veln.assign(tot, 0.0f);
  velnb.assign(tot, 0.0f);
  ttn.assign(tot, 0.0f);
  ttnr.assign(tot, 0.0f);
  nsts.assign(tot, -1);
  nstsr.assign(tot, -1);
  size_t maxbt = (size_t)std::lround((double)snb * ld * ncols) + 8;
  btg_px.assign(maxbt + 1, 0);
  btg_pz.assign(maxbt + 1, 0);
  rbint = 0;
  error = 0;
}

// CalSurfG.f90:1460-1553
void Fmm::gridder(const double *pv) {
  for (int i = 0; i <= nvz + 1; i++)
    for (int j = 0; j <= nvx + 1; j++) VV(i, j) = (float)pv[(size_t)i * (nvx + 2) + j];
  float ui[10][5], vi[10][5];
  for (int i = 1; i <= gdx + 1; i++) {
    float u = (float)gdx;
    u = (float)(i - 1) / u;
    bspline4(u, ui[i]);
  }
  for (int i = 1; i <= gdz + 1; i++) {
    float u = (float)gdz;
    u = (float)(i - 1) / u;
    bspline4(u, vi[i]);
  }
  for (int i = 1; i <= nvz - 1; i++) {
    int conz = gdz;
    if (i == nvz - 1) conz = gdz + 1;
    for (int j = 1; j <= nvx - 1; j++) {
      int conx = gdx;
      if (j == nvx - 1) conx = gdx + 1;
      for (int l = 1; l <= conz; l++) {
        int stz = gdz * (i - 1) + l;
        for (int m = 1; m <= conx; m++) {
          int stx = gdx * (j - 1) + m;
          float sumi = 0.0f;
          for (int i1 = 1; i1 <= 4; i1++) {
            float sumj = 0.0f;
            for (int j1 = 1; j1 <= 4; j1++) sumj = sumj + ui[m][j1] * VV(i - 2 + i1, j - 2 + j1);
            sumi = sumi + vi[l][i1] * sumj;
          }
          V(stz, stx) = sumi;
        }
      }
    }
  }
}

// CalSurfG.f90:1562-1628.  The reference scans every B-spline cell; cells that cannot
// intersect the refined box [vnt,vnb]x[vnl,vnr] are skipped here (identical stores).
void Fmm::bsplrefine() {
  const int nrxr = gdx * sgdl, nrzr = gdz * sgdl;
  static thread_local float ub[66][5];
  for (int j = 1; j <= nrxr + 1; j++) {
    float u = (float)nrxr;
    u = (float)(j - 1) / u;
    bspline4(u, ub[j]);  // ui(j,i,:) depends on j only; vi(j,i,:) on i only, same values
  }
  const int origx = (vnl - 1) * sgdl + 1;
  const int origz = (vnt - 1) * sgdl + 1;
  for (int i = 1; i <= nvz - 1; i++) {
    int conz = nrzr;
    if (i == nvz - 1) conz = nrzr + 1;
    if (gdz * (i - 1) + 1 > vnb || gdz * (i - 1) + (conz - 1) / sgdl + 1 < vnt) continue;
    for (int j = 1; j <= nvx - 1; j++) {
      int conx = nrxr;
      if (j == nvx - 1) conx = nrxr + 1;
      if (gdx * (j - 1) + 1 > vnr || gdx * (j - 1) + (conx - 1) / sgdl + 1 < vnl) continue;
      for (int k = 1; k <= conz; k++) {
        int st1 = gdz * (i - 1) + (k - 1) / sgdl + 1;
        if (st1 < vnt || st1 > vnb) continue;
        st1 = nrzr * (i - 1) + k;
        for (int l = 1; l <= conx; l++) {
          int st2 = gdx * (j - 1) + (l - 1) / sgdl + 1;
          if (st2 < vnl || st2 > vnr) continue;
          st2 = nrxr * (j - 1) + l;
          float sum[5];
          for (int i1 = 1; i1 <= 4; i1++) {
            sum[i1] = 0.0f;
            for (int j1 = 1; j1 <= 4; j1++) sum[i1] = sum[i1] + ub[l][j1] * VV(i - 2 + i1, j - 2 + j1);
            sum[i1] = ub[k][i1] * sum[i1];
          }
          int idm1 = st1 - origz + 1;
          int idm2 = st2 - origx + 1;
          if (idm1 < 1 || idm1 > nnz) continue;
          if (idm2 < 1 || idm2 > nnx) continue;
          V(idm1, idm2) = sum[1] + sum[2] + sum[3] + sum[4];
        }
      }
    }
  }
}

// CalSurfG.f90:2328-2349 (nv is 1-based [i][j])
float Fmm::bilinear(const float nv[3][3], float dsx, float dsz) {
  float biv = 0.0f;
  for (int i = 1; i <= 2; i++)
    for (int j = 1; j <= 2; j++) {
      float produ = (1.0f - std::fabs(((float)(i - 1) * dnx - dsx) / dnx)) *
                    (1.0f - std::fabs(((float)(j - 1) * dnz - dsz) / dnz));
      biv = biv + nv[i][j] * produ;
    }
  return biv;
}

// CalSurfG.f90:768-805
void Fmm::addtree(int iz, int ix) {
  ntr = ntr + 1;
  S(iz, ix) = ntr;
  btg_px[ntr] = ix;
  btg_pz[ntr] = iz;
  int tpc = ntr;
  int tpp = tpc / 2;
  while (tpp > 0) {
    if (T(iz, ix) < T(btg_pz[tpp], btg_px[tpp])) {
      S(iz, ix) = tpp;
      S(btg_pz[tpp], btg_px[tpp]) = tpc;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
      tpc = tpp;
      tpp = tpc / 2;
    } else {
      tpp = 0;
    }
  }
}

// CalSurfG.f90:816-885
void Fmm::downtree() {
  if (ntr == 1) {
    ntr = ntr - 1;
    return;
  }
  S(btg_pz[ntr], btg_px[ntr]) = 1;
  btg_px[1] = btg_px[ntr];
  btg_pz[1] = btg_pz[ntr];
  ntr = ntr - 1;
  int tpp = 1;
  int tpc = 2 * tpp;
  while (tpc < ntr) {
    float rd1 = T(btg_pz[tpc], btg_px[tpc]);
    float rd2 = T(btg_pz[tpc + 1], btg_px[tpc + 1]);
    if (rd1 > rd2) tpc = tpc + 1;
    rd1 = T(btg_pz[tpc], btg_px[tpc]);
    rd2 = T(btg_pz[tpp], btg_px[tpp]);
    if (rd1 < rd2) {
      S(btg_pz[tpp], btg_px[tpp]) = tpc;
      S(btg_pz[tpc], btg_px[tpc]) = tpp;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
      tpp = tpc;
      tpc = 2 * tpp;
    } else {
      tpc = ntr + 1;
    }
  }
  if (tpc == ntr) {
    float rd1 = T(btg_pz[tpc], btg_px[tpc]);
    float rd2 = T(btg_pz[tpp], btg_px[tpp]);
    if (rd1 < rd2) {
      S(btg_pz[tpp], btg_px[tpp]) = tpc;
      S(btg_pz[tpc], btg_px[tpc]) = tpp;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
    }
  }
}

// CalSurfG.f90:894-921 (sifts UP only, even if the value increased)
void Fmm::updtree(int iz, int ix) {
  int tpc = S(iz, ix);
  int tpp = tpc / 2;
  while (tpp > 0) {
    if (T(iz, ix) < T(btg_pz[tpp], btg_px[tpp])) {
      S(iz, ix) = tpp;
      S(btg_pz[tpp], btg_px[tpp]) = tpc;
      std::swap(btg_px[tpc], btg_px[tpp]);
      std::swap(btg_pz[tpc], btg_pz[tpp]);
      tpc = tpp;
      tpp = tpc / 2;
    } else {
      tpp = 0;
    }
  }
}

// CalSurfG.f90:587-759 -- mixed-order upwind update on the spherical-shell grid
void Fmm::fouds2(int iz, int ix) {
  int tsw1 = 0;
  float travm = 0.0f;
  const float slown = 1.0f / V(iz, ix);
  const float ri = earth;
  const float risti = ri * std::sin(gox + (float)(ix - 1) * dnx);
  for (int j = ix - 1; j <= ix + 1; j += 2) {
    if (j < 1 || j > nnx) continue;
    int swj = -1, j2;
    if (j == ix - 1) {
      j2 = j - 1;
      if (j2 >= 1 && S(iz, j2) == 0) swj = 0;
    } else {
      j2 = j + 1;
      if (j2 <= nnx && S(iz, j2) == 0) swj = 0;
    }
    if (S(iz, j) == 0 && swj == 0) {
      swj = -1;
      if (T(iz, j) > T(iz, j2)) swj = 0;
    } else {
      swj = -1;
    }
    for (int k = iz - 1; k <= iz + 1; k += 2) {
      if (k < 1 || k > nnz) continue;
      int swk = -1, k2;
      if (k == iz - 1) {
        k2 = k - 1;
        if (k2 >= 1 && S(k2, ix) == 0) swk = 0;
      } else {
        k2 = k + 1;
        if (k2 <= nnz && S(k2, ix) == 0) swk = 0;
      }
      if (S(k, ix) == 0 && swk == 0) {
        swk = -1;
        if (T(k, ix) > T(k2, ix)) swk = 0;
      } else {
        swk = -1;
      }
      int swsol = 0;
      float a = 0, b = 0, c = 0, u, v, em, tref = 0, tdiv = 1.0f;
      if (swj == 0) {
        swsol = 1;
        if (swk == 0) {
          u = 2.0f * ri * dnx;
          v = 2.0f * risti * dnz;
          em = 4.0f * T(iz, j) - T(iz, j2) - 4.0f * T(k, ix);
          em = em + T(k2, ix);
          a = v * v + u * u;
          b = 2.0f * em * (u * u);
          c = (u * u) * (em * em - (slown * slown) * (v * v));
          tref = 4.0f * T(iz, j) - T(iz, j2);
          tdiv = 3.0f;
        } else if (S(k, ix) == 0) {
          u = risti * dnz;
          v = 2.0f * ri * dnx;
          em = 3.0f * T(k, ix) - 4.0f * T(iz, j) + T(iz, j2);
          a = v * v + 9.0f * (u * u);
          b = 6.0f * em * (u * u);
          c = (u * u) * (em * em - (slown * slown) * (v * v));
          tref = T(k, ix);
          tdiv = 1.0f;
        } else {
          u = 2.0f * ri * dnx;
          a = 1.0f;
          b = 0.0f;
          c = -((u * u) * (slown * slown));
          tref = 4.0f * T(iz, j) - T(iz, j2);
          tdiv = 3.0f;
        }
      } else if (S(iz, j) == 0) {
        swsol = 1;
        if (swk == 0) {
          u = ri * dnx;
          v = 2.0f * risti * dnz;
          em = 3.0f * T(iz, j) - 4.0f * T(k, ix) + T(k2, ix);
          a = v * v + 9.0f * (u * u);
          b = 6.0f * em * (u * u);
          c = (u * u) * (em * em - (v * v) * (slown * slown));
          tref = T(iz, j);
          tdiv = 1.0f;
        } else if (S(k, ix) == 0) {
          u = ri * dnx;
          v = risti * dnz;
          em = T(k, ix) - T(iz, j);
          a = u * u + v * v;
          b = -(2.0f * (u * u) * em);
          c = (u * u) * (em * em - (v * v) * (slown * slown));
          tref = T(iz, j);
          tdiv = 1.0f;
        } else {
          a = 1.0f;
          b = 0.0f;
          c = -((slown * slown) * (ri * ri) * (dnx * dnx));
          tref = T(iz, j);
          tdiv = 1.0f;
        }
      } else {
        if (swk == 0) {
          swsol = 1;
          u = 2.0f * risti * dnz;
          a = 1.0f;
          b = 0.0f;
          c = -((u * u) * (slown * slown));
          tref = 4.0f * T(k, ix) - T(k2, ix);
          tdiv = 3.0f;
        } else if (S(k, ix) == 0) {
          swsol = 1;
          a = 1.0f;
          b = 0.0f;
          c = -((slown * slown) * (risti * risti) * (dnz * dnz));
          tref = T(k, ix);
          tdiv = 1.0f;
        }
      }
      if (swsol == 1) {
        float rd1 = b * b - 4.0f * a * c;
        if (rd1 < 0.0f) rd1 = 0.0f;
        float tdsh = (-b + std::sqrt(rd1)) / (2.0f * a);
        float trav = (tref + tdsh) / tdiv;
        if (tsw1 == 1) {
          travm = std::min(trav, travm);
        } else {
          travm = trav;
          tsw1 = 1;
        }
      }
    }
  }
  T(iz, ix) = travm;
}

// CalSurfG.f90:288-487
void Fmm::travel(float scx, float scz, int urg) {
  int isx = (int)((scx - gox) / dnx) + 1;
  int isz = (int)((scz - goz) / dnz) + 1;
  int sw = 0;
  if (isx < 1 || isx > nnx) sw = 1;
  if (isz < 1 || isz > nnz) sw = 1;
  if (sw == 1) {
    error = 1;  // reference: "Source lies outside bounds of model" + STOP
    return;
  }
  if (isx == nnx) isx = isx - 1;
  if (isz == nnz) isz = isz - 1;
  if (urg != 2) std::fill(nsts.begin(), nsts.end(), -1);
  ntr = 0;
  if (urg == 2) {
    for (int i = 1; i <= nnx; i++)
      for (int j = 1; j <= nnz; j++)
        if (S(j, i) > 0) addtree(j, i);
  } else {
    float vss[3][3];
    for (int i = 1; i <= 2; i++)
      for (int j = 1; j <= 2; j++) vss[i][j] = V(isz - 1 + j, isx - 1 + i);
    float dsx = (scx - gox) - (float)(isx - 1) * dnx;
    float dsz = (scz - goz) - (float)(isz - 1) * dnz;
    float vsrc = bilinear(vss, dsx, dsz);
    for (int i = 1; i <= 2; i++)
      for (int j = 1; j <= 2; j++) {
        float ex = dsx - (float)(i - 1) * dnx;
        float ez = dsz - (float)(j - 1) * dnz;
        float ds = std::sqrt(ex * ex + ez * ez);
        T(isz - 1 + j, isx - 1 + i) = 2.0f * ds / (vss[i][j] + vsrc);
        addtree(isz - 1 + j, isx - 1 + i);
      }
  }
  while (ntr > 0) {
    int ix, iz;
    if (urg == 1) {  // refined-grid exit test, :392-412
      ix = btg_px[1];
      iz = btg_pz[1];
      int swrg = 0;
      if (ix == 1 && vnl != 1) swrg = 1;
      if (ix == nnx && vnr != nnx) swrg = 1;
      if (iz == 1 && vnt != 1) swrg = 1;
      if (iz == nnz && vnb != nnz) swrg = 1;
      if (swrg == 1) {
        S(iz, ix) = 0;
        break;
      }
    }
    ix = btg_px[1];
    iz = btg_pz[1];
    S(iz, ix) = 0;
    downtree();
    for (int i = ix - 1; i <= ix + 1; i += 2) {
      if (i >= 1 && i <= nnx) {
        if (S(iz, i) == -1) {
          fouds2(iz, i);
          addtree(iz, i);
        } else if (S(iz, i) > 0) {
          fouds2(iz, i);
          updtree(iz, i);
        }
      }
    }
    for (int i = iz - 1; i <= iz + 1; i += 2) {
      if (i >= 1 && i <= nnz) {
        if (S(i, ix) == -1) {
          fouds2(i, ix);
          addtree(i, ix);
        } else if (S(i, ix) > 0) {
          fouds2(i, ix);
          updtree(i, ix);
        }
      }
    }
  }
}

// CalSurfG.f90:1186-1355 (asgr == 1 branch; the reference hard-codes asgr=1 at :1034)
void Fmm::solve_source(const double *pv, float x, float z) {
  gridder(pv);
  for (int j = 1; j <= nnx; j++)
    for (int k = 1; k <= nnz; k++) VB(k, j) = V(k, j);
  const int nnxb = nnx, nnzb = nnz;
  const float dnxb = dnx, dnzb = dnz, goxb = gox, gozb = goz;
  int isx = (int)((x - gox) / dnx) + 1;
  int isz = (int)((z - goz) / dnz) + 1;
  int sw = 0;
  if (isx < 1 || isx > nnx) sw = 1;
  if (isz < 1 || isz > nnz) sw = 1;
  if (sw == 1) {
    error = 1;
    return;
  }
  if (isx == nnx) isx = isx - 1;
  if (isz == nnz) isz = isz - 1;
  vnl = isx - sgs;
  if (vnl < 1) vnl = 1;
  vnr = isx + sgs;
  if (vnr > nnx) vnr = nnx;
  vnt = isz - sgs;
  if (vnt < 1) vnt = 1;
  vnb = isz + sgs;
  if (vnb > nnz) vnb = nnz;
  nrnx = (vnr - vnl) * sgdl + 1;
  nrnz = (vnb - vnt) * sgdl + 1;
  drnx = dvx / (float)(gdx * sgdl);
  drnz = dvz / (float)(gdz * sgdl);
  gorx = gox + dnx * (float)(vnl - 1);
  gorz = goz + dnz * (float)(vnt - 1);
  nnx = nrnx;
  nnz = nrnz;
  dnx = drnx;
  dnz = drnz;
  gox = gorx;
  goz = gorz;
  bsplrefine();
  travel(x, z, 1);
  if (error) return;
  ttnr = ttn;  // whole-array assignments, :1287-1288
  nstsr = nsts;
  const int ogx = vnl, ogz = vnt, grdfx = sgdl, grdfz = sgdl;
  std::fill(nsts.begin(), nsts.end(), -1);
  for (int k = 1; k <= nnz; k += grdfz) {
    int idm1 = ogz + (k - 1) / grdfz;
    for (int l = 1; l <= nnx; l += grdfx) {
      int idm2 = ogx + (l - 1) / grdfx;
      S(idm1, idm2) = SR(k, l);
      if (S(idm1, idm2) >= 0) T(idm1, idm2) = TR(k, l);
    }
  }
  nnxr = nnx;
  nnzr = nnz;
  goxr = gox;
  gozr = goz;
  dnxr = dnx;
  dnzr = dnz;
  nnx = nnxb;
  nnz = nnzb;
  dnx = dnxb;
  dnz = dnzb;
  gox = goxb;
  goz = gozb;
  for (int j = 1; j <= nnx; j++)
    for (int k = 1; k <= nnz; k++) V(k, j) = VB(k, j);
  for (int k = 1; k <= nnx; k++)  // narrow-band completion, :1332-1349
    for (int l = 1; l <= nnz; l++)
      if (S(l, k) == 0) {
        if (l - 1 >= 1 && S(l - 1, k) == -1) S(l, k) = 1;
        if (l + 1 <= nnz && S(l + 1, k) == -1) S(l, k) = 1;
        if (k - 1 >= 1 && S(l, k - 1) == -1) S(l, k) = 1;
        if (k + 1 <= nnx && S(l, k + 1) == -1) S(l, k) = 1;
      }
  travel(x, z, 2);
}

// CalSurfG.f90:1636-1759
float Fmm::srtimes(float scx, float scz, float rcx1, float rcz1) {
  int irx = (int)((rcx1 - gox) / dnx) + 1;
  int irz = (int)((rcz1 - goz) / dnz) + 1;
  int sw = 0;
  if (irx < 1 || irx > nnx) sw = 1;
  if (irz < 1 || irz > nnz) sw = 1;
  if (sw == 1) {
    error = 2;  // "Receiver lies outside model" + STOP
    return 0.0f;
  }
  if (irx == nnx) irx = irx - 1;
  if (irz == nnz) irz = irz - 1;
  int isx = (int)((scx - gox) / dnx) + 1;
  int isz = (int)((scz - goz) / dnz) + 1;
  float dpl = dnx * earth;
  float rd1 = dnz * earth * std::sin(gox);
  if (rd1 < dpl) dpl = rd1;
  rd1 = dnz * earth * std::sin(gox + (float)(nnx - 1) * dnx);
  if (rd1 < dpl) dpl = rd1;
  float e1 = (scx - rcx1) * earth;
  float sred = e1 * e1;
  float e2 = (scz - rcz1) * earth * std::sin(rcx1);
  sred = sred + e2 * e2;
  sred = std::sqrt(sred);
  if (sred < dpl) sw = 1;
  if (isx == irx && isz == irz) sw = 1;
  float trr;
  if (sw == 1) {
    float vss[3][3];
    for (int k = 1; k <= 2; k++)
      for (int l = 1; l <= 2; l++) vss[k][l] = V(isz - 1 + l, isx - 1 + k);
    float drx = (scx - gox) - (float)(isx - 1) * dnx;
    float drz = (scz - goz) - (float)(isz - 1) * dnz;
    float vels = bilinear(vss, drx, drz);
    for (int k = 1; k <= 2; k++)
      for (int l = 1; l <= 2; l++) vss[k][l] = V(irz - 1 + l, irx - 1 + k);
    drx = (rcx1 - gox) - (float)(irx - 1) * dnx;
    drz = (rcz1 - goz) - (float)(irz - 1) * dnz;
    float velr = bilinear(vss, drx, drz);
    trr = 2.0f * sred / (vels + velr);
  } else {
    float drx = (rcx1 - gox) - (float)(irx - 1) * dnx;
    float drz = (rcz1 - goz) - (float)(irz - 1) * dnz;
    trr = 0.0f;
    for (int k = 1; k <= 2; k++)
      for (int l = 1; l <= 2; l++) {
        float produ = (1.0f - std::fabs(((float)(l - 1) * dnz - drz) / dnz)) *
                      (1.0f - std::fabs(((float)(k - 1) * dnx - drx) / dnx));
        trr = trr + T(irz - 1 + l, irx - 1 + k) * produ;
      }
  }
  return trr;
}

// CalSurfG.f90:1771-2318 -- fdm is fdm(0:nvz+1,0:nvx+1), column-major
void Fmm::rpaths(float scx, float scz, float *fdm, float surfrcx, float surfrcz) {
  const int fld = nvz + 2;
  auto F = [&](int i, int j) -> float & { return fdm[(size_t)j * fld + i]; };
  const long maxrp = (long)nnx * nnz;
  int isx, isz;
  if (asgr == 1) {
    isx = (int)((scx - goxr) / dnxr) + 1;
    isz = (int)((scz - gozr) / dnzr) + 1;
  } else {
    isx = (int)((scx - gox) / dnx) + 1;
    isz = (int)((scz - goz) / dnz) + 1;
  }
  float dpl = dnx * earth;
  float rd1 = dnz * earth * std::sin(gox);
  if (rd1 < dpl) dpl = rd1;
  rd1 = dnz * earth * std::sin(gox + (float)(nnx - 1) * dnx);
  if (rd1 < dpl) dpl = rd1;
  dpl = 0.5f * dpl;
  std::fill(fdm, fdm + (size_t)(nvz + 2) * (nvx + 2), 0.0f);
  int ipx = (int)((surfrcx - gox) / dnx) + 1;
  int ipz = (int)((surfrcz - goz) / dnz) + 1;
  int sw = 0;
  if (ipx < 1 || ipx >= nnx) sw = 1;
  if (ipz < 1 || ipz >= nnz) sw = 1;
  if (sw == 1) {
    error = 2;  // "rpath Receiver lies outside model" + STOP
    return;
  }
  float rgx_j = surfrcx, rgz_j = surfrcz;  // rgx(j), rgz(j)
  auto rec = [&](float x, float z) {
    if (path_x) path_x->push_back(x), path_z->push_back(z);
  };
  if (path_x) path_x->clear(), path_z->clear();
  rec(rgx_j, rgz_j);  // rgx(1) = receiver, :1910-1911
  float e1 = (scx - rgx_j) * earth;
  float sred = e1 * e1;
  float e2 = (scz - rgz_j) * earth * std::sin(rgx_j);
  sred = sred + e2 * e2;
  sred = std::sqrt(sred);
  if (sred < 2.0f * dpl) sw = 1;
  int ipxr = 0, ipzr = 0, igref;
  if (asgr == 1) {
    ipxr = (int)((surfrcx - goxr) / dnxr) + 1;
    ipzr = (int)((surfrcz - gozr) / dnzr) + 1;
    igref = 1;
    if (ipxr < 1 || ipxr >= nnxr) igref = 0;
    if (ipzr < 1 || ipzr >= nnzr) igref = 0;
    if (igref == 1) {
      if (SR(ipzr, ipxr) != 0 || SR(ipzr + 1, ipxr) != 0) igref = 0;
      if (SR(ipzr, ipxr + 1) != 0 || SR(ipzr + 1, ipxr + 1) != 0) igref = 0;
    }
  } else {
    igref = 0;
  }
  if (sw == 0) {
    if (asgr == 1) {
      if (igref == 1 && ipxr == isx && ipzr == isz) sw = 1;
    } else {
      if (ipx == isx && ipz == isz) sw = 1;
    }
  }
  if (sw == 1) rec(scx, scz);  // rgx(2) = source, nrp = 2, :1919-1921 / :1949-1961
  for (long j = 1; j <= maxrp; j++) {
    if (sw == 1) break;
    float dtx, dtz;
    const float sinx = std::sin(rgx_j);
    if (igref == 1) {
      dtx = TR(ipzr, ipxr + 1) - TR(ipzr, ipxr);
      dtx = dtx + TR(ipzr + 1, ipxr + 1) - TR(ipzr + 1, ipxr);
      dtx = dtx / (2.0f * earth * dnxr);
      dtz = TR(ipzr + 1, ipxr) - TR(ipzr, ipxr);
      dtz = dtz + TR(ipzr + 1, ipxr + 1) - TR(ipzr, ipxr + 1);
      dtz = dtz / (2.0f * earth * sinx * dnzr);
    } else {
      dtx = T(ipz, ipx + 1) - T(ipz, ipx);
      dtx = dtx + T(ipz + 1, ipx + 1) - T(ipz + 1, ipx);
      dtx = dtx / (2.0f * earth * dnx);
      dtz = T(ipz + 1, ipx) - T(ipz, ipx);
      dtz = dtz + T(ipz + 1, ipx + 1) - T(ipz, ipx + 1);
      dtz = dtz / (2.0f * earth * sinx * dnz);
    }
    rd1 = std::sqrt(dtx * dtx + dtz * dtz);
    float rgx_n = rgx_j - dpl * dtx / (earth * rd1);          // rgx(j+1)
    float rgz_n = rgz_j - dpl * dtz / (earth * sinx * rd1);   // rgz(j+1)
    const int ipxo = ipx, ipzo = ipz;
    if (asgr == 1) {
      ipxr = (int)((rgx_n - goxr) / dnxr) + 1;
      ipzr = (int)((rgz_n - gozr) / dnzr) + 1;
      igref = 1;
      if (ipxr < 1 || ipxr >= nnxr) igref = 0;
      if (ipzr < 1 || ipzr >= nnzr) igref = 0;
      if (igref == 1) {
        if (SR(ipzr, ipxr) != 0 || SR(ipzr + 1, ipxr) != 0) igref = 0;
        if (SR(ipzr, ipxr + 1) != 0 || SR(ipzr + 1, ipxr + 1) != 0) igref = 0;
      }
      ipx = (int)((rgx_n - gox) / dnx) + 1;
      ipz = (int)((rgz_n - goz) / dnz) + 1;
    } else {
      ipx = (int)((rgx_n - gox) / dnx) + 1;
      ipz = (int)((rgz_n - goz) / dnz) + 1;
      igref = 0;
    }
    e1 = (scx - rgx_n) * earth;
    sred = e1 * e1;
    e2 = (scz - rgz_n) * earth * std::sin(rgx_n);
    sred = sred + e2 * e2;
    sred = std::sqrt(sred);
    sw = 0;
    if (sred < 2.0f * dpl) sw = 1;
    if (sw == 0) {
      if (asgr == 1) {
        if (igref == 1 && ipxr == isx && ipzr == isz) sw = 1;
      } else {
        if (ipx == isx && ipz == isz) sw = 1;
      }
    }
    if (ipx < 1) {
      rgx_n = gox;
      ipx = 1;
      rbint = 1;
    }
    if (ipx >= nnx) {
      rgx_n = gox + (float)(nnx - 1) * dnx;
      ipx = nnx - 1;
      rbint = 1;
    }
    if (ipz < 1) {
      rgz_n = goz;
      ipz = 1;
      rbint = 1;
    }
    if (ipz >= nnz) {
      rgz_n = goz + (float)(nnz - 1) * dnz;
      ipz = nnz - 1;
      rbint = 1;
    }
    rec(rgx_n, rgz_n);           // rgx(j+1) as left by the boundary clamp
    if (sw == 1) rec(scx, scz);  // rgx(j+2) = source, nrp = j+2, :2042-2046 / :2057-2071
    // ---- Frechet derivatives, :2110-2265
    const int ivx = (ipx - 1) / gdx + 1;
    const int ivz = (ipz - 1) / gdz + 1;
    const int ivxo = (ipxo - 1) / gdx + 1;
    const int ivzo = (ipzo - 1) / gdz + 1;
    int nhp = 0;
    float vrat[5];
    int chp[5];
    if (ivx != ivxo) {
      nhp = nhp + 1;
      float xi;
      if (ivx > ivxo)
        xi = gox + (float)(ivx - 1) * dvx;
      else
        xi = gox + (float)ivx * dvx;
      vrat[nhp] = (xi - rgx_j) / (rgx_n - rgx_j);
      chp[nhp] = 1;
    }
    if (ivz != ivzo) {
      nhp = nhp + 1;
      float zi;
      if (ivz > ivzo)
        zi = goz + (float)(ivz - 1) * dvz;
      else
        zi = goz + (float)ivz * dvz;
      rd1 = (zi - rgz_j) / (rgz_n - rgz_j);
      if (nhp == 1) {
        vrat[nhp] = rd1;
        chp[nhp] = 2;
      } else {
        if (rd1 >= vrat[nhp - 1]) {
          vrat[nhp] = rd1;
          chp[nhp] = 2;
        } else {
          vrat[nhp] = vrat[nhp - 1];
          chp[nhp] = chp[nhp - 1];
          vrat[nhp - 1] = rd1;
          chp[nhp - 1] = 2;
        }
      }
    }
    nhp = nhp + 1;
    vrat[nhp] = 1.0f;
    chp[nhp] = 0;
    float drx = (rgx_j - gox) - (float)(ipxo - 1) * dnx;
    float drz = (rgz_j - goz) - (float)(ipzo - 1) * dnz;
    float vel = 0.0f;
    for (int l = 1; l <= 2; l++)
      for (int m = 1; m <= 2; m++) {
        float produ = (1.0f - std::fabs(((float)(m - 1) * dnz - drz) / dnz));
        produ = produ * (1.0f - std::fabs(((float)(l - 1) * dnx - drx) / dnx));
        if (ipzo - 1 + m <= nnz && ipxo - 1 + l <= nnx) vel = vel + V(ipzo - 1 + m, ipxo - 1 + l) * produ;
      }
    drx = (rgx_j - gox) - (float)(ivxo - 1) * dvx;
    drz = (rgz_j - goz) - (float)(ivzo - 1) * dvz;
    float v = drx / dvx;
    float w = drz / dvz;
    float vi[5], wi[5], vio[5], wio[5];
    bspline4(v, vi);
    bspline4(w, wi);
    int ivxt = ivxo, ivzt = ivzo;
    for (int k = 1; k <= nhp; k++) {
      const float velo = vel;
      for (int q = 1; q <= 4; q++) {
        vio[q] = vi[q];
        wio[q] = wi[q];
      }
      if (k > 1) {
        if (chp[k - 1] == 1)
          ivxt = ivx;
        else if (chp[k - 1] == 2)
          ivzt = ivz;
      }
      const float rigz = rgz_j + vrat[k] * (rgz_n - rgz_j);
      const float rigx = rgx_j + vrat[k] * (rgx_n - rgx_j);
      const int ipxt = (int)((rigx - gox) / dnx) + 1;
      const int ipzt = (int)((rigz - goz) / dnz) + 1;
      drx = (rigx - gox) - (float)(ipxt - 1) * dnx;
      drz = (rigz - goz) - (float)(ipzt - 1) * dnz;
      vel = 0.0f;
      for (int m = 1; m <= 2; m++)
        for (int n = 1; n <= 2; n++) {
          float produ = (1.0f - std::fabs(((float)(n - 1) * dnz - drz) / dnz));
          produ = produ * (1.0f - std::fabs(((float)(m - 1) * dnx - drx) / dnx));
          if (ipzt - 1 + n <= nnz && ipxt - 1 + m <= nnx && ipzt - 1 + n >= 1 && ipxt - 1 + m >= 1)
            vel = vel + V(ipzt - 1 + n, ipxt - 1 + m) * produ;
        }
      drx = (rigx - gox) - (float)(ivxt - 1) * dvx;
      drz = (rigz - goz) - (float)(ivzt - 1) * dvz;
      v = drx / dvx;
      w = drz / dvz;
      bspline4(v, vi);
      bspline4(w, wi);
      float dinc;
      if (k == 1)
        dinc = vrat[k] * dpl;
      else
        dinc = (vrat[k] - vrat[k - 1]) * dpl;
      for (int l = 1; l <= 4; l++)
        for (int m = 1; m <= 4; m++) {
          float r1 = vi[m] * wi[l] / (vel * vel);
          float r2 = vio[m] * wio[l] / (velo * velo);
          r1 = -(r1 + r2) * dinc / 2.0f;
          r2 = F(ivzt - 2 + l, ivxt - 2 + m);
          F(ivzt - 2 + l, ivxt - 2 + m) = r1 + r2;
        }
    }
    rgx_j = rgx_n;
    rgz_j = rgz_n;
  }
}

}  // namespace oracle
