// K1 -- 1-D Rayleigh/Love dispersion + finite-difference depth kernels on sm_100a, replacing
// depthkernel (src/CalSurfG.f90:1-169), caldespersion (:2866-2927), refineGrid2LayerMdl
// (:2352-2411) and surfdisp96 with its helpers (src/surfdisp96.f:52-1062).
//
// Mapping: one THREAD per layered model.  For the depth kernels a model is one of the
// V = 1 + 6*nz variants of a column (base, and +-0.5% in Vs, Vp, rho at each depth node,
// CalSurfG.f90:76-160); the variants of one column sit in adjacent lanes, so a warp walks
// nearly identical root-search paths.  Periods are sequential inside a thread because each
// period's bracket search starts from the previous root (surfdisp96.f:262-265).
// The per-thread layer stack (a = Vp, b = Vs, rho after earth flattening) lives in shared
// memory as [array][layer][thread] (bank-conflict free); everything that depends only on the
// layer geometry -- refined thicknesses, the flattening factors, btp**(-2.275), btp**(-5) -- is
// identical for all columns and is tabulated once on the host with the C library the
// reference itself would call (log / powf), which removes those transcendentals from the
// device and from the parity budget.
//
// Arithmetic: fp32 model, fp64 period-equation search, fp32-rounded result, exactly as the
// reference; compiled with --fmad=false.
// Bound: FP64 ALU + DP transcendentals (sin/cos/exp/sqrt); HBM traffic is negligible
// (4*nz bytes in, 8*kmax*(1+3nz) bytes out per column).
#include <cmath>
#include <vector>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "plan.cuh"

namespace dsurf {

// ------------------------------------------------------------------------------------------
// host-side layer geometry tables (refineGrid2LayerMdl thickness part + sphere(0,0) geometry)
// ------------------------------------------------------------------------------------------
void make_layer_tables_from_thk(const float *thk, int nlayer, int iflsph, LayerTables &T) {
  T.rmax = nlayer;
  T.dflat.assign(nlayer, 0.f);
  T.tmp.assign(nlayer, 1.0);
  T.facR.assign(nlayer, 1.f);
  T.facL.assign(nlayer, 1.f);
  for (int i = 0; i < nlayer; i++) T.dflat[i] = thk[i];
  if (iflsph == 1) {  // surfdisp96.f:510-533 (iflag = 0), d(mmax) = 1.0 on entry
    std::vector<float> d(thk, thk + nlayer);
    const double ar = 6370.0;
    double dr = 0.0, r0 = ar;
    d[nlayer - 1] = 1.0f;
    for (int i = 0; i < nlayer; i++) {
      dr = dr + (double)d[i];
      const double r1 = ar - dr;
      const double z0 = ar * std::log(ar / r0);
      const double z1 = ar * std::log(ar / r1);
      T.dflat[i] = (float)(z1 - z0);
      const double tmp = (ar + ar) / (r0 + r1);
      T.tmp[i] = tmp;
      const float btp = (float)tmp;
      const float x2 = btp * btp;
      const float x5 = x2 * (btp * x2);           // gfortran powi(5)
      T.facL[i] = 1.0f / x5;                      // btp**(-5), surfdisp96.f:539
      T.facR[i] = std::pow(btp, -2.275f);         // btp**(-2.275), surfdisp96.f:541
      r0 = r1;
    }
  }
  T.dflat[nlayer - 1] = 0.0f;  // d(mmax) = 0 on exit of sphere (and thkm(mmax) = 0 for flat models)
}

void make_layer_tables(const float *depz, int nz, float minthk0, LayerTables &T) {
  // CalSurfG.f90:2369-2387: thickness / interpolation weights of the refined stack
  std::vector<float> rthk;
  T.node.clear();
  T.wnum.clear();
  T.wden.clear();
  for (int i = 1; i <= nz - 1; i++) {
    const float thk = depz[i] - depz[i - 1];
    const float minthk = thk / minthk0;
    const int nsub = (int)((thk + 1.0e-4f) / minthk) + 1;
    const float newthk = thk / (float)nsub;
    for (int j = 1; j <= nsub; j++) {
      rthk.push_back(newthk);
      T.node.push_back(i - 1);
      T.wnum.push_back((float)(2 * j - 1));
      T.wden.push_back((float)(2 * nsub));
    }
  }
  rthk.push_back(0.0f);  // half space
  T.node.push_back(nz - 1);
  T.wnum.push_back(0.0f);
  T.wden.push_back(1.0f);
  make_layer_tables_from_thk(rthk.data(), (int)rthk.size(), /*iflsph*/ 1, T);
}

// ------------------------------------------------------------------------------------------
// device: period equations
// ------------------------------------------------------------------------------------------
struct Stack {            // view of one thread's layer stack in shared memory
  float *a, *b, *rho;        // element m (1-based layer) at [(m-1)*stride]
  const float *d;            // shared flattened thicknesses, [m-1]
  int stride, mmax, llw;
  __device__ __forceinline__ float af(int m) const { return a[(m - 1) * stride]; }
  __device__ __forceinline__ float bf(int m) const { return b[(m - 1) * stride]; }
  __device__ __forceinline__ double A(int m) const { return (double)a[(m - 1) * stride]; }
  __device__ __forceinline__ double B(int m) const { return (double)b[(m - 1) * stride]; }
  __device__ __forceinline__ double R(int m) const { return (double)rho[(m - 1) * stride]; }
  __device__ __forceinline__ double D(int m) const { return (double)d[m - 1]; }
  // sphere(ifunc, 1): rho = rtp * btp**(-5 | -2.275)
  __device__ __forceinline__ void apply_fac(const float *fac) {
    if (fac)
      for (int i = 1; i <= mmax; i++) rho[(i - 1) * stride] = rho[(i - 1) * stride] * fac[i - 1];
  }
};

// On-the-fly stack of one finite-difference VARIANT of a column (depthkernel, CalSurfG.f90:76-160).
// A variant differs from its column's base model in ONE parameter of ONE depth node, i.e. in the few
// refined layers interpolated from that node.  The base stack is stored once per column in shared
// memory (396 B per column instead of per thread: the per-thread stacks were what limited the kernel to
// 16 warps per SM); a perturbed layer is recomputed from the node values with the statements of
// refineGrid2LayerMdl + sphere, in the same order, so every value is bit-identical to the stored form.
struct StackOTF {
  const float *a, *b, *rho;      // base stack of this thread's column, contiguous [m-1]; rho already * fac
  const float *d;
  const float *nval;             // node values of the perturbed parameter of this column, [nz] (base)
  const int *lnode;              // refined layer -> upper node
  const float *wnum, *wden, *fac;
  const double *tmpfac;
  int pnode, pwhich, psign;      // perturbed node (-1 none), parameter (0 vs, 1 vp, 2 rho), sign (0: -, 1: +)
  int nz, mmax, llw;
  __device__ __forceinline__ bool hit(int k) const {
    if (pnode < 0) return false;
    if (k == mmax - 1) return pnode == nz - 1;
    const int i = lnode[k];
    return i == pnode || i + 1 == pnode;
  }
  __device__ __forceinline__ float node(int i) const {  // node value with the +-0.5 % perturbation
    const float v = nval[i];
    if (i != pnode) return v;
    const float hf = 0.5f * 0.01f;
    return psign ? v + hf * v : v - hf * v;
  }
  __device__ __forceinline__ float layer(int k) const {  // refineGrid2LayerMdl interpolation
    if (k == mmax - 1) return node(nz - 1);
    const int i = lnode[k];
    const float lo = node(i), hi = node(i + 1);
    return lo + wnum[k] * (hi - lo) / wden[k];
  }
  __device__ __forceinline__ float bf(int m) const {
    const int k = m - 1;
    if (pwhich != 0 || !hit(k)) return b[k];
    return (float)((double)layer(k) * tmpfac[k]);
  }
  __device__ __forceinline__ float af(int m) const {
    const int k = m - 1;
    if (pwhich != 1 || !hit(k)) return a[k];
    return (float)((double)layer(k) * tmpfac[k]);
  }
  __device__ __forceinline__ double A(int m) const { return (double)af(m); }
  __device__ __forceinline__ double B(int m) const { return (double)bf(m); }
  __device__ __forceinline__ double R(int m) const {
    const int k = m - 1;
    if (pwhich != 2 || !hit(k)) return (double)rho[k];
    return (double)(layer(k) * fac[k]);
  }
  __device__ __forceinline__ double D(int m) const { return (double)d[m - 1]; }
  __device__ __forceinline__ void apply_fac(const float *) {}  // folded into the base stack / R()
};

__device__ __forceinline__ double dsign1(double x) { return copysign(1.0, x); }

// surfdisp96.f:704-761
template <class STK>
__device__ double dltar1(double wvno, double omega, const STK &L) {
  const int mmax = L.mmax;
  double beta1 = L.B(mmax);
  double rho1 = L.R(mmax);
  double xkb = omega / beta1;
  double wvnop = wvno + xkb;
  double wvnom = fabs(wvno - xkb);
  double rb = sqrt(wvnop * wvnom);
  double e1 = rho1 * rb;
  double e2 = 1.0 / (beta1 * beta1);
  for (int m = mmax - 1; m >= L.llw; m--) {
    beta1 = L.B(m);
    rho1 = L.R(m);
    const double xmu = rho1 * beta1 * beta1;
    xkb = omega / beta1;
    wvnop = wvno + xkb;
    wvnom = fabs(wvno - xkb);
    rb = sqrt(wvnop * wvnom);
    const double dm = L.D(m);
    const double q = dm * rb;
    double sinq, cosq, y, z;
    if (wvno < xkb) {
      sincos(q, &sinq, &cosq);
      y = sinq / rb;
      z = -rb * sinq;
    } else if (wvno == xkb) {
      cosq = 1.0;
      y = dm;
      z = 0.0;
    } else {
      double fac = 0.0;
      if (q < 16.0) fac = exp(-2.0 * q);
      cosq = (1.0 + fac) * 0.5;
      sinq = (1.0 - fac) * 0.5;
      y = sinq / rb;
      z = rb * sinq;
    }
    const double e10 = e1 * cosq + e2 * xmu * z;
    const double e20 = e1 * y / xmu + e2 * cosq;
    double xnor = fabs(e10);
    const double ynor = fabs(e20);
    if (ynor > xnor) xnor = ynor;
    if (xnor < 1.0e-40) xnor = 1.0;
    e1 = e10 / xnor;
    e2 = e20 / xnor;
  }
  return e1;
}

struct VarOut {
  double w, cosp, a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz;
};

// surfdisp96.f:868-987
__device__ __forceinline__ void var_psv(double p, double q, double ra, double rb, double wvno,
                                        double xka, double xkb, double dpth, VarOut &o) {
  double w = 0, x = 0, cosp = 0, sinp, y = 0, z = 0, cosq = 0, sinq, fac;
  double pex = 0.0, sex = 0.0;
  if (wvno < xka) {
    sincos(p, &sinp, &cosp);
    w = sinp / ra;
    x = -ra * sinp;
  } else if (wvno == xka) {
    cosp = 1.0;
    w = dpth;
    x = 0.0;
  } else if (wvno > xka) {
    pex = p;
    fac = 0.0;
    if (p < 16.0) fac = exp(-2.0 * p);
    cosp = (1.0 + fac) * 0.5;
    sinp = (1.0 - fac) * 0.5;
    w = sinp / ra;
    x = ra * sinp;
  }
  if (wvno < xkb) {
    sincos(q, &sinq, &cosq);
    y = sinq / rb;
    z = -rb * sinq;
  } else if (wvno == xkb) {
    cosq = 1.0;
    y = dpth;
    z = 0.0;
  } else if (wvno > xkb) {
    sex = q;
    fac = 0.0;
    if (q < 16.0) fac = exp(-2.0 * q);
    cosq = (1.0 + fac) * 0.5;
    sinq = (1.0 - fac) * 0.5;
    y = sinq / rb;
    z = rb * sinq;
  }
  const double exa = pex + sex;
  double a0 = 0.0;
  if (exa < 60.0) a0 = exp(-exa);
  o.a0 = a0;
  o.cpcq = cosp * cosq;
  o.cpy = cosp * y;
  o.cpz = cosp * z;
  o.cqw = cosq * w;
  o.cqx = cosq * x;
  o.xy = x * y;
  o.xz = x * z;
  o.wy = w * y;
  o.wz = w * z;
  o.w = w;
  o.cosp = cosp;
}

// surfdisp96.f:767-864 with dnka (:1018-1062) and normc (:989-1014) inlined; the 5x5 Dunkin
// matrix is formed column by column so only one column is live at a time.
template <class STK>
__device__ double dltar4(double wvno, double omga, const STK &L) {
  const int mmax = L.mmax;
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  const double wvno2 = wvno * wvno;
  double xka = omega / L.A(mmax);
  double xkb = omega / L.B(mmax);
  double wvnop = wvno + xka;
  double wvnom = fabs(wvno - xka);
  double ra = sqrt(wvnop * wvnom);
  wvnop = wvno + xkb;
  wvnom = fabs(wvno - xkb);
  double rb = sqrt(wvnop * wvnom);
  double t = L.B(mmax) / omega;
  double gammk = 2.0 * t * t;
  double gam = gammk * wvno2;
  double gamm1 = gam - 1.0;
  double rho1 = L.R(mmax);
  double e1 = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
  double e2 = -rho1 * ra;
  double e3 = rho1 * (gamm1 - gammk * ra * rb);
  double e4 = rho1 * rb;
  double e5 = wvno2 - ra * rb;
  VarOut v;
  for (int m = mmax - 1; m >= L.llw; m--) {
    xka = omega / L.A(m);
    xkb = omega / L.B(m);
    t = L.B(m) / omega;
    gammk = 2.0 * t * t;
    gam = gammk * wvno2;
    wvnop = wvno + xka;
    wvnom = fabs(wvno - xka);
    ra = sqrt(wvnop * wvnom);
    wvnop = wvno + xkb;
    wvnom = fabs(wvno - xkb);
    rb = sqrt(wvnop * wvnom);
    const double dpth = L.D(m);
    const double rho = L.R(m);
    const double p = ra * dpth;
    const double q = rb * dpth;
    var_psv(p, q, ra, rb, wvno, xka, xkb, dpth, v);
    // ---- dnka
    const double one = 1.0, two = 2.0;
    gamm1 = gam - one;
    const double twgm1 = gam + gamm1;
    const double gmgmk = gam * gammk;
    const double gmgm1 = gam * gamm1;
    const double gm1sq = gamm1 * gamm1;
    const double rho2 = rho * rho;
    const double a0pq = v.a0 - v.cpcq;
    const double ca11 = v.cpcq - two * gmgm1 * a0pq - gmgmk * v.xz - wvno2 * gm1sq * v.wy;
    const double ca12 = (wvno2 * v.cpy - v.cqx) / rho;
    const double ca13 = -(twgm1 * a0pq + gammk * v.xz + wvno2 * gamm1 * v.wy) / rho;
    const double ca14 = (v.cpz - wvno2 * v.cqw) / rho;
    const double ca15 = -(two * wvno2 * a0pq + v.xz + wvno2 * wvno2 * v.wy) / rho2;
    const double ca21 = (gmgmk * v.cpz - gm1sq * v.cqw) * rho;
    const double ca22 = v.cpcq;
    const double ca23 = gammk * v.cpz - gamm1 * v.cqw;
    const double ca24 = -v.wz;
    const double ca25 = ca14;
    const double ca41 = (gm1sq * v.cpy - gmgmk * v.cqx) * rho;
    const double ca42 = -v.xy;
    const double ca43 = gamm1 * v.cpy - gammk * v.cqx;
    const double ca44 = ca22;
    const double ca45 = ca12;
    const double ca51 = -(two * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * v.xz + gm1sq * gm1sq * v.wy) * rho2;
    const double ca52 = ca41;
    const double ca53 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * v.xz + gamm1 * gm1sq * v.wy) * rho;
    const double ca54 = ca21;
    const double ca55 = ca11;
    const double tt = -two * wvno2;
    const double ca31 = tt * ca53;
    const double ca32 = tt * ca43;
    const double ca33 = v.a0 + two * (v.cpcq - ca11);
    const double ca34 = tt * ca23;
    const double ca35 = tt * ca13;
    // ---- ee(i) = sum_j e(j) ca(j,i), j ascending from 0 (surfdisp96.f:832-838)
    double ee1 = 0.0 + e1 * ca11; ee1 = ee1 + e2 * ca21; ee1 = ee1 + e3 * ca31; ee1 = ee1 + e4 * ca41; ee1 = ee1 + e5 * ca51;
    double ee2 = 0.0 + e1 * ca12; ee2 = ee2 + e2 * ca22; ee2 = ee2 + e3 * ca32; ee2 = ee2 + e4 * ca42; ee2 = ee2 + e5 * ca52;
    double ee3 = 0.0 + e1 * ca13; ee3 = ee3 + e2 * ca23; ee3 = ee3 + e3 * ca33; ee3 = ee3 + e4 * ca43; ee3 = ee3 + e5 * ca53;
    double ee4 = 0.0 + e1 * ca14; ee4 = ee4 + e2 * ca24; ee4 = ee4 + e3 * ca34; ee4 = ee4 + e4 * ca44; ee4 = ee4 + e5 * ca54;
    double ee5 = 0.0 + e1 * ca15; ee5 = ee5 + e2 * ca25; ee5 = ee5 + e3 * ca35; ee5 = ee5 + e4 * ca45; ee5 = ee5 + e5 * ca55;
    // ---- normc
    double t1 = 0.0;
    if (fabs(ee1) > t1) t1 = fabs(ee1);
    if (fabs(ee2) > t1) t1 = fabs(ee2);
    if (fabs(ee3) > t1) t1 = fabs(ee3);
    if (fabs(ee4) > t1) t1 = fabs(ee4);
    if (fabs(ee5) > t1) t1 = fabs(ee5);
    if (t1 < 1.0e-40) t1 = 1.0;
    e1 = ee1 / t1;
    e2 = ee2 / t1;
    e3 = ee3 / t1;
    e4 = ee4 / t1;
    e5 = ee5 / t1;
  }
  if (L.llw != 1) {  // water layer on top, surfdisp96.f:844-860
    xka = omega / L.A(1);
    wvnop = wvno + xka;
    wvnom = fabs(wvno - xka);
    ra = sqrt(wvnop * wvnom);
    const double dpth = L.D(1);
    rho1 = L.R(1);
    const double p = ra * dpth;
    const double znul = 1.0e-05;
    var_psv(p, znul, ra, znul, wvno, xka, znul, dpth, v);
    const double w0 = -rho1 * v.w;
    return v.cosp * e1 + w0 * e2;
  }
  return e1;
}

template <class STK>
__device__ __forceinline__ double dltar(double wvno, double omega, int kk, const STK &L) {
  return kk == 1 ? dltar1(wvno, omega, L) : dltar4(wvno, omega, L);
}

// surfdisp96.f:551-668 (half inlined, :670-680)
template <class STK>
__device__ double nevill(double t, double c1, double c2, double del1, double del2, int ifunc,
                         const STK &L, double twopi) {
  double x[12], y[12];
  double c3, del3;
  int m = 1;
  const double omega = twopi / t;
  c3 = 0.5 * (c1 + c2);
  del3 = dltar(omega / c3, omega, ifunc, L);
  int nev = 1;
  int nctrl = 1;
  for (;;) {
    nctrl = nctrl + 1;
    if (nctrl >= 100) break;
    if (c3 < fmin(c1, c2) || c3 > fmax(c1, c2)) {
      nev = 0;
      c3 = 0.5 * (c1 + c2);
      del3 = dltar(omega / c3, omega, ifunc, L);
    }
    const double s13 = del1 - del3;
    const double s32 = del3 - del2;
    if (dsign1(del3) * dsign1(del1) < 0.0) {
      c2 = c3;
      del2 = del3;
    } else {
      c1 = c3;
      del1 = del3;
    }
    if (fabs(c1 - c2) <= 1.0e-6 * c1) break;
    if (dsign1(s13) != dsign1(s32)) nev = 0;
    const double ss1 = fabs(del1);
    const double s1 = (double)0.01f * ss1;
    const double ss2 = fabs(del2);
    const double s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || nev == 0) {
      c3 = 0.5 * (c1 + c2);
      del3 = dltar(omega / c3, omega, ifunc, L);
      nev = 1;
      m = 1;
    } else {
      if (nev == 2) {
        x[m + 1] = c3;
        y[m + 1] = del3;
      } else {
        x[1] = c1;
        y[1] = del1;
        x[2] = c2;
        y[2] = del2;
        m = 1;
      }
      bool fallback = false;
      for (int kk = 1; kk <= m; kk++) {
        const int j = m - kk + 1;
        const double denom = y[m + 1] - y[j];
        if (fabs(denom) < 1.0e-10 * fabs(y[m + 1])) {
          fallback = true;
          break;
        }
        x[j] = (-y[j] * x[j + 1] + y[m + 1] * x[j]) / denom;
      }
      if (!fallback) {
        c3 = x[1];
        del3 = dltar(omega / c3, omega, ifunc, L);
        nev = 2;
        m = m + 1;
        if (m > 10) m = 10;
      } else {
        c3 = 0.5 * (c1 + c2);
        del3 = dltar(omega / c3, omega, ifunc, L);
        nev = 1;
        m = 1;
      }
    }
  }
  return c3;
}

// surfdisp96.f:384-476
template <class STK>
__device__ void getsol(double t1, double &c1, double clow, double dc, double cm, float betmx,
                       int &iret, int ifunc, int ifirst, const STK &L, double &del1st) {
  const double twopi = 2.0 * 3.141592653589793;
  double omega = twopi / t1;
  double wvno = omega / c1;
  double del1 = dltar(wvno, omega, ifunc, L);
  if (ifirst == 1) del1st = del1;
  const double plmn = dsign1(del1st) * dsign1(del1);
  int idir;
  if (ifirst == 1)
    idir = +1;
  else if (plmn >= 0.0)
    idir = +1;
  else
    idir = -1;
  double c2, del2;
  for (;;) {
    if (idir > 0)
      c2 = c1 + dc;
    else
      c2 = c1 - dc;
    if (c2 <= clow) {
      idir = +1;
      c1 = clow;
      continue;
    }
    omega = twopi / t1;
    wvno = omega / c2;
    del2 = dltar(wvno, omega, ifunc, L);
    if (dsign1(del1) != dsign1(del2)) break;
    c1 = c2;
    del1 = del2;
    if (c1 < cm) {
      iret = -1;
      return;
    }
    if (c1 >= ((double)betmx + dc)) {
      iret = -1;
      return;
    }
  }
  const double cn = nevill(t1, c1, c2, del1, del2, ifunc, L, twopi);
  c1 = cn;
  if (c1 > (double)betmx) {
    iret = -1;
    return;
  }
  iret = 1;
}

// surfdisp96.f:361-382 (REAL*4)
__device__ float gtsolh(float a, float b) {
  float c = 0.95f * b;
  for (int i = 1; i <= 5; i++) {
    const float gamma = b / a;
    const float kappa = c / b;
    const float k2 = kappa * kappa;
    const float gk = gamma * kappa;
    const float gk2 = gk * gk;
    const float fac1 = sqrtf(1.0f - gk2);
    const float fac2 = sqrtf(1.0f - k2);
    const float tk = 2.0f - k2;
    const float fr = tk * tk - 4.0f * fac1 * fac2;
    float frp = -(4.0f * (2.0f - k2) * kappa) + 4.0f * fac2 * gamma * gamma * kappa / fac1 +
                4.0f * fac1 * kappa / fac2;
    frp = frp / b;
    c = c - fr / frp;
  }
  return c;
}

// body of surfdisp96 (surfdisp96.f:92-353) for a stack already earth-flattened in shared
// memory (a, b hold the transformed velocities, rtp the untransformed density).
// smem arrays: sa, sb, srho (work), with rho rewritten per wave type from rtp * fac.
// c_prev[] (kmax doubles, thread-private global scratch) holds c(k) for higher modes.
template <class STK>
__device__ void surfdisp_core_t(STK &L, const float *fac, int ifunc, int mode, int igr, int kmax, const double *t,
                                double *cg, double *cwork) {
  const int mmax = L.mmax;
  L.llw = 1;
  if (L.bf(1) <= 0.0f) L.llw = 2;
  int jmn = 1, jsol = 1;
  float betmx = -1.e20f, betmn = 1.e20f;
  for (int i = 1; i <= mmax; i++) {
    const float bi = L.bf(i), ai = L.af(i);
    if (bi > 0.01f && bi < betmn) {
      betmn = bi;
      jmn = i;
      jsol = 1;
    } else if (bi <= 0.01f && ai < betmn) {
      betmn = ai;
      jmn = i;
      jsol = 0;
    }
    if (bi > betmx) betmx = bi;
  }
  L.apply_fac(fac);
  const float ddc = 0.005f, sone = 1.500f, h = 0.005f;
  const double one = 1.0e-2;
  const double onea = (double)sone;
  float cc1;
  if (jsol == 0)
    cc1 = betmn;
  else
    cc1 = gtsolh(L.af(jmn), L.bf(jmn));
  cc1 = .95f * cc1;
  cc1 = .90f * cc1;
  const double cc = (double)cc1;
  const double dc = fabs((double)ddc);
  double c1 = cc;
  const double cm = cc;
  double del1st = 0.0;
  double cprev = 0.0;  // c(k-1) of the current mode
  int ift = 999;
  for (int iq = 1; iq <= mode; iq++) {
    int k;
    bool failed = false;
    for (k = 1; k <= kmax; k++) {
      if (k >= ift) {
        failed = true;
        break;
      }
      double t1 = t[k - 1];
      float t1a, t1b = 0.0f;
      if (igr > 0) {
        t1a = (float)(t1 / (double)(1.f + h));
        t1b = (float)(t1 / (double)(1.f - h));
        t1 = (double)t1a;
      } else {
        t1a = (float)t1;
      }
      double clow;
      int ifirst;
      if (k == 1 && iq == 1) {
        c1 = cc;
        clow = cc;
        ifirst = 1;
      } else if (k == 1 && iq > 1) {
        c1 = cwork[0] + one * dc;
        clow = c1;
        ifirst = 1;
      } else if (k > 1 && iq > 1) {
        ifirst = 0;
        clow = cwork[k - 1] + one * dc;
        c1 = cprev;
        if (c1 < clow) c1 = clow;
      } else {
        ifirst = 0;
        c1 = cprev - onea * dc;
        clow = cm;
      }
      int iret;
      getsol(t1, c1, clow, dc, cm, betmx, iret, ifunc, ifirst, L, del1st);
      if (iret == -1) {
        failed = true;
        break;
      }
      const double ck = c1;
      cprev = ck;
      if (mode > 1) cwork[k - 1] = ck;
      if (igr > 0) {
        t1 = (double)t1b;
        ifirst = 0;
        // cb(k) is zero on first use for every mode-1 period (surfdisp96.f:209-212)
        clow = (mode > 1 ? cwork[kmax + k - 1] : 0.0) + one * dc;
        c1 = c1 - onea * dc;
        getsol(t1, c1, clow, dc, cm, betmx, iret, ifunc, ifirst, L, del1st);
        if (iret == -1) c1 = ck;
        if (mode > 1) cwork[kmax + k - 1] = c1;
      } else {
        c1 = 0.0;
      }
      const float cc0 = (float)ck;
      const float cc1b = (float)c1;
      if (igr == 0) {
        cg[k - 1] = (double)cc0;
      } else {
        const float gvel = (1.0f / t1a - 1.0f / t1b) / (1.0f / (t1a * cc0) - 1.0f / (t1b * cc1b));
        cg[k - 1] = (double)gvel;
      }
    }
    if (failed) {
      ift = k;
      for (int i = k; i <= kmax; i++) cg[i - 1] = 0.0;
    }
  }
}

// per-thread shared-memory stacks (surfdisp96 drop-in, first-generation column kernel)
__device__ void surfdisp_core(float *sa, float *sb, float *srho, const float *d, const float *fac,
                              int stride, int mmax, int ifunc, int mode, int igr, int kmax,
                              const double *t, double *cg, double *cwork) {
  Stack L;
  L.a = sa;
  L.b = sb;
  L.rho = srho;
  L.d = d;
  L.stride = stride;
  L.mmax = mmax;
  L.llw = 1;
  surfdisp_core_t(L, fac, ifunc, mode, igr, kmax, t, cg, cwork);
}

// gfortran powi trees
__device__ __forceinline__ float pow3f(float x) { return x * (x * x); }
__device__ __forceinline__ float pow4f(float x) { const float x2 = x * x; return x2 * x2; }
__device__ __forceinline__ float pow5f(float x) { const float x2 = x * x; return x2 * (x * x2); }

// ------------------------------------------------------------------------------------------
// kernel A: column variants (depthkernel / caldespersion)
// ------------------------------------------------------------------------------------------
constexpr int kDispBlock = 64;
constexpr int kDispOtfDefault = 0;  // first generation: the on-the-fly variants measured 1.7-1.9x slower (see k_disp_columns_otf)

__global__ void __launch_bounds__(kDispBlock)
k_disp_columns(const float *__restrict__ vel, int nx, int ny, int nz, int nvar, int col0, int ncol_batch,
               const float *__restrict__ dflat, const double *__restrict__ tmpfac,
               const float *__restrict__ fac, const int *__restrict__ lnode,
               const float *__restrict__ wnum, const float *__restrict__ wden, int rmax, int ifunc,
               int igr, int kmax, const double *__restrict__ t, double *__restrict__ cgbuf) {
  extern __shared__ float smem[];
  const int stride = kDispBlock;
  float *sa = smem + threadIdx.x;
  float *sb = sa + (size_t)rmax * stride;
  float *srho = sb + (size_t)rmax * stride;
  float *sd = smem + (size_t)3 * rmax * stride;    // shared tables
  float *sfac = sd + rmax;
  for (int i = threadIdx.x; i < rmax; i += blockDim.x) {
    sd[i] = dflat[i];
    sfac[i] = fac[i];
  }
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)ncol_batch * nvar) return;
  const int col = col0 + (int)(gid / nvar);
  const int var = (int)(gid % nvar);
  // variant decoding (CalSurfG.f90:76-160): 0 = base; 1 + 6*i + j
  const int pnode = var == 0 ? -1 : (var - 1) / 6;
  const int pkind = var == 0 ? -1 : (var - 1) % 6;
  const size_t plane = (size_t)nx * ny;
  // node model -> refined, flattened layer stack
  float vs_lo, vp_lo, rho_lo, vs_hi, vp_hi, rho_hi;
  int cur = -1;
  auto node_model = [&](int i, float &vs, float &vp, float &rho) {
    const float v = vel[(size_t)i * plane + col];
    vs = v;
    vp = 0.9409f + 2.0947f * v - 0.8206f * (v * v) + 0.2683f * pow3f(v) - 0.0251f * pow4f(v);
    rho = 1.6612f * vp - 0.4721f * (vp * vp) + 0.0671f * pow3f(vp) - 0.0043f * pow4f(vp) +
          0.000106f * pow5f(vp);
    if (i == pnode) {  // +-0.5 % perturbation of one parameter at node i
      const float hf = 0.5f * 0.01f;
      if (pkind == 0) vs = v - hf * v;
      if (pkind == 1) vs = v + hf * v;
      if (pkind == 2) vp = vp - hf * vp;
      if (pkind == 3) vp = vp + hf * vp;
      if (pkind == 4) rho = rho - hf * rho;
      if (pkind == 5) rho = rho + hf * rho;
    }
  };
  for (int k = 0; k < rmax; k++) {
    const int i = lnode[k];
    float rvs, rvp, rrho;
    if (k == rmax - 1) {  // half space = last node
      node_model(nz - 1, rvs, rvp, rrho);
    } else {
      if (i != cur) {
        node_model(i, vs_lo, vp_lo, rho_lo);
        node_model(i + 1, vs_hi, vp_hi, rho_hi);
        cur = i;
      }
      rvp = vp_lo + wnum[k] * (vp_hi - vp_lo) / wden[k];
      rvs = vs_lo + wnum[k] * (vs_hi - vs_lo) / wden[k];
      rrho = rho_lo + wnum[k] * (rho_hi - rho_lo) / wden[k];
    }
    const double tm = tmpfac[k];
    sa[(size_t)k * stride] = (float)((double)rvp * tm);
    sb[(size_t)k * stride] = (float)((double)rvs * tm);
    srho[(size_t)k * stride] = rrho;
  }
  double *cg = cgbuf + (size_t)gid * kmax;
  surfdisp_core(sa, sb, srho, sd, sfac, stride, rmax, ifunc, /*mode*/ 1, igr, kmax, t, cg, nullptr);
}

// Second-generation column kernel: on-the-fly variant stacks (StackOTF).  Shared memory per block drops from
// 25.6 KB (one stack per thread) to ~2.5 KB (one base stack per column), so residency is set by registers
// alone: MINB = 8 (128 registers, 16 warps/SM like the first generation), 9, 10 (96 registers, 20 warps) or
// 12 (80 registers, 24 warps).  ncu of the first generation at cfg-3 scale: FP64 pipe 28 % busy, every warp
// waiting on dependent fixed-latency DP chains (profiles/r01_disp_cfg3.md), so more resident warps looked like the
// lever.  MEASURED (gpurun_out/s18_disp_ab.log, Rc, 17161 columns x 55 variants x 16 periods): first generation
// 1.23 s; on-the-fly 2.36 s (MINB 8), 2.08 s (10), 2.24 s (12) -- bit-identical output, but the per-access hit test,
// its divergence across the 55 variants of a warp and the spills at 96/80 registers cost more than the extra warps
// return.  Kept as an opt-in (DSURF_DISP_OTF) with its parity test; not the default.
template <int MINB>
__global__ void __launch_bounds__(kDispBlock, MINB)
k_disp_columns_otf(const float *__restrict__ vel, int nx, int ny, int nz, int nvar, int col0, int ncol_batch,
                   const float *__restrict__ dflat, const double *__restrict__ tmpfac,
                   const float *__restrict__ fac, const int *__restrict__ lnode,
                   const float *__restrict__ wnum, const float *__restrict__ wden, int rmax, int ifunc,
                   int igr, int kmax, const double *__restrict__ t, double *__restrict__ cgbuf, int nslots) {
  extern __shared__ double smem_d[];
  double *stmp = smem_d;                               // [rmax]
  float *sd = reinterpret_cast<float *>(stmp + rmax);  // [rmax]
  float *sfac = sd + rmax;
  float *swnum = sfac + rmax;
  float *swden = swnum + rmax;
  int *slnode = reinterpret_cast<int *>(swden + rmax);
  float *slots = reinterpret_cast<float *>(slnode + rmax);  // per column: a[rmax] b[rmax] rho[rmax] nodes[3][nz]
  const int slot_sz = 3 * rmax + 3 * nz;
  for (int i = threadIdx.x; i < rmax; i += blockDim.x) {
    stmp[i] = tmpfac[i];
    sd[i] = dflat[i];
    sfac[i] = fac[i];
    swnum[i] = wnum[i];
    swden[i] = wden[i];
    slnode[i] = lnode[i];
  }
  const long long gid0 = (long long)blockIdx.x * blockDim.x;
  const long long total = (long long)ncol_batch * nvar;
  const int cfirst = (int)(gid0 / nvar);
  const size_t plane = (size_t)nx * ny;
  // node models of the block's columns (CalSurfG.f90:45-52: Brocher's vp(vs), rho(vp))
  for (int it = threadIdx.x; it < nslots * nz; it += blockDim.x) {
    const int sl = it / nz, i = it % nz;
    if (cfirst + sl < ncol_batch) {
      const float v = vel[(size_t)i * plane + col0 + cfirst + sl];
      const float vp = 0.9409f + 2.0947f * v - 0.8206f * (v * v) + 0.2683f * pow3f(v) - 0.0251f * pow4f(v);
      const float rho = 1.6612f * vp - 0.4721f * (vp * vp) + 0.0671f * pow3f(vp) - 0.0043f * pow4f(vp) +
                        0.000106f * pow5f(vp);
      float *nd = slots + (size_t)sl * slot_sz + 3 * rmax;
      nd[i] = v;
      nd[nz + i] = vp;
      nd[2 * nz + i] = rho;
    }
  }
  __syncthreads();
  // base (unperturbed) refined + flattened stacks of the block's columns
  for (int it = threadIdx.x; it < nslots * rmax; it += blockDim.x) {
    const int sl = it / rmax, k = it % rmax;
    if (cfirst + sl < ncol_batch) {
      float *sb0 = slots + (size_t)sl * slot_sz;
      const float *nd = sb0 + 3 * rmax;
      float rvs, rvp, rrho;
      if (k == rmax - 1) {
        rvs = nd[nz - 1];
        rvp = nd[nz + nz - 1];
        rrho = nd[2 * nz + nz - 1];
      } else {
        const int i = slnode[k];
        rvp = nd[nz + i] + swnum[k] * (nd[nz + i + 1] - nd[nz + i]) / swden[k];
        rvs = nd[i] + swnum[k] * (nd[i + 1] - nd[i]) / swden[k];
        rrho = nd[2 * nz + i] + swnum[k] * (nd[2 * nz + i + 1] - nd[2 * nz + i]) / swden[k];
      }
      const double tm = stmp[k];
      sb0[k] = (float)((double)rvp * tm);
      sb0[rmax + k] = (float)((double)rvs * tm);
      sb0[2 * rmax + k] = rrho * sfac[k];
    }
  }
  __syncthreads();
  const long long gid = gid0 + threadIdx.x;
  if (gid >= total) return;
  const int colb = (int)(gid / nvar);
  const int var = (int)(gid % nvar);
  const float *sb0 = slots + (size_t)(colb - cfirst) * slot_sz;
  StackOTF L;
  L.a = sb0;
  L.b = sb0 + rmax;
  L.rho = sb0 + 2 * rmax;
  L.d = sd;
  L.lnode = slnode;
  L.wnum = swnum;
  L.wden = swden;
  L.fac = sfac;
  L.tmpfac = stmp;
  L.nz = nz;
  L.mmax = rmax;
  L.llw = 1;
  L.pnode = var == 0 ? -1 : (var - 1) / 6;           // variant decoding (CalSurfG.f90:76-160)
  const int pkind = var == 0 ? 0 : (var - 1) % 6;    // 0/1 vs -/+, 2/3 vp -/+, 4/5 rho -/+
  L.pwhich = pkind >> 1;
  L.psign = pkind & 1;
  L.nval = sb0 + 3 * rmax + (size_t)L.pwhich * nz;
  double *cg = cgbuf + (size_t)gid * kmax;
  surfdisp_core_t(L, nullptr, ifunc, /*mode*/ 1, igr, kmax, t, cg, nullptr);
}

// pv and the three finite-difference kernels from the variant curves (CalSurfG.f90:60,91-93,
// 121-123,148-150,163-165).  Thread per (col, period).
__global__ void k_disp_finish(const float *__restrict__ vel, int nx, int ny, int nz, int nvar, int col0,
                              int ncol_batch, int kmax, const double *__restrict__ cgbuf,
                              double *__restrict__ pv, double *__restrict__ sen_vs,
                              double *__restrict__ sen_vp, double *__restrict__ sen_rho) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)ncol_batch * kmax) return;
  const int cb = (int)(gid / kmax), n = (int)(gid % kmax);
  const int col = col0 + cb;
  const size_t ncol = (size_t)nx * ny;
  const double *cgc = cgbuf + (size_t)cb * nvar * kmax;
  pv[(size_t)n * ncol + col] = cgc[n];
  if (nvar == 1 || !sen_vs) return;
  const float dln = 0.01f;
  for (int i = 0; i < nz; i++) {
    const float v = vel[(size_t)i * ncol + col];
    const float vp = 0.9409f + 2.0947f * v - 0.8206f * (v * v) + 0.2683f * pow3f(v) - 0.0251f * pow4f(v);
    const float rho = 1.6612f * vp - 0.4721f * (vp * vp) + 0.0671f * pow3f(vp) - 0.0043f * pow4f(vp) +
                      0.000106f * pow5f(vp);
    const double *c = cgc + (size_t)(1 + 6 * i) * kmax + n;
    const size_t o = ((size_t)i * kmax + n) * ncol + col;
    sen_vs[o] = (c[(size_t)1 * kmax] - c[0]) / (double)(dln * v);
    sen_vp[o] = (c[(size_t)3 * kmax] - c[(size_t)2 * kmax]) / (double)(dln * vp);
    sen_rho[o] = (c[(size_t)5 * kmax] - c[(size_t)4 * kmax]) / (double)(dln * rho);
  }
}

// ------------------------------------------------------------------------------------------
// kernel B: explicit layer stacks (the surfdisp96 drop-in; thk shared by all models)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDispBlock)
k_disp_stacks(const float *__restrict__ vp, const float *__restrict__ vs, const float *__restrict__ rho,
              int nmodel, int rmax, const float *__restrict__ dflat, const double *__restrict__ tmpfac,
              const float *__restrict__ fac, int ifunc, int mode, int igr, int kmax,
              const double *__restrict__ t, double *__restrict__ cg, double *__restrict__ cwork) {
  extern __shared__ float smem[];
  const int stride = kDispBlock;
  float *sa = smem + threadIdx.x;
  float *sb = sa + (size_t)rmax * stride;
  float *srho = sb + (size_t)rmax * stride;
  float *sd = smem + (size_t)3 * rmax * stride;
  float *sfac = sd + rmax;
  for (int i = threadIdx.x; i < rmax; i += blockDim.x) {
    sd[i] = dflat[i];
    sfac[i] = fac ? fac[i] : 1.0f;
  }
  __syncthreads();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nmodel) return;
  for (int k = 0; k < rmax; k++) {
    const double tm = tmpfac[k];
    sa[(size_t)k * stride] = (float)((double)vp[(size_t)gid * rmax + k] * tm);
    sb[(size_t)k * stride] = (float)((double)vs[(size_t)gid * rmax + k] * tm);
    srho[(size_t)k * stride] = rho[(size_t)gid * rmax + k];
  }
  surfdisp_core(sa, sb, srho, sd, fac ? sfac : nullptr, stride, rmax, ifunc, mode, igr, kmax, t,
                cg + (size_t)gid * kmax, cwork ? cwork + (size_t)gid * 2 * kmax : nullptr);
}

static size_t disp_smem_bytes(int rmax) { return ((size_t)3 * rmax * kDispBlock + 2 * rmax) * sizeof(float); }

// Runs depthkernel/caldespersion for all columns of d_vel; outputs on device.
int run_dispersion(cudaStream_t st, const float *d_vel, int nx, int ny, int nz, const LayerTablesDev &T,
                   int iwave, int igr, int kmax, const double *d_t, bool want_kernels, double *d_pv,
                   double *d_sen_vs, double *d_sen_vp, double *d_sen_rho, DevBuf<double> &cgbuf) {
  const int nvar = want_kernels ? 1 + 6 * nz : 1;
  const int ncol = nx * ny;
  const int ifunc = (iwave == 1) ? 1 : 2;
  const size_t smem = disp_smem_bytes(T.rmax);
  DS_CUDA(cudaFuncSetAttribute(k_disp_columns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // column batches bound the variant-curve scratch (8*kmax*nvar bytes per column)
  const size_t per_col = (size_t)nvar * kmax;
  int batch = (int)std::min<size_t>(ncol, std::max<size_t>(1, ((size_t)1 << 28) / (per_col * sizeof(double))));
  if (cgbuf.reserve((size_t)batch * per_col)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (dispersion scratch)");
    return DSURF_ERR_CUDA;
  }
  const float *fac = (ifunc == 1) ? T.facL : T.facR;
  for (int c0 = 0; c0 < ncol; c0 += batch) {
    const int nb = std::min(batch, ncol - c0);
    const long long nthreads = (long long)nb * nvar;
    const int grid = (int)((nthreads + kDispBlock - 1) / kDispBlock);
    // DSURF_DISP_OTF = resident blocks per SM of the on-the-fly-stack kernel (8, 9, 10, 12); 0 = first generation
    static const int otf = getenv("DSURF_DISP_OTF") ? atoi(getenv("DSURF_DISP_OTF")) : kDispOtfDefault;
    if (otf > 0 && nvar > 1) {
      const int nslots = kDispBlock / nvar + 2;
      const size_t smo = (size_t)T.rmax * 8 + (size_t)5 * T.rmax * 4 + (size_t)nslots * (3 * T.rmax + 3 * nz) * 4;
#define DISP_OTF_LAUNCH(MB)                                                                                        \
  k_disp_columns_otf<MB><<<grid, kDispBlock, smo, st>>>(d_vel, nx, ny, nz, nvar, c0, nb, T.dflat, T.tmp, fac, T.node, \
                                                        T.wnum, T.wden, T.rmax, ifunc, igr, kmax, d_t, cgbuf.p, nslots)
      if (otf >= 12)
        DISP_OTF_LAUNCH(12);
      else if (otf >= 10)
        DISP_OTF_LAUNCH(10);
      else if (otf == 9)
        DISP_OTF_LAUNCH(9);
      else
        DISP_OTF_LAUNCH(8);
#undef DISP_OTF_LAUNCH
    } else {
      k_disp_columns<<<grid, kDispBlock, smem, st>>>(d_vel, nx, ny, nz, nvar, c0, nb, T.dflat, T.tmp, fac,
                                                     T.node, T.wnum, T.wden, T.rmax, ifunc, igr, kmax, d_t,
                                                     cgbuf.p);
    }
    const long long nf = (long long)nb * kmax;
    k_disp_finish<<<(int)((nf + 127) / 128), 128, 0, st>>>(d_vel, nx, ny, nz, nvar, c0, nb, kmax, cgbuf.p,
                                                           d_pv, d_sen_vs, d_sen_vp, d_sen_rho);
  }
  DS_CUDA(cudaGetLastError());
  return DSURF_OK;
}

int upload_tables(const LayerTables &T, LayerTablesDev &D) {
  D.rmax = T.rmax;
  const int r = T.rmax;
  if (D.b_dflat.reserve(r) || D.b_tmp.reserve(r) || D.b_facR.reserve(r) || D.b_facL.reserve(r) ||
      D.b_node.reserve(r) || D.b_wnum.reserve(r) || D.b_wden.reserve(r))
    return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemcpy(D.b_dflat.p, T.dflat.data(), r * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(D.b_tmp.p, T.tmp.data(), r * sizeof(double), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(D.b_facR.p, T.facR.data(), r * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(D.b_facL.p, T.facL.data(), r * sizeof(float), cudaMemcpyHostToDevice));
  if ((int)T.node.size() == r) {
    DS_CUDA(cudaMemcpy(D.b_node.p, T.node.data(), r * sizeof(int), cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(D.b_wnum.p, T.wnum.data(), r * sizeof(float), cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(D.b_wden.p, T.wden.data(), r * sizeof(float), cudaMemcpyHostToDevice));
  }
  D.dflat = D.b_dflat.p;
  D.tmp = D.b_tmp.p;
  D.facR = D.b_facR.p;
  D.facL = D.b_facL.p;
  D.node = D.b_node.p;
  D.wnum = D.b_wnum.p;
  D.wden = D.b_wden.p;
  return DSURF_OK;
}

}  // namespace dsurf

using namespace dsurf;

extern "C" int dsurf_depthkernel(int nx, int ny, int nz, const float *vel, double *pv, double *sen_vs,
                                 double *sen_vp, double *sen_rho, int iwave, int igr, int kmax,
                                 const double *t, const float *depz, float minthk) {
  DS_CHECK(ensure_device());
  if (!vel || !pv || !t || !depz || kmax < 1 || nz < 2) return DSURF_ERR_BAD_ARG;
  const bool want = sen_vs && sen_vp && sen_rho;
  LayerTables T;
  make_layer_tables(depz, nz, minthk, T);
  LayerTablesDev D;
  DS_CHECK(upload_tables(T, D));
  const size_t ncol = (size_t)nx * ny;
  DevBuf<float> dvel;
  DevBuf<double> dt, dpv, ds0, ds1, ds2, cgbuf;
  if (dvel.reserve(ncol * nz) || dt.reserve(kmax) || dpv.reserve(ncol * kmax)) return DSURF_ERR_CUDA;
  if (want && (ds0.reserve(ncol * kmax * nz) || ds1.reserve(ncol * kmax * nz) || ds2.reserve(ncol * kmax * nz)))
    return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemcpy(dvel.p, vel, ncol * nz * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dt.p, t, kmax * sizeof(double), cudaMemcpyHostToDevice));
  DS_CHECK(run_dispersion(0, dvel.p, nx, ny, nz, D, iwave, igr, kmax, dt.p, want, dpv.p, ds0.p, ds1.p, ds2.p, cgbuf));
  DS_CUDA(cudaMemcpy(pv, dpv.p, ncol * kmax * sizeof(double), cudaMemcpyDeviceToHost));
  if (want) {
    DS_CUDA(cudaMemcpy(sen_vs, ds0.p, ncol * kmax * nz * sizeof(double), cudaMemcpyDeviceToHost));
    DS_CUDA(cudaMemcpy(sen_vp, ds1.p, ncol * kmax * nz * sizeof(double), cudaMemcpyDeviceToHost));
    DS_CUDA(cudaMemcpy(sen_rho, ds2.p, ncol * kmax * nz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return DSURF_OK;
}

extern "C" int dsurf_surfdisp96_batch(int nmodel, const float *thkm, const float *vpm, const float *vsm,
                                      const float *rhom, int nlayer, int iflsph, int iwave, int igr,
                                      int kmax, const double *t, double *cg) {
  DS_CHECK(ensure_device());
  if (nmodel < 1 || nlayer < 2 || nlayer > 200 || kmax < 1 || kmax > 80) return DSURF_ERR_BAD_ARG;
  LayerTables T;
  make_layer_tables_from_thk(thkm, nlayer, iflsph, T);
  LayerTablesDev D;
  DS_CHECK(upload_tables(T, D));
  const size_t tot = (size_t)nmodel * nlayer;
  DevBuf<float> dvp, dvs, drho;
  DevBuf<double> dt, dcg;
  if (dvp.reserve(tot) || dvs.reserve(tot) || drho.reserve(tot) || dt.reserve(kmax) || dcg.reserve((size_t)nmodel * kmax))
    return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemcpy(dvp.p, vpm, tot * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dvs.p, vsm, tot * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(drho.p, rhom, tot * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dt.p, t, kmax * sizeof(double), cudaMemcpyHostToDevice));
  const int ifunc = (iwave == 1) ? 1 : 2;
  const size_t smem = disp_smem_bytes(nlayer);
  DS_CUDA(cudaFuncSetAttribute(k_disp_stacks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float *fac = (iflsph == 1) ? ((ifunc == 1) ? D.facL : D.facR) : nullptr;
  k_disp_stacks<<<(nmodel + kDispBlock - 1) / kDispBlock, kDispBlock, smem>>>(
      dvp.p, dvs.p, drho.p, nmodel, nlayer, D.dflat, D.tmp, fac, ifunc, 1, igr, kmax, dt.p, dcg.p, nullptr);
  DS_CUDA(cudaGetLastError());
  DS_CUDA(cudaMemcpy(cg, dcg.p, (size_t)nmodel * kmax * sizeof(double), cudaMemcpyDeviceToHost));
  return DSURF_OK;
}

extern "C" int dsurf_surfdisp96(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                                int nlayer, int iflsph, int iwave, int mode, int igr, int kmax,
                                const double *t, double *cg) {
  if (mode == 1)
    return dsurf_surfdisp96_batch(1, thkm, vpm, vsm, rhom, nlayer, iflsph, iwave, igr, kmax, t, cg);
  // higher modes: same kernel with per-model scratch for c(k), cb(k)
  DS_CHECK(ensure_device());
  if (nlayer < 2 || nlayer > 200 || kmax < 1 || kmax > 80 || mode < 1) return DSURF_ERR_BAD_ARG;
  LayerTables T;
  make_layer_tables_from_thk(thkm, nlayer, iflsph, T);
  LayerTablesDev D;
  DS_CHECK(upload_tables(T, D));
  DevBuf<float> dvp, dvs, drho;
  DevBuf<double> dt, dcg, dw;
  if (dvp.reserve(nlayer) || dvs.reserve(nlayer) || drho.reserve(nlayer) || dt.reserve(kmax) ||
      dcg.reserve(kmax) || dw.reserve(2 * kmax))
    return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemcpy(dvp.p, vpm, nlayer * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dvs.p, vsm, nlayer * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(drho.p, rhom, nlayer * sizeof(float), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemcpy(dt.p, t, kmax * sizeof(double), cudaMemcpyHostToDevice));
  DS_CUDA(cudaMemset(dw.p, 0, 2 * kmax * sizeof(double)));
  const int ifunc = (iwave == 1) ? 1 : 2;
  const size_t smem = disp_smem_bytes(nlayer);
  DS_CUDA(cudaFuncSetAttribute(k_disp_stacks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float *fac = (iflsph == 1) ? ((ifunc == 1) ? D.facL : D.facR) : nullptr;
  k_disp_stacks<<<1, kDispBlock, smem>>>(dvp.p, dvs.p, drho.p, 1, nlayer, D.dflat, D.tmp, fac, ifunc, mode,
                                         igr, kmax, dt.p, dcg.p, dw.p);
  DS_CUDA(cudaGetLastError());
  DS_CUDA(cudaMemcpy(cg, dcg.p, kmax * sizeof(double), cudaMemcpyDeviceToHost));
  return DSURF_OK;
}

extern "C" void surfdisp96_(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                            const int *nlayer, const int *iflsph, const int *iwave, const int *mode,
                            const int *igr, const int *kmax, const double *t, double *cg) {
  int rc = dsurf_surfdisp96(thkm, vpm, vsm, rhom, *nlayer, *iflsph, *iwave, *mode, *igr, *kmax, t, cg);
  if (rc != DSURF_OK) {
    fprintf(stderr, "surfdisp96: libdsurf_b200 error %d: %s\n", rc, dsurf_last_error());
    exit(1);
  }
}

extern "C" void depthkernel_(const int *nx, const int *ny, const int *nz, const float *vel, double *pvRc,
                             double *sen_vsRc, double *sen_vpRc, double *sen_rhoRc, const int *iwave,
                             const int *igr, const int *kmaxRc, const double *tRc, const float *depz,
                             const float *minthk) {
  int rc = dsurf_depthkernel(*nx, *ny, *nz, vel, pvRc, sen_vsRc, sen_vpRc, sen_rhoRc, *iwave, *igr, *kmaxRc,
                             tRc, depz, *minthk);
  if (rc != DSURF_OK) {
    fprintf(stderr, "depthkernel: libdsurf_b200 error %d: %s\n", rc, dsurf_last_error());
    exit(1);
  }
}

extern "C" void caldespersion_(const int *nx, const int *ny, const int *nz, const float *vel, double *pvRc,
                               const int *iwave, const int *igr, const int *kmaxRc, const double *tRc,
                               const float *depz, const float *minthk) {
  int rc = dsurf_depthkernel(*nx, *ny, *nz, vel, pvRc, nullptr, nullptr, nullptr, *iwave, *igr, *kmaxRc, tRc,
                             depz, *minthk);
  if (rc != DSURF_OK) {
    fprintf(stderr, "caldespersion: libdsurf_b200 error %d: %s\n", rc, dsurf_last_error());
    exit(1);
  }
}
