// ORACLE (test infrastructure, see oracle.h) -- CPU restatement of src/surfdisp96.f
// (Herrmann/CPS srfdis as modified for DSurfTomo).  Every routine cites the reference
// lines it follows.  Typing follows the Fortran exactly: surfdisp96 itself has no IMPLICIT
// statement, so betmx, betmn, cc1, cc0, t1a, t1b, gvel are REAL*4 (surfdisp96.f:52-90);
// nevill/half/dltar*/var/normc/dnka are IMPLICIT DOUBLE PRECISION.
// Integer powers follow gfortran -O (powi expansion table): x**3 = x*(x*x),
// x**4 = (x*x)*(x*x), x**5 = (x*x)*(x*(x*x)).
#include <cmath>
#include <cstring>
#include "oracle.h"

namespace {

constexpr int NL = 200;

struct Layers {  // the arrays threaded through every call (surfdisp96.f:80)
  float d[NL + 1], a[NL + 1], b[NL + 1], rho[NL + 1], rtp[NL + 1], dtp[NL + 1], btp[NL + 1];
  int mmax, llw;
  float dhalf;    // SAVE dhalf, surfdisp96.f:509 (serial semantics)
  double del1st;  // SAVE del1st, surfdisp96.f:409 (serial semantics: per surfdisp96 call)
};

inline double dsign1(double x) { return std::copysign(1.0, x); }

// surfdisp96.f:361-382 -- half-space Rayleigh start value, all REAL*4
void gtsolh(float a, float b, float &c) {
  c = 0.95f * b;
  for (int i = 1; i <= 5; i++) {
    float gamma = b / a;
    float kappa = c / b;
    float k2 = kappa * kappa;
    float gk = gamma * kappa;
    float gk2 = gk * gk;
    float fac1 = std::sqrt(1.0f - gk2);
    float fac2 = std::sqrt(1.0f - k2);
    float tk = 2.0f - k2;
    float fr = tk * tk - 4.0f * fac1 * fac2;
    float frp = -(4.0f * (2.0f - k2) * kappa) + 4.0f * fac2 * gamma * gamma * kappa / fac1 +
                4.0f * fac1 * kappa / fac2;
    frp = frp / b;
    c = c - fr / frp;
  }
}

// surfdisp96.f:480-549
void sphere(int ifunc, int iflag, Layers &L) {
  const int mmax = L.mmax;
  double ar = 6370.0, dr = 0.0, r0 = ar;
  L.d[mmax] = 1.0f;
  if (iflag == 0) {
    for (int i = 1; i <= mmax; i++) {
      L.dtp[i] = L.d[i];
      L.rtp[i] = L.rho[i];
    }
    for (int i = 1; i <= mmax; i++) {
      dr = dr + (double)L.d[i];
      double r1 = ar - dr;
      double z0 = ar * std::log(ar / r0);
      double z1 = ar * std::log(ar / r1);
      L.d[i] = (float)(z1 - z0);
      double tmp = (ar + ar) / (r0 + r1);
      L.a[i] = (float)((double)L.a[i] * tmp);
      L.b[i] = (float)((double)L.b[i] * tmp);
      L.btp[i] = (float)tmp;
      r0 = r1;
    }
    L.dhalf = L.d[mmax];
  } else {
    L.d[mmax] = L.dhalf;
    for (int i = 1; i <= mmax; i++) {
      if (ifunc == 1) {
        float x = L.btp[i];
        float x2 = x * x;
        float x5 = x2 * (x * x2);
        L.rho[i] = L.rtp[i] * (1.0f / x5);  // btp**(-5), surfdisp96.f:539
      } else if (ifunc == 2) {
        L.rho[i] = L.rtp[i] * std::pow(L.btp[i], -2.275f);  // surfdisp96.f:541 (REAL*4 pow)
      }
    }
  }
  L.d[mmax] = 0.0f;
}

// surfdisp96.f:704-761 -- SH period equation (Haskell 2-vector from the half-space up)
double dltar1(double wvno, double omega, const Layers &L) {
  const int mmax = L.mmax, llw = L.llw;
  double beta1 = (double)L.b[mmax];
  double rho1 = (double)L.rho[mmax];
  double xkb = omega / beta1;
  double wvnop = wvno + xkb;
  double wvnom = std::fabs(wvno - xkb);
  double rb = std::sqrt(wvnop * wvnom);
  double e1 = rho1 * rb;
  double e2 = 1.0 / (beta1 * beta1);
  for (int m = mmax - 1; m >= llw; m--) {
    beta1 = (double)L.b[m];
    rho1 = (double)L.rho[m];
    double xmu = rho1 * beta1 * beta1;
    xkb = omega / beta1;
    wvnop = wvno + xkb;
    wvnom = std::fabs(wvno - xkb);
    rb = std::sqrt(wvnop * wvnom);
    double q = (double)L.d[m] * rb;
    double sinq, cosq, y, z;
    if (wvno < xkb) {
      sinq = std::sin(q);
      y = sinq / rb;
      z = -rb * sinq;
      cosq = std::cos(q);
    } else if (wvno == xkb) {
      cosq = 1.0;
      y = (double)L.d[m];
      z = 0.0;
    } else {
      double fac = 0.0;
      if (q < 16.0) fac = std::exp(-2.0 * q);
      cosq = (1.0 + fac) * 0.5;
      sinq = (1.0 - fac) * 0.5;
      y = sinq / rb;
      z = rb * sinq;
    }
    double e10 = e1 * cosq + e2 * xmu * z;
    double e20 = e1 * y / xmu + e2 * cosq;
    double xnor = std::fabs(e10);
    double ynor = std::fabs(e20);
    if (ynor > xnor) xnor = ynor;
    if (xnor < 1.0e-40) xnor = 1.0;
    e1 = e10 / xnor;
    e2 = e20 / xnor;
  }
  return e1;
}

struct VarOut {
  double w, cosp, exa, a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz;
};

// surfdisp96.f:868-987
void var(double p, double q, double ra, double rb, double wvno, double xka, double xkb,
         double dpth, VarOut &o) {
  double w = 0, x = 0, cosp = 0, sinp, y = 0, z = 0, cosq = 0, sinq, fac;
  double pex = 0.0, sex = 0.0;
  if (wvno < xka) {
    sinp = std::sin(p);
    w = sinp / ra;
    x = -ra * sinp;
    cosp = std::cos(p);
  } else if (wvno == xka) {
    cosp = 1.0;
    w = dpth;
    x = 0.0;
  } else if (wvno > xka) {
    pex = p;
    fac = 0.0;
    if (p < 16.0) fac = std::exp(-2.0 * p);
    cosp = (1.0 + fac) * 0.5;
    sinp = (1.0 - fac) * 0.5;
    w = sinp / ra;
    x = ra * sinp;
  }
  if (wvno < xkb) {
    sinq = std::sin(q);
    y = sinq / rb;
    z = -rb * sinq;
    cosq = std::cos(q);
  } else if (wvno == xkb) {
    cosq = 1.0;
    y = dpth;
    z = 0.0;
  } else if (wvno > xkb) {
    sex = q;
    fac = 0.0;
    if (q < 16.0) fac = std::exp(-2.0 * q);
    cosq = (1.0 + fac) * 0.5;
    sinq = (1.0 - fac) * 0.5;
    y = sinq / rb;
    z = rb * sinq;
  }
  double exa = pex + sex;
  double a0 = 0.0;
  if (exa < 60.0) a0 = std::exp(-exa);
  o.exa = exa;
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
  // surfdisp96.f:979-985 rescale cosq,y,z by exp(sex-pex): results unused by dltar4's
  // solid-layer path (the compound-matrix products above are already formed).
}

// surfdisp96.f:1018-1062 -- Dunkin compound matrix, ca(row,col) 1-based
void dnka(double ca[6][6], double wvno2, double gam, double gammk, double rho, const VarOut &v) {
  const double one = 1.0, two = 2.0;
  double gamm1 = gam - one;
  double twgm1 = gam + gamm1;
  double gmgmk = gam * gammk;
  double gmgm1 = gam * gamm1;
  double gm1sq = gamm1 * gamm1;
  double rho2 = rho * rho;
  double a0pq = v.a0 - v.cpcq;
  ca[1][1] = v.cpcq - two * gmgm1 * a0pq - gmgmk * v.xz - wvno2 * gm1sq * v.wy;
  ca[1][2] = (wvno2 * v.cpy - v.cqx) / rho;
  ca[1][3] = -(twgm1 * a0pq + gammk * v.xz + wvno2 * gamm1 * v.wy) / rho;
  ca[1][4] = (v.cpz - wvno2 * v.cqw) / rho;
  ca[1][5] = -(two * wvno2 * a0pq + v.xz + wvno2 * wvno2 * v.wy) / rho2;
  ca[2][1] = (gmgmk * v.cpz - gm1sq * v.cqw) * rho;
  ca[2][2] = v.cpcq;
  ca[2][3] = gammk * v.cpz - gamm1 * v.cqw;
  ca[2][4] = -v.wz;
  ca[2][5] = ca[1][4];
  ca[4][1] = (gm1sq * v.cpy - gmgmk * v.cqx) * rho;
  ca[4][2] = -v.xy;
  ca[4][3] = gamm1 * v.cpy - gammk * v.cqx;
  ca[4][4] = ca[2][2];
  ca[4][5] = ca[1][2];
  ca[5][1] = -(two * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * v.xz + gm1sq * gm1sq * v.wy) * rho2;
  ca[5][2] = ca[4][1];
  ca[5][3] = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * v.xz + gamm1 * gm1sq * v.wy) * rho;
  ca[5][4] = ca[2][1];
  ca[5][5] = ca[1][1];
  double t = -two * wvno2;
  ca[3][1] = t * ca[5][3];
  ca[3][2] = t * ca[4][3];
  ca[3][3] = v.a0 + two * (v.cpcq - ca[1][1]);
  ca[3][4] = t * ca[2][3];
  ca[3][5] = t * ca[1][3];
}

// surfdisp96.f:989-1014 (the dlog(t1) result "ex" is never used by the caller)
void normc(double ee[6]) {
  double t1 = 0.0;
  for (int i = 1; i <= 5; i++)
    if (std::fabs(ee[i]) > t1) t1 = std::fabs(ee[i]);
  if (t1 < 1.0e-40) t1 = 1.0;
  for (int i = 1; i <= 5; i++) {
    double t2 = ee[i];
    t2 = t2 / t1;
    ee[i] = t2;
  }
}

// surfdisp96.f:767-864 -- P-SV period equation
double dltar4(double wvno, double omga, const Layers &L) {
  const int mmax = L.mmax, llw = L.llw;
  double e[6], ee[6], ca[6][6];
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  double wvno2 = wvno * wvno;
  double xka = omega / (double)L.a[mmax];
  double xkb = omega / (double)L.b[mmax];
  double wvnop = wvno + xka;
  double wvnom = std::fabs(wvno - xka);
  double ra = std::sqrt(wvnop * wvnom);
  wvnop = wvno + xkb;
  wvnom = std::fabs(wvno - xkb);
  double rb = std::sqrt(wvnop * wvnom);
  double t = (double)L.b[mmax] / omega;
  double gammk = 2.0 * t * t;
  double gam = gammk * wvno2;
  double gamm1 = gam - 1.0;
  double rho1 = (double)L.rho[mmax];
  e[1] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
  e[2] = -rho1 * ra;
  e[3] = rho1 * (gamm1 - gammk * ra * rb);
  e[4] = rho1 * rb;
  e[5] = wvno2 - ra * rb;
  VarOut v;
  for (int m = mmax - 1; m >= llw; m--) {
    xka = omega / (double)L.a[m];
    xkb = omega / (double)L.b[m];
    t = (double)L.b[m] / omega;
    gammk = 2.0 * t * t;
    gam = gammk * wvno2;
    wvnop = wvno + xka;
    wvnom = std::fabs(wvno - xka);
    ra = std::sqrt(wvnop * wvnom);
    wvnop = wvno + xkb;
    wvnom = std::fabs(wvno - xkb);
    rb = std::sqrt(wvnop * wvnom);
    double dpth = (double)L.d[m];
    rho1 = (double)L.rho[m];
    double p = ra * dpth;
    double q = rb * dpth;
    var(p, q, ra, rb, wvno, xka, xkb, dpth, v);
    dnka(ca, wvno2, gam, gammk, rho1, v);
    for (int i = 1; i <= 5; i++) {
      double cr = 0.0;
      for (int j = 1; j <= 5; j++) cr = cr + e[j] * ca[j][i];
      ee[i] = cr;
    }
    normc(ee);
    for (int i = 1; i <= 5; i++) e[i] = ee[i];
  }
  if (llw != 1) {  // water layer, surfdisp96.f:844-860
    xka = omega / (double)L.a[1];
    wvnop = wvno + xka;
    wvnom = std::fabs(wvno - xka);
    ra = std::sqrt(wvnop * wvnom);
    double dpth = (double)L.d[1];
    rho1 = (double)L.rho[1];
    double p = ra * dpth;
    double znul = 1.0e-05;
    var(p, znul, ra, znul, wvno, xka, znul, dpth, v);
    double w0 = -rho1 * v.w;
    return v.cosp * e[1] + w0 * e[2];
  }
  return e[1];
}

// surfdisp96.f:684-700
inline double dltar(double wvno, double omega, int kk, const Layers &L) {
  if (kk == 1) return dltar1(wvno, omega, L);
  return dltar4(wvno, omega, L);
}

// surfdisp96.f:670-680
inline void half(double c1, double c2, double &c3, double &del3, double omega, int ifunc,
                 const Layers &L) {
  c3 = 0.5 * (c1 + c2);
  double wvno = omega / c3;
  del3 = dltar(wvno, omega, ifunc, L);
}

// surfdisp96.f:551-668 -- hybrid bisection / Neville refinement
void nevill(double t, double c1, double c2, double del1, double del2, int ifunc, double &cc,
            const Layers &L, double twopi) {
  double x[21], y[21];
  double c3, del3;
  int m = 1;
  double omega = twopi / t;
  half(c1, c2, c3, del3, omega, ifunc, L);
  int nev = 1;
  int nctrl = 1;
  for (;;) {
    nctrl = nctrl + 1;
    if (nctrl >= 100) break;
    if (c3 < std::fmin(c1, c2) || c3 > std::fmax(c1, c2)) {
      nev = 0;
      half(c1, c2, c3, del3, omega, ifunc, L);
    }
    double s13 = del1 - del3;
    double s32 = del3 - del2;
    if (dsign1(del3) * dsign1(del1) < 0.0) {
      c2 = c3;
      del2 = del3;
    } else {
      c1 = c3;
      del1 = del3;
    }
    if (std::fabs(c1 - c2) <= 1.0e-6 * c1) break;
    if (dsign1(s13) != dsign1(s32)) nev = 0;
    double ss1 = std::fabs(del1);
    double s1 = (double)0.01f * ss1;  // REAL*4 literal 0.01 promoted, surfdisp96.f:621
    double ss2 = std::fabs(del2);
    double s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || nev == 0) {
      half(c1, c2, c3, del3, omega, ifunc, L);
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
        int j = m - kk + 1;
        double denom = y[m + 1] - y[j];
        if (std::fabs(denom) < 1.0e-10 * std::fabs(y[m + 1])) {
          fallback = true;
          break;
        }
        x[j] = (-y[j] * x[j + 1] + y[m + 1] * x[j]) / denom;
      }
      if (!fallback) {
        c3 = x[1];
        double wvno = omega / c3;
        del3 = dltar(wvno, omega, ifunc, L);
        nev = 2;
        m = m + 1;
        if (m > 10) m = 10;
      } else {
        half(c1, c2, c3, del3, omega, ifunc, L);
        nev = 1;
        m = 1;
      }
    }
  }
  cc = c3;
}

// surfdisp96.f:384-476 -- bracket by dc steps, then refine
void getsol(double t1, double &c1, double clow, double dc, double cm, float betmx, int &iret,
            int ifunc, int ifirst, Layers &L) {
  const double twopi = 2.0 * 3.141592653589793;
  double omega = twopi / t1;
  double wvno = omega / c1;
  double del1 = dltar(wvno, omega, ifunc, L);
  if (ifirst == 1) L.del1st = del1;
  double plmn = dsign1(L.del1st) * dsign1(del1);
  int idir = +1;
  if (ifirst == 1)
    idir = +1;
  else if (plmn >= 0.0)
    idir = +1;
  else
    idir = -1;
  double c2, del2, cn;
  for (;;) {
    if (idir > 0)
      c2 = c1 + dc;
    else
      c2 = c1 - dc;
    if (c2 <= clow) {
      idir = +1;
      c1 = clow;
      continue;  // surfdisp96.f:446-450 (del1 is NOT recomputed at clow)
    }
    omega = twopi / t1;
    wvno = omega / c2;
    del2 = dltar(wvno, omega, ifunc, L);
    if (dsign1(del1) != dsign1(del2)) break;  // root bracketed
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
  nevill(t1, c1, c2, del1, del2, ifunc, cn, L, twopi);
  c1 = cn;
  if (c1 > (double)betmx) {
    iret = -1;
    return;
  }
  iret = 1;
}

}  // namespace

// surfdisp96.f:52-354
extern "C" int oracle_surfdisp96(const float *thkm, const float *vpm, const float *vsm,
                                 const float *rhom, int nlayer, int iflsph, int iwave, int mode,
                                 int igr, int kmax, const double *t, double *cg) {
  constexpr int NP = 80;
  Layers L;
  std::memset(&L, 0, sizeof(L));
  double c[NP + 1], cb[NP + 1];
  int nfail = 0;
  const int mmax = nlayer;
  L.mmax = mmax;
  const int nsph = iflsph;
  for (int i = 1; i <= mmax; i++) {
    L.b[i] = vsm[i - 1];
    L.a[i] = vpm[i - 1];
    L.d[i] = thkm[i - 1];
    L.rho[i] = rhom[i - 1];
  }
  int idispl = 0, idispr = 0;
  if (iwave == 1) {
    idispl = kmax;
    idispr = 0;
  } else if (iwave == 2) {
    idispl = 0;
    idispr = kmax;
  }
  const float sone0 = 1.500f, ddc0 = 0.005f, h0 = 0.005f;
  L.llw = 1;
  if (L.b[1] <= 0.0f) L.llw = 2;
  const double one = 1.0e-2;
  if (nsph == 1) sphere(0, 0, L);
  int jmn = 1, jsol = 1;
  float betmx = -1.e20f, betmn = 1.e20f;
  for (int i = 1; i <= mmax; i++) {
    if (L.b[i] > 0.01f && L.b[i] < betmn) {
      betmn = L.b[i];
      jmn = i;
      jsol = 1;
    } else if (L.b[i] <= 0.01f && L.a[i] < betmn) {
      betmn = L.a[i];
      jmn = i;
      jsol = 0;
    }
    if (L.b[i] > betmx) betmx = L.b[i];
  }
  for (int ifunc = 1; ifunc <= 2; ifunc++) {
    if (ifunc == 1 && idispl <= 0) continue;
    if (ifunc == 2 && idispr <= 0) continue;
    if (nsph == 1) sphere(ifunc, 1, L);
    float ddc = ddc0, sone = sone0, h = h0;
    if (sone < 0.01f) sone = 2.0f;
    double onea = (double)sone;
    float cc1;
    if (jsol == 0)
      cc1 = betmn;
    else
      gtsolh(L.a[jmn], L.b[jmn], cc1);
    cc1 = .95f * cc1;
    cc1 = .90f * cc1;
    double cc = (double)cc1;
    double dc = (double)ddc;
    dc = std::fabs(dc);
    double c1 = cc;
    double cm = cc;
    for (int i = 1; i <= kmax; i++) {
      cb[i] = 0.0;
      c[i] = 0.0;
    }
    int ift = 999;
    for (int iq = 1; iq <= mode; iq++) {
      const int is = 1, ie = kmax;
      int k;
      bool failed = false;
      for (k = is; k <= ie; k++) {
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
        if (k == is && iq == 1) {
          c1 = cc;
          clow = cc;
          ifirst = 1;
        } else if (k == is && iq > 1) {
          c1 = c[is] + one * dc;
          clow = c1;
          ifirst = 1;
        } else if (k > is && iq > 1) {
          ifirst = 0;
          clow = c[k] + one * dc;
          c1 = c[k - 1];
          if (c1 < clow) c1 = clow;
        } else {  // k > is && iq == 1
          ifirst = 0;
          c1 = c[k - 1] - onea * dc;
          clow = cm;
        }
        int iret;
        getsol(t1, c1, clow, dc, cm, betmx, iret, ifunc, ifirst, L);
        if (iret == -1) {
          failed = true;
          break;
        }
        c[k] = c1;
        if (igr > 0) {
          t1 = (double)t1b;
          ifirst = 0;
          clow = cb[k] + one * dc;
          c1 = c1 - onea * dc;
          getsol(t1, c1, clow, dc, cm, betmx, iret, ifunc, ifirst, L);
          if (iret == -1) c1 = c[k];
          cb[k] = c1;
        } else {
          c1 = 0.0;
        }
        float cc0 = (float)c[k];
        float cc1b = (float)c1;
        if (igr == 0) {
          cg[k - 1] = (double)cc0;
        } else {
          float gvel = (1.0f / t1a - 1.0f / t1b) / (1.0f / (t1a * cc0) - 1.0f / (t1b * cc1b));
          cg[k - 1] = (double)gvel;
        }
      }
      if (failed) {
        // labels 1700/1750, surfdisp96.f:307-348: warning text to unit 66 omitted; the
        // remaining periods are zero-filled and later modes/periods skip via ift.
        ift = k;
        for (int i = k; i <= ie; i++) {
          cg[i - 1] = 0.0;
          nfail++;
        }
      }
    }
  }
  return nfail;
}
