// ORACLE (test infrastructure, see oracle.h) -- CPU restatement of the reference's sparse
// solver: aprod (src/aprod.f90:7-60), dnrm2/dscal retyped to REAL*4
// (src/lsmrblas.f90:247-359) and LSMR with local reorthogonalisation
// (src/lsmrModule.f90:36-750).  real(dp) is REAL*4 here (lsmrDataModule.f90:21), so every
// scalar and vector is float and every operation is rounded to float in source order.
#include <algorithm>
#include <cmath>
#include <vector>
#include "oracle.h"

// aprod.f90:7-60.  iw = [nar | row(1..nar) | col(1..nar)] (1-based indices)
extern "C" void oracle_aprod(int mode, int m, int n, float *x, float *y, int leniw, int lenrw,
                             const int *iw, const float *rw) {
  (void)m; (void)n; (void)leniw; (void)lenrw;
  const int kk = iw[0];
  const int *irow = iw + 1;
  const int *icol = iw + 1 + kk;
  if (mode == 1) {
    for (int k = 0; k < kk; k++) y[irow[k] - 1] = y[irow[k] - 1] + rw[k] * x[icol[k] - 1];
  } else {
    for (int k = 0; k < kk; k++) x[icol[k] - 1] = x[icol[k] - 1] + rw[k] * y[irow[k] - 1];
  }
}

// lsmrblas.f90:247-277 (scaled sum of squares, REAL*4)
extern "C" float oracle_snrm2(int n, const float *x) {
  float norm;
  if (n < 1) {
    norm = 0.0f;
  } else if (n == 1) {
    norm = std::fabs(x[0]);
  } else {
    float scale = 0.0f, ssq = 1.0f;
    for (int ix = 0; ix < n; ix++) {
      if (x[ix] != 0.0f) {
        float absxi = std::fabs(x[ix]);
        if (scale < absxi) {
          float r = scale / absxi;
          ssq = 1.0f + ssq * (r * r);
          scale = absxi;
        } else {
          float r = absxi / scale;
          ssq = ssq + r * r;
        }
      }
    }
    norm = scale * std::sqrt(ssq);
  }
  return norm;
}

namespace {
inline void sscal(int n, float sa, float *x) {  // lsmrblas.f90:317-359 (incx == 1)
  for (int i = 0; i < n; i++) x[i] = sa * x[i];
}
inline float d2norm(float a, float b) {  // lsmrModule.f90:686-709
  float scale = std::fabs(a) + std::fabs(b);
  if (scale == 0.0f) return 0.0f;
  float ra = a / scale, rb = b / scale;
  return scale * std::sqrt(ra * ra + rb * rb);
}
}  // namespace

// lsmrModule.f90:36-750 (nout is undefined in the caller, main.f90:107 -> treated as silent)
extern "C" void oracle_lsmr(int m, int n, int leniw, int lenrw, const int *iw, const float *rw,
                            const float *b, float damp, float atol, float btol, float conlim,
                            int itnlim, int localSize, float *x, int *istop_out, int *itn_out,
                            float *normA_out, float *condA_out, float *normr_out,
                            float *normAr_out, float *normx_out) {
  const float zero = 0.0f, one = 1.0f;
  std::vector<float> h(n), hbar(n), u(m), v(n), w(n);
  const int localVecs = std::min(localSize, std::min(m, n));
  std::vector<float> localV((size_t)n * std::max(localVecs, 1));
  bool localOrtho = false, localVQueueFull = false;
  int localPointer = 0;
  int istop = 0, itn = 0;
  float normA = 0, condA = 0, normr = 0, normAr = 0, normx = 0;
  const bool damped = damp > zero;

  for (int i = 0; i < m; i++) u[i] = b[i];
  for (int i = 0; i < n; i++) { v[i] = zero; x[i] = zero; }
  float alpha = zero;
  float beta = oracle_snrm2(m, u.data());
  if (beta > zero) {
    sscal(m, one / beta, u.data());
    oracle_aprod(2, m, n, v.data(), u.data(), leniw, lenrw, iw, rw);
    alpha = oracle_snrm2(n, v.data());
  }
  if (alpha > zero) {
    sscal(n, one / alpha, v.data());
    w = v;
  }
  normAr = alpha * beta;
  if (normAr != zero) {
    if (localVecs > 0) {
      localPointer = 1;
      localOrtho = true;
      localVQueueFull = false;
      for (int i = 0; i < n; i++) localV[i] = v[i];
    }
    itn = 0;
    float zetabar = alpha * beta, alphabar = alpha, rho = 1, rhobar = 1, cbar = 1, sbar = 0;
    h = v;
    for (int i = 0; i < n; i++) { hbar[i] = zero; x[i] = zero; }
    float betadd = beta, betad = 0, rhodold = 1, tautildeold = 0, thetatilde = 0, zeta = 0, d = 0;
    float normA2 = alpha * alpha, maxrbar = 0.0f, minrbar = 1e+30f;
    const float normb = beta;
    istop = 0;
    float ctol = zero;
    if (conlim > zero) ctol = one / conlim;
    normr = beta;
    normAr = alpha * beta;
    // (normAr == 0 exit at :447-452 cannot trigger here: tested above)
    for (;;) {
      itn = itn + 1;
      sscal(m, -alpha, u.data());
      oracle_aprod(1, m, n, v.data(), u.data(), leniw, lenrw, iw, rw);
      beta = oracle_snrm2(m, u.data());
      if (beta > zero) {
        sscal(m, one / beta, u.data());
        if (localOrtho) {  // localVEnqueue, :715-726
          if (localPointer < localVecs) {
            localPointer = localPointer + 1;
          } else {
            localPointer = 1;
            localVQueueFull = true;
          }
          float *q = &localV[(size_t)(localPointer - 1) * n];
          for (int i = 0; i < n; i++) q[i] = v[i];
        }
        sscal(n, -beta, v.data());
        oracle_aprod(2, m, n, v.data(), u.data(), leniw, lenrw, iw, rw);
        if (localOrtho) {  // localVOrtho, :731-748
          const int lim = localVQueueFull ? localVecs : localPointer;
          for (int c = 1; c <= lim; c++) {
            const float *q = &localV[(size_t)(c - 1) * n];
            float dd = 0.0f;  // dot_product: sequential REAL*4 accumulation
            for (int i = 0; i < n; i++) dd = dd + v[i] * q[i];
            for (int i = 0; i < n; i++) v[i] = v[i] - dd * q[i];
          }
        }
        alpha = oracle_snrm2(n, v.data());
        if (alpha > zero) sscal(n, one / alpha, v.data());
      }
      float alphahat = d2norm(alphabar, damp);
      float chat = alphabar / alphahat;
      float shat = damp / alphahat;
      float rhoold = rho;
      rho = d2norm(alphahat, beta);
      float c = alphahat / rho;
      float s = beta / rho;
      float thetanew = s * alpha;
      alphabar = c * alpha;
      float rhobarold = rhobar;
      float zetaold = zeta;
      float thetabar = sbar * rho;
      float rhotemp = cbar * rho;
      rhobar = d2norm(cbar * rho, thetanew);
      cbar = cbar * rho / rhobar;
      sbar = thetanew / rhobar;
      zeta = cbar * zetabar;
      zetabar = -sbar * zetabar;
      {
        const float f1 = thetabar * rho / (rhoold * rhobarold);
        for (int i = 0; i < n; i++) hbar[i] = h[i] - f1 * hbar[i];
        const float f2 = zeta / (rho * rhobar);
        for (int i = 0; i < n; i++) x[i] = x[i] + f2 * hbar[i];
        const float f3 = thetanew / rho;
        for (int i = 0; i < n; i++) h[i] = v[i] - f3 * h[i];
      }
      float betaacute = chat * betadd;
      float betacheck = -shat * betadd;
      float betahat = c * betaacute;
      betadd = -s * betaacute;
      float thetatildeold = thetatilde;
      float rhotildeold = d2norm(rhodold, thetabar);
      float ctildeold = rhodold / rhotildeold;
      float stildeold = thetabar / rhotildeold;
      thetatilde = stildeold * rhobar;
      rhodold = ctildeold * rhobar;
      betad = -stildeold * betad + ctildeold * betahat;
      tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold;
      float taud = (zeta - thetatilde * tautildeold) / rhodold;
      d = d + betacheck * betacheck;
      {
        float e = betad - taud;
        normr = std::sqrt(d + e * e + betadd * betadd);
      }
      normA2 = normA2 + beta * beta;
      normA = std::sqrt(normA2);
      normA2 = normA2 + alpha * alpha;
      maxrbar = std::max(maxrbar, rhobarold);
      if (itn > 1) minrbar = std::min(minrbar, rhobarold);
      condA = std::max(maxrbar, rhotemp) / std::min(minrbar, rhotemp);
      normAr = std::fabs(zetabar);
      normx = oracle_snrm2(n, x);
      float test1 = normr / normb;
      float test2 = normAr / (normA * normr);
      float test3 = one / condA;
      float t1 = test1 / (one + normA * normx / normb);
      float rtol = btol + atol * normA * normx / normb;
      if (itn >= itnlim) istop = 7;
      if (one + test3 <= one) istop = 6;
      if (one + test2 <= one) istop = 5;
      if (one + t1 <= one) istop = 4;
      if (test3 <= ctol) istop = 3;
      if (test2 <= atol) istop = 2;
      if (test1 <= rtol) istop = 1;
      if (istop != 0) break;
    }
  }
  if (damped && istop == 2) istop = 3;
  *istop_out = istop;
  *itn_out = itn;
  *normA_out = normA;
  *condA_out = condA;
  *normr_out = normr;
  *normAr_out = normAr;
  *normx_out = normx;
}
