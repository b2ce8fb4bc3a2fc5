// Device-resident scalar state of one LSMR solve (lsmrModule.f90:336-351: all REAL*4).
#pragma once
namespace dsurf {
struct LsmrScalars {
  float alpha, beta, inv_alpha, inv_beta, neg_beta;
  float alphabar, rho, rhobar, cbar, sbar, zeta, zetabar;
  float betadd, betad, rhodold, tautildeold, thetatilde, d;
  float normA2, maxrbar, minrbar, normb;
  float normA, condA, normr, normAr, normx;
  float f1, f2, f3, dot_d;
  double sum_u2;
  int beta_pos, alpha_pos;
  int localPointer, queueFull, enq_slot, orthoLimit;
  int itn, istop, stop;
};
}  // namespace dsurf
