// Analytic pin of the oracle's fouds2 (oracle/fmm.cpp, restating src/CalSurfG.f90:587-759): a travel-time field
// linear in the grid coordinates, T = T0 + a (ix - ix0) + b (iz - iz0), satisfies the eikonal equation on the
// spherical-shell grid at node (ix0, iz0) with slowness s = sqrt((a/(r dnx))^2 + (b/(r sin(x) dnz))^2); one-sided
// first- and second-order differences are exact for it, so every two-dimensional branch must return T0.
// Driven by tests/test_oracle_pins.py.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../oracle/fmm.h"

int main() {
  oracle::Fmm f;
  f.setup(18, 18, 26.5f, 120.0f, 0.015f, 0.017f);
  std::vector<double> pv((size_t)18 * 18, 1.0);
  f.gridder(pv.data());
  const int ix0 = 40, iz0 = 55;
  const float T0 = 100.0f;
  int bad = 0;
  // order of the x leg / z leg: 2 = second-order (two alive nodes, T(j) > T(j2)), 1 = first-order (one alive node)
  for (int oj = 1; oj <= 2; oj++)
    for (int ok = 1; ok <= 2; ok++)
      for (int sj = -1; sj <= 1; sj += 2)
        for (int sk = -1; sk <= 1; sk += 2) {
          const float a = 0.11f, b = 0.07f;  // time increments per node towards (ix0, iz0) along x and z
          std::fill(f.nsts.begin(), f.nsts.end(), -1);
          std::fill(f.ttn.begin(), f.ttn.end(), 0.0f);
          // upwind neighbours on the sides (sj, sk): times decrease away from the node
          f.S(iz0, ix0 + sj) = 0;
          f.T(iz0, ix0 + sj) = T0 - a;
          if (oj == 2) {
            f.S(iz0, ix0 + 2 * sj) = 0;
            f.T(iz0, ix0 + 2 * sj) = T0 - 2.0f * a;
          }
          f.S(iz0 + sk, ix0) = 0;
          f.T(iz0 + sk, ix0) = T0 - b;
          if (ok == 2) {
            f.S(iz0 + 2 * sk, ix0) = 0;
            f.T(iz0 + 2 * sk, ix0) = T0 - 2.0f * b;
          }
          const double r = f.earth, rs = (double)f.earth * std::sin((double)f.gox + (double)(ix0 - 1) * f.dnx);
          const double s = std::sqrt(std::pow(a / (r * f.dnx), 2) + std::pow(b / (rs * f.dnz), 2));
          f.V(iz0, ix0) = (float)(1.0 / s);
          f.fouds2(iz0, ix0);
          const double rel = std::fabs((double)f.T(iz0, ix0) - T0) / T0;
          printf("case oj=%d ok=%d sj=%+d sk=%+d T=%.7f rel %.3e\n", oj, ok, sj, sk, f.T(iz0, ix0), rel);
          if (!(rel <= 4e-6)) bad++;
        }
  return bad != 0;
}
