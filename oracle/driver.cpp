// ORACLE (test infrastructure, see oracle.h) -- CPU restatement of the host glue that sits
// between CalSurfG and LSMR in the reference's outer loop (src/main.f90:361-466, 518-532),
// plus delsph (src/delsph.f90) and getpercentile (src/getpercentile.f90).  Needed as the
// test driver because no Fortran compiler exists in this environment.
#include <cmath>
#include <vector>
#include "oracle.h"

// delsph.f90:1-28 (REAL*4; inputs are colatitude / longitude in radians)
extern "C" float oracle_delsph(float flat1, float flon1, float flat2, float flon2) {
  const float R = 6371.0f;
  const float pi = 3.1415926535898f;
  float dlat = flat2 - flat1;
  float dlon = flon2 - flon1;
  float lat1 = pi / 2 - flat1;
  float lat2 = pi / 2 - flat2;
  float a = std::sin(dlat / 2) * std::sin(dlat / 2) +
            std::sin(dlon / 2) * std::sin(dlon / 2) * std::cos(lat1) * std::cos(lat2);
  float c = 2 * std::atan2(std::sqrt(a), std::sqrt(1 - a));
  return R * c;
}

// getpercentile.f90:1-50 -- heapsort of a copy, then RA(int(0.25N)), RA(int(0.75N))
extern "C" void oracle_getpercentile(int N, const float *array, float *q25, float *q75) {
  std::vector<float> RA(N + 2);
  for (int i = 1; i <= N; i++) RA[i] = array[i - 1];
  int L = N / 2 + 1;
  int IR = N;
  float RRA;
  for (;;) {
    if (L > 1) {
      L = L - 1;
      RRA = RA[L];
    } else {
      RRA = RA[IR];
      RA[IR] = RA[1];
      IR = IR - 1;
      if (IR == 1) {
        RA[1] = RRA;
        int idx = (int)(0.25f * (float)N);
        *q25 = RA[idx];
        idx = (int)(0.75f * (float)N);
        *q75 = RA[idx];
        return;
      }
    }
    int I = L;
    int J = L + L;
    while (J <= IR) {
      if (J < IR) {
        if (RA[J] < RA[J + 1]) J = J + 1;
      }
      if (RRA < RA[J]) {
        RA[I] = RA[J];
        I = J;
        J = J + J;
      } else {
        J = IR + 1;
      }
    }
    RA[I] = RRA;
  }
}

// main.f90:361-466.  iw/rw/col as left by CalSurfG (iw[1+k] = row of triplet k); on return
// iw = [nar | rows | cols] and cbst has m entries.  Returns m.
extern "C" int oracle_host_glue(int nx, int ny, int nz, int dall, const float *obst,
                                const float *dsyn, float threshold0, float weight, int *iw,
                                float *rw, int *col, float *cbst, float *datweight, int *nar_io) {
  int nar = *nar_io;
  for (int i = 0; i < dall; i++) cbst[i] = obst[i] - dsyn[i];
  float q25, q75;
  oracle_getpercentile(dall, cbst, &q25, &q75);
  for (int i = 0; i < dall; i++) datweight[i] = 1.0f;
  for (int i = 0; i < dall; i++) {
    if (cbst[i] < q25 * threshold0 || cbst[i] > q75 * threshold0) {
      datweight[i] = 0.0f;
      cbst[i] = 0;
    }
  }
  for (int i = 0; i < nar; i++) rw[i] = rw[i] * datweight[iw[1 + i] - 1];
  // smoothing rows, main.f90:418-459
  int count3 = 0;
  const int nvz = ny - 2, nvx = nx - 2;
  for (int k = 1; k <= nz - 1; k++)
    for (int j = 1; j <= nvz; j++)
      for (int i = 1; i <= nvx; i++) {
        count3 = count3 + 1;
        const int c0 = (k - 1) * nvz * nvx + (j - 1) * nvx + i;
        if (i == 1 || i == nvx || j == 1 || j == nvz || k == 1 || k == nz - 1) {
          col[nar] = c0;
          rw[nar] = 2.0f * weight;
          iw[1 + nar] = dall + count3;
          cbst[dall + count3 - 1] = 0;
          nar = nar + 1;
        } else {
          const int cc[7] = {c0, c0 - 1, c0 + 1, c0 - nvx, c0 + nvx, c0 - nvz * nvx, c0 + nvz * nvx};
          for (int q = 0; q < 7; q++) {
            col[nar + q] = cc[q];
            rw[nar + q] = (q == 0 ? 6.0f : -1.0f) * weight;
            iw[1 + nar + q] = dall + count3;
          }
          cbst[dall + count3 - 1] = 0;
          nar = nar + 7;
        }
      }
  iw[0] = nar;
  for (int i = 0; i < nar; i++) iw[1 + nar + i] = col[i];
  *nar_io = nar;
  return dall + count3;
}

// main.f90:518-532
extern "C" void oracle_model_update(int nx, int ny, int nz, float *vsf, float *dv, float minvel,
                                    float maxvel) {
  for (int k = 1; k <= nz - 1; k++)
    for (int j = 1; j <= ny - 2; j++)
      for (int i = 1; i <= nx - 2; i++) {
        float &d = dv[(size_t)(k - 1) * (nx - 2) * (ny - 2) + (size_t)(j - 1) * (nx - 2) + (i - 1)];
        if (d >= 0.500f) d = 0.500f;
        if (d <= -0.500f) d = -0.500f;
        float &v = vsf[(size_t)(k - 1) * nx * ny + (size_t)j * nx + i];
        v = v + d;
        if (v < minvel) v = minvel;
        if (v > maxvel) v = maxvel;
      }
}
