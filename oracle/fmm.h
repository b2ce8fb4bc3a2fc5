// ORACLE (test infrastructure, see oracle.h) -- state of the reference's globalp/traveltime
// modules (src/CalSurfG.f90:181-247, 264-287) as one object so independent gathers can be
// timed on several host threads.  All REAL(KIND=i10) are float.
#ifndef DSURF_ORACLE_FMM_H
#define DSURF_ORACLE_FMM_H
#include <cstdint>
#include <vector>

namespace oracle {

struct Fmm {
  // ---- globalp (CalSurfG.f90:181-247)
  int nvx = 0, nvz = 0, nnx = 0, nnz = 0, fom = 1, gdx = 8, gdz = 8;
  int vnl = 0, vnr = 0, vnt = 0, vnb = 0, nrnx = 0, nrnz = 0, sgdl = 8, rbint = 0;
  int nnxr = 0, nnzr = 0, asgr = 1, sgs = 8;
  float gox = 0, goz = 0, dnx = 0, dnz = 0, dvx = 0, dvz = 0, snb = 0.5f, earth = 6371.0f;
  float goxd = 0, gozd = 0, dvxd = 0, dvzd = 0, dnxd = 0, dnzd = 0;
  float drnx = 0, drnz = 0, gorx = 0, gorz = 0;
  float dnxr = 0, dnzr = 0, goxr = 0, gozr = 0;
  const float pi = 3.1415926535898f;  // CalSurfG.f90:196, rounds to 3.14159274f
  // arrays: (iz,ix) column-major with leading dimension ld (>= max(coarse,refined) nnz)
  int ld = 0, ncols = 0;
  std::vector<float> veln, velnb, ttn, ttnr;
  std::vector<int> nsts, nstsr;
  std::vector<float> velv;  // velv(0:nvz+1,0:nvx+1), column-major, i (z) fastest
  // ---- traveltime module heap (CalSurfG.f90:266-287)
  int ntr = 0;
  std::vector<int> btg_px, btg_pz;
  int error = 0;  // 1: source outside, 2: receiver outside (reference STOPs)

  inline float &V(int iz, int ix) { return veln[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float &VB(int iz, int ix) { return velnb[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float &T(int iz, int ix) { return ttn[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float &TR(int iz, int ix) { return ttnr[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline int &S(int iz, int ix) { return nsts[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline int &SR(int iz, int ix) { return nstsr[(size_t)(ix - 1) * ld + (iz - 1)]; }
  inline float &VV(int i, int j) { return velv[(size_t)j * (nvz + 2) + i]; }

  // :1032-1094; gd = gdx = gdz: 8 in CalSurfG (:1032-1033), 5 in subroutine synthetic (:2497-2498)
  void setup(int nx, int ny, float goxdf, float gozdf, float dvxdf, float dvzdf, int gd = 8);
  void gridder(const double *pv);                                                   // :1460-1553
  void bsplrefine();                                                                // :1562-1628
  void travel(float scx, float scz, int urg);                                       // :288-487
  void fouds2(int iz, int ix);                                                      // :587-759
  void addtree(int iz, int ix);                                                     // :768-805
  void downtree();                                                                  // :816-885
  void updtree(int iz, int ix);                                                     // :894-921
  float bilinear(const float nv[3][3], float dsx, float dsz);                       // :2328-2349
  // :1186-1355: gridder + refined pass + injection + coarse pass for one source
  void solve_source(const double *pv, float x, float z);
  float srtimes(float scx, float scz, float rcx1, float rcz1);                      // :1636-1759
  void rpaths(float scx, float scz, float *fdm, float surfrcx, float surfrcz);      // :1771-2318
  // ray geometry rgx(1:nrp), rgz(1:nrp) of the last rpaths call when non-null (what the
  // commented raypath.out block :2276-2283 would print)
  std::vector<float> *path_x = nullptr, *path_z = nullptr;
};

}  // namespace oracle
#endif
