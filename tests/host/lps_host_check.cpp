// Host replay of the lane-per-sweep march (dsurftomo_b200/csrc/eik_lps.cuh) against the oracle's
// Fmm::travel (oracle/fmm.cpp, restating src/CalSurfG.f90:288-487): same seeds, travel times must be
// bit-identical on every node.  Also reports how far the lazy back-pointer chain is walked.
// Build/run: tests/test_lps_host.py (g++ -O2 -ffp-contract=off, links oracle/liboracle.so).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../dsurftomo_b200/csrc/eik_lps.cuh"
#include "../../oracle/fmm.h"
#ifndef LPS_HOST_KLG
#define LPS_HOST_KLG 7
#endif

using namespace dsurf::lps;

struct HostMem {
  GridP G;
  std::vector<uint32_t> w;  // padded (nnx + 6) x (nnz + 6), frame = far
  std::vector<float> v;     // same padded indexing
  std::vector<float> ris;
  std::vector<Ent> h;
  long long probes = 0, lookups = 0, wstores = 0, pops = 0, ntr_sum = 0, slow = 0;
  int ntr_max = 0;
  int pub_root = -1, pub_state = 0;
  uint32_t pub_key = 0;
  void stat(int k, int v_) {
    if (k == 0) probes += v_;
    if (k == 1) lookups += v_;
    if (k == 3) slow += v_;
    if (k == 2) {
      pops++;
      ntr_sum += v_;
      if (v_ > ntr_max) ntr_max = v_;
    }
  }
  uint32_t word(int i) const { return w[i]; }
  void set_word(int i, uint32_t x) {
    w[i] = x;
    wstores++;
  }
  float vel(int i) const { return v[i]; }
  float risti(int ix) const { return ris[ix]; }
  Ent hget(int p) { return h[p]; }
  void hget2(int p, Ent &a, Ent &b) {
    a = h[p];
    b = h[p + 1];
  }
  void hset(int p, Ent e) { h[p] = e; }
  bool any(bool p) const { return p; }
  void sync() const {}
  void publish(int root, uint32_t key, int state) {
    pub_root = root;
    pub_key = key;
    pub_state = state;
  }
  void collect(float tv[4]) {  // on the device other warps do this while the heap is sifted
    for (int g = 0; g < 4; g++) {
      tv[g] = 0.0f;
      if (pub_state == 1) tv[g] = eval_neighbour(G, *this, pub_root, pub_key, g).tv;
    }
  }
  int div_ld(int i) const { return i / (G.nnz + 2 * kPad); }
  static constexpr int kLg = LPS_HOST_KLG;  // cheap levels of the device layout (7 there); small values exercise the block walk
  void hblock(int q, Ent b[14], int lim) {
    int k = 0;
    for (int d = 1; d <= 3; d++)
      for (int o = 0; o < (1 << d); o++) {
        const size_t sl = ((size_t)q << d) + o;
        b[k++] = (sl <= (size_t)lim) ? h[sl] : Ent{0, -1};
      }
  }
};

static int run_case(int nx, int ny, unsigned seed, int kind, float fx, float fz) {
  oracle::Fmm f;
  f.setup(nx, ny, 26.5f, 120.0f, 0.015f, 0.017f);
  std::mt19937 rng(seed);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> pv((size_t)nx * ny);
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      double v = 1.3;
      if (kind == 1) v = 1.2 + 0.35 * sin(0.7 * i) * cos(0.5 * j) + 0.05 * nd(rng);
      if (kind == 2) v = 0.6 + 2.0 * ((i / 3 + j / 3) & 1) + 0.2 * nd(rng);  // blocky, strong contrast
      if (v < 0.3) v = 0.3;
      pv[(size_t)j * nx + i] = v;
    }
  f.gridder(pv.data());
  const int nnx = f.nnx, nnz = f.nnz;
  const float scx = f.gox + fx * (nnx - 1) * f.dnx, scz = f.goz + fz * (nnz - 1) * f.dnz;
  // ---- seeds exactly as travel (:312-375)
  int isx = (int)((scx - f.gox) / f.dnx) + 1, isz = (int)((scz - f.goz) / f.dnz) + 1;
  if (isx == nnx) isx--;
  if (isz == nnz) isz--;
  HostMem m;
  const int ld = nnz + 2 * kPad;
  m.G = GridP{nnx, nnz, f.dnx, f.dnz, f.earth};
  m.w.assign((size_t)(nnx + 2 * kPad) * ld, kFar);
  m.v.assign(m.w.size(), 1.0f);
  for (int ix = 1; ix <= nnx; ix++)
    for (int iz = 1; iz <= nnz; iz++) m.v[(size_t)(ix - 1 + kPad) * ld + (iz - 1 + kPad)] = f.V(iz, ix);
  m.ris.resize(nnx);
  for (int ix = 1; ix <= nnx; ix++) m.ris[ix - 1] = f.earth * std::sin(f.gox + (float)(ix - 1) * f.dnx);
  const int hcap = nnx * nnz / 2 + 8;
  m.h.assign(hcap + 2, Ent{0, -1});
  float vss[3][3];
  for (int i = 1; i <= 2; i++)
    for (int j = 1; j <= 2; j++) vss[i][j] = f.V(isz - 1 + j, isx - 1 + i);
  const float dsx = (scx - f.gox) - (float)(isx - 1) * f.dnx, dsz = (scz - f.goz) - (float)(isz - 1) * f.dnz;
  const float vsrc = f.bilinear(vss, dsx, dsz);
  int ntr = 0;
  for (int i = 1; i <= 2; i++)
    for (int j = 1; j <= 2; j++) {
      const float ex = dsx - (float)(i - 1) * f.dnx, ez = dsz - (float)(j - 1) * f.dnz;
      const float ds = std::sqrt(ex * ex + ez * ez);
      const float t0 = 2.0f * ds / (vss[i][j] + vsrc);
      const int nid = (isx - 1 + i - 1 + kPad) * ld + (isz - 1 + j - 1 + kPad);
      bool moved;
      ntr++;
      const int pos = sift_up(m, ntr, t0, nid, moved);
      m.set_word(nid, kCloseBit | (uint32_t)pos);
    }
  const int rc = march(m.G, m, ntr, hcap);
  f.travel(scx, scz, 0);
  if (rc != 0 || f.error) {
    printf("FAIL rc=%d err=%d\n", rc, f.error);
    return 1;
  }
  long long bad = 0;
  for (int ix = 1; ix <= nnx; ix++)
    for (int iz = 1; iz <= nnz; iz++) {
      const float t = f.T(iz, ix);
      uint32_t b;
      memcpy(&b, &t, 4);
      if (b != m.w[(size_t)(ix - 1 + kPad) * ld + (iz - 1 + kPad)]) bad++;
    }
  printf("case nx=%d ny=%d kind=%d src=(%.2f,%.2f): %dx%d nodes, mismatches=%lld, word stores/pop=%.2f, "
         "lookups/pop=%.2f slow-path updates/pop=%.3f heap mean=%.0f max=%d\n", nx, ny, kind, fx, fz, nnx, nnz, bad,
         (double)m.wstores / (double)m.pops, (double)m.lookups / (double)m.pops,
         (double)m.slow / (double)m.pops, (double)m.ntr_sum / (double)m.pops, m.ntr_max);
  return bad != 0;
}

int main(int argc, char **argv) {
  const int big = argc > 1 ? atoi(argv[1]) : 0;
  int fails = 0;
  const float src[4][2] = {{0.5f, 0.5f}, {0.02f, 0.97f}, {0.31f, 0.66f}, {1.0f, 1.0f}};
  for (int kind = 0; kind < 3; kind++)
    for (int s = 0; s < 4; s++) fails += run_case(18, 20, 100 + 7 * kind + s, kind, src[s][0], src[s][1]);
  for (int kind = 0; kind < 3; kind++) fails += run_case(35, 35, 7 + kind, kind, 0.4f, 0.55f);
  if (big)
    for (int kind = 0; kind < 3; kind++) fails += run_case(131, 131, 99 + kind, kind, 0.47f, 0.52f);
  printf(fails ? "LPS HOST CHECK FAILED (%d)\n" : "LPS HOST CHECK OK\n", fails);
  return fails != 0;
}
