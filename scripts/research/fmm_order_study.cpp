// Research harness (not product, not oracle): how often does the reference's heap-layout-dependent
// pop order differ from a plain "sorted by key" priority queue?  Links oracle/fmm.cpp.
//   g++ -O2 -std=c++17 -ffp-contract=off -I oracle scripts/research/fmm_order_study.cpp oracle/fmm.cpp -o /tmp/fmm_study
#include "fmm.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <set>
#include <algorithm>
#include <vector>
using namespace oracle;

struct Ev { long pop; float lo, hi; int node; };
struct Stats { long startup_pops = 0, pops = 0, incr = 0, noncausal = 0, ties = 0, nearties = 0, updates = 0; std::vector<Ev> ev; double maxdef = 0; };

static int g_startup = 0;  // minimum number of exact (heap-replay) pops before switching to the sorted queue
static void travel_sorted(Fmm &f, float scx, float scz, int urg, Stats &st) {
  int isx = (int)((scx - f.gox) / f.dnx) + 1, isz = (int)((scz - f.goz) / f.dnz) + 1;
  if (isx == f.nnx) isx--;
  if (isz == f.nnz) isz--;
  std::set<std::pair<float, int>> pq;
  auto id = [&](int iz, int ix) { return (ix - 1) * f.ld + (iz - 1); };
  if (urg != 2) std::fill(f.nsts.begin(), f.nsts.end(), -1);
  if (urg == 2) {
    for (int i = 1; i <= f.nnx; i++)
      for (int j = 1; j <= f.nnz; j++)
        if (f.S(j, i) > 0) pq.insert({f.T(j, i), id(j, i)});
  } else {
    float vss[3][3];
    for (int i = 1; i <= 2; i++)
      for (int j = 1; j <= 2; j++) vss[i][j] = f.V(isz - 1 + j, isx - 1 + i);
    float dsx = (scx - f.gox) - (float)(isx - 1) * f.dnx, dsz = (scz - f.goz) - (float)(isz - 1) * f.dnz;
    float vsrc = f.bilinear(vss, dsx, dsz);
    for (int i = 1; i <= 2; i++)
      for (int j = 1; j <= 2; j++) {
        float ex = dsx - (float)(i - 1) * f.dnx, ez = dsz - (float)(j - 1) * f.dnz;
        float ds = std::sqrt(ex * ex + ez * ez);
        f.T(isz - 1 + j, isx - 1 + i) = 2.0f * ds / (vss[i][j] + vsrc);
        f.S(isz - 1 + j, isx - 1 + i) = 1;
        pq.insert({f.T(isz - 1 + j, isx - 1 + i), id(isz - 1 + j, isx - 1 + i)});
      }
  }
  if (g_startup > 0) {
    // rebuild the reference heap from pq's content in the reference's insertion order, then replay exactly
    std::vector<std::pair<float,int>> init(pq.begin(), pq.end());
    pq.clear();
    f.ntr = 0;
    if (urg == 2) {
      for (int i = 1; i <= f.nnx; i++) for (int j = 1; j <= f.nnz; j++) if (f.S(j, i) > 0) f.addtree(j, i);
    } else {
      for (int i = 1; i <= 2; i++) for (int j = 1; j <= 2; j++) f.addtree(isz - 1 + j, isx - 1 + i);
    }
    std::set<int> viol;
    long np = 0;
    bool exited = false;
    while (f.ntr > 0) {
      if (np >= g_startup && viol.empty()) break;
      int ix = f.btg_px[1], iz = f.btg_pz[1];
      if (urg == 1) {
        int swrg = 0;
        if (ix == 1 && f.vnl != 1) swrg = 1;
        if (ix == f.nnx && f.vnr != f.nnx) swrg = 1;
        if (iz == 1 && f.vnt != 1) swrg = 1;
        if (iz == f.nnz && f.vnb != f.nnz) swrg = 1;
        if (swrg) { f.S(iz, ix) = 0; exited = true; break; }
      }
      f.S(iz, ix) = 0;
      viol.erase(id(iz, ix));
      f.downtree();
      np++;
      const int dx[4] = {-1, 1, 0, 0}, dz[4] = {0, 0, -1, 1};
      for (int d = 0; d < 4; d++) {
        int xx = ix + dx[d], xz = iz + dz[d];
        if (xx < 1 || xx > f.nnx || xz < 1 || xz > f.nnz) continue;
        int s0 = f.S(xz, xx);
        if (s0 == -1) { f.fouds2(xz, xx); f.addtree(xz, xx); }
        else if (s0 > 0) { float old = f.T(xz, xx); f.fouds2(xz, xx); if (f.T(xz, xx) > old) viol.insert(id(xz, xx)); f.updtree(xz, xx); }
      }
    }
    st.startup_pops += np;
    if (exited) return;
    for (int p = 1; p <= f.ntr; p++) pq.insert({f.T(f.btg_pz[p], f.btg_px[p]), id(f.btg_pz[p], f.btg_px[p])});
  }
  while (!pq.empty()) {
    auto it = pq.begin();
    float key = it->first;
    int n = it->second;
    int ix = n / f.ld + 1, iz = n % f.ld + 1;
    if (urg == 1) {
      int swrg = 0;
      if (ix == 1 && f.vnl != 1) swrg = 1;
      if (ix == f.nnx && f.vnr != f.nnx) swrg = 1;
      if (iz == 1 && f.vnt != 1) swrg = 1;
      if (iz == f.nnz && f.vnb != f.nnz) swrg = 1;
      if (swrg) { f.S(iz, ix) = 0; break; }
    }
    f.S(iz, ix) = 0;
    pq.erase(it);
    st.pops++;
    // tie statistics: other entries with the same key
    for (auto jt = pq.begin(); jt != pq.end() && jt->first == key; ++jt) {
      st.ties++;
      int m = jt->second, mx = m / f.ld + 1, mz = m % f.ld + 1;
      if (std::abs(mx - ix) + std::abs(mz - iz) <= 4) st.nearties++;
    }
    const int dx[4] = {-1, 1, 0, 0}, dz[4] = {0, 0, -1, 1};
    for (int d = 0; d < 4; d++) {
      int xx = ix + dx[d], xz = iz + dz[d];
      if (xx < 1 || xx > f.nnx || xz < 1 || xz > f.nnz) continue;
      int s = f.S(xz, xx);
      if (s == 0) continue;
      float old = f.T(xz, xx);
      if (s > 0) pq.erase({old, id(xz, xx)});
      f.fouds2(xz, xx);
      float nw = f.T(xz, xx);
      st.updates++;
      if (s > 0 && nw > old) { st.incr++; st.ev.push_back({st.pops, old, nw, id(xz, xx)}); }
      if (nw < key) { st.noncausal++; st.maxdef = std::max(st.maxdef, (double)(key - nw) / key); }
      f.S(xz, xx) = 1;
      pq.insert({nw, id(xz, xx)});
    }
  }
}

static void solve_sorted(Fmm &f, const double *pv, float x, float z, Stats &sr, Stats &sc) {
  f.gridder(pv);
  for (int j = 1; j <= f.nnx; j++)
    for (int k = 1; k <= f.nnz; k++) f.VB(k, j) = f.V(k, j);
  const int nnxb = f.nnx, nnzb = f.nnz;
  const float dnxb = f.dnx, dnzb = f.dnz, goxb = f.gox, gozb = f.goz;
  int isx = (int)((x - f.gox) / f.dnx) + 1, isz = (int)((z - f.goz) / f.dnz) + 1;
  if (isx == f.nnx) isx--;
  if (isz == f.nnz) isz--;
  f.vnl = std::max(1, isx - f.sgs); f.vnr = std::min(f.nnx, isx + f.sgs);
  f.vnt = std::max(1, isz - f.sgs); f.vnb = std::min(f.nnz, isz + f.sgs);
  f.nrnx = (f.vnr - f.vnl) * f.sgdl + 1; f.nrnz = (f.vnb - f.vnt) * f.sgdl + 1;
  f.drnx = f.dvx / (float)(f.gdx * f.sgdl); f.drnz = f.dvz / (float)(f.gdz * f.sgdl);
  f.gorx = f.gox + f.dnx * (float)(f.vnl - 1); f.gorz = f.goz + f.dnz * (float)(f.vnt - 1);
  f.nnx = f.nrnx; f.nnz = f.nrnz; f.dnx = f.drnx; f.dnz = f.drnz; f.gox = f.gorx; f.goz = f.gorz;
  f.bsplrefine();
  travel_sorted(f, x, z, 1, sr);
  f.ttnr = f.ttn; f.nstsr = f.nsts;
  const int ogx = f.vnl, ogz = f.vnt;
  std::fill(f.nsts.begin(), f.nsts.end(), -1);
  for (int k = 1; k <= f.nnz; k += f.sgdl) {
    int idm1 = ogz + (k - 1) / f.sgdl;
    for (int l = 1; l <= f.nnx; l += f.sgdl) {
      int idm2 = ogx + (l - 1) / f.sgdl;
      f.S(idm1, idm2) = f.SR(k, l);
      if (f.S(idm1, idm2) >= 0) f.T(idm1, idm2) = f.TR(k, l);
    }
  }
  f.nnx = nnxb; f.nnz = nnzb; f.dnx = dnxb; f.dnz = dnzb; f.gox = goxb; f.goz = gozb;
  for (int j = 1; j <= f.nnx; j++)
    for (int k = 1; k <= f.nnz; k++) f.V(k, j) = f.VB(k, j);
  for (int k = 1; k <= f.nnx; k++)
    for (int l = 1; l <= f.nnz; l++)
      if (f.S(l, k) == 0) {
        if (l - 1 >= 1 && f.S(l - 1, k) == -1) f.S(l, k) = 1;
        if (l + 1 <= f.nnz && f.S(l + 1, k) == -1) f.S(l, k) = 1;
        if (k - 1 >= 1 && f.S(l, k - 1) == -1) f.S(l, k) = 1;
        if (k + 1 <= f.nnx && f.S(l, k + 1) == -1) f.S(l, k) = 1;
      }
  travel_sorted(f, x, z, 2, sc);
}

int main(int argc, char **argv) {
  int nx = argc > 1 ? atoi(argv[1]) : 35, nsrc = argc > 2 ? atoi(argv[2]) : 8;
  double amp = argc > 3 ? atof(argv[3]) : 0.12;
  int rough = argc > 4 ? atoi(argv[4]) : 0;
  g_startup = argc > 5 ? atoi(argv[5]) : 0;
  int verbose = argc > 6 ? atoi(argv[6]) : 1;
  Fmm a, b;
  a.setup(nx, nx, 26.5f, 120.0f, 0.015f, 0.015f);
  b.setup(nx, nx, 26.5f, 120.0f, 0.015f, 0.015f);
  std::vector<double> pv((size_t)nx * nx);
  std::mt19937 rng(12345);
  std::uniform_real_distribution<double> U(0, 1);
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < nx; j++) {
      double lat = std::sin(0.21 * j + 0.3) * std::cos(0.17 * i) + 0.5 * std::sin(0.05 * i * j / nx + 7);
      if (rough) lat = std::sin(0.5 * i) * std::sin(0.5 * j) + (rough > 1 ? 0.5 * (U(rng) - 0.5) : 0.0);
      pv[(size_t)i * nx + j] = (double)(float)(1.5 * (1.0 + amp * lat));
    }
  float x0 = a.gox, z0 = a.goz, xl = (a.nnx - 1) * a.dnx, zl = (a.nnz - 1) * a.dnz;
  long tot_mis = 0, sweeps_mis = 0;
  for (int s = 0; s < nsrc; s++) {
    float x = x0 + (float)(0.1 + 0.8 * U(rng)) * xl, z = z0 + (float)(0.1 + 0.8 * U(rng)) * zl;
    a.solve_source(pv.data(), x, z);
    Stats sr, sc;
    solve_sorted(b, pv.data(), x, z, sr, sc);
    long mis = 0, misr = 0;
    double maxrel = 0;
    for (int ix = 1; ix <= a.nnx; ix++)
      for (int iz = 1; iz <= a.nnz; iz++) {
        float ta = a.T(iz, ix), tb = b.T(iz, ix);
        if (memcmp(&ta, &tb, 4)) { mis++; maxrel = std::max(maxrel, (double)std::fabs(ta - tb) / ta); }
      }
    for (int ix = 1; ix <= a.nrnx; ix++)
      for (int iz = 1; iz <= a.nrnz; iz++)
        if (a.SR(iz, ix) == 0 && b.SR(iz, ix) == 0) { float ta = a.TR(iz, ix), tb = b.TR(iz, ix); if (memcmp(&ta, &tb, 4)) misr++; }
        else if (a.SR(iz, ix) != b.SR(iz, ix) && (a.SR(iz, ix) == 0 || b.SR(iz, ix) == 0)) misr++;
    printf("maxdeficit %.2e ", sc.maxdef);
    printf("src %d: coarse mism %ld (maxrel %.2e) refined mism %ld | coarse pops %ld upd %ld incr %ld noncausal %ld ties %ld near %ld | refined pops %ld incr %ld noncausal %ld ties %ld near %ld\n",
           s, mis, maxrel, misr, sc.pops, sc.updates, sc.incr, sc.noncausal, sc.ties, sc.nearties, sr.pops, sr.incr, sr.noncausal, sr.ties, sr.nearties);
    {
      std::vector<float> all;
      for (int ix = 1; ix <= b.nnx; ix++) for (int iz = 1; iz <= b.nnz; iz++) all.push_back(b.T(iz, ix));
      std::sort(all.begin(), all.end());
      float hs = b.dnx * b.earth / 1.5f;
      if (verbose) for (auto &e : sc.ev) {
        long n = std::upper_bound(all.begin(), all.end(), e.hi) - std::lower_bound(all.begin(), all.end(), e.lo);
        printf("   incr at pop %ld node(ix=%d,iz=%d) lo %.6f hi %.6f  d/(hs)=%.3g  pops in window %ld\n", e.pop, e.node / b.ld + 1, e.node % b.ld + 1, e.lo, e.hi, (e.hi - e.lo) / hs, n);
      }
    }
    if (mis && verbose >= 2) {
      // earliest mismatching node
      float best = 1e30f; int bx = 0, bz = 0;
      for (int ix = 1; ix <= a.nnx; ix++) for (int iz = 1; iz <= a.nnz; iz++) {
        float ta = a.T(iz, ix), tb = b.T(iz, ix);
        if (memcmp(&ta, &tb, 4) && std::min(ta, tb) < best) { best = std::min(ta, tb); bx = ix; bz = iz; }
      }
      printf("  earliest mismatch at ix=%d iz=%d: exact %.9g sorted %.9g\n", bx, bz, a.T(bz, bx), b.T(bz, bx));
      for (int dz = -3; dz <= 3; dz++) { for (int dx = -3; dx <= 3; dx++) {
          int x = bx + dx, z = bz + dz; if (x < 1 || x > a.nnx || z < 1 || z > a.nnz) { printf("      --      "); continue; }
          printf(" %.9g%c", a.T(z, x), a.T(z, x) == b.T(z, x) ? ' ' : '*'); } printf("\n"); }
    }
    tot_mis += mis;
    if (mis || misr) sweeps_mis++;
  }
  printf("sweeps with any mismatch: %ld / %d ; total coarse mismatches %ld\n", sweeps_mis, nsrc, tot_mis);
}
