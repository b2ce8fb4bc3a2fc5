// Research harness (not product, not oracle): the CAUSAL LOCAL RULE of the reference's fast marching, iterated to its
// fixed point in arbitrary order (what a block-level fast-iterative sweep computes), against the exact heap march.
//
// In travel (CalSurfG.f90:386-486) a node's time is the LAST trial value fouds2 wrote before the node was popped; a
// trial value is (re)computed whenever one of its four neighbours is popped, from the nodes alive at that moment.
// While the heap is a valid heap and no two interacting keys are equal, nodes are popped in increasing time, so the
// final time of node C is given by a rule that looks at C's stencil only:
//     v = seed value of C (coarse pass: close nodes injected from the refined grid) or +inf
//     for the not-initially-alive neighbours J of C in increasing T(J):   if T(J) < v:  v = fouds2(C | alive = initially
//         alive nodes and nodes with T <= T(J))   else stop
// The rule is causal (v depends only on nodes with smaller times), so its fixed point is unique and any iteration order
// reaches it; every evaluation is the reference's fp32 arithmetic, so wherever the heap order equals the time order the
// result is BIT-IDENTICAL to the reference.  The heap order differs from the time order at equal keys (heap layout
// decides) and after updtree raised a key (the reference only sifts up) -- this harness measures how often that matters.
//   g++ -O2 -std=c++17 -ffp-contract=off -I oracle scripts/research/fim_rule_study.cpp oracle/fmm.cpp -o /tmp/fim_study
#include "fmm.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <random>
#include <vector>
using namespace oracle;

static const float INF = std::numeric_limits<float>::infinity();

struct Rule {
  Fmm &f;
  std::vector<float> T, seed;       // current times; seed value (INF if none)
  std::vector<char> init_alive;
  long evals = 0, fevals = 0;
  explicit Rule(Fmm &ff) : f(ff) {}
  int id(int iz, int ix) const { return (ix - 1) * f.ld + (iz - 1); }
  // fouds2 with the alive predicate "initially alive, or T <= t" evaluated through the oracle: temporarily publish
  // statuses/times of the 8 stencil nodes into f and call f.fouds2
  float F(int iz, int ix, float t) {
    fevals++;
    static const int dx[8] = {-1, -2, 1, 2, 0, 0, 0, 0}, dz[8] = {0, 0, 0, 0, -1, -2, 1, 2};
    for (int q = 0; q < 8; q++) {
      const int x = ix + dx[q], z = iz + dz[q];
      if (x < 1 || x > f.nnx || z < 1 || z > f.nnz) continue;
      const int n = id(z, x);
      const bool al = init_alive[n] || T[n] <= t;
      f.S(z, x) = al ? 0 : -1;
      f.T(z, x) = al ? T[n] : 0.0f;
    }
    f.fouds2(iz, ix);
    return f.T(iz, ix);
  }
  float G(int iz, int ix) {
    evals++;
    const int n = id(iz, ix);
    float v = seed[n];
    float tj[4];
    int nj = 0;
    const int dx[4] = {-1, 1, 0, 0}, dz[4] = {0, 0, -1, 1};
    for (int d = 0; d < 4; d++) {
      const int x = ix + dx[d], z = iz + dz[d];
      if (x < 1 || x > f.nnx || z < 1 || z > f.nnz) continue;
      const int m = id(z, x);
      if (init_alive[m] || !(T[m] < INF)) continue;
      tj[nj++] = T[m];
    }
    std::sort(tj, tj + nj);
    for (int k = 0; k < nj; k++) {
      if (!(tj[k] < v)) break;
      if (k + 1 < nj && tj[k + 1] == tj[k]) continue;  // equal neighbours are accepted back to back: evaluate once with both
      v = F(iz, ix, tj[k]);
    }
    return v;
  }
};

// coarse pass by the rule; f holds the injected state (S: 0 alive, >0 close, -1 far; T of alive/close nodes)
static void travel_rule(Fmm &f, long &evals, long &fevals, long &rounds) {
  Rule R(f);
  const size_t N = f.ttn.size();
  R.T.assign(N, INF);
  R.seed.assign(N, INF);
  R.init_alive.assign(N, 0);
  std::deque<int> q;
  std::vector<char> inq(N, 0);
  auto push = [&](int iz, int ix) {
    if (ix < 1 || ix > f.nnx || iz < 1 || iz > f.nnz) return;
    const int n = R.id(iz, ix);
    if (R.init_alive[n] || inq[n]) return;
    inq[n] = 1;
    q.push_back(n);
  };
  for (int ix = 1; ix <= f.nnx; ix++)
    for (int iz = 1; iz <= f.nnz; iz++) {
      const int n = R.id(iz, ix);
      if (f.S(iz, ix) == 0) {
        R.init_alive[n] = 1;
        R.T[n] = f.T(iz, ix);
      } else if (f.S(iz, ix) > 0) {
        R.seed[n] = f.T(iz, ix);
        R.T[n] = f.T(iz, ix);
      }
    }
  for (int ix = 1; ix <= f.nnx; ix++)
    for (int iz = 1; iz <= f.nnz; iz++)
      if (f.S(iz, ix) > 0) {
        push(iz, ix);
        push(iz, ix - 1), push(iz, ix + 1), push(iz - 1, ix), push(iz + 1, ix);
      }
  while (!q.empty()) {
    const int n = q.front();
    q.pop_front();
    inq[n] = 0;
    const int ix = n / f.ld + 1, iz = n % f.ld + 1;
    const float v = R.G(iz, ix);
    if (memcmp(&v, &R.T[n], 4)) {
      R.T[n] = v;
      // dist-1 and dist-2 stencil users
      for (int d = 1; d <= 2; d++) {
        push(iz, ix - d), push(iz, ix + d), push(iz - d, ix), push(iz + d, ix);
      }
    }
    if (R.evals > 400ll * (long)f.nnx * f.nnz) {
      fprintf(stderr, "no convergence\n");
      break;
    }
  }
  for (int ix = 1; ix <= f.nnx; ix++)
    for (int iz = 1; iz <= f.nnz; iz++) {
      f.T(iz, ix) = R.T[R.id(iz, ix)];
      f.S(iz, ix) = 0;
    }
  evals = R.evals;
  fevals = R.fevals;
  rounds = 0;
}

static void solve_rule(Fmm &f, const double *pv, float x, float z, long &evals, long &fevals) {
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
  f.travel(x, z, 1);  // refined grid: the exact heap march (early exit makes its alive SET order-dependent)
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
  f.nnxr = f.nnx; f.nnzr = f.nnz; f.goxr = f.gox; f.gozr = f.goz; f.dnxr = f.dnx; f.dnzr = f.dnz;
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
  long rounds;
  travel_rule(f, evals, fevals, rounds);
}

int main(int argc, char **argv) {
  int nx = argc > 1 ? atoi(argv[1]) : 35, nsrc = argc > 2 ? atoi(argv[2]) : 8;
  double amp = argc > 3 ? atof(argv[3]) : 0.12;
  int rough = argc > 4 ? atoi(argv[4]) : 0;
  int nrecv = argc > 5 ? atoi(argv[5]) : 16;
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
  long tot_mis = 0, sweeps_mis = 0, tot_nodes = 0, rays = 0, rays_pattern = 0, rays_val = 0, cells = 0, cells_flip = 0;
  double worst = 0, worst_t = 0;
  std::vector<double> rels;
  const int nv = (a.nvx + 2) * (a.nvz + 2);
  std::vector<float> fa(nv), fb(nv);
  for (int s = 0; s < nsrc; s++) {
    float x = x0 + (float)(0.1 + 0.8 * U(rng)) * xl, z = z0 + (float)(0.1 + 0.8 * U(rng)) * zl;
    a.solve_source(pv.data(), x, z);
    long evals, fevals;
    solve_rule(b, pv.data(), x, z, evals, fevals);
    long mis = 0;
    double maxrel = 0;
    for (int ix = 1; ix <= a.nnx; ix++)
      for (int iz = 1; iz <= a.nnz; iz++) {
        float ta = a.T(iz, ix), tb = b.T(iz, ix);
        if (memcmp(&ta, &tb, 4)) {
          mis++;
          double r = std::fabs((double)ta - tb) / ta;
          maxrel = std::max(maxrel, r);
          rels.push_back(r);
        }
      }
    // rays to random receivers through both fields
    long rp = 0, rv = 0;
    double maxdt = 0;
    for (int r = 0; r < nrecv; r++) {
      float rx = x0 + (float)(0.05 + 0.9 * U(rng)) * xl, rz = z0 + (float)(0.05 + 0.9 * U(rng)) * zl;
      float ta = a.srtimes(x, z, rx, rz), tb = b.srtimes(x, z, rx, rz);
      maxdt = std::max(maxdt, std::fabs((double)ta - tb) / ta);
      std::fill(fa.begin(), fa.end(), 0.0f);
      std::fill(fb.begin(), fb.end(), 0.0f);
      a.rpaths(x, z, fa.data(), rx, rz);
      b.rpaths(x, z, fb.data(), rx, rz);
      bool pat = false, val = false;
      for (int i = 0; i < nv; i++) {
        if ((fa[i] != 0.0f) != (fb[i] != 0.0f)) { pat = true; cells_flip++; }
        if (memcmp(&fa[i], &fb[i], 4)) val = true;
        if (fa[i] != 0.0f) cells++;
      }
      rays++;
      rp += pat;
      rv += val;
    }
    rays_pattern += rp;
    rays_val += rv;
    worst_t = std::max(worst_t, maxdt);
    printf("src %d: mismatching nodes %ld / %ld (max rel %.2e) | rule evaluations %.2f per node, fouds2 %.2f per node | rays: pattern differs %ld / %d, "
           "any value differs %ld, max rel dt %.2e\n", s, mis, (long)a.nnx * a.nnz, maxrel, (double)evals / (a.nnx * a.nnz), (double)fevals / (a.nnx * a.nnz), rp, nrecv, rv, maxdt);
    tot_mis += mis;
    tot_nodes += (long)a.nnx * a.nnz;
    worst = std::max(worst, maxrel);
    if (mis) sweeps_mis++;
  }
  std::sort(rels.begin(), rels.end());
  auto pct = [&](double p) { return rels.empty() ? 0.0 : rels[std::min(rels.size() - 1, (size_t)(p * rels.size()))]; };
  printf("SUMMARY grid %d^2: sweeps with any mismatch %ld / %d ; mismatching nodes %ld / %ld (%.4f %%) ; |dT|/T of mismatching nodes p50 %.2e p99 %.2e max %.2e ; "
         "rays with a different vertex pattern %ld / %ld, vertex entries flipped %ld / %ld, rays with any differing value %ld, max rel receiver-time difference %.2e\n",
         a.nnx, sweeps_mis, nsrc, tot_mis, tot_nodes, 100.0 * tot_mis / tot_nodes, pct(0.5), pct(0.99), worst, rays_pattern, rays, cells_flip, cells, rays_val, worst_t);
}
