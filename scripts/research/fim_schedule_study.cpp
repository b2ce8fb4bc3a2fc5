// Research harness (not product, not a test): the tile SCHEDULE of the fast-iterative coarse pass when several warps share
// one sweep.  Same per-node code and checks as tests/host/fim_host_check.cpp; FIM_SCHEDULE=rb runs every round in two
// phases (tiles with tx + tz even, then odd): tiles of one colour never read what another tile of that colour writes, so W
// warps can relax them concurrently and the result stays independent of timing.  Prints, besides the usual summary, the
// critical path in diagonal steps for W = 1, 2, 4, 8 warps per sweep (greedy assignment of a phase's tiles to warps).
//   g++ -O2 -std=c++17 -ffp-contract=off -I. scripts/research/fim_schedule_study.cpp -Loracle -loracle -Wl,-rpath,$PWD/oracle -o /tmp/fim_sched
// Host replay of the block-level fast-iterative coarse pass (dsurftomo_b200/csrc/eik_fim.cuh -- the per-node code the
// device kernel k_fim_march runs) against the oracle's heap march Fmm::travel(urg=2) (oracle/fmm.cpp, restating
// src/CalSurfG.f90:288-487).  The refined pass and the injection are the oracle's; the coarse pass is relaxed tile by
// tile in rounds exactly as the kernel schedules it (every active tile once per round, lanes of an anti-diagonal in
// turn).  Reports, per grid: nodes whose time differs from the reference (count, relative size), rays whose B-spline
// vertex pattern / values differ, and the work done (tile activations, rule evaluations per node).
// Build/run: tests/test_fim_host.py (g++ -O2 -ffp-contract=off, links oracle/liboracle.so).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#define DSURF_FIM_CROSSCHECK 1
#include "../../dsurftomo_b200/csrc/eik_fim.cuh"
#include <cstdlib>
#include "../../oracle/fmm.h"

using namespace dsurf::fim;
using oracle::Fmm;

struct Stats {
  long activations = 0, walks = 0, steps = 0, evals = 0, rounds = 0, empty_activations = 0;
  double crit[4] = {0, 0, 0, 0};  // critical path (diagonal steps + 40 per activation for load/store) with 1, 2, 4, 8 warps
};

struct HostSweep {
  Layout L;
  std::vector<uint32_t> T, bitmap;
  std::vector<unsigned char> active;
  std::vector<int> box;  // (time bits, status) pairs
  TileCtx base;
  const float *vel;  // [nnx][nnz]
  std::vector<float> ris;
  int srcx, srcz;
  Stats st;
  int max_walks = 8;

  void mark_global(int gx, int gz) {
    if (gx < 0 || gx >= base.nnx || gz < 0 || gz >= base.nnz) return;
    const int tx = gx / kT, tz = gz / kT;
    bitmap[(size_t)(tx * L.ntz + tz) * kT + (gx - tx * kT)] |= 1u << (gz - tz * kT);
    active[tx * L.ntz + tz] = 1;
  }

  void process_tile(int tx, int tz) {
    static Tile tl;
    TileCtx C = base;
    C.gx0 = tx * kT;
    C.gz0 = tz * kT;
    // ---- load
    for (int r = 0; r < kRows; r++)
      for (int c = 0; c < kPitch; c++) {
        const int gx = C.gx0 + r - kHX, gz = C.gz0 + c - kHZ;
        uint32_t w = T[(size_t)(gx + kHX) * L.pitch + (gz + kHZ)];
        if ((int)w < 0) w = kInf;
        const int bx = gx - C.bx0, bz = gz - C.bz0;
        if (bx >= 0 && bx < C.bw && bz >= 0 && bz < C.bh && C.box[2 * (bx * C.bh + bz) + 1] == 0) w |= kInit;
        tl.t[r * kPitch + c] = w;
      }
    for (int x = 0; x < kT; x++) {
      const int gx = C.gx0 + x;
      tl.risti[x] = gx < C.nnx ? ris[gx] : 0.0f;
      for (int z = 0; z < kT; z++) {
        const int gz = C.gz0 + z;
        tl.slow[x * kT + z] = (gx < C.nnx && gz < C.nnz) ? 1.0f / vel[(size_t)gx * C.nnz + gz] : 1.0f;
      }
      uint32_t &bm = bitmap[(size_t)(tx * L.ntz + tz) * kT + x];
      tl.dirty[x] = bm;
      bm = 0;
    }
    for (int i = 0; i < 4; i++) tl.hx[i] = tl.hz[i] = 0;
    st.activations++;
    bool any = false;
    for (int x = 0; x < kT; x++) any |= tl.dirty[x] != 0;
    if (!any) {
      st.empty_activations++;
      return;
    }
    // ---- relax: anti-diagonal walks, first away from the source
    const int sx0 = (C.gx0 + kT / 2 >= srcx) ? 1 : -1, sz0 = (C.gz0 + kT / 2 >= srcz) ? 1 : -1;
    bool changed = false;
    for (int w = 0; w < max_walks * 4; w++) {
      bool left = false;
      for (int x = 0; x < kT; x++) left |= tl.dirty[x] != 0;
      if (!left) break;
      const int k = w & 3;
      const int sx = (k & 1) ? -sx0 : sx0, sz = (k & 2) ? -sz0 : sz0;
      st.walks++;
      for (int d = 0; d < 2 * kT - 1; d++) {
        bool anyd = false;
        for (int x = 0; x < kT; x++) {
          const int z = diag_z(x, d, sx, sz);
          if (z >= 0 && ((tl.dirty[x] >> z) & 1u)) anyd = true;
        }
        if (!anyd) continue;
        st.steps++;
        // lanes of the diagonal: decide on the state before the step (as the warp does), then relax
        int zs[kT];
        for (int x = 0; x < kT; x++) {
          const int z = diag_z(x, d, sx, sz);
          zs[x] = (z >= 0 && ((tl.dirty[x] >> z) & 1u)) ? z : -1;
        }
        for (int x = 0; x < kT; x++)
          if (zs[x] >= 0) {
            st.evals++;
            changed |= relax_node(tl, C, x, zs[x], tl.slow[x * kT + zs[x]]);
          }
      }
    }
    // ---- unload
    if (changed)
      for (int x = 0; x < kT; x++)
        for (int z = 0; z < kT; z++) {
          const uint32_t w = *tl.at(x, z);
          if ((int)w < 0) continue;
          const int gx = C.gx0 + x, gz = C.gz0 + z;
          if (gx >= C.nnx || gz >= C.nnz) continue;
          T[L.at(gx, gz)] = (w == kInf) ? kFarG : w;
        }
    for (int i = 0; i < 4; i++) {
      const int ux = i < 2 ? i - 2 : kT + i - 2;
      for (int b = 0; b < kT; b++) {
        if ((tl.hx[i] >> b) & 1u) mark_global(C.gx0 + ux, C.gz0 + b);
        if ((tl.hz[i] >> b) & 1u) mark_global(C.gx0 + b, C.gz0 + ux);
      }
    }
    for (int x = 0; x < kT; x++)
      if (tl.dirty[x]) {  // walk limit reached: the tile stays active
        bitmap[(size_t)(tx * L.ntz + tz) * kT + x] |= tl.dirty[x];
        active[tx * L.ntz + tz] = 1;
      }
  }

  void run() {
    const bool rb = getenv("FIM_SCHEDULE") && !strcmp(getenv("FIM_SCHEDULE"), "rb");
    std::vector<int> list;
    for (;;) {
      bool any = false;
      for (int i = 0; i < L.ntx * L.ntz; i++) any |= active[i] != 0;
      if (!any) break;
      st.rounds++;
      for (int phase = 0; phase < (rb ? 2 : 1); phase++) {
        list.clear();
        for (int i = 0; i < L.ntx * L.ntz; i++)
          if (active[i] && (!rb || ((i / L.ntz + i % L.ntz) & 1) == phase)) {
            list.push_back(i);
            active[i] = 0;
          }
        std::vector<double> cost;
        for (int i : list) {
          const long s0 = st.steps;
          process_tile(i / L.ntz, i % L.ntz);
          cost.push_back((double)(st.steps - s0) + 40.0);
        }
        // greedy longest-first assignment to W warps; without phases (index order) only W = 1 is deterministic
        std::sort(cost.rbegin(), cost.rend());
        for (int wi = 0; wi < 4; wi++) {
          const int W = 1 << wi;
          std::vector<double> load(W, 0.0);
          for (double c : cost) *std::min_element(load.begin(), load.end()) += c;
          st.crit[wi] += *std::max_element(load.begin(), load.end()) + (W > 1 ? 20.0 : 0.0);  // + a block barrier per phase
        }
      }
    }
  }
};

// the oracle's solve_source up to (and including) the injection; then the coarse pass by tiles
static long g_start_pops = 0, g_start_alive = 0, g_start_alive_mismatch = 0;
static const Fmm *g_exact = nullptr;  // the oracle's finished solve of the same source (checker of the start-up hand-over)
static bool no_startup = false;
static void solve_fim(Fmm &f, const double *pv, float x, float z, Stats &st, int max_walks) {
  f.gridder(pv);
  for (int j = 1; j <= f.nnx; j++)
    for (int k = 1; k <= f.nnz; k++) f.VB(k, j) = f.V(k, j);
  const int nnxb = f.nnx, nnzb = f.nnz;
  const float dnxb = f.dnx, dnzb = f.dnz, goxb = f.gox, gozb = f.goz;
  int isx = (int)((x - f.gox) / f.dnx) + 1, isz = (int)((z - f.goz) / f.dnz) + 1;
  if (isx == f.nnx) isx--;
  if (isz == f.nnz) isz--;
  f.vnl = std::max(1, isx - f.sgs);
  f.vnr = std::min(f.nnx, isx + f.sgs);
  f.vnt = std::max(1, isz - f.sgs);
  f.vnb = std::min(f.nnz, isz + f.sgs);
  f.nrnx = (f.vnr - f.vnl) * f.sgdl + 1;
  f.nrnz = (f.vnb - f.vnt) * f.sgdl + 1;
  f.drnx = f.dvx / (float)(f.gdx * f.sgdl);
  f.drnz = f.dvz / (float)(f.gdz * f.sgdl);
  f.gorx = f.gox + f.dnx * (float)(f.vnl - 1);
  f.gorz = f.goz + f.dnz * (float)(f.vnt - 1);
  f.nnx = f.nrnx; f.nnz = f.nrnz; f.dnx = f.drnx; f.dnz = f.drnz; f.gox = f.gorx; f.goz = f.gorz;
  f.bsplrefine();
  f.travel(x, z, 1);
  f.ttnr = f.ttn;
  f.nstsr = f.nsts;
  const int ogx = f.vnl, ogz = f.vnt;
  std::fill(f.nsts.begin(), f.nsts.end(), -1);
  for (int k = 1; k <= f.nnz; k += f.sgdl) {
    const int idm1 = ogz + (k - 1) / f.sgdl;
    for (int l = 1; l <= f.nnx; l += f.sgdl) {
      const int idm2 = ogx + (l - 1) / f.sgdl;
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
  // ---- what k_refine leaves: the refined box as (time bits, status) pairs
  const int bx0 = f.vnl - 1, bz0 = f.vnt - 1, bw = f.vnr - f.vnl + 1, bh = f.vnb - f.vnt + 1;
  std::vector<int> box((size_t)2 * bw * bh, 0);
  for (int bx = 0; bx < bw; bx++)
    for (int bz = 0; bz < bh; bz++) {
      const int ix = bx0 + bx + 1, iz = bz0 + bz + 1;
      const int s = f.S(iz, ix);
      float t = f.T(iz, ix);
      int tb;
      memcpy(&tb, &t, 4);
      box[2 * (bx * bh + bz)] = s >= 0 ? tb : 0;
      box[2 * (bx * bh + bz) + 1] = s;
    }
  std::vector<float> vel((size_t)f.nnx * f.nnz);
  for (int ix = 1; ix <= f.nnx; ix++)
    for (int iz = 1; iz <= f.nnz; iz++) vel[(size_t)(ix - 1) * f.nnz + iz - 1] = f.V(iz, ix);
  std::vector<float> ris(f.nnx);
  for (int ix = 1; ix <= f.nnx; ix++) ris[ix - 1] = f.earth * std::sin(f.gox + (float)(ix - 1) * f.dnx);
  // ---- start-up: exact heap march on the region (one thread per sweep on the device)
  StartCtx SC;
  SC.nnx = f.nnx;
  SC.nnz = f.nnz;
  SC.rx0 = std::max(0, isx - 1 - kRegHalf);
  SC.rz0 = std::max(0, isz - 1 - kRegHalf);
  SC.rw = std::min(f.nnx - 1, isx + kRegHalf) - SC.rx0 + 1;
  SC.rh = std::min(f.nnz - 1, isz + kRegHalf) - SC.rz0 + 1;
  SC.ri = f.earth;
  SC.dnx = f.dnx;
  SC.dnz = f.dnz;
  SC.vel = vel.data();
  SC.risti = ris.data();
  std::vector<uint32_t> rwords((size_t)SC.rw * SC.rh);
  std::vector<dsurf::lps::Ent> rheap((size_t)SC.rw * SC.rh + 1);
  std::vector<unsigned char> rflag((size_t)SC.rw * SC.rh);
  StartMem SM{rwords.data(), rheap.data(), rflag.data()};
  int ntr = 0;
  g_start_pops += no_startup ? 0 : startup_march(SC, SM, box.data(), bx0, bz0, bw, bh, ntr, 1 << 30);
  // every node the start-up march accepted must carry the reference's final time, bit for bit (it IS the reference's march)
  if (!no_startup && g_exact)
    for (int i = 0; i < SC.rw * SC.rh; i++)
      if (dsurf::lps::alive(rwords[i])) {
        const int ix = SC.rx0 + i / SC.rh + 1, iz = SC.rz0 + i % SC.rh + 1;
        const float te = const_cast<Fmm *>(g_exact)->T(iz, ix);
        uint32_t eb;
        memcpy(&eb, &te, 4);
        g_start_alive++;
        if (eb != rwords[i]) g_start_alive_mismatch++;
      }
  // ---- hand the state to the tile solver
  HostSweep hs;
  hs.max_walks = max_walks;
  hs.L = make_layout(f.nnx, f.nnz);
  hs.T.assign(hs.L.words(), kFarG);
  hs.bitmap.assign((size_t)hs.L.ntx * hs.L.ntz * kT, 0);
  hs.active.assign((size_t)hs.L.ntx * hs.L.ntz, 0);
  TileCtx &C = hs.base;
  C.nnx = f.nnx;
  C.nnz = f.nnz;
  C.ri = f.earth;
  C.dnx = f.dnx;
  C.dnz = f.dnz;
  if (no_startup) {
    C.bx0 = bx0; C.bz0 = bz0; C.bw = bw; C.bh = bh;
    hs.box = box;
    for (size_t i = 1; i < hs.box.size(); i += 2) if (hs.box[i] == -100) hs.box[i] = 1;
  } else {
    C.bx0 = SC.rx0; C.bz0 = SC.rz0; C.bw = SC.rw; C.bh = SC.rh;
    hs.box.assign((size_t)2 * C.bw * C.bh, 0);
    for (int i = 0; i < C.bw * C.bh; i++) {
      const uint32_t w = rwords[i];
      if (dsurf::lps::alive(w)) {
        hs.box[2 * i] = (int)w;
        hs.box[2 * i + 1] = 0;
      } else if (w == dsurf::lps::kFar) {
        hs.box[2 * i + 1] = -1;
      } else {
        hs.box[2 * i] = rheap[w & 0x7FFFFFFFu].x;
        hs.box[2 * i + 1] = 1;
      }
    }
  }
  C.box = hs.box.data();
  for (int bx = 0; bx < C.bw; bx++)
    for (int bz = 0; bz < C.bh; bz++) {
      const int st = hs.box[2 * (bx * C.bh + bz) + 1];
      if (st >= 0) hs.T[hs.L.at(C.bx0 + bx, C.bz0 + bz)] = (uint32_t)hs.box[2 * (bx * C.bh + bz)];
      if (st > 0) {
        const int gx = C.bx0 + bx, gz = C.bz0 + bz;
        const int nx[5] = {gx, gx - 1, gx + 1, gx, gx}, nz[5] = {gz, gz, gz, gz - 1, gz + 1};
        for (int q = 0; q < 5; q++) {
          const int ex = nx[q] - C.bx0, ez = nz[q] - C.bz0;
          const bool inbox = ex >= 0 && ex < C.bw && ez >= 0 && ez < C.bh;
          if (inbox && hs.box[2 * (ex * C.bh + ez) + 1] == 0) continue;  // alive before the pass
          hs.mark_global(nx[q], nz[q]);
        }
      }
    }
  hs.vel = vel.data();
  hs.ris = ris;
  hs.srcx = isx - 1;
  hs.srcz = isz - 1;
  hs.run();
  for (int ix = 1; ix <= f.nnx; ix++)
    for (int iz = 1; iz <= f.nnz; iz++) {
      const uint32_t w = hs.T[hs.L.at(ix - 1, iz - 1)];
      float t;
      memcpy(&t, &w, 4);
      f.T(iz, ix) = t;
      f.S(iz, ix) = 0;
    }
  st = hs.st;
}

int main(int argc, char **argv) {
  const int nx = argc > 1 ? atoi(argv[1]) : 35, nsrc = argc > 2 ? atoi(argv[2]) : 6;
  const int rough = argc > 3 ? atoi(argv[3]) : 0, nrecv = argc > 4 ? atoi(argv[4]) : 8;
  const int max_walks = argc > 5 ? atoi(argv[5]) : 8;
  const double amp = 0.12;
  no_startup = getenv("FIM_NO_STARTUP") != nullptr;
  Fmm a, b;
  a.setup(nx, nx, 26.5f, 120.0f, 0.015f, 0.015f);
  b.setup(nx, nx, 26.5f, 120.0f, 0.015f, 0.015f);
  std::vector<double> pv((size_t)nx * nx);
  std::mt19937 rng(12345);
  std::uniform_real_distribution<double> U(0, 1);
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < nx; j++) {
      double lat = std::sin(0.21 * j + 0.3) * std::cos(0.17 * i) + 0.5 * std::sin(0.05 * i * j / nx + 7);
      if (rough == 1) lat = std::sin(0.5 * i) * std::sin(0.5 * j);
      if (rough == 2) lat = ((i / 4 + j / 4) & 1) ? 1.0 : -1.0;  // blocky
      if (rough == 3) lat = 0.0;                                  // uniform: ties everywhere
      pv[(size_t)i * nx + j] = (double)(float)(1.5 * (1.0 + amp * lat));
    }
  const float x0 = a.gox, z0 = a.goz, xl = (a.nnx - 1) * a.dnx, zl = (a.nnz - 1) * a.dnz;
  long tot_mis = 0, tot_nodes = 0, rays = 0, rays_pattern = 0, rays_val = 0, sweeps_mis = 0, unreached = 0;
  double worst = 0, worst_dt = 0;
  Stats tot;
  const int nv = (a.nvx + 2) * (a.nvz + 2);
  std::vector<float> fa(nv), fb(nv);
  for (int s = 0; s < nsrc; s++) {
    float x = x0 + (float)(0.1 + 0.8 * U(rng)) * xl, z = z0 + (float)(0.1 + 0.8 * U(rng)) * zl;
    if (s == 1) {  // a source in the corner cell: clipped refined box
      x = x0 + 0.004f * xl;
      z = z0 + 0.99f * zl;
    }
    a.error = 0;
    a.solve_source(pv.data(), x, z);
    if (a.error) {
      printf("src %d: outside the grid, skipped\n", s);
      continue;
    }
    Stats st;
    g_exact = &a;
    solve_fim(b, pv.data(), x, z, st, max_walks);
    long mis = 0;
    double maxrel = 0;
    for (int ix = 1; ix <= a.nnx; ix++)
      for (int iz = 1; iz <= a.nnz; iz++) {
        const float ta = a.T(iz, ix), tb = b.T(iz, ix);
        if (!(tb < 1e30f)) unreached++;
        if (memcmp(&ta, &tb, 4)) {
          mis++;
          maxrel = std::max(maxrel, std::fabs((double)ta - tb) / ta);
        }
      }
    if (mis && getenv("FIM_DEBUG")) {
      float best = 1e30f; int bx = 0, bz = 0;
      for (int ix = 1; ix <= a.nnx; ix++) for (int iz = 1; iz <= a.nnz; iz++) {
        float ta = a.T(iz, ix), tb = b.T(iz, ix);
        if (memcmp(&ta, &tb, 4) && std::min(ta, tb) < best) { best = std::min(ta, tb); bx = ix; bz = iz; }
      }
      printf("  earliest mismatch at ix=%d iz=%d: exact %.9g fim %.9g (box %d..%d x %d..%d)\n", bx, bz, a.T(bz, bx), b.T(bz, bx), b.vnl, b.vnr, b.vnt, b.vnb);
      for (int dz = -3; dz <= 3; dz++) { for (int dx = -3; dx <= 3; dx++) {
          int x = bx + dx, z = bz + dz; if (x < 1 || x > a.nnx || z < 1 || z > a.nnz) { printf("      --      "); continue; }
          printf(" %.9g%c", a.T(z, x), a.T(z, x) == b.T(z, x) ? ' ' : '*'); } printf("\n"); }
    }
    long rp = 0, rv = 0;
    for (int r = 0; r < nrecv; r++) {
      const float rx = x0 + (float)(0.05 + 0.9 * U(rng)) * xl, rz = z0 + (float)(0.05 + 0.9 * U(rng)) * zl;
      const float ta = a.srtimes(x, z, rx, rz), tb = b.srtimes(x, z, rx, rz);
      worst_dt = std::max(worst_dt, std::fabs((double)ta - tb) / ta);
      std::fill(fa.begin(), fa.end(), 0.0f);
      std::fill(fb.begin(), fb.end(), 0.0f);
      a.rpaths(x, z, fa.data(), rx, rz);
      b.rpaths(x, z, fb.data(), rx, rz);
      bool pat = false, val = false;
      for (int i = 0; i < nv; i++) {
        if ((fa[i] != 0.0f) != (fb[i] != 0.0f)) pat = true;
        if (memcmp(&fa[i], &fb[i], 4)) val = true;
      }
      rays++;
      rp += pat;
      rv += val;
    }
    rays_pattern += rp;
    rays_val += rv;
    const double N = (double)a.nnx * a.nnz;
    printf("src %d: mismatches=%ld of %.0f (max rel %.2e) | rounds %ld, tile activations %ld (%.2f per tile, %ld empty), walks %ld, diagonal steps %ld, "
           "rule evaluations %.2f per node | rays: pattern differs %ld / %d, value differs %ld\n",
           s, mis, N, maxrel, st.rounds, st.activations, (double)st.activations / ((a.nnx + kT - 1) / kT * ((a.nnz + kT - 1) / kT)),
           st.empty_activations, st.walks, st.steps, st.evals / N, rp, nrecv, rv);
    tot_mis += mis;
    tot_nodes += (long)N;
    worst = std::max(worst, maxrel);
    if (mis) sweeps_mis++;
    tot.activations += st.activations;
    tot.steps += st.steps;
    tot.evals += st.evals;
    tot.rounds += st.rounds;
    tot.walks += st.walks;
    for (int wi = 0; wi < 4; wi++) tot.crit[wi] += st.crit[wi];
  }
  printf("SUMMARY grid=%d rough=%d sweeps=%d sweeps_with_mismatch=%ld mismatch_nodes=%ld nodes=%ld mismatch_frac=%.3e max_rel=%.3e unreached=%ld "
         "rays=%ld rays_pattern_diff=%ld rays_value_diff=%ld max_rel_dt=%.3e | per sweep: rounds %.1f activations %.1f walks %.1f steps %.1f evals_per_node %.2f\n",
         a.nnx, rough, nsrc, sweeps_mis, tot_mis, tot_nodes, (double)tot_mis / tot_nodes, worst, unreached, rays, rays_pattern, rays_val, worst_dt,
         (double)tot.rounds / nsrc, (double)tot.activations / nsrc, (double)tot.walks / nsrc, (double)tot.steps / nsrc, (double)tot.evals / tot_nodes);
  printf("critical path per sweep (diagonal steps incl. 40 per activation): W=1 %.0f  W=2 %.0f  W=4 %.0f  W=8 %.0f  (schedule %s)\n",
         tot.crit[0] / nsrc, tot.crit[1] / nsrc, tot.crit[2] / nsrc, tot.crit[3] / nsrc, getenv("FIM_SCHEDULE") ? getenv("FIM_SCHEDULE") : "index order");
  printf("start-up pops per sweep: %.1f; nodes alive at hand-over %ld, differing from the reference's final times: %ld\n",
         (double)g_start_pops / nsrc, g_start_alive, g_start_alive_mismatch);
  printf("rule cross-check: evaluations=%ld cached_vs_plain_mismatch=%ld handed_to_generic=%ld (%.3f %%)\n", dsurf::fim::g_cross_total,
         dsurf::fim::g_cross_mismatch, dsurf::fim::g_generic_calls, 100.0 * dsurf::fim::g_generic_calls / std::max(1L, dsurf::fim::g_cross_total));
  const bool ok = dsurf::fim::g_cross_mismatch == 0 && g_start_alive_mismatch == 0 && unreached == 0 && worst <= 1e-5 && (double)tot_mis / tot_nodes <= 5e-2 && rays_pattern == 0;
  printf(ok ? "FIM HOST CHECK OK\n" : "FIM HOST CHECK FAILED\n");
  return ok ? 0 : 1;
}
