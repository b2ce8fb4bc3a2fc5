// K2/K3 -- B-spline dicing and the 2-D eikonal solve on the spherical-shell grid, replacing
// gridder (src/CalSurfG.f90:1460-1553), bsplrefine (:1562-1628), travel/fouds2/addtree/downtree/
// updtree (:288-921) and the source-grid refinement orchestration of CalSurfG (:1193-1355).
//
// EXACT-ORDER REPLAY.  The reference's fast-marching result depends on the acceptance order
// (fouds2 overwrites a trial value instead of taking a min, the mixed-order stencil is upgraded as
// neighbours become alive, updtree only sifts up, ties are broken by the heap layout -- SURVEY.md
// section 7, hard part 1), and ray cell indices flip on 1-ulp differences of T.  Every sweep is
// therefore marched in the reference's exact pop order, and the GPU is filled with thousands of
// independent (period, source) sweeps.  Two stages per sweep:
//   * k_refine: the refined source grid (<= 129 x 129, early exit), 16 lanes per sweep on packed
//     (time, status) records with a shared-memory/global (key, node) heap; then the injection into
//     the coarse grid.  ~1.5 % of the pops.
//   * k_march_lps: the coarse grid, ONE LANE PER SWEEP (32 sweeps per warp, every sweep of a batch
//     resident), one 32-bit word per node, lazy heap back-pointers -- eik_lps.cuh.  Round 1 marched
//     the coarse grid with 16 lanes per sweep (k_eikonal3 below, DSURF_EIKONAL_V3=1): 75 % of its
//     issue slots were scalar heap work executed 16 lanes wide and half of its DRAM store sectors
//     were 4-byte back-pointer updates (profiles/r01_eikonal_v3_source_regions.md).
// All fp32 arithmetic follows the reference's operand order (library built with --fmad=false);
// sin(colatitude) factors come from host-side tables computed with the C library.
//
// Bound: dependency depth / DRAM sector rate -- the algorithmic HBM bytes of a sweep are only
// 8*(Nc + Nr); the roofline fraction is reported anyway.
#include <algorithm>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "plan.cuh"
#include "eik_lps.cuh"
#include "eik_fim.cuh"
#include <cstring>

namespace dsurf {

__device__ __forceinline__ float cube(float x) { return x * (x * x); }
__device__ __forceinline__ void bspline4(float u, float o[4]) {  // CalSurfG.f90:1510-1513
  o[0] = cube(1.0f - u) / 6.0f;
  o[1] = (4.0f - 6.0f * (u * u) + 3.0f * cube(u)) / 6.0f;
  o[2] = (1.0f + 3.0f * u + 3.0f * (u * u) - 3.0f * cube(u)) / 6.0f;
  o[3] = cube(u) / 6.0f;
}

// ---------------------------------------------------------------- K2: gridder
__global__ void k_velv(const double *__restrict__ pv, float *__restrict__ velv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) velv[i] = (float)pv[i];  // velv(i,j) = real(pv(i*(nvx+2)+j+1)), :1492
}

__global__ void k_dice(const float *__restrict__ velv, float *__restrict__ veln, int nvx, int nvz,
                       int nnx, int nnz, int kGd) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nnx * nnz) return;
  const int stz = gid % nnz + 1, stx = gid / nnz + 1;
  int i = (stz - 1) / kGd + 1;
  if (i > nvz - 1) i = nvz - 1;
  const int l = stz - kGd * (i - 1);
  int j = (stx - 1) / kGd + 1;
  if (j > nvx - 1) j = nvx - 1;
  const int m = stx - kGd * (j - 1);
  float ui[4], vi[4];
  bspline4((float)(m - 1) / (float)kGd, ui);
  bspline4((float)(l - 1) / (float)kGd, vi);
  const int ldv = nvx + 2;
  float sumi = 0.0f;
#pragma unroll
  for (int i1 = 0; i1 < 4; i1++) {
    float sumj = 0.0f;
#pragma unroll
    for (int j1 = 0; j1 < 4; j1++) sumj = sumj + ui[j1] * velv[(i - 1 + i1) * ldv + (j - 1 + j1)];
    sumi = sumi + vi[i1] * sumj;
  }
  veln[gid] = sumi;
}

int launch_dice(cudaStream_t st, const Geom &g, const double *d_pv_map, float *d_velv, float *d_veln) {
  const int nv = g.nx * g.ny;
  k_velv<<<(nv + 255) / 256, 256, 0, st>>>(d_pv_map, d_velv, nv);
  const int nn = g.nnx * g.nnz;
  k_dice<<<(nn + 255) / 256, 256, 0, st>>>(d_velv, d_veln, g.nvx, g.nvz, g.nnx, g.nnz, g.gd);
  return DSURF_OK;
}

// ---------------------------------------------------------------- K3: eikonal
constexpr int kWarpsPerBlock = 8;

struct Grid {
  int2 *node;          // packed (ttn bits, nsts)
  const float *vel;    // velocity, same indexing
  const float *risti;  // earth*sin(gox+(ix-1)*dnx), [nnx]
  int nnx, nnz;
  float dnx, dnz, earth;
  unsigned mdiv;  // n / nnz == __umulhi(n, mdiv) >> sdiv for 0 <= n < 2^31 (nnz = 8k+1 is never a power of two)
  int sdiv;
  __device__ __forceinline__ void set_div() {
    sdiv = 31 - __clz(nnz);
    mdiv = (unsigned)((1ull << (32 + sdiv)) / (unsigned)nnz) + 1u;
  }
};

// one quadrant of fouds2 (:664-756): returns true and the trial time if a solution exists
__device__ __forceinline__ bool quadrant(float Tj, float Tj2, int Sj, int Sj2, float Tk, float Tk2,
                                         int Sk, int Sk2, float slown, float ri, float risti,
                                         float dnx, float dnz, float &trav) {
  // swj/swk: second-order leg usable (:620-663)
  const bool swj = (Sj == 0) && (Sj2 == 0) && (Tj > Tj2);
  const bool swk = (Sk == 0) && (Sk2 == 0) && (Tk > Tk2);
  float a, b, c, u, v, em, tref, tdiv;
  if (swj) {
    if (swk) {
      u = 2.0f * ri * dnx;
      v = 2.0f * risti * dnz;
      em = 4.0f * Tj - Tj2 - 4.0f * Tk;
      em = em + Tk2;
      a = v * v + u * u;
      b = 2.0f * em * (u * u);
      c = (u * u) * (em * em - (slown * slown) * (v * v));
      tref = 4.0f * Tj - Tj2;
      tdiv = 3.0f;
    } else if (Sk == 0) {
      u = risti * dnz;
      v = 2.0f * ri * dnx;
      em = 3.0f * Tk - 4.0f * Tj + Tj2;
      a = v * v + 9.0f * (u * u);
      b = 6.0f * em * (u * u);
      c = (u * u) * (em * em - (slown * slown) * (v * v));
      tref = Tk;
      tdiv = 1.0f;
    } else {
      u = 2.0f * ri * dnx;
      a = 1.0f;
      b = 0.0f;
      c = -((u * u) * (slown * slown));
      tref = 4.0f * Tj - Tj2;
      tdiv = 3.0f;
    }
  } else if (Sj == 0) {
    if (swk) {
      u = ri * dnx;
      v = 2.0f * risti * dnz;
      em = 3.0f * Tj - 4.0f * Tk + Tk2;
      a = v * v + 9.0f * (u * u);
      b = 6.0f * em * (u * u);
      c = (u * u) * (em * em - (v * v) * (slown * slown));
      tref = Tj;
      tdiv = 1.0f;
    } else if (Sk == 0) {
      u = ri * dnx;
      v = risti * dnz;
      em = Tk - Tj;
      a = u * u + v * v;
      b = -(2.0f * (u * u) * em);
      c = (u * u) * (em * em - (v * v) * (slown * slown));
      tref = Tj;
      tdiv = 1.0f;
    } else {
      a = 1.0f;
      b = 0.0f;
      c = -((slown * slown) * (ri * ri) * (dnx * dnx));
      tref = Tj;
      tdiv = 1.0f;
    }
  } else {
    if (swk) {
      u = 2.0f * risti * dnz;
      a = 1.0f;
      b = 0.0f;
      c = -((u * u) * (slown * slown));
      tref = 4.0f * Tk - Tk2;
      tdiv = 3.0f;
    } else if (Sk == 0) {
      a = 1.0f;
      b = 0.0f;
      c = -((slown * slown) * (risti * risti) * (dnz * dnz));
      tref = Tk;
      tdiv = 1.0f;
    } else {
      return false;
    }
  }
  float rd1 = b * b - 4.0f * a * c;
  if (rd1 < 0.0f) rd1 = 0.0f;
  const float tdsh = (-b + sqrtf(rd1)) / (2.0f * a);
  trav = (tref + tdsh) / tdiv;
  return true;
}

constexpr int kOut = -9;  // status sentinel for "outside the grid"

template <int HS>
struct Heap2 {
  int2 *sm;  // entries [1, HS)
  int2 *gm;  // entries [HS, hcap], indexed by absolute position
  __device__ __forceinline__ int2 get(int p) const { return p < HS ? sm[p] : gm[p]; }
  __device__ __forceinline__ void set(int p, int2 e) const {
    if (p < HS)
      sm[p] = e;
    else
      gm[p] = e;
  }
};
__device__ __forceinline__ float keyf(int2 e) { return __int_as_float(e.x); }


// sift-up for the (few) initial inserts of a pass, on the v2 heap layout
template <int HS>
__device__ __forceinline__ void sift_up2(const Heap2<HS> &H, const Grid &G, int tpc, float key, int xn) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const int2 pe = H.get(tpp);
    if (key < keyf(pe)) {
      H.set(tpc, pe);
      G.node[pe.y].y = tpc;
      tpc = tpp;
      tpp = tpc >> 1;
    } else {
      tpp = 0;
    }
  }
  H.set(tpc, make_int2(__float_as_int(key), xn));
  G.node[xn].y = tpc;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_fill_nodes(int2 *node, long long n) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) node[i] = make_int2(0, -1);
}


// =============================================================================================
// v3: kG lanes per sweep, 32/kG sweeps per warp.  The march is serial per sweep, so with one warp
// per sweep (v1/v2) 31 of 32 lanes replay the same scalar work (ncu: ~780 warp instructions per
// accepted node, issue-bound at full occupancy).  Here eight lanes own a sweep: the scalar heap
// work is still executed redundantly, but only 8-wide, and four sweeps share every issued
// instruction.  Lane roles inside a group (gl = lane % 8):
//   * stencil: lane gl loads stencil node gl of each of the four neighbours (4 independent loads);
//     they are exchanged through a 256-byte shared scratch so that lane gl ends up with the eight
//     stencil nodes of neighbour gl/2 and solves two of its four quadrants;
//   * sift-down below the shared-memory heap levels: three levels (14 entries) are fetched by the
//     group in one round trip into the scratch and walked from there;
//   * sift-up: lane gl fetches ancestor (gl&1) of neighbour gl/2's heap slot in one round trip.
// Pop order, arithmetic and tie behaviour are those of march<> / march2<> (bit-identical output).
// =============================================================================================
constexpr int kScr = 32;  // int2 scratch entries per sweep
// kG lanes per sweep (8 or 16); HS shared-memory heap entries per sweep
template <int kG> struct V3 {
  static constexpr int NG = 32 / kG;          // sweeps per warp
  static constexpr int LPX = kG / 4;          // lanes per neighbour
  static constexpr int NQ = 16 / kG;          // quadrants per lane
  static constexpr int HS = (kG == 16) ? 384 : 256;
  static constexpr int MINB = (kG == 16) ? 4 : 3;  // resident blocks per SM aimed at
};

// Position of node `nid` inside the shared-memory part of the heap (entries [1, lim]); the group's
// lanes scan two entries per 16-byte load.  Used by the LAZY back-pointer scheme below.
template <int kG>
__device__ __forceinline__ int find_in_smem(const int2 *sm, int lim, int nid, int gl, unsigned gm) {
  int found = 0;
  for (int p0 = 2 * gl; p0 <= lim; p0 += 2 * kG) {
    const int4 e = *reinterpret_cast<const int4 *>(sm + p0);  // entries p0 (x,y) and p0+1 (z,w)
    if (e.y == nid && p0 >= 1) found = p0;
    if (e.w == nid && p0 + 1 <= lim) found = p0 + 1;
  }
#pragma unroll
  for (int o = kG / 2; o > 0; o >>= 1) found = max(found, __shfl_xor_sync(gm, found, o));
  return found;
}

// LAZY back-pointers (optional, DSURF_EIKONAL_LAZY=1): a node's status word holds its heap slot exactly only while the slot
// lies in the global part of the heap (slot >= HS).  Moves between two shared-memory slots -- every
// level of a sift-down above the slab, most sift-ups -- do not touch the node array at all; a stored
// value in [1, HS) therefore only says "somewhere in the shared-memory part", and the slot is found
// by scanning those <= HS-1 entries when an update of such a node needs it.  ncu (profiles/
// r01_launches_v3_summary.md): the eager scheme's scattered 4-byte status stores are half of the
// kernel's store sectors, each a read-modify-write of a 32-byte DRAM sector at full occupancy.
// Pop order and arithmetic are unchanged (heap contents are identical; only the inverse map is lazy).
// Result on B200 (gpurun_out/s8_eik_*.json, cfg 3 type-block steps): 882 sweeps/s lazy vs 939 eager -- the
// DRAM sectors saved do not pay for the slot searches because the kernel is issue-bound, not DRAM-bound.
template <bool REFINED, int kG, bool LAZY>
__device__ int march3(const Grid &G, const Heap2<V3<kG>::HS> &H, int2 *scr, int ntr, int hcap, int gl, unsigned gm,
                      int gbase, unsigned wmask, int vnl, int vnr, int vnt, int vnb) {
  const int nnx = G.nnx, nnz = G.nnz;
  int2 lastE = make_int2(0, 0);
  bool lastOK = false;
  constexpr int kHS3 = V3<kG>::HS, LPX = V3<kG>::LPX, NQ = V3<kG>::NQ;
  const int Xown = gl / LPX;  // neighbour whose quadrants this lane solves
  const int sub = gl % LPX;   // index of this lane among the lanes of its neighbour
  bool active = ntr > 0;
  int err = 0;
  // The four sweeps of a warp advance in lock step: without the warp-wide barrier at the top of
  // every acceptance the groups drift apart and the warp ends up replaying each group's
  // instruction stream separately (measured: no gain over one sweep per warp).
  for (;;) {
    __syncwarp(wmask);
    if (!__any_sync(wmask, active)) break;
    if (!active) continue;
    do {
    const int root = H.sm[1].y;
    const int ixm = (int)(__umulhi((unsigned)root, G.mdiv) >> G.sdiv);  // root / nnz (exact, see Grid::set_div)
    const int ix = ixm + 1, iz = root - ixm * nnz + 1;
    if (REFINED) {
      int swrg = 0;
      if (ix == 1 && vnl != 1) swrg = 1;
      if (ix == nnx && vnr != nnx) swrg = 1;  // sic (:399-401)
      if (iz == 1 && vnt != 1) swrg = 1;
      if (iz == nnz && vnb != nnz) swrg = 1;
      if (swrg) {
        G.node[root].y = 0;
        active = false;
        break;
      }
    }
    G.node[root].y = 0;
    // ---- ids of the four neighbours (group-uniform)
    const int xid0 = (ix - 1 >= 1) ? (ix - 2) * nnz + (iz - 1) : -1;
    const int xid1 = (ix + 1 <= nnx) ? ix * nnz + (iz - 1) : -1;
    const int xid2 = (iz - 1 >= 1) ? (ix - 1) * nnz + (iz - 2) : -1;
    const int xid3 = (iz + 1 <= nnz) ? (ix - 1) * nnz + iz : -1;
    // ---- (1) loads: own neighbour record + stencil node gl of each neighbour
    const int oxx = ix + (Xown == 0 ? -1 : Xown == 1 ? 1 : 0);
    const int oxz = iz + (Xown == 2 ? -1 : Xown == 3 ? 1 : 0);
    const int oidx = (Xown == 0) ? xid0 : (Xown == 1) ? xid1 : (Xown == 2) ? xid2 : xid3;
    int2 xn = make_int2(0, 0);
    float velx = 1.0f, risti = 0.0f;
    if (oidx >= 0) {
      xn = G.node[oidx];
      velx = G.vel[oidx];
      risti = G.risti[oxx - 1];
    }
    (void)oxz;
    int2 sn[32 / kG];
    {
      const int qn = gl & 7;  // stencil node handled by this lane
      const int off = (qn & 1) ? 2 : 1;
      const int sgn = (qn & 2) ? 1 : -1;
      const int ddx = (qn < 4) ? sgn * off : 0;
      const int ddz = (qn < 4) ? 0 : sgn * off;
#pragma unroll
      for (int j = 0; j < 32 / kG; j++) {
        const int g = ((gl + kG * j) >> 3);  // neighbour of item t = gl + kG*j
        const int xx = ix + (g == 0 ? -1 : g == 1 ? 1 : 0);
        const int xz = iz + (g == 2 ? -1 : g == 3 ? 1 : 0);
        const bool xin = (xx >= 1 && xx <= nnx && xz >= 1 && xz <= nnz);
        const int sx = xx + ddx, sz = xz + ddz;
        sn[j] = make_int2(0, kOut);
        if (xin && sx >= 1 && sx <= nnx && sz >= 1 && sz <= nnz) sn[j] = G.node[(sx - 1) * nnz + (sz - 1)];
      }
    }
    // heap slot of this lane's own neighbour if the pop moves it (node ids are never -1)
    int xmown = -1;
    // LAZY: compare every moved entry with this lane's neighbour.  Eager (exact slots in xn.y): the entries a
    // sift-down moves are exactly the ancestors-or-self of the slot where the last element lands, each going to
    // its parent -- decided once after the sift-down (see below) instead of once per level.
#define TRACK_MOVE(nid, newpos) \
  {                             \
    if (LAZY && (nid) == oidx) xmown = (newpos); \
  }
    // ---- (2) downtree (:816-885)
    if (ntr == 1) {
      ntr = 0;
      lastOK = false;
    } else {
      const int2 m = lastOK ? lastE : H.get(ntr);
      const float mk = keyf(m);
      ntr = ntr - 1;
      int tpp = 1, tpc = 2;
      bool stop = false;
      while (!stop && tpc < ntr && tpc + 1 < kHS3) {
        const int4 e12 = *reinterpret_cast<const int4 *>(H.sm + tpc);  // both children in one 16-byte load (tpc is even)
        const int2 e1 = make_int2(e12.x, e12.y), e2 = make_int2(e12.z, e12.w);
        int2 ec = e1;
        if (keyf(e1) > keyf(e2)) {
          tpc = tpc + 1;
          ec = e2;
        }
        if (keyf(ec) < mk) {
          H.sm[tpp] = ec;
          if (!LAZY) G.node[ec.y].y = tpp;
          TRACK_MOVE(ec.y, tpp);
          tpp = tpc;
          tpc = 2 * tpp;
        } else {
          stop = true;
        }
      }
      while (!stop && tpc <= ntr) {
        if (tpc + 1 < kHS3) {  // single child inside shared memory (tpc == ntr)
          const int2 e1 = H.sm[tpc];
          if (keyf(e1) < mk) {
            H.sm[tpp] = e1;
            if (!LAZY) G.node[e1.y].y = tpp;
            TRACK_MOVE(e1.y, tpp);
            tpp = tpc;
          }
          stop = true;
          break;
        }
        // three levels below tpp -> scratch[(2^r - 2) + o], r = 1..3
#pragma unroll
        for (int t0 = 0; t0 < 16; t0 += kG) {
          const int t = t0 + gl;
          if (t < 14) {
            const int r = (t < 2) ? 1 : (t < 6) ? 2 : 3;
            const int o = t - ((1 << r) - 2);
            const long long pos = ((long long)tpp << r) + o;
            int2 e = make_int2(0x7f800000, -1);
            if (pos <= ntr) e = H.get((int)pos);
            scr[t] = e;
          }
        }
        __syncwarp(gm);
        const int base = tpp;
        int rel = 0;
        for (int lvl = 1; lvl <= 3; lvl++) {
          const long long c0 = ((long long)base << lvl) + 2 * rel;
          if (c0 > ntr) {
            stop = true;
            break;
          }
          const int Lc = (1 << lvl) - 2 + 2 * rel;
          const int2 e1 = scr[Lc], e2 = scr[Lc + 1];
          int pick = 0;
          if (c0 < ntr && keyf(e1) > keyf(e2)) pick = 1;
          const int2 ec = pick ? e2 : e1;
          if (keyf(ec) < mk) {
            H.set(tpp, ec);
            G.node[ec.y].y = tpp;
            TRACK_MOVE(ec.y, tpp);
            tpp = (int)c0 + pick;
            rel = 2 * rel + pick;
            if (c0 == ntr) {
              stop = true;
              break;
            }
          } else {
            stop = true;
            break;
          }
        }
        __syncwarp(gm);
        tpc = 2 * tpp;
      }
      H.set(tpp, m);
      if (!LAZY || tpp >= kHS3 || ntr + 1 >= kHS3) G.node[m.y].y = tpp;  // ntr + 1 = slot m came from
      TRACK_MOVE(m.y, tpp);
      if (!LAZY) {
        const int p = xn.y;  // slot of this lane's neighbour before the pop (exact; <= 0: not in the heap)
        if (p == ntr + 1) {
          xmown = tpp;       // it was the last element
        } else if (p >= 2) {
          const int d = __clz(p) - __clz(tpp);  // depth(tpp) - depth(p)
          if (d >= 0 && (tpp >> d) == p) xmown = p >> 1;
        }
      }
      lastOK = false;
    }
#undef TRACK_MOVE
    // ---- stencil exchange through the scratch: lane gl receives the 8 nodes of neighbour gl/2
#pragma unroll
    for (int j = 0; j < 32 / kG; j++) scr[gl + kG * j] = sn[j];
    __syncwarp(gm);
    int2 s8[8];
#pragma unroll
    for (int qq = 0; qq < 8; qq++) s8[qq] = scr[Xown * 8 + qq];
    __syncwarp(gm);
    // ---- quadrants: this lane solves (jside = gl & 1, kside = 0 and 1) of neighbour Xown
    const int proc = (oidx >= 0 && xn.y != 0) ? (xn.y == -1 ? 1 : 2) : 0;
    float trav = 3.0e38f;
    if (proc) {
      const float slown = 1.0f / velx;
#pragma unroll
      for (int qq = 0; qq < NQ; qq++) {
        const int r = sub * NQ + qq;  // quadrant index: jside = r >> 1, kside = r & 1
        const int js = r >> 1, ks = r & 1;
        const int2 nj = js ? s8[2] : s8[0], nj2 = js ? s8[3] : s8[1];
        const int2 nk = ks ? s8[6] : s8[4], nk2 = ks ? s8[7] : s8[5];
        if (nj.y != kOut && nk.y != kOut) {
          float tq;
          if (quadrant(__int_as_float(nj.x), __int_as_float(nj2.x), nj.y, nj2.y, __int_as_float(nk.x),
                       __int_as_float(nk2.x), nk.y, nk2.y, slown, G.earth, risti, G.dnx, G.dnz, tq))
            trav = fminf(trav, tq);
        }
      }
    }
#pragma unroll
    for (int o = 1; o < LPX; o <<= 1) trav = fminf(trav, __shfl_xor_sync(gm, trav, o));
    // ---- (3) planned heap positions + ancestor fetch
    int pr[4], xi[4], ppos[4];
    float tv[4];
    int nfar = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
      pr[g] = __shfl_sync(gm, proc, gbase + LPX * g);
      tv[g] = __shfl_sync(gm, trav, gbase + LPX * g);
      int st = __shfl_sync(gm, (xmown >= 0) ? xmown : (LAZY && xn.y < kHS3 ? -1 : xn.y), gbase + LPX * g);
      xi[g] = (g == 0) ? xid0 : (g == 1) ? xid1 : (g == 2) ? xid2 : xid3;
      ppos[g] = 0;
      if (pr[g] == 1) {
        nfar++;
        ppos[g] = ntr + nfar;
      } else if (pr[g] == 2) {
        if (LAZY && st < 0) st = find_in_smem<kG>(H.sm, min(ntr, kHS3 - 1), xi[g], gl, gm);
        ppos[g] = st;
      }
    }
    const int mypos = (Xown == 0) ? ppos[0] : (Xown == 1) ? ppos[1] : (Xown == 2) ? ppos[2] : ppos[3];
    const int myanc = mypos >> (sub + 1);
    int2 anc = make_int2(0, -1);
    if (myanc >= 1) anc = H.get(myanc);
    int2 cand = make_int2(0, -1);
    if (nfar == 0 && ntr >= 1) cand = H.get(ntr);
    if (ntr + nfar > hcap) {
      err = -1;
      active = false;
      break;
    }
    // ---- apply in the reference's order (:424-486)
    bool slow = false;
    int ls0 = -1, ls1 = -1, ls2 = -1, ls3 = -1;
    int2 le0 = make_int2(0, 0), le1 = le0, le2 = le0, le3 = le0;
    int nl = 0;
    bool appended = false;
#pragma unroll
    for (int g = 0; g < 4; g++) {
      if (!pr[g]) continue;
      const int xg = xi[g];
      const float tvg = tv[g];
      int tpc;
      if (pr[g] == 1) {
        ntr = ntr + 1;
        tpc = ntr;
      } else {
        tpc = slow ? G.node[xg].y : ppos[g];
        if (LAZY && slow && tpc < kHS3) tpc = find_in_smem<kG>(H.sm, min(ntr, kHS3 - 1), xg, gl, gm);
      }
      const bool use_pref = !slow && (tpc == ppos[g]);
      int a = 0;
      bool moved = false;
      int tpp = tpc >> 1;
      while (tpp > 0) {
        int2 pe;
        if (use_pref && a < LPX) {
          pe.x = __shfl_sync(gm, anc.x, gbase + LPX * g + a);
          pe.y = __shfl_sync(gm, anc.y, gbase + LPX * g + a);
          if (tpp == ls0) pe = le0;
          if (tpp == ls1) pe = le1;
          if (tpp == ls2) pe = le2;
        } else {
          pe = H.get(tpp);
        }
        if (tvg < keyf(pe)) {
          H.set(tpc, pe);
          if (!LAZY || tpc >= kHS3) G.node[pe.y].y = tpc;
          tpc = tpp;
          tpp = tpc >> 1;
          a++;
          moved = true;
        } else {
          tpp = 0;
        }
      }
      const int2 ne = make_int2(__float_as_int(tvg), xg);
      H.set(tpc, ne);
      G.node[xg] = make_int2(__float_as_int(tvg), tpc);  // trial time + heap slot in one 8-byte store
      if (moved) {
        slow = true;
      } else {
        if (nl == 0) { ls0 = tpc; le0 = ne; }
        if (nl == 1) { ls1 = tpc; le1 = ne; }
        if (nl == 2) { ls2 = tpc; le2 = ne; }
        if (nl == 3) { ls3 = tpc; le3 = ne; }
        nl++;
      }
      if (pr[g] == 1) {
        appended = true;
        lastE = ne;
        lastOK = !moved;
      }
    }
    if (!appended) {
      if (!slow && ntr >= 1) {
        lastE = cand;
        if (ntr == ls0) lastE = le0;
        if (ntr == ls1) lastE = le1;
        if (ntr == ls2) lastE = le2;
        if (ntr == ls3) lastE = le3;
        lastOK = true;
      } else {
        lastOK = false;
      }
    } else if (slow) {
      lastOK = false;
    }
    if (ntr == 0) active = false;
    } while (false);
  }
  return err ? -1 : ntr;
}


// ---- stage 1 of a sweep, shared by both pipelines: bsplrefine (:1562-1628), the refined-grid source
// cell (:312-375) and travel(urg=1) on the refined grid with its exit test (:386-412).
// Returns true if the heap slab was too small.
template <int kG, bool LAZY>
__device__ __forceinline__ bool refine_stage(const Geom &g, const SweepDesc &d, const float *velv, int2 *noder, float *velr,
                                             const float *ristr, const Heap2<V3<kG>::HS> &H, int2 *scr, int hcap, int gl,
                                             unsigned gm, int gbase, unsigned wmask) {
  const int nrnx = d.nrnx, nrnz = d.nrnz;
  {
    const int nrr = g.gd * kSgdl;
    const int origx = (d.vnl - 1) * kSgdl + 1, origz = (d.vnt - 1) * kSgdl + 1;
    const int ldv = g.nvx + 2;
    for (int n = gl; n < nrnx * nrnz; n += kG) {
      const int idm1 = n % nrnz + 1, idm2 = n / nrnz + 1;
      const int st1 = idm1 + origz - 1, st2 = idm2 + origx - 1;
      int i = (st1 - 1) / nrr + 1;
      if (i > g.nvz - 1) i = g.nvz - 1;
      const int k = st1 - nrr * (i - 1);
      int j = (st2 - 1) / nrr + 1;
      if (j > g.nvx - 1) j = g.nvx - 1;
      const int l = st2 - nrr * (j - 1);
      float ul[4], vk[4];
      bspline4((float)(l - 1) / (float)nrr, ul);
      bspline4((float)(k - 1) / (float)nrr, vk);
      float s[4];
#pragma unroll
      for (int i1 = 0; i1 < 4; i1++) {
        float t = 0.0f;
#pragma unroll
        for (int j1 = 0; j1 < 4; j1++) t = t + ul[j1] * velv[(i - 1 + i1) * ldv + (j - 1 + j1)];
        s[i1] = vk[i1] * t;
      }
      velr[n] = s[0] + s[1] + s[2] + s[3];
      noder[n] = make_int2(0, -1);
    }
  }
  __syncwarp(gm);
  Grid R;
  R.node = noder;
  R.vel = velr;
  R.risti = ristr;
  R.nnx = nrnx;
  R.nnz = nrnz;
  R.dnx = g.drnx;
  R.dnz = g.drnz;
  R.earth = g.earth;
  R.set_div();
  int ntr = 0;
  {
    const int isx = d.tsx, isz = d.tsz;
    float vss[2][2];
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) vss[i][j] = velr[(isx - 1 + i) * nrnz + (isz - 1 + j)];
    const float dsx = (d.scx - d.gorx) - (float)(isx - 1) * g.drnx;
    const float dsz = (d.scz - d.gorz) - (float)(isz - 1) * g.drnz;
    float vsrc = 0.0f;  // bilinear (:2328-2349)
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        const float produ = (1.0f - fabsf(((float)i * g.drnx - dsx) / g.drnx)) *
                            (1.0f - fabsf(((float)j * g.drnz - dsz) / g.drnz));
        vsrc = vsrc + vss[i][j] * produ;
      }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        const float ex = dsx - (float)i * g.drnx, ez = dsz - (float)j * g.drnz;
        const float ds = sqrtf(ex * ex + ez * ez);
        const float t0 = 2.0f * ds / (vss[i][j] + vsrc);
        const int xi = (isx - 1 + i) * nrnz + (isz - 1 + j);
        noder[xi].x = __float_as_int(t0);
        ntr = ntr + 1;
        sift_up2(H, R, ntr, t0, xi);
      }
  }
  const int rc = march3<true, kG, LAZY>(R, H, scr, ntr, hcap, gl, gm, gbase, wmask, d.vnl, d.vnr, d.vnt, d.vnb);
  return rc < 0;
}

// ---- legacy single-kernel pipeline (round 1, DSURF_EIKONAL_V3=1): kG lanes march the coarse grid too,
// on packed (time, status) 8-byte node records.  Kept as an A/B reference for the pipeline below.
template <int kG, bool LAZY>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, V3<kG>::MINB)
k_eikonal3(Geom g, SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ veln_all,
           const float *__restrict__ velv_all, const float *__restrict__ risti_c, BatchView bv) {
  extern __shared__ float smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kNG = V3<kG>::NG, kHS3 = V3<kG>::HS;
  const int gg = lane / kG, gl = lane % kG, gbase = gg * kG;
  const unsigned gm = ((kG == 32) ? 0xffffffffu : ((1u << kG) - 1u)) << gbase;
  const int slot = (blockIdx.x * kWarpsPerBlock + w) * kNG + gg;
  const unsigned wmask = __ballot_sync(kFull, slot < nsw);  // lanes of this warp that own a sweep
  if (slot >= nsw) return;
  int2 *wbase = (int2 *)smem + (size_t)w * kNG * (kHS3 + kScr);
  Heap2<kHS3> H;
  H.sm = wbase + (size_t)gg * kHS3;
  H.gm = bv.hent + (size_t)slot * bv.slab;
  int2 *scr = wbase + (size_t)kNG * kHS3 + (size_t)gg * kScr;
  SweepDesc d = sw[slot];
  const size_t Nc = (size_t)g.nnx * g.nnz;
  const float *veln = veln_all + (size_t)d.map * Nc;
  const float *velv = velv_all + (size_t)d.map * g.nx * g.ny;
  int2 *node = bv.node + (size_t)slot * Nc;
  int2 *noder = bv.noder + (size_t)slot * kRefMax * kRefMax;
  float *velr = bv.velr + (size_t)slot * kRefMax * kRefMax;
  const int nrnz = d.nrnz;
  const bool failed = refine_stage<kG, LAZY>(g, d, velv, noder, velr, bv.ristr + (size_t)slot * kRefMax, H, scr, bv.hcap,
                                             gl, gm, gbase, wmask);
  if (failed && gl == 0) sw[slot].status = DSURF_ERR_HEAP;
  __syncwarp(gm);
  // ---- map refined -> coarse (:1289-1303); coarse states were pre-filled with (0,-1)
  const int bw = d.vnr - d.vnl + 1, bh = d.vnb - d.vnt + 1;
  for (int n = gl; n < bw * bh; n += kG) {
    const int cz = n % bh, cx = n / bh;
    const int2 rn = noder[(cx * kSgdl) * nrnz + cz * kSgdl];
    int2 cn = make_int2(0, rn.y);
    if (rn.y >= 0) cn.x = rn.x;
    node[(size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz)] = cn;
  }
  __syncwarp(gm);
  // ---- narrow-band completion (:1332-1349): alive with a far neighbour -> close
  // (order-free: a node turned close is still "not far" for its neighbours)
  for (int n = gl; n < bw * bh; n += kG) {
    const int cz = n % bh, cx = n / bh;
    const int k = d.vnl + cx, l = d.vnt + cz;
    const size_t o = (size_t)(k - 1) * g.nnz + (l - 1);
    if (node[o].y == 0) {
      bool far = false;
      if (l - 1 >= 1 && node[o - 1].y == -1) far = true;
      if (l + 1 <= g.nnz && node[o + 1].y == -1) far = true;
      if (k - 1 >= 1 && node[o - g.nnz].y == -1) far = true;
      if (k + 1 <= g.nnx && node[o + g.nnz].y == -1) far = true;
      if (far) node[o].y = -100;  // the reference writes 1 in place; neighbours only test ".EQ.-1"
    }
  }
  __syncwarp(gm);
  for (int n = gl; n < bw * bh; n += kG) {
    const int cz = n % bh, cx = n / bh;
    const size_t o = (size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz);
    if (node[o].y == -100) node[o].y = 1;
  }
  __syncwarp(gm);
  // ---- travel(x, z, urg=2): rebuild the heap by scanning i=1..nnx, j=1..nnz (:341-347)
  Grid C;
  C.node = node;
  C.vel = veln;
  C.risti = risti_c;
  C.nnx = g.nnx;
  C.nnz = g.nnz;
  C.dnx = g.dnx;
  C.dnz = g.dnz;
  C.earth = g.earth;
  C.set_div();
  int ntr = 0;
  for (int cx = 0; cx < bw; cx++) {
    for (int base = 0; base < bh; base += kG) {
      const int cz = base + gl;
      int st = 0;
      float tt = 0.0f;
      if (cz < bh) {
        const int2 v = node[(size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz)];
        st = v.y;
        tt = __int_as_float(v.x);
      }
      unsigned mask = (__ballot_sync(gm, st > 0) >> gbase) & ((1u << kG) - 1u);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const float key = __shfl_sync(gm, tt, gbase + b);
        const int xi = (d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + base + b);
        ntr = ntr + 1;
        sift_up2(H, C, ntr, key, xi);
      }
    }
  }
  if (failed) ntr = 0;  // keep taking part in the warp-wide barriers of the coarse march
  const int rc = march3<false, kG, LAZY>(C, H, scr, ntr, bv.hcap, gl, gm, gbase, wmask, 0, 0, 0, 0);
  if (rc < 0 && gl == 0) sw[slot].status = DSURF_ERR_HEAP;
}

// =============================================================================================
// Round-2 pipeline: k_refine (refined grid, 16 lanes per sweep, as above) + k_march_lps (coarse
// grid, ONE LANE PER SWEEP, eik_lps.cuh).
//
// k_refine ends with the injection of CalSurfG.f90:1289-1349 evaluated on a (<= 17 x 17) scratch
// copy of the refined box: alive box nodes are written to the coarse word array as their time,
// close ones (refined close nodes and alive nodes with a far neighbour) become SEEDS, listed in
// the order travel(urg=2) inserts them (ix outer, iz inner, :341-347).
// =============================================================================================
constexpr int kBoxMax = (2 * kSgs + 1) * (2 * kSgs + 1);  // 289 coarse nodes in the refined box

template <int kG>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, V3<kG>::MINB)
k_refine(Geom g, SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ velv_all, BatchView bv, bool write_words) {
  extern __shared__ float smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kNG = V3<kG>::NG, kHS3 = V3<kG>::HS;
  const int gg = lane / kG, gl = lane % kG, gbase = gg * kG;
  const unsigned gm = ((kG == 32) ? 0xffffffffu : ((1u << kG) - 1u)) << gbase;
  const int slot = (blockIdx.x * kWarpsPerBlock + w) * kNG + gg;
  const unsigned wmask = __ballot_sync(kFull, slot < nsw);
  if (slot >= nsw) return;
  int2 *wbase = (int2 *)smem + (size_t)w * kNG * (kHS3 + kScr);
  Heap2<kHS3> H;
  H.sm = wbase + (size_t)gg * kHS3;
  H.gm = bv.hent + (size_t)slot * bv.slab;
  int2 *scr = wbase + (size_t)kNG * kHS3 + (size_t)gg * kScr;
  SweepDesc d = sw[slot];
  const size_t Nc = (size_t)g.nnx * g.nnz;
  const float *velv = velv_all + (size_t)d.map * g.nx * g.ny;
  const int wld = bv.wld;  // padded word array (eik_lps.cuh)
  unsigned *word = bv.word + (size_t)slot * bv.wslot;
  int2 *noder = bv.noder + (size_t)slot * kRefMax * kRefMax;
  float *velr = bv.velr + (size_t)slot * kRefMax * kRefMax;
  int2 *box = bv.box + (size_t)slot * kBoxMax;
  int2 *seed = bv.seed + (size_t)slot * kBoxMax;
  const int nrnz = d.nrnz;
  const bool failed = refine_stage<kG, false>(g, d, velv, noder, velr, bv.ristr + (size_t)slot * kRefMax, H, scr, bv.hcap,
                                              gl, gm, gbase, wmask);
  if (failed) {
    if (gl == 0) {
      sw[slot].status = DSURF_ERR_HEAP;
      bv.nseed[slot] = 0;
    }
    return;
  }
  __syncwarp(gm);
  const int bw = d.vnr - d.vnl + 1, bh = d.vnb - d.vnt + 1;
  for (int n = gl; n < bw * bh; n += kG) {  // (:1289-1303)
    const int cz = n % bh, cx = n / bh;
    const int2 rn = noder[(cx * kSgdl) * nrnz + cz * kSgdl];
    box[n] = make_int2(rn.y >= 0 ? rn.x : 0, rn.y);
  }
  __syncwarp(gm);
  for (int n = gl; n < bw * bh; n += kG) {  // (:1332-1349); every coarse node outside the box is far
    const int cz = n % bh, cx = n / bh;
    const int k = d.vnl + cx, l = d.vnt + cz;
    if (box[n].y == 0) {
      bool far = false;
      if (l - 1 >= 1 && (cz == 0 || box[n - 1].y == -1)) far = true;
      if (l + 1 <= g.nnz && (cz == bh - 1 || box[n + 1].y == -1)) far = true;
      if (k - 1 >= 1 && (cx == 0 || box[n - bh].y == -1)) far = true;
      if (k + 1 <= g.nnx && (cx == bw - 1 || box[n + bh].y == -1)) far = true;
      if (far) box[n].y = -100;
    }
  }
  __syncwarp(gm);
  if (!write_words) return;  // fast-iterative pipeline: k_fim_start picks the box up from here
  int ns = 0;
  for (int cx = 0; cx < bw; cx++) {
    for (int base = 0; base < bh; base += kG) {
      const int cz = base + gl;
      int2 v = make_int2(0, -1);
      if (cz < bh) v = box[cx * bh + cz];
      const int xi = (d.vnl - 1 + cx + bv.wpx) * wld + (d.vnt - 1 + cz + bv.wpz);  // padded node id
      const bool close = (v.y > 0 || v.y == -100);
      if (cz < bh && v.y == 0) word[xi] = (unsigned)v.x;
      const unsigned mask = (__ballot_sync(gm, close) >> gbase) & ((1u << kG) - 1u);
      if (close) seed[ns + __popc(mask & ((1u << gl) - 1u))] = make_int2(v.x, xi);
      ns += __popc(mask);
    }
  }
  if (gl == 0) bv.nseed[slot] = ns;
}

constexpr int kLpsHS = 128;  // heap slots [1, 128) of every sweep live in shared memory, interleaved by sweep
constexpr int kLpsNM = 2;    // fouds2 ("math") warps per group of 32 sweeps; each owns 4 / kLpsNM neighbours
constexpr int kLpsThreads = 32 * (1 + kLpsNM);
constexpr int kBarRoot = 1, kBarRes = 2;  // named barriers: accepted node published / trial times published

// Heap slots >= kLpsHS live in a per-sweep global slab, slot p at entry p: the two children of a slot are one
// aligned 16-byte pair, and the three levels below slot q are three contiguous runs (16, 32, 64 bytes).
int lps_slab_entries(int hcap) { return (hcap + 8 + 15) & ~15; }  // whole 128-byte lines: every sweep's slab stays 16-byte aligned

// ---- policy of the heap warp (lane = sweep)
struct LpsHeap {
  unsigned *w;
  int2 *sm;   // + lane; slot p at sm[p * 32]
  int2 *gm;
  int2 *rootbuf;  // + lane
  float *res;     // + lane; neighbour g at res[g * 32]
  unsigned mdiv;  // i / ld == __umulhi(i, mdiv) >> sdiv (Grid::set_div)
  int sdiv;
  static constexpr int kLg = 7;  // 2^7 = kLpsHS
  __device__ __forceinline__ int div_ld(int i) const { return (int)(__umulhi((unsigned)i, mdiv) >> sdiv); }
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(kFull, p); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ uint32_t word(int i) const { return w[i]; }
  __device__ __forceinline__ void set_word(int i, uint32_t x) const { w[i] = x; }
  __device__ __forceinline__ lps::Ent hget(int p) const {
    const int2 e = (p < kLpsHS) ? sm[p * 32] : gm[p];
    lps::Ent r;
    r.x = e.x;
    r.y = e.y;
    return r;
  }
  __device__ __forceinline__ void hget2(int p, lps::Ent &a, lps::Ent &b) const {  // p even, both slots valid
    if (p < kLpsHS) {
      const int2 e1 = sm[p * 32], e2 = sm[p * 32 + 32];
      a.x = e1.x; a.y = e1.y; b.x = e2.x; b.y = e2.y;
    } else {
      const int4 e = *reinterpret_cast<const int4 *>(gm + p);
      a.x = e.x; a.y = e.y; b.x = e.z; b.y = e.w;
    }
  }
  __device__ __forceinline__ void hset(int p, lps::Ent e) const {
    if (p < kLpsHS)
      sm[p * 32] = make_int2(e.x, e.y);
    else
      gm[p] = make_int2(e.x, e.y);
  }
  __device__ __forceinline__ void stat(int, int) const {}
  // the 14 descendants of slot q (q >= 64) at relative depths 1..3; slots > lim are not read
  __device__ __forceinline__ void hblock(int q, lps::Ent b[14], int lim) const {
    int4 v[7];
#pragma unroll
    for (int i = 0; i < 7; i++) v[i] = make_int4(0, -1, 0, -1);
    v[0] = *reinterpret_cast<const int4 *>(gm + 2 * q);
    if (4 * q <= lim) {
      v[1] = *reinterpret_cast<const int4 *>(gm + 4 * q);
      v[2] = *reinterpret_cast<const int4 *>(gm + 4 * q + 2);
    }
    if (8 * q <= lim) {
#pragma unroll
      for (int i = 0; i < 4; i++) v[3 + i] = *reinterpret_cast<const int4 *>(gm + 8 * q + 2 * i);
    }
#pragma unroll
    for (int i = 0; i < 7; i++) {
      b[2 * i].x = v[i].x; b[2 * i].y = v[i].y; b[2 * i + 1].x = v[i].z; b[2 * i + 1].y = v[i].w;
    }
  }
  // state: 1 = this sweep accepted `root`, 0 = this sweep is finished, -1 = every sweep of the group is finished
  __device__ __forceinline__ void publish(int root, uint32_t key, int state) const {
    *rootbuf = make_int2(state == 1 ? root : (state == 0 ? -1 : -2), (int)key);
    __threadfence_block();  // the node-word stores of the previous acceptance are visible to the math warps
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(kBarRoot), "r"(kLpsThreads) : "memory");
  }
  __device__ __forceinline__ void collect(float tv[4]) const {
    asm volatile("bar.sync %0, %1;" ::"r"(kBarRes), "r"(kLpsThreads) : "memory");
#pragma unroll
    for (int g = 0; g < 4; g++) tv[g] = res[g * 32];
  }
};
static_assert(kLpsHS == (1 << LpsHeap::kLg), "shared-memory heap levels");

struct LpsWords {  // policy of the math warps
  const unsigned *w;
  const float *v;   // unpadded velocity map (iz fastest)
  const float *ris;
  int ld, nnz;
  __device__ __forceinline__ uint32_t word(int i) const { return w[i]; }
  __device__ __forceinline__ float vel(int i) const {  // padded id -> unpadded index
    const int xp = i / ld;
    return __ldg(v + (size_t)(xp - lps::kPad) * nnz + (i - xp * ld - lps::kPad));
  }
  __device__ __forceinline__ float risti(int ix) const { return __ldg(ris + ix); }
};

// One block = one group of 32 sweeps: warp 0 marches the 32 heaps (lane = sweep, lps::march), warps 1..kLpsNM
// evaluate fouds2 for the neighbours of the 32 accepted nodes (lane = sweep as well) while the heaps are sifted.
__global__ void __launch_bounds__(kLpsThreads, 6)
k_march_lps(Geom g, SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ veln_all,
            const float *__restrict__ risti_c, BatchView bv) {
  extern __shared__ int4 lsm4[];
  int2 *heap = reinterpret_cast<int2 *>(lsm4);                       // [kLpsHS][32]
  int2 *rootbuf = heap + kLpsHS * 32;                                // [32]
  float *res = reinterpret_cast<float *>(rootbuf + 32);              // [4][32]
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * 32 + lane;
  const bool valid = slot < nsw;
  const int ld = g.nnz + 2 * lps::kPad;
  const size_t Nw = (size_t)(g.nnx + 2 * lps::kPad) * ld;
  lps::GridP G;
  G.nnx = g.nnx;
  G.nnz = g.nnz;
  G.dnx = g.dnx;
  G.dnz = g.dnz;
  G.earth = g.earth;
  if (wp == 0) {
    LpsHeap m;
    m.sm = heap + lane;
    m.rootbuf = rootbuf + lane;
    m.res = res + lane;
    m.sdiv = 31 - __clz(ld);  // ld = nnz + 6 = 8k + 7 is never a power of two
    m.mdiv = (unsigned)((1ull << (32 + m.sdiv)) / (unsigned)ld) + 1u;
    m.w = nullptr;
    m.gm = nullptr;
    int ntr = 0;
    if (valid) {
      m.w = bv.word + (size_t)slot * Nw;
      m.gm = bv.hent + (size_t)slot * bv.slab;
      const int ns = bv.nseed[slot];
      const int2 *seed = bv.seed + (size_t)slot * kBoxMax;
      for (int i = 0; i < ns; i++) {  // travel(urg=2)'s addtree scan (:341-347)
        const int2 s = seed[i];
        bool moved;
        ntr = ntr + 1;
        const int pos = lps::sift_up(m, ntr, __int_as_float(s.x), s.y, moved);
        m.set_word(s.y, lps::kCloseBit | (uint32_t)pos);
      }
    }
    const int rc = lps::march(G, m, ntr, bv.hcap);
    if (valid && rc < 0) sw[slot].status = DSURF_ERR_HEAP;
  } else {
    LpsWords wm;
    wm.w = valid ? bv.word + (size_t)slot * Nw : nullptr;
    wm.v = valid ? veln_all + (size_t)sw[slot].map * ((size_t)g.nnx * g.nnz) : nullptr;
    wm.ris = risti_c;
    wm.ld = ld;
    wm.nnz = g.nnz;
    constexpr int kPer = 4 / kLpsNM;
    for (;;) {
      asm volatile("bar.sync %0, %1;" ::"r"(kBarRoot), "r"(kLpsThreads) : "memory");
      const int2 rb = rootbuf[lane];
      if (rb.x == -2) break;  // uniform: the heap warp publishes -2 to every lane at once
#pragma unroll
      for (int k = 0; k < kPer; k++) {
        const int gq = (wp - 1) * kPer + k;
        float tv = 0.0f;
        if (rb.x >= 0) tv = lps::eval_neighbour(G, wm, rb.x, (uint32_t)rb.y, gq).tv;
        __syncwarp();
        res[gq * 32 + lane] = tv;
      }
      __syncwarp();
      asm volatile("bar.arrive %0, %1;" ::"r"(kBarRes), "r"(kLpsThreads) : "memory");
    }
  }
}

// =============================================================================================
// Fast-iterative pipeline (DSURF_EIKONAL=fim): k_refine (exact, as above) -> k_fim_start (exact start-up of the
// coarse pass, one thread per sweep) -> k_fim_march (block-level fast-iterative sweep: one warp per sweep relaxes its
// 32 x 32 tiles in shared memory).  Algorithm, its relation to the reference's heap march and the measured deviations: eik_fim.cuh.
// =============================================================================================

__global__ void __launch_bounds__(128)
k_fim_start(Geom g, const SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ veln_all,
            const float *__restrict__ risti_c, BatchView bv) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nsw) return;
  const SweepDesc d = sw[slot];
  fim::StartCtx C;
  C.nnx = g.nnx;
  C.nnz = g.nnz;
  C.rx0 = max(0, d.isx - 1 - fim::kRegHalf);
  C.rz0 = max(0, d.isz - 1 - fim::kRegHalf);
  C.rw = min(g.nnx - 1, d.isx + fim::kRegHalf) - C.rx0 + 1;
  C.rh = min(g.nnz - 1, d.isz + fim::kRegHalf) - C.rz0 + 1;
  C.ri = g.earth;
  C.dnx = g.dnx;
  C.dnz = g.dnz;
  C.vel = veln_all + (size_t)d.map * ((size_t)g.nnx * g.nnz);
  C.risti = risti_c;
  fim::StartMem M;
  M.w = bv.fim_rw + (size_t)slot * fim::kRegNodes;
  M.heap = reinterpret_cast<lps::Ent *>(bv.hent + (size_t)slot * bv.slab);  // the refined pass is done with its slab
  M.flag = bv.fim_flag + (size_t)slot * fim::kRegNodes;
  int2 *reg = bv.fim_reg + (size_t)slot * fim::kRegNodes;
  bv.fim_rect[slot] = make_int4(C.rx0, C.rz0, C.rw, C.rh);
  if (d.status != 0) {  // the refined pass failed: nothing to propagate
    for (int i = 0; i < C.rw * C.rh; i++) reg[i] = make_int2(0, -1);
    return;
  }
  const int2 *box = bv.box + (size_t)slot * kBoxMax;
  int ntr = 0;
  fim::startup_march(C, M, reinterpret_cast<const int *>(box), d.vnl - 1, d.vnt - 1, d.vnr - d.vnl + 1, d.vnb - d.vnt + 1, ntr,
                     C.rw * C.rh);
  for (int i = 0; i < C.rw * C.rh; i++) {
    const unsigned w = M.w[i];
    int2 r;
    if (lps::alive(w))
      r = make_int2((int)w, 0);
    else if (w == lps::kFar)
      r = make_int2(0, -1);
    else
      r = make_int2(M.heap[w & 0x7FFFFFFFu].x, 1);
    reg[i] = r;
  }
}

// ---- bulk asynchronous copies (TMA engine, 1-D form) with an mbarrier: tile rows global -> shared
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One WARP owns one sweep from start to finish: no block-level synchronisation anywhere, the relaxation order is a
// pure function of the sweep (results do not depend on scheduling), and 24 sweeps are resident per SM.
constexpr int kFimSweepsPerCta = 4;
constexpr int kFimMaxWordsPerLane = 5;  // tile-activity bitmask: up to 5 * 32 * 32 = 5120 tiles (71 x 71: 2272^2 nodes)

// ---- tile load: 36 rows of 40 words (16-byte aligned) into shared memory, far -> +inf, region nodes alive before the pass
// get the flag; risti column factors, the tile's dirty words, cleared halo marks
template <bool TMA>
__device__ __forceinline__ void fim_tile_load(fim::TileD &tl, const fim::TileCtx &C, const fim::Layout L, const unsigned *T,
                                              const float *risti_c, unsigned mydirty, int lane, unsigned long long *bar,
                                              unsigned &phase) {
  const bool touches = C.gx0 - fim::kHX < C.bx0 + C.bw && C.gx0 + fim::kT + fim::kHX > C.bx0 &&
                       C.gz0 - fim::kHZ < C.bz0 + C.bh && C.gz0 + fim::kT + fim::kHZ > C.bz0;
  const uint4 *src = reinterpret_cast<const uint4 *>(T + (size_t)C.gx0 * L.pitch + C.gz0);  // row gx0 - kHX, column gz0 - kHZ
  constexpr int kVecRow = fim::kPitch / 4;
  if (TMA) {
    // the 36 rows (160 bytes each, 16-byte aligned) go through the TMA engine; one mbarrier per tile buffer counts the bytes
    // generic-proxy accesses made so far -- the tile buffer in shared memory AND the previous tile's interior stored to the
    // field in global memory, which this load's halo may overlap -- before the async-proxy reads below
    asm volatile("fence.proxy.async;" ::: "memory");
    if (lane == 0) mbar_expect_tx(bar, fim::kRows * fim::kPitch * 4);
    __syncwarp();
    for (int r = lane; r < fim::kRows; r += 32)
      bulk_g2s(tl.t + r * fim::kPitch, src + (size_t)r * (L.pitch / 4), fim::kPitch * 4, bar);
    mbar_wait(bar, phase);
    phase ^= 1u;
  }
  for (int i = lane; i < fim::kRows * kVecRow; i += 32) {
    const int r = i / kVecRow, c4 = i - r * kVecRow;
    uint4 v = TMA ? *reinterpret_cast<const uint4 *>(tl.t + r * fim::kPitch + 4 * c4) : src[(size_t)r * (L.pitch / 4) + c4];
    v.x = (int)v.x < 0 ? fim::kInf : v.x;
    v.y = (int)v.y < 0 ? fim::kInf : v.y;
    v.z = (int)v.z < 0 ? fim::kInf : v.z;
    v.w = (int)v.w < 0 ? fim::kInf : v.w;
    if (touches) {
      const int bx = C.gx0 + r - fim::kHX - C.bx0, bz = C.gz0 + 4 * c4 - fim::kHZ - C.bz0;
      if (bx >= 0 && bx < C.bw) {
        unsigned *e = &v.x;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (bz + k >= 0 && bz + k < C.bh && C.box[2 * (bx * C.bh + bz + k) + 1] == 0) e[k] |= fim::kInit;
      }
    }
    *reinterpret_cast<uint4 *>(tl.t + r * fim::kPitch + 4 * c4) = v;
  }
  const int gx = C.gx0 + lane;
  tl.risti[lane] = gx < C.nnx ? __ldg(risti_c + gx) : 0.0f;
  tl.dirty[lane] = mydirty;
  if (lane < 4) tl.hx[lane] = 0;
  if (lane >= 4 && lane < 8) tl.hz[lane - 4] = 0;
}

// ---- tile store: the interior (coalesced rows) if anything changed, then the marks for the neighbour tiles
__device__ __forceinline__ void fim_tile_store(fim::TileD &tl, const fim::TileCtx &C, const fim::Layout L, int tile, unsigned *T,
                                               unsigned *bitmap, unsigned *active, int lane, bool changed) {
  const int tx = tile / L.ntz, tz = tile - tx * L.ntz;
  if (changed) {
    const int gz = C.gz0 + lane;
    for (int x = 0; x < fim::kT; x++) {
      const int gx = C.gx0 + x;
      const unsigned w = *tl.at(x, lane);
      if (gx < C.nnx && gz < C.nnz && (int)w >= 0) T[L.at(gx, gz)] = (w == fim::kInf) ? fim::kFarG : w;
    }
  }
  // x-neighbours: halo rows -2, -1 are rows 30, 31 of tile (tx - 1, tz); rows 32, 33 are rows 0, 1 of (tx + 1, tz)
  if (lane < 4) {
    const unsigned m = tl.hx[lane];
    const int ntx = lane < 2 ? tx - 1 : tx + 1;
    if (m && ntx >= 0 && ntx < L.ntx) {
      const int nt = ntx * L.ntz + tz;
      atomicOr(bitmap + (size_t)nt * fim::kT + (lane < 2 ? fim::kT - 2 + lane : lane - 2), m);
      atomicOr(active + (nt >> 5), 1u << (nt & 31));
    }
  }
  // z-neighbours: halo columns -2, -1 are bits 30, 31 of tile (tx, tz - 1); columns 32, 33 are bits 0, 1 of (tx, tz + 1)
  const unsigned lo = (((tl.hz[0] >> lane) & 1u) << 30) | (((tl.hz[1] >> lane) & 1u) << 31);
  const unsigned hi = ((tl.hz[2] >> lane) & 1u) | (((tl.hz[3] >> lane) & 1u) << 1);
  if (lo && tz - 1 >= 0) {
    atomicOr(bitmap + (size_t)(tile - 1) * fim::kT + lane, lo);
    atomicOr(active + ((tile - 1) >> 5), 1u << ((tile - 1) & 31));
  }
  if (hi && tz + 1 < L.ntz) {
    atomicOr(bitmap + (size_t)(tile + 1) * fim::kT + lane, hi);
    atomicOr(active + ((tile + 1) >> 5), 1u << ((tile + 1) & 31));
  }
  const unsigned left = ((volatile unsigned *)tl.dirty)[lane];  // walk limit reached: the tile stays active
  if (left) {
    atomicOr(bitmap + (size_t)tile * fim::kT + lane, left);
    atomicOr(active + (tile >> 5), 1u << (tile & 31));
  }
}

// the warp relaxes one tile: load (times + halo), anti-diagonal walks, store, marks for the neighbour tiles
template <bool TMA>
__device__ __forceinline__ void fim_process_tile(fim::TileD &tl, const fim::TileCtx &Cb, const fim::Layout L, int tile, unsigned *T,
                                                 unsigned *bitmap, unsigned *active, const float *vel, const float *risti_c,
                                                 int srcx, int srcz, int lane, unsigned long long *bar, unsigned &phase) {
  const int tx = tile / L.ntz, tz = tile - tx * L.ntz;
  unsigned *bm = bitmap + (size_t)tile * fim::kT;
  const unsigned mydirty = bm[lane];
  if (!__any_sync(kFull, mydirty != 0)) return;
  bm[lane] = 0;
  fim::TileCtx C = Cb;
  C.gx0 = tx * fim::kT;
  C.gz0 = tz * fim::kT;
  fim_tile_load<TMA>(tl, C, L, T, risti_c, mydirty, lane, bar, phase);
  __syncwarp();
  // ---- relax: anti-diagonal walks (lane = tile row), first away from the source
  const int sx0 = (C.gx0 + fim::kT / 2 >= srcx) ? 1 : -1, sz0 = (C.gz0 + fim::kT / 2 >= srcz) ? 1 : -1;
  bool changed = false;
  volatile unsigned *vdirty = tl.dirty;
  const float *velrow = vel + (size_t)min(C.gx0 + lane, C.nnx - 1) * C.nnz;
  for (int w = 0; w < 32; w++) {
    if (!__any_sync(kFull, vdirty[lane] != 0)) break;
    const int sx = (w & 1) ? -sx0 : sx0, sz = (w & 2) ? -sz0 : sz0;
    // slowness of this lane's node on the NEXT diagonal, loaded one step ahead (off the critical path).  A/B on B200
    // (profiles/r02_fim_source_variants_ab.log, r02_fim_marking_variants_ab.log): deferring the division to the use site, a
    // branch-free marking loop in relax_node, marking without bounds tests (out-of-grid words flagged at load), re-using the
    // stencil words in the marking loop and 7 resident CTAs per SM were all measured and rejected (same COO digest, 4-25 % slower)
    float slow_pf = 0.0f;
    int z_pf = -2;
    for (int dg = 0; dg < 2 * fim::kT - 1; dg++) {
      const int z = fim::diag_z(lane, dg, sx, sz);
      const bool mine = z >= 0 && ((vdirty[lane] >> z) & 1u);
      if (!__any_sync(kFull, mine)) continue;
      float sl = slow_pf;
      if (mine && z_pf != z) sl = 1.0f / __ldg(velrow + min(C.gz0 + z, C.nnz - 1));
      const int zn = fim::diag_z(lane, dg + 1, sx, sz);
      if (zn >= 0) {
        slow_pf = 1.0f / __ldg(velrow + min(C.gz0 + zn, C.nnz - 1));
        z_pf = zn;
      }
      if (mine) changed |= fim::relax_node(tl, C, lane, z, sl);
      __syncwarp();
    }
  }
  changed = __any_sync(kFull, changed);
  fim_tile_store(tl, C, L, tile, T, bitmap, active, lane, changed);
  __syncwarp();
}

template <int MINB, bool TMA>
__global__ void __launch_bounds__(kFimSweepsPerCta * 32, MINB)
k_fim_march(Geom g, SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ veln_all, const float *__restrict__ risti_c,
            BatchView bv, fim::Layout L) {
  extern __shared__ uint4 fsm4[];
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * kFimSweepsPerCta + wp;
  if (slot >= nsw) return;
  const int ntiles = L.ntx * L.ntz;
  const int nwords = (ntiles + 31) >> 5, nwpad = (nwords + 3) & ~3;
  fim::TileD *tiles = reinterpret_cast<fim::TileD *>(fsm4);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(tiles + kFimSweepsPerCta);
  unsigned *active = reinterpret_cast<unsigned *>(bars + kFimSweepsPerCta) + wp * nwpad;  // bit per tile: relax it next round
  fim::TileD &tl = tiles[wp];
  unsigned long long *bar = bars + wp;
  unsigned phase = 0;
  if (TMA) {
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
  }
  const SweepDesc d = sw[slot];
  unsigned *T = bv.word + (size_t)slot * bv.wslot;
  unsigned *bitmap = bv.fim_bitmap + (size_t)slot * ((size_t)ntiles * fim::kT);
  const int4 rect = bv.fim_rect[slot];
  const int *reg = reinterpret_cast<const int *>(bv.fim_reg + (size_t)slot * fim::kRegNodes);
  const float *vel = veln_all + (size_t)d.map * ((size_t)g.nnx * g.nnz);
  fim::TileCtx C;
  C.gx0 = C.gz0 = 0;
  C.nnx = g.nnx;
  C.nnz = g.nnz;
  C.ri = g.earth;
  C.dnx = g.dnx;
  C.dnz = g.dnz;
  C.bx0 = rect.x;
  C.bz0 = rect.y;
  C.bw = rect.z;
  C.bh = rect.w;
  C.box = reg;
  // ---- prologue: region times into the field, dirty marks for the seeds and their not-yet-alive neighbours
  for (int i = lane; i < nwpad; i += 32) active[i] = 0;
  for (int i = lane; i < C.bw * C.bh; i += 32) {
    const int st = reg[2 * i + 1];
    if (st >= 0) {
      const int bx = i / C.bh, bz = i - bx * C.bh;
      T[L.at(C.bx0 + bx, C.bz0 + bz)] = (unsigned)reg[2 * i];
    }
  }
  __syncwarp();
  for (int i = lane; i < C.bw * C.bh; i += 32) {
    if (reg[2 * i + 1] <= 0) continue;
    const int bx = i / C.bh, bz = i - bx * C.bh;
#pragma unroll
    for (int q = 0; q < 5; q++) {
      const int ex = bx + (q == 1 ? -1 : q == 2 ? 1 : 0), ez = bz + (q == 3 ? -1 : q == 4 ? 1 : 0);
      const int gx = C.bx0 + ex, gz = C.bz0 + ez;
      if (gx < 0 || gx >= C.nnx || gz < 0 || gz >= C.nnz) continue;
      if (ex >= 0 && ex < C.bw && ez >= 0 && ez < C.bh && reg[2 * (ex * C.bh + ez) + 1] == 0) continue;
      const int tx = gx / fim::kT, tz = gz / fim::kT, nt = tx * L.ntz + tz;
      atomicOr(bitmap + (size_t)nt * fim::kT + (gx - tx * fim::kT), 1u << (gz - tz * fim::kT));
      atomicOr(active + (nt >> 5), 1u << (nt & 31));
    }
  }
  __syncwarp();
  // ---- rounds: every tile that was active when the round started, in index order
  const int guard_max = 64 * (L.ntx + L.ntz) + 1024;
  for (int round = 0;; round++) {
    unsigned snap[kFimMaxWordsPerLane];
    bool any = false;
#pragma unroll
    for (int j = 0; j < kFimMaxWordsPerLane; j++) {
      const int wi = j * 32 + lane;
      snap[j] = 0;
      if (wi < nwords) {
        snap[j] = active[wi];
        active[wi] = 0;
      }
      any |= snap[j] != 0;
    }
    __syncwarp();
    if (!__any_sync(kFull, any)) break;
    if (round > guard_max) {  // no fixed point (cannot happen for a causal rule): report instead of hanging
      if (lane == 0) sw[slot].status = DSURF_ERR_HEAP;
      break;
    }
#pragma unroll
    for (int j = 0; j < kFimMaxWordsPerLane; j++) {
      unsigned lanes = __ballot_sync(kFull, snap[j] != 0);
      while (lanes) {
        const int src = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        unsigned w = __shfl_sync(kFull, snap[j], src);
        while (w) {
          const int b = __ffs(w) - 1;
          w &= w - 1;
          fim_process_tile<TMA>(tl, C, L, (j * 32 + src) * 32 + b, T, bitmap, active, vel, risti_c, d.isx - 1, d.isz - 1, lane, bar, phase);
        }
      }
    }
  }
}

// Pipeline selection: dsurf_set_eikonal_mode() or DSURF_EIKONAL = exact | lps | fim (DSURF_EIKONAL_LPS=1 is the older
// spelling of lps).  Default: the exact single-kernel march with 16 lanes per sweep (k_eikonal3).
//   lps  k_refine + k_march_lps (exact, lane per sweep, warp-specialised).  Measured on B200 at cfg 3 (profiles/
//        r02_eikonal_fim.md, last section): 33 s per stage against 22.6 s -- with one lane per sweep every load touches 32 distinct lines
//        and the slowest of 32 unrelated sweeps sets the pace of each acceptance; it needs 2.4x fewer DRAM bytes and 5x
//        fewer issue slots per accepted node, but the machine is latency-, not throughput-bound on this path.
//   fim  k_refine + k_fim_start + k_fim_march (eik_fim.cuh): not bit-exact by construction, deviations measured.
static int g_eik_mode = -1;
int eikonal_mode() {
  if (g_eik_mode < 0) {
    g_eik_mode = kEikExact16;
    const char *e = getenv("DSURF_EIKONAL");
    if (e && (!strcmp(e, "fim") || !strcmp(e, "FIM"))) g_eik_mode = kEikFim;
    else if (e && !strcmp(e, "lps")) g_eik_mode = kEikLps;
    else if (getenv("DSURF_EIKONAL_LPS")) g_eik_mode = kEikLps;
  }
  return g_eik_mode;
}
}  // namespace dsurf
extern "C" int dsurf_set_eikonal_mode(int mode) {
  if (mode < 0 || mode > 2) return DSURF_ERR_BAD_ARG;
  dsurf::g_eik_mode = mode;
  return DSURF_OK;
}
extern "C" int dsurf_get_eikonal_mode(void) { return dsurf::eikonal_mode(); }
namespace dsurf {
bool eikonal_uses_words(int mode) { return mode != kEikExact16; }

void eikonal_word_layout(const Geom &g, int mode, int *wld, int *wpx, int *wpz, size_t *wslot) {
  if (mode == kEikFim) {
    const fim::Layout L = fim::make_layout(g.nnx, g.nnz);
    *wld = L.pitch;
    *wpx = fim::kHX;
    *wpz = fim::kHZ;
    *wslot = L.words();
  } else {
    *wld = g.nnz + 2 * lps::kPad;
    *wpx = *wpz = lps::kPad;
    *wslot = (size_t)(g.nnx + 2 * lps::kPad) * (g.nnz + 2 * lps::kPad);
  }
}
size_t eikonal_fim_bitmap_words(const Geom &g) {
  const fim::Layout L = fim::make_layout(g.nnx, g.nnz);
  return (size_t)L.ntx * L.ntz * fim::kT;
}

// int2 entries of per-sweep heap slab the selected pipeline needs for capacity hcap
int eikonal_slab_entries(int hcap, int mode) {
  int n = hcap + 1;
  if (mode == kEikLps) n = std::max(n, lps_slab_entries(hcap));
  if (mode == kEikFim) n = std::max(n, fim::kRegNodes + 1);
  return n;
}

// packed-record pipeline: sweeps that can be resident at once (its batches are sized to whole waves);
// the word pipelines keep every sweep of a batch in memory and are limited by memory only
int eikonal_resident_sweeps(int mode) {
  if (mode != kEikExact16) return 0;
  const size_t sm = (size_t)kWarpsPerBlock * V3<16>::NG * (V3<16>::HS + kScr) * sizeof(int2);
  int nb = 0;
  cudaFuncSetAttribute(k_eikonal3<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eikonal3<16, false>, kWarpsPerBlock * 32, sm);
  cudaGetLastError();
  if (nb < 1) nb = 1;
  return nb * kWarpsPerBlock * V3<16>::NG * sm_count();
}

int launch_eikonal(cudaStream_t st, const Geom &g, const SweepDesc *d_sw, int nsw, const float *d_veln_all,
                   const float *d_velv_all, const float *d_risti, BatchView bv, int *launches, int mode) {
  if (nsw <= 0) return DSURF_OK;
  const long long ntot = (long long)nsw * g.nnx * g.nnz;
  SweepDesc *sw = const_cast<SweepDesc *>(d_sw);
  const int nt = kWarpsPerBlock * 32;
  const int per_block = kWarpsPerBlock * V3<16>::NG;
  const int grid3 = (nsw + per_block - 1) / per_block;
  const size_t smem = (size_t)kWarpsPerBlock * V3<16>::NG * (V3<16>::HS + kScr) * sizeof(int2);
  const fim::Layout L = fim::make_layout(g.nnx, g.nnz);
  const int ntiles = L.ntx * L.ntz;
  const size_t fsm = (size_t)kFimSweepsPerCta * (sizeof(fim::TileD) + sizeof(unsigned long long) + (size_t)((((ntiles + 31) >> 5) + 3) & ~3) * sizeof(unsigned));
  static bool attr = false;
  if (!attr) {
    DS_CUDA(cudaFuncSetAttribute(k_eikonal3<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DS_CUDA(cudaFuncSetAttribute(k_eikonal3<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DS_CUDA(cudaFuncSetAttribute(k_refine<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DS_CUDA(cudaFuncSetAttribute(k_march_lps, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DS_CUDA(cudaFuncSetAttribute(k_march_lps, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    DS_CUDA(cudaFuncSetAttribute(k_fim_march<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DS_CUDA(cudaFuncSetAttribute(k_fim_march<6, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DS_CUDA(cudaFuncSetAttribute(k_fim_march<8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DS_CUDA(cudaFuncSetAttribute(k_fim_march<6, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DS_CUDA(cudaFuncSetAttribute(k_fim_march<5, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  if (mode == kEikExact16) {
    k_fill_nodes<<<sm_count() * 8, nt, 0, st>>>(bv.node, ntot);
    static const bool lazy = getenv("DSURF_EIKONAL_LAZY") != nullptr;
    if (lazy)
      k_eikonal3<16, true><<<grid3, nt, smem, st>>>(g, sw, nsw, d_veln_all, d_velv_all, d_risti, bv);
    else
      k_eikonal3<16, false><<<grid3, nt, smem, st>>>(g, sw, nsw, d_veln_all, d_velv_all, d_risti, bv);
    DS_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    return DSURF_OK;
  }
  DS_CUDA(cudaMemsetAsync(bv.word, 0xFF, (size_t)nsw * bv.wslot * sizeof(unsigned), st));  // every node (and the frame) far
  if (mode == kEikFim) {
    if (ntiles > kFimMaxWordsPerLane * 32 * 32) {
      set_error(__FILE__, __LINE__, "fast-iterative eikonal: grid too large for the per-sweep tile bitmask (5120 tiles)");
      return DSURF_ERR_BAD_ARG;
    }
    static size_t fsm_set = 0;
    if (fsm > fsm_set) {
      DS_CUDA(cudaFuncSetAttribute(k_fim_march<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      DS_CUDA(cudaFuncSetAttribute(k_fim_march<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      DS_CUDA(cudaFuncSetAttribute(k_fim_march<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      DS_CUDA(cudaFuncSetAttribute(k_fim_march<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      DS_CUDA(cudaFuncSetAttribute(k_fim_march<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      fsm_set = fsm;
    }
    // resident CTAs per SM the kernel is compiled for (4 sweeps each): 6 = 80 registers per thread (default), 4 = 128, 8 = 64
    static const int minb = getenv("DSURF_FIM_MINB") ? atoi(getenv("DSURF_FIM_MINB")) : 6;
    DS_CUDA(cudaMemsetAsync(bv.fim_bitmap, 0, (size_t)nsw * ntiles * fim::kT * sizeof(unsigned), st));
    k_refine<16><<<grid3, nt, smem, st>>>(g, sw, nsw, d_velv_all, bv, false);
    k_fim_start<<<(nsw + 127) / 128, 128, 0, st>>>(g, sw, nsw, d_veln_all, d_risti, bv);
    const int fgrid = (nsw + kFimSweepsPerCta - 1) / kFimSweepsPerCta, fnt = kFimSweepsPerCta * 32;
    // tile rows through the TMA engine (cp.async.bulk + mbarrier) or with plain 16-byte loads (DSURF_FIM_TMA=0)
    static const bool tma = !(getenv("DSURF_FIM_TMA") && atoi(getenv("DSURF_FIM_TMA")) == 0);
    if (minb == 4)
      k_fim_march<4, false><<<fgrid, fnt, fsm, st>>>(g, sw, nsw, d_veln_all, d_risti, bv, L);
    else if (minb == 8)
      k_fim_march<8, false><<<fgrid, fnt, fsm, st>>>(g, sw, nsw, d_veln_all, d_risti, bv, L);
    else if (minb == 5)
      k_fim_march<5, true><<<fgrid, fnt, fsm, st>>>(g, sw, nsw, d_veln_all, d_risti, bv, L);
    else if (tma)
      k_fim_march<6, true><<<fgrid, fnt, fsm, st>>>(g, sw, nsw, d_veln_all, d_risti, bv, L);
    else
      k_fim_march<6, false><<<fgrid, fnt, fsm, st>>>(g, sw, nsw, d_veln_all, d_risti, bv, L);
    DS_CUDA(cudaGetLastError());
    if (launches) *launches += 3;
    return DSURF_OK;
  }
  k_refine<16><<<grid3, nt, smem, st>>>(g, sw, nsw, d_velv_all, bv, true);
  DS_CUDA(cudaGetLastError());
  const size_t lsm = (size_t)kLpsHS * 32 * sizeof(int2) + 32 * sizeof(int2) + 4 * 32 * sizeof(float);
  k_march_lps<<<(nsw + 31) / 32, kLpsThreads, lsm, st>>>(g, sw, nsw, d_veln_all, d_risti, bv);
  DS_CUDA(cudaGetLastError());
  if (launches) *launches += 2;
  return DSURF_OK;
}

}  // namespace dsurf
