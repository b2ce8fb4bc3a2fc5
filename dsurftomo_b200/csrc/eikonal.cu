// K2/K3 -- B-spline dicing and the 2-D eikonal solve on the spherical-shell grid, replacing
// gridder (src/CalSurfG.f90:1460-1553), bsplrefine (:1562-1628), travel/fouds2/addtree/downtree/
// updtree (:288-921) and the source-grid refinement orchestration of CalSurfG (:1193-1355).
//
// Design (round 1): EXACT-ORDER REPLAY.  The reference's fast-marching result depends on the
// acceptance order (fouds2 overwrites a trial value instead of taking a min, the mixed-order
// stencil is upgraded as neighbours become alive, updtree only sifts up -- SURVEY.md section 7,
// hard part 1), so a tile-parallel fast-iterative sweep converges to a slightly different
// field.  To guarantee bit-identical travel times (and therefore bit-identical ray cells) each
// sweep is marched by ONE WARP in the reference's exact pop order, and the GPU is filled with
// thousands of independent (period, source) sweeps:
//   * the narrow-band binary heap keeps (key, node) pairs; its first kHeapSm entries live in
//     shared memory, deeper levels in a per-sweep global slab (L1/L2 resident);
//   * when a node is accepted, the four neighbours' mixed-order updates are evaluated
//     concurrently: 8 lanes per neighbour fetch its 8-point stencil (one packed
//     (time,status) 8-byte load each), 4 lanes per neighbour solve one quadrant each and a
//     2-step warp-shuffle min-reduction yields the trial time; heap inserts/updates are then
//     applied in the reference's order (x-1, x+1, z-1, z+1);
//   * heap/state writes are performed redundantly by all lanes (same address, same value), so
//     every lane observes its own program-order writes and no intra-warp fence is needed on
//     the serial path.
// All fp32 arithmetic follows the reference's operand order (library built with --fmad=false);
// sin(colatitude) factors come from host-side tables computed with the C library.
//
// Bound: dependency depth / L2 latency -- the algorithmic HBM bytes of a sweep are only
// 8*(Nc + Nr); the roofline fraction is reported anyway.
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "plan.cuh"

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
constexpr int kHeapSm = 1024;     // heap entries [1, kHeapSm) kept in shared memory per warp
constexpr int kWarpsPerBlock = 8;

struct Heap {
  float *sk;
  int *sn;
  float *gk;
  int *gn;
  __device__ __forceinline__ float key(int p) const { return p < kHeapSm ? sk[p] : gk[p]; }
  __device__ __forceinline__ int nod(int p) const { return p < kHeapSm ? sn[p] : gn[p]; }
  __device__ __forceinline__ void set(int p, float k, int n) const {
    if (p < kHeapSm) {
      sk[p] = k;
      sn[p] = n;
    } else {
      gk[p] = k;
      gn[p] = n;
    }
  }
};

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

// sift-up from position tpc with (key,node): addtree (:768-805) / updtree (:894-921)
__device__ __forceinline__ void sift_up(const Heap &H, const Grid &G, int tpc, float key, int xn) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const float pk = H.key(tpp);
    if (key < pk) {
      const int pn = H.nod(tpp);
      H.set(tpc, pk, pn);
      G.node[pn].y = tpc;
      tpc = tpp;
      tpp = tpc >> 1;
    } else {
      tpp = 0;
    }
  }
  H.set(tpc, key, xn);
  G.node[xn].y = tpc;
}

// downtree (:816-885); the root has already been marked alive by the caller
__device__ __forceinline__ void pop_root(const Heap &H, const Grid &G, int &ntr) {
  if (ntr == 1) {
    ntr = 0;
    return;
  }
  const float mk = H.key(ntr);
  const int mn = H.nod(ntr);
  ntr = ntr - 1;
  int tpp = 1, tpc = 2;
  while (tpc < ntr) {
    const float k1 = H.key(tpc), k2 = H.key(tpc + 1);
    float kc = k1;
    if (k1 > k2) {
      tpc = tpc + 1;
      kc = k2;
    }
    if (kc < mk) {
      const int cn = H.nod(tpc);
      H.set(tpp, kc, cn);
      G.node[cn].y = tpp;
      tpp = tpc;
      tpc = 2 * tpp;
    } else {
      tpc = ntr + 1;
    }
  }
  if (tpc == ntr) {
    const float kc = H.key(tpc);
    if (kc < mk) {
      const int cn = H.nod(tpc);
      H.set(tpp, kc, cn);
      G.node[cn].y = tpp;
      tpp = tpc;
    }
  }
  H.set(tpp, mk, mn);
  G.node[mn].y = tpp;
}

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

// the march loop of travel (:386-486).  REFINED adds the refined-grid exit test (:392-412).
template <bool REFINED>
__device__ int march(const Grid &G, const Heap &H, int ntr, int hcap, int lane, int vnl, int vnr,
                     int vnt, int vnb) {
  const int nnx = G.nnx, nnz = G.nnz;
  const int grp = lane >> 3, q = lane & 7;
  while (ntr > 0) {
    const int root = H.nod(1);
    const int ix = root / nnz + 1, iz = root - (ix - 1) * nnz + 1;
    if (REFINED) {
      int swrg = 0;
      if (ix == 1 && vnl != 1) swrg = 1;
      if (ix == nnx && vnr != nnx) swrg = 1;  // sic: the reference compares with the REFINED nnx
      if (iz == 1 && vnt != 1) swrg = 1;
      if (iz == nnz && vnb != nnz) swrg = 1;
      if (swrg) {
        G.node[root].y = 0;
        break;
      }
    }
    G.node[root].y = 0;
    pop_root(H, G, ntr);
    // ---- neighbour X of this lane group and the stencil node of this lane
    int xx = ix, xz = iz;
    if (grp == 0) xx = ix - 1;
    if (grp == 1) xx = ix + 1;
    if (grp == 2) xz = iz - 1;
    if (grp == 3) xz = iz + 1;
    const bool xin = (xx >= 1 && xx <= nnx && xz >= 1 && xz <= nnz);
    const int xidx = (xx - 1) * nnz + (xz - 1);
    int sx = xx, sz = xz;
    {
      const int off = (q & 1) ? 2 : 1;
      const int sgn = (q & 2) ? 1 : -1;
      if (q < 4)
        sx = xx + sgn * off;
      else
        sz = xz + sgn * off;
    }
    const bool sin_ = xin && (sx >= 1 && sx <= nnx && sz >= 1 && sz <= nnz);
    int2 sn = make_int2(0, kOut);
    if (sin_) sn = G.node[(sx - 1) * nnz + (sz - 1)];
    int2 xn = make_int2(0, 0);
    float slown = 0.0f, risti = 0.0f;
    if (xin) {
      xn = G.node[xidx];
      slown = 1.0f / G.vel[xidx];
      risti = G.risti[xx - 1];
    }
    const int proc = (xin && xn.y != 0) ? (xn.y == -1 ? 1 : 2) : 0;
    // ---- quadrant lanes: r = q (0..3): jside = r>>1, kside = r&1
    const int gb = lane & ~7;
    const int r = q & 3;
    const int lj = gb + 2 * (r >> 1), lk = gb + 4 + 2 * (r & 1);
    const float Tj = __int_as_float(__shfl_sync(kFull, sn.x, lj));
    const float Tj2 = __int_as_float(__shfl_sync(kFull, sn.x, lj + 1));
    const int Sj = __shfl_sync(kFull, sn.y, lj);
    const int Sj2 = __shfl_sync(kFull, sn.y, lj + 1);
    const float Tk = __int_as_float(__shfl_sync(kFull, sn.x, lk));
    const float Tk2 = __int_as_float(__shfl_sync(kFull, sn.x, lk + 1));
    const int Sk = __shfl_sync(kFull, sn.y, lk);
    const int Sk2 = __shfl_sync(kFull, sn.y, lk + 1);
    float trav = 3.0e38f;
    if (proc && q < 4 && Sj != kOut && Sk != kOut) {
      float tq;
      if (quadrant(Tj, Tj2, Sj, Sj2, Tk, Tk2, Sk, Sk2, slown, G.earth, risti, G.dnx, G.dnz, tq)) trav = tq;
    }
    trav = fminf(trav, __shfl_xor_sync(kFull, trav, 1));
    trav = fminf(trav, __shfl_xor_sync(kFull, trav, 2));
    // ---- apply in the reference's order: x-1, x+1, z-1, z+1
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int pr = __shfl_sync(kFull, proc, 8 * g);
      if (!pr) continue;
      const float tv = __shfl_sync(kFull, trav, 8 * g);
      const int xi = __shfl_sync(kFull, xidx, 8 * g);
      G.node[xi].x = __float_as_int(tv);
      if (pr == 1) {
        ntr = ntr + 1;
        if (ntr > hcap) return -1;
        sift_up(H, G, ntr, tv, xi);
      } else {
        const int pos = G.node[xi].y;  // re-read: earlier sifts may have moved it
        sift_up(H, G, pos, tv, xi);
      }
    }
  }
  return ntr;
}


// =============================================================================================
// v2 march: same pop order and arithmetic as march<> above, restructured so that the dependent
// memory round trips of one acceptance overlap:
//   (1) the stencil loads of the four neighbours are issued BEFORE the root is popped (they only
//       depend on the root's coordinates; the pop changes heap positions, never times or the
//       alive/close/far category);
//   (2) heap entries are single 8-byte (key,node) words; when the sift-down leaves the
//       shared-memory levels, the next four levels below the current slot (2+4+8+16 entries) are
//       fetched by 30 lanes in ONE round trip and the walk continues on registers via shuffles;
//   (3) the ancestors of the (up to four) insert/update positions are fetched by 8 lanes each in
//       one round trip; a three-entry write log keeps earlier inserts of the same acceptance
//       coherent, and any sift that actually moves entries drops the rest of the acceptance to
//       plain loads (rare: new trial times are almost always larger than their parents');
//   (4) the moving element of the next pop (the last heap entry) is kept in registers whenever it
//       is known (it is the entry appended last), or fetched together with (3).
// Heap positions of neighbours moved by the pop itself are tracked in registers, so no status
// re-read is needed.  Results are bit-identical to march<> (tests/test_gpu_parity.py runs both).
// =============================================================================================
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

template <bool REFINED, int HS>
__device__ int march2(const Grid &G, const Heap2<HS> &H, int ntr, int hcap, int lane, int vnl, int vnr, int vnt,
                      int vnb) {
  const int nnx = G.nnx, nnz = G.nnz;
  const int grp = lane >> 3, q = lane & 7;
  int2 lastE = make_int2(0, 0);
  bool lastOK = false;
  while (ntr > 0) {
    const int root = H.sm[1].y;
    const int ix = root / nnz + 1, iz = root - (ix - 1) * nnz + 1;
    if (REFINED) {
      int swrg = 0;
      if (ix == 1 && vnl != 1) swrg = 1;
      if (ix == nnx && vnr != nnx) swrg = 1;  // sic: compared with the REFINED nnx (:399-401)
      if (iz == 1 && vnt != 1) swrg = 1;
      if (iz == nnz && vnb != nnz) swrg = 1;
      if (swrg) {
        G.node[root].y = 0;
        break;
      }
    }
    G.node[root].y = 0;
    // ---- (1) stencil loads, issued before the pop
    int xx = ix, xz = iz;
    if (grp == 0) xx = ix - 1;
    if (grp == 1) xx = ix + 1;
    if (grp == 2) xz = iz - 1;
    if (grp == 3) xz = iz + 1;
    const bool xin = (xx >= 1 && xx <= nnx && xz >= 1 && xz <= nnz);
    const int xidx = xin ? (xx - 1) * nnz + (xz - 1) : -1;
    int sx = xx, sz = xz;
    {
      const int off = (q & 1) ? 2 : 1;
      const int sgn = (q & 2) ? 1 : -1;
      if (q < 4)
        sx = xx + sgn * off;
      else
        sz = xz + sgn * off;
    }
    const bool sin_ = xin && (sx >= 1 && sx <= nnx && sz >= 1 && sz <= nnz);
    int2 sn = make_int2(0, kOut);
    if (sin_) sn = G.node[(sx - 1) * nnz + (sz - 1)];
    int2 xn = make_int2(0, 0);
    float velx = 1.0f, risti = 0.0f;
    if (xin) {
      xn = G.node[xidx];
      velx = G.vel[xidx];
      risti = G.risti[xx - 1];
    }
    // ids of the four neighbours, known to every lane (for tracking heap moves)
    const int xid0 = (ix - 1 >= 1) ? (ix - 2) * nnz + (iz - 1) : -1;
    const int xid1 = (ix + 1 <= nnx) ? ix * nnz + (iz - 1) : -1;
    const int xid2 = (iz - 1 >= 1) ? (ix - 1) * nnz + (iz - 2) : -1;
    const int xid3 = (iz + 1 <= nnz) ? (ix - 1) * nnz + iz : -1;
    int xm0 = -1, xm1 = -1, xm2 = -1, xm3 = -1;  // new positions of neighbours moved by the pop
#define TRACK_MOVE(nid, newpos)      \
  {                                  \
    if ((nid) == xid0) xm0 = (newpos); \
    if ((nid) == xid1) xm1 = (newpos); \
    if ((nid) == xid2) xm2 = (newpos); \
    if ((nid) == xid3) xm3 = (newpos); \
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
      while (!stop && tpc < ntr && tpc + 1 < HS) {  // both children in shared memory
        const int2 e1 = H.sm[tpc], e2 = H.sm[tpc + 1];
        int2 ec = e1;
        if (keyf(e1) > keyf(e2)) {
          tpc = tpc + 1;
          ec = e2;
        }
        if (keyf(ec) < mk) {
          H.sm[tpp] = ec;
          G.node[ec.y].y = tpp;
          TRACK_MOVE(ec.y, tpp);
          tpp = tpc;
          tpc = 2 * tpp;
        } else {
          stop = true;
        }
      }
      while (!stop && tpc <= ntr) {
        if (tpc + 1 < HS) {  // single child inside shared memory (tpc == ntr)
          const int2 e1 = H.sm[tpc];
          if (keyf(e1) < mk) {
            H.sm[tpp] = e1;
            G.node[e1.y].y = tpp;
            TRACK_MOVE(e1.y, tpp);
            tpp = tpc;
          }
          stop = true;
          break;
        }
        // fetch the four levels below tpp: level r (1..4), offset o -> lane (2^r - 2) + o
        const int r = (lane < 2) ? 1 : (lane < 6) ? 2 : (lane < 14) ? 3 : (lane < 30) ? 4 : 0;
        const int o = lane - ((1 << r) - 2);
        const long long pos = ((long long)tpp << r) + o;
        int2 e = make_int2(0x7f800000, -1);
        if (r > 0 && pos <= ntr) e = H.get((int)pos);
        const int base = tpp;
        int rel = 0;
        for (int lvl = 1; lvl <= 4; lvl++) {
          const long long c0 = ((long long)base << lvl) + 2 * rel;  // left child of the current slot
          if (c0 > ntr) {
            stop = true;
            break;
          }
          const int Lc = (1 << lvl) - 2 + 2 * rel;
          const int k1 = __shfl_sync(kFull, e.x, Lc), n1 = __shfl_sync(kFull, e.y, Lc);
          const int k2 = __shfl_sync(kFull, e.x, Lc + 1), n2 = __shfl_sync(kFull, e.y, Lc + 1);
          int pick = 0;
          if (c0 < ntr && __int_as_float(k1) > __int_as_float(k2)) pick = 1;
          const int2 ec = pick ? make_int2(k2, n2) : make_int2(k1, n1);
          if (keyf(ec) < mk) {
            H.set(tpp, ec);
            G.node[ec.y].y = tpp;
            TRACK_MOVE(ec.y, tpp);
            tpp = (int)c0 + pick;
            rel = 2 * rel + pick;
            if (c0 == ntr) {  // that was the single last child
              stop = true;
              break;
            }
          } else {
            stop = true;
            break;
          }
        }
        tpc = 2 * tpp;
      }
      H.set(tpp, m);
      G.node[m.y].y = tpp;
      TRACK_MOVE(m.y, tpp);
      lastOK = false;
    }
#undef TRACK_MOVE
    // ---- quadrants (as in march<>)
    const float slown = 1.0f / velx;
    const int proc = (xin && xn.y != 0) ? (xn.y == -1 ? 1 : 2) : 0;
    const int gb = lane & ~7;
    const int rq = q & 3;
    const int lj = gb + 2 * (rq >> 1), lk = gb + 4 + 2 * (rq & 1);
    const float Tj = __int_as_float(__shfl_sync(kFull, sn.x, lj));
    const float Tj2 = __int_as_float(__shfl_sync(kFull, sn.x, lj + 1));
    const int Sj = __shfl_sync(kFull, sn.y, lj);
    const int Sj2 = __shfl_sync(kFull, sn.y, lj + 1);
    const float Tk = __int_as_float(__shfl_sync(kFull, sn.x, lk));
    const float Tk2 = __int_as_float(__shfl_sync(kFull, sn.x, lk + 1));
    const int Sk = __shfl_sync(kFull, sn.y, lk);
    const int Sk2 = __shfl_sync(kFull, sn.y, lk + 1);
    float trav = 3.0e38f;
    if (proc && q < 4 && Sj != kOut && Sk != kOut) {
      float tq;
      if (quadrant(Tj, Tj2, Sj, Sj2, Tk, Tk2, Sk, Sk2, slown, G.earth, risti, G.dnx, G.dnz, tq)) trav = tq;
    }
    trav = fminf(trav, __shfl_xor_sync(kFull, trav, 1));
    trav = fminf(trav, __shfl_xor_sync(kFull, trav, 2));
    // ---- (3) planned heap positions + one-shot ancestor fetch
    int pr[4], xi[4], ppos[4];
    float tv[4];
    int nfar = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
      pr[g] = __shfl_sync(kFull, proc, 8 * g);
      tv[g] = __shfl_sync(kFull, trav, 8 * g);
      xi[g] = __shfl_sync(kFull, xidx, 8 * g);
      const int st = __shfl_sync(kFull, xn.y, 8 * g);
      const int xm = (g == 0) ? xm0 : (g == 1) ? xm1 : (g == 2) ? xm2 : xm3;
      ppos[g] = 0;
      if (pr[g] == 1) {
        nfar++;
        ppos[g] = ntr + nfar;
      } else if (pr[g] == 2) {
        ppos[g] = (xm >= 0) ? xm : st;
      }
    }
    const int mypos = (grp == 0) ? ppos[0] : (grp == 1) ? ppos[1] : (grp == 2) ? ppos[2] : ppos[3];
    const int myanc = mypos >> (q + 1);
    int2 anc = make_int2(0, -1);
    if (myanc >= 1) anc = H.get(myanc);
    int2 cand = make_int2(0, -1);  // candidate "last entry" for the next pop when nothing is appended
    if (nfar == 0 && ntr >= 1) cand = H.get(ntr);
    // ---- apply in the reference's order: x-1, x+1, z-1, z+1 (:424-486)
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
      G.node[xg].x = __float_as_int(tvg);
      int tpc;
      if (pr[g] == 1) {
        ntr = ntr + 1;
        if (ntr > hcap) return -1;
        tpc = ntr;
      } else {
        tpc = slow ? G.node[xg].y : ppos[g];
      }
      const bool use_pref = !slow && (tpc == ppos[g]);
      int a = 0;
      bool moved = false;
      int tpp = tpc >> 1;
      while (tpp > 0) {
        int2 pe;
        if (use_pref && a < 8) {
          pe.x = __shfl_sync(kFull, anc.x, 8 * g + a);
          pe.y = __shfl_sync(kFull, anc.y, 8 * g + a);
          if (tpp == ls0) pe = le0;
          if (tpp == ls1) pe = le1;
          if (tpp == ls2) pe = le2;
        } else {
          pe = H.get(tpp);
        }
        if (tvg < keyf(pe)) {
          H.set(tpc, pe);
          G.node[pe.y].y = tpc;
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
      G.node[xg].y = tpc;
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
        lastOK = !moved;  // the entry sits at slot ntr unless it was sifted up
      } else if (appended && moved) {
        // a later sift that moves entries cannot touch slot ntr (it only shifts its own ancestors)
      }
    }
    // ---- (4) moving element of the next pop
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
      lastOK = false;  // conservative: positions may have shifted
    }
  }
  return ntr;
}

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

// ---- uniform interface over the two heap layouts
__device__ __forceinline__ void heap_init(Heap &H, float *smem, int w, const BatchView &bv, int slot) {
  H.sk = smem + (size_t)w * 2 * kHeapSm;
  H.sn = (int *)(H.sk + kHeapSm);
  H.gk = bv.hkey + (size_t)slot * (bv.hcap + 1);
  H.gn = bv.hnode + (size_t)slot * (bv.hcap + 1);
}
template <int HS>
__device__ __forceinline__ void heap_init(Heap2<HS> &H, float *smem, int w, const BatchView &bv, int slot) {
  H.sm = (int2 *)smem + (size_t)w * HS;
  H.gm = bv.hent + (size_t)slot * (bv.hcap + 1);
}
__device__ __forceinline__ void heap_sift(const Heap &H, const Grid &G, int tpc, float key, int xn) {
  sift_up(H, G, tpc, key, xn);
}
template <int HS>
__device__ __forceinline__ void heap_sift(const Heap2<HS> &H, const Grid &G, int tpc, float key, int xn) {
  sift_up2(H, G, tpc, key, xn);
}
template <bool REFINED>
__device__ __forceinline__ int heap_march(const Grid &G, const Heap &H, int ntr, int hcap, int lane, int a, int b,
                                          int c, int d) {
  return march<REFINED>(G, H, ntr, hcap, lane, a, b, c, d);
}
template <bool REFINED, int HS>
__device__ __forceinline__ int heap_march(const Grid &G, const Heap2<HS> &H, int ntr, int hcap, int lane, int a,
                                          int b, int c, int d) {
  return march2<REFINED, HS>(G, H, ntr, hcap, lane, a, b, c, d);
}

template <class HeapT, int MINB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
k_eikonal(Geom g, SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ veln_all,
          const float *__restrict__ velv_all, const float *__restrict__ risti_c, BatchView bv) {
  extern __shared__ float smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * kWarpsPerBlock + w;
  if (slot >= nsw) return;
  HeapT H;
  heap_init(H, smem, w, bv, slot);
  SweepDesc d = sw[slot];
  const size_t Nc = (size_t)g.nnx * g.nnz;
  const float *veln = veln_all + (size_t)d.map * Nc;
  const float *velv = velv_all + (size_t)d.map * g.nx * g.ny;
  int2 *node = bv.node + (size_t)slot * Nc;
  int2 *noder = bv.noder + (size_t)slot * kRefMax * kRefMax;
  float *velr = bv.velr + (size_t)slot * kRefMax * kRefMax;
  const int nrnx = d.nrnx, nrnz = d.nrnz;
  // ---- bsplrefine (:1562-1628): refined velocities + reset refined node states
  {
    const int nrr = g.gd * kSgdl;  // 64 (40 for synthetic)
    const int origx = (d.vnl - 1) * kSgdl + 1, origz = (d.vnt - 1) * kSgdl + 1;
    const int ldv = g.nvx + 2;
    for (int n = lane; n < nrnx * nrnz; n += 32) {
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
  __syncwarp();
  // ---- travel(x, z, urg=1) on the refined grid (:312-375): source cell + 4 corner times
  Grid R;
  R.node = noder;
  R.vel = velr;
  R.risti = bv.ristr + (size_t)slot * kRefMax;
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
        heap_sift(H, R, ntr, t0, xi);
      }
  }
  int rc = heap_march<true>(R, H, ntr, bv.hcap, lane, d.vnl, d.vnr, d.vnt, d.vnb);
  if (rc < 0) {
    if (lane == 0) sw[slot].status = DSURF_ERR_HEAP;
    return;
  }
  __syncwarp();
  // ---- map refined -> coarse (:1289-1303); coarse states were pre-filled with (0,-1)
  const int bw = d.vnr - d.vnl + 1, bh = d.vnb - d.vnt + 1;
  for (int n = lane; n < bw * bh; n += 32) {
    const int cz = n % bh, cx = n / bh;  // offsets inside the box
    const int2 rn = noder[(cx * kSgdl) * nrnz + cz * kSgdl];
    int2 cn = make_int2(0, rn.y);
    if (rn.y >= 0) cn.x = rn.x;
    node[(size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz)] = cn;
  }
  __syncwarp();
  // ---- narrow-band completion (:1332-1349): alive with a far neighbour -> close
  // (order-free: a node turned close is still "not far" for its neighbours)
  for (int n = lane; n < bw * bh; n += 32) {
    const int cz = n % bh, cx = n / bh;
    const int k = d.vnl + cx, l = d.vnt + cz;  // 1-based (ix, iz)
    const size_t o = (size_t)(k - 1) * g.nnz + (l - 1);
    if (node[o].y == 0) {
      bool far = false;
      if (l - 1 >= 1 && node[o - 1].y == -1) far = true;
      if (l + 1 <= g.nnz && node[o + 1].y == -1) far = true;
      if (k - 1 >= 1 && node[o - g.nnz].y == -1) far = true;
      if (k + 1 <= g.nnx && node[o + g.nnz].y == -1) far = true;
      // the reference writes 1 in place; neighbours only test ".EQ.-1", so deferring is identical
      if (far) node[o].y = -100;
    }
  }
  __syncwarp();
  for (int n = lane; n < bw * bh; n += 32) {
    const int cz = n % bh, cx = n / bh;
    const size_t o = (size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz);
    if (node[o].y == -100) node[o].y = 1;
  }
  __syncwarp();
  // ---- travel(x, z, urg=2): rebuild the heap by scanning i=1..nnx, j=1..nnz (:341-347);
  // only nodes of the refined box can be close.
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
  ntr = 0;
  for (int cx = 0; cx < bw; cx++) {
    for (int base = 0; base < bh; base += 32) {
      const int cz = base + lane;
      int st = 0;
      float tt = 0.0f;
      if (cz < bh) {
        const int2 v = node[(size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz)];
        st = v.y;
        tt = __int_as_float(v.x);
      }
      unsigned mask = __ballot_sync(kFull, st > 0);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const float key = __shfl_sync(kFull, tt, b);
        const int xi = (d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + base + b);
        ntr = ntr + 1;
        heap_sift(H, C, ntr, key, xi);
      }
    }
  }
  rc = heap_march<false>(C, H, ntr, bv.hcap, lane, 0, 0, 0, 0);
  if (rc < 0 && lane == 0) sw[slot].status = DSURF_ERR_HEAP;
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
  H.gm = bv.hent + (size_t)slot * (bv.hcap + 1);
  int2 *scr = wbase + (size_t)kNG * kHS3 + (size_t)gg * kScr;
  SweepDesc d = sw[slot];
  const size_t Nc = (size_t)g.nnx * g.nnz;
  const float *veln = veln_all + (size_t)d.map * Nc;
  const float *velv = velv_all + (size_t)d.map * g.nx * g.ny;
  int2 *node = bv.node + (size_t)slot * Nc;
  int2 *noder = bv.noder + (size_t)slot * kRefMax * kRefMax;
  float *velr = bv.velr + (size_t)slot * kRefMax * kRefMax;
  const int nrnx = d.nrnx, nrnz = d.nrnz;
  // ---- bsplrefine (:1562-1628)
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
  R.risti = bv.ristr + (size_t)slot * kRefMax;
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
    float vsrc = 0.0f;
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
  int rc = march3<true, kG, LAZY>(R, H, scr, ntr, bv.hcap, gl, gm, gbase, wmask, d.vnl, d.vnr, d.vnt, d.vnb);
  const bool failed = rc < 0;
  if (failed && gl == 0) sw[slot].status = DSURF_ERR_HEAP;
  __syncwarp(gm);
  const int bw = d.vnr - d.vnl + 1, bh = d.vnb - d.vnt + 1;
  for (int n = gl; n < bw * bh; n += kG) {
    const int cz = n % bh, cx = n / bh;
    const int2 rn = noder[(cx * kSgdl) * nrnz + cz * kSgdl];
    int2 cn = make_int2(0, rn.y);
    if (rn.y >= 0) cn.x = rn.x;
    node[(size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz)] = cn;
  }
  __syncwarp(gm);
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
      if (far) node[o].y = -100;
    }
  }
  __syncwarp(gm);
  for (int n = gl; n < bw * bh; n += kG) {
    const int cz = n % bh, cx = n / bh;
    const size_t o = (size_t)(d.vnl - 1 + cx) * g.nnz + (d.vnt - 1 + cz);
    if (node[o].y == -100) node[o].y = 1;
  }
  __syncwarp(gm);
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
  ntr = 0;
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
  rc = march3<false, kG, LAZY>(C, H, scr, ntr, bv.hcap, gl, gm, gbase, wmask, 0, 0, 0, 0);
  if (rc < 0 && gl == 0) sw[slot].status = DSURF_ERR_HEAP;
}

constexpr int kHeapSm2 = 768;  // v2: 6 KB of heap per warp -> 4 blocks (32 warps) per SM

// sweeps that can be resident at once with the selected kernel variant (batches larger than this
// run as several waves of equal duration, so the plan sizes its batches to a multiple of it)
int eikonal_resident_sweeps() {
  const bool v1 = getenv("DSURF_EIKONAL_V1") != nullptr, v2 = getenv("DSURF_EIKONAL_V2") != nullptr;
  const char *gs = getenv("DSURF_EIKONAL_G");
  const bool g8 = gs && atoi(gs) == 8;
  int nb = 0;
  int per_block = kWarpsPerBlock;
  if (v1) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eikonal<Heap, 1>, kWarpsPerBlock * 32,
                                                  (size_t)kWarpsPerBlock * 2 * kHeapSm * sizeof(float));
  } else if (v2) {
    cudaFuncSetAttribute(k_eikonal<Heap2<kHeapSm2>, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)kWarpsPerBlock * kHeapSm2 * sizeof(int2)));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eikonal<Heap2<kHeapSm2>, 4>, kWarpsPerBlock * 32,
                                                  (size_t)kWarpsPerBlock * kHeapSm2 * sizeof(int2));
  } else if (g8) {
    const size_t sm = (size_t)kWarpsPerBlock * V3<8>::NG * (V3<8>::HS + kScr) * sizeof(int2);
    cudaFuncSetAttribute(k_eikonal3<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eikonal3<8, true>, kWarpsPerBlock * 32, sm);
    per_block = kWarpsPerBlock * V3<8>::NG;
  } else {
    const size_t sm = (size_t)kWarpsPerBlock * V3<16>::NG * (V3<16>::HS + kScr) * sizeof(int2);
    cudaFuncSetAttribute(k_eikonal3<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_eikonal3<16, true>, kWarpsPerBlock * 32, sm);
    per_block = kWarpsPerBlock * V3<16>::NG;
  }
  cudaGetLastError();
  if (nb < 1) nb = 1;
  return nb * per_block * sm_count();
}

int launch_eikonal(cudaStream_t st, const Geom &g, const SweepDesc *d_sw, int nsw, const float *d_veln_all,
                   const float *d_velv_all, const float *d_risti, BatchView bv, int *launches) {
  if (nsw <= 0) return DSURF_OK;
  const long long ntot = (long long)nsw * g.nnx * g.nnz;
  k_fill_nodes<<<sm_count() * 8, kWarpsPerBlock * 32, 0, st>>>(bv.node, ntot);
  static const bool use_v1 = getenv("DSURF_EIKONAL_V1") != nullptr;  // reference variants for A/B tests
  static const bool use_v2 = getenv("DSURF_EIKONAL_V2") != nullptr;
  static const char *gsel = getenv("DSURF_EIKONAL_G");  // lanes per sweep of the v3 kernel: 16 (default) or 8
  if (!use_v1 && !use_v2) {
    const bool g8 = gsel && atoi(gsel) == 8;
    const int ng = g8 ? V3<8>::NG : V3<16>::NG;
    const int hs = g8 ? V3<8>::HS : V3<16>::HS;
    const size_t smem = (size_t)kWarpsPerBlock * ng * (hs + kScr) * sizeof(int2);
    // lazy heap back-pointers (see march3): measured 6 % SLOWER than the eager scheme at cfg 3 (the kernel is
    // issue-bound, and a slot search by one sweep of a warp stalls the other), so eager is the default
    static const bool eager = getenv("DSURF_EIKONAL_LAZY") == nullptr;
    static bool attr3 = false;
    if (!attr3) {
      const int s8 = (int)((size_t)kWarpsPerBlock * V3<8>::NG * (V3<8>::HS + kScr) * sizeof(int2));
      const int s16 = (int)((size_t)kWarpsPerBlock * V3<16>::NG * (V3<16>::HS + kScr) * sizeof(int2));
      DS_CUDA(cudaFuncSetAttribute(k_eikonal3<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s8));
      DS_CUDA(cudaFuncSetAttribute(k_eikonal3<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s8));
      DS_CUDA(cudaFuncSetAttribute(k_eikonal3<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
      DS_CUDA(cudaFuncSetAttribute(k_eikonal3<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
      attr3 = true;
    }
    const int per_block = kWarpsPerBlock * ng;
    const int grid3 = (nsw + per_block - 1) / per_block;
    SweepDesc *sw = const_cast<SweepDesc *>(d_sw);
    const int nt = kWarpsPerBlock * 32;
    if (g8 && eager)
      k_eikonal3<8, false><<<grid3, nt, smem, st>>>(g, sw, nsw, d_veln_all, d_velv_all, d_risti, bv);
    else if (g8)
      k_eikonal3<8, true><<<grid3, nt, smem, st>>>(g, sw, nsw, d_veln_all, d_velv_all, d_risti, bv);
    else if (eager)
      k_eikonal3<16, false><<<grid3, nt, smem, st>>>(g, sw, nsw, d_veln_all, d_velv_all, d_risti, bv);
    else
      k_eikonal3<16, true><<<grid3, nt, smem, st>>>(g, sw, nsw, d_veln_all, d_velv_all, d_risti, bv);
    DS_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    return DSURF_OK;
  }
  const int grid = (nsw + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (use_v1) {
    const size_t smem = (size_t)kWarpsPerBlock * 2 * kHeapSm * sizeof(float);
    static bool attr = false;
    if (!attr) {
      DS_CUDA(cudaFuncSetAttribute(k_eikonal<Heap, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    k_eikonal<Heap, 1><<<grid, kWarpsPerBlock * 32, smem, st>>>(g, const_cast<SweepDesc *>(d_sw), nsw, d_veln_all,
                                                                d_velv_all, d_risti, bv);
  } else {
    const size_t smem = (size_t)kWarpsPerBlock * kHeapSm2 * sizeof(int2);
    static bool attr = false;
    if (!attr) {
      DS_CUDA(cudaFuncSetAttribute(k_eikonal<Heap2<kHeapSm2>, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
      attr = true;
    }
    k_eikonal<Heap2<kHeapSm2>, 4><<<grid, kWarpsPerBlock * 32, smem, st>>>(
        g, const_cast<SweepDesc *>(d_sw), nsw, d_veln_all, d_velv_all, d_risti, bv);
  }
  DS_CUDA(cudaGetLastError());
  if (launches) *launches += 2;
  return DSURF_OK;
}

}  // namespace dsurf
