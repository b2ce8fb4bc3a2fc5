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
                       int nnx, int nnz) {
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
  k_dice<<<(nn + 255) / 256, 256, 0, st>>>(d_velv, d_veln, g.nvx, g.nvz, g.nnx, g.nnz);
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

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_fill_nodes(int2 *node, long long n) {
  const long long tot = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += tot) node[i] = make_int2(0, -1);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_eikonal(Geom g, SweepDesc *__restrict__ sw, int nsw, const float *__restrict__ veln_all,
          const float *__restrict__ velv_all, const float *__restrict__ risti_c, BatchView bv) {
  extern __shared__ float smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * kWarpsPerBlock + w;
  if (slot >= nsw) return;
  Heap H;
  H.sk = smem + (size_t)w * 2 * kHeapSm;
  H.sn = (int *)(H.sk + kHeapSm);
  H.gk = bv.hkey + (size_t)slot * (bv.hcap + 1);
  H.gn = bv.hnode + (size_t)slot * (bv.hcap + 1);
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
    const int nrr = kGd * kSgdl;  // 64
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
        sift_up(H, R, ntr, t0, xi);
      }
  }
  int rc = march<true>(R, H, ntr, bv.hcap, lane, d.vnl, d.vnr, d.vnt, d.vnb);
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
        sift_up(H, C, ntr, key, xi);
      }
    }
  }
  rc = march<false>(C, H, ntr, bv.hcap, lane, 0, 0, 0, 0);
  if (rc < 0 && lane == 0) sw[slot].status = DSURF_ERR_HEAP;
}

int launch_eikonal(cudaStream_t st, const Geom &g, const SweepDesc *d_sw, int nsw, const float *d_veln_all,
                   const float *d_velv_all, const float *d_risti, BatchView bv, int *launches) {
  if (nsw <= 0) return DSURF_OK;
  const long long ntot = (long long)nsw * g.nnx * g.nnz;
  k_fill_nodes<<<sm_count() * 8, kWarpsPerBlock * 32, 0, st>>>(bv.node, ntot);
  const size_t smem = (size_t)kWarpsPerBlock * 2 * kHeapSm * sizeof(float);
  static bool attr = false;
  if (!attr) {
    DS_CUDA(cudaFuncSetAttribute(k_eikonal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int grid = (nsw + kWarpsPerBlock - 1) / kWarpsPerBlock;
  k_eikonal<<<grid, kWarpsPerBlock * 32, smem, st>>>(g, const_cast<SweepDesc *>(d_sw), nsw, d_veln_all,
                                                    d_velv_all, d_risti, bv);
  DS_CUDA(cudaGetLastError());
  if (launches) *launches += 2;
  return DSURF_OK;
}

}  // namespace dsurf
