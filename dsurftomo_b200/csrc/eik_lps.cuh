// K3 (round 2) -- the coarse-grid march of travel (src/CalSurfG.f90:386-486) with fouds2 (:587-759)
// and the binary heap addtree/downtree/updtree (:768-921), written as SCALAR code: one thread
// marches one whole (gather, pass) sweep in the reference's exact pop order ("lane per sweep",
// 32 independent sweeps per warp).  The same source compiles for the host, where
// tests/lps_host_check.cpp replays it against the oracle's Fmm::travel bit for bit.
//
// Node state is ONE 32-bit word per node:
//     alive  : the fp32 travel time (sign bit clear; times are never negative)
//     close  : 0x80000000 | s, where s is a heap slot with the LAZY invariant below
//     far    : 0xFFFFFFFF
// The trial time of a close node lives only in its heap entry (key, node): fouds2 reads times of
// ALIVE nodes only (:620-663), so nothing ever needs the trial time through the grid.
//
// LAZY back-pointers.  The reference stores every node's exact heap slot in nsts and rewrites it on
// every heap move (~11 scattered 4-byte stores per pop).  Here a move is recorded only when an
// entry moves DOWN or jumps: the parents a sift-up pushes down, the entry the sift-up places, and
// the last heap element that downtree re-inserts.  The entries a sift-down pulls UP one level are
// not recorded.  Invariant: the true slot of a close node is s >> u for some u >= 0, i.e. an
// ancestor-or-self of the stored slot; updtree finds it by comparing node ids along that chain,
// which it has to read anyway for the sift-up.  Heap contents, and therefore pop order, tie
// behaviour and every travel time, are exactly those of the reference.
#pragma once
#include <cstdint>
#include <cstring>
#if defined(__CUDACC__)
#define LPS_HD __host__ __device__ __forceinline__
#else
#define LPS_HD inline
#include <cmath>
#endif

namespace dsurf {
namespace lps {

constexpr uint32_t kFar = 0xFFFFFFFFu;
constexpr uint32_t kCloseBit = 0x80000000u;

struct Ent {  // heap entry: fp32 key bits + node index (iz fastest), layout-compatible with int2
  int x, y;
};

LPS_HD float bits2f(int b) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
LPS_HD int f2bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int b;
  memcpy(&b, &f, 4);
  return b;
#endif
}
LPS_HD float sqrt_rn(float x) {
#if defined(__CUDA_ARCH__)
  return sqrtf(x);
#else
  return std::sqrt(x);
#endif
}
LPS_HD bool alive(uint32_t w) { return (int)w >= 0; }

struct GridP {
  int nnx, nnz;
  float dnx, dnz, earth;
};

// one quadrant of fouds2 (:664-756).  aj/ak: the 1-away node is alive; swj/swk: the second-order
// leg is usable (:620-663).  Operand order is the reference's.
LPS_HD bool quad(float Tj, float Tj2, bool aj, bool swj, float Tk, float Tk2, bool ak, bool swk, float slown, float ri,
                 float risti, float dnx, float dnz, float &trav) {
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
    } else if (ak) {
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
  } else if (aj) {
    if (swk) {
      u = ri * dnx;
      v = 2.0f * risti * dnz;
      em = 3.0f * Tj - 4.0f * Tk + Tk2;
      a = v * v + 9.0f * (u * u);
      b = 6.0f * em * (u * u);
      c = (u * u) * (em * em - (v * v) * (slown * slown));
      tref = Tj;
      tdiv = 1.0f;
    } else if (ak) {
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
    } else if (ak) {
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
  const float tdsh = (-b + sqrt_rn(rd1)) / (2.0f * a);
  trav = (tref + tdsh) / tdiv;
  return true;
}

// fouds2 for one node: wj1[s]/wj2[s] = words of the 1-away / 2-away node on the x side s (0: ix-1,
// 1: ix+1), inj[s] = the 1-away node is inside the grid (the reference skips the side otherwise,
// :604-605); same for z.  Returns the minimum over the solvable quadrants (0 if none, as :758).
LPS_HD float fouds2_words(const uint32_t wj1[2], const uint32_t wj2[2], const bool inj[2], const uint32_t wk1[2],
                          const uint32_t wk2[2], const bool ink[2], float slown, float ri, float risti, float dnx,
                          float dnz) {
  float travm = 0.0f;
  bool have = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int sj = 0; sj < 2; sj++) {
    if (!inj[sj]) continue;
    const bool aj = alive(wj1[sj]);
    const float Tj = bits2f((int)wj1[sj]), Tj2 = bits2f((int)wj2[sj]);
    const bool swj = aj && alive(wj2[sj]) && (Tj > Tj2);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int sk = 0; sk < 2; sk++) {
      if (!ink[sk]) continue;
      const bool ak = alive(wk1[sk]);
      const float Tk = bits2f((int)wk1[sk]), Tk2 = bits2f((int)wk2[sk]);
      const bool swk = ak && alive(wk2[sk]) && (Tk > Tk2);
      float tq;
      if (quad(Tj, Tj2, aj, swj, Tk, Tk2, ak, swk, slown, ri, risti, dnx, dnz, tq)) {
        if (have) {
          travm = tq < travm ? tq : travm;  // MIN(trav, travm), :752
        } else {
          travm = tq;
          have = true;
        }
      }
    }
  }
  return travm;
}

// addtree (:768-805) / the sift-up half of updtree (:894-921) from slot tpc with key tv.
// Parents pushed down get their exact slot recorded; returns the landing slot.
template <class M>
LPS_HD int sift_up(M &m, int tpc, float tv, int node, bool &moved) {
  moved = false;
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const Ent pe = m.hget(tpp);
    if (tv < bits2f(pe.x)) {
      m.hset(tpc, pe);
      m.set_word(pe.y, kCloseBit | (uint32_t)tpc);
      tpc = tpp;
      tpp = tpc >> 1;
      moved = true;
    } else {
      tpp = 0;
    }
  }
  Ent ne;
  ne.x = f2bits(tv);
  ne.y = node;
  m.hset(tpc, ne);
  return tpc;
}

// travel's march loop (:386-486) on the coarse grid (urg = 2 or 0: no refined-grid exit test).
// Returns 0 when the heap ran empty, -1 if the narrow band outgrew hcap, -2 on a broken invariant.
//
// M::kLanes lanes cooperate on one sweep (1 on the host, 4 on the device: lane g owns neighbour g of
// the accepted node -- its stencil loads, its fouds2 and the prefetch of its heap chain); the heap
// itself is walked redundantly by the lanes of a sweep (same addresses, same values).
//
// Memory policy M:
//   word(i) / set_word(i, w)   node words;  vel(i), risti(ix)   velocity and earth*sin(colatitude)
//   hget(p) / hset(p, e)       heap slot p (1-based)
//   hget2(p, a, b)             slots p (even) and p + 1
//   hblock(q, e[14])           the 2 + 4 + 8 descendants of slot q at relative depths 1..3, q on a level
//                              M::kLg - 1 + 3k (one 128-byte line of the device slab)
//   M::kLg                     slots below 2^kLg are cheap (shared memory on the device)
//   lane(), bcast(v, src), any(pred)   lane index inside the sweep's group, value of lane src, warp-wide OR
//
// One acceptance is organised in dependent memory ROUND TRIPS:
//   trip 1   stencil words + velocity of the owned neighbour(s), and the last heap element if it is not
//            already in registers -- all issued together;
//   trip 2-3 downtree: levels below 2^kLg from shared memory, then three levels per block fetch;
//   trip 4   the ancestor chains of the (up to four) insert/update slots, fetched together; the
//            updates are then applied in the reference's order from registers.  A three-entry log
//            carries entries written by earlier neighbours of the same acceptance; a sift-up that
//            moves entries (rare) sends the remaining neighbours to plain loads.
template <class M>
LPS_HD int march(const GridP &G, M &m, int ntr, int hcap) {
  const int nnx = G.nnx, nnz = G.nnz;
  constexpr int kSm = 1 << M::kLg;  // first slot outside the cheap levels
  constexpr int kGC = 6;            // chain entries prefetched per neighbour (levels kLg .. kLg+5)
  constexpr int NL = M::kLanes, NPER = 4 / NL;
  const int me = m.lane();
  Ent last;
  last.x = 0;
  last.y = -1;
  bool lastOK = false;
  int err = 0;
  while (m.any(ntr > 0)) {  // the sweeps of a warp stay in lock step (all need ~nnx*nnz acceptances)
    if (ntr <= 0) continue;
    m.stat(2, ntr);
    const Ent r = m.hget(1);
    const int root = r.y;
    const int ix = root / nnz, iz = root - ix * nnz;  // 0-based
    const uint32_t wX = (uint32_t)r.x;                 // the accepted node: alive with its key
    // ---- trip 1: the owned neighbour(s): own word, 8-point stencil (:620-663), velocity
    uint32_t wN[NPER], wj1[NPER][2], wj2[NPER][2], wk1[NPER][2], wk2[NPER][2];
    bool inN[NPER], inj[NPER][2], ink[NPER][2];
    float vl[NPER];
    auto ldw = [&](int x, int z) -> uint32_t {
      if (x == ix && z == iz) return wX;
      return (x >= 0 && x < nnx && z >= 0 && z < nnz) ? m.word(x * nnz + z) : kFar;
    };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int gi = 0; gi < NPER; gi++) {
      const int g = me + NL * gi;
      const int nx = ix + ((g == 0) ? -1 : (g == 1) ? 1 : 0), nz = iz + ((g == 2) ? -1 : (g == 3) ? 1 : 0);
      inN[gi] = nx >= 0 && nx < nnx && nz >= 0 && nz < nnz;
      wN[gi] = ldw(nx, nz);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int sd = 0; sd < 2; sd++) {
        const int sg = sd ? 1 : -1;
        inj[gi][sd] = nx + sg >= 0 && nx + sg < nnx;
        ink[gi][sd] = nz + sg >= 0 && nz + sg < nnz;
        wj1[gi][sd] = ldw(nx + sg, nz);
        wj2[gi][sd] = ldw(nx + 2 * sg, nz);
        wk1[gi][sd] = ldw(nx, nz + sg);
        wk2[gi][sd] = ldw(nx, nz + 2 * sg);
      }
      vl[gi] = inN[gi] ? m.vel(nx * nnz + nz) : 1.0f;
    }
    if (!lastOK && ntr > 1) last = m.hget(ntr);
    m.set_word(root, wX);  // alive, time = its key (:415-417)
    // ---- trips 2-3: downtree (:816-885): the last element sinks from the root; entries pulled up are NOT recorded
    int landed = -1;  // slot where the last element landed
    if (ntr == 1) {
      ntr = 0;
    } else {
      const float mk = bits2f(last.x);
      ntr = ntr - 1;
      int tpp = 1, tpc = 2;
      bool stop = false;
      while (!stop && tpc < ntr && tpc + 1 < kSm) {  // both children in the cheap levels
        Ent e1, e2;
        m.hget2(tpc, e1, e2);
        Ent ec = e1;
        if (bits2f(e1.x) > bits2f(e2.x)) {
          tpc = tpc + 1;
          ec = e2;
        }
        if (bits2f(ec.x) < mk) {
          m.hset(tpp, ec);
          tpp = tpc;
          tpc = 2 * tpp;
        } else {
          stop = true;
        }
      }
      while (!stop && tpc <= ntr) {
        if (tpc + 1 < kSm) {  // single child inside the cheap levels (tpc == ntr)
          const Ent ec = m.hget(tpc);
          if (bits2f(ec.x) < mk) {
            m.hset(tpp, ec);
            tpp = tpc;
          }
          break;
        }
        Ent b[14];  // the three levels below tpp
        m.hblock(tpp, b);
        const int base = tpp;
        int rel = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int lvl = 1; lvl <= 3; lvl++) {
          const int c0 = (base << lvl) + 2 * rel;  // left child of the current slot
          if (stop || c0 > ntr) {
            stop = true;
            continue;
          }
          Ent e1, e2;
          if (lvl == 1) {
            e1 = b[0];
            e2 = b[1];
          } else if (lvl == 2) {
            e1 = rel ? b[4] : b[2];
            e2 = rel ? b[5] : b[3];
          } else {
            e1 = (rel == 0) ? b[6] : (rel == 1) ? b[8] : (rel == 2) ? b[10] : b[12];
            e2 = (rel == 0) ? b[7] : (rel == 1) ? b[9] : (rel == 2) ? b[11] : b[13];
          }
          int pick = 0;
          if (c0 < ntr && bits2f(e1.x) > bits2f(e2.x)) pick = 1;
          const Ent ec = pick ? e2 : e1;
          if (bits2f(ec.x) < mk) {
            m.hset(tpp, ec);
            tpp = c0 + pick;
            rel = 2 * rel + pick;
            if (c0 == ntr) stop = true;  // that was the single last child
          } else {
            stop = true;
          }
        }
        tpc = 2 * tpp;
      }
      m.hset(tpp, last);
      m.set_word(last.y, kCloseBit | (uint32_t)tpp);
      landed = tpp;
    }
    lastOK = false;
    // ---- trial times of the owned neighbour(s) (independent: only alive nodes enter fouds2)
    float tvm[NPER];
    int kindm[NPER];  // 0: nothing to do, 1: far -> addtree, 2: close -> updtree
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int gi = 0; gi < NPER; gi++) {
      const int g = me + NL * gi;
      kindm[gi] = 0;
      tvm[gi] = 0.0f;
      if (!inN[gi] || alive(wN[gi])) continue;
      kindm[gi] = (wN[gi] == kFar) ? 1 : 2;
      const int nix = ix + ((g == 0) ? -1 : (g == 1) ? 1 : 0);
      tvm[gi] = fouds2_words(wj1[gi], wj2[gi], inj[gi], wk1[gi], wk2[gi], ink[gi], 1.0f / vl[gi], G.earth, m.risti(nix),
                             G.dnx, G.dnz);
    }
    // ---- everybody learns the four results; chain heads
    float tv[4];
    int kind[4], q[4];  // q: first chain slot = stored slot of a close node, parent of the new slot of a far node
    {
      int nf = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int g = 0; g < 4; g++) {
        kind[g] = m.bcast(kindm[g / NL], g % NL);
        tv[g] = m.bcast(tvm[g / NL], g % NL);
        uint32_t w = m.bcast(wN[g / NL], g % NL);
        const int nidx = (g == 0) ? root - nnz : (g == 1) ? root + nnz : (g == 2) ? root - 1 : root + 1;
        // the words were read before downtree's store: the landed element may be this neighbour
        if (kind[g] == 2 && landed >= 0 && last.y == nidx) w = kCloseBit | (uint32_t)landed;
        q[g] = 0;
        if (kind[g] == 1) {
          nf++;
          q[g] = (ntr + nf) >> 1;
        } else if (kind[g] == 2) {
          q[g] = (int)(w & 0x7FFFFFFFu);
        }
      }
      if (ntr + nf > hcap) {
        err = -1;
        ntr = 0;
        continue;
      }
    }
    // ---- trip 4: chain slots of the owned heap operation(s), fetched together
    Ent pre[NPER][kGC];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int gi = 0; gi < NPER; gi++) {
      int q0 = q[0];
      if (me + NL * gi == 1) q0 = q[1];
      if (me + NL * gi == 2) q0 = q[2];
      if (me + NL * gi == 3) q0 = q[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int c = 0; c < kGC; c++) {
        const int sl = q0 >> c;
        pre[gi][c].x = 0;
        pre[gi][c].y = -1;
        if (sl >= kSm && sl <= hcap) pre[gi][c] = m.hget(sl);
      }
    }
    // ---- apply in the reference's order x-1, x+1, z-1, z+1 (:419-440)
    bool dirty = false;  // a sift-up of this acceptance moved entries: later neighbours read memory directly
    int ls0 = -1, ls1 = -1, ls2 = -1, nl = 0;
    Ent le0 = last, le1 = last, le2 = last;
    int lastAt = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < 4; g++) {
      if (kind[g] == 0) continue;
      const int nidx = (g == 0) ? root - nnz : (g == 1) ? root + nnz : (g == 2) ? root - 1 : root + 1;
      int q0 = q[g];
      if (dirty && kind[g] == 2) q0 = (int)(m.word(nidx) & 0x7FFFFFFFu);
      auto rd = [&](int sl, int c) -> Ent {  // chain reader: slot sl = q0 >> c
        Ent e;
        if (!dirty && sl >= kSm && c < kGC) {
          Ent eo = pre[g / NL][0];  // select instead of indexing by c (registers)
          if (c == 1) eo = pre[g / NL][1];
          if (c == 2) eo = pre[g / NL][2];
          if (c == 3) eo = pre[g / NL][3];
          if (c == 4) eo = pre[g / NL][4];
          if (c == 5) eo = pre[g / NL][5];
          e.x = m.bcast(eo.x, g % NL);
          e.y = m.bcast(eo.y, g % NL);
        } else {
          e = m.hget(sl);
        }
        if (!dirty) {  // entries written by earlier neighbours of this acceptance (none of them moved anything)
          if (sl == ls0) e = le0;
          if (sl == ls1) e = le1;
          if (sl == ls2) e = le2;
        }
        return e;
      };
      int tpc, c = 0;
      if (kind[g] == 1) {  // addtree (:768-805)
        ntr = ntr + 1;
        tpc = ntr;
      } else {  // updtree (:894-921): locate the entry on the ancestor chain of the stored slot
        int p = q0;
        while (p > 0) {
          m.stat(0, 1);
          if (p <= ntr && rd(p, c).y == nidx) break;
          p >>= 1;
          c++;
        }
        m.stat(1, 1);
        if (p == 0) {
          err = -2;
          break;
        }
        tpc = p;
        c++;  // chain index of tpc's parent
      }
      const float t = tv[g];
      bool moved = false;
      int tpp = tpc >> 1;
      while (tpp > 0) {
        const Ent pe = rd(tpp, c);
        if (t < bits2f(pe.x)) {
          m.hset(tpc, pe);
          m.set_word(pe.y, kCloseBit | (uint32_t)tpc);
          tpc = tpp;
          tpp = tpc >> 1;
          c++;
          moved = true;
        } else {
          tpp = 0;
        }
      }
      Ent ne;
      ne.x = f2bits(t);
      ne.y = nidx;
      m.hset(tpc, ne);
      if (kind[g] == 1 || moved) m.set_word(nidx, kCloseBit | (uint32_t)tpc);
      if (moved) {
        dirty = true;
        lastAt = -1;
      } else {
        if (nl == 0) { ls0 = tpc; le0 = ne; }
        if (nl == 1) { ls1 = tpc; le1 = ne; }
        if (nl == 2) { ls2 = tpc; le2 = ne; }
        nl++;
        if (kind[g] == 1 || tpc == lastAt) {  // the entry at the last slot is known
          if (kind[g] == 1) lastAt = tpc;
          last = ne;
        }
      }
    }
    if (err) ntr = 0;
    lastOK = (lastAt == ntr) && ntr > 0;
  }
  return err;
}

}  // namespace lps
}  // namespace dsurf
