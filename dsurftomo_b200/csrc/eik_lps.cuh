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

// Node ids are indices into the PADDED word array: node (ix, iz), 0-based, lives at
// (ix + kPad) * ld + (iz + kPad) with ld = nnz + 2 * kPad; the kPad-wide frame is permanently far, so
// stencil loads need no bounds tests (the reference's "outside the grid -> skip the side" tests
// (:604-605) are evaluated on the coordinates).
constexpr int kPad = 3;

struct NbrOut {  // what the accepted node's neighbour g needs from the heap: nothing / addtree / updtree
  int kind;      // 0: outside or alive, 1: far, 2: close
  float tv;      // fouds2's trial time
  uint32_t w;    // the neighbour's word as read (stored heap slot when close)
};

// fouds2 (:587-759) for neighbour g (0: ix-1, 1: ix+1, 2: iz-1, 3: iz+1) of the node `root` accepted
// with key bits wX.  W: word(i), vel(i) on padded indices, risti(ix).
template <class W>
LPS_HD NbrOut eval_neighbour(const GridP &G, const W &wm, int root, uint32_t wX, int g) {
  const int ld = G.nnz + 2 * kPad;
  const int xp = root / ld;
  const int ix = xp - kPad, iz = root - xp * ld - kPad;  // accepted node, 0-based
  const int nx = ix + ((g == 0) ? -1 : (g == 1) ? 1 : 0), nz = iz + ((g == 2) ? -1 : (g == 3) ? 1 : 0);
  const int nid = root + ((g == 0) ? -ld : (g == 1) ? ld : (g == 2) ? -1 : 1);
  NbrOut o;
  o.kind = 0;
  o.tv = 0.0f;
  o.w = kFar;
  if (nx < 0 || nx >= G.nnx || nz < 0 || nz >= G.nnz) return o;
  o.w = wm.word(nid);
  if (alive(o.w)) return o;
  o.kind = (o.w == kFar) ? 1 : 2;
  uint32_t wj1[2], wj2[2], wk1[2], wk2[2];
  bool inj[2], ink[2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int sd = 0; sd < 2; sd++) {
    const int sg = sd ? 1 : -1;
    inj[sd] = nx + sg >= 0 && nx + sg < G.nnx;
    ink[sd] = nz + sg >= 0 && nz + sg < G.nnz;
    // the accepted node is one of the 1-away nodes: it is alive with its key (its word may not be stored yet)
    wj1[sd] = (nid + sg * ld == root) ? wX : wm.word(nid + sg * ld);
    wj2[sd] = wm.word(nid + 2 * sg * ld);
    wk1[sd] = (nid + sg == root) ? wX : wm.word(nid + sg);
    wk2[sd] = wm.word(nid + 2 * sg);
  }
  o.tv = fouds2_words(wj1, wj2, inj, wk1, wk2, ink, 1.0f / wm.vel(nid), G.earth, wm.risti(nx), G.dnx, G.dnz);
  return o;
}

// travel's march loop (:386-486) on the coarse grid (urg = 2 or 0: no refined-grid exit test): the HEAP side.
// Returns 0 when the heap ran empty, -1 if the narrow band outgrew hcap, -2 on a broken invariant.
//
// The four fouds2 evaluations of an acceptance are independent of the heap (only alive nodes enter them)
// and are obtained through the policy: m.publish(root, key, state) announces the accepted node,
// m.collect(tv) returns the four trial times.  On the device they are computed by other warps of the block
// (lane = sweep everywhere) while this warp sifts the heap; on the host collect() just calls eval_neighbour.
//
// Memory policy M:
//   word(i) / set_word(i, w)   node words (padded index)
//   hget(p) / hset(p, e)       heap slot p (1-based);  hget2(p, a, b): slots p (even) and p + 1
//   hblock(q, e[14], lim)      the 2 + 4 + 8 descendants of slot q at relative depths 1..3 (slots <= lim)
//   M::kLg                     slots below 2^kLg are cheap (shared memory on the device)
//   div_ld(i)                  i / ld
//   any(pred), sync()          warp-wide OR / re-convergence point;  publish / collect as above
//
// One acceptance is organised in a FIXED schedule of dependent memory round trips.  A warp is 32 unrelated
// sweeps that advance in lock step, so the slowest lane sets the pace of every acceptance: no lane may take a
// private detour through memory, however rare (measured: with 32 lanes a 5 % event happens every time).
//   trip 1   the four neighbour words (decides addtree / updtree / nothing and gives the stored slots);
//   trip 2-3 downtree: levels below 2^kLg from shared memory, then three levels per fetch;
//   trip 4   issued while the trial times are still being computed: for each of the (up to four) heap
//            operations EVERY chain slot outside the cheap levels, plus the heap's last slot.
// The updates are then applied in the reference's order from registers and shared memory only: every heap
// write goes through put(), which also patches the later neighbours' prefetched copies and the tracked last
// element (the next downtree starts from it without a load).  The one case that re-reads memory is a later
// neighbour being pushed down by an earlier neighbour's sift-up (its chain changes altogether).
template <class M>
LPS_HD int march(const GridP &G, M &m, int ntr, int hcap) {
  const int ld = G.nnz + 2 * kPad;
  constexpr int kGC = 6;  // prefetched chain entries per neighbour: every level outside the cheap ones up to slot 2^(kLg+6) - 1
  Ent last;
  last.x = 0;
  last.y = -1;
  bool lastOK = false;
  int err = 0;
  constexpr int kSm = 1 << M::kLg;  // first slot outside the cheap levels
  for (;;) {  // the 32 sweeps of a warp stay in lock step (all need ~nnx*nnz acceptances)
    const bool active = ntr > 0;
    const bool anyact = m.any(active);
    Ent r;
    r.x = 0;
    r.y = -1;
    if (active) r = m.hget(1);
    m.publish(r.y, (uint32_t)r.x, anyact ? (active ? 1 : 0) : -1);
    if (!anyact) break;
    const int root = r.y;
    int kind[4] = {0, 0, 0, 0};  // 0: outside or alive, 1: far -> addtree, 2: close -> updtree
    int q[4] = {0, 0, 0, 0};     // chain head: stored slot of a close node, parent of the new slot of a far node
    Ent pre[4][kGC];
    if (active) {
      m.stat(2, ntr);
      // ---- trip 1
      uint32_t a1[4];
      a1[0] = m.word(root - ld);
      a1[1] = m.word(root + ld);
      a1[2] = m.word(root - 1);
      a1[3] = m.word(root + 1);
      if (!lastOK && ntr > 1) last = m.hget(ntr);  // first acceptance only: afterwards the last element is tracked
      m.set_word(root, (uint32_t)r.x);  // alive, time = its key (:415-417)
      // ---- downtree (:816-885): the last element sinks from the root; entries pulled up are NOT recorded
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
          m.hblock(tpp, b, ntr);
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
      // ---- what the four neighbours need, and where their chains start
      const int xp = m.div_ld(root), zp = root - xp * ld;
      const bool inN[4] = {xp - 1 >= kPad, xp + 1 < G.nnx + kPad, zp - 1 >= kPad, zp + 1 < G.nnz + kPad};
      int nf = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int g = 0; g < 4; g++) {
        const int nidx = root + ((g == 0) ? -ld : (g == 1) ? ld : (g == 2) ? -1 : 1);
        // the words were read before downtree's store: the landed element may be this neighbour
        if (landed >= 0 && last.y == nidx) a1[g] = kCloseBit | (uint32_t)landed;
        if (inN[g] && !alive(a1[g])) {
          if (a1[g] == kFar) {
            kind[g] = 1;
            nf++;
            q[g] = (ntr + nf) >> 1;
          } else {
            kind[g] = 2;
            q[g] = (int)(a1[g] & 0x7FFFFFFFu);
          }
        }
      }
      if (ntr + nf > hcap) {
        err = -1;
        ntr = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int g = 0; g < 4; g++) kind[g] = 0;
      }
      // ---- trip 4 (overlaps the computation of the trial times): the chains and the last slot
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int g = 0; g < 4; g++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < kGC; c++) {
          const int sl = q[g] >> c;
          pre[g][c].x = 0;
          pre[g][c].y = -1;
          if (kind[g] != 0 && sl >= kSm && sl <= hcap) pre[g][c] = m.hget(sl);
        }
      }
      if (ntr >= 1) last = m.hget(ntr);
    }
    // ---- the four trial times (computed elsewhere while the heap was sifted)
    float tv[4];
    m.collect(tv);
    // ---- apply in the reference's order x-1, x+1, z-1, z+1 (:419-440).  Every lane walks the same four
    // steps and re-converges (m.sync) before each: otherwise lanes whose neighbour needs nothing run ahead,
    // the 32 sweeps drift apart and every path is replayed for 2-3 lanes at a time (profiles/r02_*).
    bool refresh[4] = {false, false, false, false};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < 4; g++) {
      m.sync();
      if (kind[g] != 0 && err == 0) {
        const int nidx = root + ((g == 0) ? -ld : (g == 1) ? ld : (g == 2) ? -1 : 1);
        if (refresh[g]) {  // an earlier sift-up pushed this neighbour down to slot q[g]: its chain is new
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
          for (int c = 0; c < kGC; c++) {
            const int sl = q[g] >> c;
            if (sl >= kSm && sl <= hcap) pre[g][c] = m.hget(sl);
          }
        }
        auto ent_at = [&](int sl, int c) -> Ent {  // chain slot sl = q[g] >> c
          if (sl < kSm || c >= kGC) return m.hget(sl);
          Ent e = pre[g][0];  // g is a compile-time constant after unrolling; select instead of indexing by c
          if (c == 1) e = pre[g][1];
          if (c == 2) e = pre[g][2];
          if (c == 3) e = pre[g][3];
          if (c == 4) e = pre[g][4];
          if (c == 5) e = pre[g][5];
          return e;
        };
        auto put = [&](int sl, Ent e) {  // every heap write of the apply phase
          m.hset(sl, e);
          if (sl == ntr) last = e;
          if (sl >= kSm) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int g2 = g + 1; g2 < 4; g2++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
              for (int c2 = 0; c2 < kGC; c2++)
                if ((q[g2] >> c2) == sl) pre[g2][c2] = e;
            }
          }
        };
        int tpc, c;  // slot of the entry and chain index of its parent
        if (kind[g] == 1) {  // addtree (:768-805)
          ntr = ntr + 1;
          tpc = ntr;
          c = 0;
        } else {  // updtree (:894-921): locate the entry on the ancestor chain of the stored slot
          m.stat(1, 1);
          tpc = q[g];
          c = 0;
          while (tpc > 0) {
            m.stat(0, 1);
            if (tpc <= ntr && ent_at(tpc, c).y == nidx) break;
            tpc >>= 1;
            c++;
          }
          if (tpc == 0) err = -2;
          c++;
        }
        if (tpc > 0) {
          const float t = tv[g];
          bool moved = false;
          int tpp = tpc >> 1;
          while (tpp > 0) {
            const Ent pe = ent_at(tpp, c);
            if (t < bits2f(pe.x)) {
              put(tpc, pe);
              m.set_word(pe.y, kCloseBit | (uint32_t)tpc);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
              for (int g2 = g + 1; g2 < 4; g2++) {  // a later neighbour pushed down: its exact slot is tpc now
                const int n2 = root + ((g2 == 0) ? -ld : (g2 == 1) ? ld : (g2 == 2) ? -1 : 1);
                if (kind[g2] == 2 && pe.y == n2) {
                  q[g2] = tpc;
                  refresh[g2] = true;
                }
              }
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
          put(tpc, ne);
          if (kind[g] == 1 || moved) m.set_word(nidx, kCloseBit | (uint32_t)tpc);
          if (moved) m.stat(3, 1);
        }
      }
    }
    m.sync();
    if (err) ntr = 0;
    lastOK = ntr > 0;
  }
  return err;
}

}  // namespace lps
}  // namespace dsurf
