// K3-FIM -- the coarse-grid pass of travel (src/CalSurfG.f90:386-486, fouds2 :587-759) as a BLOCK-LEVEL
// FAST ITERATIVE sweep: no heap, no pop order; one warp per sweep relaxes 32 x 32 node tiles in shared memory
// and tiles re-activate their neighbours until nothing changes.  Shared by the device kernel (eikonal.cu,
// k_fim_march) and the host replay (tests/host/fim_host_check.cpp), which runs the same per-node code against the
// oracle's heap march.
//
// WHAT IS ITERATED.  In the reference a node's final time is the LAST trial value fouds2 wrote before the node was
// popped, and a trial value is recomputed whenever one of the node's four neighbours is popped, from the nodes alive
// at that moment.  As long as nodes are popped in increasing time, that history can be read off the node's stencil:
//     v = the node's injected time if it starts as a close node (travel(urg=2) seeds, :341-347), else +inf
//     for its neighbours J that are popped during this pass, in increasing T(J):
//         if T(J) < v:   v = fouds2(node | alive = nodes alive before the pass + nodes with T <= T(J))   else stop
// rule_generic() below evaluates exactly this with the reference's fp32 operation order (lps::fouds2_words, the routine
// the exact kernels use); rule() is the same function in straight-line code for the common case.  The rule is causal -- v
// depends only on nodes with smaller times -- so wherever trial values only decrease it has ONE fixed point and any
// relaxation order reaches it (in the rare non-monotone spots the order can matter in the last bits: 6 nodes of 2.1e6
// between two tile orders on the host replay; the device order is fixed, one warp per sweep).  Wherever the reference's
// heap pops in time order the fixed point is the reference's travel-time field BIT FOR BIT.  The heap deviates from time order only (a) between equal keys (heap
// layout decides) and (b) after updtree raised a key (it only sifts up, :894-921); both are rare and local, and their
// effect is at the last-bit level.  Measured (profiles/r02_fim_parity.md; scripts/fim_parity.py on the GPU, tests/
// test_fim_host.py on the host): grids up to 257^2 -- travel-time fields, G matrices and the Vs model after an outer iteration
// bit-identical to the exact kernel's on the Taipei example, cfg 2 and a 4-type problem; 1025^2 -- 1.5 % of the nodes differ,
// by <= 2.2e-6 relative (p50 3.4e-7), predicted times by <= 1.3e-6, and 0.7 % of the rays differ in their G entries (0.1 % because
// a ray point falls into another B-spline cell: a last-bit change of T is 1e-3 of a cell's traversal time there).  That is why this pipeline is opt-in.  The refined source
// grid (early exit: its alive SET depends on the pop order) stays on the exact heap kernel (k_refine).
//
// RELAXATION ORDER.  Node-level dirty bits: a node is re-evaluated only after a stencil neighbour changed in a way
// that can matter (min(old, new) < the node's time: later nodes never enter the rule).  Inside a tile the warp walks
// anti-diagonals (lane = tile row), first in the direction pointing away from the source, so one walk usually
// settles the tile (Gauss-Seidel along the characteristics); diagonals without dirty nodes cost one ballot.
#pragma once
#include "eik_lps.cuh"

namespace dsurf {
namespace fim {

constexpr int kT = 32;                 // tile edge (nodes)
constexpr int kHX = 2;                 // halo rows on each side in x (the stencil reaches 2 nodes)
constexpr int kHZ = 4;                 // halo columns in z: 4, so that every tile row starts on a 16-byte boundary
constexpr int kRows = kT + 2 * kHX;    // 36
constexpr int kPitch = kT + 2 * kHZ;   // 40 words per tile row: lane l on an anti-diagonal hits bank (7 l + d) mod 32
constexpr uint32_t kFarG = 0xFFFFFFFFu;  // global array: far node / outside the grid
constexpr uint32_t kInf = 0x7f800000u;   // tile copy: far (+inf)
constexpr uint32_t kInit = 0x80000000u;  // tile copy: flag "alive before the pass" (times are never negative)
constexpr int kTileWords = kRows * kPitch;

// global layout of one sweep's time field: node (ix, iz) 0-based at (ix + kHX) * pitch + iz + kHZ; rows
// -kHX .. ntx*kT + kHX - 1 and columns -kHZ .. ntz*kT + kHZ - 1 exist, everything outside the grid is kFarG.
struct Layout {
  int ntx, ntz;   // tiles per direction
  int pitch;      // words per row, multiple of 4
  int rows;
  LPS_HD size_t words() const { return (size_t)rows * pitch; }
  LPS_HD size_t at(int ix, int iz) const { return (size_t)(ix + kHX) * pitch + (iz + kHZ); }
};
LPS_HD Layout make_layout(int nnx, int nnz) {
  Layout L;
  L.ntx = (nnx + kT - 1) / kT;
  L.ntz = (nnz + kT - 1) / kT;
  L.pitch = L.ntz * kT + 2 * kHZ;
  L.rows = L.ntx * kT + 2 * kHX;
  return L;
}

#if !defined(__CUDA_ARCH__) && defined(DSURF_FIM_CROSSCHECK)
static long g_cross_total = 0, g_cross_mismatch = 0, g_generic_calls = 0;
#endif
LPS_HD float as_f(uint32_t w) { return lps::bits2f((int)w); }
LPS_HD uint32_t as_u(float f) { return (uint32_t)lps::f2bits(f); }

// tile-copy word -> lps word for fouds2_words under the predicate "alive before the pass, or time <= t"
LPS_HD uint32_t alive_word(uint32_t w, float t) {
  const bool al = ((int)w < 0) || (as_f(w) <= t);
  return al ? (w & 0x7FFFFFFFu) : lps::kFar;
}

// The causal rule for the node whose tile-copy word is *c (see the header), evaluated the plain way: one complete
// fouds2 per neighbour that is popped before the node.  seed: injected time of a close node, else +inf.  in[0..3]: the
// neighbour ix-1 / ix+1 / iz-1 / iz+1 lies inside the grid (the reference skips the side otherwise, :604-605).
// Returns the node's time (+inf: nothing reaches it yet).  rule() below is the same function, faster.
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
float rule_generic(const uint32_t *c, float seed, float slown, float ri, float risti, float dnx, float dnz, bool in0, bool in1,
                   bool in2, bool in3) {
  const bool in[4] = {in0, in1, in2, in3};
  const uint32_t w1[4] = {c[-kPitch], c[kPitch], c[-1], c[1]};
  const uint32_t w2[4] = {c[-2 * kPitch], c[2 * kPitch], c[-2], c[2]};
  // neighbours popped during the pass: inside the grid, not alive before it, reached (finite)
  float t0 = (in[0] && w1[0] < kInf) ? as_f(w1[0]) : as_f(kInf);
  float t1 = (in[1] && w1[1] < kInf) ? as_f(w1[1]) : as_f(kInf);
  float t2 = (in[2] && w1[2] < kInf) ? as_f(w1[2]) : as_f(kInf);
  float t3 = (in[3] && w1[3] < kInf) ? as_f(w1[3]) : as_f(kInf);
#define DSURF_FIM_CX(a, b) \
  {                        \
    const float lo = a < b ? a : b, hi = a < b ? b : a; \
    a = lo;                \
    b = hi;                \
  }
  DSURF_FIM_CX(t0, t1) DSURF_FIM_CX(t2, t3) DSURF_FIM_CX(t0, t2) DSURF_FIM_CX(t1, t3) DSURF_FIM_CX(t1, t2)
#undef DSURF_FIM_CX
  const float ts[4] = {t0, t1, t2, t3};
  const bool inj[2] = {in[0], in[1]}, ink[2] = {in[2], in[3]};
  float v = seed;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k = 0; k < 4; k++) {
    const float t = ts[k];
    if (!(t < v)) break;
    if (k < 3 && ts[k + 1] == t) continue;  // equal neighbours are popped back to back: one evaluation with both alive
    uint32_t wj1[2], wj2[2], wk1[2], wk2[2];
    wj1[0] = alive_word(w1[0], t);
    wj1[1] = alive_word(w1[1], t);
    wk1[0] = alive_word(w1[2], t);
    wk1[1] = alive_word(w1[3], t);
    wj2[0] = alive_word(w2[0], t);
    wj2[1] = alive_word(w2[1], t);
    wk2[0] = alive_word(w2[2], t);
    wk2[1] = alive_word(w2[3], t);
    v = lps::fouds2_words(wj1, wj2, inj, wk1, wk2, ink, slown, ri, risti, dnx, dnz);
  }
  return v;
}

// The same rule for the COMMON case, in straight-line code.  fouds2's result for an alive set A is the minimum over
// the (up to four) quadrants (x side, z side) of: the two-leg solution if both sides are alive, the one-leg solution of
// the alive side if only one is.  A leg is (T1, T2, second-order flag) of one side; second order needs the 2-away node
// alive and T1 > T2 (:620-663), and a 2-away node EARLIER than its 1-away node is alive whenever that one is, so a leg --
// and with it every quadrant solution -- does not depend on the moment of the evaluation.  Almost every node ends up
// with at most ONE alive side per axis (the upwind one: X1 = the earlier of ix-1 / ix+1, Z1 likewise); then the whole
// history is
//     v1 = one-leg(first of X1, Z1);   if the other one is popped before v1:   v = min(two-leg(X1, Z1), one-leg(X1) if
//     the far z side exists, one-leg(Z1) if the far x side exists)
// i.e. three quadrant solutions, each computed once (the plain form evaluates 5-7 quadrants).  Everything else -- a
// second side of an axis popped in time (colliding fronts), equal times on one axis, seeds and nodes alive before the
// pass (start-up region) -- is detected and handed to rule_generic.  Both forms call lps::quad on the same operands,
// so they agree bit for bit; the host replay (tests/host/fim_host_check.cpp) compares them on every evaluation.
LPS_HD float rule(const uint32_t *c, float seed, float slown, float ri, float risti, float dnx, float dnz, const bool in[4]) {
  const uint32_t wxm = c[-kPitch], wxp = c[kPitch], wzm = c[-1], wzp = c[1];
  const float inf = as_f(kInf);
  // popped neighbours: inside the grid, not alive before the pass (sign flag), reached (finite)
  const float txm = (in[0] && wxm < kInf) ? as_f(wxm) : inf, txp = (in[1] && wxp < kInf) ? as_f(wxp) : inf;
  const float tzm = (in[2] && wzm < kInf) ? as_f(wzm) : inf, tzp = (in[3] && wzp < kInf) ? as_f(wzp) : inf;
  const bool anyinit = (in[0] && (int)wxm < 0) || (in[1] && (int)wxp < 0) || (in[2] && (int)wzm < 0) || (in[3] && (int)wzp < 0);
  const bool xp = txp < txm, zp = tzp < tzm;                 // upwind side of each axis
  const float tx = xp ? txp : txm, tx2 = xp ? txm : txp;     // its time, the far side's time
  const float tz = zp ? tzp : tzm, tz2 = zp ? tzm : tzp;
  bool generic = anyinit || seed < inf || (tx2 == tx && tx < inf) || (tz2 == tz && tz < inf);
  float v = inf;
  if (!generic && (tx < inf || tz < inf)) {
    // legs of the two upwind sides
    const uint32_t wx2 = xp ? c[2 * kPitch] : c[-2 * kPitch], wz2 = zp ? c[2] : c[-2];
    const float Tx2 = as_f(wx2 & 0x7FFFFFFFu), Tz2 = as_f(wz2 & 0x7FFFFFFFu);
    const bool sox = tx > Tx2, soz = tz > Tz2;  // 2-away node reached (or alive before the pass) and earlier
    const bool xfar_in = xp ? in[0] : in[1], zfar_in = zp ? in[2] : in[3];  // the far side of each axis lies inside the grid
    float q1x = inf, q1z = inf, q2 = inf;
    if (tx < inf) lps::quad(tx, Tx2, true, sox, 0.0f, 0.0f, false, false, slown, ri, risti, dnx, dnz, q1x);
    if (tz < inf) lps::quad(0.0f, 0.0f, false, false, tz, Tz2, true, soz, slown, ri, risti, dnx, dnz, q1z);
    const bool xfirst = tx < tz, tie = tx == tz;
    const float v1 = xfirst ? q1x : q1z;          // after the first pop (not evaluated on a tie)
    const float tb = xfirst ? tz : tx;            // the other axis' upwind side ...
    const float toa = xfirst ? tx2 : tz2;         // ... and the first axis' far side: whichever is earlier comes next
    if (!tie && toa <= tb && toa < v1) generic = true;  // the far side of the first axis is popped in time (or ties with tb)
    const bool both = tie || (tb < toa && tb < v1);
    if (!tie && !generic && !both) v = v1;
    if (both && !generic) {
      lps::quad(tx, Tx2, true, sox, tz, Tz2, true, soz, slown, ri, risti, dnx, dnz, q2);
      float v2 = q2;
      if (zfar_in) v2 = q1x < v2 ? q1x : v2;
      if (xfar_in) v2 = q1z < v2 ? q1z : v2;
      const float t3 = tx2 < tz2 ? tx2 : tz2;
      if (t3 < v2) generic = true;  // a third side is popped in time
      v = v2;
    }
  }
#if !defined(__CUDA_ARCH__) && defined(DSURF_FIM_CROSSCHECK)
  if (generic) g_generic_calls++;
#endif
  if (generic) return rule_generic(c, seed, slown, ri, risti, dnx, dnz, in[0], in[1], in[2], in[3]);
  return v;
}

// ---- per-warp working set of a tile: host replay (with a slowness copy); the device keeps TileD in shared memory
struct Tile {
  uint32_t t[kTileWords];  // times: rows -kHX..kT+kHX-1, columns -kHZ..kT+kHZ-1
  float slow[kT * kT];     // 1 / velocity, [x][z]
  float risti[kT];         // earth * sin(colatitude of column x)  (host libm table)
  uint32_t dirty[kT];      // word x: bit z = node (x, z) must be re-evaluated
  uint32_t hx[4];          // marks for the halo rows x = -2, -1, kT, kT+1 (bit z): nodes of the x-neighbour tiles
  uint32_t hz[4];          // marks for the halo columns z = -2, -1, kT, kT+1 (bit x)
  LPS_HD uint32_t *at(int x, int z) { return t + (x + kHX) * kPitch + (z + kHZ); }
};

struct TileD {             // 6048 bytes: 24-32 warps (= sweeps) per SM
  uint32_t t[kTileWords];
  float risti[kT];
  uint32_t dirty[kT];
  uint32_t hx[4];
  uint32_t hz[4];
  LPS_HD uint32_t *at(int x, int z) { return t + (x + kHX) * kPitch + (z + kHZ); }
};

// what the tile needs to know about its place in the sweep
struct TileCtx {
  int gx0, gz0;       // grid coordinates (0-based) of tile node (0, 0)
  int nnx, nnz;
  float ri, dnx, dnz;
  // injected close nodes (seeds) live in the refined box [bx0, bx0 + bw) x [bz0, bz0 + bh) of the coarse grid;
  // box[(ix - bx0) * bh + (iz - bz0)] = (time bits, status): status > 0 or -100 = close (CalSurfG.f90:1332-1349)
  int bx0, bz0, bw, bh;
  const int *box;  // pairs (x = time bits, y = status), int2-compatible
};

LPS_HD float seed_of(const TileCtx &C, int gx, int gz) {
  const int bx = gx - C.bx0, bz = gz - C.bz0;
  if (bx < 0 || bx >= C.bw || bz < 0 || bz >= C.bh) return as_f(kInf);
  const int st = C.box[2 * (bx * C.bh + bz) + 1];
  return (st > 0 || st == -100) ? lps::bits2f(C.box[2 * (bx * C.bh + bz)]) : as_f(kInf);
}

#if defined(__CUDA_ARCH__)
#define DSURF_FIM_OR(p, v) atomicOr((p), (v))
#define DSURF_FIM_AND(p, v) atomicAnd((p), (v))
#else
#define DSURF_FIM_OR(p, v) (*(p) |= (v))
#define DSURF_FIM_AND(p, v) (*(p) &= (v))
#endif

// Re-evaluates tile node (x, z) (its dirty bit is set): clears the bit, applies the rule, and if the time changed stores
// it and marks the stencil users that can see the change.  Lanes of one anti-diagonal call this concurrently: their
// nodes are never in each other's stencils.  Returns true if the time changed.
template <class TL>
LPS_HD bool relax_node(TL &tl, const TileCtx &C, int x, int z, float slown) {
  DSURF_FIM_AND(&tl.dirty[x], ~(1u << z));
  uint32_t *c = tl.at(x, z);
  const uint32_t old = *c;
  if ((int)old < 0) return false;  // alive before the pass: never recomputed
  const int gx = C.gx0 + x, gz = C.gz0 + z;
  const bool in[4] = {gx - 1 >= 0, gx + 1 < C.nnx, gz - 1 >= 0, gz + 1 < C.nnz};
  const float v = rule(c, seed_of(C, gx, gz), slown, C.ri, tl.risti[x], C.dnx, C.dnz, in);
#if !defined(__CUDA_ARCH__) && defined(DSURF_FIM_CROSSCHECK)
  {  // host replay: the cached-quadrant rule against the plain one, on every evaluation
    const float vg = rule_generic(c, seed_of(C, gx, gz), slown, C.ri, tl.risti[x], C.dnx, C.dnz, in[0], in[1], in[2], in[3]);
    g_cross_total++;
    if (as_u(vg) != as_u(v)) g_cross_mismatch++;
  }
#endif
  const uint32_t nw = as_u(v);
  if (nw == old) return false;
  *c = nw;
  const float mn = v < as_f(old) ? v : as_f(old);
  // users: the 8 nodes whose stencil holds this node.  A user earlier than both values never looks at it (the rule only
  // takes nodes popped before the user); users alive before the pass are never recomputed.
  unsigned zmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; q++) {
    const int dz = q == 0 ? -2 : q == 1 ? -1 : q == 2 ? 1 : 2;
    const int uz = z + dz, ugz = gz + dz;
    const uint32_t uw = c[dz];
    if (ugz < 0 || ugz >= C.nnz || (int)uw < 0 || mn >= as_f(uw)) continue;
    if (uz < 0)
      DSURF_FIM_OR(&tl.hz[uz + 2], 1u << x);
    else if (uz >= kT)
      DSURF_FIM_OR(&tl.hz[uz - kT + 2], 1u << x);
    else
      zmask |= 1u << uz;
  }
  if (zmask) DSURF_FIM_OR(&tl.dirty[x], zmask);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; q++) {
    const int dx = q == 0 ? -2 : q == 1 ? -1 : q == 2 ? 1 : 2;
    const int ux = x + dx, ugx = gx + dx;
    const uint32_t uw = c[dx * kPitch];
    if (ugx < 0 || ugx >= C.nnx || (int)uw < 0 || mn >= as_f(uw)) continue;
    if (ux < 0)
      DSURF_FIM_OR(&tl.hx[ux + 2], 1u << z);
    else if (ux >= kT)
      DSURF_FIM_OR(&tl.hx[ux - kT + 2], 1u << z);
    else
      DSURF_FIM_OR(&tl.dirty[ux], 1u << z);
  }
  return true;
}

// =============================================================================================
// START-UP: the first pops of the coarse pass, marched EXACTLY (the reference's heap, :768-921) by one thread per
// sweep on a small region around the refined box.  The close nodes injected from the refined grid carry refined-grid
// times; the coarse stencil recomputes them to larger values, updtree leaves those raised keys where they are, and the
// heap pops out of time order until they are gone -- with visible effects when the refined pass stopped early (a
// source next to the grid edge: tests/host/fim_host_check.cpp, case "corner").  The march stops as soon as no injected
// seed and no raised key is left in the heap (from then on pops come in time order and the rule above applies), or
// when the front gets within 3 nodes of the region's edge.  Hand-over: alive nodes become "alive before the pass",
// heap entries become seeds with their current keys.
// =============================================================================================
constexpr int kRegHalf = 20;                 // region = source cell +- kRegHalf nodes, clipped to the grid
constexpr int kRegMax = 2 * kRegHalf + 2;    // 42 nodes per direction at most
constexpr int kRegNodes = kRegMax * kRegMax;

struct StartCtx {
  int nnx, nnz;            // coarse grid
  int rx0, rz0, rw, rh;    // region rectangle (0-based grid coordinates of its first node, extent)
  float ri, dnx, dnz;
  const float *vel;        // coarse velocity, [nnx][nnz]
  const float *risti;      // [nnx]
};

struct StartMem {
  uint32_t *w;       // [rw * rh] region words (x * rh + z): alive = time, close = kCloseBit | heap slot, far = kFar
  lps::Ent *heap;    // [1 .. rw * rh] (key bits, region node)
  unsigned char *flag;  // [rw * rh] bit 0: injected seed, bit 1: key was raised while in the heap
};

LPS_HD void start_sift_up(const StartMem &m, int tpc, float key, int node) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const lps::Ent pe = m.heap[tpp];
    if (key < lps::bits2f(pe.x)) {
      m.heap[tpc] = pe;
      m.w[pe.y] = lps::kCloseBit | (uint32_t)tpc;
      tpc = tpp;
      tpp = tpc >> 1;
    } else {
      tpp = 0;
    }
  }
  lps::Ent ne;
  ne.x = lps::f2bits(key);
  ne.y = node;
  m.heap[tpc] = ne;
  m.w[node] = lps::kCloseBit | (uint32_t)tpc;
}

// box: (time bits, status) pairs of the refined box [bx0, bx0 + bw) x [bz0, bz0 + bh) as k_refine leaves them.
// Returns the number of pops; on return m.w / m.heap[1 .. ntr] describe the hand-over state.
LPS_HD int startup_march(const StartCtx &C, const StartMem &m, const int *box, int bx0, int bz0, int bw, int bh, int &ntr_out,
                         int max_pops) {
  const int rh = C.rh;
  for (int i = 0; i < C.rw * C.rh; i++) {
    m.w[i] = lps::kFar;
    m.flag[i] = 0;
  }
  int ntr = 0, seeds_left = 0, viol = 0, pops = 0;
  for (int bx = 0; bx < bw; bx++)       // travel(urg=2) scans ix outer, iz inner (:341-347)
    for (int bz = 0; bz < bh; bz++) {
      const int st = box[2 * (bx * bh + bz) + 1];
      const int node = (bx0 + bx - C.rx0) * rh + (bz0 + bz - C.rz0);
      if (st == 0) {
        m.w[node] = (uint32_t)box[2 * (bx * bh + bz)];
      } else if (st > 0 || st == -100) {
        ntr++;
        start_sift_up(m, ntr, lps::bits2f(box[2 * (bx * bh + bz)]), node);
        m.flag[node] = 1;
        seeds_left++;
      }
    }
  while (ntr > 0 && pops < max_pops) {
    if (seeds_left == 0 && viol == 0) break;
    const lps::Ent r = m.heap[1];
    const int x = r.y / rh, z = r.y - x * rh;
    // every stencil of the four neighbours must lie inside the region (or outside the grid)
    if ((x - 3 < 0 && C.rx0 > 0) || (x + 3 >= C.rw && C.rx0 + C.rw < C.nnx) || (z - 3 < 0 && C.rz0 > 0) ||
        (z + 3 >= C.rh && C.rz0 + C.rh < C.nnz))
      break;
    pops++;
    m.w[r.y] = (uint32_t)r.x;  // alive, time = key (:415-417)
    if (m.flag[r.y] & 1) seeds_left--;
    if (m.flag[r.y] & 2) viol--;
    // ---- downtree (:816-885)
    if (ntr == 1) {
      ntr = 0;
    } else {
      const lps::Ent last = m.heap[ntr];
      const float mk = lps::bits2f(last.x);
      ntr--;
      int tpp = 1, tpc = 2;
      while (tpc <= ntr) {
        lps::Ent ec = m.heap[tpc];
        if (tpc < ntr) {
          const lps::Ent e2 = m.heap[tpc + 1];
          if (lps::bits2f(ec.x) > lps::bits2f(e2.x)) {
            tpc++;
            ec = e2;
          }
        }
        if (!(lps::bits2f(ec.x) < mk)) break;
        m.heap[tpp] = ec;
        m.w[ec.y] = lps::kCloseBit | (uint32_t)tpp;
        tpp = tpc;
        tpc = 2 * tpp;
      }
      m.heap[tpp] = last;
      m.w[last.y] = lps::kCloseBit | (uint32_t)tpp;
    }
    // ---- the four neighbours in the reference's order (:419-440)
    for (int g = 0; g < 4; g++) {
      const int nx = x + (g == 0 ? -1 : g == 1 ? 1 : 0), nz = z + (g == 2 ? -1 : g == 3 ? 1 : 0);
      const int gx = C.rx0 + nx, gz = C.rz0 + nz;
      if (gx < 0 || gx >= C.nnx || gz < 0 || gz >= C.nnz) continue;
      const int node = nx * rh + nz;
      const uint32_t wn = m.w[node];
      if (lps::alive(wn)) continue;
      uint32_t wj1[2], wj2[2], wk1[2], wk2[2];
      bool inj[2], ink[2];
      for (int sd = 0; sd < 2; sd++) {
        const int sg = sd ? 1 : -1;
        inj[sd] = gx + sg >= 0 && gx + sg < C.nnx;
        ink[sd] = gz + sg >= 0 && gz + sg < C.nnz;
        const int x1 = nx + sg, x2 = nx + 2 * sg, z1 = nz + sg, z2 = nz + 2 * sg;
        wj1[sd] = (x1 >= 0 && x1 < C.rw) ? m.w[x1 * rh + nz] : lps::kFar;
        wj2[sd] = (x2 >= 0 && x2 < C.rw) ? m.w[x2 * rh + nz] : lps::kFar;
        wk1[sd] = (z1 >= 0 && z1 < C.rh) ? m.w[nx * rh + z1] : lps::kFar;
        wk2[sd] = (z2 >= 0 && z2 < C.rh) ? m.w[nx * rh + z2] : lps::kFar;
      }
      const float tv = lps::fouds2_words(wj1, wj2, inj, wk1, wk2, ink, 1.0f / C.vel[(size_t)gx * C.nnz + gz], C.ri, C.risti[gx],
                                         C.dnx, C.dnz);
      if (wn == lps::kFar) {
        ntr++;
        start_sift_up(m, ntr, tv, node);
      } else {
        const int slot = (int)(wn & 0x7FFFFFFFu);
        if (tv > lps::bits2f(m.heap[slot].x) && !(m.flag[node] & 2)) {
          m.flag[node] |= 2;
          viol++;
        }
        start_sift_up(m, slot, tv, node);  // updtree: sifts up only, even when the key was raised (:894-921)
      }
    }
  }
  ntr_out = ntr;
  return pops;
}

// node handled by tile row x on anti-diagonal d of a walk in direction (sx, sz); -1 if none
LPS_HD int diag_z(int x, int d, int sx, int sz) {
  const int xs = sx > 0 ? x : kT - 1 - x;
  const int zs = d - xs;
  if (zs < 0 || zs >= kT) return -1;
  return sz > 0 ? zs : kT - 1 - zs;
}

}  // namespace fim
}  // namespace dsurf
