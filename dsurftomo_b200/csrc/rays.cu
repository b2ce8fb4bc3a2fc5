// K4/K5/K6 -- receiver travel times, ray back-tracing with Frechet accumulation, and assembly
// of the sparse sensitivity rows, replacing srtimes (src/CalSurfG.f90:1636-1759), rpaths
// (:1771-2318), bilinear (:2328-2349) and the row assembly of CalSurfG (:1383-1432).
//
// K4/K5: one THREAD per ray (a ray is an inherently serial gradient descent).  The 16 B-spline
// vertex accumulators of the current vertex cell are kept in registers and spilled to the ray's
// dense fdm(0:nvz+1,0:nvx+1) slab only when the ray enters another vertex cell, so the fp32
// accumulation order per vertex is exactly the reference's step order.  sin(colatitude) at ray
// points restates glibc's sinf algorithm (see sin_cr); all other arithmetic is fp32 in source order.
// K6: ordered compaction: rows come out ascending in column (depth-major, then z, then x
// vertex) exactly as the reference's scan nn = 1..nparpi emits them.
// Bound: L2 gather latency (rays), HBM streaming of sen/fdm/CSR (assembly).
#include <cub/cub.cuh>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "plan.cuh"

namespace dsurf {

__device__ __forceinline__ float cube_r(float x) { return x * (x * x); }
__device__ __forceinline__ void bsp4(float u, float o[4]) {  // CalSurfG.f90:2180-2187
  o[0] = cube_r(1.0f - u) / 6.0f;
  o[1] = (4.0f - 6.0f * (u * u) + 3.0f * cube_r(u)) / 6.0f;
  o[2] = (1.0f + 3.0f * u + 3.0f * (u * u) - 3.0f * cube_r(u)) / 6.0f;
  o[3] = cube_r(u) / 6.0f;
}
// REAL*4 SIN as the reference's runtime evaluates it.  gfortran's SIN(real(4)) is glibc's sinf,
// which is NOT correctly rounded (0.85 % of inputs differ from the rounded fp64 sine), so the
// published algorithm glibc >= 2.28 uses (Arm Optimized Routines sinf: fp64 range reduction by
// pi/2 and two short fp64 polynomials) is restated here; on the build host it agrees with
// sinf() bit for bit on all 7.97e7 floats of [2^-7, 6) (checked exhaustively).
// Evaluated in fp64 without FMA contraction (--fmad=false), like glibc's generic build.
__device__ __forceinline__ float sinf_poly(double x, double x2, int tab, int n) {
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double s1 = S2 + x2 * S3;
    const double x7 = x3 * x2;
    const double s = x + x3 * S1;
    return (float)(s + x7 * s1);
  }
  const double sg = tab ? -1.0 : 1.0;
  const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5,
               C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
  const double x4 = x2 * x2;
  const double c2 = C3 + x2 * C4;
  const double c1 = C0 + x2 * C1;
  const double x6 = x4 * x2;
  const double c = c1 + x4 * C2;
  return (float)(c + x6 * c2);
}
__device__ __forceinline__ float sin_cr(float y) {
  const double x = (double)y;
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
  if (top < (0x3f490fdbu >> 20)) {  // |y| < pi/4
    if (top < (0x39800000u >> 20)) return y;
    return sinf_poly(x, x * x, 0, 0);
  }
  if (top < (0x42f00000u >> 20)) {  // |y| < 120
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double r = x * hpi_inv;
    const int n = ((int)r + 0x800000) >> 24;
    const double xr = x - (double)n * hpi;
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sinf_poly(xr * sgn, xr * xr, (n & 2) ? 1 : 0, n);
  }
  return (float)sin(x);  // outside any colatitude; not reached
}

__device__ __forceinline__ float node_t(const int2 *n, size_t i) { return __int_as_float(n[i].x); }
// coarse travel times: one word per node (eik_lps.cuh; every node is alive after the sweep, word = time)
// or the legacy packed (time, status) records
struct CoarseT {  // i = (ix - 1) * nnz + (iz - 1); the word array is padded (BatchView::wld, wpx, wpz)
  const int2 *n;
  const unsigned *w;
  int nnz, wld, wpx, wpz;
  __device__ __forceinline__ float t(size_t i) const {
    if (!w) return __int_as_float(n[i].x);
    const size_t x = i / (size_t)nnz, z = i - x * (size_t)nnz;
    return __uint_as_float(w[(x + wpx) * (size_t)wld + z + wpz]);
  }
};

// bilinear interpolation of velocity at (drx, drz) inside cell (ipz, ipx), guards as :2153-2157
__device__ __forceinline__ float vel_at(const float *veln, int nnx, int nnz, int ipx, int ipz, float drx,
                                        float drz, float dnx, float dnz, bool xz_order) {
  float vel = 0.0f;
  // first point (:2150-2159) loops l (x) outer, m (z) inner; later points (:2213-2221) loop
  // m (x) outer, n (z) inner -- the same order.
  for (int l = 0; l < 2; l++)
    for (int m = 0; m < 2; m++) {
      float produ = (1.0f - fabsf(((float)m * dnz - drz) / dnz));
      produ = produ * (1.0f - fabsf(((float)l * dnx - drx) / dnx));
      const int iz = ipz + m, ix = ipx + l;
      if (iz <= nnz && ix <= nnx && iz >= 1 && ix >= 1) vel = vel + veln[(size_t)(ix - 1) * nnz + (iz - 1)] * produ;
    }
  (void)xz_order;
  return vel;
}

__global__ void __launch_bounds__(128)
k_rays(Geom g, const SweepDesc *__restrict__ sw, const RayDesc *__restrict__ rays, int nrays,
       const float *__restrict__ veln_all, BatchView bv, float *__restrict__ tt_out,
       float *__restrict__ fdm_all, int4 *__restrict__ bbox, int *__restrict__ rbint_flag,
       int *__restrict__ err_flag, float2 *__restrict__ path, int *__restrict__ path_n, int path_cap) {
  const int rid = blockIdx.x * blockDim.x + threadIdx.x;
  if (rid >= nrays) return;
  // optional ray-geometry export (rgx/rgz of rpaths, the reference's raypath.out block :2276-2283)
  int np = 0;
  auto rec = [&](float x, float z) {
    if (path && np < path_cap) path[(size_t)rid * path_cap + np] = make_float2(x, z);
    np++;
  };
  if (path_n) path_n[rid] = 0;
  const RayDesc rd = rays[rid];
  const SweepDesc d = sw[rd.sweep];
  bbox[rid] = make_int4(1, 0, 1, 0);  // empty until the ray has been traced
  const size_t Nc = (size_t)g.nnx * g.nnz;
  const float *veln = veln_all + (size_t)d.map * Nc;
  CoarseT node;
  node.n = bv.word ? nullptr : bv.node + (size_t)rd.sweep * Nc;
  node.w = bv.word ? bv.word + (size_t)rd.sweep * bv.wslot : nullptr;
  node.nnz = g.nnz;
  node.wld = bv.wld;
  node.wpx = bv.wpx;
  node.wpz = bv.wpz;
  const int2 *noder = bv.noder + (size_t)rd.sweep * kRefMax * kRefMax;
  const int nnx = g.nnx, nnz = g.nnz, nnxr = d.nrnx, nnzr = d.nrnz;
  const float gox = g.gox, goz = g.goz, dnx = g.dnx, dnz = g.dnz, dvx = g.dvx, dvz = g.dvz;
  const float goxr = d.gorx, gozr = d.gorz, dnxr = g.drnx, dnzr = g.drnz, earth = g.earth;
  const float scx = d.scx, scz = d.scz;
  // ------------------------------------------------------------------ srtimes (:1636-1759)
  if (d.do_times) {
    int irx = (int)((rd.rcx - gox) / dnx) + 1;
    int irz = (int)((rd.rcz - goz) / dnz) + 1;
    if (irx < 1 || irx > nnx || irz < 1 || irz > nnz) {
      atomicExch(err_flag, DSURF_ERR_RECEIVER_OUTSIDE);
      return;
    }
    if (irx == nnx) irx = irx - 1;
    if (irz == nnz) irz = irz - 1;
    const int isx = (int)((scx - gox) / dnx) + 1;
    const int isz = (int)((scz - goz) / dnz) + 1;
    const float dpl = g.dpl_sr;
    const float e1 = (scx - rd.rcx) * earth;
    float sred = e1 * e1;
    const float e2 = (scz - rd.rcz) * earth * rd.sin_rcx;
    sred = sred + e2 * e2;
    sred = sqrtf(sred);
    int sw1 = 0;
    if (sred < dpl) sw1 = 1;
    if (isx == irx && isz == irz) sw1 = 1;
    float trr;
    if (sw1) {
      float drx = (scx - gox) - (float)(isx - 1) * dnx;
      float drz = (scz - goz) - (float)(isz - 1) * dnz;
      float vels = 0.0f, velr = 0.0f;
      for (int k = 0; k < 2; k++)
        for (int l = 0; l < 2; l++) {
          const float produ = (1.0f - fabsf(((float)k * dnx - drx) / dnx)) * (1.0f - fabsf(((float)l * dnz - drz) / dnz));
          vels = vels + veln[(size_t)(isx - 1 + k) * nnz + (isz - 1 + l)] * produ;
        }
      drx = (rd.rcx - gox) - (float)(irx - 1) * dnx;
      drz = (rd.rcz - goz) - (float)(irz - 1) * dnz;
      for (int k = 0; k < 2; k++)
        for (int l = 0; l < 2; l++) {
          const float produ = (1.0f - fabsf(((float)k * dnx - drx) / dnx)) * (1.0f - fabsf(((float)l * dnz - drz) / dnz));
          velr = velr + veln[(size_t)(irx - 1 + k) * nnz + (irz - 1 + l)] * produ;
        }
      trr = 2.0f * sred / (vels + velr);
    } else {
      const float drx = (rd.rcx - gox) - (float)(irx - 1) * dnx;
      const float drz = (rd.rcz - goz) - (float)(irz - 1) * dnz;
      trr = 0.0f;
      for (int k = 0; k < 2; k++)
        for (int l = 0; l < 2; l++) {
          const float produ = (1.0f - fabsf(((float)l * dnz - drz) / dnz)) * (1.0f - fabsf(((float)k * dnx - drx) / dnx));
          trr = trr + node.t((size_t)(irx - 1 + k) * nnz + (irz - 1 + l)) * produ;
        }
    }
    tt_out[rd.row] = trr;
  }
  if (!d.do_rays) return;
  // ------------------------------------------------------------------ rpaths (:1771-2318)
  const int fld = g.nvz + 2;
  float *fdm = fdm_all + (size_t)rid * (size_t)(g.nvz + 2) * (g.nvx + 2);
  const int isx = d.rsx, isz = d.rsz;  // refined source cell, unclamped (:1853-1854)
  const float dpl = g.dpl_ray;
  int ipx = (int)((rd.rcx - gox) / dnx) + 1;
  int ipz = (int)((rd.rcz - goz) / dnz) + 1;
  if (ipx < 1 || ipx >= nnx || ipz < 1 || ipz >= nnz) {
    atomicExch(err_flag, DSURF_ERR_RECEIVER_OUTSIDE);
    return;
  }
  float rgx_j = rd.rcx, rgz_j = rd.rcz;
  rec(rgx_j, rgz_j);  // rgx(1), rgz(1) = receiver (:1910-1911)
  int sw1 = 0;
  {
    const float e1 = (scx - rgx_j) * earth;
    float sred = e1 * e1;
    const float e2 = (scz - rgz_j) * earth * rd.sin_rcx;
    sred = sred + e2 * e2;
    sred = sqrtf(sred);
    if (sred < 2.0f * dpl) sw1 = 1;
  }
  auto alive_cell = [&](int ipxr, int ipzr) -> int {
    int igref = 1;
    if (ipxr < 1 || ipxr >= nnxr) igref = 0;
    if (ipzr < 1 || ipzr >= nnzr) igref = 0;
    if (igref == 1) {
      const size_t o = (size_t)(ipxr - 1) * nnzr + (ipzr - 1);
      if (noder[o].y != 0 || noder[o + 1].y != 0) igref = 0;
      if (noder[o + nnzr].y != 0 || noder[o + nnzr + 1].y != 0) igref = 0;
    }
    return igref;
  };
  int ipxr = (int)((rd.rcx - goxr) / dnxr) + 1;
  int ipzr = (int)((rd.rcz - gozr) / dnzr) + 1;
  int igref = alive_cell(ipxr, ipzr);
  if (sw1 == 0 && igref == 1 && ipxr == isx && ipzr == isz) sw1 = 1;
  if (sw1 == 1) rec(scx, scz);  // nrp = 2 (:1919-1921, 1949-1951)
  // register cache of the 4x4 vertex accumulators
  float acc[16];
  int civz = -1000, civx = -1000;
  int bz0 = 1 << 30, bz1 = -1, bx0 = 1 << 30, bx1 = -1;
  int rb = 0;
  const long long maxrp = (long long)nnx * nnz;
  for (long long j = 1; j <= maxrp; j++) {
    if (sw1 == 1) break;
    float dtx, dtz;
    const float sinx = sin_cr(rgx_j);
    if (igref == 1) {
      const size_t o = (size_t)(ipxr - 1) * nnzr + (ipzr - 1);
      const float t00 = node_t(noder, o), t10 = node_t(noder, o + 1);            // (ipzr,ipxr),(ipzr+1,ipxr)
      const float t01 = node_t(noder, o + nnzr), t11 = node_t(noder, o + nnzr + 1);  // (ipzr,ipxr+1),(ipzr+1,ipxr+1)
      dtx = t01 - t00;
      dtx = dtx + t11 - t10;
      dtx = dtx / (2.0f * earth * dnxr);
      dtz = t10 - t00;
      dtz = dtz + t11 - t01;
      dtz = dtz / (2.0f * earth * sinx * dnzr);
    } else {
      const size_t o = (size_t)(ipx - 1) * nnz + (ipz - 1);
      const float t00 = node.t(o), t10 = node.t(o + 1);
      const float t01 = node.t(o + nnz), t11 = node.t(o + nnz + 1);
      dtx = t01 - t00;
      dtx = dtx + t11 - t10;
      dtx = dtx / (2.0f * earth * dnx);
      dtz = t10 - t00;
      dtz = dtz + t11 - t01;
      dtz = dtz / (2.0f * earth * sinx * dnz);
    }
    float rd1 = sqrtf(dtx * dtx + dtz * dtz);
    float rgx_n = rgx_j - dpl * dtx / (earth * rd1);
    float rgz_n = rgz_j - dpl * dtz / (earth * sinx * rd1);
    const int ipxo = ipx, ipzo = ipz;
    ipxr = (int)((rgx_n - goxr) / dnxr) + 1;
    ipzr = (int)((rgz_n - gozr) / dnzr) + 1;
    igref = alive_cell(ipxr, ipzr);
    ipx = (int)((rgx_n - gox) / dnx) + 1;
    ipz = (int)((rgz_n - goz) / dnz) + 1;
    {
      const float e1 = (scx - rgx_n) * earth;
      float sred = e1 * e1;
      const float e2 = (scz - rgz_n) * earth * sin_cr(rgx_n);
      sred = sred + e2 * e2;
      sred = sqrtf(sred);
      sw1 = 0;
      if (sred < 2.0f * dpl) sw1 = 1;
    }
    if (sw1 == 0 && igref == 1 && ipxr == isx && ipzr == isz) sw1 = 1;
    if (ipx < 1) {
      rgx_n = gox;
      ipx = 1;
      rb = 1;
    }
    if (ipx >= nnx) {
      rgx_n = g.x_last;
      ipx = nnx - 1;
      rb = 1;
    }
    if (ipz < 1) {
      rgz_n = goz;
      ipz = 1;
      rb = 1;
    }
    if (ipz >= nnz) {
      rgz_n = g.z_last;
      ipz = nnz - 1;
      rb = 1;
    }
    rec(rgx_n, rgz_n);            // rgx(j+1) after the boundary clamp (:2082-2101)
    if (sw1 == 1) rec(scx, scz);  // rgx(j+2) = source, nrp = j+2 (:2042-2046, 2057-2061)
    // ---- Frechet derivatives (:2110-2265)
    const int ivx = (ipx - 1) / kGd + 1, ivz = (ipz - 1) / kGd + 1;
    const int ivxo = (ipxo - 1) / kGd + 1, ivzo = (ipzo - 1) / kGd + 1;
    int nhp = 0;
    float vrat[3];
    int chp[3];
    if (ivx != ivxo) {
      float xi;
      if (ivx > ivxo)
        xi = gox + (float)(ivx - 1) * dvx;
      else
        xi = gox + (float)ivx * dvx;
      vrat[nhp] = (xi - rgx_j) / (rgx_n - rgx_j);
      chp[nhp] = 1;
      nhp = nhp + 1;
    }
    if (ivz != ivzo) {
      float zi;
      if (ivz > ivzo)
        zi = goz + (float)(ivz - 1) * dvz;
      else
        zi = goz + (float)ivz * dvz;
      rd1 = (zi - rgz_j) / (rgz_n - rgz_j);
      if (nhp == 0) {
        vrat[0] = rd1;
        chp[0] = 2;
      } else {
        if (rd1 >= vrat[0]) {
          vrat[1] = rd1;
          chp[1] = 2;
        } else {
          vrat[1] = vrat[0];
          chp[1] = chp[0];
          vrat[0] = rd1;
          chp[0] = 2;
        }
      }
      nhp = nhp + 1;
    }
    vrat[nhp] = 1.0f;
    chp[nhp] = 0;
    nhp = nhp + 1;
    float drx = (rgx_j - gox) - (float)(ipxo - 1) * dnx;
    float drz = (rgz_j - goz) - (float)(ipzo - 1) * dnz;
    float vel = vel_at(veln, nnx, nnz, ipxo, ipzo, drx, drz, dnx, dnz, true);
    drx = (rgx_j - gox) - (float)(ivxo - 1) * dvx;
    drz = (rgz_j - goz) - (float)(ivzo - 1) * dvz;
    float vi[4], wi[4], vio[4], wio[4];
    bsp4(drx / dvx, vi);
    bsp4(drz / dvz, wi);
    int ivxt = ivxo, ivzt = ivzo;
    for (int k = 0; k < nhp; k++) {
      const float velo = vel;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        vio[q] = vi[q];
        wio[q] = wi[q];
      }
      if (k > 0) {
        if (chp[k - 1] == 1)
          ivxt = ivx;
        else if (chp[k - 1] == 2)
          ivzt = ivz;
      }
      const float rigz = rgz_j + vrat[k] * (rgz_n - rgz_j);
      const float rigx = rgx_j + vrat[k] * (rgx_n - rgx_j);
      const int ipxt = (int)((rigx - gox) / dnx) + 1;
      const int ipzt = (int)((rigz - goz) / dnz) + 1;
      drx = (rigx - gox) - (float)(ipxt - 1) * dnx;
      drz = (rigz - goz) - (float)(ipzt - 1) * dnz;
      vel = vel_at(veln, nnx, nnz, ipxt, ipzt, drx, drz, dnx, dnz, false);
      drx = (rigx - gox) - (float)(ivxt - 1) * dvx;
      drz = (rigz - goz) - (float)(ivzt - 1) * dvz;
      bsp4(drx / dvx, vi);
      bsp4(drz / dvz, wi);
      float dinc;
      if (k == 0)
        dinc = vrat[k] * dpl;
      else
        dinc = (vrat[k] - vrat[k - 1]) * dpl;
      if (ivzt != civz || ivxt != civx) {  // switch the register cache to vertex cell (ivzt, ivxt)
        if (civz > -1000) {
#pragma unroll
          for (int l = 0; l < 4; l++)
#pragma unroll
            for (int m = 0; m < 4; m++) fdm[(size_t)(civx - 1 + m) * fld + (civz - 1 + l)] = acc[l * 4 + m];
        }
        civz = ivzt;
        civx = ivxt;
#pragma unroll
        for (int l = 0; l < 4; l++)
#pragma unroll
          for (int m = 0; m < 4; m++) acc[l * 4 + m] = fdm[(size_t)(civx - 1 + m) * fld + (civz - 1 + l)];
        bz0 = min(bz0, civz - 1);
        bz1 = max(bz1, civz + 2);
        bx0 = min(bx0, civx - 1);
        bx1 = max(bx1, civx + 2);
      }
      const float vel2 = vel * vel, velo2 = velo * velo;
#pragma unroll
      for (int l = 0; l < 4; l++)
#pragma unroll
        for (int m = 0; m < 4; m++) {
          float r1 = vi[m] * wi[l] / vel2;
          const float r2 = vio[m] * wio[l] / velo2;
          r1 = -(r1 + r2) * dinc / 2.0f;
          acc[l * 4 + m] = r1 + acc[l * 4 + m];
        }
    }
    rgx_j = rgx_n;
    rgz_j = rgz_n;
  }
  if (civz > -1000) {
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
      for (int m = 0; m < 4; m++) fdm[(size_t)(civx - 1 + m) * fld + (civz - 1 + l)] = acc[l * 4 + m];
  }
  bbox[rid] = make_int4(bz0, bz1, bx0, bx1);  // vertex index ranges (i: z, j: x), empty if bz1 < bz0
  if (rb) atomicExch(rbint_flag, 1);
  if (path_n) path_n[rid] = np;
}

int launch_rays(cudaStream_t st, const Geom &g, const SweepDesc *d_sw, const RayDesc *d_rays, int nrays,
                const float *d_veln_all, BatchView bv, float *d_tt, float *d_fdm, int4 *d_bbox,
                int *d_rbint, int *d_err, float2 *d_path, int *d_path_n, int path_cap) {
  if (nrays <= 0) return DSURF_OK;
  k_rays<<<(nrays + 127) / 128, 128, 0, st>>>(g, d_sw, d_rays, nrays, d_veln_all, bv, d_tt, d_fdm, d_bbox,
                                              d_rbint, d_err, d_path, d_path_n, path_cap);
  DS_CUDA(cudaGetLastError());
  return DSURF_OK;
}

// ---------------------------------------------------------------- K6: row assembly
// S[k][col] = sen_vp*coe_a + sen_rho*coe_rho + sen_vs for one period-type (fp64, source order of
// CalSurfG.f90:1396-1399); coe_a/coe_rho are REAL*4 polynomials of vels (:1388-1395 / :1404-1411).
__device__ __forceinline__ float p3(float x) { return x * (x * x); }
__device__ __forceinline__ float p4(float x) { const float x2 = x * x; return x2 * x2; }

__global__ void k_coef(const float *__restrict__ vels, int nx, int ny, int nz, int brocher,
                       float *__restrict__ coe_a, float *__restrict__ coe_rho) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = nx * ny;
  if (gid >= ncol * (nz - 1)) return;
  const float v = vels[gid];
  float ca, vp;
  if (brocher) {
    ca = (2.0947f - 0.8206f * 2 * v + 0.2683f * 3 * (v * v) - 0.0251f * 4 * p3(v));
    vp = 0.9409f + 2.0947f * v - 0.8206f * (v * v) + 0.2683f * p3(v) - 0.0251f * p4(v);
  } else {
    ca = (2.2110f - 0.8984f * 2 * v + 0.2786f * 3 * (v * v) - 0.02412f * 4 * p3(v));
    vp = 0.9098f + 2.2110f * v - 0.8984f * (v * v) + 0.2786f * p3(v) - 0.02412f * p4(v);
  }
  coe_a[gid] = ca;
  coe_rho[gid] = ca * (1.6612f - 0.4721f * 2 * vp + 0.0671f * 3 * (vp * vp) - 0.0043f * 4 * p3(vp) +
                       0.000106f * 5 * p4(vp));
}

// S for all periods of one type: layout [k (nz-1)][period][col]
__global__ void k_combine(const double *__restrict__ sen_vs, const double *__restrict__ sen_vp,
                          const double *__restrict__ sen_rho, const float *__restrict__ coe_a,
                          const float *__restrict__ coe_rho, int ncol, int kmax_t, int nzm1,
                          double *__restrict__ S) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tot = (long long)ncol * kmax_t * nzm1;
  if (gid >= tot) return;
  const int col = (int)(gid % ncol);
  const int k = (int)(gid / ((long long)ncol * kmax_t));
  // sen layout [i (nz)][period][col]: same linear index for i = k < nz-1
  const double r = sen_vp[gid] * (double)coe_a[(size_t)k * ncol + col] +
                   sen_rho[gid] * (double)coe_rho[(size_t)k * ncol + col] + sen_vs[gid];
  S[gid] = r;
}

constexpr float kFtol = 1e-4f;  // CalSurfG.f90:1029

// pass 1: per ray, count / list the fdm entries with |fdm| >= ftol inside the bbox, in (jj, kk)
// order (jj = z vertex 1..nvz outer, kk = x vertex 1..nvx inner).  One warp per ray.
template <bool FILL>
__global__ void k_list(const float *__restrict__ fdm_all, const int4 *__restrict__ bbox, int nrays, int nvx,
                       int nvz, const long long *__restrict__ off, int *__restrict__ cnt,
                       int *__restrict__ lpos, float *__restrict__ lval) {
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= nrays) return;
  const int4 b = bbox[ray];
  const int j0 = max(b.x, 1), j1 = min(b.y, nvz), k0 = max(b.z, 1), k1 = min(b.w, nvx);
  const int fld = nvz + 2;
  const float *fdm = fdm_all + (size_t)ray * (size_t)(nvz + 2) * (nvx + 2);
  long long base = FILL ? off[ray] : 0;
  int total = 0;
  if (j1 >= j0 && k1 >= k0) {
    const int w = k1 - k0 + 1;
    const int n = (j1 - j0 + 1) * w;
    for (int s = 0; s < n; s += 32) {
      const int t = s + lane;
      bool hit = false;
      float fd = 0.0f;
      int jj = 0, kk = 0;
      if (t < n) {
        jj = j0 + t / w;
        kk = k0 + t % w;
        fd = fdm[(size_t)kk * fld + jj];
        hit = fabsf(fd) >= kFtol;
      }
      const unsigned m = __ballot_sync(kFull, hit);
      if (FILL && hit) {
        const long long p = base + total + __popc(m & ((1u << lane) - 1));
        lpos[p] = (jj - 1) * nvx + kk;  // (jj-1)*nvx + kk, 1-based position inside a depth slab
        lval[p] = fd;
      }
      total += __popc(m);
    }
  }
  if (!FILL && lane == 0) cnt[ray] = total;
}

// pass 2: per ray, walk depth k = 0..nz-2 outer, listed vertices inner; value =
// real(S(k,period,col) * fdm); keep |value| > ftol.  One warp per ray.
template <bool FILL>
__global__ void k_rows(const long long *__restrict__ loff, const int *__restrict__ lcnt,
                       const int *__restrict__ lpos, const float *__restrict__ lval,
                       const int *__restrict__ ray_S, const double *const *__restrict__ S_ptr,
                       const long long *__restrict__ S_stride, int nrays,
                       int nx, int ny, int nz, const long long *__restrict__ roff, int *__restrict__ rcnt,
                       const int *__restrict__ ray_row, float *__restrict__ rw, int *__restrict__ col,
                       int *__restrict__ rowidx, long long nar_base, long long cap, int *__restrict__ err) {
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= nrays) return;
  const int nvx = nx - 2, nvz = ny - 2;
  const long long lb = loff[ray];
  const int n = lcnt[ray];
  // S slab of this ray's period-type: [k][ncol] with stride given by the table entry
  const int sid = ray_S[ray];
  const double *S = S_ptr[sid];
  const long long kstride = S_stride[sid];
  long long base = FILL ? nar_base + roff[ray] : 0;
  const int grow = FILL ? ray_row[ray] : 0;
  int total = 0;
  for (int k = 0; k < nz - 1; k++) {
    for (int s = 0; s < n; s += 32) {
      const int t = s + lane;
      bool hit = false;
      float r = 0.0f;
      int pos = 0;
      if (t < n) {
        pos = lpos[lb + t];
        const float fd = lval[lb + t];
        const int jj = (pos - 1) / nvx + 1, kk = (pos - 1) % nvx + 1;
        const size_t cidx = (size_t)jj * nx + kk;  // model column (kk+1, jj+1), 0-based
        r = (float)(S[(size_t)k * kstride + cidx] * (double)fd);
        hit = fabsf(r) > kFtol;
      }
      const unsigned m = __ballot_sync(kFull, hit);
      if (FILL && hit) {
        const long long p = base + total + __popc(m & ((1u << lane) - 1));
        if (p < cap) {
          rw[p] = r;
          col[p] = k * nvz * nvx + pos;
          rowidx[p] = grow + 1;  // count11, 1-based
        } else {
          atomicExch(err, DSURF_ERR_CAPACITY);
        }
      }
      total += __popc(m);
    }
  }
  if (!FILL && lane == 0) rcnt[ray] = total;
}

__global__ void k_clear_fdm(float *__restrict__ fdm_all, const int4 *__restrict__ bbox, int nrays, int nvx,
                            int nvz) {
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= nrays) return;
  const int4 b = bbox[ray];
  if (b.y < b.x || b.w < b.z) return;
  const int fld = nvz + 2;
  float *fdm = fdm_all + (size_t)ray * (size_t)(nvz + 2) * (nvx + 2);
  const int h = b.y - b.x + 1, w = b.w - b.z + 1;
  for (int t = lane; t < h * w; t += 32) fdm[(size_t)(b.z + t / h) * fld + (b.x + t % h)] = 0.0f;
}

__global__ void k_widen_i(const int *in, long long *out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

int exclusive_scan_ll(cudaStream_t st, const int *d_cnt, long long *d_wide, long long *d_off, int n,
                      DevBuf<char> &tmp, long long *h_total) {
  k_widen_i<<<(n + 255) / 256, 256, 0, st>>>(d_cnt, d_wide, n);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_wide, d_off, n, st);
  if (tmp.reserve(tb + 16)) return DSURF_ERR_CUDA;
  cub::DeviceScan::ExclusiveSum(tmp.p, tb, d_wide, d_off, n, st);
  long long last_off = 0, last_cnt = 0;
  DS_CUDA(cudaMemcpyAsync(&last_off, d_off + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaMemcpyAsync(&last_cnt, d_wide + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  *h_total = last_off + last_cnt;
  return DSURF_OK;
}

// host drivers used by plan.cu
int launch_coef(cudaStream_t st, const float *d_vels, int nx, int ny, int nz, int brocher, float *coe_a,
                float *coe_rho) {
  const int n = nx * ny * (nz - 1);
  k_coef<<<(n + 255) / 256, 256, 0, st>>>(d_vels, nx, ny, nz, brocher, coe_a, coe_rho);
  return DSURF_OK;
}
int launch_combine(cudaStream_t st, const double *sen_vs, const double *sen_vp, const double *sen_rho,
                   const float *coe_a, const float *coe_rho, int ncol, int kmax_t, int nzm1, double *S) {
  const long long tot = (long long)ncol * kmax_t * nzm1;
  if (tot <= 0) return DSURF_OK;
  k_combine<<<(int)((tot + 255) / 256), 256, 0, st>>>(sen_vs, sen_vp, sen_rho, coe_a, coe_rho, ncol, kmax_t,
                                                      nzm1, S);
  return DSURF_OK;
}

int launch_assembly(cudaStream_t st, const Geom &g, int nz, const float *d_fdm, const int4 *d_bbox, int nrays,
                    const int *d_ray_S, const double *const *d_S_ptr, const long long *d_S_stride,
                    const int *d_ray_row,
                    DevBuf<int> &cnt, DevBuf<long long> &wide, DevBuf<long long> &loff,
                    DevBuf<long long> &roff, DevBuf<int> &lpos, DevBuf<float> &lval, DevBuf<int> &lcnt,
                    DevBuf<char> &tmp, DevBuf<float> &rw, DevBuf<int> &col, DevBuf<int> &rowidx,
                    long long &nar, int *d_err, int *launches) {
  if (nrays <= 0) return DSURF_OK;
  const int nvx = g.nvx, nvz = g.nvz;
  if (cnt.reserve(nrays) || wide.reserve(nrays) || loff.reserve(nrays) || roff.reserve(nrays) ||
      lcnt.reserve(nrays))
    return DSURF_ERR_CUDA;
  const int grid = (nrays * 32 + 127) / 128;
  k_list<false><<<grid, 128, 0, st>>>(d_fdm, d_bbox, nrays, nvx, nvz, nullptr, lcnt.p, nullptr, nullptr);
  long long ltot = 0;
  DS_CHECK(exclusive_scan_ll(st, lcnt.p, wide.p, loff.p, nrays, tmp, &ltot));
  if (lpos.reserve(ltot + 1) || lval.reserve(ltot + 1)) return DSURF_ERR_CUDA;
  k_list<true><<<grid, 128, 0, st>>>(d_fdm, d_bbox, nrays, nvx, nvz, loff.p, nullptr, lpos.p, lval.p);
  k_rows<false><<<grid, 128, 0, st>>>(loff.p, lcnt.p, lpos.p, lval.p, d_ray_S, d_S_ptr, d_S_stride, nrays, g.nx, g.ny, nz,
                                      nullptr, cnt.p, nullptr, nullptr, nullptr, nullptr, 0, 0, d_err);
  long long rtot = 0;
  DS_CHECK(exclusive_scan_ll(st, cnt.p, wide.p, roff.p, nrays, tmp, &rtot));
  const size_t need = (size_t)(nar + rtot);
  if (need > rw.cap) {
    const size_t ncap = std::max(need, rw.cap * 2);
    if (rw.reserve(ncap, true, st) || col.reserve(ncap, true, st) || rowidx.reserve(ncap, true, st)) {
      set_error(__FILE__, __LINE__, "cudaMalloc failed (COO growth)");
      return DSURF_ERR_CUDA;
    }
  }
  k_rows<true><<<grid, 128, 0, st>>>(loff.p, lcnt.p, lpos.p, lval.p, d_ray_S, d_S_ptr, d_S_stride, nrays, g.nx, g.ny, nz,
                                     roff.p, nullptr, d_ray_row, rw.p, col.p, rowidx.p, nar, (long long)rw.cap,
                                     d_err);
  k_clear_fdm<<<grid, 128, 0, st>>>(const_cast<float *>(d_fdm), d_bbox, nrays, nvx, nvz);
  nar += rtot;
  if (launches) *launches += 9;
  DS_CUDA(cudaGetLastError());
  return DSURF_OK;
}

}  // namespace dsurf
