// Host orchestration of the forward/sensitivity path, replacing subroutine CalSurfG
// (src/CalSurfG.f90:939-1459): K1 dispersion + depth kernels per data type, then for batches of
// (gather, ig) sweeps K2 dicing, K3 eikonal, K4/K5 receiver times + rays, K6 row assembly into a
// device-resident COO in the reference's row/column order.
//
// What the reference redoes per source and this plan hoists (identical values):
//   * gridder per source (:1186)            -> each velocity map is diced once per model;
//   * slab copies of sen_* per source (:1146-1168) -> per-type S = sen_vp*coe_a+sen_rho*coe_rho
//     +sen_vs combined once (same fp64 operation order as :1396-1399);
//   * dense row(nparpi) zero + scan per ray (:1383,1425) -> ordered compaction of the touched
//     vertices.
// Geometry scalars and sin(colatitude) tables are computed on the host in REAL*4 with the C
// library exactly as the reference computes them, so the device never evaluates a sine on the
// eikonal path.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/dsurftomo_b200.h"
#include "common.cuh"
#include "plan.cuh"
#include "glue.cuh"
#include "eik_fim.cuh"

namespace dsurf {
int launch_coef(cudaStream_t st, const float *d_vels, int nx, int ny, int nz, int brocher, float *coe_a,
                float *coe_rho);
int launch_combine(cudaStream_t st, const double *sen_vs, const double *sen_vp, const double *sen_rho,
                   const float *coe_a, const float *coe_rho, int ncol, int kmax_t, int nzm1, double *S);
int launch_assembly(cudaStream_t st, const Geom &g, int nz, const float *d_fdm, const int4 *d_bbox, int nrays,
                    const int *d_ray_S, const double *const *d_S_ptr, const long long *d_S_stride,
                    const int *d_ray_row, DevBuf<int> &cnt, DevBuf<long long> &wide, DevBuf<long long> &loff,
                    DevBuf<long long> &roff, DevBuf<int> &lpos, DevBuf<float> &lval, DevBuf<int> &lcnt,
                    DevBuf<char> &tmp, DevBuf<float> &rw, DevBuf<int> &col, DevBuf<int> &rowidx,
                    long long &nar, int *d_err, int *launches);
}  // namespace dsurf

using namespace dsurf;

namespace {
struct GatherInfo {
  int knumi, srcnum;  // 1-based loop indices of CalSurfG.f90:1144-1145
  int type;           // 0 Rc, 1 Rg, 2 Lc, 3 Lg
  int per;            // periods(srcnum,knumi), 1-based inside the type
  int igr;            // 0 phase, 1 group
  int nrc;
  int first_row;      // 0-based global row of the first receiver
  float scx, scz;
};
struct SweepRef {
  int gather, ig;
};
}  // namespace

struct dsurf_plan {
  Geom g{};
  int nz = 0, kmaxT[4] = {0, 0, 0, 0}, kmax = 0, nsrc = 0, nrcf = 0;
  bool forward_only = false;  // subroutine synthetic (CalSurfG.f90:2412-2865): times only, gdx = gdz = 5
  float minthk = 0;
  std::vector<float> depz;
  std::vector<double> tper[4];
  std::vector<GatherInfo> gathers;
  std::vector<float> rcx, rcz;  // per gather receivers, flattened by first_row
  int dall = 0;
  cudaStream_t st = nullptr;
  // model + dispersion
  DevBuf<float> vels, coe_a, coe_rho;
  LayerTablesDev tables;
  DevBuf<double> dt[4], pv[4], sen[4][3], S[4], cgbuf;
  int pvcols[4] = {0, 0, 0, 0};
  bool disp_done = false;
  // velocity maps: slot = mapslot[array][col]
  std::vector<int> mapslot[4];
  int nmaps = 0;
  DevBuf<float> velv_all, veln_all, risti_c;
  bool maps_diced = false;
  // S table (one entry per (type, period))
  DevBuf<const double *> S_ptr;
  DevBuf<long long> S_stride;
  std::vector<int> S_id_base = std::vector<int>(4, 0);
  // batch workspace
  int maxslots = 0, maxrays = 0, hcap = 0;
  bool words = true;          // node state: one word per node (eik_lps.cuh / eik_fim.cuh) or packed (time, status) records
  int mode = 0;               // eikonal pipeline this plan was sized for (kEikExact16 / kEikLps / kEikFim)
  int wld = 0, wpx = 0, wpz = 0;
  size_t wslot = 0;           // word layout (BatchView)
  DevBuf<unsigned> fim_bitmap, fim_rw;
  DevBuf<int2> fim_reg;
  DevBuf<int4> fim_rect;
  DevBuf<unsigned char> fim_flag;
  DevBuf<int2> node, noder, box, seed;
  DevBuf<unsigned> word;
  DevBuf<int> nseed;
  DevBuf<float> velr, ristr;
  DevBuf<int2> hent;
  DevBuf<SweepDesc> d_sw;
  DevBuf<RayDesc> d_rays;
  DevBuf<float> fdm;
  DevBuf<int4> bbox;
  DevBuf<int> ray_S, ray_row;
  DevBuf<int> cnt, lcnt, lpos;
  DevBuf<long long> wide, loff, roff;
  DevBuf<float> lval;
  DevBuf<char> tmp;
  // outputs
  DevBuf<float> dsurf, rw;
  DevBuf<int> col, rowidx, flags;  // flags[0] err, flags[1] rbint
  long long nar = 0;
  int last_g0 = 0, last_g1 = 0;  // gather range of the rows currently in the COO (dsurf_plan_sweeps)
  // multi-GPU: COO of every rank in rank order (dsurf_plan_allgather)
  DevBuf<float> G_rw;
  DevBuf<int> G_col, G_row;
  DevBuf<long long> G_cnt;
  long long G_nar = -1;
  double ms_gather = 0;
  double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evt0 = nullptr, evt1 = nullptr;
  double ms_sweeps_total = 0;
  // optional ray-path export (raypath.out, CalSurfG.f90:2276-2283)
  FILE *path_fh = nullptr;
  int path_cap = 0;
  long long path_rays = 0;
  DevBuf<float2> path;
  DevBuf<int> path_n;
  // device-resident host glue (glue.cu)
  DevBuf<float> g_obst, g_cbst, g_datw, g_sorted, g_sval;
  DevBuf<double> g_norm;
  DevBuf<int> g_srow, g_scol;
  DevBuf<char> g_tmp;
  long long g_nsm = 0;
  int g_count3 = 0;
  float g_weight = -1.0f;
  bool g_valid = false;
  GlueStats g_stats{};
};

static const float kPi = 3.1415926535898f;  // CalSurfG.f90:196

static void make_geom(Geom &g, int nx, int ny, float goxd, float gozd, float dvxd, float dvzd, int gd) {
  g.gd = gd;
  g.nx = nx;
  g.ny = ny;
  g.nvx = nx - 2;
  g.nvz = ny - 2;
  g.earth = 6371.0f;
  g.dvx = dvxd * kPi / 180.0f;
  g.dvz = dvzd * kPi / 180.0f;
  g.gox = (90.0f - goxd) * kPi / 180.0f;
  g.goz = gozd * kPi / 180.0f;
  g.nnx = (g.nvx - 1) * gd + 1;
  g.nnz = (g.nvz - 1) * gd + 1;
  g.dnx = g.dvx / (float)gd;
  g.dnz = g.dvz / (float)gd;
  g.drnx = g.dvx / (float)(gd * kSgdl);
  g.drnz = g.dvz / (float)(gd * kSgdl);
  float dpl = g.dnx * g.earth;                        // :1705-1709 / :1864-1869
  float rd1 = g.dnz * g.earth * sinf(g.gox);
  if (rd1 < dpl) dpl = rd1;
  rd1 = g.dnz * g.earth * sinf(g.gox + (float)(g.nnx - 1) * g.dnx);
  if (rd1 < dpl) dpl = rd1;
  g.dpl_sr = dpl;
  g.dpl_ray = 0.5f * dpl;
  g.x_last = g.gox + (float)(g.nnx - 1) * g.dnx;
  g.z_last = g.goz + (float)(g.nnz - 1) * g.dnz;
}

// host part of CalSurfG.f90:1209-1240 + travel's :312-325 + rpaths' :1853-1854 for one source
static int make_sweep(const Geom &g, float x, float z, SweepDesc &d, float *ristr /*129*/) {
  int isx = (int)((x - g.gox) / g.dnx) + 1;
  int isz = (int)((z - g.goz) / g.dnz) + 1;
  if (isx < 1 || isx > g.nnx || isz < 1 || isz > g.nnz) return DSURF_ERR_SOURCE_OUTSIDE;
  if (isx == g.nnx) isx = isx - 1;
  if (isz == g.nnz) isz = isz - 1;
  d.isx = isx;
  d.isz = isz;
  d.scx = x;
  d.scz = z;
  d.vnl = std::max(isx - kSgs, 1);
  d.vnr = std::min(isx + kSgs, g.nnx);
  d.vnt = std::max(isz - kSgs, 1);
  d.vnb = std::min(isz + kSgs, g.nnz);
  d.nrnx = (d.vnr - d.vnl) * kSgdl + 1;
  d.nrnz = (d.vnb - d.vnt) * kSgdl + 1;
  d.gorx = g.gox + g.dnx * (float)(d.vnl - 1);
  d.gorz = g.goz + g.dnz * (float)(d.vnt - 1);
  int tsx = (int)((x - d.gorx) / g.drnx) + 1;
  int tsz = (int)((z - d.gorz) / g.drnz) + 1;
  d.rsx = tsx;
  d.rsz = tsz;
  if (tsx < 1 || tsx > d.nrnx || tsz < 1 || tsz > d.nrnz) return DSURF_ERR_SOURCE_OUTSIDE;
  if (tsx == d.nrnx) tsx = tsx - 1;
  if (tsz == d.nrnz) tsz = tsz - 1;
  d.tsx = tsx;
  d.tsz = tsz;
  for (int ix = 1; ix <= d.nrnx; ix++) ristr[ix - 1] = g.earth * sinf(d.gorx + (float)(ix - 1) * g.drnx);
  d.status = 0;
  return DSURF_OK;
}

static int plan_create_impl(dsurf_plan **out, int nx, int ny, int nz, const float *vels, float goxdf,
                            float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                            int kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                            const double *tLg, const int *wavetype, const int *igrt, const int *periods,
                            const float *depz, float minthk, const float *scxf, const float *sczf,
                            const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                            int kmax, int nsrcsurf, int nrcf, int gd, bool forward_only) {
  DS_CHECK(ensure_device());
  if (!out || nx < 4 || ny < 4 || nz < 2 || kmax != kmaxRc + kmaxRg + kmaxLc + kmaxLg) return DSURF_ERR_BAD_ARG;
  if (!forward_only) dsurf_lsmr_hint_geometry(nx, ny, nz);
  auto *p = new dsurf_plan();
  p->forward_only = forward_only;
  make_geom(p->g, nx, ny, goxdf, gozdf, dvxdf, dvzdf, gd);
  p->nz = nz;
  p->kmaxT[0] = kmaxRc;
  p->kmaxT[1] = kmaxRg;
  p->kmaxT[2] = kmaxLc;
  p->kmaxT[3] = kmaxLg;
  p->kmax = kmax;
  p->nsrc = nsrcsurf;
  p->nrcf = nrcf;
  p->minthk = minthk;
  p->depz.assign(depz, depz + nz);
  const double *tp[4] = {tRc, tRg, tLc, tLg};
  for (int t = 0; t < 4; t++)
    if (p->kmaxT[t] > 0) p->tper[t].assign(tp[t], tp[t] + p->kmaxT[t]);  // never touch t when kmaxX == 0
  // pv arrays: Rc and Lc are dimensioned with kmax columns in the reference (:1003-1004)
  // (subroutine synthetic dimensions every pv array by its own kmaxX, :2481)
  p->pvcols[0] = forward_only ? std::max(kmaxRc, 1) : kmax;
  p->pvcols[1] = std::max(kmaxRg, 1);
  p->pvcols[2] = forward_only ? std::max(kmaxLc, 1) : kmax;
  p->pvcols[3] = std::max(kmaxLg, 1);
  for (int t = 0; t < 4; t++) p->mapslot[t].assign(p->pvcols[t], -1);
  // ---- flatten the gather loop nest and assign velocity-map slots
  const int koff[4] = {0, kmaxRc, kmaxRc + kmaxRg, kmaxRc + kmaxRg + kmaxLc};
  int row = 0;
  auto slot_of = [&](int arr, int col0) {
    if (col0 < 0 || col0 >= p->pvcols[arr]) return -1;
    if (p->mapslot[arr][col0] < 0) p->mapslot[arr][col0] = p->nmaps++;
    return p->mapslot[arr][col0];
  };
  for (int knumi = 1; knumi <= kmax; knumi++) {
    for (int srcnum = 1; srcnum <= nsrcsurf1[knumi - 1]; srcnum++) {
      const size_t i2 = (size_t)(knumi - 1) * nsrcsurf + (srcnum - 1);
      const int wt = wavetype[i2], gr = igrt[i2];
      int type = -1;
      if (wt == 2 && gr == 0) type = 0;
      if (wt == 2 && gr == 1) type = 1;
      if (wt == 1 && gr == 0) type = 2;
      if (wt == 1 && gr == 1) type = 3;
      GatherInfo gi;
      gi.knumi = knumi;
      gi.srcnum = srcnum;
      gi.type = type;
      gi.per = periods[i2];
      gi.igr = gr;
      gi.nrc = nrc1[i2];
      gi.first_row = row;
      gi.scx = scxf[i2];
      gi.scz = sczf[i2];
      if (type < 0 || gi.per < 1 || gi.per > p->kmaxT[type] || knumi - koff[type] < 1 ||
          knumi - koff[type] > p->kmaxT[type]) {
        set_error(__FILE__, __LINE__, "gather wavetype/igrt/periods inconsistent with its knumi block");
        delete p;
        return DSURF_ERR_BAD_ARG;
      }
      slot_of(type, gi.per - 1);
      if (gr == 1 && !forward_only) slot_of(type == 1 ? 0 : 2, gi.per - 1);  // ig = 2: phase map (:1177-1185)
      for (int r = 0; r < gi.nrc; r++) {
        const size_t i3 = i2 * nrcf + r;
        p->rcx.push_back(rcxf[i3]);
        p->rcz.push_back(rczf[i3]);
      }
      row += gi.nrc;
      p->gathers.push_back(gi);
    }
  }
  p->dall = row;
  const size_t ncol = (size_t)nx * ny;
  const size_t Nc = (size_t)p->g.nnx * p->g.nnz;
  bool bad = false;
  bad |= p->vels.reserve(ncol * nz) != cudaSuccess;
  bad |= p->coe_a.reserve(ncol * (nz - 1)) != cudaSuccess;
  bad |= p->coe_rho.reserve(ncol * (nz - 1)) != cudaSuccess;
  for (int t = 0; t < 4; t++) {
    const int kt = p->kmaxT[t];
    bad |= p->pv[t].reserve(ncol * p->pvcols[t]) != cudaSuccess;
    if (kt > 0) {
      bad |= p->dt[t].reserve(kt) != cudaSuccess;
      if (!forward_only) {
        for (int q = 0; q < 3; q++) bad |= p->sen[t][q].reserve(ncol * kt * nz) != cudaSuccess;
        bad |= p->S[t].reserve(ncol * kt * (nz - 1)) != cudaSuccess;
      }
    }
  }
  bad |= p->velv_all.reserve(std::max<size_t>(1, p->nmaps) * ncol) != cudaSuccess;
  bad |= p->veln_all.reserve(std::max<size_t>(1, p->nmaps) * Nc) != cudaSuccess;
  bad |= p->risti_c.reserve(p->g.nnx) != cudaSuccess;
  bad |= p->dsurf.reserve(std::max(1, p->dall)) != cudaSuccess;
  bad |= p->flags.reserve(4) != cudaSuccess;
  bad |= p->S_ptr.reserve(std::max(1, kmax)) != cudaSuccess;
  bad |= p->S_stride.reserve(std::max(1, kmax)) != cudaSuccess;
  if (bad) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (plan)");
    delete p;
    return DSURF_ERR_CUDA;
  }
  cudaMemcpy(p->vels.p, vels, ncol * nz * sizeof(float), cudaMemcpyHostToDevice);
  for (int t = 0; t < 4; t++) {
    if (p->kmaxT[t] > 0) cudaMemcpy(p->dt[t].p, p->tper[t].data(), p->kmaxT[t] * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(p->pv[t].p, 0, ncol * p->pvcols[t] * sizeof(double));
  }
  {
    std::vector<float> r(p->g.nnx);
    for (int ix = 1; ix <= p->g.nnx; ix++) r[ix - 1] = p->g.earth * sinf(p->g.gox + (float)(ix - 1) * p->g.dnx);
    cudaMemcpy(p->risti_c.p, r.data(), r.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  cudaMemset(p->flags.p, 0, 4 * sizeof(int));
  cudaMemset(p->dsurf.p, 0, std::max(1, p->dall) * sizeof(float));
  {  // S table: id = koff[type] + (period-1)
    std::vector<const double *> sp(std::max(1, kmax), nullptr);
    std::vector<long long> ss(std::max(1, kmax), 0);
    for (int t = 0; t < 4; t++)
      for (int c = 0; c < p->kmaxT[t]; c++) {
        sp[koff[t] + c] = p->S[t].p + (size_t)c * ncol;
        ss[koff[t] + c] = (long long)p->kmaxT[t] * (long long)ncol;
      }
    for (int t = 0; t < 4; t++) p->S_id_base[t] = koff[t];
    cudaMemcpy(p->S_ptr.p, sp.data(), sp.size() * sizeof(const double *), cudaMemcpyHostToDevice);
    cudaMemcpy(p->S_stride.p, ss.data(), ss.size() * sizeof(long long), cudaMemcpyHostToDevice);
  }
  LayerTables T;
  make_layer_tables(depz, nz, minthk, T);
  if (upload_tables(T, p->tables) != DSURF_OK) {
    delete p;
    return DSURF_ERR_CUDA;
  }
  for (auto &e : p->ev) cudaEventCreate(&e);
  cudaEventCreate(&p->evt0);
  cudaEventCreate(&p->evt1);
  // ---- batch workspace sizing
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  p->hcap = 8 * (p->g.nnx + p->g.nnz) + 1024;
  p->mode = eikonal_mode();
  if (p->mode == kEikLps) {
    // the blocked heap slab (eikonal.cu: lps_gaddr) is allocated in whole level groups: capacities 2^(10+3k) - 1.
    // Measured narrow bands at 1025^2: mean 1.8-2.8 k, max 4.1 k entries (tests/host/lps_host_check.cpp).
    int cap = 1023;
    while (cap < 2 * (p->g.nnx + p->g.nnz)) cap = cap * 8 + 7;
    p->hcap = cap;
  }
  {
    const long long maxbt = (long long)std::lround(0.5 * (double)p->g.nnx * p->g.nnz);  // the reference's limit (:1093)
    if (p->hcap > maxbt) p->hcap = (int)std::max<long long>(maxbt, 16);
  }
  if (const char *hc = getenv("DSURF_HCAP")) p->hcap = std::max(16, atoi(hc));  // test hook: force heap-slab growth
  const size_t fdm_per_ray = forward_only ? sizeof(float) : (size_t)(p->g.nvz + 2) * (p->g.nvx + 2) * sizeof(float);
  int maxnrc = 1;
  for (auto &gi : p->gathers) maxnrc = std::max(maxnrc, gi.nrc);
  p->words = eikonal_uses_words(p->mode);
  const size_t slab = (size_t)eikonal_slab_entries(p->hcap, p->mode);
  const size_t kBox = (size_t)(2 * kSgs + 1) * (2 * kSgs + 1);
  eikonal_word_layout(p->g, p->mode, &p->wld, &p->wpx, &p->wpz, &p->wslot);
  const size_t Nw = p->wslot;  // padded word array of the word pipelines
  const size_t fim_bm = p->mode == kEikFim ? eikonal_fim_bitmap_words(p->g) : 0;
  const size_t fim_slot = p->mode == kEikFim ? fim_bm * sizeof(unsigned) + (size_t)fim::kRegNodes * (sizeof(int2) + sizeof(unsigned) + 1) + sizeof(int4) : 0;
  const size_t per_slot = fim_slot + (p->words ? Nw * sizeof(unsigned) : Nc * sizeof(int2)) +
                          (size_t)kRefMax * kRefMax * (sizeof(int2) + sizeof(float)) + slab * sizeof(int2) +
                          kRefMax * sizeof(float) + sizeof(SweepDesc) + kBox * 2 * sizeof(int2) + sizeof(int);
  // rays are traced and assembled in chunks of at most maxrays (their dense fdm slabs dominate otherwise)
  const size_t per_ray = fdm_per_ray + sizeof(RayDesc) + 64;
  long long nsw_total = 0, nray_total = 0;
  for (auto &gi : p->gathers) {
    const int ns = (gi.igr == 1 && !forward_only) ? 2 : 1;
    nsw_total += ns;
    nray_total += (long long)ns * gi.nrc;
  }
  long long mr = std::max<long long>(4096, (long long)(((size_t)4 << 30) / per_ray));
  mr = std::max<long long>(std::min<long long>(mr, nray_total), maxnrc);
  if (const char *e = getenv("DSURF_MAXRAYS")) mr = std::max<long long>(maxnrc, atoll(e));  // test hook: force ray chunking
  p->maxrays = (int)mr;
  // the lane-per-sweep march wants every sweep resident at once: take what the device has, leaving room for the
  // COO output and the LSMR system that follow
  const size_t reserve_b = std::min<size_t>((size_t)40 << 30, freeb / 4);
  size_t budget = freeb > reserve_b + mr * per_ray ? freeb - reserve_b - (size_t)mr * per_ray : freeb / 2;
  if (!p->words) budget = std::min<size_t>((size_t)(freeb * 0.6), (size_t)100 << 30);
  long long ms = (long long)(budget / per_slot);
  ms = std::min<long long>(ms, std::max<long long>(nsw_total, 1));
  {  // legacy pipeline: batches of at most one resident wave (larger launches would run as equal-length waves)
    const long long res = eikonal_resident_sweeps(p->mode);
    if (res > 0) ms = std::min<long long>(ms, res);
  }
  if (const char *e = getenv("DSURF_MAXSLOTS")) ms = std::min<long long>(ms, std::max(1, atoi(e)));  // test hook: force batching
  ms = std::max<long long>(ms, 1);
  p->maxslots = (int)ms;
  bad = false;
  if (p->words) {
    bad |= p->word.reserve((size_t)p->maxslots * Nw) != cudaSuccess;
    bad |= p->box.reserve((size_t)p->maxslots * kBox) != cudaSuccess;
    bad |= p->seed.reserve((size_t)p->maxslots * kBox) != cudaSuccess;
    bad |= p->nseed.reserve((size_t)p->maxslots) != cudaSuccess;
    if (p->mode == kEikFim) {
      bad |= p->fim_bitmap.reserve((size_t)p->maxslots * fim_bm) != cudaSuccess;
      bad |= p->fim_rw.reserve((size_t)p->maxslots * fim::kRegNodes) != cudaSuccess;
      bad |= p->fim_reg.reserve((size_t)p->maxslots * fim::kRegNodes) != cudaSuccess;
      bad |= p->fim_flag.reserve((size_t)p->maxslots * fim::kRegNodes) != cudaSuccess;
      bad |= p->fim_rect.reserve((size_t)p->maxslots) != cudaSuccess;
    }
  } else {
    bad |= p->node.reserve((size_t)p->maxslots * Nc) != cudaSuccess;
  }
  bad |= p->noder.reserve((size_t)p->maxslots * kRefMax * kRefMax) != cudaSuccess;
  bad |= p->velr.reserve((size_t)p->maxslots * kRefMax * kRefMax) != cudaSuccess;
  bad |= p->hent.reserve((size_t)p->maxslots * slab) != cudaSuccess;
  bad |= p->ristr.reserve((size_t)p->maxslots * kRefMax) != cudaSuccess;
  bad |= p->d_sw.reserve(p->maxslots) != cudaSuccess;
  bad |= p->d_rays.reserve(p->maxrays) != cudaSuccess;
  bad |= p->fdm.reserve((size_t)p->maxrays * (fdm_per_ray / sizeof(float))) != cudaSuccess;
  bad |= p->bbox.reserve(p->maxrays) != cudaSuccess;
  bad |= p->ray_S.reserve(p->maxrays) != cudaSuccess;
  bad |= p->ray_row.reserve(p->maxrays) != cudaSuccess;
  if (bad) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (batch workspace)");
    delete p;
    return DSURF_ERR_CUDA;
  }
  cudaMemset(p->fdm.p, 0, (size_t)p->maxrays * fdm_per_ray);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
    delete p;
    return DSURF_ERR_CUDA;
  }
  *out = p;
  return DSURF_OK;
}

extern "C" int dsurf_plan_create(dsurf_plan **out, int nx, int ny, int nz, const float *vels, float goxdf,
                                 float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                                 int kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                                 const double *tLg, const int *wavetype, const int *igrt, const int *periods,
                                 const float *depz, float minthk, const float *scxf, const float *sczf,
                                 const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                                 int kmax, int nsrcsurf, int nrcf) {
  return plan_create_impl(out, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, kmaxRg, kmaxLc, kmaxLg, tRc, tRg,
                          tLc, tLg, wavetype, igrt, periods, depz, minthk, scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1,
                          kmax, nsrcsurf, nrcf, kGd, false);
}

// forward-only plan of subroutine synthetic (CalSurfG.f90:2412-2865): gdx = gdz = 5, dispersion
// maps by caldespersion (group maps for the group types), one sweep per gather, times only
extern "C" int dsurf_plan_create_forward(dsurf_plan **out, int nx, int ny, int nz, const float *vels, float goxdf,
                                         float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                                         int kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                                         const double *tLg, const int *wavetype, const int *igrt,
                                         const int *periods, const float *depz, float minthk, const float *scxf,
                                         const float *sczf, const float *rcxf, const float *rczf, const int *nrc1,
                                         const int *nsrcsurf1, int kmax, int nsrcsurf, int nrcf) {
  return plan_create_impl(out, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, kmaxRg, kmaxLc, kmaxLg, tRc, tRg,
                          tLc, tLg, wavetype, igrt, periods, depz, minthk, scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1,
                          kmax, nsrcsurf, nrcf, 5, true);
}

extern "C" int dsurf_plan_destroy(dsurf_plan *p) {
  if (!p) return DSURF_OK;
  for (auto &e : p->ev)
    if (e) cudaEventDestroy(e);
  if (p->path_fh) fclose(p->path_fh);
  delete p;
  return DSURF_OK;
}

extern "C" int dsurf_plan_set_model(dsurf_plan *p, const float *vels) {
  if (!p || !vels) return DSURF_ERR_BAD_ARG;
  DS_CUDA(cudaMemcpy(p->vels.p, vels, (size_t)p->g.nx * p->g.ny * p->nz * sizeof(float), cudaMemcpyHostToDevice));
  p->disp_done = false;
  p->maps_diced = false;
  return DSURF_OK;
}

static int dice_maps(dsurf_plan *p) {
  const size_t ncol = (size_t)p->g.nx * p->g.ny, Nc = (size_t)p->g.nnx * p->g.nnz;
  for (int arr = 0; arr < 4; arr++)
    for (int c = 0; c < p->pvcols[arr]; c++) {
      const int s = p->mapslot[arr][c];
      if (s < 0) continue;
      DS_CHECK(launch_dice(p->st, p->g, p->pv[arr].p + (size_t)c * ncol, p->velv_all.p + (size_t)s * ncol,
                           p->veln_all.p + (size_t)s * Nc));
    }
  DS_CUDA(cudaGetLastError());
  p->maps_diced = true;
  return DSURF_OK;
}

// CalSurfG.f90:1098-1133 in the reference's call order (pvRc / pvLc overwrite quirk included)
extern "C" int dsurf_plan_dispersion(dsurf_plan *p) {
  if (!p) return DSURF_ERR_BAD_ARG;
  DS_CHECK(ensure_device());
  const Geom &g = p->g;
  cudaEventRecord(p->ev[0], p->st);
  const int iw[4] = {2, 2, 1, 1}, ig[4] = {0, 1, 0, 1};
  if (p->forward_only) {  // subroutine synthetic: caldespersion per data type (:2552-2613)
    for (int t = 0; t < 4; t++)
      if (p->kmaxT[t] > 0)
        DS_CHECK(run_dispersion(p->st, p->vels.p, g.nx, g.ny, p->nz, p->tables, iw[t], ig[t], p->kmaxT[t], p->dt[t].p,
                                false, p->pv[t].p, nullptr, nullptr, nullptr, p->cgbuf));
    cudaEventRecord(p->ev[1], p->st);
    DS_CHECK(dice_maps(p));
    cudaEventRecord(p->ev[2], p->st);
    DS_CUDA(cudaStreamSynchronize(p->st));
    DS_CUDA(cudaGetLastError());
    float msf = 0;
    cudaEventElapsedTime(&msf, p->ev[0], p->ev[1]);
    p->ms[0] = msf;
    cudaEventElapsedTime(&msf, p->ev[1], p->ev[2]);
    p->ms[1] = msf;
    p->disp_done = true;
    return DSURF_OK;
  }
  for (int t = 0; t < 4; t++) {
    const int kt = p->kmaxT[t];
    if (kt <= 0) continue;
    if (ig[t] == 1) {  // caldespersion: phase maps at the group periods into pvRc / pvLc
      const int ph = (t == 1) ? 0 : 2;
      DS_CHECK(run_dispersion(p->st, p->vels.p, g.nx, g.ny, p->nz, p->tables, iw[t], 0, kt, p->dt[t].p, false,
                              p->pv[ph].p, nullptr, nullptr, nullptr, p->cgbuf));
    }
    DS_CHECK(run_dispersion(p->st, p->vels.p, g.nx, g.ny, p->nz, p->tables, iw[t], ig[t], kt, p->dt[t].p, true,
                            p->pv[t].p, p->sen[t][0].p, p->sen[t][1].p, p->sen[t][2].p, p->cgbuf));
  }
  // coe_a / coe_rho and the combined kernels S (row assembly operands)
  const int brocher = p->depz[p->nz - 2] < 35.0f ? 1 : 0;  // :1385
  DS_CHECK(launch_coef(p->st, p->vels.p, g.nx, g.ny, p->nz, brocher, p->coe_a.p, p->coe_rho.p));
  for (int t = 0; t < 4; t++)
    if (p->kmaxT[t] > 0)
      DS_CHECK(launch_combine(p->st, p->sen[t][0].p, p->sen[t][1].p, p->sen[t][2].p, p->coe_a.p, p->coe_rho.p,
                              g.nx * g.ny, p->kmaxT[t], p->nz - 1, p->S[t].p));
  cudaEventRecord(p->ev[1], p->st);
  DS_CHECK(dice_maps(p));
  cudaEventRecord(p->ev[2], p->st);
  DS_CUDA(cudaStreamSynchronize(p->st));
  DS_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]);
  p->ms[0] = ms;
  cudaEventElapsedTime(&ms, p->ev[1], p->ev[2]);
  p->ms[1] = ms;
  p->disp_done = true;
  return DSURF_OK;
}

// Caller-provided dispersion results for one data type (same layouts as dsurf_depthkernel):
// pv[pvcols][ncol], sen_*[nz][kmax_t][ncol].  Followed by dsurf_plan_finalize_dispersion().
extern "C" int dsurf_plan_set_dispersion(dsurf_plan *p, int type, const double *pv, const double *sen_vs,
                                         const double *sen_vp, const double *sen_rho) {
  if (!p || type < 0 || type > 3) return DSURF_ERR_BAD_ARG;
  const size_t ncol = (size_t)p->g.nx * p->g.ny;
  if (pv) DS_CUDA(cudaMemcpy(p->pv[type].p, pv, ncol * p->pvcols[type] * sizeof(double), cudaMemcpyHostToDevice));
  const size_t ns = ncol * p->kmaxT[type] * p->nz;
  const double *src[3] = {sen_vs, sen_vp, sen_rho};
  for (int q = 0; q < 3; q++)
    if (src[q] && ns > 0) {
      if (p->forward_only) return DSURF_ERR_BAD_ARG;  // a forward-only plan holds no depth kernels
      DS_CUDA(cudaMemcpy(p->sen[type][q].p, src[q], ns * sizeof(double), cudaMemcpyHostToDevice));
    }
  p->maps_diced = false;
  return DSURF_OK;
}
extern "C" int dsurf_plan_finalize_dispersion(dsurf_plan *p) {
  if (!p) return DSURF_ERR_BAD_ARG;
  const Geom &g = p->g;
  if (p->forward_only) {
    DS_CHECK(dice_maps(p));
    DS_CUDA(cudaStreamSynchronize(p->st));
    p->disp_done = true;
    return DSURF_OK;
  }
  const int brocher = p->depz[p->nz - 2] < 35.0f ? 1 : 0;
  DS_CHECK(launch_coef(p->st, p->vels.p, g.nx, g.ny, p->nz, brocher, p->coe_a.p, p->coe_rho.p));
  for (int t = 0; t < 4; t++)
    if (p->kmaxT[t] > 0)
      DS_CHECK(launch_combine(p->st, p->sen[t][0].p, p->sen[t][1].p, p->sen[t][2].p, p->coe_a.p, p->coe_rho.p,
                              g.nx * g.ny, p->kmaxT[t], p->nz - 1, p->S[t].p));
  DS_CHECK(dice_maps(p));
  DS_CUDA(cudaStreamSynchronize(p->st));
  p->disp_done = true;
  return DSURF_OK;
}

extern "C" int dsurf_plan_set_map(dsurf_plan *p, int type, int period0, const double *pvh) {
  if (!p || type < 0 || type > 3 || period0 < 0 || period0 >= p->pvcols[type]) return DSURF_ERR_BAD_ARG;
  const size_t ncol = (size_t)p->g.nx * p->g.ny;
  DS_CUDA(cudaMemcpy(p->pv[type].p + (size_t)period0 * ncol, pvh, ncol * sizeof(double), cudaMemcpyHostToDevice));
  p->maps_diced = false;
  return DSURF_OK;
}

// Ray-path export: every later dsurf_plan_sweeps call appends the traced ray
// geometry to `file` in the format of the reference's raypath.out.  file == NULL closes it.
extern "C" int dsurf_plan_set_raypath(dsurf_plan *p, const char *file, int max_points) {
  if (!p) return DSURF_ERR_BAD_ARG;
  if (p->path_fh) fclose(p->path_fh);
  p->path_fh = nullptr;
  p->path_cap = 0;
  if (!file) return DSURF_OK;
  if (p->forward_only) return DSURF_ERR_BAD_ARG;  // subroutine synthetic traces no rays
  p->path_fh = fopen(file, "w");
  if (!p->path_fh) {
    set_error(__FILE__, __LINE__, "cannot open the ray-path file for writing");
    return DSURF_ERR_BAD_ARG;
  }
  p->path_cap = max_points > 0 ? max_points : 4 * (p->g.nnx + p->g.nnz);
  p->path_rays = 0;
  return DSURF_OK;
}
extern "C" int64_t dsurf_plan_raypath_count(const dsurf_plan *p) { return p ? p->path_rays : 0; }

extern "C" int dsurf_plan_reset_rows(dsurf_plan *p) {
  if (!p) return DSURF_ERR_BAD_ARG;
  p->nar = 0;
  DS_CUDA(cudaMemset(p->flags.p, 0, 4 * sizeof(int)));
  return DSURF_OK;
}
extern "C" int dsurf_plan_num_gathers(const dsurf_plan *p) { return p ? (int)p->gathers.size() : 0; }
extern "C" int dsurf_plan_num_sweeps(const dsurf_plan *p, int g0, int g1) {
  if (!p) return 0;
  int n = 0;
  for (int g = std::max(g0, 0); g < std::min(g1, (int)p->gathers.size()); g++) n += (p->gathers[g].igr == 1 && !p->forward_only) ? 2 : 1;
  return n;
}
extern "C" int64_t dsurf_plan_nar(const dsurf_plan *p) { return p ? p->nar : 0; }
extern "C" int dsurf_plan_nrows(const dsurf_plan *p) { return p ? p->dall : 0; }

// gfortran list-directed REAL*4 item: one separator blank + G16.9E2
static void ld_real(char *buf, size_t n, float v) {
  int e = 0;
  if (v != 0.0f) {
    char t[32];
    snprintf(t, sizeof t, "%.8E", (double)v);
    e = atoi(strchr(t, 'E') + 1);
  }
  if (e >= -1 && e < 9)
    snprintf(buf, n, " %12.*f    ", 8 - e, (double)v);
  else
    snprintf(buf, n, " %16.8E", (double)v);
}

// raypath.out records of one batch, in the reference's order (gather, receiver): "# nrp" then nrp
// lines "latitude longitude" in degrees, rayx = (pi/2 - rgx)*180/pi, rayz = rgz*180/pi in REAL*4
// (CalSurfG.f90:2276-2283; consumer scripts/plotpath.py)
static int write_paths(dsurf_plan *p, const std::vector<SweepDesc> &hsw, const std::vector<RayDesc> &hrays) {
  const int nrays = (int)hrays.size(), cap = p->path_cap;
  std::vector<int> hn(nrays);
  std::vector<float2> hp((size_t)nrays * cap);
  DS_CUDA(cudaMemcpy(hn.data(), p->path_n.p, nrays * sizeof(int), cudaMemcpyDeviceToHost));
  DS_CUDA(cudaMemcpy(hp.data(), p->path.p, hp.size() * sizeof(float2), cudaMemcpyDeviceToHost));
  char a[40], b[40];
  for (int r = 0; r < nrays; r++) {
    if (!hsw[hrays[r].sweep].do_rays) continue;
    if (hn[r] > cap) {
      set_error(__FILE__, __LINE__, "ray-path export: a ray has more points than max_points");
      return DSURF_ERR_CAPACITY;
    }
    fprintf(p->path_fh, " #%12d\n", hn[r]);
    for (int j = 0; j < hn[r]; j++) {
      const float2 q = hp[(size_t)r * cap + j];
      const float rayx = (kPi / 2 - q.x) * 180.0f / kPi, rayz = q.y * 180.0f / kPi;
      ld_real(a, sizeof a, rayx);
      ld_real(b, sizeof b, rayz);
      fprintf(p->path_fh, "%s%s\n", a, b);
    }
    p->path_rays++;
  }
  fflush(p->path_fh);
  return DSURF_OK;
}

static BatchView batch_view(dsurf_plan *p) {
  BatchView bv;
  bv.node = p->node.p;
  bv.word = p->words ? p->word.p : nullptr;
  bv.box = p->box.p;
  bv.seed = p->seed.p;
  bv.nseed = p->nseed.p;
  bv.slab = eikonal_slab_entries(p->hcap, p->mode);
  bv.wld = p->wld;
  bv.wpx = p->wpx;
  bv.wpz = p->wpz;
  bv.wslot = p->wslot;
  bv.fim_bitmap = p->fim_bitmap.p;
  bv.fim_rw = p->fim_rw.p;
  bv.fim_reg = p->fim_reg.p;
  bv.fim_rect = p->fim_rect.p;
  bv.fim_flag = p->fim_flag.p;
  bv.noder = p->noder.p;
  bv.velr = p->velr.p;
  bv.hent = p->hent.p;
  bv.ristr = p->ristr.p;
  bv.hcap = p->hcap;
  return bv;
}

// runs one batch of sweeps (already described in hsw / hrays); assemble = false keeps fdm.
// Rays are traced and assembled in chunks of at most p->maxrays, in row order.
static int run_batch(dsurf_plan *p, std::vector<SweepDesc> &hsw, std::vector<RayDesc> &hrays,
                     std::vector<float> &hristr, std::vector<int> &hrayS, std::vector<int> &hrayrow,
                     bool assemble, int *launches) {
  const int nsw = (int)hsw.size(), nrays = (int)hrays.size();
  if (nsw == 0) return DSURF_OK;
  cudaStream_t st = p->st;
  DS_CUDA(cudaMemcpyAsync(p->d_sw.p, hsw.data(), nsw * sizeof(SweepDesc), cudaMemcpyHostToDevice, st));
  DS_CUDA(cudaMemcpyAsync(p->ristr.p, hristr.data(), (size_t)nsw * kRefMax * sizeof(float), cudaMemcpyHostToDevice, st));
  float ms;
  for (int attempt = 0; attempt < 3; attempt++) {
    const BatchView bv = batch_view(p);
    cudaEventRecord(p->ev[0], st);
    DS_CHECK(launch_eikonal(st, p->g, p->d_sw.p, nsw, p->veln_all.p, p->velv_all.p, p->risti_c.p, bv, launches, p->mode));
    cudaEventRecord(p->ev[1], st);
    DS_CUDA(cudaMemcpyAsync(hsw.data(), p->d_sw.p, nsw * sizeof(SweepDesc), cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    DS_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]);
    p->ms[2] += ms;
    p->ms[5] += 1;
    bool heap_err = false;
    for (auto &s : hsw)
      if (s.status == DSURF_ERR_HEAP) heap_err = true;
    if (!heap_err) break;
    // narrow band larger than the default slab: grow towards the reference's maxbt = 0.5*Nc
    const long long maxbt = (long long)std::lround(0.5 * (double)p->g.nnx * p->g.nnz);
    if (p->hcap >= maxbt || attempt == 2) {
      set_error(__FILE__, __LINE__, "narrow-band heap exceeded maxbt = NINT(0.5*nnx*nnz)");
      return DSURF_ERR_HEAP;
    }
    p->hcap = (int)std::min<long long>(maxbt, (long long)p->hcap * 8 + 7);
    if (p->hent.reserve((size_t)p->maxslots * (size_t)eikonal_slab_entries(p->hcap, p->mode))) {
      set_error(__FILE__, __LINE__, "cudaMalloc failed (heap growth)");
      return DSURF_ERR_CUDA;
    }
    for (auto &s : hsw) s.status = 0;
    DS_CUDA(cudaMemcpyAsync(p->d_sw.p, hsw.data(), nsw * sizeof(SweepDesc), cudaMemcpyHostToDevice, st));
  }
  const bool want_paths = p->path_fh != nullptr && p->path_cap > 0;
  for (int r0 = 0; r0 < nrays; r0 += p->maxrays) {
    const int nr = std::min(p->maxrays, nrays - r0);
    const BatchView bv = batch_view(p);
    DS_CUDA(cudaMemcpyAsync(p->d_rays.p, hrays.data() + r0, nr * sizeof(RayDesc), cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaMemcpyAsync(p->ray_S.p, hrayS.data() + r0, nr * sizeof(int), cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaMemcpyAsync(p->ray_row.p, hrayrow.data() + r0, nr * sizeof(int), cudaMemcpyHostToDevice, st));
    cudaEventRecord(p->ev[2], st);
    if (want_paths && (p->path.reserve((size_t)nr * p->path_cap) || p->path_n.reserve(nr))) {
      set_error(__FILE__, __LINE__, "cudaMalloc failed (ray-path export buffers)");
      return DSURF_ERR_CUDA;
    }
    DS_CHECK(launch_rays(st, p->g, p->d_sw.p, p->d_rays.p, nr, p->veln_all.p, bv, p->dsurf.p, p->fdm.p,
                         p->bbox.p, p->flags.p + 1, p->flags.p, want_paths ? p->path.p : nullptr,
                         want_paths ? p->path_n.p : nullptr, p->path_cap));
    if (launches) *launches += 1;
    cudaEventRecord(p->ev[3], st);
    if (assemble) {
      DS_CHECK(launch_assembly(st, p->g, p->nz, p->fdm.p, p->bbox.p, nr, p->ray_S.p, p->S_ptr.p,
                               p->S_stride.p, p->ray_row.p, p->cnt, p->wide, p->loff, p->roff, p->lpos, p->lval,
                               p->lcnt, p->tmp, p->rw, p->col, p->rowidx, p->nar, p->flags.p, launches));
    }
    cudaEventRecord(p->ev[4], st);
    DS_CUDA(cudaStreamSynchronize(st));
    DS_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]);
    p->ms[3] += ms;
    cudaEventElapsedTime(&ms, p->ev[3], p->ev[4]);
    p->ms[4] += ms;
    if (want_paths) {
      std::vector<RayDesc> chunk(hrays.begin() + r0, hrays.begin() + r0 + nr);
      DS_CHECK(write_paths(p, hsw, chunk));
    }
  }
  return DSURF_OK;
}

static int append_gather(dsurf_plan *p, int gidx, int only_ig, std::vector<SweepDesc> &hsw,
                         std::vector<RayDesc> &hrays, std::vector<float> &hristr, std::vector<int> &hrayS,
                         std::vector<int> &hrayrow) {
  const GatherInfo &gi = p->gathers[gidx];
  const int igroup = (gi.igr == 1 && !p->forward_only) ? 2 : 1;
  for (int ig = 1; ig <= igroup; ig++) {
    if (only_ig && ig != only_ig) continue;
    SweepDesc d;
    memset(&d, 0, sizeof(d));
    hristr.resize(hristr.size() + kRefMax, 0.0f);
    int rc = make_sweep(p->g, gi.scx, gi.scz, d, hristr.data() + hristr.size() - kRefMax);
    if (rc != DSURF_OK) return rc;
    int arr = gi.type;
    if (ig == 2) arr = (gi.type == 1) ? 0 : 2;
    d.map = p->mapslot[arr][gi.per - 1];
    d.gather = gidx;
    d.first_row = gi.first_row;
    d.nrc = gi.nrc;
    d.do_times = (ig == 1) ? 1 : 0;
    d.do_rays = (gi.igr == 0 || ig == 2) ? 1 : 0;
    if (p->forward_only) d.do_rays = 0;
    const int slot = (int)hsw.size();
    hsw.push_back(d);
    for (int r = 0; r < gi.nrc; r++) {
      RayDesc rd;
      rd.sweep = slot;
      rd.row = gi.first_row + r;
      rd.rcx = p->rcx[gi.first_row + r];
      rd.rcz = p->rcz[gi.first_row + r];
      rd.sin_rcx = sinf(rd.rcx);
      hrays.push_back(rd);
      hrayS.push_back(p->S_id_base[gi.type] + (gi.knumi - p->S_id_base[gi.type] - 1));  // sen_*(:,knumi,:)
      hrayrow.push_back(rd.row);
    }
  }
  return DSURF_OK;
}

extern "C" int dsurf_plan_sweeps(dsurf_plan *p, int g0, int g1) {
  if (!p) return DSURF_ERR_BAD_ARG;
  DS_CHECK(ensure_device());
  if (!p->maps_diced) DS_CHECK(dice_maps(p));
  g0 = std::max(g0, 0);
  g1 = std::min(g1, (int)p->gathers.size());
  for (int i = 2; i < 8; i++) p->ms[i] = 0;
  if (p->nar == 0) p->last_g0 = g0;  // rows accumulate until dsurf_plan_reset_rows
  p->last_g1 = g1;
  p->G_nar = -1;
  int launches = 0, nsolved = 0;
  cudaEventRecord(p->evt0, p->st);
  std::vector<SweepDesc> hsw;
  std::vector<RayDesc> hrays;
  std::vector<float> hristr;
  std::vector<int> hrayS, hrayrow;
  int g = g0;
  while (g < g1) {
    hsw.clear();
    hrays.clear();
    hristr.clear();
    hrayS.clear();
    hrayrow.clear();
    while (g < g1) {
      const GatherInfo &gi = p->gathers[g];
      const int ns = (gi.igr == 1 && !p->forward_only) ? 2 : 1;
      if (!hsw.empty() && (int)hsw.size() + ns > p->maxslots) break;
      DS_CHECK(append_gather(p, g, 0, hsw, hrays, hristr, hrayS, hrayrow));
      g++;
    }
    // group gathers trace rays only in pass 2 and times only in pass 1: drop the unused RayDescs
    // per sweep inside the kernels via do_times/do_rays (both kept so that rows stay aligned).
    DS_CHECK(run_batch(p, hsw, hrays, hristr, hrayS, hrayrow, !p->forward_only, &launches));
    nsolved += (int)hsw.size();
  }
  cudaEventRecord(p->evt1, p->st);
  int hflags[4];
  DS_CUDA(cudaMemcpy(hflags, p->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost));
  {
    float t = 0;
    cudaEventSynchronize(p->evt1);
    cudaEventElapsedTime(&t, p->evt0, p->evt1);
    p->ms_sweeps_total = t;
  }
  p->ms[6] = launches;
  p->ms[7] = nsolved;
  if (hflags[0] != 0) return hflags[0];
  return DSURF_OK;
}

extern "C" int dsurf_plan_download(dsurf_plan *p, int *iw_rows, float *rw, int *col, float *dsurf, int *rbint) {
  if (!p) return DSURF_ERR_BAD_ARG;
  if (p->nar > 0) {
    if (iw_rows) DS_CUDA(cudaMemcpy(iw_rows, p->rowidx.p, p->nar * sizeof(int), cudaMemcpyDeviceToHost));
    if (rw) DS_CUDA(cudaMemcpy(rw, p->rw.p, p->nar * sizeof(float), cudaMemcpyDeviceToHost));
    if (col) DS_CUDA(cudaMemcpy(col, p->col.p, p->nar * sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (dsurf && p->dall > 0) DS_CUDA(cudaMemcpy(dsurf, p->dsurf.p, p->dall * sizeof(float), cudaMemcpyDeviceToHost));
  if (rbint) {
    int f[4];
    DS_CUDA(cudaMemcpy(f, p->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost));
    *rbint = f[1];
  }
  return DSURF_OK;
}

extern "C" double dsurf_plan_last_sweeps_ms(const dsurf_plan *p) { return p ? p->ms_sweeps_total : 0.0; }

extern "C" int dsurf_plan_timings(const dsurf_plan *p, double *ms8) {
  if (!p || !ms8) return DSURF_ERR_BAD_ARG;
  for (int i = 0; i < 8; i++) ms8[i] = p->ms[i];
  return DSURF_OK;
}

extern "C" int dsurf_plan_get_dispersion(dsurf_plan *p, int type, double *pv, double *sen_vs, double *sen_vp,
                                         double *sen_rho) {
  if (!p || type < 0 || type > 3) return DSURF_ERR_BAD_ARG;
  const size_t ncol = (size_t)p->g.nx * p->g.ny;
  if (pv) DS_CUDA(cudaMemcpy(pv, p->pv[type].p, ncol * p->pvcols[type] * sizeof(double), cudaMemcpyDeviceToHost));
  const size_t ns = ncol * p->kmaxT[type] * p->nz;
  double *dst[3] = {sen_vs, sen_vp, sen_rho};
  for (int q = 0; q < 3; q++)
    if (dst[q] && ns > 0) DS_CUDA(cudaMemcpy(dst[q], p->sen[type][q].p, ns * sizeof(double), cudaMemcpyDeviceToHost));
  return DSURF_OK;
}

extern "C" int dsurf_plan_debug_sweep(dsurf_plan *p, int gidx, int ig, float *veln, float *ttn, float *ttnr,
                                      int *nstsr, float *rgeom, float *fdm) {
  if (!p || gidx < 0 || gidx >= (int)p->gathers.size() || ig < 1 || ig > 2) return DSURF_ERR_BAD_ARG;
  DS_CHECK(ensure_device());
  if (!p->maps_diced) DS_CHECK(dice_maps(p));
  std::vector<SweepDesc> hsw;
  std::vector<RayDesc> hrays;
  std::vector<float> hristr;
  std::vector<int> hrayS, hrayrow;
  DS_CHECK(append_gather(p, gidx, ig, hsw, hrays, hristr, hrayS, hrayrow));
  if (hsw.empty()) return DSURF_ERR_BAD_ARG;
  hsw[0].do_rays = 1;
  int launches = 0;
  DS_CHECK(run_batch(p, hsw, hrays, hristr, hrayS, hrayrow, false, &launches));
  const Geom &g = p->g;
  const size_t Nc = (size_t)g.nnx * g.nnz;
  const SweepDesc &d = hsw[0];
  if (veln) DS_CUDA(cudaMemcpy(veln, p->veln_all.p + (size_t)d.map * Nc, Nc * sizeof(float), cudaMemcpyDeviceToHost));
  if (ttn && p->words) {  // every node alive: word = time; strip the 3-node frame
    const size_t wld = (size_t)p->wld;
    DS_CUDA(cudaMemcpy2D(ttn, g.nnz * sizeof(float), p->word.p + p->wpx * wld + p->wpz, wld * sizeof(float), g.nnz * sizeof(float),
                         g.nnx, cudaMemcpyDeviceToHost));
  } else if (ttn) {
    std::vector<int2> tmp(Nc);
    DS_CUDA(cudaMemcpy(tmp.data(), p->node.p, Nc * sizeof(int2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < Nc; i++) memcpy(&ttn[i], &tmp[i].x, 4);
  }
  if (ttnr || nstsr) {
    const size_t nr = (size_t)d.nrnx * d.nrnz;
    std::vector<int2> tmp(nr);
    DS_CUDA(cudaMemcpy(tmp.data(), p->noder.p, nr * sizeof(int2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nr; i++) {
      if (ttnr) memcpy(&ttnr[i], &tmp[i].x, 4);
      if (nstsr) nstsr[i] = tmp[i].y;
    }
  }
  if (rgeom) {
    rgeom[0] = d.gorx;
    rgeom[1] = d.gorz;
    rgeom[2] = g.drnx;
    rgeom[3] = g.drnz;
    rgeom[4] = (float)d.nrnx;
    rgeom[5] = (float)d.nrnz;
  }
  const size_t fsz = (size_t)(g.nvz + 2) * (g.nvx + 2);
  if (fdm) DS_CUDA(cudaMemcpy(fdm, p->fdm.p, fsz * hrays.size() * sizeof(float), cudaMemcpyDeviceToHost));
  DS_CUDA(cudaMemset(p->fdm.p, 0, fsz * hrays.size() * sizeof(float)));
  return DSURF_OK;
}

// ------------------------------------------------------------------------------- CalSurfG drop-in
extern "C" int dsurf_calsurfg(int nx, int ny, int nz, int nparpi, const float *vels, int *iw, float *rw,
                              int *col, float *dsurf, float goxdf, float gozdf, float dvxdf, float dvzdf,
                              int kmaxRc, int kmaxRg, int kmaxLc, int kmaxLg, const double *tRc,
                              const double *tRg, const double *tLc, const double *tLg, const int *wavetype,
                              const int *igrt, const int *periods, const float *depz, float minthk,
                              const float *scxf, const float *sczf, const float *rcxf, const float *rczf,
                              const int *nrc1, const int *nsrcsurf1, int kmax, int nsrcsurf, int nrcf,
                              int64_t maxnar, int *nar, int *rbint) {
  (void)nparpi;
  dsurf_plan *p = nullptr;
  DS_CHECK(dsurf_plan_create(&p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, kmaxRg, kmaxLc, kmaxLg, tRc,
                             tRg, tLc, tLg, wavetype, igrt, periods, depz, minthk, scxf, sczf, rcxf, rczf, nrc1,
                             nsrcsurf1, kmax, nsrcsurf, nrcf));
  int rc = dsurf_plan_dispersion(p);
  if (rc == DSURF_OK) rc = dsurf_plan_reset_rows(p);
  if (rc == DSURF_OK) rc = dsurf_plan_sweeps(p, 0, dsurf_plan_num_gathers(p));
  if (rc == DSURF_OK && maxnar >= 0 && p->nar > maxnar) {
    set_error(__FILE__, __LINE__, "nar exceeds the caller's COO capacity (increase sparsity fraction)");
    rc = DSURF_ERR_CAPACITY;
  }
  if (rc == DSURF_OK) rc = dsurf_plan_download(p, iw ? iw + 1 : nullptr, rw, col, dsurf, rbint);
  if (rc == DSURF_OK && nar) *nar = (int)p->nar;
  dsurf_plan_destroy(p);
  return rc;
}

extern "C" void dsurf_fatal_(const int *rc);

extern "C" void calsurfg_(const int *nx, const int *ny, const int *nz, const int *nparpi, const float *vels,
                          int *iw, float *rw, int *col, float *dsurf, const float *goxdf, const float *gozdf,
                          const float *dvxdf, const float *dvzdf, const int *kmaxRc, const int *kmaxRg,
                          const int *kmaxLc, const int *kmaxLg, const double *tRc, const double *tRg,
                          const double *tLc, const double *tLg, const int *wavetype, const int *igrt,
                          const int *periods, const float *depz, const float *minthk, const float *scxf,
                          const float *sczf, const float *rcxf, const float *rczf, const int *nrc1,
                          const int *nsrcsurf1, const int *kmax, const int *nsrcsurf, const int *nrcf, int *nar) {
  int rbint = 0;
  // the Fortran caller does not pass its COO capacity (maxnar); it is unchecked in the reference too
  int rc = dsurf_calsurfg(*nx, *ny, *nz, *nparpi, vels, iw, rw, col, dsurf, *goxdf, *gozdf, *dvxdf, *dvzdf, *kmaxRc,
                          *kmaxRg, *kmaxLc, *kmaxLg, tRc, tRg, tLc, tLg, wavetype, igrt, periods, depz, *minthk,
                          scxf, sczf, rcxf, rczf, nrc1, nsrcsurf1, *kmax, *nsrcsurf, *nrcf, -1, nar, &rbint);
  if (rc != DSURF_OK) dsurf_fatal_(&rc);
  if (rbint) {  // CalSurfG.f90:1447-1454
    printf(" Note that at least one two-point ray path\n tracked along the boundary of the model.\n"
           " This class of path is unlikely to be\n a true path, and it is STRONGLY RECOMMENDED\n"
           " that you adjust the dimensions of your grid\n to prevent this from occurring.\n");
  }
}

// ------------------------------------------------------------------------------- synthetic drop-in
// subroutine synthetic (CalSurfG.f90:2412-2865): forward times through `vels` on the gd = 5
// propagation grid, obst(i) = t + t*gaussian()*noiselevel, and the velmap2d{Rc,Rg,Lc,Lg}.dat maps.
namespace {
// gaussian.f90: Box-Muller on two uniform deviates, second deviate discarded (use_last is reset on
// every call).  The reference draws from gfortran's random_number, whose stream is not reproducible
// outside libgfortran; this is a splitmix64-seeded xoshiro256** -- same distribution, other draws.
struct Rng {
  uint64_t s[4];
  explicit Rng(uint64_t seed) {
    for (auto &v : s) {
      seed += 0x9e3779b97f4a7c15ull;
      uint64_t z = seed;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      v = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  float uniform() {  // [0, 1)
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return (float)(r >> 40) * (1.0f / 16777216.0f);
  }
  float gaussian() {
    float x1 = 0, w = 2.0f;
    while (w >= 1.0f) {
      const float n1 = uniform(), n2 = uniform();
      x1 = 2.0f * n1 - 1.0f;
      const float x2 = 2.0f * n2 - 1.0f;
      w = x1 * x1 + x2 * x2;
    }
    w = powf((-2.0f * logf(w)) / w, 0.5f);
    return x1 * w;
  }
};
}  // namespace

extern "C" int dsurf_synthetic(int nx, int ny, int nz, int nparpi, const float *vels, float *obst, float goxdf,
                               float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc, int kmaxLg,
                               const double *tRc, const double *tRg, const double *tLc, const double *tLg,
                               const int *wavetype, const int *igrt, const int *periods, const float *depz,
                               float minthk, const float *scxf, const float *sczf, const float *rcxf,
                               const float *rczf, const int *nrc1, const int *nsrcsurf1, int kmax, int nsrcsurf,
                               int nrcf, float noiselevel, const char *outdir, uint64_t seed, int *rbint) {
  (void)nparpi;
  if (!obst) return DSURF_ERR_BAD_ARG;
  dsurf_plan *p = nullptr;
  DS_CHECK(dsurf_plan_create_forward(&p, nx, ny, nz, vels, goxdf, gozdf, dvxdf, dvzdf, kmaxRc, kmaxRg, kmaxLc, kmaxLg,
                                     tRc, tRg, tLc, tLg, wavetype, igrt, periods, depz, minthk, scxf, sczf, rcxf, rczf,
                                     nrc1, nsrcsurf1, kmax, nsrcsurf, nrcf));
  int rc = dsurf_plan_dispersion(p);
  if (rc == DSURF_OK && outdir) {  // velmap2dXX.dat, format (5f8.4) (:2557-2613)
    static const char *names[4] = {"velmap2dRc.dat", "velmap2dRg.dat", "velmap2dLc.dat", "velmap2dLg.dat"};
    const size_t ncol = (size_t)nx * ny;
    for (int t = 0; t < 4 && rc == DSURF_OK; t++) {
      const int kt = p->kmaxT[t];
      if (kt <= 0) continue;
      std::vector<double> pv(ncol * kt);
      if (cudaMemcpy(pv.data(), p->pv[t].p, pv.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = DSURF_ERR_CUDA;
        break;
      }
      std::string path = std::string(outdir);
      if (!path.empty() && path.back() != '/') path += "/";
      path += names[t];
      FILE *f = fopen(path.c_str(), "w");
      if (!f) {
        set_error(__FILE__, __LINE__, "cannot open velmap2d output file");
        rc = DSURF_ERR_BAD_ARG;
        break;
      }
      for (int k = 1; k <= kt; k++)
        for (int j = 1; j <= ny - 2; j++)
          for (int i = 1; i <= nx - 2; i++)
            fprintf(f, "%8.4f%8.4f%8.4f%8.4f\n", gozdf + (float)(j - 1) * dvzdf, goxdf - (float)(i - 1) * dvxdf,
                    p->tper[t][k - 1], pv[(size_t)(k - 1) * ncol + (size_t)(j + 1) * nx + i]);
      fclose(f);
    }
  }
  if (rc == DSURF_OK) rc = dsurf_plan_reset_rows(p);
  if (rc == DSURF_OK) rc = dsurf_plan_sweeps(p, 0, dsurf_plan_num_gathers(p));
  if (rc == DSURF_OK) rc = dsurf_plan_download(p, nullptr, nullptr, nullptr, obst, rbint);
  if (rc == DSURF_OK && noiselevel != 0.0f) {
    Rng rng(seed);
    for (int i = 0; i < p->dall; i++) obst[i] = obst[i] + obst[i] * rng.gaussian() * noiselevel;
  }
  dsurf_plan_destroy(p);
  return rc;
}

extern "C" void dsurf_fatal_(const int *rc);
extern "C" void synthetic_(const int *nx, const int *ny, const int *nz, const int *nparpi, const float *vels,
                           float *obst, const float *goxdf, const float *gozdf, const float *dvxdf,
                           const float *dvzdf, const int *kmaxRc, const int *kmaxRg, const int *kmaxLc,
                           const int *kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                           const double *tLg, const int *wavetype, const int *igrt, const int *periods,
                           const float *depz, const float *minthk, const float *scxf, const float *sczf,
                           const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                           const int *kmax, const int *nsrcsurf, const int *nrcf, const float *noiselevel) {
  int rbint = 0;
  int rc = dsurf_synthetic(*nx, *ny, *nz, *nparpi, vels, obst, *goxdf, *gozdf, *dvxdf, *dvzdf, *kmaxRc, *kmaxRg,
                           *kmaxLc, *kmaxLg, tRc, tRg, tLc, tLg, wavetype, igrt, periods, depz, *minthk, scxf, sczf,
                           rcxf, rczf, nrc1, nsrcsurf1, *kmax, *nsrcsurf, *nrcf, *noiselevel, ".", 20150131ull,
                           &rbint);
  if (rc != DSURF_OK) dsurf_fatal_(&rc);
  if (rbint) {  // :2843-2850 (printed per gather in the reference)
    printf(" Note that at least one two-point ray path\n tracked along the boundary of the model.\n"
           " This class of path is unlikely to be\n a true path, and it is STRONGLY RECOMMENDED\n"
           " that you adjust the dimensions of your grid\n to prevent this from occurring.\n");
  }
}

// ------------------------------------------------------------------- multi-GPU exchange
namespace dsurf {
int nccl_allgather_bytes(void *comm, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st);
int nccl_allgatherv_words(void *comm, int rank, int nranks, const void *send, void *recv, const long long *off4,
                          const long long *count4, cudaStream_t st);
}
static int first_row_of(const dsurf_plan *p, int g) { return g < (int)p->gathers.size() ? p->gathers[g].first_row : p->dall; }

// Ranks hold contiguous gather blocks in rank order (the loop nest CalSurfG.f90:1144-1145 cut into blocks), so the
// reference's output is the concatenation of the ranks' outputs: one all-gather of the predicted times (every rank
// needs them for the percentile weights of main.f90:361-376) and, when the caller wants the full COO that
// main.f90:355-359 receives, one of the row blocks -- counts first, then grouped broadcasts straight into place.
extern "C" int dsurf_plan_allgather(dsurf_plan *p, void *comm, int rank, int nranks, int want_coo, int64_t *nar_total) {
  if (!p || !comm || rank < 0 || rank >= nranks || nranks > 64) return DSURF_ERR_BAD_ARG;
  cudaStream_t st = p->st;
  cudaEventRecord(p->ev[0], st);
  if (p->G_cnt.reserve(4 * 66)) return DSURF_ERR_CUDA;
  long long mine[4] = {p->nar, first_row_of(p, p->last_g0), first_row_of(p, p->last_g1), 0};
  DS_CUDA(cudaMemcpyAsync(p->G_cnt.p + 4 * 64, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  DS_CHECK(nccl_allgather_bytes(comm, p->G_cnt.p + 4 * 64, p->G_cnt.p, sizeof(mine), st));
  long long all[4 * 64];
  DS_CUDA(cudaMemcpyAsync(all, p->G_cnt.p, sizeof(long long) * 4 * nranks, cudaMemcpyDeviceToHost, st));
  DS_CUDA(cudaStreamSynchronize(st));
  long long off[64], cnt[64], roff[64], rcnt[64], tot = 0;
  for (int r = 0; r < nranks; r++) {
    off[r] = tot;
    cnt[r] = all[4 * r];
    tot += cnt[r];
    roff[r] = all[4 * r + 1];
    rcnt[r] = all[4 * r + 2] - all[4 * r + 1];
    if (r > 0 && roff[r] != all[4 * (r - 1) + 2]) {
      set_error(__FILE__, __LINE__, "allgather: the ranks' gather blocks are not contiguous in rank order");
      return DSURF_ERR_BAD_ARG;
    }
  }
  // predicted times: in place, every rank's slice of the full-length vector
  DS_CHECK(nccl_allgatherv_words(comm, rank, nranks, p->dsurf.p + roff[rank], p->dsurf.p, roff, rcnt, st));
  p->G_nar = -1;
  if (want_coo) {
    if (p->G_rw.reserve((size_t)std::max<long long>(tot, 1)) || p->G_col.reserve((size_t)std::max<long long>(tot, 1)) ||
        p->G_row.reserve((size_t)std::max<long long>(tot, 1))) {
      set_error(__FILE__, __LINE__, "cudaMalloc failed (gathered COO)");
      return DSURF_ERR_CUDA;
    }
    DS_CHECK(nccl_allgatherv_words(comm, rank, nranks, p->rw.p, p->G_rw.p, off, cnt, st));
    DS_CHECK(nccl_allgatherv_words(comm, rank, nranks, p->col.p, p->G_col.p, off, cnt, st));
    DS_CHECK(nccl_allgatherv_words(comm, rank, nranks, p->rowidx.p, p->G_row.p, off, cnt, st));
    p->G_nar = tot;
  }
  cudaEventRecord(p->ev[1], st);
  DS_CUDA(cudaStreamSynchronize(st));
  DS_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]);
  p->ms_gather = ms;
  if (nar_total) *nar_total = tot;
  return DSURF_OK;
}
extern "C" double dsurf_plan_last_gather_ms(const dsurf_plan *p) { return p ? p->ms_gather : 0.0; }

extern "C" int dsurf_plan_download_gathered(dsurf_plan *p, int *iw_rows, float *rw, int *col) {
  if (!p || p->G_nar < 0) return DSURF_ERR_BAD_ARG;
  const size_t n = (size_t)p->G_nar;
  if (n == 0) return DSURF_OK;
  if (iw_rows) DS_CUDA(cudaMemcpy(iw_rows, p->G_row.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  if (rw) DS_CUDA(cudaMemcpy(rw, p->G_rw.p, n * sizeof(float), cudaMemcpyDeviceToHost));
  if (col) DS_CUDA(cudaMemcpy(col, p->G_col.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  return DSURF_OK;
}

// position-keyed 64-bit digest of a COO (order-sensitive): equal digests on 1 and on N GPUs <=> same triplets in the same order
__global__ void k_coo_digest(const int *__restrict__ row, const int *__restrict__ col, const float *__restrict__ rw,
                             long long n, unsigned long long *out) {
  unsigned long long acc = 0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)k * 0x9E3779B97F4A7C15ull;
    h ^= ((unsigned long long)(unsigned)row[k] << 32) | (unsigned)col[k];
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= (unsigned long long)__float_as_uint(rw[k]) * 0x94D049BB133111EBull;
    h ^= h >> 29;
    acc += h;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
extern "C" int dsurf_plan_digest(dsurf_plan *p, int gathered, uint64_t *digest, int64_t *n_out) {
  if (!p || !digest || (gathered && p->G_nar < 0)) return DSURF_ERR_BAD_ARG;
  const long long n = gathered ? p->G_nar : p->nar;
  if (p->G_cnt.reserve(4 * 66)) return DSURF_ERR_CUDA;
  unsigned long long *d = reinterpret_cast<unsigned long long *>(p->G_cnt.p + 4 * 64 + 4);
  DS_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), p->st));
  if (n > 0)
    k_coo_digest<<<sm_count() * 8, 256, 0, p->st>>>(gathered ? p->G_row.p : p->rowidx.p, gathered ? p->G_col.p : p->col.p,
                                                    gathered ? p->G_rw.p : p->rw.p, n, d);
  unsigned long long h = 0;
  DS_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, p->st));
  DS_CUDA(cudaStreamSynchronize(p->st));
  *digest = h;
  if (n_out) *n_out = n;
  return DSURF_OK;
}

// ------------------------------------------------------------------- device-resident host glue
// main.f90:361-466 on the device: residual, percentile outlier weights, row scaling, DWS statistics,
// smoothing rows appended behind the data rows, then the LSMR system is built straight from the
// plan's COO in HBM.
extern "C" int dsurf_lsmr_create_from_plan(dsurf_lsmr_sys **sys, dsurf_plan *p, const float *obst, float threshold0,
                                           float weight) {
  if (!sys || !p || !obst) return DSURF_ERR_BAD_ARG;
  DS_CHECK(ensure_device());
  const int dall = p->dall;
  const int maxvp = p->g.nvx * p->g.nvz * (p->nz - 1);
  if (dall < 2) return DSURF_ERR_BAD_ARG;
  cudaStream_t st = p->st;
  if (!p->g_srow.p || p->g_weight != weight) {  // smoothing rows: static per geometry and weight
    std::vector<int> r, c;
    std::vector<float> v;
    glue_smoothing_rows(p->g.nx, p->g.ny, p->nz, dall, weight, r, c, v, &p->g_count3);
    p->g_nsm = (long long)r.size();
    if (p->g_srow.reserve(r.size()) || p->g_scol.reserve(r.size()) || p->g_sval.reserve(r.size())) {
      set_error(__FILE__, __LINE__, "cudaMalloc failed (smoothing rows)");
      return DSURF_ERR_CUDA;
    }
    DS_CUDA(cudaMemcpy(p->g_srow.p, r.data(), r.size() * sizeof(int), cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(p->g_scol.p, c.data(), c.size() * sizeof(int), cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(p->g_sval.p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    p->g_weight = weight;
  }
  const int m = dall + p->g_count3;
  if (p->g_obst.reserve(dall) || p->g_cbst.reserve(m) || p->g_datw.reserve(dall) || p->g_norm.reserve((size_t)maxvp + 2)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (glue vectors)");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cudaMemcpyAsync(p->g_obst.p, obst, (size_t)dall * sizeof(float), cudaMemcpyHostToDevice, st));
  DS_CUDA(cudaMemsetAsync(p->g_cbst.p, 0, (size_t)m * sizeof(float), st));
  DS_CHECK(glue_apply(st, dall, maxvp, p->nar, p->g_obst.p, p->dsurf.p, threshold0, p->rowidx.p, p->col.p, p->rw.p,
                      p->g_cbst.p, p->g_datw.p, p->g_norm.p, p->g_sorted, p->g_tmp, &p->g_stats));
  // append the smoothing rows behind the data rows (in place: the tail is scratch of the plan)
  const long long nnz = p->nar + p->g_nsm;
  if (nnz >= (1ll << 31)) {
    set_error(__FILE__, __LINE__, "nar exceeds the int32 triplet count of the reference boundary");
    return DSURF_ERR_CAPACITY;
  }
  if (p->rw.reserve((size_t)nnz, true, st) || p->col.reserve((size_t)nnz, true, st) ||
      p->rowidx.reserve((size_t)nnz, true, st)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (COO growth for smoothing rows)");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cudaMemcpyAsync(p->rowidx.p + p->nar, p->g_srow.p, (size_t)p->g_nsm * sizeof(int), cudaMemcpyDeviceToDevice, st));
  DS_CUDA(cudaMemcpyAsync(p->col.p + p->nar, p->g_scol.p, (size_t)p->g_nsm * sizeof(int), cudaMemcpyDeviceToDevice, st));
  DS_CUDA(cudaMemcpyAsync(p->rw.p + p->nar, p->g_sval.p, (size_t)p->g_nsm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  DS_CUDA(cudaStreamSynchronize(st));
  p->g_valid = true;
  return lsmr_sys_create_dev(sys, m, maxvp, nnz, p->rowidx.p, p->col.p, p->rw.p, p->g_cbst.p);
}

// Row-partitioned variant for the distributed LSMR: this rank keeps the data rows it produced
// (gathers [last_g0, last_g1)) plus a contiguous share of the smoothing rows; rows are renumbered 1..m_local.  The
// predicted times of ALL rows must be present (dsurf_plan_allgather) because the outlier weights use the percentiles
// of every residual (main.f90:361-376).  With nranks == 1 this is dsurf_lsmr_create_from_plan.
__global__ void k_shift_rows(int *rows, long long n, int delta) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) rows[k] += delta;
}
extern "C" int dsurf_lsmr_create_from_plan_shard(dsurf_lsmr_sys **sys, dsurf_plan *p, const float *obst, float threshold0,
                                                 float weight, int rank, int nranks, int *m_local, int64_t *nnz_local) {
  if (!sys || !p || !obst || rank < 0 || rank >= nranks) return DSURF_ERR_BAD_ARG;
  DS_CHECK(ensure_device());
  const int dall = p->dall;
  const int maxvp = p->g.nvx * p->g.nvz * (p->nz - 1);
  if (dall < 2) return DSURF_ERR_BAD_ARG;
  cudaStream_t st = p->st;
  const int R0 = first_row_of(p, p->last_g0), R1 = first_row_of(p, p->last_g1);
  std::vector<int> r, c;
  std::vector<float> v;
  int count3 = 0;
  glue_smoothing_rows(p->g.nx, p->g.ny, p->nz, dall, weight, r, c, v, &count3);
  const int base = count3 / nranks, rem = count3 % nranks;
  const int s0 = rank * base + std::min(rank, rem), s1 = s0 + base + (rank < rem ? 1 : 0);
  size_t k0 = 0, k1 = 0;  // triplets of smoothing rows (s0, s1]: rows are ascending
  while (k0 < r.size() && r[k0] <= dall + s0) k0++;
  k1 = k0;
  while (k1 < r.size() && r[k1] <= dall + s1) k1++;
  const long long nsm = (long long)(k1 - k0);
  const int mdata = R1 - R0, m = mdata + (s1 - s0);
  for (size_t k = k0; k < k1; k++) r[k] = r[k] - dall - s0 + mdata;
  if (p->g_obst.reserve(dall) || p->g_cbst.reserve((size_t)dall + count3) || p->g_datw.reserve(dall) ||
      p->g_norm.reserve((size_t)maxvp + 2)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (glue vectors)");
    return DSURF_ERR_CUDA;
  }
  DS_CUDA(cudaMemcpyAsync(p->g_obst.p, obst, (size_t)dall * sizeof(float), cudaMemcpyHostToDevice, st));
  DS_CUDA(cudaMemsetAsync(p->g_cbst.p, 0, ((size_t)dall + count3) * sizeof(float), st));
  DS_CHECK(glue_apply(st, dall, maxvp, p->nar, p->g_obst.p, p->dsurf.p, threshold0, p->rowidx.p, p->col.p, p->rw.p,
                      p->g_cbst.p, p->g_datw.p, p->g_norm.p, p->g_sorted, p->g_tmp, &p->g_stats));
  const long long nnz = p->nar + nsm;
  if (nnz >= (1ll << 31)) {
    set_error(__FILE__, __LINE__, "nar exceeds the int32 triplet count of the reference boundary");
    return DSURF_ERR_CAPACITY;
  }
  if (p->rw.reserve((size_t)nnz, true, st) || p->col.reserve((size_t)nnz, true, st) || p->rowidx.reserve((size_t)nnz, true, st)) {
    set_error(__FILE__, __LINE__, "cudaMalloc failed (COO growth for smoothing rows)");
    return DSURF_ERR_CUDA;
  }
  if (nsm > 0) {
    DS_CUDA(cudaMemcpyAsync(p->rowidx.p + p->nar, r.data() + k0, (size_t)nsm * sizeof(int), cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaMemcpyAsync(p->col.p + p->nar, c.data() + k0, (size_t)nsm * sizeof(int), cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaMemcpyAsync(p->rw.p + p->nar, v.data() + k0, (size_t)nsm * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  // right-hand side: residuals of the local data rows, zeros for the smoothing rows (main.f90:461)
  DevBuf<float> bl;
  if (bl.reserve((size_t)std::max(m, 1))) return DSURF_ERR_CUDA;
  DS_CUDA(cudaMemsetAsync(bl.p, 0, (size_t)std::max(m, 1) * sizeof(float), st));
  if (mdata > 0) DS_CUDA(cudaMemcpyAsync(bl.p, p->g_cbst.p + R0, (size_t)mdata * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (p->nar > 0 && R0 != 0) k_shift_rows<<<sm_count() * 4, 256, 0, st>>>(p->rowidx.p, p->nar, -R0);
  DS_CUDA(cudaStreamSynchronize(st));
  p->g_valid = true;
  p->g_count3 = count3;
  p->g_nsm = nsm;
  const int rc = lsmr_sys_create_dev(sys, m, maxvp, nnz, p->rowidx.p, p->col.p, p->rw.p, bl.p);
  if (p->nar > 0 && R0 != 0) k_shift_rows<<<sm_count() * 4, 256, 0, st>>>(p->rowidx.p, p->nar, R0);  // global numbering again
  DS_CUDA(cudaStreamSynchronize(st));
  if (m_local) *m_local = m;
  if (nnz_local) *nnz_local = nnz;
  return rc;
}

extern "C" int dsurf_plan_glue_results(dsurf_plan *p, float *cbst, float *datweight, float *stats4, int *m_out,
                                       int64_t *nar_out) {
  if (!p || !p->g_valid) return DSURF_ERR_BAD_ARG;
  if (cbst) DS_CUDA(cudaMemcpy(cbst, p->g_cbst.p, (size_t)p->dall * sizeof(float), cudaMemcpyDeviceToHost));
  if (datweight) DS_CUDA(cudaMemcpy(datweight, p->g_datw.p, (size_t)p->dall * sizeof(float), cudaMemcpyDeviceToHost));
  if (stats4) {
    stats4[0] = p->g_stats.q25;
    stats4[1] = p->g_stats.q75;
    stats4[2] = p->g_stats.maxnorm;
    stats4[3] = p->g_stats.averdws;
  }
  if (m_out) *m_out = p->dall + p->g_count3;
  if (nar_out) *nar_out = p->nar + p->g_nsm;
  return DSURF_OK;
}

// main.f90:518-532 on the device: dv = solution of the last solve of `sys` (clipped to +-0.5 in
// place), model clamped to [minvel, maxvel]; the plan's dispersion/maps are invalidated.
extern "C" int dsurf_plan_update_model(dsurf_plan *p, dsurf_lsmr_sys *sys, float minvel, float maxvel, float *dv_host,
                                       float *vels_host) {
  if (!p || !sys) return DSURF_ERR_BAD_ARG;
  float *d_dv = lsmr_x_dev(sys);
  if (!d_dv) return DSURF_ERR_BAD_ARG;
  const int maxvp = p->g.nvx * p->g.nvz * (p->nz - 1);
  DS_CUDA(cudaStreamSynchronize(p->st));
  DS_CHECK(glue_model_update(p->st, p->vels.p, d_dv, p->g.nx, p->g.ny, p->nz, minvel, maxvel));
  if (dv_host) DS_CUDA(cudaMemcpyAsync(dv_host, d_dv, (size_t)maxvp * sizeof(float), cudaMemcpyDeviceToHost, p->st));
  if (vels_host)
    DS_CUDA(cudaMemcpyAsync(vels_host, p->vels.p, (size_t)p->g.nx * p->g.ny * p->nz * sizeof(float),
                            cudaMemcpyDeviceToHost, p->st));
  DS_CUDA(cudaStreamSynchronize(p->st));
  p->disp_done = false;
  p->maps_diced = false;
  return DSURF_OK;
}
