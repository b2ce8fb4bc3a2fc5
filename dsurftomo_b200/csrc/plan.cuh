// Internal structures shared by the kernels of the forward/sensitivity path.
#pragma once
#include <vector>
#include "common.cuh"

namespace dsurf {

// ---- layer-geometry tables of the dispersion stage (disp.cu)
struct LayerTables {
  int rmax = 0;
  std::vector<float> dflat, facR, facL, wnum, wden;
  std::vector<double> tmp;
  std::vector<int> node;
};
struct LayerTablesDev {
  int rmax = 0;
  DevBuf<float> b_dflat, b_facR, b_facL, b_wnum, b_wden;
  DevBuf<double> b_tmp;
  DevBuf<int> b_node;
  const float *dflat = nullptr, *facR = nullptr, *facL = nullptr, *wnum = nullptr, *wden = nullptr;
  const double *tmp = nullptr;
  const int *node = nullptr;
};
void make_layer_tables(const float *depz, int nz, float minthk0, LayerTables &T);
void make_layer_tables_from_thk(const float *thk, int nlayer, int iflsph, LayerTables &T);
int upload_tables(const LayerTables &T, LayerTablesDev &D);
int run_dispersion(cudaStream_t st, const float *d_vel, int nx, int ny, int nz, const LayerTablesDev &T,
                   int iwave, int igr, int kmax, const double *d_t, bool want_kernels, double *d_pv,
                   double *d_sen_vs, double *d_sen_vp, double *d_sen_rho, DevBuf<double> &cgbuf);

// ---- propagation-grid geometry (globalp, CalSurfG.f90:1032-1063), all REAL*4, host-computed
constexpr int kGd = 8;     // gdx = gdz of CalSurfG (CalSurfG.f90:1032-1033); subroutine synthetic uses 5 (:2497-2498)
constexpr int kSgdl = 8;   // source grid dicing level (:1035)
constexpr int kSgs = 8;    // extent of refined source grid (:1036)
constexpr int kRefMax = 2 * kSgs * kSgdl + 1;  // 129

struct Geom {
  int nx, ny, nvx, nvz, nnx, nnz;
  int gd;                    // grid dicing gdx = gdz: 8 for CalSurfG, 5 for synthetic
  float gox, goz, dnx, dnz, dvx, dvz, earth;
  float drnx, drnz;          // refined spacing dvx/REAL(gdx*sgdl)
  float dpl_sr;              // srtimes' dpl (:1705-1709)
  float dpl_ray;             // rpaths' dpl = 0.5 * min cell edge (:1864-1869)
  float x_last, z_last;      // gox+(nnx-1)*dnx, goz+(nnz-1)*dnz (:2088,2098)
};

// One eikonal solve = one (gather, ig) pass (CalSurfG.f90:1186-1355); host-computed in REAL*4
struct SweepDesc {
  int map;                    // velocity-map slot to propagate through
  float scx, scz;             // source (colatitude, longitude) in radians
  int isx, isz;               // coarse source cell after clamping (:1209-1222)
  int vnl, vnr, vnt, vnb;     // refined box in coarse node indices (:1227-1234)
  int nrnx, nrnz;             // refined extent
  float gorx, gorz;           // refined origin (:1239-1240)
  int tsx, tsz;               // travel()'s source cell on the refined grid (clamped, :312-325)
  int rsx, rsz;               // rpaths' unclamped refined source cell (:1853-1854)
  int gather;                 // flattened gather index
  int first_row;              // 0-based row (ray) index of the gather's first receiver
  int nrc;                    // receivers of the gather
  int do_times, do_rays;      // ig==1 -> times; phase or ig==2 -> rays (:1367,1380)
  int status;                 // out: 0 ok, DSURF_ERR_*
};

struct RayDesc {              // one receiver of one sweep
  int sweep;                  // slot of the sweep inside the current batch
  int row;                    // global 0-based row (= count11 - 1)
  float rcx, rcz;
  float sin_rcx;              // SIN(rcx) with the host libm (:1712, :1916)
};

// per-sweep state of a batch slot (device pointers)
struct BatchView {
  int2 *node;        // legacy pipeline: [slot][nnx*nnz] packed (float bits of ttn, nsts), iz fastest
  unsigned *word;    // [slot][wslot] one word per coarse node (alive = fp32 time), node (ix, iz) 0-based at
                     // (ix + wpx) * wld + iz + wpz; null in the packed-record mode.  Lane-per-sweep march: 3-node frame
                     // (eik_lps.cuh); fast-iterative march: tile-aligned rows (eik_fim.cuh, Layout)
  int wld, wpx, wpz;
  size_t wslot;
  int2 *box;         // [slot][289] injection scratch (status, time) of the refined box
  int2 *seed;        // [slot][289] (key bits, node) of the coarse close nodes in travel(urg=2)'s insertion order
  int *nseed;        // [slot]
  int slab;          // int2 entries of heap slab per slot
  int2 *noder;       // [slot][129*129] refined (ttnr, nstsr), leading dimension nrnz
  float *velr;       // [slot][129*129] refined velocity
  int2 *hent;        // [slot][slab] (key,node) heap entries beyond the shared-memory part
  const float *ristr; // [slot][129] earth*sin(gorx+(ix-1)*drnx), host libm
  int hcap;
  // fast-iterative march (eik_fim.cuh)
  unsigned *fim_bitmap;     // [slot][tiles * 32] node-level dirty bits, one word per tile row
  int2 *fim_reg;            // [slot][kRegNodes] start-up region after hand-over: (time bits, status 0 alive / 1 seed / -1 far)
  int4 *fim_rect;           // [slot] region rectangle (rx0, rz0, rw, rh)
  unsigned *fim_rw;         // [slot][kRegNodes] start-up scratch: region words
  unsigned char *fim_flag;  // [slot][kRegNodes] start-up scratch: seed / raised-key flags
};

// eikonal pipelines (dsurf_set_eikonal_mode / DSURF_EIKONAL): 0 exact, 16 lanes per sweep, packed records (default);
// 1 exact, lane per sweep, one word per node; 2 block-level fast-iterative sweep after an exact start-up
enum { kEikExact16 = 0, kEikLps = 1, kEikFim = 2 };
int eikonal_mode();   // mode newly created plans take

int launch_dice(cudaStream_t st, const Geom &g, const double *d_pv_map, float *d_velv, float *d_veln);
int launch_eikonal(cudaStream_t st, const Geom &g, const SweepDesc *d_sw, int nsw, const float *d_veln_all,
                   const float *d_velv_all, const float *d_risti, BatchView bv, int *launches, int mode);
int eikonal_resident_sweeps(int mode);        // packed-record pipeline only; 0 = limited by memory only
int eikonal_slab_entries(int hcap, int mode);
bool eikonal_uses_words(int mode);
void eikonal_word_layout(const Geom &g, int mode, int *wld, int *wpx, int *wpz, size_t *wslot);
size_t eikonal_fim_bitmap_words(const Geom &g);
int launch_rays(cudaStream_t st, const Geom &g, const SweepDesc *d_sw, const RayDesc *d_rays, int nrays,
                const float *d_veln_all, BatchView bv, float *d_tt, float *d_fdm, int4 *d_bbox,
                int *d_rbint, int *d_err, float2 *d_path = nullptr, int *d_path_n = nullptr, int path_cap = 0);

}  // namespace dsurf
