/*
 * dsurftomo_b200.h -- C ABI of libdsurf_b200.so, the B200 (sm_100a) implementation of
 * DSurfTomo's forward/sensitivity + LSMR hot path.
 *
 * Two layers:
 *
 *  (1) gfortran-convention drop-ins.  Same symbol names, argument order and argument meaning
 *      as the reference subroutines, every argument by reference, arrays column-major and
 *      1-based in content, no hidden arguments (there are no CHARACTER dummies).  A
 *      DSurfTomo build links main.o against this library instead of CalSurfG.o / surfdisp96.o /
 *      lsmrModule.o / aprod.o and nothing else changes.  On the reference's
 *      fatal conditions these print the reference's message and exit(1) (Fortran STOP).
 *
 *        calsurfg_               replaces  subroutine CalSurfG        src/CalSurfG.f90:939-943
 *        depthkernel_            replaces  subroutine depthkernel     src/CalSurfG.f90:1-2
 *        caldespersion_          replaces  subroutine caldespersion   src/CalSurfG.f90:2866-2867
 *        surfdisp96_             replaces  subroutine surfdisp96      src/surfdisp96.f:52-53
 *        __lsmrmodule_MOD_lsmr   replaces  LSMRmodule::LSMR           src/lsmrModule.f90:36-38
 *        aprod_                  replaces  subroutine aprod           src/aprod.f90:7
 *
 *  (2) neutral C entry points (dsurf_*): scalars by value, an int status instead of STOP.
 *      All pointers are HOST pointers unless the name ends in _dev; the callee keeps no
 *      caller pointer after return.  Status codes below.
 *
 * Nothing in this header mentions torch; device memory, streams and NCCL are internal.
 */
#ifndef DSURFTOMO_B200_H
#define DSURFTOMO_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  DSURF_OK = 0,
  DSURF_ERR_SOURCE_OUTSIDE = 1,   /* CalSurfG.f90:317-323,1214-1220: "Source lies outside bounds of model" */
  DSURF_ERR_RECEIVER_OUTSIDE = 2, /* CalSurfG.f90:1686-1692,1898-1904 */
  DSURF_ERR_NO_CUDA = 3,          /* no usable sm_100 device: the library has NO CPU fallback */
  DSURF_ERR_CUDA = 4,             /* a CUDA call failed; dsurf_last_error() has the text */
  DSURF_ERR_CAPACITY = 5,         /* caller-provided rw/iw/col capacity (maxnar) too small */
  DSURF_ERR_BAD_ARG = 6,
  DSURF_ERR_HEAP = 7,             /* narrow-band heap exceeded the reference's maxbt = 0.5*nnx*nnz */
  DSURF_ERR_NCCL = 8
};

const char *dsurf_last_error(void);
/* device selection (default: CUDA device 0, or LOCAL_RANK if set) */
int dsurf_set_device(int device);
/* library build info: "sm_100a fmad=off ..." */
const char *dsurf_build_info(void);
/* Eikonal pipeline used by plans (and drop-in calls) created AFTER the call; initial value from the environment
 * variable DSURF_EIKONAL = exact | lps | fim.
 *   0 exact  the reference's fast marching replayed in its exact heap pop order (travel/fouds2/addtree/downtree/updtree,
 *            CalSurfG.f90:288-921): travel times bit-identical to the reference.  Default.
 *   1 lps    the same exact march, one lane per sweep (slower on B200, kept for A/B runs).
 *   2 fim    block-level fast-iterative sweep of north_star (no heap; 32 x 32 tiles relaxed in shared memory) after an
 *            exact start-up around the source.  Iterates the reference's own causal update rule with the reference's
 *            fp32 arithmetic, so most nodes come out bit-identical; where the reference's heap leaves time order (equal
 *            keys, raised keys) times differ in the last bits (measured: <= 4e-6 relative on smooth models,
 *            profiles/r02_fim_parity.md).  Several times faster. */
int dsurf_set_eikonal_mode(int mode);
int dsurf_get_eikonal_mode(void);

/* ------------------------------------------------------------------ (1) Fortran drop-ins */

void calsurfg_(const int *nx, const int *ny, const int *nz, const int *nparpi, const float *vels,
               int *iw, float *rw, int *col, float *dsurf, const float *goxdf, const float *gozdf,
               const float *dvxdf, const float *dvzdf, const int *kmaxRc, const int *kmaxRg,
               const int *kmaxLc, const int *kmaxLg, const double *tRc, const double *tRg,
               const double *tLc, const double *tLg, const int *wavetype, const int *igrt,
               const int *periods, const float *depz, const float *minthk, const float *scxf,
               const float *sczf, const float *rcxf, const float *rczf, const int *nrc1,
               const int *nsrcsurf1, const int *kmax, const int *nsrcsurf, const int *nrcf, int *nar);

void depthkernel_(const int *nx, const int *ny, const int *nz, const float *vel, double *pvRc,
                  double *sen_vsRc, double *sen_vpRc, double *sen_rhoRc, const int *iwave,
                  const int *igr, const int *kmaxRc, const double *tRc, const float *depz,
                  const float *minthk);

void caldespersion_(const int *nx, const int *ny, const int *nz, const float *vel, double *pvRc,
                    const int *iwave, const int *igr, const int *kmaxRc, const double *tRc,
                    const float *depz, const float *minthk);

void surfdisp96_(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                 const int *nlayer, const int *iflsph, const int *iwave, const int *mode,
                 const int *igr, const int *kmax, const double *t, double *cg);

void __lsmrmodule_MOD_lsmr(const int *m, const int *n, const int *leniw, const int *lenrw,
                           const int *iw, const float *rw, const float *b, const float *damp,
                           const float *atol, const float *btol, const float *conlim,
                           const int *itnlim, const int *localSize, const int *nout, float *x,
                           int *istop, int *itn, float *normA, float *condA, float *normr,
                           float *normAr, float *normx);

void aprod_(const int *mode, const int *m, const int *n, float *x, float *y, const int *leniw,
            const int *lenrw, const int *iw, const float *rw);

/* replaces subroutine synthetic, src/CalSurfG.f90:2412-2415 (checkerboard forward run, ifsyn = 1,
 * main.f90:338-342); writes velmap2d{Rc,Rg,Lc,Lg}.dat into the working directory like the reference */
void synthetic_(const int *nx, const int *ny, const int *nz, const int *nparpi, const float *vels,
                float *obst, const float *goxdf, const float *gozdf, const float *dvxdf,
                const float *dvzdf, const int *kmaxRc, const int *kmaxRg, const int *kmaxLc,
                const int *kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                const double *tLg, const int *wavetype, const int *igrt, const int *periods,
                const float *depz, const float *minthk, const float *scxf, const float *sczf,
                const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                const int *kmax, const int *nsrcsurf, const int *nrcf, const float *noiselevel);

/* ------------------------------------------------------------------ (2) neutral C entry points */

/* CalSurfG with an explicit COO capacity (maxnar = spfra*dall*nx*ny*nz in main.f90:287) and a
 * status instead of STOP.  rbint (may be NULL) receives the ray-boundary warning flag of
 * CalSurfG.f90:1447-1454. */
int dsurf_calsurfg(int nx, int ny, int nz, int nparpi, const float *vels, int *iw, float *rw,
                   int *col, float *dsurf, float goxdf, float gozdf, float dvxdf, float dvzdf,
                   int kmaxRc, int kmaxRg, int kmaxLc, int kmaxLg, const double *tRc,
                   const double *tRg, const double *tLc, const double *tLg, const int *wavetype,
                   const int *igrt, const int *periods, const float *depz, float minthk,
                   const float *scxf, const float *sczf, const float *rcxf, const float *rczf,
                   const int *nrc1, const int *nsrcsurf1, int kmax, int nsrcsurf, int nrcf,
                   int64_t maxnar, int *nar, int *rbint);

/* subroutine synthetic with a status instead of STOP.  Forward-only: caldespersion maps of every
 * data type (group maps for the group types), ONE eikonal solve per gather on the gdx = gdz = 5
 * propagation grid (CalSurfG.f90:2497-2498), receiver times only.  obst(i) = t + t*gaussian()*noiselevel
 * (gaussian.f90; the normal deviates come from a xoshiro256** stream seeded with `seed`, because
 * gfortran's random_number stream cannot be reproduced outside libgfortran -- noiselevel = 0 is
 * deterministic and parity-tested).  outdir: where velmap2dXX.dat are written (format 5f8.4,
 * :2557-2613), NULL = do not write. */
int dsurf_synthetic(int nx, int ny, int nz, int nparpi, const float *vels, float *obst, float goxdf,
                    float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                    int kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                    const double *tLg, const int *wavetype, const int *igrt, const int *periods,
                    const float *depz, float minthk, const float *scxf, const float *sczf,
                    const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                    int kmax, int nsrcsurf, int nrcf, float noiselevel, const char *outdir,
                    uint64_t seed, int *rbint);

/* depthkernel / caldespersion (igr: 0 phase, 1 group; iwave: 1 Love, 2 Rayleigh).
 * sen_* may all be NULL (dispersion map only == caldespersion). */
int dsurf_depthkernel(int nx, int ny, int nz, const float *vel, double *pv, double *sen_vs,
                      double *sen_vp, double *sen_rho, int iwave, int igr, int kmax,
                      const double *t, const float *depz, float minthk);

/* One layered model (the reference's per-column call).  Batched form: nmodel stacks of
 * nlayer layers each, thk shared. */
int dsurf_surfdisp96(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                     int nlayer, int iflsph, int iwave, int mode, int igr, int kmax,
                     const double *t, double *cg);
int dsurf_surfdisp96_batch(int nmodel, const float *thkm, const float *vpm, const float *vsm,
                           const float *rhom, int nlayer, int iflsph, int iwave, int igr, int kmax,
                           const double *t, double *cg /* [nmodel][kmax] */);

int dsurf_lsmr(int m, int n, int leniw, int lenrw, const int *iw, const float *rw, const float *b,
               float damp, float atol, float btol, float conlim, int itnlim, int localSize,
               float *x, int *istop, int *itn, float *normA, float *condA, float *normr,
               float *normAr, float *normx);

int dsurf_aprod(int mode, int m, int n, float *x, float *y, int leniw, int lenrw, const int *iw,
                const float *rw);

/* ------------------------------------------------------------------ staged / device-resident API
 * Used by bench.py and the stage-level parity tests: the same kernels the drop-ins run, with
 * inputs kept resident in HBM and per-stage CUDA-event timings. */

typedef struct dsurf_plan dsurf_plan;

/* Uploads model, geometry and the gather tables (same arrays as dsurf_calsurfg). */
int dsurf_plan_create(dsurf_plan **plan, int nx, int ny, int nz, const float *vels, float goxdf,
                      float gozdf, float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc,
                      int kmaxLg, const double *tRc, const double *tRg, const double *tLc,
                      const double *tLg, const int *wavetype, const int *igrt, const int *periods,
                      const float *depz, float minthk, const float *scxf, const float *sczf,
                      const float *rcxf, const float *rczf, const int *nrc1, const int *nsrcsurf1,
                      int kmax, int nsrcsurf, int nrcf);
/* forward-only plan (what dsurf_synthetic runs): gd = 5, maps only, one sweep per gather, times only */
int dsurf_plan_create_forward(dsurf_plan **plan, int nx, int ny, int nz, const float *vels,
                              float goxdf, float gozdf, float dvxdf, float dvzdf, int kmaxRc,
                              int kmaxRg, int kmaxLc, int kmaxLg, const double *tRc,
                              const double *tRg, const double *tLc, const double *tLg,
                              const int *wavetype, const int *igrt, const int *periods,
                              const float *depz, float minthk, const float *scxf, const float *sczf,
                              const float *rcxf, const float *rczf, const int *nrc1,
                              const int *nsrcsurf1, int kmax, int nsrcsurf, int nrcf);
int dsurf_plan_destroy(dsurf_plan *plan);
/* replace the model (vels[nz][ny][nx]) for the next outer iteration */
int dsurf_plan_set_model(dsurf_plan *plan, const float *vels);
/* K1: dispersion maps + depth kernels for every data type (CalSurfG.f90:1098-1133) */
int dsurf_plan_dispersion(dsurf_plan *plan);
/* caller-provided dispersion results of one data type instead of K1 (layouts of dsurf_depthkernel;
 * pv has kmax columns for Rc/Lc, kmaxRg/kmaxLg for Rg/Lg), then finalize (combine + dice) */
int dsurf_plan_set_dispersion(dsurf_plan *plan, int type, const double *pv, const double *sen_vs,
                              const double *sen_vp, const double *sen_rho);
int dsurf_plan_finalize_dispersion(dsurf_plan *plan);
/* test hook: overwrite the velocity map (nx*ny doubles) of period-type slot `map` */
int dsurf_plan_set_map(dsurf_plan *plan, int type /*0 Rc,1 Rg,2 Lc,3 Lg*/, int period0, const double *pv);
/* K2-K6 for gathers [g0, g1) of the flattened (knumi, srcnum) loop nest; results are appended
 * to the plan's device COO / dsurf buffers in reference order.  dsurf_plan_reset_rows() rewinds. */
int dsurf_plan_reset_rows(dsurf_plan *plan);
int dsurf_plan_sweeps(dsurf_plan *plan, int g0, int g1);
int dsurf_plan_num_gathers(const dsurf_plan *plan);
int dsurf_plan_num_sweeps(const dsurf_plan *plan, int g0, int g1);
int64_t dsurf_plan_nar(const dsurf_plan *plan);
int dsurf_plan_nrows(const dsurf_plan *plan);
/* download everything produced so far (any pointer may be NULL) */
int dsurf_plan_download(dsurf_plan *plan, int *iw_rows /* nar */, float *rw, int *col, float *dsurf,
                        int *rbint);
/* Ray-path export (optional debug output of the ray tracer; the reference's raypath.out block,
 * CalSurfG.f90:2276-2283, consumer scripts/plotpath.py): every later dsurf_plan_sweeps call appends,
 * per traced ray in (gather, receiver) order, a list-directed "# nrp" record and nrp records
 * "latitude longitude" (degrees) from the receiver to the source.  max_points <= 0 selects
 * 4*(nnx+nnz) points per ray; a longer ray makes dsurf_plan_sweeps return DSURF_ERR_CAPACITY.
 * file == NULL stops the export and closes the file. */
int dsurf_plan_set_raypath(dsurf_plan *plan, const char *file, int max_points);
int64_t dsurf_plan_raypath_count(const dsurf_plan *plan);
/* stage-level outputs for parity tests: last solved sweep of gather g, pass ig (1 or 2) */
int dsurf_plan_debug_sweep(dsurf_plan *plan, int g, int ig, float *veln, float *ttn, float *ttnr,
                           int *nstsr, float *rgeom, float *fdm /* [nrc][nx][ny] or NULL */);
/* copy device dispersion results to the host (any pointer may be NULL) */
int dsurf_plan_get_dispersion(dsurf_plan *plan, int type, double *pv, double *sen_vs, double *sen_vp,
                              double *sen_rho);
/* timings (ms, CUDA events on the plan's stream) of the last call of each stage:
 * [0] dispersion, [1] dice, [2] eikonal (refined+coarse FMM), [3] receivers+rays,
 * [4] row assembly, [5] eikonal launches, [6] total kernel launches of last sweeps call,
 * [7] sweeps solved in the last call */
int dsurf_plan_timings(const dsurf_plan *plan, double *ms8);
/* device time (ms, CUDA events on the launching stream) of the whole last dsurf_plan_sweeps call */
double dsurf_plan_last_sweeps_ms(const dsurf_plan *plan);

/* Multi-GPU exchange of the path (one process per GPU; comm = ncclComm_t from dsurf_nccl_comm_init).  The units of
 * the loop nest CalSurfG.f90:1144-1145 are sharded over ranks in contiguous gather blocks, so the reference's
 * output is the rank-order concatenation of the ranks' outputs.  One call gathers the predicted times of all rows
 * (in place in the plan's full-length dsurf) and, with want_coo != 0, the COO row blocks of every rank into a
 * second device buffer -- the full (rw, iw, col) that main.f90:355-359 receives.  ncclAllGather of the counts, then
 * grouped ncclBroadcast of each block straight to its offset. */
int dsurf_plan_allgather(dsurf_plan *plan, void *nccl_comm, int rank, int nranks, int want_coo, int64_t *nar_total);
double dsurf_plan_last_gather_ms(const dsurf_plan *plan);
int dsurf_plan_download_gathered(dsurf_plan *plan, int *iw_rows, float *rw, int *col);
/* order-sensitive 64-bit digest of the plan's own COO (gathered == 0) or of the gathered one: equal digests on
 * 1 and on N GPUs <=> identical triplets in identical order */
int dsurf_plan_digest(dsurf_plan *plan, int gathered, uint64_t *digest, int64_t *n);

/* Device-resident LSMR: build from host COO once, then run iterations with everything in HBM. */
typedef struct dsurf_lsmr_sys dsurf_lsmr_sys;
int dsurf_lsmr_create(dsurf_lsmr_sys **sys, int m, int n, int64_t nar, const int *rows1,
                      const int *cols1, const float *vals, const float *b);
/* same, taking the COO a plan holds in HBM: the host glue of main.f90:361-466 (residual
 * cbst = obst - dsyn, getpercentile outlier weights, rw *= datweight(row), DWS statistics,
 * smoothing rows appended, right-hand side) runs on the device, so the matrix never leaves HBM
 * between CalSurfG and LSMR.  obst: dall observed times (host).  The plan's rw is scaled in
 * place, as the reference's main program does. */
int dsurf_lsmr_create_from_plan(dsurf_lsmr_sys **sys, dsurf_plan *plan, const float *obst,
                                float threshold0, float weight);
/* row-partitioned variant for the distributed LSMR: this rank's system holds the data rows its plan produced
 * (renumbered from 1) and a contiguous share of the smoothing rows; every rank must hold the predicted times of
 * all rows (dsurf_plan_allgather).  nranks == 1 is dsurf_lsmr_create_from_plan. */
int dsurf_lsmr_create_from_plan_shard(dsurf_lsmr_sys **sys, dsurf_plan *plan, const float *obst, float threshold0,
                                      float weight, int rank, int nranks, int *m_local, int64_t *nnz_local);
/* results of the last dsurf_lsmr_create_from_plan (any pointer may be NULL): cbst(1:dall) after
 * outlier rejection, datweight(1:dall), stats4 = {q25, q75, maxnorm, averdws} (main.f90:364,386-394),
 * m = dall + count3, nar including the smoothing rows */
int dsurf_plan_glue_results(dsurf_plan *plan, float *cbst, float *datweight, float *stats4, int *m,
                            int64_t *nar);
/* main.f90:518-532 on the device: dv = solution of the last dsurf_lsmr_solve of `sys`, clipped to
 * +-0.5, added to the interior nodes of the plan's model, model clamped to [minvel, maxvel].
 * dv_host (maxvp, clipped) and vels_host (nx*ny*nz, updated model) may be NULL. */
int dsurf_plan_update_model(dsurf_plan *plan, dsurf_lsmr_sys *sys, float minvel, float maxvel,
                            float *dv_host, float *vels_host);
/* column-order hint (P = (nx-2)(ny-2) vertices, K = nz-1 depths) enabling the depth-blocked
 * sparse layout when n == P*K; set automatically by every dsurf_plan_create / CalSurfG call */
int dsurf_lsmr_hint_geometry(int nx, int ny, int nz);
int dsurf_lsmr_destroy(dsurf_lsmr_sys *sys);
/* multi-GPU: rows are partitioned over ranks; comm is an ncclComm_t created by the host side
 * (see dsurftomo_b200/dist.py); the per-iteration exchange is one all-reduce of n+1 floats. */
int dsurf_lsmr_set_comm(dsurf_lsmr_sys *sys, void *nccl_comm, int rank, int nranks);
/* peer-memory exchange of the distributed LSMR (replaces the per-iteration NCCL all-reduce by direct NVLink loads
 * from every rank's exchange buffer; the iteration becomes a CUDA graph): every rank exports the CUDA IPC handle of
 * its buffer (64 bytes), the host side all-gathers the handles (dsurftomo_b200/dist.py) and attaches them.
 * dsurf_lsmr_set_comm is still needed (the two set-up reductions of a solve use NCCL).  Destroy the systems only after
 * every rank has returned from its last solve (a barrier on the host side): peers read each other's buffers. */
int dsurf_lsmr_xchg_export(dsurf_lsmr_sys *sys, void *handle64);
int dsurf_lsmr_xchg_attach(dsurf_lsmr_sys *sys, const void *handles /* nranks x 64 bytes */, int rank, int nranks);
int dsurf_nccl_unique_id(void *id128);
int dsurf_nccl_comm_init(void **comm, const void *id128, int rank, int nranks);
int dsurf_nccl_comm_destroy(void *comm);
/* run; max_iters<=0 -> itnlim.  force_iters != 0 disables the stopping tests (benchmarks). */
int dsurf_lsmr_solve(dsurf_lsmr_sys *sys, float damp, float atol, float btol, float conlim,
                     int itnlim, int localSize, int force_iters, float *x_host, int *istop, int *itn,
                     float *normA, float *condA, float *normr, float *normAr, float *normx,
                     double *ms_total, double *ms_spmv, double *ms_spmtv);
int64_t dsurf_lsmr_nnz(const dsurf_lsmr_sys *sys);
/* CTAs of the thread-block cluster running the fused small-vector phases of the last solve
 * (16 or 8), 0 = unfused kernels, -1 = no solve yet */
int dsurf_lsmr_fused_cluster(const dsurf_lsmr_sys *sys);

#ifdef __cplusplus
}
#endif
#endif
