/*
 * oracle.h -- C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This directory is a plain C++ restatement of the
 * reference's (HongjianFang/DSurfTomo) forward/sensitivity + LSMR hot path.  It is
 * the checker for the CUDA product in dsurftomo_b200/: only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * Nothing under dsurftomo_b200/ links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs
 *  and cannot be compiled here (no Fortran compiler in the image),
 * so this restatement is pinned only by the analytic known-answer tests in
 * tests/test_oracle_*.py and by statement-by-statement review against the cited lines.
 *
 * Numeric contract: REAL / REAL(KIND=i10) / real(dp) -> float,
 * real*8 / double precision -> double, INT() -> truncation, strict IEEE evaluation
 * (build with -ffp-contract=off -fno-fast-math), serial semantics for the SAVEd
 * del1st/dhalf of surfdisp96.f:409,509.
 */
#ifndef DSURF_ORACLE_H
#define DSURF_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* surfdisp96.f:52  (status: number of periods for which no root was found) */
int oracle_surfdisp96(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                      int nlayer, int iflsph, int iwave, int mode, int igr, int kmax,
                      const double *t, double *cg);

/* CalSurfG.f90:2352 */
void oracle_refine_grid2layer(float minthk0, int mmax, const float *dep, const float *vp,
                              const float *vs, const float *rho, int *rmax, float *rdep,
                              float *rvp, float *rvs, float *rrho, float *rthk);

/* CalSurfG.f90:1-169; arrays column-major as in the reference:
 * vel[nz][ny][nx], pv[kmax][nx*ny], sen_*[nz][kmax][nx*ny]; nthreads<=0 -> serial. */
void oracle_depthkernel(int nx, int ny, int nz, const float *vel, double *pv, double *sen_vs,
                        double *sen_vp, double *sen_rho, int iwave, int igr, int kmax,
                        const double *t, const float *depz, float minthk, int nthreads);
/* CalSurfG.f90:2866-2927 */
void oracle_caldespersion(int nx, int ny, int nz, const float *vel, double *pv, int iwave, int igr,
                          int kmax, const double *t, const float *depz, float minthk, int nthreads);

/* One (gather, ig) eikonal solve: gridder + refined-source FMM + injection + coarse FMM
 * (CalSurfG.f90:1186-1355).  pv = velocity on the nx*ny model columns (double, as velf).
 * Outputs (any may be NULL): veln[nnx][nnz], ttn[nnx][nnz] (iz fastest), refined
 * ttnr/nstsr [nnxr][nnzr], refined geometry in rgeom[6] = goxr,gozr,dnxr,dnzr,(float)nnxr,(float)nnzr.
 * returns 0, or 1 if the source lies outside the grid. */
int oracle_fmm_sweep(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                     const double *pv, float scx, float scz, float *veln, float *ttn,
                     float *ttnr, int *nstsr, float *rgeom);

/* Same sweep followed by srtimes + rpaths for nrc receivers (CalSurfG.f90:1366-1382):
 * tt[nrc], fdm[nrc][(nvx+2)][(nvz+2)] (column-major fdm(0:nvz+1,0:nvx+1) per ray). */
int oracle_sweep_rays(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                      const double *pv, float scx, float scz, int nrc, const float *rcx,
                      const float *rcz, float *tt, float *fdm);

/* Ray geometry rgx/rgz(1:nrp) (colatitude, longitude in radians, receiver first, source last) of
 * every receiver of one sweep -- the content of the reference's raypath.out block
 * (CalSurfG.f90:2276-2283).  npts[nrc]; px, pz are [nrc][cap]. */
int oracle_sweep_paths(int nx, int ny, float goxd, float gozd, float dvxd, float dvzd,
                       const double *pv, float scx, float scz, int nrc, const float *rcx,
                       const float *rcz, int cap, int *npts, float *px, float *pz);

/* CalSurfG.f90:939-1459; same argument list as the Fortran subroutine (by value where
 * scalar).  nthreads: OpenMP threads; mode 0 = reference-faithful threading (dispersion
 * only), 1 = additionally thread the independent gathers.  Returns 0 or an error code
 * (1 source outside, 2 receiver outside).  rbint_out (may be NULL) receives the
 * ray-boundary warning flag. */
int oracle_calsurfg(int nx, int ny, int nz, int nparpi, const float *vels, int *iw, float *rw,
                    int *col, float *dsurf, float goxdf, float gozdf, float dvxdf, float dvzdf,
                    int kmaxRc, int kmaxRg, int kmaxLc, int kmaxLg, const double *tRc,
                    const double *tRg, const double *tLc, const double *tLg, const int *wavetype,
                    const int *igrt, const int *periods, const float *depz, float minthk,
                    const float *scxf, const float *sczf, const float *rcxf, const float *rczf,
                    const int *nrc1, const int *nsrcsurf1, int kmax, int nsrcsurf, int nrcf,
                    int *nar, int nthreads, int mode, int *rbint_out, double *stage_seconds);

/* subroutine synthetic, CalSurfG.f90:2412-2865 (noise-free part): forward times on the gd = 5
 * propagation grid through caldespersion maps; pv_out (may be NULL) receives the four maps. */
int oracle_synthetic(int nx, int ny, int nz, const float *vels, float *obst, float goxdf, float gozdf,
                     float dvxdf, float dvzdf, int kmaxRc, int kmaxRg, int kmaxLc, int kmaxLg,
                     const double *tRc, const double *tRg, const double *tLc, const double *tLg,
                     const int *wavetype, const int *igrt, const int *periods, const float *depz,
                     float minthk, const float *scxf, const float *sczf, const float *rcxf,
                     const float *rczf, const int *nrc1, const int *nsrcsurf1, int kmax, int nsrcsurf,
                     int nrcf, int nthreads, double *pv_out);

/* aprod.f90:7 */
void oracle_aprod(int mode, int m, int n, float *x, float *y, int leniw, int lenrw, const int *iw,
                  const float *rw);
/* lsmrModule.f90:36 */
void oracle_lsmr(int m, int n, int leniw, int lenrw, const int *iw, const float *rw, const float *b,
                 float damp, float atol, float btol, float conlim, int itnlim, int localSize,
                 float *x, int *istop, int *itn, float *normA, float *condA, float *normr,
                 float *normAr, float *normx);
/* lsmrblas.f90:247 */
float oracle_snrm2(int n, const float *x);

/* delsph.f90, getpercentile.f90 */
float oracle_delsph(float flat1, float flon1, float flat2, float flon2);
void oracle_getpercentile(int n, const float *array, float *q25, float *q75);

/* main.f90:361-466: residuals, outlier weights, smoothing rows, iw packing.
 * In/out exactly as the main program's arrays; returns m (=dall+count3); nar updated. */
int oracle_host_glue(int nx, int ny, int nz, int dall, const float *obst, const float *dsyn,
                     float threshold0, float weight, int *iw, float *rw, int *col, float *cbst,
                     float *datweight, int *nar);
/* main.f90:518-532 */
void oracle_model_update(int nx, int ny, int nz, float *vsf, float *dv, float minvel, float maxvel);

#ifdef __cplusplus
}
#endif
#endif
