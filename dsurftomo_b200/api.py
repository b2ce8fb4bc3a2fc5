"""Python host-side mirror of the reference's interface for the hot path.

Same names, argument meaning and error behaviour as the Fortran subroutines (a Python host
cannot be the Fortran caller itself -- no Fortran toolchain exists in this image -- so this module
is what the parity tests and bench.py drive; the Fortran drop-in symbols live in the same
shared library, see include/dsurftomo_b200.h and INTEGRATION.md):

    CalSurfG(...)        src/CalSurfG.f90:939     -> dsurf_calsurfg
    depthkernel(...)     src/CalSurfG.f90:1       -> dsurf_depthkernel
    caldespersion(...)   src/CalSurfG.f90:2866    -> dsurf_depthkernel(sen = NULL)
    surfdisp96(...)      src/surfdisp96.f:52      -> dsurf_surfdisp96
    LSMR(...)            src/lsmrModule.f90:36    -> dsurf_lsmr
    aprod(...)           src/aprod.f90:7          -> dsurf_aprod

Arrays use the Fortran memory layout expressed as C-ordered numpy arrays with reversed axes
(see dsurftomo_b200.inputs).  Everything runs on the GPU through the C ABI; host buffers in,
host buffers out.  :class:`Plan` and :class:`LsmrSystem` expose the staged, device-resident
form of the same kernels used by bench.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DsurfError, check, lib, ptr

F32, F64, I32 = np.float32, np.float64, np.int32
_dummy64 = np.zeros(1, F64)


EIKONAL_MODES = {"exact": 0, "lps": 1, "fim": 2}


def set_eikonal_mode(mode):
    """Eikonal pipeline of plans / CalSurfG calls created after this call (dsurf_set_eikonal_mode):
    "exact" (default: the reference's heap pop order, bit-identical times), "lps" (exact, lane per sweep) or
    "fim" (block-level fast-iterative sweep, last-bit deviations where the reference's heap leaves time order).
    Returns the previous mode name."""
    prev = lib().dsurf_get_eikonal_mode()
    m = EIKONAL_MODES[mode] if isinstance(mode, str) else int(mode)
    check(lib().dsurf_set_eikonal_mode(C.c_int(m)), "set_eikonal_mode")
    return [k for k, v in EIKONAL_MODES.items() if v == prev][0]


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def _t(t):
    t = _c(t, F64)
    return t if t.size else _dummy64


def surfdisp96(thkm, vpm, vsm, rhom, nlayer, iflsph, iwave, mode, igr, kmax, t):
    """surfdisp96.f:52 -- returns cg[kmax] (float64 holding REAL*4-rounded velocities)."""
    thkm, vpm, vsm, rhom = (_c(a, F32) for a in (thkm, vpm, vsm, rhom))
    t = _c(t, F64)
    cg = np.zeros(kmax, F64)
    check(lib().dsurf_surfdisp96(ptr(thkm, C.c_float), ptr(vpm, C.c_float), ptr(vsm, C.c_float),
                                 ptr(rhom, C.c_float), C.c_int(nlayer), C.c_int(iflsph), C.c_int(iwave),
                                 C.c_int(mode), C.c_int(igr), C.c_int(kmax), ptr(t, C.c_double),
                                 ptr(cg, C.c_double)), "surfdisp96")
    return cg


def surfdisp96_batch(thkm, vpm, vsm, rhom, iflsph, iwave, igr, t):
    """Many layer stacks sharing one thickness vector: vpm/vsm/rhom are [nmodel, nlayer]."""
    thkm = _c(thkm, F32)
    vpm, vsm, rhom = (_c(a, F32) for a in (vpm, vsm, rhom))
    nmodel, nlayer = vpm.shape
    t = _c(t, F64)
    cg = np.zeros((nmodel, len(t)), F64)
    check(lib().dsurf_surfdisp96_batch(C.c_int(nmodel), ptr(thkm, C.c_float), ptr(vpm, C.c_float),
                                       ptr(vsm, C.c_float), ptr(rhom, C.c_float), C.c_int(nlayer),
                                       C.c_int(iflsph), C.c_int(iwave), C.c_int(igr), C.c_int(len(t)),
                                       ptr(t, C.c_double), ptr(cg, C.c_double)), "surfdisp96_batch")
    return cg


def depthkernel(nx, ny, nz, vel, iwave, igr, kmax, t, depz, minthk):
    """CalSurfG.f90:1 -- returns pv[kmax, nx*ny], sen_vs/sen_vp/sen_rho[nz, kmax, nx*ny]."""
    vel = _c(vel, F32)
    t = _c(t, F64)
    depz = _c(depz, F32)
    pv = np.zeros((kmax, nx * ny), F64)
    sen = [np.zeros((nz, kmax, nx * ny), F64) for _ in range(3)]
    check(lib().dsurf_depthkernel(C.c_int(nx), C.c_int(ny), C.c_int(nz), ptr(vel, C.c_float),
                                  ptr(pv, C.c_double), ptr(sen[0], C.c_double), ptr(sen[1], C.c_double),
                                  ptr(sen[2], C.c_double), C.c_int(iwave), C.c_int(igr), C.c_int(kmax),
                                  ptr(t, C.c_double), ptr(depz, C.c_float), C.c_float(minthk)), "depthkernel")
    return pv, sen[0], sen[1], sen[2]


def caldespersion(nx, ny, nz, vel, iwave, igr, kmax, t, depz, minthk):
    """CalSurfG.f90:2866 -- dispersion map only."""
    vel = _c(vel, F32)
    t = _c(t, F64)
    depz = _c(depz, F32)
    pv = np.zeros((kmax, nx * ny), F64)
    check(lib().dsurf_depthkernel(C.c_int(nx), C.c_int(ny), C.c_int(nz), ptr(vel, C.c_float),
                                  ptr(pv, C.c_double), None, None, None, C.c_int(iwave), C.c_int(igr),
                                  C.c_int(kmax), ptr(t, C.c_double), ptr(depz, C.c_float),
                                  C.c_float(minthk)), "caldespersion")
    return pv


def _plan_args(pb, vels):
    vels = _c(pb.vsf if vels is None else vels, F32)
    keep = dict(vels=vels, tRc=_t(pb.tRc), tRg=_t(pb.tRg), tLc=_t(pb.tLc), tLg=_t(pb.tLg),
                wavetype=_c(pb.wavetype, I32), igrt=_c(pb.igrt, I32), periods=_c(pb.periods, I32),
                depz=_c(pb.depz, F32), scxf=_c(pb.scxf, F32), sczf=_c(pb.sczf, F32),
                rcxf=_c(pb.rcxf, F32), rczf=_c(pb.rczf, F32), nrc1=_c(pb.nrc1, I32),
                nsrc1=_c(pb.nsrc1, I32))
    return keep


def CalSurfG(pb, vels=None, maxnar=None):
    """CalSurfG (CalSurfG.f90:939) on a :class:`inputs.Problem`: host buffers in and out through
    the C ABI.  Returns dict(dsurf, iw, rw, col, nar, rbint) with iw/rw/col sized maxnar like the
    reference's caller-owned arrays (iw[1+k] = row of triplet k)."""
    k = _plan_args(pb, vels)
    if maxnar is None:
        maxnar = max(pb.maxnar(), 1)
    iw = np.zeros(2 * maxnar + 1, I32)
    rw = np.zeros(maxnar, F32)
    col = np.zeros(maxnar, I32)
    dsurf = np.zeros(pb.dall, F32)
    nar = C.c_int(0)
    rbint = C.c_int(0)
    check(lib().dsurf_calsurfg(
        C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), C.c_int(pb.maxvp), ptr(k["vels"], C.c_float),
        ptr(iw, C.c_int), ptr(rw, C.c_float), ptr(col, C.c_int), ptr(dsurf, C.c_float),
        C.c_float(pb.goxd), C.c_float(pb.gozd), C.c_float(pb.dvxd), C.c_float(pb.dvzd),
        C.c_int(pb.kmaxRc), C.c_int(pb.kmaxRg), C.c_int(pb.kmaxLc), C.c_int(pb.kmaxLg),
        ptr(k["tRc"], C.c_double), ptr(k["tRg"], C.c_double), ptr(k["tLc"], C.c_double),
        ptr(k["tLg"], C.c_double), ptr(k["wavetype"], C.c_int), ptr(k["igrt"], C.c_int),
        ptr(k["periods"], C.c_int), ptr(k["depz"], C.c_float), C.c_float(pb.minthk),
        ptr(k["scxf"], C.c_float), ptr(k["sczf"], C.c_float), ptr(k["rcxf"], C.c_float),
        ptr(k["rczf"], C.c_float), ptr(k["nrc1"], C.c_int), ptr(k["nsrc1"], C.c_int), C.c_int(pb.kmax),
        C.c_int(pb.nsrc), C.c_int(pb.nrc), C.c_int64(maxnar), C.byref(nar), C.byref(rbint)), "CalSurfG")
    n = nar.value
    return dict(nar=n, dsurf=dsurf, iw=iw, rw_full=rw, col_full=col, rw=rw[:n], row=iw[1:n + 1], col=col[:n],
                rbint=rbint.value)


def synthetic(pb, vels=None, noiselevel=0.0, outdir=None, seed=20150131):
    """subroutine synthetic (CalSurfG.f90:2412-2865): forward times through `vels` (the "true"
    model) on the gd = 5 propagation grid; returns obst (dall,) and the boundary-ray flag."""
    k = _plan_args(pb, vels)
    obst = np.zeros(max(pb.dall, 1), F32)
    rb = C.c_int(0)
    check(lib().dsurf_synthetic(
        C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), C.c_int(pb.maxvp), ptr(k["vels"], C.c_float),
        ptr(obst, C.c_float), C.c_float(pb.goxd), C.c_float(pb.gozd), C.c_float(pb.dvxd), C.c_float(pb.dvzd),
        C.c_int(pb.kmaxRc), C.c_int(pb.kmaxRg), C.c_int(pb.kmaxLc), C.c_int(pb.kmaxLg),
        ptr(k["tRc"], C.c_double), ptr(k["tRg"], C.c_double), ptr(k["tLc"], C.c_double), ptr(k["tLg"], C.c_double),
        ptr(k["wavetype"], C.c_int), ptr(k["igrt"], C.c_int), ptr(k["periods"], C.c_int), ptr(k["depz"], C.c_float),
        C.c_float(pb.minthk), ptr(k["scxf"], C.c_float), ptr(k["sczf"], C.c_float), ptr(k["rcxf"], C.c_float),
        ptr(k["rczf"], C.c_float), ptr(k["nrc1"], C.c_int), ptr(k["nsrc1"], C.c_int), C.c_int(pb.kmax),
        C.c_int(pb.nsrc), C.c_int(pb.nrc), C.c_float(noiselevel),
        C.c_char_p(outdir.encode()) if outdir is not None else None, C.c_uint64(seed), C.byref(rb)), "synthetic")
    return dict(obst=obst[: pb.dall], rbint=rb.value)


def aprod(mode, m, n, x, y, leniw, lenrw, iw, rw):
    """aprod.f90:7 -- mode 1: y += A x; mode 2: x += A' y.  Returns (x, y)."""
    x = _c(x, F32).copy()
    y = _c(y, F32).copy()
    iw = _c(iw, I32)
    rw = _c(rw, F32)
    check(lib().dsurf_aprod(C.c_int(mode), C.c_int(m), C.c_int(n), ptr(x, C.c_float), ptr(y, C.c_float),
                            C.c_int(leniw), C.c_int(lenrw), ptr(iw, C.c_int), ptr(rw, C.c_float)), "aprod")
    return x, y


def LSMR(m, n, leniw, lenrw, iw, rw, b, damp, atol, btol, conlim, itnlim, localSize, nout=0):
    """LSMRmodule::LSMR (lsmrModule.f90:36).  Returns dict(x, istop, itn, normA, condA, normr,
    normAr, normx).  nout is ignored like in the reference's caller (undefined there)."""
    iw = _c(iw, I32)
    rw = _c(rw, F32)
    b = _c(b, F32)
    x = np.zeros(n, F32)
    istop, itn = C.c_int(0), C.c_int(0)
    vals = [C.c_float(0) for _ in range(5)]
    check(lib().dsurf_lsmr(C.c_int(m), C.c_int(n), C.c_int(leniw), C.c_int(lenrw), ptr(iw, C.c_int),
                           ptr(rw, C.c_float), ptr(b, C.c_float), C.c_float(damp), C.c_float(atol),
                           C.c_float(btol), C.c_float(conlim), C.c_int(itnlim), C.c_int(localSize),
                           ptr(x, C.c_float), C.byref(istop), C.byref(itn), *[C.byref(v) for v in vals]), "LSMR")
    return dict(x=x, istop=istop.value, itn=itn.value, normA=vals[0].value, condA=vals[1].value,
                normr=vals[2].value, normAr=vals[3].value, normx=vals[4].value)


class Plan:
    """Device-resident forward/sensitivity plan (the kernels CalSurfG runs, staged)."""

    def __init__(self, pb, vels=None, forward=False):
        """forward=True: the forward-only plan of subroutine synthetic (gd = 5, times only)."""
        self.pb = pb
        self.forward = forward
        k = self._keep = _plan_args(pb, vels)
        h = C.c_void_p()
        create = lib().dsurf_plan_create_forward if forward else lib().dsurf_plan_create
        check(create(
            C.byref(h), C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), ptr(k["vels"], C.c_float),
            C.c_float(pb.goxd), C.c_float(pb.gozd), C.c_float(pb.dvxd), C.c_float(pb.dvzd),
            C.c_int(pb.kmaxRc), C.c_int(pb.kmaxRg), C.c_int(pb.kmaxLc), C.c_int(pb.kmaxLg),
            ptr(k["tRc"], C.c_double), ptr(k["tRg"], C.c_double), ptr(k["tLc"], C.c_double),
            ptr(k["tLg"], C.c_double), ptr(k["wavetype"], C.c_int), ptr(k["igrt"], C.c_int),
            ptr(k["periods"], C.c_int), ptr(k["depz"], C.c_float), C.c_float(pb.minthk),
            ptr(k["scxf"], C.c_float), ptr(k["sczf"], C.c_float), ptr(k["rcxf"], C.c_float),
            ptr(k["rczf"], C.c_float), ptr(k["nrc1"], C.c_int), ptr(k["nsrc1"], C.c_int),
            C.c_int(pb.kmax), C.c_int(pb.nsrc), C.c_int(pb.nrc)), "plan_create")
        self.h = h

    def close(self):
        if self.h:
            lib().dsurf_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_gathers(self):
        return lib().dsurf_plan_num_gathers(self.h)

    def num_sweeps(self, g0=0, g1=None):
        return lib().dsurf_plan_num_sweeps(self.h, g0, self.num_gathers if g1 is None else g1)

    def set_model(self, vels):
        vels = _c(vels, F32)
        check(lib().dsurf_plan_set_model(self.h, ptr(vels, C.c_float)), "plan_set_model")

    def dispersion(self):
        check(lib().dsurf_plan_dispersion(self.h), "plan_dispersion")

    def set_dispersion(self, type_, pv, sen_vs, sen_vp, sen_rho):
        """Caller-provided dispersion results of one data type (0 Rc, 1 Rg, 2 Lc, 3 Lg)."""
        arrs = [_c(a, F64) if a is not None else None for a in (pv, sen_vs, sen_vp, sen_rho)]
        check(lib().dsurf_plan_set_dispersion(self.h, C.c_int(type_), *[ptr(a, C.c_double) for a in arrs]),
              "plan_set_dispersion")

    def finalize_dispersion(self):
        check(lib().dsurf_plan_finalize_dispersion(self.h), "plan_finalize_dispersion")

    def set_map(self, type_, period0, pv):
        pv = _c(pv, F64)
        check(lib().dsurf_plan_set_map(self.h, C.c_int(type_), C.c_int(period0), ptr(pv, C.c_double)), "plan_set_map")

    def set_raypath(self, file, max_points=0):
        """Append the traced ray geometry of every later sweeps() call to `file` in the format of the
        reference's raypath.out (CalSurfG.f90:2276-2283); file=None stops and closes."""
        check(lib().dsurf_plan_set_raypath(self.h, file.encode() if file is not None else None,
                                           C.c_int(max_points)), "plan_set_raypath")

    def reset_rows(self):
        check(lib().dsurf_plan_reset_rows(self.h), "plan_reset_rows")

    def sweeps(self, g0=0, g1=None):
        check(lib().dsurf_plan_sweeps(self.h, g0, self.num_gathers if g1 is None else g1), "plan_sweeps")

    @property
    def nar(self):
        return int(lib().dsurf_plan_nar(self.h))

    def download(self, out=None):
        """Device -> host copy of everything produced so far.  `out` may hold preallocated
        (e.g. pinned) numpy buffers 'row', 'rw', 'col', 'dsurf' of sufficient size."""
        n = self.nar
        out = out or {}
        rows = out.get("row") if out.get("row") is not None else np.zeros(max(n, 1), I32)
        rw = out.get("rw") if out.get("rw") is not None else np.zeros(max(n, 1), F32)
        col = out.get("col") if out.get("col") is not None else np.zeros(max(n, 1), I32)
        dsurf = out.get("dsurf") if out.get("dsurf") is not None else np.zeros(max(self.pb.dall, 1), F32)
        assert len(rows) >= n and len(rw) >= n and len(col) >= n and len(dsurf) >= self.pb.dall
        rb = C.c_int(0)
        check(lib().dsurf_plan_download(self.h, ptr(rows, C.c_int), ptr(rw, C.c_float), ptr(col, C.c_int),
                                        ptr(dsurf, C.c_float), C.byref(rb)), "plan_download")
        return dict(nar=n, row=rows[:n], rw=rw[:n], col=col[:n], dsurf=dsurf[: self.pb.dall], rbint=rb.value)

    def allgather(self, comm, want_coo=True):
        """Multi-GPU exchange (dsurf_plan_allgather): predicted times of all rows, and the COO row blocks of every
        rank in rank order when want_coo.  Returns the total number of triplets."""
        tot = C.c_int64(0)
        check(lib().dsurf_plan_allgather(self.h, comm.h, C.c_int(comm.rank), C.c_int(comm.world),
                                         C.c_int(1 if want_coo else 0), C.byref(tot)), "plan_allgather")
        return int(tot.value)

    @property
    def gather_ms(self):
        return float(lib().dsurf_plan_last_gather_ms(self.h))

    def download_gathered(self, nar):
        rows, rw, col = np.zeros(max(nar, 1), I32), np.zeros(max(nar, 1), F32), np.zeros(max(nar, 1), I32)
        check(lib().dsurf_plan_download_gathered(self.h, ptr(rows, C.c_int), ptr(rw, C.c_float), ptr(col, C.c_int)),
              "plan_download_gathered")
        return dict(nar=nar, row=rows[:nar], rw=rw[:nar], col=col[:nar])

    def digest(self, gathered=False):
        """(64-bit order-sensitive digest, number of triplets) of the plan's COO or of the gathered COO."""
        d, n = C.c_uint64(0), C.c_int64(0)
        check(lib().dsurf_plan_digest(self.h, C.c_int(1 if gathered else 0), C.byref(d), C.byref(n)), "plan_digest")
        return int(d.value), int(n.value)

    def glue_results(self, want_vectors=True):
        """cbst / datweight / statistics left by LsmrSystem.from_plan (main.f90:361-394)."""
        dall = self.pb.dall
        cb = np.zeros(dall, F32) if want_vectors else None
        dw = np.zeros(dall, F32) if want_vectors else None
        st = np.zeros(4, F32)
        m, nar = C.c_int(0), C.c_int64(0)
        check(lib().dsurf_plan_glue_results(self.h, ptr(cb, C.c_float), ptr(dw, C.c_float), ptr(st, C.c_float),
                                            C.byref(m), C.byref(nar)), "plan_glue_results")
        return dict(cbst=cb, datweight=dw, q25=float(st[0]), q75=float(st[1]), maxnorm=float(st[2]),
                    averdws=float(st[3]), m=m.value, nar=nar.value)

    def update_model(self, lsmr_system, minvel=None, maxvel=None):
        """main.f90:518-532 on the device; returns (updated model [nz][ny][nx], clipped dv)."""
        pb = self.pb
        dv = np.zeros(pb.maxvp, F32)
        vs = np.zeros((pb.nz, pb.ny, pb.nx), F32)
        check(lib().dsurf_plan_update_model(
            self.h, lsmr_system.h, C.c_float(pb.minvel if minvel is None else minvel),
            C.c_float(pb.maxvel if maxvel is None else maxvel), ptr(dv, C.c_float), ptr(vs, C.c_float)),
            "plan_update_model")
        return vs, dv

    def timings(self):
        ms = np.zeros(8, F64)
        check(lib().dsurf_plan_timings(self.h, ptr(ms, C.c_double)), "plan_timings")
        return dict(dispersion_ms=ms[0], dice_ms=ms[1], eikonal_ms=ms[2], rays_ms=ms[3], assembly_ms=ms[4],
                    eikonal_launches=int(ms[5]), launches=int(ms[6]), sweeps=int(ms[7]),
                    total_ms=float(lib().dsurf_plan_last_sweeps_ms(self.h)))

    def get_dispersion(self, type_):
        pb = self.pb
        ncol = pb.nx * pb.ny
        kt = (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)[type_]
        pvcols = pb.kmax if (type_ in (0, 2) and not self.forward) else max(kt, 1)
        pv = np.zeros((pvcols, ncol), F64)
        sen = [np.zeros((pb.nz, max(kt, 1), ncol), F64) for _ in range(3)]
        check(lib().dsurf_plan_get_dispersion(self.h, C.c_int(type_), ptr(pv, C.c_double),
                                              *[ptr(s, C.c_double) for s in sen]), "plan_get_dispersion")
        return pv, sen[0], sen[1], sen[2]

    def debug_sweep(self, g, ig=1, want_fdm=True):
        pb = self.pb
        nnx, nnz = (pb.nx - 3) * 8 + 1, (pb.ny - 3) * 8 + 1
        veln = np.zeros((nnx, nnz), F32)
        ttn = np.zeros((nnx, nnz), F32)
        ttnr = np.zeros(129 * 129, F32)
        nstsr = np.zeros(129 * 129, I32)
        rgeom = np.zeros(6, F32)
        fdm = np.zeros((pb.nrc, pb.nx, pb.ny), F32) if want_fdm else None
        check(lib().dsurf_plan_debug_sweep(self.h, C.c_int(g), C.c_int(ig), ptr(veln, C.c_float),
                                           ptr(ttn, C.c_float), ptr(ttnr, C.c_float), ptr(nstsr, C.c_int),
                                           ptr(rgeom, C.c_float), ptr(fdm, C.c_float) if want_fdm else None),
              "plan_debug_sweep")
        nnxr, nnzr = int(rgeom[4]), int(rgeom[5])
        return dict(veln=veln, ttn=ttn, ttnr=ttnr[: nnxr * nnzr].reshape(nnxr, nnzr),
                    nstsr=nstsr[: nnxr * nnzr].reshape(nnxr, nnzr), rgeom=rgeom, fdm=fdm)


class LsmrSystem:
    """Device-resident sparse system for LSMR (CSR + CSC copies in HBM)."""

    def __init__(self, m, n, rows1, cols1, vals, b):
        rows1, cols1 = _c(rows1, I32), _c(cols1, I32)
        vals, b = _c(vals, F32), _c(b, F32)
        h = C.c_void_p()
        check(lib().dsurf_lsmr_create(C.byref(h), C.c_int(m), C.c_int(n), C.c_int64(len(vals)),
                                      ptr(rows1, C.c_int), ptr(cols1, C.c_int), ptr(vals, C.c_float),
                                      ptr(b, C.c_float)), "lsmr_create")
        self.h, self.m, self.n, self.nnz = h, m, n, len(vals)

    @classmethod
    def from_plan(cls, plan, obst=None, threshold0=None, weight=None):
        """Device-resident host glue (main.f90:361-466): the plan's COO stays in HBM."""
        pb = plan.pb
        obst = _c(pb.obst if obst is None else obst, F32)
        h = C.c_void_p()
        check(lib().dsurf_lsmr_create_from_plan(
            C.byref(h), plan.h, ptr(obst, C.c_float),
            C.c_float(pb.threshold if threshold0 is None else threshold0),
            C.c_float(pb.weight if weight is None else weight)), "lsmr_create_from_plan")
        self = cls.__new__(cls)
        g = plan.glue_results(want_vectors=False)
        self.h, self.m, self.n, self.nnz = h, g["m"], pb.maxvp, g["nar"]
        return self

    @classmethod
    def from_plan_shard(cls, plan, rank, world, obst=None, threshold0=None, weight=None):
        """Row-partitioned system of rank `rank` (its own data rows + a share of the smoothing rows); the plan must
        hold the predicted times of all rows (Plan.allgather)."""
        pb = plan.pb
        obst = _c(pb.obst if obst is None else obst, F32)
        h, m, nnz = C.c_void_p(), C.c_int(0), C.c_int64(0)
        check(lib().dsurf_lsmr_create_from_plan_shard(
            C.byref(h), plan.h, ptr(obst, C.c_float), C.c_float(pb.threshold if threshold0 is None else threshold0),
            C.c_float(pb.weight if weight is None else weight), C.c_int(rank), C.c_int(world), C.byref(m),
            C.byref(nnz)), "lsmr_create_from_plan_shard")
        self = cls.__new__(cls)
        self.h, self.m, self.n, self.nnz = h, m.value, pb.maxvp, nnz.value
        return self

    def close(self):
        if self.h:
            lib().dsurf_lsmr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, damp, atol=1e-6, btol=1e-6, conlim=100.0, itnlim=400, localSize=10, force_iters=False,
              want_x=True):
        x = np.zeros(self.n, F32) if want_x else None
        istop, itn = C.c_int(0), C.c_int(0)
        vals = [C.c_float(0) for _ in range(5)]
        ms = [C.c_double(0) for _ in range(3)]
        check(lib().dsurf_lsmr_solve(self.h, C.c_float(damp), C.c_float(atol), C.c_float(btol),
                                     C.c_float(conlim), C.c_int(itnlim), C.c_int(localSize),
                                     C.c_int(1 if force_iters else 0), ptr(x, C.c_float) if want_x else None,
                                     C.byref(istop), C.byref(itn), *[C.byref(v) for v in vals],
                                     *[C.byref(v) for v in ms]), "lsmr_solve")
        return dict(x=x, istop=istop.value, itn=itn.value, normA=vals[0].value, condA=vals[1].value,
                    normr=vals[2].value, normAr=vals[3].value, normx=vals[4].value, ms_total=ms[0].value,
                    ms_spmv=ms[1].value, ms_spmtv=ms[2].value)


def lsmr_hint_geometry(nx, ny, nz):
    """Column-order hint for LSMR (columns = k*P + vertex, P=(nx-2)(ny-2), K=nz-1): enables the
    depth-blocked sparse layout for systems with n == P*K.  CalSurfG/Plan set it automatically."""
    check(lib().dsurf_lsmr_hint_geometry(C.c_int(nx), C.c_int(ny), C.c_int(nz)), "lsmr_hint_geometry")
