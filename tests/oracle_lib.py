"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke()
import this module.  The product package dsurftomo_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(_HERE, "..", "oracle")
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_snrm2.restype = C.c_float
        _LIB.oracle_delsph.restype = C.c_float
        _LIB.oracle_delsph.argtypes = [C.c_float] * 4
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def surfdisp96(thk, vp, vs, rho, iflsph, iwave, mode, igr, t):
    thk, vp, vs, rho = (np.ascontiguousarray(a, np.float32) for a in (thk, vp, vs, rho))
    t = np.ascontiguousarray(t, np.float64)
    cg = np.zeros(len(t), np.float64)
    nf = lib().oracle_surfdisp96(
        _p(thk, C.c_float), _p(vp, C.c_float), _p(vs, C.c_float), _p(rho, C.c_float),
        C.c_int(len(thk)), C.c_int(iflsph), C.c_int(iwave), C.c_int(mode), C.c_int(igr),
        C.c_int(len(t)), _p(t, C.c_double), _p(cg, C.c_double))
    return cg, nf


def refine_grid2layer(minthk, dep, vp, vs, rho):
    dep, vp, vs, rho = (np.ascontiguousarray(a, np.float32) for a in (dep, vp, vs, rho))
    out = [np.zeros(200, np.float32) for _ in range(5)]
    rmax = C.c_int(0)
    lib().oracle_refine_grid2layer(
        C.c_float(minthk), C.c_int(len(dep)), _p(dep, C.c_float), _p(vp, C.c_float),
        _p(vs, C.c_float), _p(rho, C.c_float), C.byref(rmax), *[_p(o, C.c_float) for o in out])
    n = rmax.value
    rdep, rvp, rvs, rrho, rthk = (o[:n] for o in out)
    return rdep, rvp, rvs, rrho, rthk


def depthkernel(vel, iwave, igr, t, depz, minthk, nthreads=1):
    vel = np.ascontiguousarray(vel, np.float32)
    nz, ny, nx = vel.shape
    t = np.ascontiguousarray(t, np.float64)
    depz = np.ascontiguousarray(depz, np.float32)
    k = len(t)
    pv = np.zeros((k, ny * nx), np.float64)
    sen = [np.zeros((nz, k, ny * nx), np.float64) for _ in range(3)]
    lib().oracle_depthkernel(
        C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(vel, C.c_float), _p(pv, C.c_double),
        _p(sen[0], C.c_double), _p(sen[1], C.c_double), _p(sen[2], C.c_double), C.c_int(iwave),
        C.c_int(igr), C.c_int(k), _p(t, C.c_double), _p(depz, C.c_float), C.c_float(minthk),
        C.c_int(nthreads))
    return pv, sen[0], sen[1], sen[2]


def caldespersion(vel, iwave, igr, t, depz, minthk, nthreads=1):
    vel = np.ascontiguousarray(vel, np.float32)
    nz, ny, nx = vel.shape
    t = np.ascontiguousarray(t, np.float64)
    depz = np.ascontiguousarray(depz, np.float32)
    pv = np.zeros((len(t), ny * nx), np.float64)
    lib().oracle_caldespersion(
        C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(vel, C.c_float), _p(pv, C.c_double),
        C.c_int(iwave), C.c_int(igr), C.c_int(len(t)), _p(t, C.c_double), _p(depz, C.c_float),
        C.c_float(minthk), C.c_int(nthreads))
    return pv


def grid_dims(nx, ny):
    return (nx - 3) * 8 + 1, (ny - 3) * 8 + 1  # nnx, nnz


def fmm_sweep(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz):
    """Returns dict(veln[nnx,nnz], ttn[nnx,nnz], ttnr, nstsr, rgeom)."""
    nnx, nnz = grid_dims(nx, ny)
    pv = np.ascontiguousarray(pv, np.float64)
    veln = np.zeros((nnx, nnz), np.float32)
    ttn = np.zeros((nnx, nnz), np.float32)
    ttnr = np.zeros((129, 129), np.float32)
    nstsr = np.zeros((129, 129), np.int32)
    rgeom = np.zeros(6, np.float32)
    err = lib().oracle_fmm_sweep(
        C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd), C.c_float(dvzd),
        _p(pv, C.c_double), C.c_float(scx), C.c_float(scz), _p(veln, C.c_float), _p(ttn, C.c_float),
        _p(ttnr, C.c_float), _p(nstsr, C.c_int), _p(rgeom, C.c_float))
    nnxr, nnzr = int(rgeom[4]), int(rgeom[5])
    return dict(err=err, veln=veln, ttn=ttn, ttnr=ttnr.ravel()[: nnxr * nnzr].reshape(nnxr, nnzr),
                nstsr=nstsr.ravel()[: nnxr * nnzr].reshape(nnxr, nnzr), rgeom=rgeom)


def sweep_rays(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, rcx, rcz):
    pv = np.ascontiguousarray(pv, np.float64)
    rcx = np.ascontiguousarray(rcx, np.float32)
    rcz = np.ascontiguousarray(rcz, np.float32)
    nrc = len(rcx)
    tt = np.zeros(nrc, np.float32)
    fdm = np.zeros((nrc, nx, ny), np.float32)  # [ray][j (x vertex 0..nvx+1)][i (z vertex)]
    err = lib().oracle_sweep_rays(
        C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd), C.c_float(dvzd),
        _p(pv, C.c_double), C.c_float(scx), C.c_float(scz), C.c_int(nrc), _p(rcx, C.c_float),
        _p(rcz, C.c_float), _p(tt, C.c_float), _p(fdm, C.c_float))
    return err, tt, fdm


def sweep_paths(nx, ny, goxd, gozd, dvxd, dvzd, pv, scx, scz, rcx, rcz, cap=4096):
    """Ray geometry (colatitude, longitude in radians) per receiver: list of (n,2) arrays."""
    pv = np.ascontiguousarray(pv, np.float64)
    rcx = np.ascontiguousarray(rcx, np.float32)
    rcz = np.ascontiguousarray(rcz, np.float32)
    nrc = len(rcx)
    npts = np.zeros(nrc, np.int32)
    px = np.zeros((nrc, cap), np.float32)
    pz = np.zeros((nrc, cap), np.float32)
    err = lib().oracle_sweep_paths(
        C.c_int(nx), C.c_int(ny), C.c_float(goxd), C.c_float(gozd), C.c_float(dvxd), C.c_float(dvzd),
        _p(pv, C.c_double), C.c_float(scx), C.c_float(scz), C.c_int(nrc), _p(rcx, C.c_float),
        _p(rcz, C.c_float), C.c_int(cap), _p(npts, C.c_int), _p(px, C.c_float), _p(pz, C.c_float))
    assert err == 0 and npts.max() <= cap
    return [np.stack([px[r, : npts[r]], pz[r, : npts[r]]], axis=1) for r in range(nrc)]


def calsurfg(pb, vels=None, nthreads=1, mode=0, maxnar=None):
    """Run the oracle CalSurfG on a Problem.  Returns dict(dsurf, rw, row, col, nar, ...)."""
    vels = np.ascontiguousarray(pb.vsf if vels is None else vels, np.float32)
    if maxnar is None:
        maxnar = max(pb.maxnar(), 1)
    iw = np.zeros(2 * maxnar + 1, np.int32)
    rw = np.zeros(maxnar, np.float32)
    col = np.zeros(maxnar, np.int32)
    dsurf = np.zeros(pb.dall, np.float32)
    nar = C.c_int(0)
    rbint = C.c_int(0)
    st = np.zeros(8, np.float64)
    dummy = np.zeros(1, np.float64)

    def tp(t):
        return _p(np.ascontiguousarray(t if len(t) else dummy, np.float64), C.c_double)

    tRc, tRg, tLc, tLg = (np.ascontiguousarray(t if len(t) else dummy, np.float64)
                          for t in (pb.tRc, pb.tRg, pb.tLc, pb.tLg))
    err = lib().oracle_calsurfg(
        C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), C.c_int(pb.maxvp), _p(vels, C.c_float),
        _p(iw, C.c_int), _p(rw, C.c_float), _p(col, C.c_int), _p(dsurf, C.c_float),
        C.c_float(pb.goxd), C.c_float(pb.gozd), C.c_float(pb.dvxd), C.c_float(pb.dvzd),
        C.c_int(pb.kmaxRc), C.c_int(pb.kmaxRg), C.c_int(pb.kmaxLc), C.c_int(pb.kmaxLg),
        _p(tRc, C.c_double), _p(tRg, C.c_double), _p(tLc, C.c_double), _p(tLg, C.c_double),
        _p(pb.wavetype, C.c_int), _p(pb.igrt, C.c_int), _p(pb.periods, C.c_int),
        _p(pb.depz, C.c_float), C.c_float(pb.minthk), _p(pb.scxf, C.c_float), _p(pb.sczf, C.c_float),
        _p(pb.rcxf, C.c_float), _p(pb.rczf, C.c_float), _p(pb.nrc1, C.c_int), _p(pb.nsrc1, C.c_int),
        C.c_int(pb.kmax), C.c_int(pb.nsrc), C.c_int(pb.nrc), C.byref(nar), C.c_int(nthreads),
        C.c_int(mode), C.byref(rbint), _p(st, C.c_double))
    n = nar.value
    return dict(err=err, nar=n, dsurf=dsurf, rw=rw[:n].copy(), row=iw[1 : n + 1].copy(),
                col=col[:n].copy(), iw=iw, rw_full=rw, col_full=col, rbint=rbint.value,
                t_disp=st[0], t_gather=st[1], t_fmm=st[2], t_ray=st[3], nsweeps=int(st[4]),
                nrays=int(st[5]))


def synthetic(pb, vels=None, nthreads=1):
    """Oracle `synthetic` (noise-free): forward times through `vels` on the gd = 5 grid."""
    vels = np.ascontiguousarray(pb.vsf if vels is None else vels, np.float32)
    obst = np.zeros(pb.dall, np.float32)
    dummy = np.zeros(1, np.float64)
    tRc, tRg, tLc, tLg = (np.ascontiguousarray(t if len(t) else dummy, np.float64)
                          for t in (pb.tRc, pb.tRg, pb.tLc, pb.tLg))
    pv = np.zeros(max(1, pb.kmax) * pb.nx * pb.ny, np.float64)
    err = lib().oracle_synthetic(
        C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), _p(vels, C.c_float), _p(obst, C.c_float),
        C.c_float(pb.goxd), C.c_float(pb.gozd), C.c_float(pb.dvxd), C.c_float(pb.dvzd),
        C.c_int(pb.kmaxRc), C.c_int(pb.kmaxRg), C.c_int(pb.kmaxLc), C.c_int(pb.kmaxLg),
        _p(tRc, C.c_double), _p(tRg, C.c_double), _p(tLc, C.c_double), _p(tLg, C.c_double),
        _p(pb.wavetype, C.c_int), _p(pb.igrt, C.c_int), _p(pb.periods, C.c_int),
        _p(pb.depz, C.c_float), C.c_float(pb.minthk), _p(pb.scxf, C.c_float), _p(pb.sczf, C.c_float),
        _p(pb.rcxf, C.c_float), _p(pb.rczf, C.c_float), _p(pb.nrc1, C.c_int), _p(pb.nsrc1, C.c_int),
        C.c_int(pb.kmax), C.c_int(pb.nsrc), C.c_int(pb.nrc), C.c_int(nthreads), _p(pv, C.c_double))
    return dict(err=err, obst=obst, pv=pv.reshape(max(1, pb.kmax), pb.nx * pb.ny))


def calsurfg_pre(pb, pv4, sen12, g_lo=-1, g_hi=-1, vels=None, nthreads=1, mode=1, maxnar=None):
    """Gather loop of CalSurfG on caller-provided dispersion results (bench CPU arm).
    pv4: 4 arrays [cols, ncol] (Rc, Rg, Lc, Lg); sen12: 12 arrays [nz, kmax_t, ncol] in the order
    Rc(vs,vp,rho), Rg(...), Lc(...), Lg(...).  Entries of absent types may be None."""
    vels = np.ascontiguousarray(pb.vsf if vels is None else vels, np.float32)
    if maxnar is None:
        maxnar = max(pb.maxnar(), 1)
    iw = np.zeros(2 * maxnar + 1, np.int32)
    rw = np.zeros(maxnar, np.float32)
    col = np.zeros(maxnar, np.int32)
    dsurf = np.zeros(pb.dall, np.float32)
    nar, rbint = C.c_int(0), C.c_int(0)
    st = np.zeros(8, np.float64)
    dummy = np.zeros(1, np.float64)
    keep = [np.ascontiguousarray(a if a is not None else dummy, np.float64) for a in list(pv4) + list(sen12)]
    pvp = (C.POINTER(C.c_double) * 4)(*[_p(a, C.c_double) for a in keep[:4]])
    senp = (C.POINTER(C.c_double) * 12)(*[_p(a, C.c_double) for a in keep[4:]])
    err = lib().oracle_calsurfg_pre(
        C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), _p(vels, C.c_float), _p(iw, C.c_int), _p(rw, C.c_float),
        _p(col, C.c_int), _p(dsurf, C.c_float), C.c_float(pb.goxd), C.c_float(pb.gozd), C.c_float(pb.dvxd),
        C.c_float(pb.dvzd), C.c_int(pb.kmaxRc), C.c_int(pb.kmaxRg), C.c_int(pb.kmaxLc), C.c_int(pb.kmaxLg),
        pvp, senp, _p(pb.wavetype, C.c_int), _p(pb.igrt, C.c_int), _p(pb.periods, C.c_int),
        _p(pb.depz, C.c_float), _p(pb.scxf, C.c_float), _p(pb.sczf, C.c_float), _p(pb.rcxf, C.c_float),
        _p(pb.rczf, C.c_float), _p(pb.nrc1, C.c_int), _p(pb.nsrc1, C.c_int), C.c_int(pb.kmax), C.c_int(pb.nsrc),
        C.c_int(pb.nrc), C.byref(nar), C.c_int(nthreads), C.c_int(mode), C.byref(rbint), _p(st, C.c_double),
        C.c_int(g_lo), C.c_int(g_hi))
    n = nar.value
    return dict(err=err, nar=n, dsurf=dsurf, rw=rw[:n].copy(), row=iw[1:n + 1].copy(), col=col[:n].copy(),
                rbint=rbint.value, t_gather=st[1], t_fmm=st[2], t_ray=st[3], nsweeps=int(st[4]), nrays=int(st[5]))


def pack_iw(row, col):
    """iw = [nar | rows | cols] as main.f90:463-466."""
    nar = len(row)
    iw = np.zeros(2 * nar + 1, np.int32)
    iw[0] = nar
    iw[1 : nar + 1] = row
    iw[nar + 1 :] = col
    return iw


def aprod(mode, m, n, x, y, iw, rw):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    iw = np.ascontiguousarray(iw, np.int32)
    rw = np.ascontiguousarray(rw, np.float32)
    lib().oracle_aprod(C.c_int(mode), C.c_int(m), C.c_int(n), _p(x, C.c_float), _p(y, C.c_float),
                       C.c_int(len(iw)), C.c_int(len(rw)), _p(iw, C.c_int), _p(rw, C.c_float))
    return x, y


def lsmr(m, n, iw, rw, b, damp, atol=1e-6, btol=1e-6, conlim=100.0, itnlim=400, localSize=10):
    iw = np.ascontiguousarray(iw, np.int32)
    rw = np.ascontiguousarray(rw, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    x = np.zeros(n, np.float32)
    istop, itn = C.c_int(0), C.c_int(0)
    nA, cA, nr, nAr, nx_ = (C.c_float(0) for _ in range(5))
    lib().oracle_lsmr(C.c_int(m), C.c_int(n), C.c_int(len(iw)), C.c_int(len(rw)), _p(iw, C.c_int),
                      _p(rw, C.c_float), _p(b, C.c_float), C.c_float(damp), C.c_float(atol),
                      C.c_float(btol), C.c_float(conlim), C.c_int(itnlim), C.c_int(localSize),
                      _p(x, C.c_float), C.byref(istop), C.byref(itn), C.byref(nA), C.byref(cA),
                      C.byref(nr), C.byref(nAr), C.byref(nx_))
    return dict(x=x, istop=istop.value, itn=itn.value, normA=nA.value, condA=cA.value,
                normr=nr.value, normAr=nAr.value, normx=nx_.value)


def snrm2(x):
    x = np.ascontiguousarray(x, np.float32)
    return float(lib().oracle_snrm2(C.c_int(len(x)), _p(x, C.c_float)))


def getpercentile(a):
    a = np.ascontiguousarray(a, np.float32)
    q25, q75 = C.c_float(0), C.c_float(0)
    lib().oracle_getpercentile(C.c_int(len(a)), _p(a, C.c_float), C.byref(q25), C.byref(q75))
    return q25.value, q75.value


def host_glue(pb, dsyn, iw, rw, col, nar):
    """main.f90:361-466 on the arrays returned by calsurfg (capacity must allow the
    smoothing rows).  Returns m, nar, cbst, datweight (iw/rw/col modified in place)."""
    cbst = np.zeros(pb.dall + pb.maxvp, np.float32)
    datw = np.zeros(pb.dall, np.float32)
    nar_c = C.c_int(nar)
    m = lib().oracle_host_glue(
        C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), C.c_int(pb.dall), _p(pb.obst, C.c_float),
        _p(np.ascontiguousarray(dsyn, np.float32), C.c_float), C.c_float(pb.threshold),
        C.c_float(pb.weight), _p(iw, C.c_int), _p(rw, C.c_float), _p(col, C.c_int),
        _p(cbst, C.c_float), _p(datw, C.c_float), C.byref(nar_c))
    return m, nar_c.value, cbst, datw


def model_update(pb, vsf, dv):
    vsf = np.ascontiguousarray(vsf, np.float32).copy()
    dv = np.ascontiguousarray(dv, np.float32).copy()
    lib().oracle_model_update(C.c_int(pb.nx), C.c_int(pb.ny), C.c_int(pb.nz), _p(vsf, C.c_float),
                              _p(dv, C.c_float), C.c_float(pb.minvel), C.c_float(pb.maxvel))
    return vsf, dv
