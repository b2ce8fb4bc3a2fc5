"""Calls the gfortran-mangled drop-in symbols exactly as a gfortran-compiled main.o would: every
argument by reference, column-major arrays, no hidden arguments -- and compares with the neutral
dsurf_* entry points.  (ctypes plays the role of the Fortran caller; no Fortran compiler exists in
this image.)"""
import ctypes as C

import numpy as np
import pytest

from dsurftomo_b200 import _lib, api, hostglue, inputs

pytestmark = pytest.mark.gpu


def ri(v):
    return C.byref(C.c_int(v))


def rf(v):
    return C.byref(C.c_float(v))


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_surfdisp96_mangled():
    L = _lib.lib()
    thk = np.array([1.0, 2.0, 0.0], np.float32)
    vs = np.array([1.5, 2.5, 3.5], np.float32)
    vp = (vs * 1.8).astype(np.float32)
    rho = np.array([2.2, 2.5, 2.9], np.float32)
    t = np.array([1.0, 2.0, 4.0])
    cg = np.zeros(3)
    L.surfdisp96_.restype = None
    L.surfdisp96_(p(thk, C.c_float), p(vp, C.c_float), p(vs, C.c_float), p(rho, C.c_float), ri(3), ri(1), ri(2),
                  ri(1), ri(0), ri(3), p(t, C.c_double), p(cg, C.c_double))
    ref = api.surfdisp96(thk, vp, vs, rho, 3, 1, 2, 1, 0, 3, t)
    assert np.array_equal(cg, ref) and np.all(cg > 1.0)


def test_calsurfg_and_lsmr_mangled():
    L = _lib.lib()
    pb = inputs.synthetic_problem(10, 2, 4, ("Rc",), nrecv=3, name="abi")
    k = api._plan_args(pb, None)
    maxnar = 400000
    iw = np.zeros(2 * maxnar + 1, np.int32)
    rw = np.zeros(maxnar, np.float32)
    col = np.zeros(maxnar, np.int32)
    dsurf = np.zeros(pb.dall, np.float32)
    nar = C.c_int(0)
    L.calsurfg_.restype = None
    L.calsurfg_(ri(pb.nx), ri(pb.ny), ri(pb.nz), ri(pb.maxvp), p(k["vels"], C.c_float), p(iw, C.c_int),
                p(rw, C.c_float), p(col, C.c_int), p(dsurf, C.c_float), rf(pb.goxd), rf(pb.gozd), rf(pb.dvxd),
                rf(pb.dvzd), ri(pb.kmaxRc), ri(0), ri(0), ri(0), p(k["tRc"], C.c_double), None, None, None,
                p(k["wavetype"], C.c_int), p(k["igrt"], C.c_int), p(k["periods"], C.c_int), p(k["depz"], C.c_float),
                rf(pb.minthk), p(k["scxf"], C.c_float), p(k["sczf"], C.c_float), p(k["rcxf"], C.c_float),
                p(k["rczf"], C.c_float), p(k["nrc1"], C.c_int), p(k["nsrc1"], C.c_int), ri(pb.kmax), ri(pb.nsrc),
                ri(pb.nrc), C.byref(nar))
    ref = api.CalSurfG(pb, maxnar=maxnar)
    n = nar.value
    assert n == ref["nar"] > 0
    assert np.array_equal(iw[1:n + 1], ref["row"]) and np.array_equal(col[:n], ref["col"])
    assert np.array_equal(rw[:n], ref["rw"]) and np.array_equal(dsurf, ref["dsurf"])
    s = hostglue.host_glue(pb, dsurf, iw[1:n + 1], col[:n], rw[:n])
    iwp = hostglue.pack_iw(s["rows"], s["cols"])
    x = np.zeros(s["n"], np.float32)
    istop, itn = C.c_int(0), C.c_int(0)
    outs = [C.c_float(0) for _ in range(5)]
    f = getattr(L, "__lsmrmodule_MOD_lsmr")
    f.restype = None
    f(ri(s["m"]), ri(s["n"]), ri(len(iwp)), ri(len(s["vals"])), p(iwp, C.c_int), p(s["vals"], C.c_float),
      p(s["cbst"], C.c_float), rf(pb.damp), rf(1e-6), rf(1e-6), rf(100.0), ri(400), ri(10), ri(-12345), p(x, C.c_float),
      C.byref(istop), C.byref(itn), *[C.byref(o) for o in outs])
    refl = api.LSMR(s["m"], s["n"], len(iwp), len(s["vals"]), iwp, s["vals"], s["cbst"], pb.damp, 1e-6, 1e-6, 100.0,
                    400, 10)
    assert itn.value == refl["itn"] and istop.value == refl["istop"] and np.array_equal(x, refl["x"])
    # aprod_: y += A x
    xx = np.ones(s["n"], np.float32)
    yy = np.zeros(s["m"], np.float32)
    L.aprod_.restype = None
    L.aprod_(ri(1), ri(s["m"]), ri(s["n"]), p(xx, C.c_float), p(yy, C.c_float), ri(len(iwp)), ri(len(s["vals"])),
             p(iwp, C.c_int), p(s["vals"], C.c_float))
    _, yref = api.aprod(1, s["m"], s["n"], np.ones(s["n"], np.float32), np.zeros(s["m"], np.float32), len(iwp),
                        len(s["vals"]), iwp, s["vals"])
    assert np.array_equal(yy, yref)
