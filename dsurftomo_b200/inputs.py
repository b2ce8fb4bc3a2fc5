"""Host-side input handling that mirrors the reference's main program.

* :func:`read_problem` follows ``src/main.f90:134-321`` (list-directed parse of
  ``DSurfTomo.in``, the '#'-gather data file and ``MOD``), including the REAL*4 conversion of
  station coordinates to colatitude/longitude radians (main.f90:258-259, 271-272) and
  ``obst = delsph distance / velocity`` (main.f90:275-277, delsph.f90).
* :func:`synthetic_problem` builds the seeded synthetic configurations of SURVEY.md section 8(d)
  / BASELINE.md section 3 (cfg 2, 3, 5 and scaled-down variants for tests).

Array layout: every array is stored exactly as the Fortran caller would hold it in memory
(column-major), expressed as C-ordered numpy arrays with the axes reversed, e.g. the Fortran
``scxf(nsrc,kmax)`` is ``scxf[kmax, nsrc]`` here and ``vsf(nx,ny,nz)`` is ``vsf[nz, ny, nx]``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32
PI32 = F32(3.1415926535898)  # main.f90:57, delsph.f90:4 (rounds to 3.14159274)


@dataclass
class Problem:
    nx: int
    ny: int
    nz: int
    goxd: float
    gozd: float
    dvxd: float
    dvzd: float
    nsrc: int  # = nrc (main.f90:213)
    weight: float
    damp: float
    minthk: float
    minvel: float
    maxvel: float
    maxiter: int
    spfra: float
    kmaxRc: int
    kmaxRg: int
    kmaxLc: int
    kmaxLg: int
    tRc: np.ndarray
    tRg: np.ndarray
    tLc: np.ndarray
    tLg: np.ndarray
    ifsyn: int
    noiselevel: float
    threshold: float
    scxf: np.ndarray  # [kmax, nsrc] f32 colatitude (rad)
    sczf: np.ndarray  # [kmax, nsrc] f32 longitude (rad)
    rcxf: np.ndarray  # [kmax, nsrc, nrc] f32
    rczf: np.ndarray
    periods: np.ndarray  # [kmax, nsrc] i32, 1-based index inside the type's period list
    wavetype: np.ndarray  # [kmax, nsrc] i32 (2 Rayleigh, 1 Love)
    igrt: np.ndarray  # [kmax, nsrc] i32 (0 phase, 1 group)
    nrc1: np.ndarray  # [kmax, nsrc] i32 receivers per gather
    nsrc1: np.ndarray  # [kmax] i32 gathers per period-type
    obst: np.ndarray  # [dall] f32 observed times
    dist: np.ndarray  # [dall] f32
    depz: np.ndarray  # [nz] f32
    vsf: np.ndarray  # [nz, ny, nx] f32
    name: str = "problem"
    meta: dict = field(default_factory=dict)

    @property
    def kmax(self) -> int:
        return self.kmaxRc + self.kmaxRg + self.kmaxLc + self.kmaxLg

    @property
    def nrc(self) -> int:
        return self.nsrc

    @property
    def dall(self) -> int:
        return int(self.obst.shape[0])

    @property
    def maxvp(self) -> int:
        return (self.nx - 2) * (self.ny - 2) * (self.nz - 1)

    @property
    def ngathers(self) -> int:
        return int(self.nsrc1.sum())

    @property
    def nsweeps(self) -> int:
        """FMM solves per CalSurfG call: one per phase gather, two per group gather
        (CalSurfG.f90:1172-1185)."""
        n = 0
        for k in range(self.kmax):
            for s in range(int(self.nsrc1[k])):
                n += 2 if self.igrt[k, s] == 1 else 1
        return n

    def maxnar(self) -> int:
        # main.f90:287 (REAL*4 product converted to INTEGER)
        return int(F32(self.spfra) * F32(self.dall) * F32(self.nx) * F32(self.ny) * F32(self.nz))


_LIBM = None


def _libm_f32(name, nargs):
    """REAL*4 intrinsic of gfortran == the C library's float routine (sinf, cosf, atan2f), applied
    element-wise; numpy's own float32 kernels differ from glibc by an ulp on some arguments."""
    global _LIBM
    import ctypes
    import ctypes.util

    if _LIBM is None:
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    fn = getattr(_LIBM, name)
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float] * nargs
    uf = np.frompyfunc(lambda *a: fn(*[float(v) for v in a]), nargs, 1)
    return lambda *a: np.asarray(uf(*a), dtype=np.float64).astype(F32)


def delsph(flat1, flon1, flat2, flon2):
    """delsph.f90:1-28 in REAL*4 (inputs: colatitude, longitude in radians)."""
    flat1, flon1, flat2, flon2 = (np.asarray(v, dtype=F32) for v in (flat1, flon1, flat2, flon2))
    sin, cos, atan2 = _libm_f32("sinf", 1), _libm_f32("cosf", 1), _libm_f32("atan2f", 2)
    R = F32(6371.0)
    dlat = flat2 - flat1
    dlon = flon2 - flon1
    lat1 = PI32 / F32(2) - flat1
    lat2 = PI32 / F32(2) - flat2
    s1, s2 = sin(dlat / F32(2)), sin(dlon / F32(2))
    a = (s1 * s1 + s2 * s2 * cos(lat1) * cos(lat2)).astype(F32)
    c = (F32(2) * atan2(np.sqrt(a), np.sqrt(F32(1) - a))).astype(F32)
    return (R * c).astype(F32)


def _tokens(line: str):
    return line.replace(",", " ").split()


def read_problem(inputfile: str) -> Problem:
    """Parse DSurfTomo.in + data file + MOD exactly as main.f90:134-321 does."""
    base = os.path.dirname(os.path.abspath(inputfile))
    with open(inputfile) as fh:
        lines = fh.read().splitlines()
    it = iter(lines[3:])  # three comment lines, main.f90:135-137
    datafile = _tokens(next(it))[0]
    nx, ny, nz = (int(v) for v in _tokens(next(it))[:3])
    goxd, gozd = (float(v) for v in _tokens(next(it))[:2])
    dvxd, dvzd = (float(v) for v in _tokens(next(it))[:2])
    nsrc = int(_tokens(next(it))[0])
    weight, damp = (float(v) for v in _tokens(next(it))[:2])
    minthk = float(_tokens(next(it))[0])
    minvel, maxvel = (float(v) for v in _tokens(next(it))[:2])
    maxiter = int(_tokens(next(it))[0])
    spfra = float(_tokens(next(it))[0])

    def read_type():
        k = int(_tokens(next(it))[0])
        t = np.zeros(0, dtype=np.float64)
        if k > 0:
            vals = []
            while len(vals) < k:
                vals += [float(v) for v in _tokens(next(it)) if _isnum(v)]
            t = np.array(vals[:k], dtype=np.float64)
        return k, t

    kmaxRc, tRc = read_type()
    kmaxRg, tRg = read_type()
    kmaxLc, tLc = read_type()
    kmaxLg, tLg = read_type()
    ifsyn = int(_tokens(next(it))[0])
    noiselevel = float(_tokens(next(it))[0])
    threshold = float(_tokens(next(it))[0])
    kmax = kmaxRc + kmaxRg + kmaxLc + kmaxLg
    nrc = nsrc

    scxf = np.zeros((kmax, nsrc), F32)
    sczf = np.zeros((kmax, nsrc), F32)
    rcxf = np.zeros((kmax, nsrc, nrc), F32)
    rczf = np.zeros((kmax, nsrc, nrc), F32)
    periods = np.zeros((kmax, nsrc), np.int32)
    wavetype = np.zeros((kmax, nsrc), np.int32)
    igrt = np.zeros((kmax, nsrc), np.int32)
    nrc1 = np.zeros((kmax, nsrc), np.int32)
    nsrc1 = np.zeros((kmax,), np.int32)
    obst, dist = [], []
    istep = 0
    istep1 = 0
    knumo = 12345
    knum = 0
    s_lat = s_lon = F32(0)
    with open(os.path.join(base, datafile)) as fh:
        for line in fh:
            if not line.strip():
                continue
            if line[0] == "#":  # gather header, main.f90:245-266
                tk = _tokens(line[1:])
                sta1_lat, sta1_lon = F32(tk[0]), F32(tk[1])
                period, wavetp, veltp = int(tk[2]), int(tk[3]), int(tk[4])
                if wavetp == 2 and veltp == 0:
                    knum = period
                if wavetp == 2 and veltp == 1:
                    knum = kmaxRc + period
                if wavetp == 1 and veltp == 0:
                    knum = kmaxRg + kmaxRc + period
                if wavetp == 1 and veltp == 1:
                    knum = kmaxLc + kmaxRg + kmaxRc + period
                if knum != knumo:
                    istep = 0
                istep += 1
                istep1 = 0
                s_lat = (F32(90.0) - sta1_lat) * PI32 / F32(180.0)
                s_lon = sta1_lon * PI32 / F32(180.0)
                scxf[knum - 1, istep - 1] = s_lat
                sczf[knum - 1, istep - 1] = s_lon
                periods[knum - 1, istep - 1] = period
                wavetype[knum - 1, istep - 1] = wavetp
                igrt[knum - 1, istep - 1] = veltp
                nsrc1[knum - 1] = istep
                knumo = knum
            else:  # receiver line, main.f90:267-279
                tk = _tokens(line)
                sta2_lat, sta2_lon, velvalue = F32(tk[0]), F32(tk[1]), F32(tk[2])
                istep1 += 1
                r_lat = (F32(90.0) - sta2_lat) * PI32 / F32(180.0)
                r_lon = sta2_lon * PI32 / F32(180.0)
                rcxf[knum - 1, istep - 1, istep1 - 1] = r_lat
                rczf[knum - 1, istep - 1, istep1 - 1] = r_lon
                d1 = F32(delsph(s_lat, s_lon, r_lat, r_lon))
                dist.append(d1)
                obst.append(F32(d1 / velvalue))
                nrc1[knum - 1, istep - 1] = istep1
    with open(os.path.join(base, "MOD")) as fh:
        vals = np.array(fh.read().split(), dtype=np.float64)
    depz = vals[:nz].astype(F32)
    vsf = vals[nz : nz + nx * ny * nz].astype(F32).reshape(nz, ny, nx)
    return Problem(
        nx, ny, nz, goxd, gozd, dvxd, dvzd, nsrc, weight, damp, minthk, minvel, maxvel, maxiter,
        spfra, kmaxRc, kmaxRg, kmaxLc, kmaxLg, tRc, tRg, tLc, tLg, ifsyn, noiselevel, threshold,
        scxf, sczf, rcxf, rczf, periods, wavetype, igrt, nrc1, nsrc1,
        np.array(obst, F32), np.array(dist, F32), depz, vsf,
        name=os.path.basename(base),
    )


def _isnum(tok: str) -> bool:
    try:
        float(tok)
        return True
    except ValueError:
        return False


TAIPEI_DEPZ = (0.0, 0.2, 0.4, 0.6, 0.8, 1.1, 1.4, 1.8, 2.5)


def synthetic_problem(
    nxy: int,
    nperiods: int,
    sources_per_period: int,
    types=("Rc",),
    nrecv: int = 16,
    seed: int = 20150131,
    perturb: bool = True,
    name: str | None = None,
) -> Problem:
    """Seeded synthetic configuration.

    nxy            model nodes per side (nx = ny); propagation grid = (nxy-3)*8+1 per side
    nperiods       periods per data type, 0.5 + 0.2 k seconds (same list for every type)
    types          subset of ("Rc", "Rg", "Lc", "Lg") in the reference's block order
    The *current* model (``vsf``) is the laterally uniform start model unless ``perturb``; the
    observations are straight-ray times through a smooth pseudo-velocity so that the residual
    vector is non-trivial without needing a forward solve.
    """
    rng = np.random.default_rng(seed)
    nx = ny = nxy
    nz = len(TAIPEI_DEPZ)
    depz = np.array(TAIPEI_DEPZ, F32)
    dvxd = dvzd = 0.015
    goxd, gozd = 26.5, 120.0
    kk = {t: (nperiods if t in types else 0) for t in ("Rc", "Rg", "Lc", "Lg")}
    tper = 0.5 + 0.2 * np.arange(nperiods, dtype=np.float64)
    kmax = sum(kk.values())
    nst = sources_per_period + 1
    nrecv = min(nrecv, nst - 1)
    nsrc = max(sources_per_period, nrecv)
    # stations in the central 80 % of the propagation box
    nvx, nvz = nx - 2, ny - 2
    lat_hi, lat_lo = goxd, goxd - (nvx - 1) * dvxd
    lon_lo, lon_hi = gozd, gozd + (nvz - 1) * dvzd
    lat = lat_lo + (0.1 + 0.8 * rng.random(nst)) * (lat_hi - lat_lo)
    lon = lon_lo + (0.1 + 0.8 * rng.random(nst)) * (lon_hi - lon_lo)
    lat = lat.astype(F32)
    lon = lon.astype(F32)
    cx = ((F32(90.0) - lat) * PI32 / F32(180.0)).astype(F32)
    cz = (lon * PI32 / F32(180.0)).astype(F32)

    scxf = np.zeros((kmax, nsrc), F32)
    sczf = np.zeros((kmax, nsrc), F32)
    rcxf = np.zeros((kmax, nsrc, nsrc), F32)
    rczf = np.zeros((kmax, nsrc, nsrc), F32)
    periods = np.zeros((kmax, nsrc), np.int32)
    wavetype = np.zeros((kmax, nsrc), np.int32)
    igrt = np.zeros((kmax, nsrc), np.int32)
    nrc1 = np.zeros((kmax, nsrc), np.int32)
    nsrc1 = np.zeros((kmax,), np.int32)
    obst, dist = [], []
    knum = 0
    for tname, (wt, gr) in (("Rc", (2, 0)), ("Rg", (2, 1)), ("Lc", (1, 0)), ("Lg", (1, 1))):
        for p in range(kk[tname]):
            for g in range(sources_per_period):
                scxf[knum, g] = cx[g]
                sczf[knum, g] = cz[g]
                periods[knum, g] = p + 1
                wavetype[knum, g] = wt
                igrt[knum, g] = gr
                ridx = (g + 1 + np.arange(nrecv)) % nst
                rcxf[knum, g, :nrecv] = cx[ridx]
                rczf[knum, g, :nrecv] = cz[ridx]
                nrc1[knum, g] = nrecv
                d = delsph(cx[g], cz[g], cx[ridx], cz[ridx])
                # smooth pseudo-velocity: grows with period, mild lateral variation
                vmid = (
                    0.85 + 0.12 * tper[p] + (0.1 if wt == 1 else 0.0) - (0.08 if gr == 1 else 0.0)
                    + 0.03 * np.sin(40.0 * (lat[g] + lat[ridx])) * np.cos(40.0 * (lon[g] + lon[ridx]))
                )
                dist.append(d)
                obst.append((d / vmid.astype(F32)).astype(F32))
            nsrc1[knum] = sources_per_period
            knum += 1
    z = depz.astype(np.float64)
    vs = (0.9 + 0.6 * z)[:, None, None] * np.ones((nz, ny, nx))
    if perturb:
        ii = np.arange(1, nx + 1)[None, None, :]
        jj = np.arange(1, ny + 1)[None, :, None]
        vs = vs + 0.15 * np.sin(0.5 * ii) * np.sin(0.5 * jj)
    vs = np.clip(vs, 0.5, 2.8).astype(F32)
    pb = Problem(
        nx, ny, nz, goxd, gozd, dvxd, dvzd, nsrc, 4.0, 1.0, 3.0, 0.5, 2.8, 1, 0.2,
        kk["Rc"], kk["Rg"], kk["Lc"], kk["Lg"],
        tper[: kk["Rc"]].copy(), tper[: kk["Rg"]].copy(), tper[: kk["Lc"]].copy(), tper[: kk["Lg"]].copy(),
        0, 0.0, 3.0, scxf, sczf, rcxf, rczf, periods, wavetype, igrt, nrc1, nsrc1,
        np.concatenate(obst).astype(F32), np.concatenate(dist).astype(F32), depz, vs,
        name=name or f"synthetic_{(nxy - 3) * 8 + 1}sq_{nperiods}p_{sources_per_period}s_{'+'.join(types)}",
    )
    pb.meta = dict(seed=seed, nrecv=nrecv, stations=nst, propagation_grid=(nxy - 3) * 8 + 1)
    return pb


def config(n: int) -> Problem:
    """BASELINE.json configs by number (interpretation A of SURVEY.md section 8: the quoted
    'W x W grid' is the FMM propagation grid)."""
    if n == 1:
        here = os.path.dirname(os.path.abspath(__file__))
        return read_problem(os.path.join(here, "..", "tests", "golden", "taipei", "DSurfTomo.in"))
    if n == 2:
        return synthetic_problem(35, 8, 64, ("Rc",), name="cfg2_257sq_8p_64s_Rc")
    if n == 3:
        return synthetic_problem(131, 16, 256, ("Rc", "Rg", "Lc", "Lg"), name="cfg3_1025sq_16p_256s_RcRgLcLg")
    if n == 5:
        return synthetic_problem(259, 32, 1024, ("Rc",), name="cfg5_2049sq_32p_1024s_Rc")
    raise ValueError(n)


def synthetic_dispersion(pb: Problem, seed: int = 7):
    """Deterministic synthetic dispersion results with the shapes CalSurfG's K1 stage produces,
    for benchmarks that time the sweep stage alone (both bench arms consume exactly these arrays):
    pv4   = [pvRc[kmax, ncol], pvRg[kmaxRg, ncol], pvLc[kmax, ncol], pvLg[kmaxLg, ncol]]
    sen12 = 12 arrays [nz, kmax_t, ncol] ordered Rc(vs, vp, rho), Rg(...), Lc(...), Lg(...).
    Velocities grow with period and vary smoothly (+-12 %) laterally; kernels are smooth,
    depth-localised bumps of realistic magnitude (dc/dVs ~ 0.1-0.5, dc/dVp, dc/drho ~ +-0.05)."""
    ncol = pb.nx * pb.ny
    ii = np.arange(pb.nx)[None, :]
    jj = np.arange(pb.ny)[:, None]
    lat = (np.sin(0.21 * ii + 0.3) * np.cos(0.17 * jj) + 0.5 * np.sin(0.05 * ii * jj / max(pb.nx, 1) + seed)).ravel()
    depth = np.asarray(pb.depz, np.float64)
    kk = (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)
    tt = (pb.tRc, pb.tRg, pb.tLc, pb.tLg)
    base = (0.0, -0.08, 0.10, 0.02)
    pv4, sen12 = [], []
    for t in range(4):
        cols = pb.kmax if t in (0, 2) else max(kk[t], 1)
        pv = np.zeros((cols, ncol), np.float64)
        for c in range(cols):
            per = tt[t][c] if c < kk[t] else (0.5 + 0.2 * (c % max(1, max(kk))))
            pv[c] = np.float32(0.85 + 0.12 * per + base[t]) * (1.0 + 0.12 * lat * (0.6 + 0.4 * np.cos(0.9 * c)))
        pv4.append(pv.astype(np.float32).astype(np.float64))  # values are REAL*4-representable
        if kk[t] == 0:
            sen12 += [None, None, None]
            continue
        zc = 0.25 * np.asarray(tt[t], np.float64)[None, :, None] + 0.1  # sensitivity deepens with period
        bump = np.exp(-(((depth[:, None, None] - zc) / (0.35 + 0.2 * zc)) ** 2))
        mod = (1.0 + 0.2 * lat)[None, None, :]
        sen12.append(0.45 * bump * mod)
        sen12.append(0.05 * bump * (1.0 - 0.3 * lat)[None, None, :])
        sen12.append(-0.04 * bump * mod + 0.01)
    return pv4, sen12
