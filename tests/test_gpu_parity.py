"""GPU parity tests: the CUDA path (through the C ABI of libdsurf_b200.so) against the CPU oracle
on identical inputs.  Bit-exact where the arithmetic is fp32 in reference order (velocity dicing,
travel times, ray cells / Frechet sums, column patterns); tolerance-checked where the reference
itself is only defined up to libm / summation order (fp64 root search, LSMR)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import ROOT
from dsurftomo_b200 import api, hostglue, inputs

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


def test_library_reports_build():
    from dsurftomo_b200 import _lib

    assert b"sm_100a" in _lib.lib().dsurf_build_info()


# ------------------------------------------------------------------ K1 dispersion
def _random_stacks(rng, nmodel, nlayer):
    vs = np.sort(rng.uniform(0.6, 3.8, (nmodel, nlayer)), axis=1).astype(np.float32)
    vs[:, 1:3] = vs[:, 1:3][:, ::-1]  # a mild low-velocity zone
    vp = (vs * rng.uniform(1.65, 1.9, (nmodel, 1))).astype(np.float32)
    rho = (1.6 + 0.3 * vs).astype(np.float32)
    thk = np.concatenate([rng.uniform(0.3, 2.0, nlayer - 1), [0.0]]).astype(np.float32)
    return thk, vp, vs, rho


@pytest.mark.parametrize("iwave,igr,iflsph", [(2, 0, 1), (2, 1, 1), (1, 0, 1), (1, 1, 1), (2, 0, 0)])
def test_surfdisp96_batch_matches_oracle(iwave, igr, iflsph):
    rng = np.random.default_rng(7 + iwave * 10 + igr)
    thk, vp, vs, rho = _random_stacks(rng, 96, 12)
    t = np.array([0.5, 0.8, 1.2, 2.0, 3.5, 6.0])
    got = api.surfdisp96_batch(thk, vp, vs, rho, iflsph, iwave, igr, t)
    ref = np.stack([O.surfdisp96(thk, vp[i], vs[i], rho[i], iflsph, iwave, 1, igr, t)[0] for i in range(len(vp))])
    # the root search is fp64 with libm-dependent sin/cos/exp: equal after rounding to fp32 except for
    # rare 1-ulp flips (phase); group velocity differentiates numerically and amplifies them
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)
    tol = 2e-7 if igr == 0 else 2e-4
    assert rel.max() <= tol, rel.max()
    if igr == 0:
        assert (got != ref).mean() < 0.02


def test_surfdisp96_single_and_higher_mode():
    rng = np.random.default_rng(3)
    thk, vp, vs, rho = _random_stacks(rng, 1, 10)
    t = np.array([0.4, 0.6, 0.9, 1.3])
    for mode in (1, 2):
        got = api.surfdisp96(thk, vp[0], vs[0], rho[0], 10, 1, 2, mode, 0, len(t), t)
        ref, _ = O.surfdisp96(thk, vp[0], vs[0], rho[0], 1, 2, mode, 0, t)
        np.testing.assert_allclose(got, ref, rtol=3e-7, atol=0)


def _ulp32(x):
    """Spacing of REAL*4 numbers at |x| (the reference rounds every dispersion value to REAL*4, surfdisp96.f:292-297)."""
    return np.spacing(np.abs(x).astype(np.float32)).astype(np.float64)


def test_depthkernel_matches_oracle(small_problem):
    """Phase velocities may differ from the oracle by one REAL*4 ulp in a few per cent of the values (CUDA's and
    glibc's fp64 sin/cos/exp differ in the last bit, which occasionally moves the REAL*4 rounding of a root).  The
    finite-difference kernels (c+ - c-)/(0.01 p) (CalSurfG.f90:76-160) inherit exactly that: every kernel entry must
    equal the oracle's up to a whole number of ulp flips of its two dispersion values (<= 2 ulps in total), and only
    a few per cent of the entries may differ at all.  Group velocities are a REAL*4 difference quotient over a 1 %
    period step (surfdisp96.f:281-300): one ulp of a root moves U by ~200 ulps, so they are budgeted in ulps of U."""
    pb = small_problem
    t = np.array([0.6, 1.0, 1.6])
    nz = pb.nz
    vs = pb.vsf.reshape(nz, -1).astype(np.float64)                       # [nz][ncol]
    vp = 0.9409 + 2.0947 * vs - 0.8206 * vs ** 2 + 0.2683 * vs ** 3 - 0.0251 * vs ** 4   # Brocher (CalSurfG.f90:49-53)
    rho = 1.6612 * vp - 0.4721 * vp ** 2 + 0.0671 * vp ** 3 - 0.0043 * vp ** 4 + 0.000106 * vp ** 5
    for iwave, igr in ((2, 0), (1, 0), (2, 1), (1, 1)):
        pv, svs, svp, srho = api.depthkernel(pb.nx, pb.ny, pb.nz, pb.vsf, iwave, igr, len(t), t, pb.depz, pb.minthk)
        rpv, rvs, rvp, rrho = O.depthkernel(pb.vsf, iwave, igr, t, pb.depz, pb.minthk, nthreads=8)
        d_ulps = np.abs(pv - rpv) / _ulp32(rpv)
        if igr == 0:
            assert d_ulps.max() <= 1.0 and (d_ulps > 0).mean() <= 0.03, (d_ulps.max(), (d_ulps > 0).mean())
        else:
            assert d_ulps.max() <= 600 and np.median(d_ulps) <= 8, (d_ulps.max(), np.median(d_ulps))
        ulp_c = _ulp32(rpv)[None, :, :]                                   # [1][k][ncol]
        budget = (2.0 if igr == 0 else 1200.0)
        for a, b, par in ((svs, rvs, vs), (svp, rvp, vp), (srho, rrho, rho)):
            flips = np.abs(a - b) * (0.01 * par[:, None, :]) / ulp_c      # difference in ulps of the dispersion values
            assert flips.max() <= budget * 1.001, (iwave, igr, flips.max())
            if igr == 0:
                assert (flips > 0).mean() <= 0.06, (flips > 0).mean()
                assert np.abs(flips - np.round(flips)).max() <= 0.02     # whole ulp flips, nothing else
    pv2 = api.caldespersion(pb.nx, pb.ny, pb.nz, pb.vsf, 2, 0, len(t), t, pb.depz, pb.minthk)
    rpv2 = O.caldespersion(pb.vsf, 2, 0, t, pb.depz, pb.minthk, nthreads=8)
    assert (np.abs(pv2 - rpv2) / _ulp32(rpv2)).max() <= 1.0


# ------------------------------------------------------------------ K2-K5 eikonal + rays
def _hetero_map(pb, seed=5):
    rng = np.random.default_rng(seed)
    ii = np.arange(pb.nx)[None, :]
    jj = np.arange(pb.ny)[:, None]
    v = 1.2 + 0.35 * np.sin(0.7 * ii) * np.cos(0.5 * jj) + 0.05 * rng.standard_normal((pb.ny, pb.nx))
    return v.ravel().astype(np.float64)


@pytest.mark.parametrize("which", ["uniform", "hetero"])
def test_sweep_bit_exact(taipei, which):
    pb = taipei
    plan = api.Plan(pb)
    pv = np.full(pb.nx * pb.ny, 1.3) if which == "uniform" else _hetero_map(pb)
    for per in range(pb.kmaxRc):
        plan.set_map(0, per, pv)
    gathers = [0, 7, 100, 448]
    for g in gathers:
        # locate (knumi, srcnum) of the flattened gather
        k = int(np.searchsorted(np.cumsum(pb.nsrc1), g, side="right"))
        s = g - int(np.concatenate([[0], np.cumsum(pb.nsrc1)])[k])
        got = plan.debug_sweep(g, 1)
        ref = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, pb.scxf[k, s], pb.sczf[k, s])
        assert ref["err"] == 0
        assert np.array_equal(_bits(got["veln"]), _bits(ref["veln"]))
        assert np.array_equal(got["nstsr"] == 0, ref["nstsr"] == 0)
        alive = ref["nstsr"] == 0
        assert np.array_equal(_bits(got["ttnr"])[alive], _bits(ref["ttnr"])[alive])
        assert np.array_equal(_bits(got["ttn"]), _bits(ref["ttn"])), np.abs(got["ttn"] - ref["ttn"]).max()
        nrc = int(pb.nrc1[k, s])
        err, tt, fdm = O.sweep_rays(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, pb.scxf[k, s],
                                    pb.sczf[k, s], pb.rcxf[k, s, :nrc], pb.rczf[k, s, :nrc])
        assert err == 0
        gf = got["fdm"][:nrc]
        # identical vertex pattern (ray cells) and identical fp32 sums
        assert np.array_equal(gf != 0, fdm != 0)
        assert np.array_equal(_bits(gf), _bits(fdm)), np.abs(gf - fdm).max()
    plan.close()


# ------------------------------------------------------------------ CalSurfG end to end
def _compare_calsurfg(pb, got, ref, disp_exact_frac=0.9):
    assert got["nar"] > 0
    # predicted times: fp32-identical velocity maps give identical times; a 1-ulp flip of a
    # dispersion value perturbs them at ~1e-7
    np.testing.assert_allclose(got["dsurf"], ref["dsurf"], rtol=1e-5, atol=0)
    same_pattern = got["nar"] == ref["nar"] and np.array_equal(got["row"], ref["row"]) and \
        np.array_equal(got["col"], ref["col"])
    if not same_pattern:
        # threshold crossings of |row| > 1e-4 may differ where kernels differ by an ulp: bound them
        a = set(zip(got["row"].tolist(), got["col"].tolist()))
        b = set(zip(ref["row"].tolist(), ref["col"].tolist()))
        assert len(a ^ b) <= 1e-3 * len(b), (len(a ^ b), len(b))
    else:
        scale = np.abs(ref["rw"]).max()
        assert np.abs(got["rw"] - ref["rw"]).max() <= 5e-3 * scale
        assert (np.abs(got["rw"] - ref["rw"]) <= 1e-5 * scale).mean() > disp_exact_frac
    return same_pattern


def test_calsurfg_small_all_types(small_problem):
    pb = small_problem
    got = api.CalSurfG(pb)
    ref = O.calsurfg(pb, nthreads=8, mode=1)
    assert ref["err"] == 0
    _compare_calsurfg(pb, got, ref, disp_exact_frac=0.5)


def test_calsurfg_taipei(taipei):
    pb = taipei
    got = api.CalSurfG(pb)
    ref = O.calsurfg(pb, nthreads=8, mode=1)
    assert ref["err"] == 0
    same = _compare_calsurfg(pb, got, ref)
    assert same, "Taipei sparsity pattern differs from the oracle"
    assert np.array_equal(_bits(got["dsurf"]), _bits(ref["dsurf"])) or \
        np.abs(got["dsurf"] / ref["dsurf"] - 1).max() < 1e-6


def test_source_outside_is_reported(small_problem):
    import copy

    pb = copy.deepcopy(small_problem)
    pb.scxf = pb.scxf.copy()
    pb.scxf[0, 0] = 0.1  # far outside the grid
    with pytest.raises(api.DsurfError) as e:
        api.CalSurfG(pb)
    assert e.value.code == 1


# ------------------------------------------------------------------ K7 LSMR / aprod
def _taipei_system(pb):
    ref = O.calsurfg(pb, nthreads=8, mode=1)
    sysd = hostglue.host_glue(pb, ref["dsurf"], ref["row"], ref["col"], ref["rw"])
    return sysd


def test_aprod_matches_oracle(taipei):
    s = _taipei_system(taipei)
    iw = hostglue.pack_iw(s["rows"], s["cols"])
    rng = np.random.default_rng(0)
    x = rng.standard_normal(s["n"]).astype(np.float32)
    y = rng.standard_normal(s["m"]).astype(np.float32)
    for mode in (1, 2):
        gx, gy = api.aprod(mode, s["m"], s["n"], x, y, len(iw), len(s["vals"]), iw, s["vals"])
        rx, ry = O.aprod(mode, s["m"], s["n"], x.copy(), y.copy(), iw, s["vals"])
        a, b = (gy, ry) if mode == 1 else (gx, rx)
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()


def test_lsmr_matches_oracle(taipei):
    pb = taipei
    s = _taipei_system(pb)
    iw = hostglue.pack_iw(s["rows"], s["cols"])
    got = api.LSMR(s["m"], s["n"], len(iw), len(s["vals"]), iw, s["vals"], s["cbst"], pb.damp, 1e-6, 1e-6, 100.0,
                   400, 10)
    ref = O.lsmr(s["m"], s["n"], iw, s["vals"], s["cbst"], pb.damp)
    assert abs(got["itn"] - ref["itn"]) <= 2, (got["itn"], ref["itn"])
    assert got["istop"] == ref["istop"]
    # Vs model within 1e-5 relative: dv is added to Vs ~ 1 km/s
    assert np.abs(got["x"] - ref["x"]).max() <= 1e-5
    for k in ("normr", "normx"):
        assert abs(got[k] - ref[k]) <= 1e-4 * abs(ref[k])
    # normA / condA are running estimates that grow by one (alpha, beta) pair per iteration: only
    # comparable at equal iteration counts (the fp32 stopping test may fire one iteration apart)
    if got["itn"] == ref["itn"]:
        for k in ("normA", "condA"):
            assert abs(got[k] - ref[k]) <= 1e-4 * abs(ref[k])
    else:
        assert abs(got["normA"] - ref["normA"]) <= 0.03 * ref["normA"]


def test_lsmr_random_system_vs_dense():
    rng = np.random.default_rng(11)
    m, n, nnz = 300, 40, 2500
    rows = np.sort(rng.integers(1, m + 1, nnz)).astype(np.int32)
    cols = rng.integers(1, n + 1, nnz).astype(np.int32)
    vals = rng.standard_normal(nnz).astype(np.float32)
    b = rng.standard_normal(m).astype(np.float32)
    iw = hostglue.pack_iw(rows, cols)
    damp = 0.3
    got = api.LSMR(m, n, len(iw), nnz, iw, vals, b, damp, 1e-7, 1e-7, 1e8, 400, 10)
    A = np.zeros((m, n))
    np.add.at(A, (rows - 1, cols - 1), vals.astype(np.float64))
    xs = np.linalg.solve(A.T @ A + damp ** 2 * np.eye(n), A.T @ b.astype(np.float64))
    assert np.abs(got["x"] - xs).max() <= 2e-4 * np.abs(xs).max()
    # unsorted triplets are accepted like the reference's aprod
    perm = rng.permutation(nnz)
    iw2 = hostglue.pack_iw(rows[perm], cols[perm])
    got2 = api.LSMR(m, n, len(iw2), nnz, iw2, vals[perm], b, damp, 1e-7, 1e-7, 1e8, 400, 10)
    assert np.abs(got2["x"] - xs).max() <= 2e-4 * np.abs(xs).max()


def _outer_iteration_vs_oracle(pb, nthreads=8):
    """One full outer iteration (main.f90:348-546 without the file output): CalSurfG -> residuals, outlier weights,
    smoothing rows -> LSMR -> model update, on the GPU through the C ABI and on the oracle chain."""
    got = api.CalSurfG(pb)
    s = hostglue.host_glue(pb, got["dsurf"], got["row"], got["col"], got["rw"])
    iw = hostglue.pack_iw(s["rows"], s["cols"])
    sol = api.LSMR(s["m"], s["n"], len(iw), len(s["vals"]), iw, s["vals"], s["cbst"], pb.damp, 1e-6, 1e-6, 100.0,
                   400, 10)
    vs_new, _ = hostglue.model_update(pb, pb.vsf, sol["x"])
    ref = O.calsurfg(pb, nthreads=nthreads, mode=1)
    assert ref["err"] == 0
    iwr, rwr, colr = ref["iw"].copy(), ref["rw_full"].copy(), ref["col_full"].copy()
    m, nar, cbst, _ = O.host_glue(pb, ref["dsurf"], iwr, rwr, colr, ref["nar"])
    L = O.lsmr(m, pb.maxvp, iwr[: 2 * nar + 1], rwr[:nar], cbst[:m], pb.damp)
    vs_ref, _ = O.model_update(pb, pb.vsf, L["x"])
    assert m == s["m"]
    a = set(zip(got["row"].tolist(), got["col"].tolist()))
    b = set(zip(ref["row"].tolist(), ref["col"].tolist()))
    return dict(got=got, ref=ref, vs_new=vs_new, vs_ref=vs_ref, itn=(sol["itn"], L["itn"]),
                pattern_mismatch=len(a ^ b), pattern_total=len(b),
                dsurf_rel=float(np.abs(got["dsurf"] / ref["dsurf"] - 1).max()),
                vs_rel=float(np.abs(vs_new / vs_ref - 1).max()))


def test_outer_iteration_model_update(taipei):
    """Taipei (Rayleigh phase): Vs model after one outer iteration within 1e-5 relative (north_star)."""
    r = _outer_iteration_vs_oracle(taipei)
    assert r["vs_rel"] <= 1e-5, r["vs_rel"]


def test_outer_iteration_all_four_data_types(small_problem):
    """Rayleigh + Love, phase + group (Rc, Rg, Lc, Lg) through CalSurfG -> glue -> LSMR -> update: the north_star
    tolerances hold for every data type, not only for Rayleigh phase.  Pattern mismatches (entries whose |row| sits
    within an ulp of the 1e-4 threshold, CalSurfG.f90:1425) are counted, not assumed zero."""
    r = _outer_iteration_vs_oracle(small_problem)
    print("4 types:", {k: r[k] for k in ("dsurf_rel", "vs_rel", "pattern_mismatch", "pattern_total", "itn")})
    assert r["dsurf_rel"] <= 1e-5
    assert r["pattern_mismatch"] <= 2e-4 * r["pattern_total"], (r["pattern_mismatch"], r["pattern_total"])
    assert abs(r["itn"][0] - r["itn"][1]) <= 2
    assert r["vs_rel"] <= 1e-5, r["vs_rel"]


def test_cfg2_whole_configuration_vs_oracle():
    """BASELINE configs[1] run whole (257 x 257 propagation grid, 8 periods x 64 sources, Rayleigh phase, 512 sweeps,
    8192 rays): predicted times within 1e-5 relative, sparsity-pattern mismatches counted, and the Vs model after one
    outer iteration within 1e-5 relative."""
    pb = inputs.config(2)
    r = _outer_iteration_vs_oracle(pb, nthreads=os.cpu_count() or 8)
    print("cfg 2:", {k: r[k] for k in ("dsurf_rel", "vs_rel", "pattern_mismatch", "pattern_total", "itn")})
    assert r["dsurf_rel"] <= 1e-5
    assert r["pattern_mismatch"] <= 2e-4 * r["pattern_total"], (r["pattern_mismatch"], r["pattern_total"])
    assert abs(r["itn"][0] - r["itn"][1]) <= 2
    assert r["vs_rel"] <= 1e-5, r["vs_rel"]


def test_synthetic_matches_oracle(taipei, tmp_path):
    """subroutine synthetic (forward times on the gd = 5 grid, ifsyn = 1).
    (a) through the drop-in entry with the library's own dispersion maps: <=1e-5 relative (the maps
    differ by 1 ulp in a few columns, see test_surfdisp96_batch_matches_oracle); (b) the forward-only
    plan fed with the oracle's maps: bit-exact times; (c) velmap2dRc.dat format and content."""
    pb = taipei
    vtrue = (pb.vsf * (1.0 + 0.05 * np.sin(np.arange(pb.nx))[None, None, :])).astype(np.float32)
    ref = O.synthetic(pb, vels=vtrue, nthreads=8)
    assert ref["err"] == 0
    got = api.synthetic(pb, vels=vtrue, outdir=str(tmp_path))
    assert np.abs(got["obst"] / ref["obst"] - 1).max() <= 1e-5
    plan = api.Plan(pb, vels=vtrue, forward=True)
    plan.set_dispersion(0, ref["pv"][: pb.kmaxRc], None, None, None)
    plan.finalize_dispersion()
    plan.reset_rows()
    plan.sweeps()
    assert plan.num_sweeps() == plan.num_gathers
    exact = plan.download()
    assert exact["nar"] == 0 and np.array_equal(exact["dsurf"], ref["obst"])
    lines = open(tmp_path / "velmap2dRc.dat").read().splitlines()
    assert len(lines) == pb.kmaxRc * (pb.nx - 2) * (pb.ny - 2) and all(len(l) == 32 for l in lines[:50])
    first = [float(lines[0][8 * i:8 * i + 8]) for i in range(4)]
    assert abs(first[0] - pb.gozd) < 1e-4 and abs(first[1] - pb.goxd) < 1e-4 and abs(first[2] - pb.tRc[0]) < 1e-4
    assert abs(first[3] - ref["pv"][0][2 * pb.nx + 1]) < 1e-4
    # noise: reproducible for a seed, zero-mean relative perturbation of the requested size
    n1 = api.synthetic(pb, vels=vtrue, noiselevel=0.02, seed=7)["obst"]
    n2 = api.synthetic(pb, vels=vtrue, noiselevel=0.02, seed=7)["obst"]
    rel = n1 / got["obst"] - 1
    assert np.array_equal(n1, n2) and abs(rel.mean()) < 3e-3 and 0.015 < rel.std() < 0.025


def test_device_glue_matches_host_glue(taipei):
    """residual / percentile weights / row scaling / smoothing rows / model update
    on the device (LsmrSystem.from_plan, Plan.update_model) against the host mirror of
    main.f90:361-466,518-532 fed with the same CalSurfG output: identical system, identical solve."""
    pb = taipei
    plan = api.Plan(pb)
    plan.dispersion()
    plan.reset_rows()
    plan.sweeps()
    raw = plan.download()
    raw = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in raw.items()}
    s = hostglue.host_glue(pb, raw["dsurf"], raw["row"], raw["col"], raw["rw"])
    sysd = api.LsmrSystem.from_plan(plan)
    g = plan.glue_results()
    assert g["m"] == s["m"] and g["nar"] == len(s["vals"])
    assert np.array_equal(g["cbst"], s["cbst"][: pb.dall])
    assert np.array_equal(g["datweight"], s["datweight"])
    q25, q75 = hostglue.getpercentile((pb.obst - raw["dsurf"]).astype(np.float32))
    assert g["q25"] == q25 and g["q75"] == q75
    norm = np.zeros(pb.maxvp, np.float64)
    np.add.at(norm, s["cols"][: raw["nar"]] - 1, np.abs(s["vals"][: raw["nar"]]).astype(np.float64))
    assert abs(g["maxnorm"] / norm.max() - 1) < 1e-6 and abs(g["averdws"] / norm.mean() - 1) < 1e-6
    scaled = plan.download()  # the plan's rw is scaled in place like the reference's
    assert np.array_equal(scaled["rw"], s["vals"][: raw["nar"]])
    dev = sysd.solve(pb.damp)
    sysh = api.LsmrSystem(s["m"], s["n"], s["rows"], s["cols"], s["vals"], s["cbst"])
    host = sysh.solve(pb.damp)
    assert dev["itn"] == host["itn"] and dev["istop"] == host["istop"]
    assert np.array_equal(dev["x"], host["x"])
    vs_dev, dv_dev = plan.update_model(sysd)
    vs_host, dv_host = hostglue.model_update(pb, pb.vsf, host["x"])
    assert np.array_equal(dv_dev, dv_host) and np.array_equal(vs_dev, vs_host)
    # second outer iteration runs from the device-resident model
    plan.dispersion()
    plan.reset_rows()
    plan.sweeps()
    second = plan.download()
    plan2 = api.Plan(pb, vels=vs_host)
    plan2.dispersion()
    plan2.reset_rows()
    plan2.sweeps()
    ref2 = plan2.download()
    assert second["nar"] == ref2["nar"] and np.array_equal(second["dsurf"], ref2["dsurf"])
    assert np.array_equal(second["rw"], ref2["rw"])


@pytest.mark.parametrize("var", ["DSURF_EIKONAL_LPS", "DSURF_EIKONAL_LAZY", "DSURF_MAXSLOTS=3,DSURF_MAXRAYS=20",
                                 "DSURF_EIKONAL_LPS,DSURF_MAXSLOTS=40,DSURF_MAXRAYS=20", "DSURF_EIKONAL_LPS,DSURF_HCAP=600"])
def test_reference_eikonal_variants_also_bit_exact(var):
    """The lane-per-sweep pipeline (k_refine + warp-specialised k_march_lps on one word per node with
    lazy heap back-pointers, DSURF_EIKONAL_LPS) and the lazy-back-pointer variant of the default
    16-lanes-per-sweep march must reproduce the oracle exactly like the default; so must both
    pipelines when sweeps are forced into several batches, rays into several chunks, and the heap
    slab through its growth path.  Variants are selected by environment variables at process start."""
    import os
    import subprocess
    import sys

    env = dict(os.environ)
    for kv in var.split(","):
        k, _, v = kv.partition("=")
        env[k] = v or "1"
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", "sweep_bit_exact or calsurfg_small or full_size_grid"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_lsmr_blocked_layout_matches_scalar_layout(taipei):
    """The depth-blocked (1 index + 8 values per vertex) layout and the scalar CSR/CSC layout are
    two storage schemes of the same operator: identical solution up to fp32 round-off."""
    pb = taipei
    s = _taipei_system(pb)
    iw = hostglue.pack_iw(s["rows"], s["cols"])
    args = (s["m"], s["n"], len(iw), len(s["vals"]), iw, s["vals"], s["cbst"], pb.damp, 1e-6, 1e-6, 100.0, 400, 10)
    api.lsmr_hint_geometry(pb.nx, pb.ny, pb.nz)      # n == P*K -> blocked
    blk = api.LSMR(*args)
    api.lsmr_hint_geometry(3, 3, 2)                  # P*K = 1 != n -> scalar
    sca = api.LSMR(*args)
    ref = O.lsmr(s["m"], s["n"], iw, s["vals"], s["cbst"], pb.damp)
    assert abs(blk["itn"] - sca["itn"]) <= 1
    assert np.abs(blk["x"] - sca["x"]).max() <= 2e-6
    assert np.abs(blk["x"] - ref["x"]).max() <= 1e-5
    # aprod-like check of both products through a 1-iteration solve is implicit; also run a system
    # whose depth count is below 8 (padding path): nz-1 = 3
    rng = np.random.default_rng(5)
    P, K, m = 30, 3, 200
    nnz = 3000
    rows = np.sort(rng.integers(1, m + 1, nnz)).astype(np.int32)
    cols = rng.integers(1, P * K + 1, nnz).astype(np.int32)
    vals = rng.standard_normal(nnz).astype(np.float32)
    b = rng.standard_normal(m).astype(np.float32)
    iw2 = hostglue.pack_iw(rows, cols)
    api.lsmr_hint_geometry(8, 7, K + 1)              # (8-2)*(7-2) = 30 vertices, 3 depths
    g1 = api.LSMR(m, P * K, len(iw2), nnz, iw2, vals, b, 0.2, 1e-7, 1e-7, 1e8, 400, 10)
    api.lsmr_hint_geometry(3, 3, 2)
    g2 = api.LSMR(m, P * K, len(iw2), nnz, iw2, vals, b, 0.2, 1e-7, 1e-7, 1e8, 400, 10)
    A = np.zeros((m, P * K))
    np.add.at(A, (rows - 1, cols - 1), vals.astype(np.float64))
    xs = np.linalg.solve(A.T @ A + 0.04 * np.eye(P * K), A.T @ b.astype(np.float64))
    assert np.abs(g1["x"] - xs).max() <= 2e-4 * np.abs(xs).max()
    assert np.abs(g2["x"] - xs).max() <= 2e-4 * np.abs(xs).max()
    api.lsmr_hint_geometry(pb.nx, pb.ny, pb.nz)


def test_heap_slab_growth_path_is_exact():
    """A deliberately tiny narrow-band slab (DSURF_HCAP) must trigger the grow-and-retry path and
    still reproduce the oracle bit for bit."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, DSURF_HCAP="600")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", "sweep_bit_exact"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


# ------------------------------------------------------------------ BASELINE sizes (cfg 3 grid)
@pytest.fixture(scope="module")
def cfg3_slice():
    """The 1025 x 1025 propagation grid of BASELINE configs[2] with one period and a few gathers."""
    pb = inputs.synthetic_problem(131, 1, 6, ("Rc",), nrecv=5, name="cfg3_grid_slice")
    pv4, sen12 = inputs.synthetic_dispersion(pb)
    return pb, pv4, sen12


def test_full_size_grid_sweep_bit_exact(cfg3_slice):
    pb, pv4, sen12 = cfg3_slice
    plan = api.Plan(pb)
    plan.set_dispersion(0, pv4[0], *sen12[0:3])
    plan.finalize_dispersion()
    g = 3
    got = plan.debug_sweep(g, 1)
    ref = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv4[0][0], pb.scxf[0, g], pb.sczf[0, g])
    assert ref["err"] == 0 and got["ttn"].shape == (1025, 1025)
    assert np.array_equal(_bits(got["veln"]), _bits(ref["veln"]))
    assert np.array_equal(_bits(got["ttn"]), _bits(ref["ttn"]))
    nrc = int(pb.nrc1[0, g])
    err, tt, fdm = O.sweep_rays(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv4[0][0], pb.scxf[0, g],
                                pb.sczf[0, g], pb.rcxf[0, g, :nrc], pb.rczf[0, g, :nrc])
    assert err == 0
    assert np.array_equal(_bits(got["fdm"][:nrc]), _bits(fdm))
    # size-independent properties of the travel-time field
    t = got["ttn"]
    assert np.isfinite(t).all() and t.min() >= 0.0
    dx = np.abs(np.diff(t, axis=0)).max()
    dz = np.abs(np.diff(t, axis=1)).max()
    vmin = got["veln"].min()
    h = 6371.0 * np.deg2rad(pb.dvxd) / 8
    assert max(dx, dz) <= 1.5 * h / vmin  # |grad T| = 1/v: neighbours differ by at most ~h/v
    plan.close()


def test_full_size_rows_match_oracle_and_normal_equations(cfg3_slice):
    pb, pv4, sen12 = cfg3_slice
    plan = api.Plan(pb)
    plan.set_dispersion(0, pv4[0], *sen12[0:3])
    plan.finalize_dispersion()
    plan.reset_rows()
    plan.sweeps()
    got = plan.download()
    ref = O.calsurfg_pre(pb, pv4, sen12, nthreads=8, maxnar=4_000_000)
    assert ref["err"] == 0 and got["nar"] == ref["nar"]
    assert np.array_equal(got["row"], ref["row"]) and np.array_equal(got["col"], ref["col"])
    assert np.array_equal(_bits(got["rw"]), _bits(ref["rw"])) and np.array_equal(_bits(got["dsurf"]), _bits(ref["dsurf"]))
    # LSMR at n = 133 128 unknowns: the damped normal equations are satisfied
    s = hostglue.host_glue(pb, got["dsurf"], got["row"], got["col"], got["rw"])
    iw = hostglue.pack_iw(s["rows"], s["cols"])
    sol = api.LSMR(s["m"], s["n"], len(iw), len(s["vals"]), iw, s["vals"], s["cbst"], pb.damp, 1e-6, 1e-6, 100.0, 400, 10)
    import scipy.sparse as sp

    A = sp.coo_matrix((s["vals"].astype(np.float64), (s["rows"] - 1, s["cols"] - 1)), shape=(s["m"], s["n"])).tocsr()
    x = sol["x"].astype(np.float64)
    g = A.T @ (A @ x - s["cbst"].astype(np.float64)) + pb.damp ** 2 * x
    assert np.linalg.norm(g) <= 1e-4 * np.linalg.norm(A.T @ s["cbst"].astype(np.float64))
    plan.close()


def test_raypath_export_matches_oracle(taipei, tmp_path):
    """ray-path export in the reference's raypath.out format (CalSurfG.f90:2276-2283):
    '# nrp' + nrp 'latitude longitude' records per traced ray, receiver first, source last.  With the
    same velocity map the geometry is bit-identical to the oracle's rgx/rgz."""
    pb = taipei
    plan = api.Plan(pb)
    plan.dispersion()
    pv_all = plan.get_dispersion(0)[0]  # [period][nx*ny] maps the sweeps propagate through (Rc only)
    f = tmp_path / "raypath.out"
    g0, g1 = 3, 9
    plan.set_raypath(str(f))
    plan.reset_rows()
    plan.sweeps(g0, g1)
    plan.set_raypath(None)
    lines = open(f).read().splitlines()
    pi = np.float32(3.1415926535898)
    cum = np.concatenate([[0], np.cumsum(pb.nsrc1)])
    pos = 0
    nrays = 0
    for g in range(g0, g1):
        k = int(np.searchsorted(np.cumsum(pb.nsrc1), g, side="right"))
        s = g - int(cum[k])
        nrc = int(pb.nrc1[k, s])
        ref = O.sweep_paths(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv_all[pb.periods[k, s] - 1],
                            pb.scxf[k, s], pb.sczf[k, s], pb.rcxf[k, s, :nrc], pb.rczf[k, s, :nrc])
        for p in ref:
            hdr = lines[pos].split()
            assert hdr[0] == "#" and int(hdr[1]) == len(p) and len(lines[pos]) == 14  # ' #' + I12
            got = np.array([[np.float32(v) for v in l.split()] for l in lines[pos + 1: pos + 1 + len(p)]], np.float32)
            want_lat = (pi / np.float32(2) - p[:, 0]) * np.float32(180.0) / pi
            want_lon = p[:, 1] * np.float32(180.0) / pi
            assert np.array_equal(got[:, 0], want_lat) and np.array_equal(got[:, 1], want_lon)
            assert all(len(l) == 34 for l in lines[pos + 1: pos + 1 + len(p)])        # two list-directed REAL*4
            pos += 1 + len(p)
            nrays += 1
    assert pos == len(lines) and nrays > 0
    # the export does not disturb the rows: same COO as a run without it
    with_paths = plan.download()
    with_paths = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in with_paths.items()}
    plan.reset_rows()
    plan.sweeps(g0, g1)
    plain = plan.download()
    assert with_paths["nar"] == plain["nar"] and np.array_equal(with_paths["rw"], plain["rw"])
    plan.close()


def test_lsmr_fused_cluster_kernels_match_unfused(taipei):
    """The fused small-vector phases (one thread-block cluster: k_fused_beta / k_fused_tail, register path)
    against the unfused kernel sequence (DSURF_LSMR_NO_FUSE=1, separate process): same iteration count
    and stopping reason, solution equal to fp32 rounding of differently ordered fp64 partial sums."""
    import json
    import os
    import subprocess
    import sys

    code = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from dsurftomo_b200 import api, inputs, hostglue\n"
        "from dsurftomo_b200._lib import lib\n"
        "pb = inputs.config(1)\n"
        "plan = api.Plan(pb); plan.dispersion(); plan.reset_rows(); plan.sweeps()\n"
        "sysd = api.LsmrSystem.from_plan(plan)\n"
        "r = sysd.solve(pb.damp)\n"
        "print(json.dumps(dict(cl=int(lib().dsurf_lsmr_fused_cluster(sysd.h)), itn=r['itn'], istop=r['istop'],"
        " normr=r['normr'], x=r['x'].tolist())))\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    out = {}
    for name, extra in (("fused", {}), ("unfused", {"DSURF_LSMR_NO_FUSE": "1"})):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["fused"]["cl"] in (8, 16) and out["unfused"]["cl"] == 0
    assert out["fused"]["itn"] == out["unfused"]["itn"] and out["fused"]["istop"] == out["unfused"]["istop"]
    xf, xu = np.array(out["fused"]["x"]), np.array(out["unfused"]["x"])
    assert np.abs(xf - xu).max() <= 1e-6 * np.abs(xu).max()


@pytest.mark.parametrize("otf", ["8", "10", "12"])
def test_dispersion_on_the_fly_stacks_bit_identical(small_problem, otf):
    """Second-generation column kernel (k_disp_columns_otf: one base stack per column in shared memory, perturbed
    layers recomputed on the fly) against the first generation (one stack per thread): identical bits for
    every variant curve, Rayleigh group and Love phase (separate processes: the variant is chosen at first use)."""
    import json
    import os
    import subprocess
    import sys

    code = (
        "import sys, json, hashlib, numpy as np; sys.path.insert(0, %r)\n"
        "from dsurftomo_b200 import api, inputs\n"
        "pb = inputs.synthetic_problem(12, 3, 6, ('Rc', 'Rg', 'Lc', 'Lg'), nrecv=5, name='small_4types')\n"
        "rng = np.random.default_rng(11)\n"
        "vs = (pb.vsf * (1.0 + 0.04 * rng.standard_normal(pb.vsf.shape))).astype(np.float32)\n"
        "h = hashlib.sha256()\n"
        "for iwave, igr, t in ((2, 1, pb.tRg), (1, 0, pb.tLc), (2, 0, pb.tRc)):\n"
        "    for a in api.depthkernel(pb.nx, pb.ny, pb.nz, vs, iwave, igr, len(t), t, pb.depz, pb.minthk):\n"
        "        assert np.isfinite(a).all(); h.update(np.ascontiguousarray(a).tobytes())\n"
        "print(h.hexdigest())\n"
    ) % (ROOT,)
    out = {}
    for name, val in (("first", "0"), ("otf", otf)):
        env = dict(os.environ, DSURF_DISP_OTF=val)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        out[name] = r.stdout.strip().splitlines()[-1]
    assert out["first"] == out["otf"]
