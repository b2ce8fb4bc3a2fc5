"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden vectors,
SURVEY.md section 4 / 8c): analytic solutions per stage plus independent numerical references."""
import numpy as np
import pytest
import scipy.optimize
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import oracle_lib as O
from dsurftomo_b200 import hostglue, inputs


# ---------------------------------------------------------------- dispersion
def test_rayleigh_halfspace_poisson():
    """Homogeneous Poisson solid, flat earth: c = 0.9194 beta at every period, U = c."""
    vs = np.array([3.0, 3.0, 3.0], np.float32)
    vp = (vs * np.sqrt(3.0)).astype(np.float32)
    rho = np.full(3, 2.7, np.float32)
    thk = np.array([2.0, 5.0, 0.0], np.float32)
    t = np.array([0.5, 1.0, 2.0, 5.0, 10.0])
    c, nf = O.surfdisp96(thk, vp, vs, rho, 0, 2, 1, 0, t)
    assert nf == 0
    np.testing.assert_allclose(c / 3.0, 0.9194, atol=2e-4)
    u, _ = O.surfdisp96(thk, vp, vs, rho, 0, 2, 1, 1, t)
    np.testing.assert_allclose(u, c, rtol=3e-4)


def test_love_one_layer_closed_form():
    """Layer (b1, rho1, H) over half-space (b2, rho2): tan(k s1 H) = mu2 s2 / (mu1 s1)."""
    b1, b2, r1, r2, H = 2.0, 3.5, 2.3, 2.9, 4.0
    thk = np.array([H, 0.0], np.float32)
    vs = np.array([b1, b2], np.float32)
    vp = (vs * 1.75).astype(np.float32)
    rho = np.array([r1, r2], np.float32)
    t = np.array([2.0, 3.0, 5.0, 8.0])
    c, nf = O.surfdisp96(thk, vp, vs, rho, 0, 1, 1, 0, t)
    assert nf == 0
    for T, ck in zip(t, c):
        w = 2 * np.pi / T

        def f(cc):
            k = w / cc
            s1 = np.sqrt((cc / b1) ** 2 - 1.0)
            s2 = np.sqrt(1.0 - (cc / b2) ** 2)
            return np.tan(k * s1 * H) - (r2 * b2 ** 2 * s2) / (r1 * b1 ** 2 * s1)

        # fundamental branch: k s1 H in (0, pi/2)
        cs = np.linspace(b1 * 1.0001, b2 * 0.9999, 20000)
        k = w / cs
        ok = (k * np.sqrt((cs / b1) ** 2 - 1.0) * H) < np.pi / 2
        vals = np.array([f(x) for x in cs[ok]])
        i = np.nonzero(np.sign(vals[:-1]) != np.sign(vals[1:]))[0][0]
        root = scipy.optimize.brentq(f, cs[ok][i], cs[ok][i + 1])
        assert abs(ck - root) <= 2e-5 * root


def test_refine_grid2layer():
    dep = np.array(inputs.TAIPEI_DEPZ, np.float32)
    vs = (0.9 + 0.6 * dep).astype(np.float32)
    rdep, rvp, rvs, rrho, rthk = O.refine_grid2layer(3.0, dep, vs * 1.7, vs, vs * 0 + 2.5)
    assert len(rthk) == (len(dep) - 1) * 4 + 1  # int((thk+1e-4)/(thk/3))+1 = 4 sublayers
    assert rthk[-1] == 0.0
    np.testing.assert_allclose(rthk[:-1].sum(), dep[-1] - dep[0], rtol=1e-6)
    # mid-point interpolation of a linear profile reproduces it at layer centres
    centres = np.concatenate([[0], np.cumsum(rthk[:-1])])[:-1] + rthk[:-1] / 2
    np.testing.assert_allclose(rvs[:-1], 0.9 + 0.6 * centres, rtol=2e-6)


def test_depth_kernels_consistent_with_direct_difference():
    pb = inputs.synthetic_problem(6, 2, 2, ("Rc",), nrecv=1)
    t = np.array([0.8, 1.5])
    pv, svs, svp, srho = O.depthkernel(pb.vsf, 2, 0, t, pb.depz, pb.minthk, nthreads=4)
    assert np.all(pv > 0.3) and np.all(pv < 3.0)
    # dc/dVs dominates and is positive for the nodes the wave samples; dc/dVp is smaller
    col = 2 * pb.nx + 3
    assert svs[:, :, col].sum() > 0.3
    assert np.abs(svp[:, :, col]).sum() < np.abs(svs[:, :, col]).sum()
    # a uniform 1 % increase of Vs at all nodes changes c by about sum_i dc/dVs_i * 0.01 Vs_i
    # (Vp, rho follow Brocher in the model but are held fixed by the partial derivative)
    vs_col = pb.vsf[:, 2, 3]
    pred = (svs[:, :, col] * (0.01 * vs_col)[:, None]).sum(axis=0)
    assert np.all(pred > 0)


# ---------------------------------------------------------------- eikonal / rays
def _uniform(pb, v=1.1):
    return np.full(pb.nx * pb.ny, v)


def test_fmm_uniform_velocity_is_distance(taipei):
    pb = taipei
    v = 1.1
    k, s = 3, 2
    scx, scz = pb.scxf[k, s], pb.sczf[k, s]
    r = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, _uniform(pb, v), scx, scz)
    assert r["err"] == 0
    nnx, nnz = r["ttn"].shape
    g = O  # geometry of the propagation grid (CalSurfG.f90:1050-1063)
    pi = np.float32(3.1415926535898)
    gox = (90.0 - pb.goxd) * pi / 180.0
    goz = pb.gozd * pi / 180.0
    dnx = pb.dvxd * pi / 180.0 / 8
    dnz = pb.dvzd * pi / 180.0 / 8
    X = gox + dnx * np.arange(nnx)[:, None] + 0 * np.arange(nnz)[None, :]
    Z = goz + dnz * np.arange(nnz)[None, :] + 0 * np.arange(nnx)[:, None]
    d = np.array([[inputs.delsph(scx, scz, X[i, j], Z[i, j]) for j in range(0, nnz, 7)] for i in range(0, nnx, 7)])
    t = r["ttn"][::7, ::7]
    far = d > 1.0
    # first-arrival times on a uniform medium: distance / v with O(h) discretisation error
    assert np.abs(t[far] * v / d[far] - 1).max() < 0.02
    assert r["veln"].min() == pytest.approx(v, rel=1e-6) and r["veln"].max() == pytest.approx(v, rel=1e-6)


def test_fmm_reciprocity_and_receiver_times(taipei):
    pb = taipei
    k = 5
    a, b = 0, 1
    pv = _uniform(pb, 1.3)
    ax, az = pb.scxf[k, a], pb.sczf[k, a]
    bx, bz = pb.rcxf[k, a, 0], pb.rczf[k, a, 0]
    _, t_ab, _ = O.sweep_rays(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, ax, az, [bx], [bz])
    _, t_ba, _ = O.sweep_rays(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, bx, bz, [ax], [az])
    d = float(inputs.delsph(ax, az, bx, bz))
    assert abs(t_ab[0] - t_ba[0]) < 1.5e-2 * t_ab[0]  # O(h) scheme on a 1.7 km grid
    assert abs(t_ab[0] * 1.3 / d - 1) < 0.02


def test_ray_frechet_mass_uniform(taipei):
    """Uniform velocity: sum over the full fdm(0:nvz+1,0:nvx+1) = -L'/v^2 with L' the traced length
    (great-circle distance minus at most ~2 dpl)."""
    pb = taipei
    v = 1.25
    k, s = 10, 1
    nrc = int(pb.nrc1[k, s])
    err, tt, fdm = O.sweep_rays(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, _uniform(pb, v),
                                pb.scxf[k, s], pb.sczf[k, s], pb.rcxf[k, s, :nrc], pb.rczf[k, s, :nrc])
    assert err == 0
    dist = inputs.delsph(pb.scxf[k, s], pb.sczf[k, s], pb.rcxf[k, s, :nrc], pb.rczf[k, s, :nrc])
    mass = -fdm.reshape(nrc, -1).sum(axis=1) * v * v
    dpl = 0.5 * 6371.0 * np.deg2rad(min(pb.dvxd, pb.dvzd * np.cos(np.deg2rad(pb.goxd)))) / 8
    assert np.all(mass <= dist * 1.01 + 1e-3)
    assert np.all(mass >= dist - 4 * dpl - 0.01 * dist)
    assert np.all(fdm <= 1e-7)  # all contributions are negative


# ---------------------------------------------------------------- solver
def _random_system(rng, m, n, nnz):
    rows = np.sort(rng.integers(1, m + 1, nnz)).astype(np.int32)
    cols = rng.integers(1, n + 1, nnz).astype(np.int32)
    vals = rng.standard_normal(nnz).astype(np.float32)
    return rows, cols, vals


def test_synthetic_forward_times(taipei):
    """subroutine synthetic (CalSurfG.f90:2412-2865) on the oracle: times through the start model on
    the gd = 5 grid agree with CalSurfG's predicted times on the gd = 8 grid to the O(h)
    discretisation error, the dispersion maps are caldespersion's, and every time is ~ distance /
    (a velocity inside the map's range)."""
    pb = taipei
    syn = O.synthetic(pb, nthreads=8)
    assert syn["err"] == 0
    full = O.calsurfg(pb, nthreads=8, mode=1)
    assert np.abs(syn["obst"] / full["dsurf"] - 1).max() < 0.03
    pv = O.caldespersion(pb.vsf, 2, 0, pb.tRc, pb.depz, pb.minthk, nthreads=8)
    assert np.array_equal(np.asarray(pv).reshape(pb.kmaxRc, -1), syn["pv"][: pb.kmaxRc])
    vapp = pb.dist / syn["obst"]
    interior = syn["pv"][: pb.kmaxRc].reshape(pb.kmaxRc, pb.ny, pb.nx)[:, 1:-1, 1:-1]
    assert vapp.min() > 0.97 * interior.min() and vapp.max() < 1.03 * interior.max()


def test_aprod_vs_scipy():
    rng = np.random.default_rng(1)
    m, n, nnz = 200, 50, 1500
    rows, cols, vals = _random_system(rng, m, n, nnz)
    A = sp.coo_matrix((vals.astype(np.float64), (rows - 1, cols - 1)), shape=(m, n)).tocsr()
    iw = O.pack_iw(rows, cols)
    x = rng.standard_normal(n).astype(np.float32)
    y = rng.standard_normal(m).astype(np.float32)
    _, y1 = O.aprod(1, m, n, x.copy(), y.copy(), iw, vals)
    np.testing.assert_allclose(y1, y + A @ x, rtol=2e-5, atol=2e-5)
    x2, _ = O.aprod(2, m, n, x.copy(), y.copy(), iw, vals)
    np.testing.assert_allclose(x2, x + A.T @ y, rtol=2e-5, atol=2e-5)


def test_lsmr_vs_scipy_and_dense():
    rng = np.random.default_rng(2)
    m, n, nnz = 400, 60, 4000
    rows, cols, vals = _random_system(rng, m, n, nnz)
    b = rng.standard_normal(m).astype(np.float32)
    damp = 0.5
    iw = O.pack_iw(rows, cols)
    r = O.lsmr(m, n, iw, vals, b, damp, atol=1e-7, btol=1e-7, conlim=1e8)
    A = sp.coo_matrix((vals.astype(np.float64), (rows - 1, cols - 1)), shape=(m, n)).tocsr()
    xs = spl.lsmr(A, b.astype(np.float64), damp=damp, atol=1e-10, btol=1e-10, conlim=1e10)[0]
    Ad = A.toarray()
    xd = np.linalg.solve(Ad.T @ Ad + damp ** 2 * np.eye(n), Ad.T @ b.astype(np.float64))
    assert np.abs(xs - xd).max() < 1e-8
    assert np.abs(r["x"] - xd).max() <= 1e-4 * np.abs(xd).max()
    assert r["istop"] in (1, 2, 3)


def test_snrm2_and_percentile():
    rng = np.random.default_rng(3)
    a = rng.standard_normal(1000).astype(np.float32)
    assert abs(O.snrm2(a) - np.linalg.norm(a.astype(np.float64))) < 1e-4
    q25, q75 = O.getpercentile(a)
    srt = np.sort(a)
    assert q25 == srt[int(0.25 * 1000) - 1] and q75 == srt[int(0.75 * 1000) - 1]
    g25, g75 = hostglue.getpercentile(a)
    assert (g25, g75) == (q25, q75)


def test_delsph_matches_host_reader(taipei):
    pb = taipei
    d = O.lib().oracle_delsph(float(pb.scxf[0, 0]), float(pb.sczf[0, 0]), float(pb.rcxf[0, 0, 0]), float(pb.rczf[0, 0, 0]))
    assert abs(d - pb.dist[0]) <= 2e-6 * d + 1e-6


# ---------------------------------------------------------------- reader + end to end
def test_taipei_fixture_facts(taipei):
    pb = taipei
    assert (pb.nx, pb.ny, pb.nz, pb.kmaxRc, pb.kmax) == (18, 18, 9, 26, 26)
    assert pb.dall == 2061 and pb.ngathers == 449 and pb.nsweeps == 449
    assert int(pb.nrc1.max()) <= 19 and pb.maxvp == 2048
    assert np.all(pb.wavetype[pb.nrc1 > 0] == 2) and np.all(pb.igrt[pb.nrc1 > 0] == 0)


def test_host_glue_matches_oracle(small_problem):
    pb = small_problem
    ref = O.calsurfg(pb, nthreads=8, mode=1)
    assert ref["err"] == 0 and ref["nar"] > 0
    s = hostglue.host_glue(pb, ref["dsurf"], ref["row"], ref["col"], ref["rw"])
    cap = ref["nar"] + 7 * pb.maxvp + 8
    iw = np.zeros(2 * cap + 1, np.int32)
    rw = np.zeros(cap, np.float32)
    col = np.zeros(cap, np.int32)
    iw[1:ref["nar"] + 1] = ref["row"]
    rw[: ref["nar"]] = ref["rw"]
    col[: ref["nar"]] = ref["col"]
    m, nar, cbst, datw = O.host_glue(pb, ref["dsurf"], iw, rw, col, ref["nar"])
    assert m == s["m"] and nar == len(s["vals"])
    assert np.array_equal(iw[1:nar + 1], s["rows"]) and np.array_equal(iw[nar + 1:2 * nar + 1], s["cols"])
    assert np.array_equal(rw[:nar], s["vals"]) and np.array_equal(cbst[:m], s["cbst"])
    assert np.array_equal(datw, s["datweight"])


def test_oracle_threading_modes_agree(small_problem):
    pb = small_problem
    a = O.calsurfg(pb, nthreads=1, mode=0)
    b = O.calsurfg(pb, nthreads=8, mode=1)
    assert a["nar"] == b["nar"]
    for k in ("dsurf", "rw", "row", "col"):
        assert np.array_equal(a[k], b[k])
    # group gathers: two sweeps per gather, rays only on pass 2
    assert a["nsweeps"] == pb.nsweeps and a["nrays"] == pb.dall


@pytest.mark.slow
def test_taipei_two_outer_iterations_reduce_misfit(taipei):
    pb = taipei
    vs = pb.vsf.copy()
    stds = []
    for _ in range(2):
        ref = O.calsurfg(pb, vels=vs, nthreads=8, mode=1)
        s = hostglue.host_glue(pb, ref["dsurf"], ref["row"], ref["col"], ref["rw"])
        res = s["cbst"][: pb.dall]
        stds.append(float(res.std()))
        L = O.lsmr(s["m"], s["n"], O.pack_iw(s["rows"], s["cols"]), s["vals"], s["cbst"], pb.damp)
        vs, _ = hostglue.model_update(pb, vs, L["x"])
        assert vs.min() >= pb.minvel - 1e-6 and vs.max() <= pb.maxvel + 1e-6
    assert stds[1] < stds[0]


def test_ray_geometry_uniform_medium(taipei):
    """Ray geometry (the reference's raypath.out content, CalSurfG.f90:2276-2283) in a uniform medium:
    starts at the receiver, ends exactly at the source, steps of ~dpl = half the smallest cell edge,
    distance to the source decreases monotonically, total length ~ great-circle distance."""
    pb = taipei
    k, s = 10, 1
    nrc = int(pb.nrc1[k, s])
    sx, sz = pb.scxf[k, s], pb.sczf[k, s]
    paths = O.sweep_paths(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, _uniform(pb, 1.25), sx, sz,
                          pb.rcxf[k, s, :nrc], pb.rczf[k, s, :nrc])
    assert len(paths) == nrc
    pi = np.float32(3.1415926535898)
    dpl = 0.5 * min(pb.dvxd, pb.dvzd * np.sin(np.radians(90 - pb.goxd))) * float(pi) / 180.0 / 8 * 6371.0
    for r, p in enumerate(paths):
        assert p[0, 0] == pb.rcxf[k, s, r] and p[0, 1] == pb.rczf[k, s, r]
        assert p[-1, 0] == sx and p[-1, 1] == sz
        d = np.array([float(inputs.delsph(sx, sz, x, z)) for x, z in p])
        assert np.all(np.diff(d[:-1]) < 0)
        seg = np.array([float(inputs.delsph(p[i, 0], p[i, 1], p[i + 1, 0], p[i + 1, 1])) for i in range(len(p) - 2)])
        if len(seg):
            assert np.abs(seg / dpl - 1).max() < 0.05
        assert abs((seg.sum() + d[-2]) / d[0] - 1) < 0.02
