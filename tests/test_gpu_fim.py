"""GPU tests of the fast-iterative eikonal pipeline (DSURF_EIKONAL=fim: k_refine -> k_fim_start -> k_fim_march) against
the oracle and against the exact-order kernel: travel times within 1e-5 relative (north_star), deviations confined to
a small fraction of the nodes, identical B-spline vertex patterns of the rays on the tested problems."""
import numpy as np
import pytest

import oracle_lib as O
from dsurftomo_b200 import api, inputs

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _hetero(pb):
    ii = np.arange(pb.nx)[None, :]
    jj = np.arange(pb.ny)[:, None]
    return (1.3 * (1.0 + 0.15 * np.sin(0.9 * ii + 0.2) * np.cos(0.7 * jj))).ravel().astype(np.float32).astype(np.float64)


@pytest.fixture()
def fim_mode():
    prev = api.set_eikonal_mode("fim")
    yield
    api.set_eikonal_mode(prev)


@pytest.mark.parametrize("which", ["uniform", "hetero"])
def test_fim_sweep_fields_vs_oracle(fim_mode, which):
    pb = inputs.config(1)
    plan = api.Plan(pb)
    pv = np.full(pb.nx * pb.ny, 1.3) if which == "uniform" else _hetero(pb)
    for per in range(pb.kmaxRc):
        plan.set_map(0, per, pv)
    cum = np.concatenate([[0], np.cumsum(pb.nsrc1)])
    tot = dif = 0
    for g in [0, 7, 100, 448]:
        k = int(np.searchsorted(cum, g, side="right")) - 1
        s = g - int(cum[k])
        got = plan.debug_sweep(g, 1)
        ref = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, pb.scxf[k, s], pb.sczf[k, s])
        assert ref["err"] == 0
        # the refined pass is the exact kernel: bit-identical alive set and times
        assert np.array_equal(got["nstsr"] == 0, ref["nstsr"] == 0)
        alive = ref["nstsr"] == 0
        assert np.array_equal(_bits(got["ttnr"])[alive], _bits(ref["ttnr"])[alive])
        assert np.isfinite(got["ttn"]).all() and (got["ttn"] < 1e30).all()
        d = _bits(got["ttn"]) != _bits(ref["ttn"])
        tot += d.size
        dif += int(d.sum())
        assert np.abs(got["ttn"] / ref["ttn"] - 1)[ref["ttn"] > 0].max() <= 1e-5
        nrc = int(pb.nrc1[k, s])
        err, tt, fdm = O.sweep_rays(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, pb.scxf[k, s], pb.sczf[k, s],
                                    pb.rcxf[k, s, :nrc], pb.rczf[k, s, :nrc])
        assert err == 0
        assert np.array_equal(got["fdm"][:nrc] != 0, fdm != 0)  # ray cells / vertex pattern
    assert dif <= 2e-2 * tot, (dif, tot)
    if which == "uniform":
        assert dif == 0
    plan.close()


@pytest.mark.parametrize("name", ["taipei", "small4"])
def test_fim_stage_vs_exact_stage(name):
    pb = inputs.config(1) if name == "taipei" else inputs.synthetic_problem(12, 3, 6, ("Rc", "Rg", "Lc", "Lg"), nrecv=5,
                                                                          name="small_4types")
    outs = {}
    for mode in ("exact", "fim"):
        prev = api.set_eikonal_mode(mode)
        try:
            outs[mode] = api.CalSurfG(pb)
        finally:
            api.set_eikonal_mode(prev)
    e, f = outs["exact"], outs["fim"]
    assert np.abs(f["dsurf"] / e["dsurf"] - 1).max() <= 1e-5
    a = set(zip(e["row"].tolist(), e["col"].tolist()))
    b = set(zip(f["row"].tolist(), f["col"].tolist()))
    assert len(a ^ b) <= 1e-4 * len(a), (len(a ^ b), len(a))
    if e["nar"] == f["nar"] and np.array_equal(e["col"], f["col"]):
        assert np.abs(e["rw"] - f["rw"]).max() <= 1e-4 * np.abs(e["rw"]).max()


def test_fim_batched_and_chunked_equals_unbatched(fim_mode, monkeypatch):
    """forcing small batches (several launches, partially filled slots) must give the same result; the relaxation
    order depends on warp scheduling, the fixed point does not -- allow last-bit differences only"""
    pb = inputs.synthetic_problem(12, 2, 6, ("Rc", "Rg"), nrecv=5, name="fim_batch")
    ref = api.CalSurfG(pb)
    monkeypatch.setenv("DSURF_MAXSLOTS", "5")
    got = api.CalSurfG(pb)
    assert np.abs(got["dsurf"] / ref["dsurf"] - 1).max() <= 1e-6
    assert got["nar"] == ref["nar"] and np.array_equal(got["col"], ref["col"])
    assert np.abs(got["rw"] - ref["rw"]).max() <= 1e-5 * np.abs(ref["rw"]).max()


def test_fim_full_size_field_within_north_star_tolerance(fim_mode):
    """1025 x 1025 propagation grid (BASELINE configs[2]): every node reached, travel times within 1e-5 relative of the
    oracle's heap march (measured <= 2.2e-6), only a few per cent of the nodes differ at all, |grad T| bounded"""
    pb = inputs.synthetic_problem(131, 1, 8, ("Rc",), nrecv=16, name="fim_full_size")
    pv4, sen12 = inputs.synthetic_dispersion(pb)
    plan = api.Plan(pb)
    plan.set_dispersion(0, pv4[0], *sen12[0:3])
    plan.finalize_dispersion()
    tot = dif = 0
    for g in (1, 6):
        got = plan.debug_sweep(g, 1, want_fdm=False)
        ref = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv4[0][0], pb.scxf[0, g], pb.sczf[0, g])
        assert ref["err"] == 0 and got["ttn"].shape == (1025, 1025)
        assert np.array_equal(_bits(got["veln"]), _bits(ref["veln"]))
        t = got["ttn"]
        assert np.isfinite(t).all() and t.min() >= 0.0 and t.max() < 1e30
        pos = ref["ttn"] > 0
        assert np.abs(t[pos] / ref["ttn"][pos] - 1).max() <= 1e-5
        d = _bits(t) != _bits(ref["ttn"])
        tot += d.size
        dif += int(d.sum())
        h = 6371.0 * np.deg2rad(pb.dvxd) / 8
        assert max(np.abs(np.diff(t, axis=0)).max(), np.abs(np.diff(t, axis=1)).max()) <= 1.5 * h / got["veln"].min()
    assert dif <= 0.1 * tot, (dif, tot)
    plan.close()
