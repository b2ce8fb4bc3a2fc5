"""Measured deviation of the block-level fast-iterative eikonal (DSURF_EIKONAL=fim, eik_fim.cuh) from the exact-order
kernel, on the GPU, end to end (the table VERDICT r01 item 1b asks for; summarised in profiles/r02_fim_parity.md):

  * travel-time fields of sampled sweeps against the CPU oracle's heap march: fraction of nodes that differ, |dT|/T
    percentiles of those that do;
  * the whole sweep stage in both modes: predicted times (dsurf), rays whose B-spline vertex pattern differs, (row, col)
    entries in the symmetric difference of the two sparsity patterns, largest relative difference of the Frechet values;
  * Vs model after one outer iteration (device glue -> LSMR -> update) from both G matrices.

usage: python scripts/fim_parity.py [taipei] [small] [cfg2] [cfg3slice[:periods]] [checker]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as O  # noqa: E402
from dsurftomo_b200 import api, inputs  # noqa: E402


def stage(pb, mode, disp=None):
    prev = api.set_eikonal_mode(mode)
    plan = api.Plan(pb)
    api.set_eikonal_mode(prev)
    if disp is None:
        plan.dispersion()
    else:
        pv4, sen12 = disp
        for t in range(4):
            if (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)[t] > 0:
                plan.set_dispersion(t, pv4[t], *sen12[3 * t:3 * t + 3])
            elif t in (0, 2):
                plan.set_dispersion(t, pv4[t], None, None, None)
        plan.finalize_dispersion()
    plan.reset_rows()
    t0 = time.perf_counter()
    plan.sweeps()
    out = plan.download()
    out["wall_s"] = time.perf_counter() - t0
    out["timings"] = plan.timings()
    return plan, out


def model_after_iteration(plan, pb):
    sysl = api.LsmrSystem.from_plan(plan)
    L = sysl.solve(pb.damp)
    vs, _ = plan.update_model(sysl)
    sysl.close()
    return np.asarray(vs), L["itn"]


def field_stats(pb, plan_fim, ngather=6, seed=1):
    """travel-time fields of sampled phase-velocity sweeps of the fim plan against the oracle's exact heap march on the
    same phase-velocity map (read back from the plan)"""
    rng = np.random.default_rng(seed)
    cum = np.concatenate([[0], np.cumsum(pb.nsrc1)])
    cand = [g for g in rng.permutation(plan_fim.num_gathers)[: 8 * ngather]
            if pb.igrt[int(np.searchsorted(cum, g, side="right")) - 1, g - int(cum[int(np.searchsorted(cum, g, side="right")) - 1])] == 0]
    picks = sorted(int(g) for g in cand[:ngather])
    pvs = {}
    differing, total, rels, per_sweep = 0, 0, [], []
    for g in picks:
        k = int(np.searchsorted(cum, g, side="right")) - 1
        s = g - int(cum[k])
        typ = 0 if int(pb.wavetype[k, s]) == 2 else 2
        if typ not in pvs:
            pvs[typ] = plan_fim.get_dispersion(typ)[0]
        pv = pvs[typ][int(pb.periods[k, s]) - 1]
        got = plan_fim.debug_sweep(g, 1, want_fdm=False)
        ref = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pv, pb.scxf[k, s], pb.sczf[k, s])
        assert ref["err"] == 0 and np.array_equal(got["veln"].view(np.uint32), ref["veln"].view(np.uint32))
        d = got["ttn"].view(np.uint32) != ref["ttn"].view(np.uint32)
        differing += int(d.sum())
        total += d.size
        rels.append(np.abs(got["ttn"][d].astype(np.float64) - ref["ttn"][d]) / ref["ttn"][d])
        per_sweep.append(int(d.sum()))
    rels = np.concatenate(rels) if rels else np.zeros(0)
    pct = (lambda q: float(np.quantile(rels, q))) if rels.size else (lambda q: 0.0)
    return dict(sweeps=len(picks), nodes=total, nodes_differing=differing, frac=differing / max(total, 1),
                differing_per_sweep=per_sweep, rel_p50=pct(0.5), rel_p99=pct(0.99), rel_max=pct(1.0))


def compare(pb, name, disp=None, fields=True, model=True):
    pe, e = stage(pb, "exact", disp)
    pf, f = stage(pb, "fim", disp)
    out = dict(workload=name, sweeps=int(pe.num_sweeps()), rays=int(len(e["dsurf"])),
               exact_eikonal_ms=e["timings"]["eikonal_ms"], fim_eikonal_ms=f["timings"]["eikonal_ms"])
    ds = np.abs(f["dsurf"].astype(np.float64) - e["dsurf"]) / e["dsurf"]
    out["dsurf"] = dict(bit_identical_frac=float((f["dsurf"].view(np.uint32) == e["dsurf"].view(np.uint32)).mean()),
                        rel_max=float(ds.max()), rel_p99=float(np.quantile(ds, 0.99)))
    ka = e["row"].astype(np.int64) * (pb.maxvp + 1) + e["col"]
    kb = f["row"].astype(np.int64) * (pb.maxvp + 1) + f["col"]
    only_a, only_b = np.setdiff1d(ka, kb, assume_unique=True), np.setdiff1d(kb, ka, assume_unique=True)
    rows_diff = np.union1d(only_a // (pb.maxvp + 1), only_b // (pb.maxvp + 1))
    both_a, both_b = np.isin(ka, kb, assume_unique=True), np.isin(kb, ka, assume_unique=True)
    va, vb = e["rw"][both_a], f["rw"][both_b]
    scale = np.abs(va).max() if va.size else 1.0
    out["G"] = dict(nar_exact=int(e["nar"]), nar_fim=int(f["nar"]), entries_only_in_one=int(len(only_a) + len(only_b)),
                    rays_with_pattern_difference=int(len(rows_diff)),
                    value_bit_identical_frac=float((va.view(np.uint32) == vb.view(np.uint32)).mean()) if va.size else 1.0,
                    value_absdiff_max_over_scale=float(np.abs(va - vb).max() / scale) if va.size else 0.0)
    if model:
        ve, ie = model_after_iteration(pe, pb)
        vf, if_ = model_after_iteration(pf, pb)
        out["vs_model"] = dict(rel_max=float(np.abs(vf / ve - 1).max()), lsmr_itn=(int(ie), int(if_)))
    if fields:
        out["fields_vs_oracle"] = field_stats(pb, pf)
    pe.close()
    pf.close()
    return out


def checker_problem():
    """checkerboard velocity model with 12 % anomalies every 4th vertex (sharp contrasts: the worst case for the
    heap-order effects)"""
    pb = inputs.synthetic_problem(35, 4, 32, ("Rc",), nrecv=12, name="checker")
    i, j = np.meshgrid(np.arange(pb.nx), np.arange(pb.ny), indexing="ij")
    sign = np.where(((i // 4 + j // 4) & 1) == 1, 1.0, -1.0).astype(np.float32)
    v = pb.vsf.reshape(pb.nz, pb.ny, pb.nx) if pb.vsf.ndim == 1 else pb.vsf
    pb.vsf = (v * (1.0 + 0.12 * sign.T[None])).astype(np.float32).reshape(pb.vsf.shape)
    return pb


def main():
    want = sys.argv[1:] or ["taipei", "small"]
    res = []
    for w in want:
        if w == "taipei":
            res.append(compare(inputs.config(1), "cfg1 Taipei 121^2"))
        elif w == "small":
            res.append(compare(inputs.synthetic_problem(12, 3, 6, ("Rc", "Rg", "Lc", "Lg"), nrecv=5, name="small_4types"),
                               "small 73^2, 4 data types"))
        elif w == "cfg2":
            res.append(compare(inputs.config(2), "cfg2 257^2 x 8 periods x 64 sources"))
        elif w == "checker":
            res.append(compare(checker_problem(), "checkerboard 257^2 (12 % anomalies, 4-vertex blocks)"))
        elif w.startswith("cfg3slice"):
            nper = int(w.split(":")[1]) if ":" in w else 1
            pb = inputs.synthetic_problem(131, nper, 256, ("Rc", "Rg", "Lc", "Lg"), name=f"cfg3slice_{nper}p")
            disp = inputs.synthetic_dispersion(pb)
            res.append(compare(pb, f"cfg3 1025^2, {nper} period(s) x 4 types x 256 sources", disp=disp, model=False))
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
