#!/usr/bin/env python
"""Generates tests/golden/oracle_golden.npz: frozen outputs of the CPU oracle (the C++ restatement of the
reference -- the reference itself has no golden vectors and cannot be built here) for the
Taipei fixture and three layered models.  tests/test_golden.py checks (CPU) that the oracle still reproduces
them and (GPU) that the B200 path matches them at the parity tolerances of DESIGN.md section 5.

    python tests/golden/make_golden.py        # rewrites oracle_golden.npz next to this script
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
from dsurftomo_b200 import hostglue, inputs  # noqa: E402

GATHERS = (0, 100, 448)


def layered_models():
    """(thk, vp, vs, rho) of three stacks: gradient crust, low-velocity zone, water-free two-layer."""
    out = []
    thk = np.array([0.5, 1.0, 2.0, 4.0, 8.0, 0.0], np.float32)
    vs = np.array([1.0, 1.6, 2.3, 3.0, 3.5, 4.2], np.float32)
    out.append((thk, (vs * 1.75).astype(np.float32), vs, (1.7 + 0.3 * vs).astype(np.float32)))
    vs2 = np.array([2.0, 1.4, 2.6, 3.2, 3.3, 4.0], np.float32)
    out.append((thk, (vs2 * 1.8).astype(np.float32), vs2, (1.9 + 0.25 * vs2).astype(np.float32)))
    out.append((np.array([3.0, 0.0], np.float32), np.array([4.0, 6.5], np.float32), np.array([2.2, 3.7], np.float32),
                np.array([2.4, 3.0], np.float32)))
    return out


def gather_ks(pb, g):
    cum = np.concatenate([[0], np.cumsum(pb.nsrc1)])
    k = int(np.searchsorted(np.cumsum(pb.nsrc1), g, side="right"))
    return k, g - int(cum[k])


def build():
    g = {}
    t = np.array([0.6, 1.0, 1.7, 3.0, 5.0, 8.0])
    for i, (thk, vp, vs, rho) in enumerate(layered_models()):
        for iwave in (1, 2):
            for igr in (0, 1):
                c, nf = O.surfdisp96(thk, vp, vs, rho, 1, iwave, 1, igr, t)
                g[f"disp_m{i}_w{iwave}_g{igr}"] = c
    g["disp_periods"] = t
    pb = inputs.config(1)
    ref = O.calsurfg(pb, nthreads=8, mode=1)
    assert ref["err"] == 0
    g["taipei_dsurf"] = ref["dsurf"]
    g["taipei_nar"] = np.array([ref["nar"]])
    pat = np.stack([ref["row"], ref["col"]]).astype(np.int32)
    g["taipei_pattern_sha256"] = np.frombuffer(hashlib.sha256(pat.tobytes()).digest(), np.uint8)
    g["taipei_rows_per_ray"] = np.bincount(ref["row"], minlength=pb.dall + 1)[1:].astype(np.int32)
    g["taipei_rw_abs_sum_per_ray"] = np.bincount(ref["row"], weights=np.abs(ref["rw"]).astype(np.float64),
                                                 minlength=pb.dall + 1)[1:]
    s = hostglue.host_glue(pb, ref["dsurf"], ref["row"], ref["col"], ref["rw"])
    L = O.lsmr(s["m"], s["n"], O.pack_iw(s["rows"], s["cols"]), s["vals"], s["cbst"], pb.damp)
    g["taipei_lsmr_x"] = L["x"]
    g["taipei_lsmr_itn_istop"] = np.array([L["itn"], L["istop"]])
    vs1, _ = hostglue.model_update(pb, pb.vsf, L["x"])
    g["taipei_vs_iter1"] = vs1
    pvd, _, _, _ = O.depthkernel(pb.vsf, 2, 0, pb.tRc, pb.depz, pb.minthk, nthreads=8)
    g["taipei_pv_Rc"] = pvd
    for gi in GATHERS:
        k, sidx = gather_ks(pb, gi)
        r = O.fmm_sweep(pb.nx, pb.ny, pb.goxd, pb.gozd, pb.dvxd, pb.dvzd, pvd[pb.periods[k, sidx] - 1], pb.scxf[k, sidx],
                        pb.sczf[k, sidx])
        assert r["err"] == 0
        g[f"taipei_ttn_g{gi}"] = r["ttn"]
    return g


if __name__ == "__main__":
    g = build()
    path = os.path.join(HERE, "oracle_golden.npz")
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path), "bytes;", len(g), "arrays")
