"""Measured parity budget of the GPU path against the oracle (numbers quoted in DESIGN.md section 5):
ulp flips of dispersion values and depth kernels per data type, and dsurf / pattern / Vs-model deviations of one
outer iteration on the 4-type problem, on Taipei and on the whole BASELINE configs[1].
usage: python scripts/parity_stats.py [cfg2]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import test_gpu_parity as T
from dsurftomo_b200 import api, inputs

out = {}
pb = inputs.synthetic_problem(12, 3, 6, ("Rc", "Rg", "Lc", "Lg"), nrecv=5, name="small_4types")
t = np.array([0.6, 1.0, 1.6])
vs = pb.vsf.reshape(pb.nz, -1).astype(np.float64)
for iwave, igr in ((2, 0), (1, 0), (2, 1), (1, 1)):
    pv, svs, svp, srho = api.depthkernel(pb.nx, pb.ny, pb.nz, pb.vsf, iwave, igr, len(t), t, pb.depz, pb.minthk)
    rpv, rvs, rvp, rrho = O.depthkernel(pb.vsf, iwave, igr, t, pb.depz, pb.minthk, nthreads=8)
    d = np.abs(pv - rpv) / T._ulp32(rpv)
    fl = np.abs(svs - rvs) * (0.01 * vs[:, None, :]) / T._ulp32(rpv)[None]
    out[f"disp iwave={iwave} igr={igr}"] = dict(value_ulps_max=float(d.max()), value_ulps_median=float(np.median(d)),
                                                value_frac_differing=float((d > 0).mean()),
                                                kernel_flips_max=float(fl.max()), kernel_frac_differing=float((fl > 0).mean()),
                                                kernel_rel_of_scale_max=float(np.abs(svs - rvs).max() / np.abs(rvs).max()))
for name, p in (("small_4types", pb), ("taipei", inputs.config(1))) + ((("cfg2", inputs.config(2)),) if "cfg2" in sys.argv else ()):
    r = T._outer_iteration_vs_oracle(p, nthreads=os.cpu_count() or 8)
    out[f"outer iteration {name}"] = {k: r[k] for k in ("dsurf_rel", "vs_rel", "pattern_mismatch", "pattern_total", "itn")}
print(json.dumps(out, indent=1))
