"""Small driver for ncu captures: python scripts/profile_target.py <config> <ngathers> [lsmr_iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dsurftomo_b200 import api, inputs, hostglue

cfg, ng = int(sys.argv[1]), int(sys.argv[2])
its = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pb = bench.build_problem(cfg)
pv4, sen12 = inputs.synthetic_dispersion(pb)
plan = api.Plan(pb)
for t in range(4):
    if (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)[t] > 0:
        plan.set_dispersion(t, pv4[t], *sen12[3 * t:3 * t + 3])
    elif t in (0, 2):
        plan.set_dispersion(t, pv4[t], None, None, None)
plan.finalize_dispersion()
plan.reset_rows(); plan.sweeps(0, ng)
print(plan.timings())
if its:
    res = plan.download()
    nrows = int(res["row"].max())
    srow, scol, sval, cnt3 = hostglue.smoothing_rows(pb.nx, pb.ny, pb.nz, nrows, pb.weight)
    R = np.concatenate([res["row"], srow]); Cc = np.concatenate([res["col"], scol]); V = np.concatenate([res["rw"], sval])
    b = np.concatenate([(pb.obst[:nrows] - res["dsurf"][:nrows]).astype(np.float32), np.zeros(cnt3, np.float32)])
    sysl = api.LsmrSystem(nrows + cnt3, pb.maxvp, R, Cc, V, b)
    print({k: v for k, v in sysl.solve(pb.damp, itnlim=its, force_iters=True, want_x=False).items() if k != "x"}, sysl.nnz)
