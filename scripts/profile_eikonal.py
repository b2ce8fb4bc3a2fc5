"""ncu target: one resident wave of the eikonal kernel on a synthetic problem.
usage: python scripts/profile_eikonal.py <nxy> <sources_per_period> <nperiods> [ngathers]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsurftomo_b200 import api, inputs

nxy, nsrc, nper = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pb = inputs.synthetic_problem(nxy, nper, nsrc, ("Rc",), nrecv=16, name="prof")
pv4, sen12 = inputs.synthetic_dispersion(pb)
plan = api.Plan(pb)
plan.set_dispersion(0, pv4[0], *sen12[0:3])
plan.finalize_dispersion()
ng = int(sys.argv[4]) if len(sys.argv) > 4 else plan.num_gathers
plan.reset_rows()
t0 = time.perf_counter()
plan.sweeps(0, ng)
print("gathers", ng, "sweeps", plan.num_sweeps(0, ng), "wall", time.perf_counter() - t0, plan.timings())
