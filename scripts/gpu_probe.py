"""Quick GPU probe: stage timings on cfg1/cfg2 and a slice of cfg3 (not a benchmark)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsurftomo_b200 import api, inputs, hostglue

out = {}
def stage(pb, tag, g1=None):
    t0 = time.time(); plan = api.Plan(pb); t1 = time.time()
    plan.dispersion(); t2 = time.time()
    plan.reset_rows(); plan.sweeps(0, g1); t3 = time.time()
    tm = plan.timings(); tm.update(create_s=t1 - t0, disp_wall_s=t2 - t1, sweeps_wall_s=t3 - t2, nar=plan.nar)
    tm["sweeps_per_s_eikonal"] = tm["sweeps"] / (tm["eikonal_ms"] / 1e3) if tm["eikonal_ms"] else None
    out[tag] = tm; print(tag, json.dumps(tm), flush=True)
    return plan

pb1 = inputs.config(1)
t0 = time.time(); r = api.CalSurfG(pb1); print("taipei CalSurfG host-buffer call s:", time.time() - t0, "nar", r["nar"], flush=True)
t0 = time.time(); r = api.CalSurfG(pb1); print("taipei CalSurfG 2nd call s:", time.time() - t0, flush=True)
s = hostglue.host_glue(pb1, r["dsurf"], r["row"], r["col"], r["rw"]); iw = hostglue.pack_iw(s["rows"], s["cols"])
t0 = time.time(); L = api.LSMR(s["m"], s["n"], len(iw), len(s["vals"]), iw, s["vals"], s["cbst"], pb1.damp, 1e-6, 1e-6, 100.0, 400, 10)
print("taipei LSMR s:", time.time() - t0, {k: v for k, v in L.items() if k != "x"}, flush=True)
stage(pb1, "cfg1").close()
pb2 = inputs.config(2)
p2 = stage(pb2, "cfg2")
d = p2.download(); p2.close()
s = hostglue.host_glue(pb2, d["dsurf"], d["row"], d["col"], d["rw"])
sysl = api.LsmrSystem(s["m"], s["n"], s["rows"], s["cols"], s["vals"], s["cbst"])
L = sysl.solve(pb2.damp, itnlim=50, force_iters=True, want_x=False)
print("cfg2 lsmr nnz", sysl.nnz, {k: v for k, v in L.items() if k != "x"}, flush=True)
out["cfg2_lsmr"] = dict(nnz=sysl.nnz, **{k: v for k, v in L.items() if k != "x"}); sysl.close()
if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
    pb3 = inputs.synthetic_problem(131, 16, 256, ("Rc",), name="cfg3_Rc_only")
    p3 = stage(pb3, "cfg3_Rc")
    d = p3.download(); p3.close()
    s = hostglue.host_glue(pb3, d["dsurf"], d["row"], d["col"], d["rw"])
    sysl = api.LsmrSystem(s["m"], s["n"], s["rows"], s["cols"], s["vals"], s["cbst"])
    L = sysl.solve(pb3.damp, itnlim=20, force_iters=True, want_x=False)
    print("cfg3_Rc lsmr nnz", sysl.nnz, {k: v for k, v in L.items() if k != "x"}, flush=True)
    out["cfg3_Rc_lsmr"] = dict(nnz=sysl.nnz, **{k: v for k, v in L.items() if k != "x"})
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
