"""torchrun --nproc-per-node N scripts/dist_check.py [taipei|small] : the multi-GPU path against the single-GPU one.

 1. every rank runs the sweep stage on its contiguous share of the gathers (dist.shard_gathers), then one
    dsurf_plan_allgather (NCCL all-gather of the counts + grouped broadcasts of the predicted times and the COO row
    blocks); rank 0 also runs all gathers alone.  The gathered (row, col, rw) and dsurf must equal the single-GPU
    ones BIT FOR BIT (array comparison and the order-sensitive device digest).
 2. distributed LSMR on the true row partition (each rank keeps the rows it produced + a share of the smoothing rows,
    device glue, one all-reduce per iteration) against the single-GPU solve of the full system.
Prints DIST_CHECK_OK on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dsurftomo_b200 import api, inputs, dist as ddist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[1] if len(sys.argv) > 1 else "taipei"
pb = inputs.config(1) if which == "taipei" else inputs.synthetic_problem(19, 3, 10, ("Rc", "Rg", "Lc", "Lg"), nrecv=6, name="dist_small")
comm = ddist.NcclComm(rank, world, local)

plan = api.Plan(pb)
plan.dispersion()
g0, g1 = ddist.shard_gathers(pb, rank, world)
plan.reset_rows()
plan.sweeps(g0, g1)
ntot = plan.allgather(comm, want_coo=True)
got = plan.download_gathered(ntot)
got_dsurf = plan.download()["dsurf"]
dig_g = plan.digest(gathered=True)

ok = True
if rank == 0:
    ref_plan = api.Plan(pb)
    ref_plan.dispersion()
    ref_plan.reset_rows()
    ref_plan.sweeps()
    ref = ref_plan.download()
    dig_r = ref_plan.digest()
    same = (ref["nar"] == ntot and np.array_equal(ref["row"], got["row"]) and np.array_equal(ref["col"], got["col"]) and
            np.array_equal(ref["rw"].view(np.uint32), got["rw"].view(np.uint32)) and
            np.array_equal(ref["dsurf"].view(np.uint32), got_dsurf.view(np.uint32)) and dig_g == dig_r)
    print(f"gather: nar {ntot} vs {ref['nar']}, digest {dig_g[0]:016x} vs {dig_r[0]:016x}, identical={same}", flush=True)
    ok &= bool(same)

# ---- distributed LSMR on the row partition vs the single-GPU full system
sysl = api.LsmrSystem.from_plan_shard(plan, rank, world)
ddist.attach(sysl, comm)
L = sysl.solve(pb.damp)
if rank == 0:
    full = api.LsmrSystem.from_plan(ref_plan)
    R = full.solve(pb.damp)
    err = float(np.abs(L["x"] - R["x"]).max())
    print(f"lsmr: dist itn={L['itn']} istop={L['istop']} (m_local={sysl.m}, nnz_local={sysl.nnz}) | single itn={R['itn']} "
          f"istop={R['istop']} (m={full.m}, nnz={full.nnz}) | max|dx|={err:.3e}", flush=True)
    ok &= abs(L["itn"] - R["itn"]) <= 2 and err <= 1e-5
    full.close()
    ref_plan.close()
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
dist.barrier()
sysl.close()
plan.close()
comm.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
