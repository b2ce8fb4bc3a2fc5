"""torchrun --nproc-per-node 2 scripts/dist_lsmr_check.py : distributed LSMR (row-partitioned, one
NCCL all-reduce per iteration) against the single-GPU solve of the same Taipei system."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dsurftomo_b200 import api, inputs, hostglue, dist as ddist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pb = inputs.config(1)
g = api.CalSurfG(pb)
s = hostglue.host_glue(pb, g["dsurf"], g["row"], g["col"], g["rw"])
part = ddist.partition_system(s, rank, world)
sysl = api.LsmrSystem(part["m"], part["n"], part["rows"], part["cols"], part["vals"], part["cbst"])
comm = ddist.NcclComm(rank, world, local)
ddist.attach(sysl, comm)
L = sysl.solve(pb.damp)
if rank == 0:
    full = api.LsmrSystem(s["m"], s["n"], s["rows"], s["cols"], s["vals"], s["cbst"])
    R = full.solve(pb.damp)
    err = float(np.abs(L["x"] - R["x"]).max())
    print(f"dist itn={L['itn']} istop={L['istop']} single itn={R['itn']} istop={R['istop']} max|dx|={err:.3e}", flush=True)
    assert abs(L["itn"] - R["itn"]) <= 2 and err <= 1e-5, "distributed LSMR deviates"
    print("DIST_LSMR_OK")
dist.barrier()
sysl.close(); comm.close()
dist.destroy_process_group()
