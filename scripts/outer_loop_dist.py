"""Device-resident outer Gauss-Newton loop of main.f90:348-546 on N GPUs (torchrun), the way BASELINE configs[4]
(2049 x 2049 grid, 32 periods x 1024 sources, 8 GPUs) has to run: the COO never passes through the reference's int32
(rw, iw, col) boundary and is NEVER gathered -- every rank keeps the row block it produced (64-bit counts globally,
int32 inside a rank), the only exchanges are the predicted times (4 bytes per ray, one all-gather per iteration) and
the per-iteration all-reduce inside LSMR; the model update is applied identically on every rank.

  torchrun --nproc-per-node N scripts/outer_loop_dist.py --nxy 259 --periods 32 --sources 1024 --iters 10   # cfg 5
Prints one JSON line per outer iteration on rank 0 (sweeps/s, LSMR it/s, SpMV/SpMTV GB/s per GPU, residual statistics)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dsurftomo_b200 import api, inputs, dist as ddist

ap = argparse.ArgumentParser()
ap.add_argument("--nxy", type=int, default=259)
ap.add_argument("--periods", type=int, default=32)
ap.add_argument("--sources", type=int, default=1024)
ap.add_argument("--types", default="Rc")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--nrecv", type=int, default=16)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = ddist.NcclComm(rank, world, local)

pb = inputs.synthetic_problem(args.nxy, args.periods, args.sources, tuple(args.types.split(",")), nrecv=args.nrecv,
                              name=f"outer_{(args.nxy - 3) * 8 + 1}sq_{args.periods}p_{args.sources}s_{args.types}")
plan = api.Plan(pb)
g0, g1 = ddist.shard_gathers(pb, rank, world)
vs = pb.vsf.copy()
for it in range(1, args.iters + 1):
    t0 = time.perf_counter()
    plan.dispersion()                       # K1 on the current model (replicated: every rank needs every map it sweeps)
    tdisp = plan.timings()["dispersion_ms"]
    plan.reset_rows()
    plan.sweeps(g0, g1)
    tm = plan.timings()
    if comm is not None:
        plan.allgather(comm, want_coo=False)  # predicted times only
    dsurf = plan.download()["dsurf"]
    res = pb.obst - dsurf
    sysl = api.LsmrSystem.from_plan_shard(plan, rank, world)
    assert sysl.nnz < 2 ** 31
    if comm is not None:
        ddist.attach(sysl, comm)
    L = sysl.solve(pb.damp)
    vs, dv = plan.update_model(sysl)
    nnz_g = torch.tensor([float(sysl.nnz), float(sysl.m), float(tm["sweeps"]), tm["total_ms"]], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([tm["total_ms"], L["ms_total"]], dtype=torch.float64, device="cuda")
    if comm is not None:
        dist.all_reduce(nnz_g, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps(dict(
            iter=it, n_gpus=world, workload=pb.name, eikonal=os.environ.get('DSURF_EIKONAL', 'exact'), sweeps=int(nnz_g[2].item()),
            sweeps_per_s=nnz_g[2].item() / (tmax[0].item() / 1e3), sweep_stage_ms=tmax[0].item(), dispersion_ms=tdisp,
            nnz_global=int(nnz_g[0].item()), nnz_global_exceeds_int32=bool(nnz_g[0].item() >= 2 ** 31), m_global=int(nnz_g[1].item()),
            nnz_this_rank=int(sysl.nnz), lsmr_itn=L["itn"], lsmr_istop=L["istop"], lsmr_iters_per_s=L["itn"] / (tmax[1].item() / 1e3),
            spmv_gbs_per_gpu=8.0 * sysl.nnz * L["itn"] / (L["ms_spmv"] / 1e3) / 1e9,
            spmtv_gbs_per_gpu=8.0 * sysl.nnz * L["itn"] / (L["ms_spmtv"] / 1e3) / 1e9,
            residual_mean=float(res.mean()), residual_std=float(res.std()), residual_rms=float(np.sqrt((res ** 2).mean())),
            dv_min=float(dv.min()), dv_max=float(dv.max()), wall_s=time.perf_counter() - t0)), flush=True)
    sysl.close()
plan.close()
if comm is not None:
    comm.close()
    dist.destroy_process_group()
