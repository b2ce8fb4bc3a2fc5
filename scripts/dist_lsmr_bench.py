#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/dist_lsmr_bench.py [--rows R] [--iters K] [--strong]

Weak-scaling probe of the distributed LSMR (lsmrModule.f90:475-616 with aprod.f90:40-55 split by rows): every rank
owns R ray rows of the cfg-3 shape (scripts/lsmr_bench.py's generator, a different seed per rank; rank 0 also holds the
smoothing rows), the n-vectors are replicated, and each iteration exchanges the partial A'u (n floats) + ||u||^2.
--strong: the SAME global system (R rows in total) is cut into N row blocks instead.
The exchange is the peer-memory one-shot reduction inside a CUDA graph (default) or NCCL's all-reduce
(DSURF_LSMR_NCCL_ONLY=1).  Rank 0 prints one JSON line; x must be identical on every rank."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from dsurftomo_b200 import api, dist as ddist  # noqa: E402
import lsmr_bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--strong", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = ddist.NcclComm(rank, world, local)
    if a.strong:
        m, n, R, Cc, V, b, geom = lsmr_bench.build(a.rows, seed=3)
        sysd = dict(m=m, n=n, rows=R, cols=Cc, vals=V, cbst=b)
        part = ddist.partition_system(sysd, rank, world)
        m, R, Cc, V, b = part["m"], part["rows"], part["cols"], part["vals"], part["cbst"]
    else:
        m, n, R, Cc, V, b, geom = lsmr_bench.build(a.rows, seed=3 + rank)
        if rank != 0:  # smoothing rows (the tail of the generator's system) live on rank 0 only
            keep = R <= a.rows
            R, Cc, V, b, m = R[keep], Cc[keep], V[keep], b[: a.rows], a.rows
    api.lsmr_hint_geometry(*geom)
    sysl = api.LsmrSystem(m, n, R, Cc, V, b)
    ddist.attach(sysl, comm)
    sysl.solve(1.0, itnlim=6, force_iters=True, want_x=False)
    best = None
    for _ in range(3):
        dist.barrier()
        L = sysl.solve(1.0, itnlim=a.iters, force_iters=True, want_x=True)
        t = torch.tensor([L["ms_total"]], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        L["ms_max"] = float(t.item())
        if best is None or L["ms_max"] < best["ms_max"]:
            best = L
    # replicated vectors must stay bit-identical on every rank
    x = torch.from_numpy(best["x"].view(np.int32).copy()).cuda()
    x0 = x.clone()
    dist.broadcast(x0, src=0)
    same = torch.tensor([int(torch.equal(x, x0))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    nnz_local = torch.tensor([len(V)], device="cuda", dtype=torch.int64)
    m_local = torch.tensor([m], device="cuda", dtype=torch.int64)
    dist.all_reduce(nnz_local)
    dist.all_reduce(m_local)
    if rank == 0:
        nnz, mg = int(nnz_local.item()), int(m_local.item())
        it_s = best["itn"] / (best["ms_max"] / 1e3)
        bytes_it = 16 * nnz + 8 * (mg + 1) + 12 * mg + 80 * n * world
        print(json.dumps(dict(n_gpus=world, scaling="strong" if a.strong else "weak",
                              exchange="nccl" if os.environ.get("DSURF_LSMR_NCCL_ONLY") else "peer-memory",
                              m_global=mg, n=n, nnz_global=nnz, nnz_per_gpu=nnz / world, iters=best["itn"],
                              iters_per_s=it_s, us_per_iter=1e6 / it_s, b_lsmr_gbs_per_gpu=bytes_it * it_s / 1e9 / world,
                              x_identical_on_all_ranks=bool(same.item()), normx=best["normx"], normr=best["normr"])),
              flush=True)
    dist.barrier()
    sysl.close()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
