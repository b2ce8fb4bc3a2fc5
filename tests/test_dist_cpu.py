"""world_size-2 gloo test of the multi-GPU host logic: gathers are sharded in contiguous blocks,
each rank evaluates its shard (here with the CPU oracle as the compute stand-in), and the
rank-ordered concatenation reproduces the single-process CalSurfG output exactly."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib as O
    from dsurftomo_b200 import dist as ddist, inputs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb = inputs.synthetic_problem(10, 2, 5, ("Rc", "Lg"), nrecv=3, name="dist_small")
    g0, g1 = ddist.shard_gathers(pb, rank, world)
    pv4, sen12 = inputs.synthetic_dispersion(pb)
    loc = O.calsurfg_pre(pb, pv4, sen12, g0, g1, nthreads=2, maxnar=400000)
    r0, r1 = ddist.rows_of_gathers(pb, g0, g1)
    local = dict(row=loc["row"], col=loc["col"], rw=loc["rw"], dsurf=loc["dsurf"][r0:r1], nar=loc["nar"])
    full = ddist.all_gather_rows(local)
    if rank == 0:
        ref = O.calsurfg_pre(pb, pv4, sen12, -1, -1, nthreads=2, maxnar=400000)
        ok = (full["nar"] == ref["nar"] and np.array_equal(full["row"], ref["row"]) and
              np.array_equal(full["col"], ref["col"]) and np.array_equal(full["rw"], ref["rw"]) and
              np.array_equal(full["dsurf"], ref["dsurf"]))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gathers_concatenate_to_full_result():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from dsurftomo_b200 import dist as ddist, inputs

    pb = inputs.synthetic_problem(10, 4, 5, ("Rc", "Rg"), nrecv=3)
    spans = [ddist.shard_gathers(pb, r, 4) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == pb.ngathers
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all((b - a) % 5 == 0 for a, b in spans)  # whole period-types per rank (kmax >= world)
    assert ddist.shard_range(10, 0, 3) == (0, 4) and ddist.shard_range(10, 2, 3) == (7, 10)
    sysd = dict(m=6, n=3, rows=np.array([1, 1, 2, 3, 4, 5, 6, 6], np.int32), cols=np.ones(8, np.int32),
                vals=np.ones(8, np.float32), cbst=np.arange(6, dtype=np.float32))
    parts = [ddist.partition_system(sysd, r, 2) for r in range(2)]
    assert sum(p["m"] for p in parts) == 6 and sum(len(p["vals"]) for p in parts) == 8
    assert parts[1]["rows"].min() == 1
