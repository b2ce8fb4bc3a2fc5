"""Multi-GPU host side: one process per GPU, launched by torchrun.

The forward/sensitivity path shards embarrassingly: the independent units are the (period-type,
gather) pairs of the loop nest CalSurfG.f90:1144-1145, partitioned in contiguous blocks so that
concatenating the ranks' outputs in rank order reproduces the reference's row order exactly
(SURVEY.md section 8e).  No collective is needed while rows are produced.  LSMR has one real
exchange per iteration (all-reduce of the partial A'u and ||u||^2), done by NCCL inside
libdsurf_b200.so; torch.distributed is used only to bootstrap (broadcast of the ncclUniqueId)
and to gather results on rank 0 when the host caller wants the full COO.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check, lib


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition [lo, hi) of n_items over world ranks (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_gathers(pb, rank: int, world: int):
    """Gather range of this rank.  Whole period-types are kept together when there are at least
    `world` of them (each rank then needs dispersion maps only for its own periods); otherwise
    gathers are split inside period-types."""
    cum = np.concatenate([[0], np.cumsum(pb.nsrc1)]).astype(np.int64)
    if pb.kmax >= world:
        klo, khi = shard_range(pb.kmax, rank, world)
        return int(cum[klo]), int(cum[khi])
    return shard_range(int(cum[-1]), rank, world)


def rows_of_gathers(pb, g0: int, g1: int):
    """Global 0-based row range [r0, r1) produced by gathers [g0, g1)."""
    nrc = []
    for k in range(pb.kmax):
        nrc.extend(int(v) for v in pb.nrc1[k, : int(pb.nsrc1[k])])
    cum = np.concatenate([[0], np.cumsum(nrc)]).astype(np.int64)
    return int(cum[g0]), int(cum[g1])


def all_gather_rows(local: dict, group=None):
    """Concatenate per-rank COO blocks (row, col, rw) and dsurf slices in rank order with
    torch.distributed (works with gloo on CPU and nccl on GPU)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    objs = [None] * world
    dist.all_gather_object(objs, local, group=group)
    out = {}
    for key in ("row", "col", "rw"):
        out[key] = np.concatenate([o[key] for o in objs])
    out["dsurf"] = np.concatenate([o["dsurf"] for o in objs])
    out["nar"] = int(sum(o["nar"] for o in objs))
    return out


class NcclComm:
    """ncclComm_t owned by libdsurf_b200.so; the unique id travels through torch.distributed."""

    def __init__(self, rank: int, world: int, device=None):
        import torch
        import torch.distributed as dist

        self.rank, self.world = rank, world
        idbuf = (C.c_char * 128)()
        if rank == 0:
            check(lib().dsurf_nccl_unique_id(idbuf), "nccl_unique_id")
        t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        self.h = C.c_void_p()
        check(lib().dsurf_nccl_comm_init(C.byref(self.h), C.c_char_p(raw), C.c_int(rank), C.c_int(world)),
              "nccl_comm_init")

    def close(self):
        if self.h:
            lib().dsurf_nccl_comm_destroy(self.h)
            self.h = None


def attach(lsmr_system, comm: NcclComm, peer_exchange: bool = True, device=None):
    """Binds a row-partitioned LSMR system to the communicator.  With peer_exchange (default) the per-iteration
    all-reduce runs over peer-mapped exchange buffers (attach_peer_exchange); DSURF_LSMR_NCCL_ONLY=1 keeps NCCL."""
    check(lib().dsurf_lsmr_set_comm(lsmr_system.h, comm.h, C.c_int(comm.rank), C.c_int(comm.world)),
          "lsmr_set_comm")
    if peer_exchange and comm.world > 1:
        attach_peer_exchange(lsmr_system, comm, device)


def attach_peer_exchange(lsmr_system, comm: NcclComm, device=None):
    """Maps every rank's LSMR exchange buffer into every other rank (CUDA IPC over NVLink): the per-iteration
    all-reduce becomes direct peer loads inside a CUDA graph.  Call after attach()."""
    import torch
    import torch.distributed as dist

    hb = (C.c_char * 64)()
    check(lib().dsurf_lsmr_xchg_export(lsmr_system.h, hb), "lsmr_xchg_export")
    mine = torch.frombuffer(bytearray(bytes(hb)), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        mine = mine.cuda(device)
    allh = [torch.empty_like(mine) for _ in range(comm.world)]
    dist.all_gather(allh, mine)
    raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
    check(lib().dsurf_lsmr_xchg_attach(lsmr_system.h, C.c_char_p(raw), C.c_int(comm.rank), C.c_int(comm.world)),
          "lsmr_xchg_attach")


def partition_system(sysd: dict, rank: int, world: int):
    """Row partition of the full system (data rows + smoothing rows) for the distributed LSMR:
    contiguous row blocks balanced by non-zeros; rows are renumbered 1..m_local."""
    rows, cols, vals, b = sysd["rows"], sysd["cols"], sysd["vals"], sysd["cbst"]
    m = sysd["m"]
    counts = np.bincount(rows - 1, minlength=m).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    target = cum[-1] * np.arange(world + 1) / world
    cuts = np.searchsorted(cum, target, side="left")
    cuts[0], cuts[-1] = 0, m
    if m < world:
        raise ValueError(f"partition_system: {m} rows cannot be split over {world} ranks (every rank needs a row)")
    for r in range(1, world):  # strictly increasing cuts: an empty rank would fail alone and hang the others
        cuts[r] = min(max(cuts[r], cuts[r - 1] + 1), m - (world - r))
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    sel = (rows > r0) & (rows <= r1)
    return dict(m=r1 - r0, n=sysd["n"], rows=(rows[sel] - r0).astype(np.int32), cols=cols[sel], vals=vals[sel],
                cbst=b[r0:r1], row0=r0)
