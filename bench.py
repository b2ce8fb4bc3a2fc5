#!/usr/bin/env python
"""bench.py -- DSurfTomo hot path on B200: FMM source-period sweeps/s (+ LSMR iterations/s).

One "step" = the sweep stage of CalSurfG (CalSurfG.f90:1186-1432: B-spline dicing, refined +
coarse eikonal solve, receiver times, ray tracing, Frechet row assembly) over one block of the
workload: all periods x all sources of one data type of BASELINE.json configs[2]
(1025 x 1025 propagation grid, 16 periods x (Rc, Rg, Lc, Lg) x 256 sources, 16 receivers per
gather).  Four consecutive steps are exactly one CalSurfG sweep stage (24 576 sweeps).  The
dispersion stage feeding it is replaced by deterministic synthetic maps/kernels of the same
shape (dsurftomo_b200.inputs.synthetic_dispersion) so that the GPU arm and the CPU reference arm
consume identical inputs; "data": "synthetic".

  python bench.py [--gpus N --steps K --warmup W]          our arm (one process per GPU)
  python bench.py --impl reference ...                      CPU restatement of the reference

See the JSON keys documented in DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dsurftomo_b200 import inputs  # noqa: E402

METRIC = "fmm_source_period_sweeps_per_sec"


def build_problem(cfg: int):
    if cfg == 3:
        return inputs.config(3)
    if cfg == 2:
        return inputs.config(2)
    if cfg == 0:  # tiny smoke configuration (not a benchmark)
        return inputs.synthetic_problem(19, 2, 24, ("Rc", "Rg", "Lc", "Lg"), nrecv=8, name="mini_129sq")
    raise SystemExit("--config must be 2 or 3")


def stage_blocks(pb, mode):
    """Steps: 'stage' = every gather of the CalSurfG call (one step = one full sweep stage);
    'type' = one data-type block per step."""
    if mode == "type":
        return type_blocks(pb)
    return [("all", 0, int(pb.nsrc1.sum()))]


def type_blocks(pb):
    """Gather ranges of the data types present, in the reference's block order."""
    cum = np.concatenate([[0], np.cumsum(pb.nsrc1)]).astype(int)
    ks = [0, pb.kmaxRc, pb.kmaxRc + pb.kmaxRg, pb.kmaxRc + pb.kmaxRg + pb.kmaxLc, pb.kmax]
    names = ["Rc", "Rg", "Lc", "Lg"]
    return [(names[t], int(cum[ks[t]]), int(cum[ks[t + 1]])) for t in range(4) if ks[t + 1] > ks[t]]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_traffic():
    """DRAM bytes measured with ncu (profiles/r01_traffic.json): per sweep of the eikonal kernel at cfg 3 and
    per non-zero per LSMR iteration.  Scaled by the units of one launch in the roofline objects."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else None


def lsmr_bytes(nnz, m, n):
    return 16 * nnz + 8 * (m + 1) + 12 * m + 80 * n


# ------------------------------------------------------------------------------------ CPU arm
def cpu_sweep_sample(pb, pv4, sen12, blocks, per_block, step, nthreads):
    """Oracle sweep stage on `per_block` gathers of every data-type block; returns (sweeps, s)."""
    import oracle_lib as O

    nsw, t = 0, 0.0
    for name, g0, g1 in blocks:
        lo = g0 + (step * per_block) % max(1, (g1 - g0 - per_block + 1))
        hi = min(lo + per_block, g1)
        nrays = 16 * (hi - lo) * 2 + 16
        t0 = time.perf_counter()
        r = O.calsurfg_pre(pb, pv4, sen12, lo, hi, nthreads=nthreads, mode=1, maxnar=nrays * 9000)
        t += time.perf_counter() - t0
        assert r["err"] == 0
        nsw += r["nsweeps"]
    return nsw, t


def run_reference(args, pb, pv4, sen12, blocks):
    """--impl reference: the C++ restatement of the reference (NOT gfortran: no Fortran compiler
    exists in this image) on all host cores, bounded sample per step."""
    import oracle_lib as O

    O.lib()
    cores = os.cpu_count() or 1
    per_block = 128  # gathers of every data type per step: ~10 s of work on 16 host threads, less on more
    nsw_tot, t_tot = 0, 0.0
    for s in range(args.warmup + args.steps):
        nsw, t = cpu_sweep_sample(pb, pv4, sen12, blocks, per_block, s, cores)
        if s >= args.warmup:
            nsw_tot += nsw
            t_tot += t
        if s == 0 and t * (args.warmup + args.steps) > 240:  # keep the whole run within minutes
            per_block = max(1, per_block // 2)
    value = nsw_tot / t_tot
    sample = f"{per_block} gathers of each of {len(blocks)} data types per step, all {cores} host threads"
    out = dict(metric=METRIC, value=value, unit="sweeps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1e3 * t_tot / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f32", data="synthetic", impl="reference",
               config=dict(workload=pb.name, note="C++ restatement of reference (g++ -O3 -fopenmp, strict IEEE), "
                           "not gfortran; bounded sample of the same workload"),
               cpu_baseline=dict(value=value, unit="sweeps/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=value, unit="sweeps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args, pb, pv4, sen12, blocks, tblocks):
    import torch

    from dsurftomo_b200 import api, dist as ddist, hostglue

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    plan = api.Plan(pb)
    for t in range(4):
        if (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)[t] > 0:
            plan.set_dispersion(t, pv4[t], *sen12[3 * t:3 * t + 3])
        elif t in (0, 2):
            plan.set_dispersion(t, pv4[t], None, None, None)
    plan.finalize_dispersion()
    nb = len(blocks)

    def block_of(step):
        return blocks[(step * world + rank) % nb]

    # ---------------- timed region 1: device-resident sweep stage (CUDA events on the launching stream)
    sampler = ClockSampler(local)
    dev_ms, eik_ms, nsw_local, launches, wall, launches_eik, nrays_local = 0.0, 0.0, 0, 0, 0.0, 0, 0
    stage = dict(eikonal_ms=0.0, rays_ms=0.0, assembly_ms=0.0)
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier()
            if rank == 0:
                sampler.start()
            w0 = time.perf_counter()
        _, g0, g1 = block_of(s)
        plan.reset_rows()
        plan.sweeps(g0, g1)
        if s >= args.warmup:
            r0_, r1_ = ddist.rows_of_gathers(pb, g0, g1)
            nrays_local += r1_ - r0_
            tm = plan.timings()
            dev_ms += tm["total_ms"]
            eik_ms += tm["eikonal_ms"]
            nsw_local += tm["sweeps"]
            launches += tm["launches"]
            launches_eik += tm["eikonal_launches"]
            for k in stage:
                stage[k] += tm[k]
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop() if rank == 0 else None
    t_max_ms = maxreduce(dev_ms)
    nsw_total = sumreduce(nsw_local)
    value = nsw_total / (t_max_ms / 1e3)
    Nc = ((pb.nx - 3) * 8 + 1) * ((pb.ny - 3) * 8 + 1)
    b_sweep = 8 * (Nc + 129 * 129)  # veln read + ttn write, coarse + refined
    peak, peak_src = peaks()
    achieved = b_sweep * nsw_local / (eik_ms / 1e3) / 1e9
    n_eik_launches = max(1, int(round(launches_eik)))
    mt = measured_traffic()
    traffic = None
    if mt and pb.nx == 131:  # measured at cfg 3 only
        traffic = mt["eikonal"]["dram_bytes_per_sweep"] * nsw_local / n_eik_launches
    roofline = dict(bound="hbm", kernel="k_eikonal3<16> (+ node-state fill)", achieved=achieved, peak=peak, unit="GB/s",
                    frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                    algorithmic_bytes_per_sweep=b_sweep, algorithmic_bytes_per_launch=b_sweep * nsw_local / n_eik_launches,
                    launches=n_eik_launches, avg_launch_ms=eik_ms / n_eik_launches,
                    traffic_source=(mt["eikonal"]["source"] if traffic else None),
                    note="issue-bound exact-order FMM replay; DRAM traffic ~250x the algorithmic bytes; see DESIGN.md section 4")

    # ---------------- timed region 2: end to end through the host-buffer API (H2D + compute + D2H)
    last = plan.download()  # sizes the pinned output buffers from the last timed step
    cap = int(last["nar"] * (1.05 if args.step_mode == "stage" else 1.6)) + 1024
    pin = dict(row=torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
               col=torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
               rw=torch.empty(cap, dtype=torch.float32, pin_memory=True).numpy(),
               dsurf=torch.empty(max(pb.dall, 1), dtype=torch.float32, pin_memory=True).numpy())
    h2d = pb.vsf.nbytes + sum(a.nbytes for a in pv4) + sum(a.nbytes for a in sen12 if a is not None) + \
        pb.scxf.nbytes * 2 + pb.rcxf.nbytes * 2
    e2e_steps = max(1, min(args.steps, 2))
    e2e_t, e2e_sw, d2h = 0.0, 0, 0
    barrier()
    for s in range(e2e_steps):
        _, g0, g1 = block_of(args.warmup + s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan.set_model(pb.vsf)                                   # H2D: model
        for t in range(4):                                       # H2D: this step's maps + kernels
            if (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)[t] > 0:
                plan.set_dispersion(t, pv4[t], *sen12[3 * t:3 * t + 3])
            elif t in (0, 2):
                plan.set_dispersion(t, pv4[t], None, None, None)
        plan.finalize_dispersion()
        plan.reset_rows()
        plan.sweeps(g0, g1)
        res = plan.download(out=pin)                             # D2H: predicted times + COO rows
        torch.cuda.synchronize()
        e2e_t += time.perf_counter() - t0
        e2e_sw += plan.timings()["sweeps"]
        d2h = 12 * res["nar"] + 4 * pb.dall
    e2e_tmax = maxreduce(e2e_t)
    e2e_value = sumreduce(e2e_sw) / e2e_tmax

    # ---------------- LSMR on the rows of the last block (device-resident, then host-buffer call)
    lsmr = None
    if args.lsmr_iters > 0:
        res = plan.download(out=pin)
        # system = the data rows of one data-type block (the first one in stage mode) + smoothing rows
        lb = tblocks[0] if args.step_mode == "stage" else block_of(args.warmup + e2e_steps - 1)
        r0, r1 = ddist.rows_of_gathers(pb, lb[1], lb[2])
        lo = int(np.searchsorted(res["row"], r0, side="right"))
        hi = int(np.searchsorted(res["row"], r1, side="right"))
        nrows = r1 - r0
        rows = (res["row"][lo:hi] - r0).astype(np.int32)
        res = dict(res, col=res["col"][lo:hi], rw=res["rw"][lo:hi])
        obst = pb.obst[r0:r0 + nrows]
        cb = (obst - res["dsurf"][r0:r0 + nrows]).astype(np.float32)
        srow, scol, sval, cnt3 = hostglue.smoothing_rows(pb.nx, pb.ny, pb.nz, nrows, pb.weight)
        if world > 1:  # smoothing rows are shared out round-robin
            keep = ((srow - nrows - 1) % world) == rank
            srow, scol, sval = srow[keep], scol[keep], sval[keep]
            _, srow = np.unique(srow, return_inverse=True)
            srow = (srow + nrows + 1).astype(np.int32)
            cnt3 = int(srow.max() - nrows) if len(srow) else 0
        R = np.concatenate([rows, srow])
        Cc = np.concatenate([res["col"], scol])
        V = np.concatenate([res["rw"], sval])
        b = np.concatenate([cb, np.zeros(cnt3, np.float32)])
        m, n = nrows + cnt3, pb.maxvp
        t0 = time.perf_counter()
        sysl = api.LsmrSystem(m, n, R, Cc, V, b)
        t_build = time.perf_counter() - t0
        comm = None
        if world > 1:
            comm = ddist.NcclComm(rank, world, local)
            ddist.attach(sysl, comm)
        sysl.solve(pb.damp, itnlim=3, force_iters=True, want_x=False)  # warm-up
        barrier()
        L = sysl.solve(pb.damp, itnlim=args.lsmr_iters, force_iters=True, want_x=False)
        barrier()
        t_it = maxreduce(L["ms_total"]) / 1e3
        nnz_tot, m_tot = sumreduce(float(sysl.nnz)), sumreduce(float(m))
        it_s = L["itn"] / t_it
        ach = lsmr_bytes(sysl.nnz, m, n) * L["itn"] / (L["ms_total"] / 1e3) / 1e9
        ach_spmv = 8.0 * sysl.nnz * L["itn"] / (L["ms_spmv"] / 1e3) / 1e9
        ach_spmtv = 8.0 * sysl.nnz * L["itn"] / (L["ms_spmtv"] / 1e3) / 1e9
        # host-buffer LSMR call (H2D of the COO, CSR/CSC build, iterations, D2H of x)
        t0 = time.perf_counter()
        x = None
        if world == 1:
            iw = hostglue.pack_iw(R, Cc)
            Lh = api.LSMR(m, n, len(iw), len(V), iw, V, b, pb.damp, 1e-6, 1e-6, 100.0, 400, 10)
            t_host = time.perf_counter() - t0
            host = dict(iters=Lh["itn"], istop=Lh["istop"], seconds=t_host, iters_per_s=Lh["itn"] / t_host,
                        h2d_bytes=12 * len(V) + 4 * m, d2h_bytes=4 * n)
        else:
            host = None
        lsmr = dict(iters_per_s=it_s, iters=L["itn"], nnz=int(nnz_tot), m=int(m_tot), n=n, build_s=t_build,
                    roofline=dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                                  spmv_gbs=ach_spmv, spmtv_gbs=ach_spmtv,
                                  traffic=(mt["lsmr"]["dram_bytes_per_nnz_per_iter"] * sysl.nnz if mt else None),
                                  traffic_source=(mt["lsmr"]["source"] if mt else None),
                                  algorithmic_bytes_per_iter=lsmr_bytes(sysl.nnz, m, n),
                                  note="depth-blocked layout stores 1 index + 8 values per vertex: real bytes are "
                                       "~0.58x the algorithmic 16 B per non-zero"),
                    launches_per_iter=(6 if world == 1 else 9), host_buffer_call=host)
        sysl.close()
        if comm:
            comm.close()

    # ---------------- K1: dispersion + depth kernels of the whole model (FP64-bound)
    disp = None
    if args.dispersion and world == 1:
        plan.set_model(pb.vsf)
        plan.dispersion()
        torch.cuda.synchronize()
        dms = plan.timings()["dispersion_ms"]
        ncol = pb.nx * pb.ny
        kts = (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)
        # curve evaluations: every column x (1 + 6 nz) model variants x periods; group-velocity
        # types evaluate two phase curves per period (surfdisp96.f:227-229)
        curves = ncol * (1 + 6 * pb.nz) * sum(k * (2 if t in (1, 3) else 1) for t, k in enumerate(kts))
        disp = dict(ms=dms, columns=ncol, column_types_per_s=ncol * sum(1 for k in kts if k) / (dms / 1e3),
                    period_roots_per_s=curves / (dms / 1e3), bound="fp64 alu/sfu",
                    note="depthkernel for all data types of the model (one CalSurfG dispersion stage), 1 launch set")

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores
    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle_lib as O

        cores = os.cpu_count() or 1
        per_block = 128  # ~10 s of CPU work on 16 host threads (fewer seconds on more cores)
        nsw, t = cpu_sweep_sample(pb, pv4, sen12, tblocks, per_block, 0, cores)
        cpu = dict(value=nsw / t, unit="sweeps/s", cores=cores, kind="port",
                   sample=f"{per_block} gathers of each of {len(tblocks)} data types ({nsw} sweeps, {t:.1f} s), "
                          "C++ restatement of the reference, g++ -O3 -fopenmp strict IEEE (no Fortran compiler here)")
        if lsmr is not None and lsmr["nnz"] <= 3e8:
            iw = hostglue.pack_iw(R, Cc)
            t0 = time.perf_counter()
            Lc = O.lsmr(m, n, iw, V, b, pb.damp, itnlim=3)
            tl = time.perf_counter() - t0
            cpu["lsmr_iters_per_s"] = Lc["itn"] / tl
            cpu["lsmr_sample"] = f"{Lc['itn']} iterations of the same system, 1 thread (the reference's LSMR is serial)"

    if rank == 0:
        out = dict(metric=METRIC, value=value, unit="sweeps/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=t_max_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype="f32", data="synthetic",
                   config=dict(workload=pb.name, grid=f"{(pb.nx - 3) * 8 + 1}x{(pb.ny - 3) * 8 + 1}",
                               step=("one full CalSurfG sweep stage (every period x type x source) per rank per step"
                                     if args.step_mode == "stage" else
                                     "all periods x sources of one data type per rank per step "
                                     f"({[b[0] for b in blocks]} in rotation)"),
                               receivers_per_gather=int(pb.nrc1.max()), parallelism=f"gathers sharded over {world} GPU(s)",
                               l2="per-step working set (node states of thousands of sweeps, GBs) >> 126 MB L2",
                               interpretation="A: quoted grid = FMM propagation grid (SURVEY.md section 8)"),
                   roofline=roofline, cpu_baseline=cpu,
                   e2e=dict(value=e2e_value, unit="sweeps/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                            steps=e2e_steps, api="Plan.set_model/set_dispersion/sweeps/download through the C ABI "
                                                 "with host buffers (pinned outputs)"),
                   gpu_launches=int(launches), clocks=clocks,
                   stage_ms_per_step={k: v / args.steps for k, v in stage.items()}, wall_s_timed=wall,
                   rays=dict(rays_per_s=(nrays_local / (stage["rays_ms"] / 1e3) if stage["rays_ms"] > 0 else None),
                             rows_per_s=(nrays_local / (stage["assembly_ms"] / 1e3) if stage["assembly_ms"] > 0 else None),
                             note="rank 0: receiver times + ray back-trace (k_rays) and Frechet row assembly, rays per second "
                                  "of their own stage time (SURVEY.md 8d: latency-bound gathers, no HBM fraction quoted)"),
                   lsmr=lsmr, dispersion=disp, impl="b200")
        emit(out)
    plan.close()
    if dist is not None:
        dist.destroy_process_group()


_JSON_FD = None


def emit(obj):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr in main()."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)  # C-level writers to fd 1 (e.g. "NCCL version ...") now land on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--step-mode", default="stage", choices=["stage", "type"])
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--lsmr-iters", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dispersion", dest="dispersion", action="store_false")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    pb = build_problem(args.config)
    pv4, sen12 = inputs.synthetic_dispersion(pb)
    blocks = stage_blocks(pb, args.step_mode)
    tblocks = type_blocks(pb)
    if args.impl == "reference":
        run_reference(args, pb, pv4, sen12, tblocks)
    else:
        run_b200(args, pb, pv4, sen12, blocks, tblocks)
    return 0


if __name__ == "__main__":
    sys.exit(main())
