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
    """DRAM bytes measured with ncu (profiles/r02_traffic.json, falling back to r01): per sweep of the eikonal kernels at
    cfg 3 (exact and fast-iterative) and per non-zero per LSMR iteration.  Scaled by the units of one launch in the
    roofline objects."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return None


def lsmr_bytes(nnz, m, n):
    return 16 * nnz + 8 * (m + 1) + 12 * m + 80 * n


def make_config(pb, world, step_mode):
    """The `config` object of the JSON line -- identical keys and values in both arms."""
    return dict(workload=pb.name, grid=f"{(pb.nx - 3) * 8 + 1}x{(pb.ny - 3) * 8 + 1}",
                step=("one full CalSurfG sweep stage: every (period, type, source) gather of the workload, "
                      "sharded over the GPUs in contiguous gather blocks" if step_mode == "stage" else
                      "all periods x sources of one data type per step (A/B runs only)"),
                sweeps_per_step=int(pb.nsrc1[: pb.kmaxRc].sum() + 2 * pb.nsrc1[pb.kmaxRc: pb.kmaxRc + pb.kmaxRg].sum() +
                                    pb.nsrc1[pb.kmaxRc + pb.kmaxRg: pb.kmaxRc + pb.kmaxRg + pb.kmaxLc].sum() +
                                    2 * pb.nsrc1[pb.kmaxRc + pb.kmaxRg + pb.kmaxLc:].sum()),
                receivers_per_gather=int(pb.nrc1.max()),
                l2="per-step working set (node states of thousands of sweeps, tens of GB) >> 126 MB L2",
                interpretation="A: quoted grid = FMM propagation grid (SURVEY.md section 8)")


# ------------------------------------------------------------------------------------ CPU arm
def cpu_sweep_sample(pb, pv4, sen12, blocks, per_block, step, nthreads, mode=1):
    """Oracle sweep stage on `per_block` gathers of every data-type block; returns (sweeps, s).
    mode 1: independent gathers on all host threads; mode 0: the reference's own threading (sweeps serial)."""
    import oracle_lib as O

    nsw, t = 0, 0.0
    for name, g0, g1 in blocks:
        lo = g0 + (step * per_block) % max(1, (g1 - g0 - per_block + 1))
        hi = min(lo + per_block, g1)
        nrays = 16 * (hi - lo) * 2 + 16
        t0 = time.perf_counter()
        r = O.calsurfg_pre(pb, pv4, sen12, lo, hi, nthreads=nthreads, mode=mode, maxnar=nrays * 9000)
        t += time.perf_counter() - t0
        assert r["err"] == 0
        nsw += r["nsweeps"]
    return nsw, t


def cpu_dispersion_sample(pb, cores):
    """Oracle depthkernel (K1) on a strip of the model: 4 model rows x nx columns, Rayleigh phase, the workload's
    periods.  The reference parallelises over model rows (OpenMP, CalSurfG.f90:39-44), so mode (i) here = 4 threads;
    mode (ii) = the same strip cut into single columns over all host threads."""
    import oracle_lib as O

    rows = 4
    vel = np.ascontiguousarray(pb.vsf.reshape(pb.nz, pb.ny, pb.nx)[:, :rows, :])
    t = np.asarray(pb.tRc if pb.kmaxRc else pb.tLc, np.float64)
    out = {}
    for label, nth in (("reference_threading", min(rows, cores)), ("all_cores", cores)):
        # (nz, ny, nx): the oracle, like the reference, runs its OpenMP loop over the ny model rows
        v = vel if label == "reference_threading" else np.ascontiguousarray(vel.reshape(pb.nz, rows * pb.nx, 1))
        t0 = time.perf_counter()
        O.depthkernel(v, 2, 0, t, pb.depz, pb.minthk, nthreads=nth)
        dt = time.perf_counter() - t0
        roots = rows * pb.nx * (1 + 6 * pb.nz) * len(t)
        out[label] = dict(period_roots_per_s=roots / dt, threads=nth, seconds=dt)
    out["sample"] = f"depthkernel (Rayleigh phase, {len(t)} periods) on {rows} model rows x {pb.nx} columns"
    return out


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
    sample = (f"{per_block} gathers of each of {len(blocks)} data types per step (a bounded sample of the step, not a "
              f"whole step), all {cores} host threads over independent gathers (mode ii); C++ restatement of the "
              "reference, g++ -O3 -fopenmp strict IEEE -- not gfortran (no Fortran compiler in this image)")
    out = dict(metric=METRIC, value=value, unit="sweeps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1e3 * t_tot / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f32", data="synthetic", impl="reference",
               config=make_config(pb, args.gpus, args.step_mode),
               cpu_baseline=dict(value=value, unit="sweeps/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=value, unit="sweeps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args, pb, pv4, sen12, blocks, tblocks):
    import torch

    from dsurftomo_b200 import api, dist as ddist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = ddist.NcclComm(rank, world, local)
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

    MODE = args.eikonal

    def make_plan(mode):
        prev = api.set_eikonal_mode(mode)
        plan = api.Plan(pb)
        api.set_eikonal_mode(prev)
        for t in range(4):
            if (pb.kmaxRc, pb.kmaxRg, pb.kmaxLc, pb.kmaxLg)[t] > 0:
                plan.set_dispersion(t, pv4[t], *sen12[3 * t:3 * t + 3])
            elif t in (0, 2):
                plan.set_dispersion(t, pv4[t], None, None, None)
        plan.finalize_dispersion()
        return plan

    plan = make_plan(MODE)
    nb = len(blocks)

    def block_of(step):
        """This rank's contiguous share of the step's gathers (strong scaling: the step is the same for every N)."""
        _, b0, b1 = blocks[step % nb]
        if args.step_mode == "stage":
            return ddist.shard_gathers(pb, rank, world)
        lo, hi = ddist.shard_range(b1 - b0, rank, world)
        return b0 + lo, b0 + hi

    def measure(plan, mode, want_clocks=True, nwarm=None, nsteps=None):
        """timed regions 1 (device-resident) and 2 (host buffers) of the sweep stage for one eikonal pipeline"""
        nwarm = args.warmup if nwarm is None else nwarm
        nsteps = args.steps if nsteps is None else nsteps
        # ---------------- timed region 1: device-resident sweep stage + the NCCL gather of predicted times and COO row
        # blocks to every rank (CUDA events on the launching stream, max over ranks)
        sampler = ClockSampler(local) if want_clocks else None
        dev_ms, eik_ms, gat_ms, nsw_local, launches, wall, launches_eik, nrays_local = 0.0, 0.0, 0.0, 0, 0, 0.0, 0, 0
        stage = dict(eikonal_ms=0.0, rays_ms=0.0, assembly_ms=0.0)
        nar_total, digest = 0, None
        for s in range(nwarm + nsteps):
            if s == nwarm:
                barrier()
                if rank == 0 and sampler is not None:
                    sampler.start()
                w0 = time.perf_counter()
            g0, g1 = block_of(s)
            plan.reset_rows()
            plan.sweeps(g0, g1)
            if comm is not None:
                nar_total = plan.allgather(comm, want_coo=True)
            else:
                nar_total = plan.nar
            if s >= nwarm:
                r0_, r1_ = ddist.rows_of_gathers(pb, g0, g1)
                nrays_local += r1_ - r0_
                tm = plan.timings()
                dev_ms += tm["total_ms"] + (plan.gather_ms if comm is not None else 0.0)
                gat_ms += plan.gather_ms if comm is not None else 0.0
                eik_ms += tm["eikonal_ms"]
                nsw_local += tm["sweeps"]
                launches += tm["launches"]
                launches_eik += tm["eikonal_launches"]
                for k in stage:
                    stage[k] += tm[k]
        barrier()
        wall = time.perf_counter() - w0
        clocks = sampler.stop() if (rank == 0 and sampler is not None) else None
        digest = plan.digest(gathered=comm is not None)  # equal on 1 and on N GPUs <=> identical COO in identical order
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
        traffic_src = None
        if mt and pb.nx == 131 and mode == "exact":  # measured at cfg 3 only
            traffic = mt["eikonal"]["dram_bytes_per_sweep"] * nsw_local / n_eik_launches
            traffic_src = mt["eikonal"]["source"]
        if mt and pb.nx == 131 and mode == "fim" and "eikonal_fim" in mt:
            traffic = mt["eikonal_fim"]["dram_bytes_per_sweep"] * nsw_local / n_eik_launches
            traffic_src = mt["eikonal_fim"]["source"]
        kname = {"exact": "k_eikonal3<16> (+ node-state fill)", "lps": "k_refine<16> + k_march_lps",
                 "fim": "k_fim_march (+ k_refine<16>, k_fim_start, field fill)"}[mode]
        note = {"fim": "block-level fast-iterative sweep: ~50 warp instructions per node (fp32 IEEE sqrt/div chains), issue-bound; "
                       "DRAM traffic ~2x B_sweep (tiles are re-read from L2); see DESIGN.md section 4",
                "lps": "exact-order FMM replay, lane per sweep: latency-bound",
                "exact": "exact-order FMM replay: bound by dependent scattered accesses (issue slots of the serial heap "
                         "work, then DRAM sector rate), not by the algorithmic bytes; DRAM traffic ~250x B_sweep; see "
                         "DESIGN.md section 4"}[mode]
        roofline = dict(bound="hbm", kernel=kname,
                        achieved=achieved, peak=peak, unit="GB/s",
                        frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                        algorithmic_bytes_per_sweep=b_sweep, algorithmic_bytes_per_launch=b_sweep * nsw_local / n_eik_launches,
                        launches=n_eik_launches, avg_launch_ms=eik_ms / n_eik_launches,
                        traffic_source=traffic_src, note=note)

        # ---------------- timed region 2: end to end through the host-buffer API (H2D + compute + gather + D2H on rank 0)
        cap = int(nar_total * (1.05 if args.step_mode == "stage" else 1.6)) + 1024
        pin = None
        if rank == 0:
            pin = dict(row=torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
                       col=torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
                       rw=torch.empty(cap, dtype=torch.float32, pin_memory=True).numpy(),
                       dsurf=torch.empty(max(pb.dall, 1), dtype=torch.float32, pin_memory=True).numpy())
        h2d = pb.vsf.nbytes + sum(a.nbytes for a in pv4) + sum(a.nbytes for a in sen12 if a is not None) + \
            pb.scxf.nbytes * 2 + pb.rcxf.nbytes * 2
        e2e_steps = max(1, min(nsteps, 5))
        e2e_t, e2e_sw, d2h = 0.0, 0, 0
        barrier()
        for s in range(e2e_steps):
            g0, g1 = block_of(nwarm + s)
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
            if comm is not None:
                ntot = plan.allgather(comm, want_coo=True)
                if rank == 0:                                        # D2H: the caller's full (rw, iw, col) + predicted times
                    lib_ = api.lib()
                    import ctypes as C
                    api.check(lib_.dsurf_plan_download_gathered(plan.h, api.ptr(pin["row"], C.c_int), api.ptr(pin["rw"], C.c_float),
                                                                api.ptr(pin["col"], C.c_int)), "download_gathered")
                    api.check(lib_.dsurf_plan_download(plan.h, None, None, None, api.ptr(pin["dsurf"], C.c_float), None),
                              "download_dsurf")
            else:
                ntot = plan.download(out=pin)["nar"]                 # D2H: predicted times + COO rows
            torch.cuda.synchronize()
            barrier()
            e2e_t += time.perf_counter() - t0
            e2e_sw += plan.timings()["sweeps"]
            d2h = 12 * ntot + 4 * pb.dall
        e2e_tmax = maxreduce(e2e_t)
        e2e_value = sumreduce(e2e_sw) / e2e_tmax

        return dict(value=value, t_max_ms=t_max_ms, roofline=roofline, e2e_value=e2e_value, e2e_steps=e2e_steps, h2d=h2d,
                    d2h=d2h, launches=launches, clocks=clocks, stage=stage, gat_ms=gat_ms, nar_total=nar_total, digest=digest,
                    wall=wall, nrays_local=nrays_local, cap=cap)

    M = measure(plan, MODE)
    value, t_max_ms, roofline, e2e_value, e2e_steps, h2d, d2h = (M[k] for k in ("value", "t_max_ms", "roofline", "e2e_value",
                                                                               "e2e_steps", "h2d", "d2h"))
    launches, clocks, stage, gat_ms, nar_total, digest, wall, nrays_local, cap = (M[k] for k in (
        "launches", "clocks", "stage", "gat_ms", "nar_total", "digest", "wall", "nrays_local", "cap"))
    peak, peak_src = peaks()
    mt = measured_traffic()

    # ---------------- LSMR on the FULL system of the step (every data row + every smoothing row, main.f90:418-489),
    # row-partitioned over the ranks exactly as the rows were produced; built on the device from the plans' COO
    lsmr = None
    if args.lsmr_iters > 0 and args.step_mode == "stage":
        g0, g1 = block_of(0)
        # the last end-to-end step left this rank's rows of the stage on the device, unscaled: LSMR is built from them
        t0 = time.perf_counter()
        sysl = api.LsmrSystem.from_plan_shard(plan, rank, world)
        t_build = time.perf_counter() - t0
        m, n = sysl.m, sysl.n
        if comm is not None:
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
        # convergence run with the reference's stopping rules (main.f90:474-485)
        Lc = sysl.solve(pb.damp, itnlim=400, want_x=False)
        t_conv = maxreduce(Lc["ms_total"]) / 1e3
        lsmr = dict(iters_per_s=it_s, iters=L["itn"], nnz=int(nnz_tot), m=int(m_tot), n=n, build_s=t_build,
                    system="full: all data rows of the step + all smoothing rows, row-partitioned over the ranks",
                    per_rank=dict(nnz=int(sysl.nnz), m=int(m)),
                    roofline=dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                                  spmv_gbs=ach_spmv, spmtv_gbs=ach_spmtv,
                                  traffic=(mt["lsmr"]["dram_bytes_per_nnz_per_iter"] * sysl.nnz if mt else None),
                                  frac_on_measured_traffic=(mt["lsmr"]["dram_bytes_per_nnz_per_iter"] * sysl.nnz * L["itn"] /
                                                            (L["ms_total"] / 1e3) / 1e9 / peak if mt else None),
                                  traffic_source=(mt["lsmr"]["source"] if mt else None),
                                  algorithmic_bytes_per_iter=lsmr_bytes(sysl.nnz, m, n),
                                  note="per rank; `frac` is on the ALGORITHMIC bytes of SURVEY.md 8(d) (16 B per non-zero) and can "
                                       "exceed 1: the depth-blocked layout stores 1 index + 8 values per vertex, real DRAM bytes "
                                       "are ~0.58x of that; frac_on_measured_traffic is the fraction of the copy peak really used"),
                    launches_per_iter=(6 if world == 1 else 9),
                    to_convergence=dict(iters=Lc["itn"], istop=Lc["istop"], seconds=t_conv, normr=Lc["normr"]))
        sysl.close()

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
        dp_per_root = 3.4e5  # FP64 operations per period root, from profiles/r01_disp_cfg3.md (pipe-active cycles x lanes)
        disp = dict(ms=dms, columns=ncol, column_types_per_s=ncol * sum(1 for k in kts if k) / (dms / 1e3),
                    period_roots_per_s=curves / (dms / 1e3), bound="fp64 alu/sfu",
                    dp_gflops_est=curves / (dms / 1e3) * dp_per_root / 1e9,
                    dp_gflops_note="period roots/s x 3.4e5 FP64 operations per root (ncu FP64-pipe-active cycles of the "
                                   "Rayleigh-phase run in profiles/r01_disp_cfg3.md; no FMA: --fmad=false)",
                    note="depthkernel for all data types of the model (one CalSurfG dispersion stage), 1 launch set")

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded samples on the host cores
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        per_block = 128  # ~10 s of CPU work on 16 host threads (fewer seconds on more cores)
        nsw, t = cpu_sweep_sample(pb, pv4, sen12, tblocks, per_block, 0, cores, mode=1)
        nsw1, t1 = cpu_sweep_sample(pb, pv4, sen12, tblocks, 2, 0, 1, mode=1)
        cpu = dict(value=nsw / t, unit="sweeps/s", cores=cores, kind="port",
                   sample=f"{per_block} gathers of each of {len(tblocks)} data types ({nsw} sweeps, {t:.1f} s), all host "
                          "threads over independent gathers (mode ii); C++ restatement of the reference, g++ -O3 -fopenmp "
                          "strict IEEE (no Fortran compiler here)",
                   reference_threading=dict(value=nsw1 / t1, unit="sweeps/s", cores=1,
                                            sample=f"{nsw1} sweeps on one thread ({t1:.1f} s): the reference marches its "
                                                   "sources serially (mode i, CalSurfG.f90:1144-1145)"))
        if disp is not None:
            cpu["dispersion"] = cpu_dispersion_sample(pb, cores)

    # ---------------- the other eikonal pipeline on the same steps: the block-level fast-iterative sweep (north_star's
    # design, not bit-exact: profiles/r02_fim_parity.md) next to the exact-order kernel, or the other way round
    other = None
    if args.both and args.step_mode == "stage":
        om = "fim" if MODE != "fim" else "exact"
        try:
            plan.close()
            plan = make_plan(om)
            M2 = measure(plan, om, want_clocks=False, nwarm=min(args.warmup, 3 if om == "fim" else 1),
                         nsteps=max(1, min(args.steps, 3 if om == "fim" else 1)))
            if rank == 0:
                other = dict(pipeline=om, value=M2["value"], unit="sweeps/s", ms_per_step=M2["t_max_ms"] / max(1, min(args.steps, 3 if om == "fim" else 1)),
                             e2e=dict(value=M2["e2e_value"], unit="sweeps/s", h2d_bytes_per_step=int(M2["h2d"]),
                                      d2h_bytes_per_step=int(M2["d2h"]), steps=M2["e2e_steps"]),
                             roofline=M2["roofline"], gpu_launches=int(M2["launches"]),
                             stage_ms_per_step={k: v / max(1, min(args.steps, 3 if om == "fim" else 1)) for k, v in M2["stage"].items()},
                             coo=dict(nar=int(M2["nar_total"]), digest=f"{M2['digest'][0]:016x}"),
                             parity=("travel times bit-identical to the reference" if om == "exact" else
                                     "iterates the reference's own update rule: <= 2.2e-6 relative on travel times at 1025^2 (1.5 % of "
                                     "the nodes, last bits); 0.7 % of the rays differ in their G entries at 1025^2 (0.1 % through another "
                                     "B-spline cell, the rest through the ftol thresholds), none at <= 257^2; profiles/r02_fim_parity.md"))
        except Exception as e:  # the headline numbers are complete: report the failure instead of losing the line
            if rank == 0:
                other = dict(pipeline=om, error=f"{type(e).__name__}: {e}")
    if rank == 0:
        cfg = make_config(pb, world, args.step_mode)
        out = dict(metric=METRIC, value=value, unit="sweeps/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=t_max_ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                   dtype="f32", data="synthetic", config=cfg,
                   parallelism=(f"gathers sharded over {world} GPU(s) in contiguous blocks; per step one NCCL all-gather of "
                                "the predicted times and of the COO row blocks (counts, then grouped broadcasts) inside the "
                                "timed region" if world > 1 else "1 GPU: no exchange"),
                   roofline=roofline, cpu_baseline=cpu,
                   e2e=dict(value=e2e_value, unit="sweeps/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                            steps=e2e_steps, api="Plan.set_model/set_dispersion/sweeps/[allgather]/download through the "
                                                 "C ABI with host buffers (pinned outputs on rank 0)"),
                   gpu_launches=int(launches), clocks=clocks,
                   stage_ms_per_step=dict({k: v / args.steps for k, v in stage.items()}, gather_ms=gat_ms / args.steps),
                   coo=dict(nar=int(nar_total), digest=f"{digest[0]:016x}", digest_n=digest[1],
                            note="order-sensitive digest of (row, col, rw): equal for every --gpus N"),
                   wall_s_timed=wall,
                   rays=dict(rays_per_s=(nrays_local / (stage["rays_ms"] / 1e3) if stage["rays_ms"] > 0 else None),
                             rows_per_s=(nrays_local / (stage["assembly_ms"] / 1e3) if stage["assembly_ms"] > 0 else None),
                             note="rank 0: receiver times + ray back-trace (k_rays) and Frechet row assembly, rays per second "
                                  "of their own stage time (SURVEY.md 8d: latency-bound gathers, no HBM fraction quoted)"),
                   lsmr=lsmr, dispersion=disp, impl="b200", eikonal_pipeline=MODE,
                   eikonal_pipeline_note=("exact: the reference's heap pop order replayed, travel times bit-identical (library "
                                          "default)" if MODE == "exact" else "fim: block-level fast-iterative sweep, not bit-exact, "
                                          "see other_pipeline.parity" if MODE == "fim" else MODE),
                   other_pipeline=other)
    plan.close()
    # ---------------- e2e through dsurf_calsurfg itself (K1 included), host buffers in and out: the call a Fortran
    # main program makes (main.f90:355-359).  N = 1 only; its own plan needs the memory the bench plan just released.
    if rank == 0 and world == 1 and args.calsurfg_e2e and args.step_mode == "stage":
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prev_mode = api.set_eikonal_mode(MODE)
        try:
            res = api.CalSurfG(pb, maxnar=cap)
        finally:
            api.set_eikonal_mode(prev_mode)
        dt = time.perf_counter() - t0
        out["e2e_calsurfg"] = dict(value=out["config"]["sweeps_per_step"] / dt, unit="sweeps/s", seconds=dt, steps=1,
                                   nar=int(res["nar"]),
                                   note="one dsurf_calsurfg call with host buffers: H2D model, dispersion + depth kernels "
                                        "(K1) of the real model, sweeps, rays, rows, D2H of (rw, iw, col, dsurf)")
    if rank == 0:
        emit(out)
    if comm is not None:
        comm.close()
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
    ap.add_argument("--no-calsurfg-e2e", dest="calsurfg_e2e", action="store_false")
    ap.add_argument("--eikonal", default=os.environ.get("DSURF_EIKONAL", "exact"), choices=["exact", "lps", "fim"],
                    help="eikonal pipeline of the headline numbers (library default: exact)")
    ap.add_argument("--no-both", dest="both", action="store_false",
                    help="skip the second pass with the other eikonal pipeline (other_pipeline key)")
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
