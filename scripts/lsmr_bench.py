#!/usr/bin/env python
"""Stand-alone LSMR iteration-rate probe (no eikonal stage): a synthetic system with the shape of one
data-type block of cfg 3 (65 536 ray rows x ~220 B-spline vertices x 8 depths + the smoothing rows of
main.f90:413-455; n = 133 128, nnz ~ 1.2e8).  Prints iterations/s and B_lsmr GB/s for the fused
(default) and unfused (DSURF_LSMR_NO_FUSE=1) small-vector paths -- run each in its own process:

    python scripts/lsmr_bench.py [--rows 65536] [--iters 60]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsurftomo_b200 import api, hostglue  # noqa: E402


def build(nrows, nx=131, ny=131, nz=9, nvert=130, seed=3):
    rng = np.random.default_rng(seed)
    nvx, nvz, K = nx - 2, ny - 2, nz - 1
    P = nvx * nvz
    # each ray: a straight-ish walk over the vertex grid
    x0 = rng.integers(0, nvx, nrows)
    z0 = rng.integers(0, nvz, nrows)
    ang = rng.uniform(0, 2 * np.pi, nrows)
    t = np.arange(nvert)[None, :] * 0.9
    vx = np.clip((x0[:, None] + t * np.cos(ang)[:, None]).astype(np.int64), 0, nvx - 2)
    vz = np.clip((z0[:, None] + t * np.sin(ang)[:, None]).astype(np.int64), 0, nvz - 2)
    vert = vz * nvx + vx                                  # [nrows, nvert]
    vert = np.concatenate([vert, vert + nvx + 1], axis=1)  # two-vertex-wide band like the B-spline footprint
    vert = np.sort(vert, axis=1)
    keep = np.ones_like(vert, bool)
    keep[:, 1:] = vert[:, 1:] != vert[:, :-1]
    rows, cols, vals = [], [], []
    ridx = np.repeat(np.arange(nrows), vert.shape[1]).reshape(nrows, vert.shape[1])
    for k in range(K):
        rows.append(ridx[keep])
        cols.append((k * P + vert)[keep])
    rows = np.concatenate(rows).astype(np.int32) + 1
    cols = np.concatenate(cols).astype(np.int32) + 1
    vals = rng.normal(0, 1e-2, len(rows)).astype(np.float32)
    srow, scol, sval, cnt3 = hostglue.smoothing_rows(nx, ny, nz, nrows, 4.0)
    R = np.concatenate([rows, srow]).astype(np.int32)
    Cc = np.concatenate([cols, scol]).astype(np.int32)
    V = np.concatenate([vals, sval]).astype(np.float32)
    b = np.concatenate([rng.normal(0, 1, nrows), np.zeros(cnt3)]).astype(np.float32)
    return nrows + cnt3, P * K, R, Cc, V, b, (nx, ny, nz)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=60)
    a = ap.parse_args()
    m, n, R, Cc, V, b, geom = build(a.rows)
    api.lsmr_hint_geometry(*geom)
    sysl = api.LsmrSystem(m, n, R, Cc, V, b)
    sysl.solve(1.0, itnlim=5, force_iters=True, want_x=False)
    best = None
    for _ in range(3):
        L = sysl.solve(1.0, itnlim=a.iters, force_iters=True, want_x=True)
        if best is None or L["ms_total"] < best["ms_total"]:
            best = L
    nnz = len(V)
    bytes_it = 16 * nnz + 8 * (m + 1) + 12 * m + 80 * n
    it_s = best["itn"] / (best["ms_total"] / 1e3)
    from dsurftomo_b200._lib import lib
    out = dict(fused=os.environ.get("DSURF_LSMR_NO_FUSE") is None, fused_cluster=int(lib().dsurf_lsmr_fused_cluster(sysl.h)), fork=os.environ.get("DSURF_LSMR_NO_FORK") is None,
               m=m, n=n, nnz=nnz, iters=best["itn"], iters_per_s=it_s, us_per_iter=1e6 / it_s,
               b_lsmr_gbs=bytes_it * it_s / 1e9, spmv_us=best["ms_spmv"] * 1e3 / best["itn"],
               spmtv_us=best["ms_spmtv"] * 1e3 / best["itn"], normx=best["normx"], normr=best["normr"],
               x_checksum=float(np.abs(best["x"]).sum()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
