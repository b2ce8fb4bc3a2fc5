"""Host glue between CalSurfG and LSMR, mirroring the reference's main program
(src/main.f90:361-466 and :518-532) in REAL*4 numpy arithmetic.  It stays on the host exactly as
in the reference (north_star: "host code stays ... unchanged"); a device-resident version is the
first "next" row of SURVEY.md section 8(f).
"""
from __future__ import annotations

import numpy as np

F32, I32 = np.float32, np.int32


def getpercentile(a):
    """getpercentile.f90: RA(int(0.25*N)), RA(int(0.75*N)) of the ascending sort (1-based)."""
    n = len(a)
    ra = np.sort(np.asarray(a, F32), kind="stable")
    # N < 4 makes the reference read RA(0) (out of bounds): clamp to the first element like the device glue
    # (glue.cu: glue_percentiles) instead of wrapping to the largest one
    i25 = max(int(F32(0.25) * F32(n)), 1)
    i75 = max(int(F32(0.75) * F32(n)), 1)
    return ra[i25 - 1], ra[i75 - 1]


def smoothing_rows(nx, ny, nz, dall, weight):
    """main.f90:418-459: first-order Laplacian rows; returns (rows1, cols1, vals, count3)."""
    nvx, nvz = nx - 2, ny - 2
    k, j, i = np.meshgrid(np.arange(1, nz), np.arange(1, nvz + 1), np.arange(1, nvx + 1), indexing="ij")
    k, j, i = k.ravel(), j.ravel(), i.ravel()
    c0 = (k - 1) * nvz * nvx + (j - 1) * nvx + i
    edge = (i == 1) | (i == nvx) | (j == 1) | (j == nvz) | (k == 1) | (k == nz - 1)
    count3 = np.arange(1, len(c0) + 1)
    nper = np.where(edge, 1, 7)
    start = np.concatenate([[0], np.cumsum(nper)[:-1]])
    tot = int(nper.sum())
    rows = np.repeat(dall + count3, nper).astype(I32)
    cols = np.zeros(tot, I32)
    vals = np.zeros(tot, F32)
    w = F32(weight)
    e = np.nonzero(edge)[0]
    cols[start[e]] = c0[e]
    vals[start[e]] = F32(2.0) * w
    q = np.nonzero(~edge)[0]
    offs = [0, -1, +1, -nvx, +nvx, -nvz * nvx, +nvz * nvx]
    for t, o in enumerate(offs):
        cols[start[q] + t] = c0[q] + o
        vals[start[q] + t] = (F32(6.0) if t == 0 else F32(-1.0)) * w
    return rows, cols, vals, len(c0)


def host_glue(pb, dsyn, row, col, rw):
    """Residuals, outlier weights, row scaling, smoothing rows (main.f90:361-466).

    row/col are 1-based as CalSurfG returns them.  Returns dict(m, n, rows, cols, vals, cbst,
    datweight) with the smoothing rows appended after the data rows."""
    dall = pb.dall
    cb = (np.asarray(pb.obst, F32) - np.asarray(dsyn, F32)).astype(F32)
    q25, q75 = getpercentile(cb)
    thr = F32(pb.threshold)
    out = (cb < q25 * thr) | (cb > q75 * thr)
    datw = np.where(out, F32(0), F32(1)).astype(F32)
    cb = np.where(out, F32(0), cb).astype(F32)
    vals = (np.asarray(rw, F32) * datw[np.asarray(row) - 1]).astype(F32)
    srow, scol, sval, count3 = smoothing_rows(pb.nx, pb.ny, pb.nz, dall, pb.weight)
    rows = np.concatenate([np.asarray(row, I32), srow])
    cols = np.concatenate([np.asarray(col, I32), scol])
    vals = np.concatenate([vals, sval])
    cbst = np.concatenate([cb, np.zeros(count3, F32)])
    return dict(m=dall + count3, n=pb.maxvp, rows=rows, cols=cols, vals=vals, cbst=cbst, datweight=datw)


def pack_iw(rows, cols):
    """iw = [nar | rows | cols] (main.f90:463-466)."""
    nar = len(rows)
    iw = np.zeros(2 * nar + 1, I32)
    iw[0] = nar
    iw[1:nar + 1] = rows
    iw[nar + 1:] = cols
    return iw


def model_update(pb, vsf, dv):
    """main.f90:518-532: clip dv to +-0.5, add to the interior nodes, clamp to [Minvel, Maxvel]."""
    nx, ny, nz = pb.nx, pb.ny, pb.nz
    dv = np.clip(np.asarray(dv, F32), F32(-0.5), F32(0.5)).reshape(nz - 1, ny - 2, nx - 2)
    out = np.array(vsf, F32, copy=True)
    upd = (out[: nz - 1, 1:ny - 1, 1:nx - 1] + dv).astype(F32)
    out[: nz - 1, 1:ny - 1, 1:nx - 1] = np.clip(upd, F32(pb.minvel), F32(pb.maxvel))
    return out, dv.ravel()
